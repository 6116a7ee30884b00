"""Device-resident plane+parallax passes: the arithmetic of ``pipeline/compute_structure``,
``pipeline/build_cost_volume`` and ``pipeline/structure2flow`` of the reference, composed from the
C-ABI kernels with every dense array kept in HBM.  The reference-signature functions in
``mrflow_b200.pipeline.*`` are thin host wrappers (upload -> these passes -> download).

Row bands: every pass works on a band ``[y0, y0+rows)`` of a frame ``frame_h`` rows tall and takes
a ``stats`` object that turns per-band partial statistics (sums, counts, order statistics, min/max)
into whole-frame ones -- ``LocalStats`` for one GPU, ``mrflow_b200.sharding.BandStats`` for row
bands over several GPUs.

All citations are paths inside the reference repository.
"""
import numpy as np
import torch

from . import _lib, device

N_LABELS = 51                         # pipeline/build_cost_volume.py:50
CHANNEL_WEIGHTS = (0.01, 1.0, 1.0)    # structure_refinement/optimize_structure_single_scale.py:350,371

# parameters.py:4-55 -- defaults of the fields read on the path
PARAM_DEFAULTS = dict(debug_use_rigidity_gt=0, debug_save_frames=0, debug_compute_figure=0,
                      rigidity_sigma_direction=1.0, rigidity_weight_cnn=0.25, rigidity_sigma_structure=1.0,
                      rigidity_use_structure=1, scale_structure=1, nonlinear_structure_initialization=1,
                      occlusion_reasoning=1, tempdir=".")


def param(params, name):
    return getattr(params, name, PARAM_DEFAULTS[name])


class LocalStats:
    """Whole-frame statistics when one GPU holds the whole frame."""
    world = 1

    def total(self, value):
        """sum of a per-band host scalar over the bands"""
        return value

    def device_sum(self, t):
        """sum over bands of a small device tensor (returns a device tensor)"""
        return t

    def median(self, keys, n):
        return device.median(keys, n)

    def minmax(self, mm):
        """mm: fp64[2] device tensor (min, max) of this band -> host (min, max) over all bands"""
        a = mm.cpu().numpy()
        return float(a[0]), float(a[1])

    def window(self, win):
        """sum over bands of the (disjoint) epipole windows"""
        return win


def _as_u8(t):
    return t if t.dtype == torch.uint8 else t.to(torch.uint8)


class Band:
    """The inputs of one row band of a triplet, resident on one GPU.

    images: 3 x fp64 [rows(+halo), w, 3] (index 0 = t-1, 1 = t, 2 = t+1; the neighbours may carry a
    halo: ``nb_y0`` is the global row of their first row); flow: [[u,v] bwd, [u,v] fwd] fp32
    [rows,w]; rigidity fp64 [rows,w]; occlusions 2 x uint8 [rows,w] or None.
    """

    def __init__(self, images, flow, rigidity, occlusions=None, y0=0, frame_h=None, nb_y0=None, nb_rows=None,
                 img_y0=None):
        self.images = images
        self.flow = flow
        self.rigidity = rigidity
        self.rows, self.w = rigidity.shape
        self.y0 = int(y0)
        self.frame_h = self.rows if frame_h is None else int(frame_h)
        self.nb_y0 = self.y0 if nb_y0 is None else int(nb_y0)          # first neighbour-texel row kept
        self.nb_rows = self.rows if nb_rows is None else int(nb_rows)
        self.img_y0 = self.nb_y0 if img_y0 is None else int(img_y0)    # first uploaded image row
        dev = rigidity.device
        if occlusions is None:
            z = torch.zeros((self.rows, self.w), dtype=torch.uint8, device=dev)
            occlusions = [z, z]
        self.occ = [_as_u8(o) for o in occlusions]
        self.device = dev

    @classmethod
    def from_host(cls, images, flow, rigidity, occlusions=None, device_name="cuda", rows=None, halo=0):
        """Upload a whole frame, or rows ``rows=(y0, y1)`` of it with ``halo`` extra neighbour rows
        (``images=None``: a stage that does not sample the frames, e.g. compute_structure, skips their 32 MB).
        Images are uploaded with 2 more rows on each cut side so that the gradient + 3x3 blur of the
        channel construction is exact on every row that is kept."""
        h, w = np.asarray(rigidity).shape
        y0, y1 = (0, h) if rows is None else rows
        n0, n1 = max(0, y0 - halo), min(h, y1 + halo)
        m0, m1 = max(0, n0 - 2), min(h, n1 + 2)

        def up(a, dtype, lo, hi):
            a = np.ascontiguousarray(np.asarray(a)[lo:hi], dtype=dtype)
            return torch.from_numpy(a).to(device_name, non_blocking=True)

        imgs = [up(images[i], np.float64, m0, m1) for i in range(3)] if images is not None else None
        fl = [[up(flow[f][0], np.float32, y0, y1), up(flow[f][1], np.float32, y0, y1)] for f in range(2)]
        rg = up(rigidity, np.float64, y0, y1)
        oc = None
        if occlusions is not None:
            oc = [up(np.asarray(o) == 1, np.uint8, y0, y1) for o in occlusions]
        return cls(imgs, fl, rg, oc, y0=y0, frame_h=h, nb_y0=n0, nb_rows=n1 - n0, img_y0=m0)


# ----------------------------------------------------------------------------------------------
def epipole_window(u_res, v_res, rigid_mask, q, y0=0, frame_h=None):
    """The 21x21 window of residual flow and rigidity round the epipole that ``correct_epipole`` looks at
    (pipeline/compute_structure.py:463-487), fetched from the device: fp64 [3,21,21] = (u_res, v_res,
    rigid).  ``None`` when the window leaves the image (:476-477: the epipole stays).  Rows of the window
    outside this band ``[y0, y0+rows)`` are zero, so the windows of all row bands add up to the whole one."""
    q = np.asarray(q, dtype=np.float64)
    rows, w = u_res.shape
    h = rows if frame_h is None else frame_h
    xmin, xmax = int(q[0]) - 10, int(q[0]) + 10
    ymin, ymax = int(q[1]) - 10, int(q[1]) + 10
    if xmin < 0 or xmax >= w or ymin < 0 or ymax >= h:
        return None
    win = np.zeros((3, 21, 21))
    lo, hi = max(ymin, y0), min(ymax + 1, y0 + rows)
    if hi > lo:
        sl = (slice(lo - y0, hi - y0), slice(xmin, xmax + 1))
        dst = slice(lo - ymin, hi - ymin)
        win[0, dst] = u_res[sl].cpu().numpy().astype(np.float64)
        win[1, dst] = v_res[sl].cpu().numpy().astype(np.float64)
        win[2, dst] = rigid_mask[sl].cpu().numpy() > 0
    return win


def epipole_from_window(win, q):
    """pipeline/compute_structure.py:487-511 on the window ``epipole_window`` returns.  The reference's own
    objective sum(|u*(qx-x) + v*(qy-y)| / max(1e-3, |q-x|^2)) over the rigid pixels of the window is
    minimised with scipy's BFGS exactly as the reference does (:507) -- the gradient chumpy derives by
    autodiff (:495-504) written out.  Returns ``q`` unchanged when more than half of the window is non-rigid
    (:489-491) or when the optimum moved 10 px or more (:509)."""
    from scipy import optimize
    q = np.asarray(q, dtype=np.float64)
    xmin, ymin = int(q[0]) - 10, int(q[1]) - 10
    r = win[2] > 0
    if (~r).sum() > (r.size / 2):
        return q
    yy, xx = np.mgrid[ymin:ymin + 21, xmin:xmin + 21]
    uw, vw, xw, yw = win[0][r], win[1][r], xx[r].astype(np.float64), yy[r].astype(np.float64)

    def fun(qe):
        fx, fy = qe[0] - xw, qe[1] - yw
        d2 = fx ** 2 + fy ** 2
        den = np.maximum(1e-3, d2)
        par = uw * fx + vw * fy
        live = (d2 > 1e-3).astype(np.float64)
        sgn = np.sign(par)
        gx = sgn * uw / den - np.abs(par) / den ** 2 * (2 * fx * live)
        gy = sgn * vw / den - np.abs(par) / den ** 2 * (2 * fy * live)
        return (np.abs(par) / den).sum(), np.array([gx.sum(), gy.sum()])

    res = optimize.minimize(fun, x0=q, jac=True, method="BFGS", options={"disp": False})
    return res.x if np.linalg.norm(res.x - q) < 10 else q


def correct_epipole(u_res, v_res, q, rigid_mask, y0=0, frame_h=None, stats=None):
    """pipeline/compute_structure.py:453-511.  ``u_res``/``v_res`` are device tensors; only the 21x21
    window round the epipole (441 pixels) comes to the host (``epipole_window`` / ``epipole_from_window``).
    With row bands, ``stats.window`` adds the bands' windows up (disjoint rows, so the sum is exact) and every
    rank solves the same 441-pixel problem -- also when the window straddles a band boundary."""
    win = epipole_window(u_res, v_res, rigid_mask, q, y0, frame_h)
    if win is None:
        return np.asarray(q, dtype=np.float64)
    if stats is not None:
        win = stats.window(win)
    elif y0 != 0 or (frame_h is not None and frame_h != u_res.shape[0]):
        raise ValueError("a row band needs stats= to assemble the epipole window")
    return epipole_from_window(win, q)


def first_pass(band, homographies, epipoles, rigid_mask, stats, fix_epipole=True):
    """pipeline/compute_structure.py:67-91 for both neighbour frames: residual flow, refined epipole,
    parallax, mu.  Returns dict(residual, parallax, q, mu)."""
    n_px = band.frame_h * band.w
    res, par, qs, mus = [], [], [], []
    for f in range(2):
        Hinv = np.linalg.inv(np.asarray(homographies[f], dtype=np.float64))   # utils/flow_homography.py:12
        q = np.asarray(epipoles[f], dtype=np.float64)
        ur, vr, p, _ = device.flow_to_parallax(band.flow[f][0], band.flow[f][1], Hinv, q, y0=band.y0, want_dists=False)
        if fix_epipole:
            q_new = correct_epipole(ur, vr, q, rigid_mask, band.y0, band.frame_h, stats)
            if not np.array_equal(q_new, q):
                _, _, p, _ = device.flow_to_parallax(band.flow[f][0], band.flow[f][1], Hinv, q_new, y0=band.y0,
                                                     want_residual=False, want_dists=False)
                q = q_new
        # mu = foe_dists.mean() (:84) depends only on (q, h, w): every band evaluates the whole-frame sum
        # itself (same grid and summation tree everywhere), so no exchange and the bits of a one-GPU run
        dsum = device.epipole_dist_sum(band.frame_h, band.w, q, y0=0, device=band.device)
        mu = float(dsum.cpu()[0]) / n_px
        res.append((ur, vr)); par.append(p); qs.append(q); mus.append(mu)
    return dict(residual=res, parallax=par, q=qs, mu=mus)


def scales(band, fp, p_thr, stats, scale_structure=1, B_fwd_fixed=None):
    """pipeline/compute_structure.py:172-221 (or pipeline/build_cost_volume.py:116-131 with
    ``B_fwd_fixed=1.0``): valid mask, B_fwd (MAD scale) and B_b (median ratio)."""
    par, q, mu = fp["parallax"], fp["q"], fp["mu"]
    pv, cand, cnt = device.structure_candidates(par[0], par[1], p_thr, q[0], q[1], mu[0], mu[1], 1.0,
                                                1 if B_fwd_fixed is None else 0, y0=band.y0)
    n_local = None
    if B_fwd_fixed is not None:
        B_fwd = float(B_fwd_fixed)
        few = False
    else:
        n_local = int(cnt.cpu()[0])
        n_all = int(stats.total(n_local))
        few = n_all < 100                                                     # :183
        if few or scale_structure not in (1, 2):
            B_fwd = 1.0                                                       # :184,:198
        elif scale_structure == 1:
            med = stats.median(cand, n_local)
            B_fwd = 1.426 * stats.median(device.abs_deviation(cand, med, n_local), n_local)   # :190
        else:
            if stats.world != 1:
                raise NotImplementedError("scale_structure=2 (std) is not sharded")
            B_fwd = float(cand[:n_local].std(unbiased=False).cpu())           # :194
    _, cand2, cnt2 = device.structure_candidates(par[0], par[1], p_thr, q[0], q[1], mu[0], mu[1], B_fwd, 2,
                                                 y0=band.y0, pv_out=False)
    n2 = int(cnt2.cpu()[0])
    if B_fwd_fixed is None and few:
        B_b = -B_fwd                                                          # :206
    else:
        B_b = stats.median(cand2, n2)                                         # :221
    return pv, [B_b, B_fwd]


def structures(band, fp, B_ar):
    """pipeline/compute_structure.py:240-246."""
    return [device.parallax_to_structure(fp["parallax"][f], fp["q"][f], B_ar[f], fp["mu"][f], y0=band.y0)
            for f in range(2)]


def label_range(band, S, zero_mask, q_ar, stats, n_labels=N_LABELS):
    """pipeline/build_cost_volume.py:154-169,190-192: linspace(1.5*min, 1.5*max, 51) over both normed
    structures with the zeroing rules applied.  Returns a host fp64 array."""
    vals = []
    h, w = band.frame_h, band.w
    for f in range(2):
        q = q_ar[f]
        bx0 = max(0, min(w - 1, int(q[0] - 10))); bx1 = max(0, min(w - 1, int(q[0] + 10)))
        by0 = max(0, min(h - 1, int(q[1] - 10))); by1 = max(0, min(h - 1, int(q[1] + 10)))
        box = None if (bx0 == w - 1 or bx1 == 0 or by0 == h - 1 or by1 == 0) else (bx0, bx1, by0, by1)   # :162-165
        vals.append(stats.minmax(device.masked_minmax(S[f], zero_mask, box, y0=band.y0)))
    vals = np.array(vals)                                    # NaN propagates like ndarray.min()
    return np.linspace(1.5 * vals[:, 0].min(), 1.5 * vals[:, 1].max(), n_labels)


def channels(band):
    """structure_refinement/optimize_structure_single_scale.py:350-371 for the three frames: fp64
    channels of the reference frame (band rows only) and packed fp32 texels of both neighbours (band
    rows + halo)."""
    def cut(t, first, n):
        off = first - band.img_y0
        return t if (off == 0 and t.shape[0] == n) else t[off:off + n].contiguous()

    if band.img_y0 != 0 or band.images[1].shape[0] != band.frame_h:
        # the uploaded rows must either start at the frame border or carry the 2-row stencil margin
        assert band.img_y0 == 0 or band.img_y0 <= min(band.y0, band.nb_y0) - 2
    ref64, _ = device.prepare_channels(band.images[1], want_texels=False)
    tex = [cut(device.prepare_channels(band.images[i], want_ch64=False)[1], band.nb_y0, band.nb_rows) for i in (0, 2)]
    return cut(ref64, band.y0, band.rows), tex


def cost_volumes(band, ref64, tex, labels_dev, homographies, q_ar, mu_ar, B_ar, nonrigid, occluded=(None, None),
                 rho="lorentzian", rho_param=1.0, interp="cubic", layout="LHW_F32", variant="staged",
                 out=(None, None), want_cost=True, i32_scale=1000.0, halo_miss=None, frames=(0, 1), events=None,
                 out_ptr=(None, None), plane_stride=0, state=None, n_labels=N_LABELS, pair=True):
    """The per-label data cost of both neighbour frames (see device.cost_volume).  The backward
    frame is evaluated at ``label * mu0/mu1`` with the index-0 parameters
    (optimize_structure_single_scale.py:435-442).  Both frames go out as ONE launch (``device.cost_volume_pair``)
    unless ``pair=False`` or a single frame is asked for; ``events`` then receives one (start, end) pair that
    spans both frames."""
    res = [None, None]
    if state is not None:                      # whole-frame scalars stay on the device
        mu_ar, B_ar = (0.0, 1.0), (0.0, 0.0)
    scales = ((mu_ar[0] / mu_ar[1]), 1.0)
    if rho == "lorentzian_auto":
        # the reference's auto-sigma Lorentzian: sigma is a median over each whole label plane
        # (utils/robust_functions.py:55-58) -- a second mode with three launches per label, whole frames only
        if band.rows != band.frame_h or layout != "LHW_F32" or out_ptr[0] is not None or out_ptr[1] is not None:
            raise NotImplementedError("the auto-sigma cost volume needs a whole frame on one GPU and the LHW_F32 layout")
        for f in frames:
            c, mn, am = device.cost_volume_auto(ref64, tex[f], labels_dev, homographies[f], q_ar[f], mu_ar[f], B_ar[f],
                                                label_scale=scales[f], nonrigid=nonrigid, occluded=occluded[f], cw=CHANNEL_WEIGHTS,
                                                interp=interp, want_cost=want_cost, halo_miss=halo_miss, state=state, frame=f,
                                                n_labels=n_labels)
            if out[f] is not None and c is not None:
                out[f].copy_(c); c = out[f]
            res[f] = (c, mn, am)
        return res
    if pair and tuple(frames) == (0, 1) and tex[0].shape == tex[1].shape:
        if events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        res = device.cost_volume_pair(ref64, tex, labels_dev, homographies, q_ar, mu_ar, B_ar, label_scale=scales,
                                      nonrigid=nonrigid, occluded=occluded, rho=rho, rho_param=rho_param, cw=CHANNEL_WEIGHTS,
                                      interp=interp, layout=layout, i32_scale=i32_scale, variant=variant, y0=band.y0,
                                      frame_h=band.frame_h, nb_y0=band.nb_y0, out=out, want_cost=want_cost, halo_miss=halo_miss,
                                      out_ptr=out_ptr, plane_stride=plane_stride, state=state, n_labels=n_labels)
        if events is not None:
            ev[1].record()
            events.append(ev)
        return res
    for f in frames:
        if events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        res[f] = device.cost_volume(ref64, tex[f], labels_dev, homographies[f], q_ar[f], mu_ar[f], B_ar[f],
                                    label_scale=scales[f], nonrigid=nonrigid, occluded=occluded[f], rho=rho,
                                    rho_param=rho_param, cw=CHANNEL_WEIGHTS, interp=interp, layout=layout,
                                    i32_scale=i32_scale, variant=variant, y0=band.y0, frame_h=band.frame_h,
                                    nb_y0=band.nb_y0, out=out[f], want_cost=want_cost, halo_miss=halo_miss,
                                    out_ptr=out_ptr[f], plane_stride=plane_stride, state=state, frame=f,
                                    n_labels=n_labels)
        if events is not None:
            ev[1].record()
            events.append(ev)
    return res


def refine_epipoles(band, homographies, epipoles, rigid_mask, stats=None):
    """compute_structure.py:77 for both frames.  Touches the device only when the 21x21 window round an
    epipole lies inside the image (otherwise correct_epipole returns q unchanged, :476-477)."""
    out = []
    for f in range(2):
        q = np.asarray(epipoles[f], dtype=np.float64)
        xmin, xmax, ymin, ymax = int(q[0]) - 10, int(q[0]) + 10, int(q[1]) - 10, int(q[1]) + 10
        if xmin < 0 or xmax >= band.w or ymin < 0 or ymax >= band.frame_h:
            out.append(q)
            continue
        Hinv = np.linalg.inv(np.asarray(homographies[f], dtype=np.float64))
        ur, vr, _, _ = device.flow_to_parallax(band.flow[f][0], band.flow[f][1], Hinv, q, y0=band.y0,
                                               want_parallax=False, want_dists=False)
        out.append(correct_epipole(ur, vr, q, rigid_mask, band.y0, band.frame_h, stats))
    return out


def state_to_host(state, n_labels=N_LABELS):
    """(mu_array, B_array, labels) from the device state (one synchronising copy)."""
    s = state.cpu().numpy()
    return ([float(s[_lib.STATE_MU]), float(s[_lib.STATE_MU + 1])], [float(s[_lib.STATE_B]), float(s[_lib.STATE_B + 1])],
            s[_lib.STATE_LABELS:_lib.STATE_LABELS + n_labels].copy())
