"""numpy/OpenCV restatement of the reference's dense plane+parallax path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All citations are relative to
``/root/reference``.  Array conventions follow the reference: images fp64 [h,w,3] in [0,1], flows
fp32 [h,w], homographies fp64 3x3 mapping reference-frame coordinates to neighbour-frame
coordinates, epipoles fp64 [2] = (x, y); frame index 0 = backward (t-1), 1 = forward (t+1).
"""
import numpy as np
import cv2
from scipy import optimize, special

N_LABELS = 51            # pipeline/build_cost_volume.py:50
CHANNEL_WEIGHTS = np.array([0.01, 1.0, 1.0])   # optimize_structure_single_scale.py:350,371


def _grid(h, w):
    y, x = np.mgrid[:h, :w]
    return x, y


def _ptransform(pts32, M):
    return cv2.perspectiveTransform(pts32[:, None, :], M).reshape(-1, 2)


# --------------------------------------------------------------------------- a1, a6 tail
def get_residual_flow(x, y, u, v, H):
    """utils/flow_homography.py:4-16 -- remove the homography from a flow field."""
    dst = np.stack([(x + u).ravel(), (y + v).ravel()], axis=1).astype(np.float32)
    src = np.stack([x.ravel(), y.ravel()], axis=1).astype(np.float32)
    r = _ptransform(dst, np.linalg.inv(H)) - src
    return r[:, 0].reshape(x.shape), r[:, 1].reshape(x.shape)


def get_full_flow(u_res, v_res, H, x=None, y=None, shape=None):
    """utils/flow_homography.py:20-37 -- add the homography back."""
    if x is None and y is None:
        x, y = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
    dst = np.stack([(x + u_res).ravel(), (y + v_res).ravel()], axis=1).astype(np.float32)
    src = np.stack([x.ravel(), y.ravel()], axis=1).astype(np.float32)
    r = _ptransform(dst, H) - src
    return r[:, 0].reshape(x.shape), r[:, 1].reshape(x.shape)


# --------------------------------------------------------------------------- a2, a3
def flow2parallax(u, v, q):
    """pipeline/compute_structure.py:517-537."""
    h, w = u.shape
    x, y = _grid(h, w)
    fx = q[0] - x
    fy = q[1] - y
    dists = np.sqrt(fx ** 2 + fy ** 2)
    nx = fx / np.maximum(dists, 1e-3)
    ny = fy / np.maximum(dists, 1e-3)
    return u * nx + v * ny, np.dstack((nx, ny)), dists


def parallax2normedstructure(parallax, q, B, mu):
    """pipeline/compute_structure.py:25-33."""
    h, w = parallax.shape
    x, y = _grid(h, w)
    dst = np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        return parallax / (B * (parallax / mu - dst / mu))


# --------------------------------------------------------------------------- a6
def structure_to_flow(structure, q, mu, H, b):
    """pipeline/structure2flow.py:14-38."""
    h, w = structure.shape
    x, y = _grid(h, w)
    dists = np.sqrt((x - q[0]) ** 2 + (y - q[1]) ** 2)
    n = dists / mu
    with np.errstate(divide="ignore", invalid="ignore"):
        parallax = (b * structure * n) / (b * structure / mu - 1)
        xv = (q[0] - x) / np.maximum(1e-3, dists)
        yv = (q[1] - y) / np.maximum(1e-3, dists)
        return get_full_flow(xv * parallax, yv * parallax, H, shape=(h, w))


def combine_flow(structure, flow_init, rigidity, q_ar, mu_ar, H_ar, b_ar):
    """pipeline/structure2flow.py:41-82 (method 'simple')."""
    u, v = structure_to_flow(structure, q_ar[1], mu_ar[1], H_ar[1], b_ar[1])
    nr = rigidity == 0
    u[nr] = flow_init[0][nr]
    v[nr] = flow_init[1][nr]
    return rigidity, u, v


# --------------------------------------------------------------------------- a7
def get_unaries_rigid(x, y, u_res, v_res, qx, qy, sigma=1.0, prior_rigid=0.5, prior_nonrigid=0.5):
    """rigidity/inference.py:28-59.  Note the reference's dtypes: ``u_res``/``v_res`` are fp32,
    so ``dist``, ``ang_uv`` and the Bessel argument are fp32 quantities."""
    dist = np.sqrt(u_res ** 2 + v_res ** 2)
    ang = np.arctan2(v_res, u_res) - np.arctan2(qy - y, qx - x)
    off_line = dist * np.sin(ang)
    p = 1.0 / np.sqrt(2 * np.pi * sigma ** 2) * np.exp(-off_line ** 2 / (2 * sigma ** 2))
    c = np.sqrt(2 * np.pi) / sigma * special.ive(0, dist ** 2 / (sigma ** 2 * 4))
    return prior_rigid * p / (prior_rigid * p + prior_nonrigid * c / (2 * np.pi))


# --------------------------------------------------------------------------- a8
def interpolate_through_H(I, H, xn, yn, mode="cubic"):
    """structure_refinement/interpolate_through_H.py:7-51 with compute_derivs=False, sampling by
    structure_refinement/interpolation.py:76-80 (cv2.remap, INTER_CUBIC, BORDER_REPLICATE).
    ``mode='linear'`` swaps in INTER_LINEAR (not a reference mode; north_star's 'bilinear')."""
    with np.errstate(divide="ignore", invalid="ignore"):
        den = xn * H[2, 0] + yn * H[2, 1] + H[2, 2]
        X = (xn * H[0, 0] + yn * H[0, 1] + H[0, 2]) / den
        Y = (xn * H[1, 0] + yn * H[1, 1] + H[1, 2]) / den
        h, w = I.shape[:2]
        mask = ((X >= 0) * (X <= w - 1) * (Y >= 0) * (Y <= h - 1)).astype("float")
    interp = cv2.INTER_CUBIC if mode == "cubic" else cv2.INTER_LINEAR
    J = cv2.remap(I.astype("float32"), X.astype("float32"), Y.astype("float32"),
                  borderMode=cv2.BORDER_REPLICATE, interpolation=interp)
    return J, mask


# --------------------------------------------------------------------------- a10
def _auto_sigma(x):
    # robust_functions.py:23-29: 1.4826 * MAD of the symmetrised sample (+sqrt(x), -sqrt(x)).
    s = np.sqrt(x)
    z = np.c_[s, -s]
    return np.maximum(1e-6, 1.4826 * np.median(np.abs(z - np.median(z))))


def rho(kind, x, sigma=0.0, eps=0.0, a=0.5):
    """utils/robust_functions.py:26-70 -- rho(x^2); ``x`` is already squared.  sigma/eps == 0
    selects the reference's auto-sigma variants."""
    if kind == "quadratic":
        return x
    if kind == "lorentzian":
        if sigma > 0:
            return np.log(1.0 + x / (sigma ** 2))
        s = _auto_sigma(x)
        return s * np.log(1.0 + 0.5 * x / (s ** 2))
    if kind == "charbonnier":
        e = eps if eps != 0 else _auto_sigma(x)
        return np.sqrt(x + e ** 2) if a == 0.5 else (x + e ** 2) ** a
    if kind == "geman_mcclure":
        s = sigma if sigma > 0 else _auto_sigma(x)
        return s * x / (x + s ** 2)
    raise ValueError(kind)


def drho(kind, x, sigma=0.0, eps=0.0, a=0.5):
    """utils/robust_functions.py:26-70 -- d rho / d(x^2)."""
    if kind == "quadratic":
        return np.ones_like(x)
    if kind == "lorentzian":
        if sigma > 0:
            return 1.0 / (sigma ** 2 + x)
        s = _auto_sigma(x)
        return s / (x + 2 * s ** 2)
    if kind == "charbonnier":
        e = eps if eps != 0 else _auto_sigma(x)
        return 1.0 / (2 * np.sqrt(x + e ** 2)) if a == 0.5 else a * (x + e ** 2) ** (a - 1)
    if kind == "geman_mcclure":
        s = sigma if sigma > 0 else _auto_sigma(x)
        return s ** 3 / ((x + s ** 2) ** 2)
    raise ValueError(kind)


# --------------------------------------------------------------------------- a11
def prepare_channels(rgb):
    """optimize_structure_single_scale.py:350-371 with variational_dataterm_grayscale=1
    (parameters.py:46): (gray*255, blur3(d/dy), blur3(d/dx)) as fp64 [h,w,3]."""
    g = cv2.cvtColor(rgb.astype("float32"), cv2.COLOR_RGB2GRAY).astype("float64") * 255.0
    gy, gx = np.gradient(g)
    gy = cv2.GaussianBlur(gy, ksize=(3, 3), sigmaX=-1)
    gx = cv2.GaussianBlur(gx, ksize=(3, 3), sigmaX=-1)
    return np.dstack((g, gy, gx))


def warp_coords(A, q, mu, B, shape):
    """Geometry of computeDataTerm, optimize_structure_single_scale.py:210-235: the point each
    reference pixel moves to (before the homography) for structure ``A`` (scalar or [h,w])."""
    h, w = shape
    x, y = _grid(h, w)
    vx = q[0] - x
    vy = q[1] - y
    nrm = np.maximum(0.5, np.sqrt(vx ** 2 + vy ** 2))
    n = np.sqrt(vx ** 2 + vy ** 2) / mu
    with np.errstate(divide="ignore", invalid="ignore"):
        r = A * B * n / (B * A / mu - 1)
        return x + r * (vx / nrm), y + r * (vy / nrm)


def data_residual(A, ch1, ch2, rigidity, occlusions, H, B, mu, q, cw=CHANNEL_WEIGHTS, mode="cubic"):
    """Residual image Iz of computeDataTerm, optimize_structure_single_scale.py:210-286."""
    xn, yn = warp_coords(A, q, mu, B, ch1.shape[:2])
    J, inside = interpolate_through_H(ch2, H, xn, yn, mode=mode)
    Iz = ((J - ch1) * cw[None, None, :]).mean(axis=2)
    Iz[inside == 0] = 0
    Iz[occlusions == 1] = 0
    Iz[rigidity == 0] = 0
    return Iz


def cost_volume(ch, rigidity, occ, H_ar, B_ar, mu_ar, q_ar, structure_range,
                kind="lorentzian", sigma=1.0, eps=1e-3, mode="cubic", labels=None, frames=(0, 1)):
    """SURVEY.md section 8c composition -- one ``data_residual`` + rho per (frame, label).

    ``ch`` = [channels(I0), channels(I1), channels(I2)].  Backward frame evaluated at
    ``A*mu0/mu1`` with the index-0 parameters (optimize_structure_single_scale.py:435-442).
    Returns cost fp32 [F, L, h, w], min cost fp32 [F,h,w], argmin int32 [F,h,w] (first minimum).
    """
    labels = range(len(structure_range)) if labels is None else labels
    nb = [ch[0], ch[2]]
    out = []
    for f in frames:
        planes = []
        for l in labels:
            a = structure_range[l] * (mu_ar[0] / mu_ar[1] if f == 0 else 1.0)
            Iz = data_residual(a, ch[1], nb[f], rigidity, occ[f], H_ar[f], B_ar[f], mu_ar[f], q_ar[f], mode=mode)
            planes.append(rho(kind, Iz ** 2, sigma=sigma, eps=eps).astype(np.float32))
        out.append(np.stack(planes))
    cost = np.stack(out)
    return cost, cost.min(axis=1), cost.argmin(axis=1).astype(np.int32)


# --------------------------------------------------------------------------- a5
def correct_epipole(u, v, q, rigidity):
    """pipeline/compute_structure.py:453-511.  The reference differentiates its objective with
    chumpy (absent here); the objective is restated with its analytic gradient."""
    h, w = u.shape
    x0, x1 = int(q[0]) - 10, int(q[0]) + 10
    y0, y1 = int(q[1]) - 10, int(q[1]) + 10
    if x0 < 0 or x1 >= w or y0 < 0 or y1 >= h:
        return q
    yy, xx = np.mgrid[y0:y1 + 1, x0:x1 + 1]
    r = rigidity[y0:y1 + 1, x0:x1 + 1]
    if (r == 0).sum() > (r.size / 2):
        return q
    sel = r > 0
    uw = u[y0:y1 + 1, x0:x1 + 1][sel].astype(np.float64)
    vw = v[y0:y1 + 1, x0:x1 + 1][sel].astype(np.float64)
    xw = xx[sel].astype(np.float64)
    yw = yy[sel].astype(np.float64)

    def fun(qe):
        fx = qe[0] - xw
        fy = qe[1] - yw
        d2 = fx ** 2 + fy ** 2
        den = np.maximum(1e-3, d2)
        par = uw * fx + vw * fy
        live = (d2 > 1e-3).astype(np.float64)
        sgn = np.sign(par)
        gx = sgn * uw / den - np.abs(par) / den ** 2 * (2 * fx * live)
        gy = sgn * vw / den - np.abs(par) / den ** 2 * (2 * fy * live)
        return (np.abs(par) / den).sum(), np.array([gx.sum(), gy.sum()])

    res = optimize.minimize(fun, x0=np.asarray(q, dtype=np.float64), jac=True, method="BFGS",
                            options={"disp": False})
    return res.x if np.linalg.norm(res.x - q) < 10 else q


# --------------------------------------------------------------------------- a4
class Params:
    """The fields of parameters.default_vals (parameters.py:4-55) read on the hot path."""
    debug_use_rigidity_gt = 0
    debug_save_frames = 0
    debug_compute_figure = 0
    rigidity_sigma_direction = 1.0
    rigidity_weight_cnn = 0.25
    rigidity_sigma_structure = 1.0
    rigidity_use_structure = 1
    scale_structure = 1
    nonlinear_structure_initialization = 0   # refine_A needs chumpy (absent): SURVEY.md 8c
    occlusion_reasoning = 1
    tempdir = "."

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def compute_structure(images, flow, rigidity, occlusions_precomputed, homographies, epipoles, params,
                      infer_mrf=None, return_intermediates=False):
    """pipeline/compute_structure.py:37-449 with nonlinear_structure_initialization=0.

    ``infer_mrf(I, unaries, lambd)`` is the consumer at :397 (rigidity/inference.py:99-135, the
    C++ solver in extern/eff_trws); pass the callable (``oracle.trws.infer_mrf`` when built).
    """
    if not params.debug_use_rigidity_gt:
        rig_t = cv2.erode((rigidity > 0.5).astype("uint8"), np.ones((3, 3), np.uint8)) > 0
    else:
        rig_t = rigidity
    h, w = images[0].shape[:2]
    x, y = _grid(h, w)
    res, q_new, par, mu_ar, dn = [], [], [], [], []
    for f in range(2):                                                        # :67-91
        u_res, v_res = get_residual_flow(x, y, flow[f][0], flow[f][1], homographies[f])
        qn = correct_epipole(u_res, v_res, epipoles[f], rig_t > 0)
        p, _, d = flow2parallax(u_res, v_res, qn)
        mu = d.mean()
        res.append([u_res, v_res]); q_new.append(qn); par.append(p); mu_ar.append(mu); dn.append(d / mu)

    p_cnn = rigidity.copy()                                                   # :104-146
    sd = params.rigidity_sigma_direction
    small = [np.sqrt(res[f][0] ** 2 + res[f][1] ** 2) < 0.1 for f in range(2)]
    pd = [get_unaries_rigid(x, y, res[f][0], res[f][1], q_new[f][0], q_new[f][1], sigma=sd) for f in range(2)]
    pd[0][small[0]] = 0.6
    pd[1][small[1]] = 0.6
    occ_b = occlusions_precomputed[0] == 1
    occ_f = occlusions_precomputed[1] == 1
    pd[1][occ_f] = 0.5
    pd[0][occ_b] = 0.5
    p_dir = (pd[0] + pd[1]) / 2.0
    p_dir[occ_b] = 0.25 + 0.5 * pd[1][occ_b]
    p_dir[occ_f] = 0.25 + 0.5 * pd[0][occ_f]
    p_dir[occ_b * occ_f] = 0.5

    wc = params.rigidity_weight_cnn                                            # :151-165
    p_thr = (wc * p_cnn + (1 - wc) * p_dir) > 0.5
    if p_thr.sum() < p_thr.size * 0.25:
        raise Exception("TooMuchNonRigid")

    pv = np.logical_and(np.logical_and(np.abs(par[0]) > 0.1, np.abs(par[1]) > 0.1), p_thr > 0)   # :172
    mu0, mu1 = mu_ar
    if pv.sum() < 100:                                                        # :183-198
        B_fwd = 1.0
    else:
        s1 = parallax2normedstructure(par[1], q_new[1], 1.0, mu1)
        if params.scale_structure == 1:
            B_fwd = 1.426 * np.median(np.abs(s1[pv] - np.median(s1[pv])))
        elif params.scale_structure == 2:
            B_fwd = s1[pv].std()
        else:
            B_fwd = 1.0
    if pv.sum() < 100:                                                        # :204-221
        B_b = -B_fwd
    else:
        p0, p1 = par[0][pv], par[1][pv]
        target = p1 / (B_fwd * (p1 / mu1 - dn[1][pv]))
        template = mu1 / mu0 * p0 / ((p0 / mu0 - dn[0][pv]))
        B_b = np.median(template / target)
    B_ar = [B_b, B_fwd]
    S = [parallax2normedstructure(par[f], q_new[f], B_ar[f], mu_ar[f]) for f in range(2)]   # :240-246

    if params.occlusion_reasoning > 0:                                        # :364-367
        occ = occlusions_precomputed
    else:
        occ = [np.ones((h, w)) > 0, np.ones((h, w)) > 0]

    inter = dict(p_dir=p_dir, p_dir_f=pd, p_thr=p_thr, pv=pv, parallax=par, residual=res)
    rig_ref = None
    if params.rigidity_use_structure:                                         # :374-398
        ss = params.rigidity_sigma_structure
        A_diff = S[0] * mu1 / mu0 - S[1]
        p_str = np.exp(-(A_diff / ss) ** 2)
        bad = ((occ[0] == 0) * (occ[1] == 0)) == 0
        p_str[bad] = 0.5
        p_mot = p_dir * p_str
        p_mot[bad] = 0.25 + 0.5 * p_dir[bad]
        p_rig = wc * p_cnn + (1 - wc) * p_mot
        unaries = np.dstack((1 - p_rig, p_rig))
        inter.update(p_str=p_str, p_rig=p_rig, unaries=unaries)
        rig_ref = infer_mrf(images[1], unaries, 1.1) > 0.5
    if params.debug_use_rigidity_gt:
        rig_ref = rigidity
    if (rig_ref == 1).sum() < 0.25 * rig_ref.size:                            # :427-429
        raise Exception("TooFewStructureMatches")
    out = [S, homographies, mu_ar, B_ar, q_new, rig_ref, occ]
    return (out, inter) if return_intermediates else out


def optimize_override(structures, mu, occlusions, occlusion_reasoning=1):
    """pipeline/optimize_variational_refinement.py:20-42,72-73 -- what reaches combine_flow when
    ``override_optimization=1`` (BASELINE.json configs[0]): 3x3 median of the two occlusion maps, then
    bwd structure (rescaled) where only the fwd frame is occluded, fwd structure where only the bwd frame
    is occluded, their mean elsewhere."""
    occ_b = cv2.medianBlur(occlusions[0].astype("uint8"), ksize=3) > 0
    occ_f = cv2.medianBlur(occlusions[1].astype("uint8"), ksize=3) > 0
    fwd_only = (occ_b == 0) * (occ_f > 0)
    bwd_only = (occ_f == 0) * (occ_b > 0)
    none = (fwd_only == 0) * (bwd_only == 0)
    if occlusion_reasoning == 0:
        bwd_only[:] = True
        none[:] = False
    with np.errstate(invalid="ignore"):
        S0 = structures[0] * mu[1] / mu[0]
        S1 = structures[1]
        return fwd_only * S0 + bwd_only * S1 + none * (S0 + S1) / 2.0


# --------------------------------------------------------------------------- a12
def structure_range(S, rigidity, q_ar, n_labels=N_LABELS, range_mask="as_written"):
    """Label set of pipeline/build_cost_volume.py:154-169,190-192: 1.5x the min/max of both
    normed structures after zeroing a +-10 px box round the epipole and (as written, :159) every
    pixel with rigidity > 0; ``range_mask='rigid_only'`` zeroes the non-rigid ones instead."""
    h, w = S[0].shape
    rem = []
    for f in range(2):
        s = S[f].copy()
        q = q_ar[f]
        bx0 = max(0, min(w - 1, int(q[0] - 10)))
        bx1 = max(0, min(w - 1, int(q[0] + 10)))
        by0 = max(0, min(h - 1, int(q[1] - 10)))
        by1 = max(0, min(h - 1, int(q[1] + 10)))
        if range_mask == "as_written":
            s[rigidity > 0] = 0
        else:
            s[~(rigidity > 0)] = 0
        if not (bx0 == w - 1 or bx1 == 0 or by0 == h - 1 or by1 == 0):
            s[by0:by1, bx0:bx1] = 0
        rem.append(s)
    rem = np.array(rem)
    return np.linspace(1.5 * rem.min(), 1.5 * rem.max(), n_labels)


def build_cost_volume(images, flow, rigidity, homographies, epipoles, params,
                      kind="lorentzian", sigma=1.0, mode="cubic", range_mask="as_written"):
    """pipeline/build_cost_volume.py:42-231 signature and 6-list return, with the live
    conventions of compute_structure.py for parallax -> structure (SURVEY.md section 8a note on
    a12) and the section-8c composition standing in for the missing extension (:209-219).
    Occlusions are not an input of this stage: none are applied."""
    h, w = images[0].shape[:2]
    x, y = _grid(h, w)
    par, q_new, mu_ar, dn = [], [], [], []
    for f in range(2):
        u_res, v_res = get_residual_flow(x, y, flow[f][0], flow[f][1], homographies[f])
        qn = correct_epipole(u_res, v_res, epipoles[f], rigidity > 0)
        p, _, d = flow2parallax(u_res, v_res, qn)
        par.append(p); q_new.append(qn); mu_ar.append(d.mean()); dn.append(d / mu_ar[-1])
    pv = np.logical_and(np.logical_and(np.abs(par[0]) > 0.1, np.abs(par[1]) > 0.1), rigidity > 0)  # :116
    mu0, mu1 = mu_ar
    with np.errstate(divide="ignore", invalid="ignore"):
        target = par[1][pv] / (1.0 * (par[1][pv] / mu1 - dn[1][pv]))
        template = mu1 / mu0 * par[0][pv] / (par[0][pv] / mu0 - dn[0][pv])
        B_b = np.median(template / target)                                    # :120-131
    B_ar = [B_b, 1.0]
    S = [parallax2normedstructure(par[f], q_new[f], B_ar[f], mu_ar[f]) for f in range(2)]
    rng = structure_range(S, rigidity, q_new, range_mask=range_mask)
    ch = [prepare_channels(I) for I in images]
    noocc = [np.zeros((h, w), bool), np.zeros((h, w), bool)]
    cost, _, _ = cost_volume(ch, rigidity, noocc, homographies, B_ar, mu_ar, q_new, rng,
                             kind=kind, sigma=sigma, mode=mode)
    return [[cost[0], cost[1]], S, mu_ar, B_ar, q_new, rng]


# --------------------------------------------------------------------------- section 8f rows
def compute_occlusions_from_consistency(flow, backflow, threshold=7.0):
    """rigidity/occlusions.py:10-74 -- forward/backward consistency occlusion maps (cv2.remap,
    INTER_LINEAR, default BORDER_CONSTANT).  Returns (occ_backward, occ_forward, occ_both)."""
    def warped_error(uf, vf, ub, vb):
        y, x = np.mgrid[:uf.shape[0], :uf.shape[1]]
        mx, my = (x + uf).astype("float32"), (y + vf).astype("float32")
        uw = cv2.remap(ub, mx, my, interpolation=cv2.INTER_LINEAR)
        vw = cv2.remap(vb, mx, my, interpolation=cv2.INTER_LINEAR)
        valid = (y + vf >= 0) * (y + vf < y.shape[0]) * (x + uf >= 0) * (x + uf < x.shape[1])
        return np.sqrt((uw + uf) ** 2 + (vw + vf) ** 2), valid == 0

    eb, ib = warped_error(flow[0][0], flow[0][1], backflow[0][0], backflow[0][1])
    ef, i_f = warped_error(flow[1][0], flow[1][1], backflow[1][0], backflow[1][1])
    ob = np.logical_or(eb > threshold, ib)
    of = np.logical_or(ef > threshold, i_f)
    ob[np.logical_and(i_f, ib == 0)] = 0
    of[np.logical_and(ib, i_f == 0)] = 0
    return ob, of, ob * of


def get_image_weights(I):
    """rigidity/inference.py:140-167 -- contrast-sensitive 8-neighbourhood edge weights for the MRF."""
    if I.ndim == 2:
        I = I[:, :, np.newaxis]
    h, w, c = I.shape
    d0 = np.vstack((np.diff(I, axis=0).sum(axis=2) / c, np.zeros((1, w))))
    d1 = np.hstack((np.diff(I, axis=1).sum(axis=2) / c, np.zeros((h, 1))))
    dne = np.zeros((h, w)); dne[1:, :-1] = (I[1:, :-1, :] - I[:-1, 1:, :]).sum(axis=2) / c
    dse = np.zeros((h, w)); dse[:-1, :-1] = (I[:-1, :-1, :] - I[1:, 1:, :]).sum(axis=2) / c
    out = []
    for d in (d1, d0, dne, dse):                      # returned order: horizontal, vertical, ne, se
        beta = 1.0 / (2 * max(1e-6, (d ** 2).mean()))
        out.append(np.exp(-d ** 2 * beta).astype("float32"))
    return tuple(out)


def pack_unaries(unaries):
    """rigidity/inference.py:113 -- what Eff_TRWS.solve is fed."""
    return ((-unaries) * 1000).astype("int32").copy("C")


# --------------------------------------------------------------------------- section 8f.1: a9 + a11 linearisation
def compute_derivs_through_H(I, H):
    """structure_refinement/interpolate_through_H.py:53-97 -- derivatives of J(x,y) = I(H(x,y)) by four
    cv2.warpPerspective calls (bilinear, BORDER_REPLICATE) and finite-difference scalings."""
    h, w = I.shape[:2]
    Hinv = np.linalg.inv(H)
    Ts = [np.array([[1.0, 0, -1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), np.array([[1.0, 0, 1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
          np.array([[1.0, 0, 0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]]), np.array([[1.0, 0, 0], [0.0, 1.0, 1.0], [0.0, 0.0, 1.0]])]
    Ms = [Hinv.dot(T).dot(H) for T in Ts]
    y, x = np.mgrid[:h, :w]
    xy = np.c_[x.flatten(), y.flatten()].astype("float")
    ds, Iw = [], []
    for M in Ms:
        p = cv2.perspectiveTransform(xy[:, np.newaxis, :], M).squeeze()
        d = np.linalg.norm(xy - p, axis=1).reshape((h, w))
        ds.append(d[:, :, np.newaxis] if I.ndim > 2 else d)
        Iw.append(cv2.warpPerspective(I, M, (w, h), borderMode=cv2.BORDER_REPLICATE))
    dx = 0.5 * ((Iw[0] - I) * ds[0] + (I - Iw[1]) * ds[1])
    dy = 0.5 * ((Iw[2] - I) * ds[2] + (I - Iw[3]) * ds[3])
    return dx, dy


def interpolate_through_H_derivs(I, H, xn, yn, dIdx=None, dIdy=None):
    """interpolate_through_H.py:7-51 with compute_derivs=True.  Returns (J, mask, dJdx, dJdy)."""
    J, mask = interpolate_through_H(I, H, xn, yn)
    if dIdx is None or dIdy is None:
        dIdx, dIdy = compute_derivs_through_H(I, H)
    dJdy, _ = interpolate_through_H(dIdy, H, xn, yn)
    dJdx, _ = interpolate_through_H(dIdx, H, xn, yn)
    return J, mask, dJdx, dJdy


def compute_data_term(A, I1, I2, rigidity, occlusions, H, B, mu, q, kind="lorentzian", sigma=0.0,
                      cw=CHANNEL_WEIGHTS, return_intermediates=False):
    """optimize_structure_single_scale.py:186-305 (computeDataTerm): the linearised data term of the
    variational refinement.  Returns (diag_data_image, b_data_image, E_data) -- the two images after
    their 5x5 median filters (:299-300), i.e. the diagonal of A_data and b_data before flattening."""
    from scipy import ndimage
    h, w = I1.shape[:2]
    x, y = _grid(h, w)
    vx = q[0] - x
    vy = q[1] - y
    nrm = np.maximum(0.5, np.sqrt(vx ** 2 + vy ** 2))
    t1 = vx / nrm
    t2 = vy / nrm
    n = np.sqrt(vx ** 2 + vy ** 2) / mu
    with np.errstate(divide="ignore", invalid="ignore"):
        r = A * B * n / (B * A / mu - 1)
        dr_da = - (n * B) / ((B * A / mu - 1) ** 2)
    xn = x + r * t1
    yn = y + r * t2
    I2w, inside, dI2w_dx, dI2w_dy = interpolate_through_H_derivs(I2, H, xn, yn)
    dI2w_dx[inside == 0, :] = 0
    dI2w_dy[inside == 0, :] = 0
    _, _, dI1w_dx, dI1w_dy = interpolate_through_H_derivs(I1, np.eye(3), x, y)
    dI2w_dx += dI1w_dx
    dI2w_dy += dI1w_dy
    dI2w_dx /= (inside[:, :, np.newaxis] + 1)
    dI2w_dy /= (inside[:, :, np.newaxis] + 1)
    c = cw[np.newaxis, np.newaxis, :]
    dx = (dI2w_dx * c).mean(axis=2)
    dy = (dI2w_dy * c).mean(axis=2)
    Iz = ((I2w - I1) * c).mean(axis=2)
    dphi = t1 * dx + t2 * dy
    Iz[inside == 0] = 0
    Iz[occlusions == 1] = 0
    Iz[rigidity == 0] = 0
    psi = drho(kind, Iz ** 2, sigma=sigma)
    diag = psi * dphi ** 2 * dr_da ** 2
    b = psi * Iz * dphi * dr_da
    diag_m = ndimage.median_filter(diag, size=5)
    b_m = ndimage.median_filter(b, size=5)
    E = rho(kind, Iz ** 2, sigma=sigma).sum()
    if return_intermediates:
        return diag_m, b_m, E, dict(Iz=Iz, dphi=dphi, dr_da=dr_da, psi=psi, diag=diag, b=b)
    return diag_m, b_m, E


def flow_read(filename):
    """utils/flow_io.py:33-48 (flow_read): tag, width, height, then height x (2 width) fp32 with u in the even and v
    in the odd columns.  Same assertions."""
    with open(filename, "rb") as f:
        check = np.fromfile(f, dtype=np.float32, count=1)[0]
        assert check == 202021.25, " flow_read:: Wrong tag in flow file"                        # :38-39
        width = np.fromfile(f, dtype=np.int32, count=1)[0]
        height = np.fromfile(f, dtype=np.int32, count=1)[0]
        size = int(width) * int(height)
        assert width > 0 and height > 0 and size > 1 and size < 100000000                       # :43
        tmp = np.fromfile(f, dtype=np.float32, count=-1).reshape((height, width * 2))           # :44
    return tmp[:, np.arange(width) * 2], tmp[:, np.arange(width) * 2 + 1]                       # :45-46
