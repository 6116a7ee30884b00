"""Device-level wrappers: torch CUDA tensors in, torch CUDA tensors out, one C-ABI call each.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libmrflow_b200.so.
Every function runs on the current CUDA stream of the tensors' device and is asynchronous.
Names follow the reference's domain; each docstring cites the reference lines it replaces
(paths relative to the reference repository root).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CostVolParams, INTERP, LAYOUT, RHO, VARIANT, check, dvec

F32, F64, U8, I32 = torch.float32, torch.float64, torch.uint8, torch.int32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA %s tensor" % (name, dtype))


def _mat(H):
    return dvec(np.asarray(H, dtype=np.float64).reshape(9))


def _pt(q):
    return dvec(np.asarray(q, dtype=np.float64).reshape(2))


def launch_count():
    return int(_lib.lib().mrf_launch_count())


def reset_launch_count():
    _lib.lib().mrf_reset_launch_count()


# ------------------------------------------------------------------------------------ a1 + a2
def flow_to_parallax(u, v, Hinv, q, y0=0, want_residual=True, want_parallax=True, want_dists=True):
    """utils/flow_homography.py:4-16 (get_residual_flow, ``Hinv`` = inv(H)) fused with
    pipeline/compute_structure.py:517-537 (flow2parallax).  Returns (u_res, v_res, parallax, dists)."""
    _chk(u, F32, "u"); _chk(v, F32, "v")
    rows, w = u.shape
    ur = torch.empty_like(u) if want_residual else None
    vr = torch.empty_like(u) if want_residual else None
    par = torch.empty((rows, w), dtype=F64, device=u.device) if want_parallax else None
    dst = torch.empty((rows, w), dtype=F64, device=u.device) if want_dists else None
    check(_lib.lib().mrf_flow_to_parallax(_p(u), _p(v), rows, w, y0, _mat(Hinv), _pt(q), _p(ur), _p(vr), _p(par),
                                          _p(dst), _stream()), "mrf_flow_to_parallax")
    return ur, vr, par, dst


def flow2parallax(u, v, q, y0=0):
    """pipeline/compute_structure.py:517-537 for a given flow (fp64).  Returns (parallax, foe_vectors
    [rows,w,2], dists)."""
    _chk(u, F64, "u"); _chk(v, F64, "v")
    rows, w = u.shape
    par = torch.empty((rows, w), dtype=F64, device=u.device)
    vec = torch.empty((rows, w, 2), dtype=F64, device=u.device)
    dst = torch.empty((rows, w), dtype=F64, device=u.device)
    check(_lib.lib().mrf_flow2parallax(_p(u), _p(v), rows, w, y0, _pt(q), _p(par), _p(vec), _p(dst), _stream()),
          "mrf_flow2parallax")
    return par, vec, dst


def epipole_dist_sum(rows, w, q, y0=0, device="cuda"):
    """Sum over the band of ||p - q|| (the numerator of mu, pipeline/compute_structure.py:84).
    Returns a 1-element fp64 tensor."""
    L = _lib.lib()
    out = torch.empty(1, dtype=F64, device=device)
    scratch = torch.empty(L.mrf_reduce_scratch_bytes(), dtype=U8, device=device)
    check(L.mrf_epipole_dist_sum(rows, w, y0, _pt(q), _p(out), _p(scratch), _stream()), "mrf_epipole_dist_sum")
    return out


def parallax_to_structure(parallax, q, B, mu, y0=0):
    """pipeline/compute_structure.py:25-33 (parallax2normedstructure)."""
    _chk(parallax, F64, "parallax")
    rows, w = parallax.shape
    out = torch.empty_like(parallax)
    check(_lib.lib().mrf_parallax_to_structure(_p(parallax), rows, w, y0, _pt(q), float(B), float(mu), _p(out),
                                               _stream()), "mrf_parallax_to_structure")
    return out


# ------------------------------------------------------------------------------------ a6
def structure_to_flow(structure, q, mu, H, b, y0=0, nonrigid=None, init_u=None, init_v=None, state=None, frame=1):
    """pipeline/structure2flow.py:14-38 incl. utils/flow_homography.py:20-37; with ``nonrigid`` also
    the paste of combine_flow_simple (:72-73)."""
    _chk(structure, F64, "structure"); _chk(nonrigid, U8, "nonrigid"); _chk(init_u, F32, "init_u"); _chk(init_v, F32, "init_v")
    rows, w = structure.shape
    u = torch.empty((rows, w), dtype=F32, device=structure.device)
    v = torch.empty_like(u)
    check(_lib.lib().mrf_structure_to_flow(_p(structure), rows, w, y0, _pt(q), float(mu), _mat(H), float(b),
                                           _p(nonrigid), _p(init_u), _p(init_v), _p(u), _p(v), _p(state), int(frame),
                                           _stream()),
          "mrf_structure_to_flow")
    return u, v


def full_flow(u_res, v_res, H, y0=0):
    """utils/flow_homography.py:20-37 (get_full_flow); residual fp64 -> full flow fp32."""
    _chk(u_res, F64, "u_res"); _chk(v_res, F64, "v_res")
    rows, w = u_res.shape
    u = torch.empty((rows, w), dtype=F32, device=u_res.device)
    v = torch.empty_like(u)
    check(_lib.lib().mrf_full_flow(_p(u_res), _p(v_res), rows, w, y0, _mat(H), _p(u), _p(v), _stream()), "mrf_full_flow")
    return u, v


# ------------------------------------------------------------------------------------ a7, a4
def rigidity_unaries(u_res, v_res, q, sigma=1.0, prior_rigid=0.5, prior_nonrigid=0.5, y0=0):
    """rigidity/inference.py:28-59 (get_unaries_rigid)."""
    _chk(u_res, F32, "u_res"); _chk(v_res, F32, "v_res")
    rows, w = u_res.shape
    out = torch.empty((rows, w), dtype=F64, device=u_res.device)
    check(_lib.lib().mrf_rigidity_unaries(_p(u_res), _p(v_res), rows, w, y0, _pt(q), float(sigma), float(prior_rigid),
                                          float(prior_nonrigid), _p(out), _stream()), "mrf_rigidity_unaries")
    return out


def rigidity_direction(res0, res1, occ0, occ1, p_cnn, q0, q1, sigma_direction, weight_cnn, y0=0):
    """pipeline/compute_structure.py:104-153.  Returns (p_dir fp64, p_thr uint8, count int32[1] =
    p_thr.sum())."""
    for t in (*res0, *res1):
        _chk(t, F32, "residual flow")
    _chk(occ0, U8, "occ0"); _chk(occ1, U8, "occ1"); _chk(p_cnn, F64, "p_cnn")
    rows, w = res0[0].shape
    p_dir = torch.empty((rows, w), dtype=F64, device=p_cnn.device)
    p_thr = torch.empty((rows, w), dtype=U8, device=p_cnn.device)
    cnt = torch.empty(1, dtype=I32, device=p_cnn.device)
    check(_lib.lib().mrf_rigidity_direction(_p(res0[0]), _p(res0[1]), _p(res1[0]), _p(res1[1]), _p(occ0), _p(occ1),
                                            _p(p_cnn), rows, w, y0, _pt(q0), _pt(q1), float(sigma_direction),
                                            float(weight_cnn), _p(p_dir), _p(p_thr), _p(cnt), _stream()),
          "mrf_rigidity_direction")
    return p_dir, p_thr, cnt


def structure_candidates(par0, par1, p_thr, q0, q1, mu0, mu1, B_fwd, want, y0=0, pv_out=True):
    """pipeline/compute_structure.py:172-221: valid mask ``pv`` and the compacted candidate list whose
    median gives B_fwd (want=1) or B_b (want=2).  Returns (pv uint8, cand fp64 [capacity], count int32[1])."""
    _chk(par0, F64, "par0"); _chk(par1, F64, "par1"); _chk(p_thr, U8, "p_thr")
    rows, w = par0.shape
    dev = par0.device
    pv = torch.empty((rows, w), dtype=U8, device=dev) if pv_out else None
    cand = torch.empty(rows * w, dtype=F64, device=dev) if want else None
    count = torch.zeros(1, dtype=I32, device=dev)
    check(_lib.lib().mrf_structure_candidates(_p(par0), _p(par1), _p(p_thr), rows, w, y0, _pt(q0), _pt(q1), float(mu0),
                                              float(mu1), float(B_fwd), int(want), _p(pv), _p(cand), _p(count),
                                              _stream()), "mrf_structure_candidates")
    return pv, cand, count


def rigidity_structure(S0, S1, p_dir, p_cnn, occ0, occ1, mu0, mu1, sigma_structure, weight_cnn, want_i32=False):
    """pipeline/compute_structure.py:374-394 (+ the int32 packing of rigidity/inference.py:113).
    Returns (unaries fp64 [rows,w,2], unaries int32 [rows,w,2] or None)."""
    for t in (S0, S1, p_dir, p_cnn):
        _chk(t, F64, "fp64 plane")
    _chk(occ0, U8, "occ0"); _chk(occ1, U8, "occ1")
    rows, w = S0.shape
    un = torch.empty((rows, w, 2), dtype=F64, device=S0.device)
    ui = torch.empty((rows, w, 2), dtype=I32, device=S0.device) if want_i32 else None
    check(_lib.lib().mrf_rigidity_structure(_p(S0), _p(S1), _p(p_dir), _p(p_cnn), _p(occ0), _p(occ1), rows, w,
                                            float(mu0), float(mu1), float(sigma_structure), float(weight_cnn), _p(un),
                                            _p(ui), _stream()), "mrf_rigidity_structure")
    return un, ui


def merge_structures(S0, S1, occ_bwd, occ_fwd, mu0, mu1, occlusion_reasoning=1, want_blurred=False):
    """pipeline/optimize_variational_refinement.py:20-42 -- occlusion-aware merge of the two structures
    (and, optionally, the 3x3-median-filtered occlusion maps of :23-24)."""
    _chk(S0, F64, "S0"); _chk(S1, F64, "S1"); _chk(occ_bwd, U8, "occ_bwd"); _chk(occ_fwd, U8, "occ_fwd")
    h, w = S0.shape
    out = torch.empty_like(S0)
    ob = torch.empty_like(occ_bwd) if want_blurred else None
    of = torch.empty_like(occ_fwd) if want_blurred else None
    check(_lib.lib().mrf_merge_structures(_p(S0), _p(S1), _p(occ_bwd), _p(occ_fwd), h, w, float(mu0), float(mu1),
                                          int(occlusion_reasoning), _p(out), _p(ob), _p(of), _stream()),
          "mrf_merge_structures")
    return (out, ob, of) if want_blurred else out


def erode3x3(mask_u8):
    """pipeline/compute_structure.py:44 -- cv2.erode(uint8, ones((3,3)))."""
    _chk(mask_u8, U8, "mask")
    h, w = mask_u8.shape
    out = torch.empty_like(mask_u8)
    check(_lib.lib().mrf_erode3x3_u8(_p(mask_u8), h, w, _p(out), _stream()), "mrf_erode3x3_u8")
    return out


# ------------------------------------------------------------------------------------ a4 with device-resident scalars
class StructureBuffers:
    """Reusable HBM buffers of the device-resident compute_structure stage for one frame size."""

    def __init__(self, rows, w, dev="cuda"):
        L = _lib.lib()
        n = rows * w
        self.state = torch.zeros(_lib.STATE_WORDS, dtype=F64, device=dev)
        self.cand = torch.empty(n, dtype=F64, device=dev)
        self.dev_abs = torch.empty(n, dtype=F64, device=dev)
        self.count1 = torch.zeros(1, dtype=I32, device=dev)
        self.count2 = torch.zeros(1, dtype=I32, device=dev)
        self.med = torch.zeros(3, dtype=F64, device=dev)
        self.scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=U8, device=dev)
        self.sel = torch.empty(L.mrf_select_scratch_bytes(n), dtype=U8, device=dev)


def structure_scales_dev(par0, par1, p_thr, q0, q1, frame_h, buf, scale_structure=1):
    """pipeline/compute_structure.py:84,172-223 with nothing coming back to the host: mu of both frames, the
    valid mask, B_fwd = 1.426 * MAD of the forward structure at B = 1 (or 1.0), B_b = the median ratio (or
    -B_fwd) -- all written to ``buf.state``.  Returns the valid mask ``pv`` (uint8).  ``scale_structure`` 1 (MAD)
    or anything but 2 (-> 1.0); the std variant (2) is not offered here."""
    L = _lib.lib()
    rows, w = par0.shape
    p = _lib.FrontParams()
    p.rows, p.w, p.y0, p.frame_h = rows, w, 0, int(frame_h)
    p.q0 = _pt(q0); p.q1 = _pt(q1)
    st = _stream()
    check(L.mrf_front_mu_whole(C.byref(p), _p(buf.state), _p(buf.scratch), st), "mrf_front_mu_whole")            # :84
    pv = torch.empty((rows, w), dtype=U8, device=par0.device)
    cap = rows * w
    # B_fwd (:183-198): MAD of parallax2normedstructure(par1, q1, 1.0, mu1) over the valid pixels
    check(L.mrf_structure_candidates_dev(_p(par0), _p(par1), _p(p_thr), rows, w, 0, _pt(q0), _pt(q1), _p(buf.state), 1, _p(pv),
                                         _p(buf.cand), _p(buf.count1), st), "mrf_structure_candidates_dev")
    check(L.mrf_select_median_dev(_p(buf.cand), _p(buf.count1), cap, _p(buf.med[0:1]), _p(buf.sel), st), "mrf_select_median_dev")
    check(L.mrf_abs_deviation_dev(_p(buf.cand), _p(buf.count1), cap, _p(buf.med[0:1]), _p(buf.dev_abs), st), "mrf_abs_deviation_dev")
    check(L.mrf_select_median_dev(_p(buf.dev_abs), _p(buf.count1), cap, _p(buf.med[1:2]), _p(buf.sel), st), "mrf_select_median_dev")
    check(L.mrf_scales_finalize(_p(buf.state), _p(buf.count1), _p(buf.med[1:2]), 1, int(scale_structure), st), "mrf_scales_finalize")
    # B_b (:204-221): median of template / target with that B_fwd
    check(L.mrf_structure_candidates_dev(_p(par0), _p(par1), _p(p_thr), rows, w, 0, _pt(q0), _pt(q1), _p(buf.state), 2, None,
                                         _p(buf.cand), _p(buf.count2), st), "mrf_structure_candidates_dev")
    check(L.mrf_select_median_dev(_p(buf.cand), _p(buf.count2), cap, _p(buf.med[2:3]), _p(buf.sel), st), "mrf_select_median_dev")
    check(L.mrf_scales_finalize(_p(buf.state), _p(buf.count1), _p(buf.med[2:3]), 2, int(scale_structure), st), "mrf_scales_finalize")
    return pv


def parallax_to_structure_dev(parallax, q, state, frame):
    """pipeline/compute_structure.py:25-33 with B and mu of ``frame`` read from the device state."""
    _chk(parallax, F64, "parallax"); _chk(state, F64, "state")
    rows, w = parallax.shape
    out = torch.empty_like(parallax)
    check(_lib.lib().mrf_parallax_to_structure_dev(_p(parallax), rows, w, 0, _pt(q), _p(state), int(frame), _p(out), _stream()),
          "mrf_parallax_to_structure_dev")
    return out


def rigidity_structure_dev(S0, S1, p_dir, p_cnn, occ0, occ1, state, sigma_structure, weight_cnn, want_i32=False):
    """pipeline/compute_structure.py:374-394 with mu of both frames read from the device state."""
    rows, w = S0.shape
    un = torch.empty((rows, w, 2), dtype=F64, device=S0.device)
    ui = torch.empty((rows, w, 2), dtype=I32, device=S0.device) if want_i32 else None
    check(_lib.lib().mrf_rigidity_structure_dev(_p(S0), _p(S1), _p(p_dir), _p(p_cnn), _p(occ0), _p(occ1), rows, w, _p(state),
                                                float(sigma_structure), float(weight_cnn), _p(un), _p(ui), _stream()),
          "mrf_rigidity_structure_dev")
    return un, ui


# ------------------------------------------------------------------------------------ statistics
def select_kth(keys, k, n=None):
    """k-th smallest (0-based) of the first ``n`` fp64 keys, its successor and the NaN count, as a
    3-element fp64 device tensor (no host synchronisation)."""
    _chk(keys, F64, "keys")
    L = _lib.lib()
    n = keys.numel() if n is None else int(n)
    out = torch.empty(3, dtype=F64, device=keys.device)
    scratch = torch.empty(L.mrf_select_scratch_bytes(n), dtype=U8, device=keys.device)
    check(L.mrf_select_kth_f64(_p(keys), n, int(k), _p(out), _p(scratch), _stream()), "mrf_select_kth_f64")
    return out


def median_dev(keys, n=None, out=None):
    """numpy.median of the first ``n`` keys as a 1-element fp64 device tensor (no host synchronisation): the
    fused radix select of the front (8 histogram passes that each resolve the previous byte, one successor
    pass, mean of the two middle order statistics for even n, NaN if any key is NaN)."""
    _chk(keys, F64, "keys")
    L = _lib.lib()
    n = keys.numel() if n is None else int(n)
    cnt = torch.full((1,), n, dtype=torch.int32, device=keys.device)      # the kernels read n on the device (no H2D copy)
    out = torch.empty(1, dtype=F64, device=keys.device) if out is None else out
    scratch = torch.empty(L.mrf_select_scratch_bytes(n), dtype=U8, device=keys.device)
    check(L.mrf_select_median_dev(_p(keys), _p(cnt), max(n, 1), _p(out), _p(scratch), _stream()), "mrf_select_median_dev")
    return out


def median(keys, n=None):
    """numpy.median of the first ``n`` keys (mean of the two middle order statistics for even n,
    NaN if any key is NaN).  Synchronises (returns a Python float)."""
    n = keys.numel() if n is None else int(n)
    if n == 0:
        return float("nan")
    return float(median_dev(keys, n).cpu()[0])


def abs_deviation(x, c, n=None):
    """|x - c| (MAD, utils/robust_functions.py:23 / pipeline/compute_structure.py:190)."""
    _chk(x, F64, "x")
    n = x.numel() if n is None else int(n)
    out = torch.empty(n, dtype=F64, device=x.device)
    check(_lib.lib().mrf_abs_deviation_f64(_p(x), n, float(c), _p(out), _stream()), "mrf_abs_deviation_f64")
    return out


def masked_minmax(S, zero_mask=None, box=None, y0=0):
    """min / max of S with the pixels of ``zero_mask`` and of the half-open global box
    (x0, x1, y0, y1) replaced by 0 (pipeline/build_cost_volume.py:154-169,190-191).  fp64[2] tensor."""
    _chk(S, F64, "S"); _chk(zero_mask, U8, "zero_mask")
    L = _lib.lib()
    rows, w = S.shape
    out = torch.empty(2, dtype=F64, device=S.device)
    scratch = torch.empty(L.mrf_minmax_scratch_bytes(), dtype=U8, device=S.device)
    bx = (C.c_int * 4)(*[int(b) for b in box]) if box is not None else None
    check(L.mrf_masked_minmax_f64(_p(S), _p(zero_mask), rows, w, y0, bx, _p(out), _p(scratch), _stream()),
          "mrf_masked_minmax_f64")
    return out


def robust_apply(x, kind, deriv, p1, p2=0.0):
    """utils/robust_functions.py:26-70 on a device array (x = squared residual)."""
    _chk(x, F64, "x")
    out = torch.empty_like(x)
    check(_lib.lib().mrf_robust_apply_f64(_p(x), x.numel(), RHO[kind], int(deriv), float(p1), float(p2), _p(out),
                                          _stream()), "mrf_robust_apply_f64")
    return out


# ------------------------------------------------------------------------------------ a11 / a8 / a12
def prepare_channels(rgb, want_ch64=True, want_texels=True):
    """structure_refinement/optimize_structure_single_scale.py:350-371 (grayscale data term):
    (gray*255, blur3(d/dy), blur3(d/dx)).  ``rgb`` [h,w,3] fp64 or fp32.  Returns
    (ch64 fp64 [h,w,3] or None, texels fp32 [h,w,4] or None)."""
    if not (rgb.is_cuda and rgb.is_contiguous() and rgb.dtype in (F32, F64) and rgb.dim() == 3 and rgb.shape[2] == 3):
        raise TypeError("rgb must be a contiguous CUDA [h,w,3] fp32/fp64 tensor")
    h, w = rgb.shape[:2]
    dev = rgb.device
    ch64 = torch.empty((h, w, 3), dtype=F64, device=dev) if want_ch64 else None
    tex = torch.empty((h, w, 4), dtype=F32, device=dev) if want_texels else None
    check(_lib.lib().mrf_prepare_channels(_p(rgb), int(rgb.dtype == F64), h, w, _p(ch64), _p(tex), None, _stream()),
          "mrf_prepare_channels")
    return ch64, tex


def warp_sample(img, H, xn, yn, interp="cubic", want_mask=True):
    """structure_refinement/interpolate_through_H.py:7-51 (compute_derivs=False).  ``img`` fp32
    [h,w] or [h,w,C<=4]; xn, yn fp64 of any (equal) shape.  Returns (J fp32 xn.shape(+[C]), mask fp64)."""
    _chk(img, F32, "img"); _chk(xn, F64, "xn"); _chk(yn, F64, "yn")
    h, w = img.shape[:2]
    Cn = 1 if img.dim() == 2 else img.shape[2]
    n = xn.numel()
    J = torch.empty(tuple(xn.shape) + ((Cn,) if img.dim() == 3 else ()), dtype=F32, device=img.device)
    mask = torch.empty(tuple(xn.shape), dtype=F64, device=img.device) if want_mask else None
    Hh = _mat(H) if H is not None else None
    check(_lib.lib().mrf_warp_sample(_p(img), h, w, Cn, Hh, _p(xn), _p(yn), n, INTERP[interp], _p(J), _p(mask),
                                     _stream()), "mrf_warp_sample")
    return J, mask


def _cv_params(rows, w, nb_rows, H, q, mu, B, label_scale, Ln, rho, rho_param, cw, interp, layout, i32_scale, variant, y0, frame_h,
               nb_y0, plane_stride, frame):
    p = CostVolParams()
    p.rows, p.w, p.y0, p.frame_h, p.nb_rows, p.nb_y0 = rows, w, int(y0), frame_h, nb_rows, int(nb_y0)
    p.H = _mat(H); p.q = _pt(q)
    p.mu, p.B, p.label_scale = float(mu), float(B), float(label_scale)
    p.n_labels, p.interp, p.rho, p.rho_param = Ln, INTERP[interp], RHO[rho], float(rho_param)
    p.cw = dvec(cw)
    p.layout, p.i32_scale, p.variant = LAYOUT[layout], float(i32_scale), VARIANT[variant]
    p.cost_plane_stride = int(plane_stride)
    p.state_frame = int(frame)
    return p


def cost_volume(ref_ch, nb_texels, labels, H, q, mu, B, label_scale=1.0, nonrigid=None, occluded=None,
                rho="lorentzian", rho_param=1.0, cw=(0.01, 1.0, 1.0), interp="cubic", layout="LHW_F32",
                i32_scale=1000.0, variant="staged", y0=0, frame_h=None, nb_y0=0, want_cost=True,
                want_min=True, want_argmin=True, out=None, halo_miss=None, out_ptr=None, plane_stride=0,
                state=None, frame=1, n_labels=None):
    """Per-label data cost of ONE neighbour frame (the missing extension of
    pipeline/build_cost_volume.py:209-219, defined per SURVEY.md section 8c from
    optimize_structure_single_scale.py:210-302).  ``ref_ch`` fp64 [rows,w,3] reference-frame channels
    (band starting at global row ``y0``), ``nb_texels`` fp32 [nb_rows,w,4] neighbour texels (band starting
    at ``nb_y0``), ``labels`` fp64 [L].  Returns (cost, min_cost fp32 [rows,w], argmin int32 [rows,w])."""
    _chk(ref_ch, F64, "ref_ch"); _chk(nb_texels, F32, "nb_texels"); _chk(labels, F64, "labels")
    _chk(nonrigid, U8, "nonrigid"); _chk(occluded, U8, "occluded"); _chk(state, F64, "state")
    rows, w = ref_ch.shape[:2]
    nb_rows = nb_texels.shape[0]
    frame_h = rows if frame_h is None else int(frame_h)
    Ln = labels.numel() if labels is not None else int(n_labels)
    dev = ref_ch.device
    p = _cv_params(rows, w, nb_rows, H, q, mu, B, label_scale, Ln, rho, rho_param, cw, interp, layout, i32_scale, variant, y0,
                   frame_h, nb_y0, plane_stride, frame)
    cost = out
    if out_ptr is not None:
        cost = None                      # raw (possibly peer-mapped) destination
    elif cost is None and want_cost:
        cost = (torch.empty((Ln, rows, w), dtype=F32, device=dev) if layout == "LHW_F32"
                else torch.empty((rows, w, Ln), dtype=I32, device=dev))
    mn = torch.empty((rows, w), dtype=F32, device=dev) if want_min else None
    am = torch.empty((rows, w), dtype=I32, device=dev) if want_argmin else None
    dst = C.c_void_p(int(out_ptr)) if out_ptr is not None else _p(cost)
    check(_lib.lib().mrf_cost_volume(C.byref(p), _p(ref_ch), _p(nb_texels), _p(labels), _p(nonrigid), _p(occluded),
                                     dst, _p(mn), _p(am), _p(halo_miss), _p(state), _stream()), "mrf_cost_volume")
    return cost, mn, am


def cost_volume_auto(ref_ch, nb_texels, labels, H, q, mu, B, label_scale=1.0, nonrigid=None, occluded=None,
                     cw=(0.01, 1.0, 1.0), interp="cubic", want_cost=True, halo_miss=None, state=None, frame=1, n_labels=None):
    """The cost volume of one neighbour frame with the reference's AUTO-SIGMA Lorentzian
    (utils/robust_functions.py:55-58, what generate_lorentzian() without an argument gives): per label plane,
    sigma = max(1e-6, 1.4826 * median |Iz|) over the whole plane and cost = sigma * log(1 + 0.5 Iz^2 / sigma^2).
    Whole frames only (the median is over the plane).  Three launches per label (fp64 residual, exact median,
    rho); nothing comes back to the host.  Returns (cost fp32 [L,h,w] or None, min_cost, argmin)."""
    _chk(ref_ch, F64, "ref_ch"); _chk(nb_texels, F32, "nb_texels"); _chk(labels, F64, "labels")
    _chk(nonrigid, U8, "nonrigid"); _chk(occluded, U8, "occluded"); _chk(state, F64, "state")
    L = _lib.lib()
    rows, w = ref_ch.shape[:2]
    if nb_texels.shape[0] != rows:
        raise ValueError("the auto-sigma mode needs whole frames (sigma is a median over the whole label plane)")
    Ln = labels.numel() if labels is not None else int(n_labels)
    dev = ref_ch.device
    n = rows * w
    p = _cv_params(rows, w, rows, H, q, mu, B, label_scale, Ln, "lorentzian_auto", 1.0, cw, interp, "LHW_F32", 1000.0, "gather", 0,
                   rows, 0, 0, frame)
    cost = torch.empty((Ln, rows, w), dtype=F32, device=dev) if want_cost else None
    mn = torch.empty((rows, w), dtype=F32, device=dev)
    am = torch.empty((rows, w), dtype=I32, device=dev)
    Iz = torch.empty((rows, w), dtype=F64, device=dev)
    aIz = torch.empty((rows, w), dtype=F64, device=dev)
    med = torch.empty(1, dtype=F64, device=dev)
    cnt = torch.full((1,), n, dtype=I32, device=dev)
    scratch = torch.empty(L.mrf_select_scratch_bytes(n), dtype=U8, device=dev)
    st = _stream()
    for l in range(Ln):
        check(L.mrf_cost_residual_plane(C.byref(p), _p(ref_ch), _p(nb_texels), _p(labels), _p(nonrigid), _p(occluded), l, _p(Iz),
                                        _p(aIz), _p(halo_miss), _p(state), st), "mrf_cost_residual_plane")
        check(L.mrf_select_median_dev(_p(aIz), _p(cnt), n, _p(med), _p(scratch), st), "mrf_select_median_dev")
        check(L.mrf_cost_auto_rho_plane(_p(Iz), n, _p(med), l, _p(cost[l]) if cost is not None else None, _p(mn), _p(am), st),
              "mrf_cost_auto_rho_plane")
    return cost, mn, am


def cost_volume_pair(ref_ch, nb_texels, labels, H, q, mu, B, label_scale=(1.0, 1.0), nonrigid=None, occluded=(None, None),
                     rho="lorentzian", rho_param=1.0, cw=(0.01, 1.0, 1.0), interp="cubic", layout="LHW_F32", i32_scale=1000.0,
                     variant="staged", y0=0, frame_h=None, nb_y0=0, want_cost=True, want_min=True, want_argmin=True,
                     out=(None, None), halo_miss=None, out_ptr=(None, None), plane_stride=0, state=None, n_labels=None):
    """``cost_volume`` for BOTH neighbour frames of a triplet in one launch (grid.z = frame; the tail of one
    frame's blocks overlaps the head of the other's).  Per-frame arguments are pairs indexed by frame (0 = bwd,
    1 = fwd).  Returns [(cost, min, argmin), (cost, min, argmin)]."""
    _chk(ref_ch, F64, "ref_ch"); _chk(labels, F64, "labels"); _chk(nonrigid, U8, "nonrigid"); _chk(state, F64, "state")
    for f in range(2):
        _chk(nb_texels[f], F32, "nb_texels"); _chk(occluded[f], U8, "occluded")
    rows, w = ref_ch.shape[:2]
    assert nb_texels[0].shape == nb_texels[1].shape
    nb_rows = nb_texels[0].shape[0]
    frame_h = rows if frame_h is None else int(frame_h)
    Ln = labels.numel() if labels is not None else int(n_labels)
    dev = ref_ch.device
    ps, cost, mn, am, dst = [], [], [], [], []
    for f in range(2):
        ps.append(_cv_params(rows, w, nb_rows, H[f], q[f], mu[f], B[f], label_scale[f], Ln, rho, rho_param, cw, interp, layout,
                             i32_scale, variant, y0, frame_h, nb_y0, plane_stride, f))
        c = out[f]
        if out_ptr[f] is not None:
            c = None
        elif c is None and want_cost:
            c = (torch.empty((Ln, rows, w), dtype=F32, device=dev) if layout == "LHW_F32"
                 else torch.empty((rows, w, Ln), dtype=I32, device=dev))
        cost.append(c)
        mn.append(torch.empty((rows, w), dtype=F32, device=dev) if want_min else None)
        am.append(torch.empty((rows, w), dtype=I32, device=dev) if want_argmin else None)
        dst.append(C.c_void_p(int(out_ptr[f])) if out_ptr[f] is not None else _p(c))
    check(_lib.lib().mrf_cost_volume_pair(C.byref(ps[0]), C.byref(ps[1]), _p(ref_ch), _p(nb_texels[0]), _p(nb_texels[1]),
                                          _p(labels), _p(nonrigid), _p(occluded[0]), _p(occluded[1]), dst[0], dst[1],
                                          _p(mn[0]), _p(mn[1]), _p(am[0]), _p(am[1]), _p(halo_miss), _p(state), _stream()),
          "mrf_cost_volume_pair")
    return [(cost[0], mn[0], am[0]), (cost[1], mn[1], am[1])]


# ------------------------------------------------------------------------------------ neighbours of the path
def flow_consistency(flow, backflow, threshold):
    """rigidity/occlusions.py:10-74.  flow / backflow: [[u,v] bwd, [u,v] fwd] fp32 CUDA tensors.
    Returns (occ_backward, occ_forward, occ_both) uint8."""
    for f in range(2):
        for t in (*flow[f], *backflow[f]):
            _chk(t, F32, "flow")
    h, w = flow[0][0].shape
    outs = [torch.empty((h, w), dtype=U8, device=flow[0][0].device) for _ in range(3)]
    check(_lib.lib().mrf_flow_consistency(_p(flow[0][0]), _p(flow[0][1]), _p(backflow[0][0]), _p(backflow[0][1]),
                                          _p(flow[1][0]), _p(flow[1][1]), _p(backflow[1][0]), _p(backflow[1][1]), h, w,
                                          float(threshold), _p(outs[0]), _p(outs[1]), _p(outs[2]), _stream()),
          "mrf_flow_consistency")
    return tuple(outs)


def flo_decode(file_bytes_dev, h, w):
    """utils/flow_io.py:44-46 on the device: the whole .flo file as a uint8 CUDA tensor -> (u, v) fp32 [h,w]."""
    _chk(file_bytes_dev, U8, "file")
    u = torch.empty((h, w), dtype=F32, device=file_bytes_dev.device)
    v = torch.empty((h, w), dtype=F32, device=file_bytes_dev.device)
    check(_lib.lib().mrf_flo_decode(_p(file_bytes_dev), file_bytes_dev.numel(), int(h), int(w), _p(u), _p(v), _stream()),
          "mrf_flo_decode")
    return u, v


def image_weights(I):
    """rigidity/inference.py:140-167.  I fp64 [h,w,c] (or [h,w]).  Returns (w_horizontal, w_vertical, w_ne,
    w_se) fp32 [h,w]."""
    _chk(I, F64, "I")
    if I.dim() == 2:
        I = I.unsqueeze(2)
    h, w, c = I.shape
    L = _lib.lib()
    outs = [torch.empty((h, w), dtype=F32, device=I.device) for _ in range(4)]
    scratch = torch.empty(L.mrf_image_weights_scratch_bytes(), dtype=U8, device=I.device)
    check(L.mrf_image_weights(_p(I), h, w, c, _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]), _p(scratch), _stream()),
          "mrf_image_weights")
    return tuple(outs)


# ------------------------------------------------------------------------------------ linearised data term
_T4 = [np.array([[1.0, 0, -1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), np.array([[1.0, 0, 1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
       np.array([[1.0, 0, 0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]]), np.array([[1.0, 0, 0], [0.0, 1.0, 1.0], [0.0, 0.0, 1.0]])]


def _cv_invert3x3(M):
    """cv::invert for a 3x3 fp64 matrix (closed form, OpenCV's operation order) -- what
    cv2.warpPerspective applies to its matrix argument."""
    S = np.asarray(M, dtype=np.float64)
    det = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
           + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    d = 1.0 / det
    return np.array([(S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d, (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d,
                     (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d, (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d,
                     (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d, (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d,
                     (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d, (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d,
                     (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d]).reshape(3, 3)


def derivs_through_H(I, H):
    """structure_refinement/interpolate_through_H.py:53-97.  I fp64 [h,w,C] (C <= 4) or [h,w].  Returns
    (dx, dy) fp64 of I's shape."""
    _chk(I, F64, "I")
    two_d = I.dim() == 2
    Iv = I.unsqueeze(2) if two_d else I
    h, w, Cn = Iv.shape
    H = np.asarray(H, dtype=np.float64)
    Hinv = np.linalg.inv(H)
    Ms = [Hinv.dot(T).dot(H) for T in _T4]                          # :68
    M4 = dvec(np.concatenate([M.reshape(9) for M in Ms]))
    Mi4 = dvec(np.concatenate([_cv_invert3x3(M).reshape(9) for M in Ms]))
    bh0 = min(16, h); bw0 = min(1024 // bh0, w)                     # cv::warpPerspective's block shape
    dx = torch.empty_like(Iv); dy = torch.empty_like(Iv)
    check(_lib.lib().mrf_derivs_through_H(_p(Iv), h, w, Cn, M4, Mi4, bw0, _p(dx), _p(dy), _stream()), "mrf_derivs_through_H")
    return (dx[..., 0], dy[..., 0]) if two_d else (dx, dy)


def pack_texels(ch64):
    """fp64 [h,w,3] -> fp32 texels [h,w,4]."""
    _chk(ch64, F64, "ch64")
    h, w = ch64.shape[:2]
    tex = torch.empty((h, w, 4), dtype=F32, device=ch64.device)
    check(_lib.lib().mrf_pack_texels(_p(ch64), h, w, _p(tex), _stream()), "mrf_pack_texels")
    return tex


def data_term_prepare(I1, I2, H):
    """The parts of computeDataTerm that do not depend on the structure: image derivatives through H
    (:238-242, :255-256) and the fp32 texel planes the warp samples.  Reusable across the outer iterations
    of the variational refinement."""
    _chk(I1, F64, "I1"); _chk(I2, F64, "I2")
    d2x, d2y = derivs_through_H(I2, H)                              # :238-242 -> interpolate_through_H.py:44-46
    d1x, d1y = derivs_through_H(I1, np.eye(3))                      # :255-256
    return dict(d1x=d1x, d1y=d1y, tex=pack_texels(I2), tdx=pack_texels(d2x), tdy=pack_texels(d2y))


def data_term(A, I1, I2, nonrigid, occluded, H, B, mu, q, rho="lorentzian_auto", sigma=0.0, cw=(0.01, 1.0, 1.0),
              prepared=None):
    """optimize_structure_single_scale.py:186-305 (computeDataTerm) on device tensors: A fp64 [h,w], I1 / I2
    fp64 [h,w,3] channel stacks, masks uint8.  Returns (diag, b, E, intermediates): the two images after
    their 5x5 median filters, the energy as a 1-element tensor, and dict(Iz, dphi, dr_da, sigma) -- with the
    auto-sigma Lorentzian, ``sigma`` is the device-resident median of |Iz| (sigma = max(1e-6, 1.4826 * median));
    nothing in here synchronises with the host."""
    _chk(A, F64, "A"); _chk(I1, F64, "I1"); _chk(I2, F64, "I2"); _chk(nonrigid, U8, "nonrigid"); _chk(occluded, U8, "occluded")
    L = _lib.lib()
    h, w = A.shape
    dev = A.device
    pr = prepared if prepared is not None else data_term_prepare(I1, I2, H)
    d1x, d1y, tex, tdx, tdy = pr["d1x"], pr["d1y"], pr["tex"], pr["tdx"], pr["tdy"]
    Iz = torch.empty((h, w), dtype=F64, device=dev); dphi = torch.empty_like(Iz); drda = torch.empty_like(Iz)
    check(L.mrf_data_term_residual(_p(A), _p(I1), _p(tex), _p(tdx), _p(tdy), _p(d1x), _p(d1y), _p(nonrigid), _p(occluded),
                                   h, w, _mat(H), _pt(q), float(mu), float(B), dvec(cw), _p(Iz), _p(dphi), _p(drda),
                                   _stream()), "mrf_data_term_residual")
    med = None
    if rho == "lorentzian_auto":                                    # utils/robust_functions.py:55: MAD auto-sigma,
        med = median_dev(abs_deviation(Iz.view(-1), 0.0))           # formed in the kernel from the device-resident median
        sigma = med
    diag = torch.empty_like(Iz); b = torch.empty_like(Iz); E = torch.empty(1, dtype=F64, device=dev)
    scratch = torch.empty(L.mrf_data_term_scratch_bytes(), dtype=U8, device=dev)
    check(L.mrf_data_term_system(_p(Iz), _p(dphi), _p(drda), h * w, RHO[rho], 0.0 if med is not None else float(sigma), _p(med),
                                 _p(diag), _p(b), _p(E), _p(scratch), _stream()), "mrf_data_term_system")
    dm = torch.empty_like(diag); bm = torch.empty_like(b)
    check(L.mrf_median5x5_f64(_p(diag), h, w, _p(dm), _stream()), "mrf_median5x5_f64")      # :299
    check(L.mrf_median5x5_f64(_p(b), h, w, _p(bm), _stream()), "mrf_median5x5_f64")         # :300
    return dm, bm, E, dict(Iz=Iz, dphi=dphi, dr_da=drda, sigma=sigma, diag_raw=diag, b_raw=b)


# ------------------------------------------------------------------------------------ device-resident pre-pass
class FrontBuffers:
    """Reusable HBM buffers of cost_volume_front for one frame size (keeps the step allocation-free
    so it can be captured in a CUDA graph)."""

    def __init__(self, rows, w, dev="cuda"):
        L = _lib.lib()
        n = rows * w
        self.par = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.S = [torch.empty((rows, w), dtype=F64, device=dev) for _ in range(2)]
        self.cand = torch.empty(n, dtype=F64, device=dev)
        self.count = torch.zeros(1, dtype=I32, device=dev)
        self.state = torch.zeros(_lib.STATE_WORDS, dtype=F64, device=dev)
        self.scratch = torch.empty(L.mrf_front_scratch_bytes(), dtype=U8, device=dev)


def cost_volume_front(flow, rigid_mask, zero_mask, homographies, epipoles, frame_h, buf, n_labels=51, B_fwd=1.0, fused=False):
    """pipeline/build_cost_volume.py:78-192 without leaving the device: parallax, mu, B_b, structures,
    label set; every whole-frame scalar is written to ``buf.state`` (see MRF_STATE_* in the header).
    ``fused=True`` runs the same pre-pass as one persistent cooperative kernel (identical bits; see the header for
    why it is not the default)."""
    for f in range(2):
        _chk(flow[f][0], F32, "u"); _chk(flow[f][1], F32, "v")
    _chk(rigid_mask, U8, "rigid_mask"); _chk(zero_mask, U8, "zero_mask")
    rows, w = flow[0][0].shape
    p = _lib.FrontParams()
    p.rows, p.w, p.y0, p.frame_h = rows, w, 0, int(frame_h)
    p.Hinv0 = _mat(np.linalg.inv(np.asarray(homographies[0], dtype=np.float64)))
    p.Hinv1 = _mat(np.linalg.inv(np.asarray(homographies[1], dtype=np.float64)))
    p.q0 = _pt(epipoles[0]); p.q1 = _pt(epipoles[1])
    p.B_fwd = float(B_fwd)
    p.n_labels = int(n_labels)
    p.fused = int(bool(fused))
    for f in range(2):
        q = epipoles[f]
        bx0 = max(0, min(w - 1, int(q[0] - 10))); bx1 = max(0, min(w - 1, int(q[0] + 10)))
        by0 = max(0, min(frame_h - 1, int(q[1] - 10))); by1 = max(0, min(frame_h - 1, int(q[1] + 10)))
        has = not (bx0 == w - 1 or bx1 == 0 or by0 == frame_h - 1 or by1 == 0)      # build_cost_volume.py:162-165
        p.have_box[f] = int(has)
        for k, v in enumerate((bx0, bx1, by0, by1)):
            p.box[f][k] = v
    check(_lib.lib().mrf_cost_volume_front(C.byref(p), _p(flow[0][0]), _p(flow[0][1]), _p(flow[1][0]), _p(flow[1][1]),
                                           _p(rigid_mask), _p(zero_mask), _p(buf.par[0]), _p(buf.par[1]), _p(buf.S[0]),
                                           _p(buf.S[1]), _p(buf.cand), _p(buf.count), _p(buf.state), _p(buf.scratch),
                                           _stream()), "mrf_cost_volume_front")
    return buf
