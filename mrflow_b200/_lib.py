"""ctypes binding of libmrflow_b200.so (include/mrflow_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.  Build it
with ``python -m mrflow_b200._build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRF_LIB_PATH") or os.path.join(HERE, "libmrflow_b200.so")    # MRF_LIB_PATH: an experimental build

INTERP = {"cubic": 0, "linear": 1}
RHO = {"lorentzian": 0, "charbonnier": 1, "geman_mcclure": 2, "quadratic": 3, "lorentzian_auto": 4, "none": 5}
LAYOUT = {"LHW_F32": 0, "HWL_I32": 1}
VARIANT = {"staged": 0, "gather": 1, "staged_occ2": 16, "gather_occ2": 17, "staged_occ4": 32, "gather_occ4": 33,
           "staged_grouped": 128, "staged_occ2_grouped": 144, "staged_occ4_grouped": 160, "staged_single": 256}
SELECT_HIST_OFFSET = 64
SELECT13_BINS, SELECT13_HIST_OFFSET, SELECT13_SUCC_OFFSET, SELECT13_PASSES = 8192, 256, 24, 5


class CostVolParams(C.Structure):
    """mrf_costvol_params"""
    _fields_ = [("rows", C.c_int), ("w", C.c_int), ("y0", C.c_int), ("frame_h", C.c_int),
                ("nb_rows", C.c_int), ("nb_y0", C.c_int),
                ("H", C.c_double * 9), ("q", C.c_double * 2),
                ("mu", C.c_double), ("B", C.c_double), ("label_scale", C.c_double),
                ("n_labels", C.c_int), ("interp", C.c_int), ("rho", C.c_int),
                ("rho_param", C.c_double), ("cw", C.c_double * 3),
                ("layout", C.c_int), ("i32_scale", C.c_double), ("variant", C.c_int),
                ("state_frame", C.c_int), ("cost_plane_stride", C.c_longlong)]


class PeerSyncArg(C.Structure):
    """mrf_peer_sync"""
    _fields_ = [("ctx", C.c_void_p * 8), ("world", C.c_int), ("rank", C.c_int)]


class FrontParams(C.Structure):
    """mrf_front_params"""
    _fields_ = [("rows", C.c_int), ("w", C.c_int), ("y0", C.c_int), ("frame_h", C.c_int),
                ("Hinv0", C.c_double * 9), ("Hinv1", C.c_double * 9), ("q0", C.c_double * 2), ("q1", C.c_double * 2),
                ("B_fwd", C.c_double), ("have_box", C.c_int * 2), ("box", (C.c_int * 4) * 2), ("n_labels", C.c_int), ("fused", C.c_int)]


STATE_MU, STATE_B, STATE_MINMAX, STATE_NANFLAG, STATE_LABELS, STATE_WORDS = 0, 2, 4, 8, 16, 16 + 256
# MRF_MINRES_* of the header: the device-resident MINRES state block
MINRES_STATE_WORDS, MINRES_ITN, MINRES_ISTOP, MINRES_ANORM, MINRES_ACOND, MINRES_RNORM, MINRES_YNORM = 32, 18, 19, 20, 21, 22, 23


_P, _I, _D, _Z = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/mrflow_b200.h declares
SIGNATURES = {
    "mrf_version": (C.c_int, []),
    "mrf_last_error": (C.c_char_p, []),
    "mrf_launch_count": (C.c_longlong, []),
    "mrf_reset_launch_count": (None, []),
    "mrf_flow_to_parallax": (_I, [_P, _P, _I, _I, _I, _DP, _DP, _P, _P, _P, _P, _P]),
    "mrf_flow2parallax": (_I, [_P, _P, _I, _I, _I, _DP, _P, _P, _P, _P]),
    "mrf_reduce_scratch_bytes": (_Z, []),
    "mrf_epipole_dist_sum": (_I, [_I, _I, _I, _DP, _P, _P, _P]),
    "mrf_parallax_to_structure": (_I, [_P, _I, _I, _I, _DP, _D, _D, _P, _P]),
    "mrf_structure_to_flow": (_I, [_P, _I, _I, _I, _DP, _D, _DP, _D, _P, _P, _P, _P, _P, _P, _I, _P]),
    "mrf_full_flow": (_I, [_P, _P, _I, _I, _I, _DP, _P, _P, _P]),
    "mrf_rigidity_unaries": (_I, [_P, _P, _I, _I, _I, _DP, _D, _D, _D, _P, _P]),
    "mrf_rigidity_direction": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _DP, _DP, _D, _D, _P, _P, _P, _P]),
    "mrf_structure_candidates": (_I, [_P, _P, _P, _I, _I, _I, _DP, _DP, _D, _D, _D, _I, _P, _P, _P, _P]),
    "mrf_rigidity_structure": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _D, _D, _D, _D, _P, _P, _P]),
    "mrf_structure_candidates_dev": (_I, [_P, _P, _P, _I, _I, _I, _DP, _DP, _P, _I, _P, _P, _P, _P]),
    "mrf_abs_deviation_dev": (_I, [_P, _P, _Z, _P, _P, _P]),
    "mrf_scales_finalize": (_I, [_P, _P, _P, _I, _I, _P]),
    "mrf_parallax_to_structure_dev": (_I, [_P, _I, _I, _I, _DP, _P, _I, _P, _P]),
    "mrf_rigidity_structure_dev": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _D, _D, _P, _P, _P]),
    "mrf_select_scratch_bytes": (_Z, [_Z]),
    "mrf_select_kth_f64": (_I, [_P, _Z, _Z, _P, _P, _P]),
    "mrf_select_begin": (_I, [_Z, _P, _P]),
    "mrf_select_hist_f64": (_I, [_P, _Z, _P, _P]),
    "mrf_select_pick": (_I, [_P, _P]),
    "mrf_select_successor_f64": (_I, [_P, _Z, _P, _P]),
    "mrf_select_result": (_I, [_Z, _P, _P, _P]),
    "mrf_select_hist_prefix_f64": (_I, [_P, _Z, C.c_ulonglong, _I, _P, _P]),
    "mrf_select_successor_key_f64": (_I, [_P, _Z, C.c_ulonglong, _P, _P]),
    "mrf_peer_alloc": (_I, [_Z, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    "mrf_peer_open": (_I, [C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "mrf_peer_close": (_I, [_P]),
    "mrf_peer_free": (_I, [_P]),
    "mrf_download_rows": (_I, [_P, _P, _Z, _I, _I, _I, _I, _I, _P]),
    "mrf_mailbox_bytes": (_Z, [_I]),
    "mrf_mailbox_allreduce": (_I, [_P, C.POINTER(C.c_void_p), _I, _I, C.c_longlong, _I, _P, _I, _P]),
    "mrf_abs_deviation_f64": (_I, [_P, _Z, _D, _P, _P]),
    "mrf_minmax_scratch_bytes": (_Z, []),
    "mrf_masked_minmax_f64": (_I, [_P, _P, _I, _I, _I, _IP, _P, _P, _P]),
    "mrf_prepare_channels": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "mrf_warp_sample": (_I, [_P, _I, _I, _I, _DP, _P, _P, _Z, _I, _P, _P, _P]),
    "mrf_cost_volume": (_I, [C.POINTER(CostVolParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_cost_volume_pair": (_I, [C.POINTER(CostVolParams), C.POINTER(CostVolParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P, _P, _P, _P]),
    "mrf_cost_residual_plane": (_I, [C.POINTER(CostVolParams), _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "mrf_cost_auto_rho_plane": (_I, [_P, _Z, _P, _I, _P, _P, _P, _P]),
    "mrf_front_scratch_bytes": (_Z, []),
    "mrf_cost_volume_front": (_I, [C.POINTER(FrontParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_front_mu_whole": (_I, [C.POINTER(FrontParams), _P, _P, _P]),
    "mrf_front_parallax_band": (_I, [C.POINTER(FrontParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_front_candidates_band": (_I, [C.POINTER(FrontParams), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_front_structures_band": (_I, [C.POINTER(FrontParams), _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "mrf_peer_sync_bytes": (_Z, []),
    "mrf_peer_sync_begin": (_I, [_P, _P]),
    "mrf_peer_push_hist": (_I, [_P, _P, _I, _P]),
    "mrf_select13_begin": (_I, [_P, _P]),
    "mrf_select13_hist_dev": (_I, [_P, _P, _Z, _I, _P, _P, _P]),
    "mrf_select13_successor_dev": (_I, [_P, _P, _Z, _P, _P, _P]),
    "mrf_select13_median": (_I, [_P, _P, _I, _P, _P]),
    "mrf_select_median_dev": (_I, [_P, _P, _Z, _P, _P, _P]),
    "mrf_label_range_dev": (_I, [_P, _I, _P, _I, _P, _P]),
    "mrf_flow_consistency": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _D, _P, _P, _P, _P]),
    "mrf_flo_decode": (_I, [_P, _Z, _I, _I, _P, _P, _P]),
    "mrf_image_weights_scratch_bytes": (_Z, []),
    "mrf_image_weights": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "mrf_derivs_through_H": (_I, [_P, _I, _I, _I, _DP, _DP, _I, _P, _P, _P]),
    "mrf_pack_texels": (_I, [_P, _I, _I, _P, _P]),
    "mrf_data_term_residual": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _DP, _DP, _D, _D, _DP, _P, _P, _P, _P]),
    "mrf_data_term_scratch_bytes": (_Z, []),
    "mrf_data_term_system": (_I, [_P, _P, _P, _Z, _I, _D, _P, _P, _P, _P, _P, _P]),
    "mrf_median5x5_f64": (_I, [_P, _I, _I, _P, _P]),
    "mrf_gray32": (_I, [_P, _I, _Z, _P, _P]),
    "mrf_edge_weights_local": (_I, [_P, _P, _I, _I, _I, _D, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_var_scratch_bytes": (_Z, []),
    "mrf_var_smooth_args": (_I, [_P, _P, _P, _I, _I, _D, _P, _P, _P, _P, _P]),
    "mrf_var_smooth_weights": (_I, [_P, _P, _P, _P, _Z, _D, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrf_var_data_merge": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _D, _D, _D, _P, _P, _P, _P, _P, _P]),
    "mrf_var_consistency_weights": (_I, [_P, _Z, _D, _D, _P, _P, _P, _P, _P]),
    "mrf_var_apply": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _I, _I, _P, _P]),
    "mrf_var_rhs": (_I, [_P, _P, _P, _P, _P, _Z, _D, _P, _P]),
    "mrf_vec_scale": (_I, [_P, _Z, _D, _P, _P]),
    "mrf_minres_scratch_bytes": (_Z, [_I, _I]),
    "mrf_minres_init": (_I, [_P, _Z, _P, _P, _P, _P, _P, _P]),
    "mrf_minres_iterate": (_I, [_P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _D,
                                C.c_longlong, _P]),
    "mrf_var_clip_update": (_I, [_P, _P, _I, _I, _DP, _D, _D, _D, _P, _P]),
    "mrf_median7x7_f64": (_I, [_P, _I, _I, _P, _P]),
    "mrf_var_advance": (_I, [_P, _P, _Z, _P]),
    "mrf_merge_structures": (_I, [_P, _P, _P, _P, _I, _I, _D, _D, _I, _P, _P, _P, _P]),
    "mrf_erode3x3_u8": (_I, [_P, _I, _I, _P, _P]),
    "mrf_robust_apply_f64": (_I, [_P, _Z, _I, _I, _D, _D, _P, _P]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmrflow_b200.so is not built (%s); run `python -m mrflow_b200._build`. "
                               "mrflow_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class MrfError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().mrf_last_error()
        raise MrfError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


def dvec(values):
    """host double array argument"""
    values = [float(v) for v in values]
    return (C.c_double * len(values))(*values)
