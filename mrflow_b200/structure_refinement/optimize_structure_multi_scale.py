"""``structure_refinement/optimize_structure_multi_scale.py`` of the reference: the coarse-to-fine driver of
the variational refinement (:31-143).  With the reference's defaults (parameters.py:43-45: one pyramid level
at scale 1) it is a single call of ``optimizeStructureSingleScale`` at full resolution and everything runs
on the GPU.  For other pyramid settings the per-level resampling of images, structures and masks is done on
the host with the same OpenCV calls as the reference (:58-95) -- resampling is not part of the hot path --
and each level's solve runs on the GPU."""
import sys

import numpy as np

from .._host import require_cuda
from .._lib import MrfError
from . import optimize_structure_single_scale


def _resize(a, s, interpolation=None, **kw):
    import cv2
    if s == 1.0 and not kw:
        return np.asarray(a)            # cv::resize only copies when the size does not change; the solver does not write its inputs
    if interpolation is None:
        return cv2.resize(a, dsize=None, fx=s, fy=s, **kw)
    return cv2.resize(a, dsize=None, fx=s, fy=s, interpolation=interpolation, **kw)


def optimizeStructureMultiScale(A_init, A_bwd, A_fwd, occlusions_bwd, occlusions_fwd, rigidity, images, H_ar, B_ar, q_ar,
                                mu_ar, params):
    """optimize_structure_multi_scale.py:31-143.  Returns ``(A, energies)``."""
    import cv2
    require_cuda()
    scale_factor = params.variational_scale_factor
    scales = scale_factor ** np.arange(params.variational_last_scale, params.variational_N_scales)    # :52
    energies = []
    A_scale_factor = 1.0
    A_scaled = _resize(A_init, float(scales[-1]), cv2.INTER_AREA)                                      # :59
    for s in reversed(scales):
        s = float(s)
        images_scaled = [_resize(I, s, cv2.INTER_AREA) for I in images]                                # :64
        A_fwd_scaled = _resize(A_fwd, s, cv2.INTER_AREA)
        A_bwd_scaled = _resize(A_bwd, s, cv2.INTER_AREA)
        if A_scaled.shape[:2] != A_fwd_scaled.shape[:2]:                                               # :70
            A_scaled = cv2.resize(A_scaled, dsize=(A_fwd_scaled.shape[1], A_fwd_scaled.shape[0]), interpolation=cv2.INTER_CUBIC)
        occ_f = _resize(np.asarray(occlusions_fwd).astype("float32"), s) > 0                           # :74-75
        occ_b = _resize(np.asarray(occlusions_bwd).astype("float32"), s) > 0
        both = occ_f * occ_b                                                                           # :80-83
        occ_f[both] = 0
        occ_b[both] = 0
        rigidity_scaled = _resize(np.asarray(rigidity).astype("float32"), s) == 1                      # :88
        B_s = [B / s for B in B_ar]                                                                    # :96-98
        mu_s = [mu * s for mu in mu_ar]
        q_s = [np.asarray(q) * s for q in q_ar]
        scaling = np.array([[1.0, 1.0, s], [1.0, 1.0, s], [1.0 / s, 1.0 / s, 1.0]])                    # :101-104
        H_s = [np.asarray(H) * scaling for H in H_ar]
        try:
            A_scaled, energies_cur = optimize_structure_single_scale.optimizeStructureSingleScale(
                A_scaled, A_bwd_scaled, A_fwd_scaled, occ_b, occ_f, both, rigidity_scaled, images_scaled, H_s, B_s, q_s, mu_s,
                params.variational_lambda_1storder, params.variational_lambda_2ndorder, params.variational_lambda_consistency,
                A_scale_factor=A_scale_factor, N_outer=params.variational_N_outer, N_inner=params.variational_N_inner,
                params=params)
            energies += energies_cur
        except (MrfError, RuntimeError, NotImplementedError):
            raise                                          # a missing / failing GPU path is never papered over
        except Exception as e:                             # noqa: BLE001 -- :129-134: the reference warns and carries on
            sys.stderr.write("\n\n (EE) Warning in {} \n\n{!r}\n".format(getattr(params, "tempdir", "."), e))
            sys.stderr.flush()
    if A_scaled.shape[:2] != np.asarray(A_init).shape[:2]:                                            # :140-141
        A_scaled = cv2.resize(A_scaled, dsize=(A_init.shape[1], A_init.shape[0]), interpolation=cv2.INTER_LANCZOS4)
    return A_scaled, energies
