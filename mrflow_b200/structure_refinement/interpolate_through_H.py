"""``structure_refinement/interpolate_through_H.py`` of the reference on the B200 (sampling part).

``interpolate_through_H(I, H, xn, yn, compute_derivs, dIdx, dIdy)`` and
``compute_derivs_through_H(I, H)`` keep the reference signatures (:7, :53).
"""
import numpy as np

from .. import device
from .._host import down, require_cuda, up


def _sample(I32, H, xn_d, yn_d, want_mask):
    J, m = device.warp_sample(I32, H, xn_d, yn_d, interp="cubic", want_mask=want_mask)
    return J, m


def interpolate_through_H(I, H, xn, yn, compute_derivs=True, dIdx=None, dIdy=None):
    """J(x,y) = I(H(xn(x,y), yn(x,y))) by cv2.remap-exact bicubic sampling (interpolation.py:76-80)
    and the inside-frame mask (:39-40).  Returns ``(J, mask)`` or ``(J, mask, dJdx, dJdy)``."""
    require_cuda()
    if compute_derivs and (dIdx is None or dIdy is None):
        dIdx, dIdy = compute_derivs_through_H(I, H)                    # :45-46
    xn_d, yn_d = up(xn, np.float64), up(yn, np.float64)
    I32 = up(I, np.float32)                                 # interpolation.py:76  I.astype('float32')
    Hh = np.asarray(H, dtype=np.float64)
    J, m = _sample(I32, Hh, xn_d, yn_d, True)
    if not compute_derivs:
        return down(J), down(m)
    dJdy, _ = _sample(up(dIdy, np.float32), Hh, xn_d, yn_d, False)     # :47
    dJdx, _ = _sample(up(dIdx, np.float32), Hh, xn_d, yn_d, False)     # :48
    return down(J), down(m), down(dJdx), down(dJdy)


def compute_derivs_through_H(I, H):
    """structure_refinement/interpolate_through_H.py:53-97 -- dJ/dx, dJ/dy of J(x,y) = I(H(x,y)): four
    bilinear cv2.warpPerspective-exact warps and fp64 finite-difference scalings.  Returns fp64 arrays of
    I's shape."""
    require_cuda()
    dx, dy = device.derivs_through_H(up(I, np.float64), H)
    return down(dx), down(dy)
