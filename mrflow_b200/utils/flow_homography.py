"""``utils/flow_homography.py`` of the reference on the B200: same names, arguments and returns.

The reference adds / removes a homography from a dense flow field with cv2.perspectiveTransform on
fp32 points; the kernels reproduce OpenCV's fp64-then-round arithmetic bit for bit.
"""
import numpy as np

from .. import device
from .._host import down, require_cuda, up


def _grid_offsets(x, y):
    """The kernels use the integer pixel grid x = column, y = row (what every caller in the
    reference passes: np.mgrid / np.meshgrid).  Anything else is rejected rather than mis-computed."""
    x = np.asarray(x); y = np.asarray(y)
    h, w = x.shape
    if not (x[0, 0] == 0 and y[0, 0] == 0 and x[0, -1] == w - 1 and y[-1, 0] == h - 1
            and np.array_equal(x[0], np.arange(w)) and np.array_equal(y[:, 0], np.arange(h))):
        raise ValueError("x, y must be the integer pixel grid (np.mgrid[:h, :w])")


def get_residual_flow(x, y, u, v, H):
    """utils/flow_homography.py:4-16 -- remove the homography ``H`` from the flow ``(u, v)``."""
    require_cuda()
    _grid_offsets(x, y)
    Hinv = np.linalg.inv(np.asarray(H, dtype=np.float64))
    # (x + u) is formed in the flow's own precision promoted to fp64 (int64 + fp32/fp64 -> fp64, :9-10), rounded to
    # fp32 ONCE (:11), sent through inv(H) and reduced by the fp32 grid (:15-16) -- the arithmetic of get_full_flow
    # with inv(H), so fp64 flows are not rounded twice
    ur, vr = device.full_flow(up(u, np.float64), up(v, np.float64), Hinv)
    return down(ur), down(vr)


def get_full_flow(u_res, v_res, H, x=None, y=None, shape=None):
    """utils/flow_homography.py:20-37 -- add the homography back to a residual flow."""
    require_cuda()
    if (x is None) != (y is None):
        raise ValueError("give both x and y or neither")
    if x is not None:
        _grid_offsets(x, y)
    # x + u_res is int64 + fp32/fp64 -> fp64 in numpy (:32-33), then .astype('float32')
    u, v = device.full_flow(up(u_res, np.float64), up(v_res, np.float64), H)
    return down(u), down(v)


def homography2flow(x, y, H):
    """utils/flow_homography.py:41-48 -- the flow field a homography induces."""
    x = np.asarray(x)
    z = np.zeros(x.shape, np.float64)
    return get_full_flow(z, z, H, x=x, y=y)
