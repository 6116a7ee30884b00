"""``pipeline/structure2flow.py`` of the reference on the B200: structure -> parallax -> residual
flow -> full flow in one kernel, same names / arguments / returns."""
import numpy as np

from .. import device
from .._host import down, flag, require_cuda, up


def structure_to_flow(structure, q, mu, H, b, params=None):
    """pipeline/structure2flow.py:14-38.  Returns (u, v) fp32."""
    require_cuda()
    u, v = device.structure_to_flow(up(structure, np.float64), q, mu, H, b)
    return down(u), down(v)


def combine_flow_simple(structure, flow_init, rigidity, q_ar, mu_ar, H_ar, b_ar, images=None, params=None):
    """pipeline/structure2flow.py:57-82: forward-frame (index 1) structure flow where rigid, initial
    flow where ``rigidity == 0`` (:72-73), pasted inside the kernel.  Returns (rigidity, u, v)."""
    require_cuda()
    rigidity = np.asarray(rigidity)
    u, v = device.structure_to_flow(up(structure, np.float64), q_ar[1], mu_ar[1], H_ar[1], b_ar[1],
                                    nonrigid=flag(rigidity == 0), init_u=up(flow_init[0], np.float32),
                                    init_v=up(flow_init[1], np.float32))
    return rigidity, down(u), down(v)


def combine_flow(structure, flow_init, rigidity, q_ar, mu_ar, H_ar, b_ar, images=None, params=None, method="simple"):
    """pipeline/structure2flow.py:41-54 (the reference ignores ``method`` and always uses 'simple')."""
    return combine_flow_simple(structure, flow_init, rigidity, q_ar, mu_ar, H_ar, b_ar, images, params)
