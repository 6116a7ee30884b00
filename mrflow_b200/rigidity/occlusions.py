"""``rigidity/occlusions.py`` of the reference on the B200."""
import numpy as np

from .. import device
from .._host import require_cuda, up


def computeOcclusionsFromConsistency(flow, backflow, threshold=7.0):
    """rigidity/occlusions.py:10-74 -- same arguments, returns ``(occ_backward, occ_forward, occ_both)``
    as boolean numpy arrays."""
    require_cuda()
    fl = [[up(flow[f][0], np.float32), up(flow[f][1], np.float32)] for f in range(2)]
    bf = [[up(backflow[f][0], np.float32), up(backflow[f][1], np.float32)] for f in range(2)]
    ob, of, both = device.flow_consistency(fl, bf, threshold)
    return ob.cpu().numpy() > 0, of.cpu().numpy() > 0, both.cpu().numpy() > 0
