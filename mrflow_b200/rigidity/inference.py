"""The device front-end of ``rigidity/inference.py``: everything ``infer_mrf`` (:99-135) computes before
it hands over to the C++ solver in extern/eff_trws -- the contrast-sensitive edge weights
(``get_image_weights``, :140-167) and the int32 unaries (:113).  The solver itself (Kovtun / alpha-expansion /
TRW-S, sequential integer max-flow) is the unchanged consumer."""
import numpy as np

from .. import device
from .._host import down, require_cuda, up


def get_image_weights(I):
    """rigidity/inference.py:140-167.  Returns (w_horizontal, w_vertical, w_ne, w_se) fp32 numpy arrays."""
    require_cuda()
    return tuple(down(t) for t in device.image_weights(up(I, np.float64)))


def pack_unaries(unaries):
    """rigidity/inference.py:113: ((-unaries) * 1000).astype('int32'), C-contiguous [h,w,nlabels]."""
    return ((-np.asarray(unaries)) * 1000).astype("int32").copy("C")


def infer_mrf(I, unaries, lambd=1.1):
    """rigidity/inference.py:99-135 with the weights computed on the GPU; needs the reference's
    extern/eff_trws extension (build_trws.sh) importable as ``extern.eff_trws.eff_trws``."""
    from extern.eff_trws import eff_trws           # the reference's own solver
    h, w, nlabels = unaries.shape
    w_e, w_n, w_ne, w_se = get_image_weights(I)
    solver = eff_trws.Eff_TRWS(w, h, nlabels, truncation=1, neighborhood=8)
    labels_out = np.zeros((h, w), dtype=np.int32)
    solver.solve(pack_unaries(unaries), int(lambd * 1000), labels_out, weights_horizontal=w_e, weights_vertical=w_n,
                 weights_ne=w_ne, weights_se=w_se, use_trws=False, effective_part_opt=True)
    return labels_out
