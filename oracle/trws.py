"""rigidity/inference.py:99-135 (infer_mrf) with the reference's own C++ solver compiled into
oracle/_ref/ by oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.  ``available()`` is False where the
extension has not been built (it travels to the GPU box as a prebuilt file)."""
import importlib.util
import os
import sysconfig

import numpy as np

from . import hotpath as hp

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "eff_trws" + sysconfig.get_config_var("EXT_SUFFIX"))
_mod = None


def available():
    return os.path.exists(_SO)


def _ext():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("eff_trws", _SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def infer_mrf(I, unaries, lambd=1.1, weights=None):
    """The reference's infer_mrf: image weights (inference.py:140-167), int32 unaries (:113), then
    Eff_TRWS(w, h, nlabels, truncation=1, neighborhood=8).solve(..., use_trws=False) (:119-133)."""
    h, w, nlabels = unaries.shape
    w_e, w_n, w_ne, w_se = weights if weights is not None else hp.get_image_weights(I)
    solver = _ext().Eff_TRWS(w, h, nlabels, truncation=1, neighborhood=8)
    labels_out = np.zeros((h, w), dtype=np.int32)
    solver.solve(hp.pack_unaries(unaries), int(lambd * 1000), labels_out, weights_horizontal=w_e, weights_vertical=w_n,
                 weights_ne=w_ne, weights_se=w_se, use_trws=False, effective_part_opt=True)
    return labels_out
