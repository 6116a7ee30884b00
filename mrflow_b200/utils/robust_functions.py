"""``utils/robust_functions.py`` of the reference on the B200: generators returning ``(f, df)``
closures over the squared residual, fixed-parameter and MAD auto-sigma forms.

The closures accept numpy arrays (returned as numpy, like the reference's) or CUDA fp64 tensors
(returned as CUDA tensors, no host round trip).
"""
import numpy as np
import torch

from .. import device
from .._host import require_cuda


def _run(x, kind, deriv, p1, p2=0.0):
    require_cuda()
    if isinstance(x, torch.Tensor):
        return device.robust_apply(x.contiguous(), kind, deriv, p1, p2)
    a = np.asarray(x, dtype=np.float64)
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return device.robust_apply(t, kind, deriv, p1, p2).cpu().numpy().reshape(a.shape)


def _auto_sigma(x):
    """utils/robust_functions.py:23-29: max(1e-6, 1.4826 * MAD(c_[sqrt(x), -sqrt(x)])).  The
    symmetrised sample has median 0 and its absolute values are sqrt(x) twice over, so the MAD is
    the median of sqrt(x)."""
    require_cuda()
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float64))).cuda()
    s = device.robust_apply(t.contiguous().view(-1), "none", 0, 0.0)
    return max(1e-6, 1.4826 * device.median(s))


def _tag(f, df, kind, param):
    """lets device-side consumers (computeDataTerm) recognise the pair without calling it"""
    f.mrf_kind = df.mrf_kind = (kind, param)
    return f, df


def generate_charbonnier(eps=0, a=0.5):
    """utils/robust_functions.py:26-43"""
    if eps == 0:
        f = lambda x: _run(x, "charbonnier", 0, _auto_sigma(x), a)
        df = lambda x: _run(x, "charbonnier", 1, _auto_sigma(x), a)
        return f, df
    f = lambda x: _run(x, "charbonnier", 0, eps, a)
    df = lambda x: _run(x, "charbonnier", 1, eps, a)
    return _tag(f, df, "charbonnier", eps) if a == 0.5 else (f, df)


def generate_quadratic():
    """utils/robust_functions.py:45-48"""
    return _tag(lambda x: _run(x, "quadratic", 0, 0.0), lambda x: _run(x, "quadratic", 1, 0.0), "quadratic", 0.0)


def generate_lorentzian(sigma=0):
    """utils/robust_functions.py:50-60"""
    if sigma > 0:
        f = lambda x: _run(x, "lorentzian", 0, sigma)
        df = lambda x: _run(x, "lorentzian", 1, sigma)
        return _tag(f, df, "lorentzian", sigma)
    f = lambda x: _run(x, "lorentzian_auto", 0, _auto_sigma(x))
    df = lambda x: _run(x, "lorentzian_auto", 1, _auto_sigma(x))
    return _tag(f, df, "lorentzian_auto", 0.0)


def generate_geman_mcclure(sigma=0):
    """utils/robust_functions.py:62-70"""
    def pair(x):
        s = sigma if sigma > 0 else _auto_sigma(x)
        return s, s ** 3
    f = lambda x: _run(x, "geman_mcclure", 0, *pair(x))
    df = lambda x: _run(x, "geman_mcclure", 1, *pair(x))
    return _tag(f, df, "geman_mcclure", sigma) if sigma > 0 else (f, df)
