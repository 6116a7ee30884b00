"""The C-ABI library builds, loads, and exports every symbol include/mrflow_b200.h declares (no
compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mrflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrf_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mrflow_b200 import _build, _lib
    path = _build.build()
    assert os.path.exists(path)
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(path)
    for n in names:
        assert hasattr(raw, n), "libmrflow_b200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    assert L.mrf_version() >= 100
    assert L.mrf_reduce_scratch_bytes() > 0 and L.mrf_select_scratch_bytes(10) > 0
    assert ctypes.sizeof(_lib.CostVolParams) == 216      # sizeof(mrf_costvol_params), checked with gcc
    assert ctypes.sizeof(_lib.FrontParams) == 248        # sizeof(mrf_front_params)


def test_bad_arguments_are_rejected_without_a_gpu():
    from mrflow_b200 import _lib
    L = _lib.lib()
    rc = L.mrf_parallax_to_structure(None, 4, 4, 0, _lib.dvec([0, 0]), 1.0, 1.0, None, None)
    assert rc == -1 and b"bad argument" in L.mrf_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mrflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
