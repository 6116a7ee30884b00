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


def test_bench_reference_arm_contract():
    """bench.py --impl reference (the CPU arm, the one leg that runs without a GPU) prints one JSON line with
    the contract's keys and the same `config` keys as the GPU arm uses."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--shape", "48x64", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cost_volume_pixel_labels_per_s" and d["higher_is_better"] is True
    assert sorted(d["config"]) == ["frames", "interp", "l2", "labels", "workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    src = open(os.path.join(root, "bench.py")).read()
    assert src.count('"config": bench_config(args, hw)') == 2          # both arms build it with the same function
