#!/usr/bin/env python
"""Compile the reference's own MRF solver (extern/eff_trws: Cython wrapper + C++ TRW-S / alpha-expansion)
from the sources where they lie under /root/reference into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The solver is the *consumer* of the path's unaries (pipeline/compute_structure.py:397 ->
rigidity/inference.py:99-135); having the real thing lets the parity tests compare the complete
7-tuple of compute_structure, rigidity_refined included.  Nothing is copied into the repository: the
Cython-generated translation unit and the shared object go to oracle/_ref/ (git-ignored); the reference's
own build (distutils setup.py, `-pg`) is not used.  Run in the authoring container:

    python oracle/build_ref.py
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/extern/eff_trws"


def build(verbose=False):
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "eff_trws" + ext)
    srcs = [os.path.join(SRC, "eff_trws.pyx"), os.path.join(SRC, "image.cpp"), os.path.join(SRC, "image.h")]
    if os.path.exists(so) and all(os.path.getmtime(s) <= os.path.getmtime(so) for s in srcs):
        return so
    cpp = os.path.join(OUT, "eff_trws_cython.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-2", "-o", cpp, srcs[0]],
                          stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
    import numpy
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-I" + SRC, "-I" + os.path.join(SRC, "eff_mlabel_ver1")]
    cmd = ["g++", "-O2", "-w", "-fPIC", "-shared", "-std=c++14"] + inc + [cpp, srcs[1], "-o", so]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
    return so


if __name__ == "__main__":
    print(build(verbose=True))
