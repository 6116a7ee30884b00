"""Build libmrflow_b200.so (sm_100a only) in-tree with nvcc.

``-fmad=false`` is part of the contract, not a tuning flag: the kernels reproduce the reference's
numpy / OpenCV arithmetic operation by operation, and a contracted multiply-add would change the
rounding (see csrc/common.cuh).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrflow_b200.so")
SOURCES = ["geometry.cu", "sampling.cu", "costvol.cu", "select.cu", "robust.cu", "adjacent.cu", "dataterm.cu", "variational.cu", "peercoll.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), lib=None, objdir_name="build"):
    """Compile every CUDA translation unit and link the shared library.  Returns its path.  ``extra`` (nvcc
    flags), ``lib`` and ``objdir_name`` build an experimental copy beside the product library
    (tools/bench_kernel.py --lib)."""
    global LIB
    lib_path = lib or LIB
    objdir = os.path.join(HERE, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "mrflow_b200.h"))
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(lib_path, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
