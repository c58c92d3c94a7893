"""Build libdecaes_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdecaes_cuda.so")
SOURCES = ["decaes_cuda.cu"]
HEADERS = ["common.cuh", "nnls.cuh", "gram.cuh", "legacy.cuh", "voxel.cuh", os.path.join("..", "..", "include", "decaes_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # only the explicit fma() calls fuse: the scalar control code must follow the reference's
    # unfused arithmetic ("Arithmetic model")
    "--fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-cudart", "static",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=(), out=None):
    if out is None and not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and libdecaes_cuda.so has not been built")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    args = sys.argv[1:]
    out = None
    if "--out" in args:  # e.g. --out /tmp/libdecaes_prof.so -DDECAES_PROFILE
        i = args.index("--out")
        out = args[i + 1]
        del args[i:i + 2]
    build(force=True, verbose=True, extra=args, out=out)
