"""Builds libtmap_b200.so (CUDA kernels + host runtime + C ABI) in-tree with nvcc for sm_100a.

kernels_bin.cu is compiled with -fmad=false: its float32/float64 operation order is part of the
bit-exactness contract with the reference (no FMA contraction).  No torch involvement: the library
is a plain C-ABI shared object.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtmap_b200.so")
OBJ = os.path.join(HERE, "_build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]
UNITS = [
    ("kernels_bin.cu", ["-fmad=false"]),
    ("kernels_classify.cu", []),
    ("voxel.cu", ["-fmad=false"]),
    ("runtime.cu", []),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "tmap.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.run(cmd, check=True)
    if force or _stale(OUT, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lpthread"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
