"""In-tree build of libualm.so (CUDA kernels + C ABI + host tools) for sm_100a with nvcc.

Usage: python -m uneven_planner_b200.build [--force]
The .so is written next to the sources (uneven_planner_b200/libualm.so): it is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libualm.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["ualm_api.cu"]
CXX_SOURCES = ["host_tools.cpp"]
HEADERS = ["ualm_kernels.cuh", "map_cell.h", "map_prep.h", "../../include/ualm_detmath.h"]



def _flags():
    # -fmad=false: no FMA contraction, so every double result is bit-identical to the CPU oracle's
    # (gcc -ffp-contract=off); fp64 division and sqrt are IEEE by default (no fast-math anywhere)
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
            "-Xcompiler", "-fPIC,-O3,-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
            "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    for name in CU_SOURCES + CXX_SOURCES + HEADERS + ["../../include/ualm.h"]:
        p = os.path.join(CSRC, name)
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = OUT + ".stamp"
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES + CXX_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [NVCC] + _flags() + ["-shared", "-o", OUT] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libualm.so")
    if verbose:
        print(log)
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", OUT)
