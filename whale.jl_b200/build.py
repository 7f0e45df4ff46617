"""Build libwhalecuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "whalecuda.cu")
OUT = os.path.join(HERE, "libwhalecuda.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(HERE, "..", "include", "whalecuda.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    # WHALE_DEV_BUILD=1: only the default launch shape of k_dp / k_dp_rev (a quarter of the compile time; experiments)
    dev = ["-DWHALE_DEV_BUILD"] if os.environ.get("WHALE_DEV_BUILD") else []
    cmd = [nvcc, *NVCC_FLAGS, *dev, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
