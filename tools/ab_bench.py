#!/usr/bin/env python
"""A/B harness for kernel experiments: run bench.py against an alternative build of libwhalecuda.

    python tools/ab_bench.py build/ab/variant.so [bench.py arguments...]

Development aid only (the package itself always loads whale.jl_b200/libwhalecuda.so).  The alternative library is
a CUDA build of the same sources with different compile-time switches (see tools/gpu_round.sh)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path = os.path.abspath(sys.argv[1])
    from whale_jl_b200 import lib as wlib
    wlib._default = wlib.Lib(path)
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")


if __name__ == "__main__":
    main()
