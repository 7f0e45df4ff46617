#!/bin/bash
# Build an experiment variant of libwhalecuda (default k_dp launch shape only: fast) into build/ab/<name>.so
#   tools/build_variant.sh odd_stride -DWHALE_ODD_STRIDE
#   tools/build_variant.sh tab_proj   -DWHALE_TAB_PROJ
#   tools/build_variant.sh slice_v1   -DWHALE_SLICE_V1
# then A/B on the GPU with  tools/gpu_quick.sh build/ab/<name>.so ...   (tools/ab_bench.py runs bench.py against it).
set -eu
name=$1; shift
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DWHALE_DEV_BUILD "$@" \
    -o build/ab/$name.so whale.jl_b200/csrc/whalecuda.cu
echo build/ab/$name.so
