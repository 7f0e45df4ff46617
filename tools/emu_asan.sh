#!/bin/bash
# Memory check of the kernel logic without a GPU: the emulation build (tests/emu) compiled with AddressSanitizer + UBSan (aborting),
# driven by the emulation tests.  Catches out-of-bounds reads/writes of the packed arena, the scratch buffers and
# the dynamic shared-memory block as a whole (not between regions inside it).  compute-sanitizer on the B200 is
# the real thing; this is the check that runs in the GPU-less build container.
#   tools/emu_asan.sh [pytest -k expression]
set -eu
cd "$(dirname "$0")/.."
mkdir -p build
/usr/bin/g++ -x c++ -std=c++17 -O1 -g -ffp-contract=off -DWHALE_EMU -Itests/emu -shared -fPIC -pthread \
    -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -o build/libwhalecuda_emu_asan.so whale.jl_b200/csrc/whalecuda.cu
export WHALE_EMU_LIB=$PWD/build/libwhalecuda_emu_asan.so
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_emu.py -x -q \
    -k "${1:-not variant and not chain_tables and not mle}"
