#!/bin/bash
# Memory-safety check of the kernels without a GPU: the host emulation of the CUDA sources
# (tests/host_emulation/) built with AddressSanitizer, driven by the GPU parity tests.
# A bounded selection (single updates of every equation / dimension / boundary id, stage weights, the folded
# scalar kernels): the long-running cycles add no new kernel code paths.  Not run: the streaming-transfer test and
# the restart and numeric-error tests (a C++ exception is thrown, which trips ASan's __cxa_throw interceptor
# under LD_PRELOAD -- unrelated to the kernels).
set -eu
cd "$(dirname "$0")/.."
make -s -C tests/host_emulation asan
ASAN=$(/usr/bin/g++ -print-file-name=libasan.so)
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
RYUJIN_B200_LIB=$PWD/tests/host_emulation/build_asan/libryujin_b200_hostemu_asan.so \
timeout ${ASAN_TIMEOUT:-1500} python -m pytest tests/test_gpu_parity.py tests/test_gpu_shallow_water.py tests/test_gpu_euler_aeos.py -q -m gpu -p no:cacheprovider -n 4 \
  -k "single_substep or stage_weights or folded or interpolate or device_layout or euler_1d or (shallow and (single or substep or bathymetry or friction)) or (euler_aeos and single)"
