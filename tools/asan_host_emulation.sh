#!/bin/bash
# Memory-safety check of the kernels without a GPU: the host emulation of the CUDA sources
# (tests/host_emulation/) built with AddressSanitizer, driven by the GPU parity tests.
# Excluded: tests that take minutes on one host thread and the
# restart and numeric-error tests (a C++ exception is thrown, which trips ASan's __cxa_throw
# interceptor under LD_PRELOAD -- unrelated to the kernels).
set -eu
cd "$(dirname "$0")/.."
make -s -C tests/host_emulation asan
ASAN=$(/usr/bin/g++ -print-file-name=libasan.so)
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
RYUJIN_B200_LIB=$PWD/tests/host_emulation/build_asan/libryujin_b200_hostemu_asan.so \
python -m pytest tests -q -m gpu -p no:cacheprovider \
  -k "not (numeric_error or restart_and_warning or conservation_and_symmetry or config1 or vortex_3d_cycles or golden_on_gpu or verification_goldens or decks_reproduce)"
