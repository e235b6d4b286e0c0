#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_fmad.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_fmad.json 2> gpurun_out/var_fmad.err
RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_fmad.so python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "vortex_3d_cycles or config1 or single_substep or reference_verification or forward_facing" > gpurun_out/pytest_fmad.log 2>&1
grep -E "rel Linf|gpu_port|passed|failed|^E  " gpurun_out/pytest_fmad.log | head -30
