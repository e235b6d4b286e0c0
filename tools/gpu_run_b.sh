#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python tools/small_mesh_stepping.py > gpurun_out/small_mesh_stepping.log 2>&1
tail -7 gpurun_out/small_mesh_stepping.log
for m in host device graph; do
  RYUJIN_B200_STEPPING=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
done
RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_fmad.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_fmad.json 2> gpurun_out/var_fmad.err
RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_fmad.so python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "vortex_3d_cycles or config1 or single_substep or reference_verification" > gpurun_out/pytest_fmad.log 2>&1
tail -5 gpurun_out/pytest_fmad.log
