#!/bin/bash
# round 2, GPU session F: concurrency experiments on one GPU -- step 2 as two concurrent kernels
# (RYUJIN_B200_STEP2_CONCURRENT = resident k_dij blocks per SM), steps 4/5 pipelined over two halves
# (RYUJIN_B200_PIPELINE), parity of both on hardware; k_dij occupancy; strong-scaling base (320^3, N = 1).
set -u
mkdir -p gpurun_out
( time RYUJIN_B200_PIPELINE=1 RYUJIN_B200_STEP2_CONCURRENT=6 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -x \
    -k "single_substep or stage_weights or config3 or restart or stepping_modes or config1 or config2" ) > gpurun_out/r2f_pytest_concurrent.log 2>&1
echo "pytest concurrent rc=$?"; tail -3 gpurun_out/r2f_pytest_concurrent.log
one() { # one <tag> <env...>
  tag=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2f_$tag.json')); print('%-28s %7.1f MDoF/s  %.2f ms/cycle  e2e %.1f' % ('$tag', d['value'], d['ms_per_step'], d['e2e']['value']))"
}
one base X=0
one step2c4 RYUJIN_B200_STEP2_CONCURRENT=4
one step2c6 RYUJIN_B200_STEP2_CONCURRENT=6
one step2c8 RYUJIN_B200_STEP2_CONCURRENT=8
one step2c10 RYUJIN_B200_STEP2_CONCURRENT=10
one pipeline RYUJIN_B200_PIPELINE=1
one pipeline_step2c6 RYUJIN_B200_PIPELINE=1 RYUJIN_B200_STEP2_CONCURRENT=6
one pipeline_step2c8 RYUJIN_B200_PIPELINE=1 RYUJIN_B200_STEP2_CONCURRENT=8
VARIANT_NAMES="dij_mb14 dij_mb16" bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2f_variants.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scaling strong --nodes 320 > gpurun_out/r2f_strong320_n1.json 2> gpurun_out/r2f_strong320_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_strong320_n1.json')); print('strong320 n1 %7.1f MDoF/s  %.2f ms/cycle  setup %.1f s' % (d['value'], d['ms_per_step'], d['run_info']['setup_seconds']))"
