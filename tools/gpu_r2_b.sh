#!/bin/bash
# round 2, GPU session B: the slice-tiled kernels (bulk-copied tiles, kernels_tile.cuh) on real
# hardware -- parity first (RYUJIN_B200_TILE=7: steps 2, 5, 6 tiled), then per-kernel times for every
# tile mask and for compile-time flavours; also the two-library parity test and test_pow on the device.
set -u
mkdir -p gpurun_out
( time RYUJIN_B200_TILE=7 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shallow_water.py -q -m gpu -x \
    -k "single_substep or stage_weights or config3 or restart or limiter_iterations or 1d_and_dynamic or stepping_modes or sw_ or config1" ) > gpurun_out/r2b_pytest_tile7.log 2>&1
echo "pytest tile7 rc=$?"; tail -3 gpurun_out/r2b_pytest_tile7.log
( time timeout 900 python -m pytest tests/test_pow.py tests/test_gpu_second_pow_form.py -q -m gpu -x ) > gpurun_out/r2b_pytest_pow.log 2>&1
echo "pytest pow rc=$?"; tail -3 gpurun_out/r2b_pytest_pow.log
VARIANT_NAMES="base_tile0 base_tile1 base_tile2 base_tile4 base_tile7 step2mb6_tile4 step2mb5_tile4 tilew8_tile7 tilemb4_tile3 nopf_tile0 low_u2_tile0 low_u4_tile0 low_mb5_tile0 pij_mb6_tile0 pij_u2_tile0" \
  bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2b_variants.txt
