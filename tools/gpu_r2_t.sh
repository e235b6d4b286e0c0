#!/bin/bash
# round 2, compute-sanitizer on the real device: memcheck and racecheck (shared-memory hazards of the tiled kernels,
# the block reduction of k_dii_tau_final) over a small folded / separate ERK33 run and a 3D stage-weights update.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  ( time timeout 200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
      -k "(folded and sod2d and erk33) or (stage_weights and vortex3d)" ) > gpurun_out/r2t_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/r2t_sanitizer_$tool.log | head -5
done
