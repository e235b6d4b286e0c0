#!/bin/bash
# round 2, GPU session H: edge cache (the symmetric midpoint entropy of the Euler limiter evaluated once
# per edge by the edge-parallel kernel instead of twice by step 4) in its two forms, with occupancy
# flavours of step 4; k_dij re-reading the row state per solve; parity of the edge-cache builds on hardware.
set -u
mkdir -p gpurun_out
VARIANT_NAMES="base ec1 ec2 ec2_low5 ec2_low6 ec2_low5_u2 ec2_ec12 ec1_low5 reload ec1_reload base" bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2h_variants.txt
for v in ec1 ec2; do
( time RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -x \
    -k "single_substep or stage_weights or config3 or restart or stepping_modes or config1 or config2 or streaming" ) > gpurun_out/r2h_pytest_$v.log 2>&1
echo "pytest $v rc=$?"; tail -3 gpurun_out/r2h_pytest_$v.log
done
