#!/bin/bash
# round 2, GPU session D: the whole -m gpu suite (new: unstructured stencils, C++ shim on the device,
# both pow forms), occupancy flavours on the default tile mask, ncu --set full of the sweep kernels
# (exported as CSV; the .ncu-rep itself is too large to bring back).
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x -s ) > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log
bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2d_variants.txt
timeout 1200 ncu --set full --clock-control none \
    -k regex:'k_dij|k_alpha|k_low_order|k_pij_lij|k_high_order|k_dii_tau' -c 14 -f -o /tmp/r2d_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r2d_full.ncu-rep --page raw --csv > gpurun_out/r2d_full_raw.csv 2> /dev/null
ncu -i /tmp/r2d_full.ncu-rep --page details --csv > gpurun_out/r2d_full_details.csv 2> /dev/null
ls -la /tmp/r2d_full.ncu-rep gpurun_out/r2d_full*; du -sh gpurun_out
