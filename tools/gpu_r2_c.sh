#!/bin/bash
# round 2, GPU session C: default bench with the new CPU arm + in-bench parity check, alpha / d_ij /
# tile occupancy flavours, then one ncu --set full capture of the sweep kernels of the default build.
set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_default.json')); print(d['value'], d['e2e']['value'], d['e2e']['result_check'], d['roofline']['frac'], d['roofline']['whole_update']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "reference rc=$?"; tail -4 gpurun_out/r2c_bench_reference.err
bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2c_variants.txt
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_dij|k_alpha|k_low_order|k_pij_lij|k_high_order|k_dii_tau' -c 21 -f -o gpurun_out/r2c_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2c_full.ncu-rep --page raw --csv > gpurun_out/r2c_full_raw.csv 2> /dev/null
ls -la gpurun_out/r2c_full* | head
