#!/bin/bash
# round 2, GPU session A: parity suite with the reference's pow (VCL), smoke, bench of the default
# build and of the prepared kernel variants (gpurun_variants/), ncu launch list of the default.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_box.txt 2>&1
free -g >> gpurun_out/r2a_box.txt; nproc >> gpurun_out/r2a_box.txt
( time timeout 1500 python -m pytest tests -q -m gpu -x -s ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2a_variants.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
echo "ncu rc=$?"
