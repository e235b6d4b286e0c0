#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['whole_update']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
