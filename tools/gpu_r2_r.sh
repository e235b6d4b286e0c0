#!/bin/bash
# round 2, last single-GPU session: the reference arm as the driver launches it, and one more default bench line.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
echo "reference rc=$?"; cut -c1-700 gpurun_out/r2r_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_default.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['fp64'])"
