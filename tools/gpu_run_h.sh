#!/bin/bash
set -u
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multi_gpu_parity.py > gpurun_out/mgpu_n$N.log 2>&1
grep -E "rel Linf|parity ok|False|Error|error" gpurun_out/mgpu_n$N.log | head -12
timeout 300 $TR --master-port 29521 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench${N}_default.json 2> gpurun_out/bench${N}_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench${N}_default.json')); print('N=$N default', d['value'], d['ms_per_step'], d['e2e']['value'])"
