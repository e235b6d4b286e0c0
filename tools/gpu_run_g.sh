#!/bin/bash
# N GPUs (interior ranks have two neighbours): parity incl. stepping modes, bench with overlap on/off
set -u
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/multi_gpu_parity.py > gpurun_out/mgpu_n$N.log 2>&1
grep -E "rel Linf|parity ok|False|Error|error" gpurun_out/mgpu_n$N.log | head -12
for ov in 1 0; do
  RYUJIN_B200_OVERLAP=$ov timeout 300 $TR --master-port 2952$ov bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench${N}_ov$ov.json 2> gpurun_out/bench${N}_ov$ov.err
  python -c "
import json; d=json.load(open('gpurun_out/bench${N}_ov$ov.json')); print('N=$N overlap $ov', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
