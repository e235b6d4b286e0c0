#!/bin/bash
# 2 GPUs: overlap on/off -- parity (incl. stepping modes) and bench
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ov in 1 0; do
  RYUJIN_B200_OVERLAP=$ov timeout 300 $TR --master-port 2951$ov tests/multi_gpu_parity.py > gpurun_out/mgpu_ov$ov.log 2>&1
  grep -E "rel Linf|parity ok|False|Error|error" gpurun_out/mgpu_ov$ov.log | head -12
  RYUJIN_B200_OVERLAP=$ov timeout 300 $TR --master-port 2952$ov bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_ov$ov.json 2> gpurun_out/bench2_ov$ov.err
  python -c "
import json; d=json.load(open('gpurun_out/bench2_ov$ov.json')); print('overlap $ov', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
