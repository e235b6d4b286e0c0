#!/bin/bash
# round 2, 2-GPU session: the three overlap modes of the halo exchange on the weak-scaling bench
# (RYUJIN_B200_OVERLAP=1: export rows, then exchange || interior rows; 2: (export rows, exchange) || interior
# rows; 0: no second stream), multi-GPU parity against the partition oracle, the strong-scaling point at N = 2.
set -u
mkdir -p gpurun_out
run() { # run <tag> <n> <env> <bench args...>
  tag=$1; n=$2; envs=$3; shift 3
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2k_%s.json" % tag))
    k = d["roofline"]["kernel_ms_per_update"]
    print("%-26s %8.1f MDoF/s  %7.2f ms/cycle  e2e %7.1f  " % (tag, d["value"], d["ms_per_step"], d["e2e"]["value"])
          + "  ".join("%s %.2f" % (a, b) for a, b in k.items() if b > 0.05))
except Exception as e:  # noqa: BLE001
    print(tag, "failed:", e)
PY
}
( time timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -s ) > gpurun_out/r2k_pytest_multi_gpu.log 2>&1
echo "pytest rc=$?"; grep -c "vs ORACLE" gpurun_out/r2k_pytest_multi_gpu.log; tail -3 gpurun_out/r2k_pytest_multi_gpu.log
run weak_n2_ov1 2 RYUJIN_B200_OVERLAP=1
run weak_n2_ov2 2 RYUJIN_B200_OVERLAP=2
run weak_n2_ov0 2 RYUJIN_B200_OVERLAP=0
run weak_n2_ov1_b 2 RYUJIN_B200_OVERLAP=1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_n1.json 2> gpurun_out/r2k_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_n1.json')); print('n1 %.1f MDoF/s %.2f ms/cycle e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
