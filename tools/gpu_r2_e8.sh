#!/bin/bash
# round 2, 8-GPU session: multi-GPU parity against the partition oracle (pytest wrapper, 8 ranks incl.
# 2x2x2 blocks), BASELINE config 4 as stated (400^3 in 2x2x2 blocks), strong scaling of fixed meshes,
# and the exposed-communication breakdown (comm_wait / allreduce profile slots).
set -u
mkdir -p gpurun_out
run() { # run <tag> <n> <bench args...>
  tag=$1; n=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2e_%s.json" % tag))
    k = d["roofline"]["kernel_ms_per_update"]
    print("%-26s %8.1f MDoF/s  %7.2f ms/cycle  n_dofs %d  " % (tag, d["value"], d["ms_per_step"], d["config"]["n_dofs"])
          + "  ".join("%s %.2f" % (a, b) for a, b in k.items() if b > 0.05))
except Exception as e:  # noqa: BLE001
    print(tag, "failed:", e)
PY
}
( time timeout 1200 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -s ) > gpurun_out/r2e_pytest_multi_gpu.log 2>&1
echo "pytest rc=$?"; grep -c "vs ORACLE" gpurun_out/r2e_pytest_multi_gpu.log; tail -3 gpurun_out/r2e_pytest_multi_gpu.log
run weak_slabs_n2 2
run weak_slabs_n8 8
run weak_blocks_n8 8 --partition blocks
run strong200_blocks_n2 2 --scaling strong --partition blocks
run strong200_blocks_n4 4 --scaling strong --partition blocks
run strong200_blocks_n8 8 --scaling strong --partition blocks
run strong200_slabs_n8 8 --scaling strong
run strong320_blocks_n8 8 --scaling strong --partition blocks --nodes 320
run strong320_blocks_n4 4 --scaling strong --partition blocks --nodes 320
