#!/bin/bash
# round 2, GPU session N: L2 prefetch distance of the row sweeps (2 .. 5 columns; d_ij as well), and step 2
# as ONE launch of interleaved d_ij / indicator blocks on the large mesh (RYUJIN_B200_SMALL_MESH_ROWS lifted).
set -u
mkdir -p gpurun_out
one() { # one <tag> <lib> <env...>
  tag=$1; lib=$2; shift 2
  env "$@" RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_$lib.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2n_%s.json" % v))
    k = d["roofline"]["kernel_ms_per_update"]
    print("%-22s %7.1f MDoF/s  " % (v, d["value"]) + "  ".join("%s %.2f" % (a, b) for a, b in k.items() if b > 0.2))
except Exception as e:  # noqa: BLE001
    print(v, "failed:", e)
PY
}
one base base X=0
one pf2 pf2 X=0
one pf3 pf3 X=0
one pf4 pf4_mix8 X=0
one pf5 pf5 X=0
one pf4_dij1 pf4_dij1 X=0
one pf4_mix8 pf4_mix8 RYUJIN_B200_SMALL_MESH_ROWS=100000000
one pf4_mix6 pf4_mix6 RYUJIN_B200_SMALL_MESH_ROWS=100000000
one pf4_mix10 pf4_mix10 RYUJIN_B200_SMALL_MESH_ROWS=100000000
one base_b base X=0
( RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_pf4_mix8.so RYUJIN_B200_SMALL_MESH_ROWS=100000000 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -x \
    -k "single_substep or stage_weights or config3 or restart or stepping_modes or config1 or config2 or folded" ) > gpurun_out/r2n_pytest_mix.log 2>&1
echo "pytest mix rc=$?"; tail -2 gpurun_out/r2n_pytest_mix.log
