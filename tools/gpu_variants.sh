#!/bin/bash
# On the GPU box: bench every flavour in gpurun_variants/ (5 steps, no CPU baseline) and print the
# per-kernel times next to each other.  Results: gpurun_out/var_<name>.json
# A name ending in _tileN runs with RYUJIN_B200_TILE=N (which sweeps run as slice-tiled kernels);
# several _tileN names may share one library: "lib_<stem>.so" is looked up without the suffix.
set -u
mkdir -p gpurun_out
for v in ${VARIANT_NAMES:-$(ls gpurun_variants/lib_*.so | sed 's#.*/lib_##; s#\.so$##')}; do
  stem=${v%_tile[0-9]*}
  f=gpurun_variants/lib_$stem.so
  unset RYUJIN_B200_TILE
  case "$v" in *_tile[0-9]*) export RYUJIN_B200_TILE=${v##*_tile};; esac
  RYUJIN_B200_LIB=$PWD/$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/var_%s.json" % v))
    k = d["roofline"]["kernel_ms_per_update"]
    print("%-22s %7.1f MDoF/s  " % (v, d["value"]) + "  ".join("%s %.2f" % (a, b) for a, b in k.items() if b > 0.2))
except Exception as e:  # noqa: BLE001
    print(v, "failed:", e)
PY
done
