#!/bin/bash
# Build experimental flavours of libryujin_b200.so into gpurun_variants/ (they travel to the GPU
# box with the snapshot; the directory is git-ignored through *.so).  Each line: name:nvcc flags.
# Run here (no GPU needed), then `gpurun -- 'bash tools/gpu_variants.sh'` on the box.
set -eu
cd "$(dirname "$0")/../ryujin_b200/csrc"
OUT=../../gpurun_variants
mkdir -p "$OUT"
VARIANTS=${VARIANTS:-"
base:
prefetch:-DRJ_PREFETCH_INDEX=1
prefetch_alpha8:-DRJ_PREFETCH_INDEX=1 -DRJ_MB_ALPHA=8
prefetch_alpha8_u2:-DRJ_PREFETCH_INDEX=1 -DRJ_MB_ALPHA=8 -DRJ_UNROLL_ALPHA=2
prefetch_pij6:-DRJ_PREFETCH_INDEX=1 -DRJ_MB_PIJ=6
prefetch_low_u2:-DRJ_PREFETCH_INDEX=1 -DRJ_UNROLL_LOW=2
prefetch_low_u2_mb5:-DRJ_PREFETCH_INDEX=1 -DRJ_UNROLL_LOW=2 -DRJ_MB_LOW=5
alpha_edge:-DRJ_ALPHA_EDGE=1
alpha_edge_w8:-DRJ_ALPHA_EDGE=1 -DRJ_ALPHA_EDGE_W=8 -DRJ_MB_ALPHA_EDGE=2
alpha_edge_mb4:-DRJ_ALPHA_EDGE=1 -DRJ_MB_ALPHA_EDGE=4
alpha_edge_prefetch:-DRJ_ALPHA_EDGE=1 -DRJ_PREFETCH_INDEX=1
pij_edge:-DRJ_PIJ_EDGE=1
pij_edge_w8:-DRJ_PIJ_EDGE=1 -DRJ_PIJ_EDGE_W=8 -DRJ_MB_PIJ_EDGE=2
ho_edge:-DRJ_HO_EDGE=1
ho_edge_mb4:-DRJ_HO_EDGE=1 -DRJ_MB_HO_EDGE=4
edge_all:-DRJ_ALPHA_EDGE=1 -DRJ_PIJ_EDGE=1 -DRJ_HO_EDGE=1
edge_all_prefetch:-DRJ_ALPHA_EDGE=1 -DRJ_PIJ_EDGE=1 -DRJ_HO_EDGE=1 -DRJ_PREFETCH_INDEX=1
"}
echo "$VARIANTS" | while IFS= read -r line; do
  [ -z "$line" ] && continue
  name=${line%%:*}; flags=${line#*:}
  rm -f build/vcl_fma/module.o
  make -s EXTRA_NVFLAGS="$flags" OUT="$OUT/lib_$name.so" 2>&1 | grep -E "error" && exit 1
  echo "built $name ($flags)"
done
rm -f build/vcl_fma/module.o
make -s -j2   # restore the default object and library
