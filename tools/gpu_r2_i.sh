#!/bin/bash
# round 2, GPU session I: the fused per-update kernel on hardware (parity of fused vs separate kernels,
# stepping modes, streaming transfers), small-mesh cycle times over mesh sizes, edge cache inside k_alpha
# (RJ_EDGE_CACHE=3) with occupancy flavours, the default build after the body/wrapper refactor.
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "fused_update or stepping_modes or streaming or restart or numeric_error or config1" ) > gpurun_out/r2i_pytest_fused.log 2>&1
echo "pytest fused rc=$?"; tail -3 gpurun_out/r2i_pytest_fused.log
( time timeout 900 python tools/small_mesh_stepping.py ) > gpurun_out/r2i_small_mesh.txt 2>&1
echo "small mesh rc=$?"; grep -v "^$" gpurun_out/r2i_small_mesh.txt | tail -40
VARIANT_NAMES="base ec3 ec3_a6 ec3_a5 ec3_a6_u2 base" bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2i_variants.txt
python -c "
import json; d=json.load(open('gpurun_out/var_base.json')); print('e2e', d['e2e'])"
