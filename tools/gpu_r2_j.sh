#!/bin/bash
# round 2, GPU session J: launch diet for small meshes (scalar kernels folded into the sweeps, d_ij + alpha as
# one launch) against the separate kernels and the fused kernel (4 and 8 resident blocks per SM); pow
# constants in constant memory; edge cache in k_alpha (final form); the build before the body/wrapper refactor.
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "fused_update or stepping_modes or streaming or restart or numeric_error or config1 or other_schemes" ) > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
( time timeout 900 python tools/small_mesh_stepping.py ) > gpurun_out/r2j_small_mesh.txt 2>&1
echo "small mesh rc=$?"; grep "dofs" gpurun_out/r2j_small_mesh.txt
( RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_fused8.so timeout 300 python tools/small_mesh_stepping.py --quick ) > gpurun_out/r2j_small_mesh_fused8.txt 2>&1
echo "fused8 rc=$?"; grep "fused" gpurun_out/r2j_small_mesh_fused8.txt
VARIANT_NAMES="base prerefactor constmem ec ec_constmem base" bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/r2j_variants.txt
