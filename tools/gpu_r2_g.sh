#!/bin/bash
# round 2, GPU session G: the whole -m gpu suite after euler_aeos landed (first run of the aeos tests on
# hardware), smoke(), the default bench line with the CPU arm, the small-mesh stepping baseline.
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2g_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2g_smoke.log
timeout 900 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
echo "bench rc=$?"; cat gpurun_out/r2g_bench_default.json | cut -c1-600
timeout 300 python tools/small_mesh_stepping.py > gpurun_out/r2g_small_mesh.txt 2>&1
echo "small rc=$?"; cat gpurun_out/r2g_small_mesh.txt | tail -12
