#!/bin/bash
# round 2, final single-GPU session: the whole -m gpu suite on the final build (edge cache on, folded scalar
# kernels), smoke(), the default bench line with the CPU arm, the ncu launch list of a bench run and one
# ncu --set full capture of the sweep kernels (three updates: stages 0, 1, 2), exported as CSV.
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2l_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2l_smoke.log
timeout 900 python bench.py > gpurun_out/r2l_bench_default.json 2> gpurun_out/r2l_bench_default.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2l_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_ncu_bench.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r2l_launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_dij|k_alpha|k_low_order|k_pij_lij|k_high_order|k_dii_tau' -c 21 -f -o /tmp/r2l_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/r2l_full.ncu-rep --page raw --csv > gpurun_out/r2l_full_raw.csv 2> /dev/null
ncu -i /tmp/r2l_full.ncu-rep --page details --csv > gpurun_out/r2l_full_details.csv 2> /dev/null
ls -la /tmp/r2l_full.ncu-rep gpurun_out/r2l_full*; du -sh gpurun_out
