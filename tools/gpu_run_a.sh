#!/bin/bash
# One GPU-box session: full GPU test suite, bench, K5 variants, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/box_info.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/box_info.txt 2>&1; uptime >> gpurun_out/box_info.txt
( time python -m pytest tests -q -m gpu --durations=20 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
for v in rev fwd rev6 rev4 rev10; do
  RYUJIN_B200_LIB=$PWD/gpurun_variants/lib_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dij|k_alpha|k_low_order|k_pij_lij|k_high_order' --launch-skip 6 -c 12 -o /tmp/prof_r1b -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ncu -i /tmp/prof_r1b.ncu-rep --page raw --csv > gpurun_out/prof_r1b_raw.csv 2>/dev/null
ncu -i /tmp/prof_r1b.ncu-rep --page details --csv > gpurun_out/prof_r1b_details.csv 2>/dev/null
sz=$(stat -c %s /tmp/prof_r1b.ncu-rep); if [ "$sz" -lt 50000000 ]; then cp /tmp/prof_r1b.ncu-rep gpurun_out/; fi
ls -la gpurun_out | head -40
