#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -14 gpurun_out/pytest_gpu.log
python - > gpurun_out/prm_run.log 2>&1 <<'PY'
import json, tempfile, os, subprocess, sys
d = json.load(open("tests/golden/prm_runs.json"))
run = d["runs"][1]
path = os.path.join(tempfile.mkdtemp(), "deck.prm")
open(path, "w").write(run["prm"].replace("set terminal update interval  = 0", "set terminal update interval  = 1"))
print(subprocess.run([sys.executable, "-m", "ryujin_b200", path], capture_output=True, text=True).stdout)
print("--- golden ---")
print(run["output"])
PY
tail -40 gpurun_out/prm_run.log
