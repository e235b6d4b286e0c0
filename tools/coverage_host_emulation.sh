#!/bin/bash
# Line coverage of the kernel sources (euler.cuh, shallow_water.cuh, kernels.cuh, ...) by the GPU
# parity tests, measured on the host emulation build with gcov.  Prints the lines that no template
# instantiation executed.
set -eu
cd "$(dirname "$0")/../tests/host_emulation"
make -s
rm -rf build_cov && mkdir -p build_cov
for f in module device_offline; do
  /usr/bin/g++ -O0 -g --coverage -std=c++17 -fPIC -fopenmp -ffp-contract=off -I. -I../../ryujin_b200/csrc \
    -Wno-unknown-pragmas -x c++ -c ../../ryujin_b200/csrc/$f.cu -o build_cov/$f.o
done
/usr/bin/g++ -shared -fopenmp -pthread --coverage -o build_cov/libemu_cov.so build_cov/module.o build_cov/device_offline.o \
  build/host_offline.o build/host_emulation.o
( cd ../.. && RYUJIN_B200_LIB=$PWD/tests/host_emulation/build_cov/libemu_cov.so \
  python -m pytest tests -q -m gpu -p no:cacheprovider \
  -k "not (conservation_and_symmetry or config1 or vortex_3d_cycles or golden_on_gpu or verification_goldens or decks_reproduce)" )
gcov -b -o build_cov build_cov/module.gcda > build_cov/gcov.log 2>&1
python - <<'PY'
import re
for f in ("euler.cuh.gcov", "shallow_water.cuh.gcov", "kernels.cuh.gcov", "initial_values.cuh.gcov", "accurate_pow.h.gcov"):
    merged = {}
    for l in open(f):
        m = re.match(r"^\s*([0-9#=\-\*]+):\s*(\d+):(.*)$", l)
        if not m or int(m.group(2)) == 0 or m.group(1) == "-":
            continue
        c = 0 if m.group(1)[0] in "#=" else int(m.group(1).rstrip("*"))
        ln = int(m.group(2))
        merged[ln] = (max(merged.get(ln, (0, ""))[0], c), m.group(3))
    dead = [(ln, src) for ln, (c, src) in sorted(merged.items()) if c == 0]
    print("== %s: %d lines with code, %d never executed" % (f, len(merged), len(dead)))
    for ln, src in dead:
        print("   %4d:%s" % (ln, src[:110]))
PY
rm -f *.gcov
