#!/usr/bin/env python3
"""profiles/<tag>_sass_ptxas_summary.txt: per kernel of libryujin_b200.so the static instruction mix from
`cuobjdump -sass` (FP64 arithmetic, memory, bulk-copy / mbarrier / prefetch instructions), the bulk-copy
instructions of one tiled kernel as evidence, and the `ptxas -v` resources of the 3D Euler kernels.

usage: python tools/sass_summary.py r2l [filter=Euler<3>]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else "Euler<3>"
lib = os.path.join(ROOT, "ryujin_b200", "libryujin_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
names = subprocess.run(["c++filt"], input="\n".join(f.split("\n", 1)[0] for f in funcs), capture_output=True, text=True).stdout.split("\n")
COLS = ["DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDL", "STL", "LDS", "STS", "UBLKCP", "SYNCS", "CCTL"]
out = ["# cuobjdump -sass ryujin_b200/libryujin_b200.so (sm_100a), per kernel: static instruction mix;",
       "# UBLKCP = cp.async.bulk (1D TMA bulk copy), SYNCS = mbarrier arrive/expect_tx/try_wait, CCTL = prefetch.global.L2,",
       "# LDS/STS = shared memory, LDL/STL = local memory (spills)",
       "%-58s %6s %6s " % ("kernel (%s and equation independent ones)" % flt, "instr", "fp64") + " ".join("%5s" % c for c in COLS)]
bulk = []
for f, name in zip(funcs, names):
    short = re.sub(r"\(rjb::StencilView.*|\(unsigned.*|\(rjb::.*|\(int.*|\(double.*", "", name.replace("rjb::", "").replace("(int)", ""))
    short = re.sub(r"^void ", "", short)
    if not (flt in short or "<" not in short):
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(2)] += 1
            if "k_pij_lij_tile" in short and re.search(r"UBLKCP|SYNCS|FENCE|UTMACMDFLUSH", line):
                bulk.append("    " + line.strip()[:100])
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    out.append("%-58s %6d %6d " % (short[:58], sum(ops.values()), fp64) + " ".join("%5d" % ops[c] for c in COLS))
out += ["", "# bulk-copy instructions of k_pij_lij_tile (loads complete on the mbarrier, stores are bulk_group):"] + bulk[:24]
log = os.path.join(ROOT, "ryujin_b200", "csrc", "build", "vcl_fma", "module.ptxas.log")
ptx = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ptxas_summary.py"), flt.replace("<", "<").replace("3", "3"), log],
                     capture_output=True, text=True).stdout
out += ["", "# ptxas -v (registers, stack, spills; dynamic shared memory at launch is sized by the widest slice):", ptx]
open(os.path.join(ROOT, "profiles", "%s_sass_ptxas_summary.txt" % tag), "w").write("\n".join(out))
print("\n".join(out[:40]))
