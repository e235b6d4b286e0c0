#!/usr/bin/env python3
"""Condense ryujin_b200/csrc/build/module.ptxas.log (nvcc -Xptxas -v) into one line per kernel:
registers, stack frame, spill stores/loads.  Usage: tools/ptxas_summary.py [filter] [log]"""
import re
import subprocess
import sys

flt = sys.argv[1] if len(sys.argv) > 1 else "Euler<(int)3>"
log = sys.argv[2] if len(sys.argv) > 2 else "ryujin_b200/csrc/build/vcl_fma/module.ptxas.log"
s = open(log).read()
items = re.findall(
    r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(\d+) bytes stack frame, "
    r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?",
    s,
    flags=re.S,
)
names = subprocess.run(["c++filt"], input="\n".join(i[0] for i in items), capture_output=True, text=True).stdout.split("\n")
for (name, stack, ss, sl, regs, smem), dem in zip(items, names):
    dem = re.sub(r"\(rjb::StencilView.*|\(unsigned.*|\(unsigned long.*", "", dem).replace("(int)", "").replace("rjb::", "")
    dem = re.sub(r"^void ", "", dem)
    if flt in dem.replace("(int)", "") or flt == "all":
        print("%-78s regs=%-4s stack=%-5s spill=%s/%s smem=%s" % (dem[:78], regs, stack, ss, sl, smem or 0))
