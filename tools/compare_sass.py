#!/usr/bin/env python
"""Compare the SASS of two builds of libryujin_b200.so kernel by kernel (addresses stripped).
Used to show that a source change that is meant to be performance neutral (a macro, a test hook,
an option that defaults to off) leaves every kernel of the default build bit-identical.

    python tools/compare_sass.py old/libryujin_b200.so ryujin_b200/libryujin_b200.so
"""
import hashlib
import re
import subprocess
import sys


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"^\s+/\*[0-9a-f]+\*/", line):
            funcs[name].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest(), len(v)) for k, v in funcs.items()}


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:90]


def main():
    a, b = sass(sys.argv[1]), sass(sys.argv[2])
    changed = [k for k in b if k in a and a[k][0] != b[k][0]]
    print("%d kernels in %s, %d in %s, %d changed, %d added, %d removed"
          % (len(a), sys.argv[1], len(b), sys.argv[2], len(changed), len(set(b) - set(a)), len(set(a) - set(b))))
    for k in changed:
        print("  changed: %s (%d -> %d instructions)" % (demangle(k), a[k][1], b[k][1]))
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
