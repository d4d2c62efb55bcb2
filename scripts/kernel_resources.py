#!/usr/bin/env python
"""Registers / stack / shared memory of every kernel instantiation in the shipped library
(cuobjdump --dump-resource-usage, no GPU needed): usage  python scripts/kernel_resources.py r02
-> profiles/r02_kernel_resources.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "distributedjets.jl_b200", "libdjets_b200.so")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True, check=True).stdout
    lines = txt.splitlines()
    rows = []
    for k, line in enumerate(lines):
        m = re.match(r"\s*Function\s+(\S+):", line)
        if not m or k + 1 >= len(lines):
            continue
        res = {a: int(b) for a, b in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", lines[k + 1])}
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void\s+", "", name).replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*$", "", name).replace("djets::", "")
        rows.append((name, res))
    rows.sort(key=lambda r: r[0])
    out = os.path.join(ROOT, "profiles", f"{tag}_kernel_resources.txt")
    with open(out, "w") as f:
        f.write(f"# cuobjdump --dump-resource-usage of distributedjets.jl_b200/libdjets_b200.so (sm_100a): {len(rows)} kernel instantiations\n")
        f.write("# REG = registers per thread, STACK = bytes of per-thread stack frame, SHARED = static shared memory, LOCAL = local memory\n")
        hist = collections.Counter(r["REG"] for _, r in rows)
        f.write("# register histogram: " + ", ".join(f"{k}:{v}" for k, v in sorted(hist.items())) + "\n")
        stk = [(n, r) for n, r in rows if r.get("STACK", 0) or r.get("LOCAL", 0)]
        f.write(f"# kernels with a stack frame or local memory: {len(stk)} of {len(rows)} (largest frame "
                f"{max((r.get('STACK', 0) for _, r in rows), default=0)} bytes; LOCAL is 0 everywhere: "
                f"{all(r.get('LOCAL', 0) == 0 for _, r in rows)})\n")
        f.write(f"{'REG':>4} {'STACK':>5} {'SHARED':>6}  kernel\n")
        for n, r in rows:
            f.write(f"{r.get('REG', 0):>4} {r.get('STACK', 0):>5} {r.get('SHARED', 0):>6}  {n}\n")
    print(out, len(rows))


if __name__ == "__main__":
    main()
