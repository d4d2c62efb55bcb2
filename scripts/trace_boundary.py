"""Per-CTA timeline of back-to-back forward / adjoint launches (DJETS_TRACE instrumentation): where the time
goes at kernel boundaries.   python scripts/trace_boundary.py <blocks> [name=value ...]"""
import os
import struct
import sys

import numpy as np

PREFIX = "/tmp/djets_trace"
os.environ["DJETS_TRACE"] = PREFIX
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
nblk = int(sys.argv[1])
ctx = dj.init(world=1, devices=[0])
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_param(k, int(v))
rng = np.random.default_rng(0)
pool = [np.asfortranarray(rng.random((4096, 4096), dtype=np.float32)) for _ in range(4)]
A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[i % 4]) if i == j else dj.JopZeroBlock(np.float32, 4096, 4096),
                 dims=(nblk, nblk), pids=[0], dist=[1, 1], isdiag=True, shapes=([(4096,)] * nblk, [(4096,)] * nblk),
                 dtype=np.float32)
m, d, m2 = A.domain().rand(rng), A.range().zeros(), A.domain().zeros()
for _ in range(10):
    A.mul(d, m); A.mul_adjoint(m2, d)
ctx.sync()
ctx.reset_stats()
for _ in range(4):
    A.mul(d, m); A.mul_adjoint(m2, d)
ctx.sync()
del A, m, d, m2
ctx.close()

with open(PREFIX + ".rank0.bin", "rb") as f:
    n = struct.unpack("i", f.read(4))[0]
    meta = struct.unpack(f"{2 * n}i", f.read(8 * n))
    launches = []
    for l in range(n):
        ctas = meta[2 * l + 1]
        launches.append((meta[2 * l], np.frombuffer(f.read(32 * ctas), dtype=np.uint64).reshape(ctas, 4).astype(np.int64)))
t0 = min(int(a[:, 0].min()) for _, a in launches)
print("launch kind ctas | first start, last start of wave 1 | dependency released (min,max) | staged (median after release) | "
      "first end, median end, last end   [us from first start of launch 0]")
prev_end = None
for l, (kind, a) in enumerate(launches):
    us = (a - t0) / 1e3
    start, rel, staged, end = us[:, 0], us[:, 1], us[:, 2], us[:, 3]
    gap = "" if prev_end is None else f" | prev last end {prev_end:8.1f}"
    print(f"{l:2d} {'fwd' if kind == 0 else 'adj'} {len(a):5d} | {start.min():8.1f} {np.sort(start)[min(len(a), 296) - 1]:8.1f} | "
          f"{rel.min():8.1f} {rel.max():8.1f} | {np.median(staged - rel):5.2f} | {end.min():8.1f} {np.median(end):8.1f} {end.max():8.1f}{gap}")
    prev_end = end.max()
# bandwidth-free view: number of CTAs past their dependency wait and not yet done, sampled every 1 us around boundary 1->2
k0, a0 = launches[2]
k1, a1 = launches[3]
lo = (a0[:, 3].max() - t0) / 1e3 - 12
print("t(us)  active(prev)  waiting(next)  running(next)")
for t in np.arange(lo, lo + 24, 1.0):
    tn = t * 1e3 + t0
    act_prev = int(((a0[:, 0] <= tn) & (a0[:, 3] > tn)).sum())
    wait_next = int(((a1[:, 0] <= tn) & (a1[:, 1] > tn)).sum())
    run_next = int(((a1[:, 1] <= tn) & (a1[:, 3] > tn)).sum())
    print(f"{t:7.1f} {act_prev:6d} {wait_next:6d} {run_next:6d}")
