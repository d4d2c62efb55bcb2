"""A/B of one tuning switch inside one process (alternating, several rounds) on the benchmark operator.
    python scripts/ab_param.py <param> <blocks> [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
param, nblk = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ctx = dj.init(world=1, devices=[0])
rng = np.random.default_rng(0)
pool = [np.asfortranarray(rng.random((4096, 4096), dtype=np.float32)) for _ in range(4)]
A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[i % 4]) if i == j else dj.JopZeroBlock(np.float32, 4096, 4096),
                 dims=(nblk, nblk), pids=[0], dist=[1, 1], isdiag=True, shapes=([(4096,)] * nblk, [(4096,)] * nblk),
                 dtype=np.float32)
m, d, m2 = A.domain().rand(rng), A.range().zeros(), A.domain().zeros()
nbytes = 2 * A.algorithmic_bytes()


def run():
    for _ in range(5):
        A.mul(d, m); A.mul_adjoint(m2, d)
    ctx.sync()
    ctx.timer_start()
    for _ in range(steps):
        A.mul(d, m); A.mul_adjoint(m2, d)
    return ctx.timer_stop() / steps


for trial in range(5):
    row = []
    for v in (0, 1):
        ctx.set_param(param, v)
        if param not in ("pdl",):
            A.replan()
        ms = run()
        row.append(f"{param}={v}: {ms * 1e3:8.1f} us {nbytes / ms / 1e6:8.1f} GB/s")
    print(" | ".join(row), flush=True)
