"""Reduction kinds side by side on one GPU (dot, norms, sum, max)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402
dj = djets_loader.load()
ctx = dj.init(world=1, devices=[0])
for logn in (27, 28):
    nb, bl = 1 << (logn - 20), 1 << 20
    x = dj.DBArray(lambda i: np.full(bl, 1.0 + i), (nb,), [0])
    y = x.similar().fill(2.0)
    def timed(fn, reps=20):
        for _ in range(3): fn()
        ctx.sync(); ctx.timer_start()
        for _ in range(reps): fn()
        return ctx.timer_stop() / reps
    n = nb * bl
    for name, fn, nbytes in (("dot", lambda: x.dot(y), 2*n*8), ("norm2", lambda: x.norm(2), n*8), ("norm1", lambda: x.norm(1), n*8),
                             ("sum", lambda: x.sum(), n*8), ("max", lambda: x.maximum(), n*8), ("dot_xx", lambda: x.dot(x), n*8)):
        ms = timed(fn)
        print(f"2^{logn} {name:7s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.0f} GB/s")
