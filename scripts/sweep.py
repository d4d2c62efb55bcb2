"""Tuning sweep of the tile-schedule parameters on one GPU (run under gpurun).
    python scripts/sweep.py [f32|f64] [blocks] [m] [n]
Prints one line per parameter combination: forward / adjoint GB/s from CUDA events around 10 applies."""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
dtype = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
n = int(sys.argv[4]) if len(sys.argv) > 4 else m
ctx = dj.init(world=1, devices=[0])
rng = np.random.default_rng(0)
base = np.asfortranarray(rng.random((m, n), dtype=dtype))
A = dj.JopDBlock(lambda i, j: dj.JopDense(base) if i == j else dj.JopZeroBlock(dtype, m, n), dims=(nblk, nblk),
                 pids=[0], dist=[1, 1], isdiag=True, shapes=([(m,)] * nblk, [(n,)] * nblk), dtype=dtype)
x, y = A.domain().rand(rng), A.range().rand(rng)
yo, xo = A.range().zeros(), A.domain().zeros()
nbytes = A.algorithmic_bytes()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


print(f"# dtype={np.dtype(dtype).name} blocks={nblk} {m}x{n} bytes/apply={nbytes}")
best = (0, None)
for tx, tc, pol in itertools.product((32, 64, 128, 256), (128, 256, 512, 1024), (0, 1)):
    ctx.set_param("fwd_tx", tx); ctx.set_param("fwd_tc", tc); ctx.set_param("ld_policy", pol)
    A.replan()
    ms = timed(lambda: A.mul(yo, x))
    gbs = nbytes / ms / 1e6
    best = max(best, (gbs, (tx, tc, pol)))
    print(f"fwd tx={tx:3d} tc={tc:4d} pol={pol} tiles={A.schedule_info(False)} {ms*1e3:8.1f} us {gbs:8.1f} GB/s", flush=True)
print("# best fwd", best)
best = (0, None)
for ru, tc, pol in itertools.product((1, 2, 4), (8, 16, 32, 64, 128), (0, 1)):
    ctx.set_param("adj_ru", ru); ctx.set_param("adj_tc", tc); ctx.set_param("ld_policy", pol)
    A.replan()
    ms = timed(lambda: A.mul_adjoint(xo, y))
    gbs = nbytes / ms / 1e6
    best = max(best, (gbs, (ru, tc, pol)))
    print(f"adj ru={ru} tc={tc:4d} pol={pol} tiles={A.schedule_info(True)} {ms*1e3:8.1f} us {gbs:8.1f} GB/s", flush=True)
print("# best adj", best)
