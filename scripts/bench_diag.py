"""Throughput of diagonal-block operators (the reference's own fixture kind: d .= diag .* m).
    python scripts/bench_diag.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ctx = dj.init(world=1, devices=[0])
only = None
for kv in sys.argv[1:]:                                   # only=2 picks a case; anything else is a context parameter
    k, v = kv.split("=")
    if k == "only":
        only = int(v)
    else:
        ctx.set_param(k, int(v))
for case, (name, nb, bl, full) in enumerate((("f64", 512, 1 << 20, False), ("f32", 512, 1 << 21, False), ("f64", 24, 1 << 20, True))):
    if only is not None and case != only:
        continue
    dtype = np.float32 if name == "f32" else np.float64
    rng = np.random.default_rng(0)
    pool = [rng.random(bl, dtype=dtype) for _ in range(4)]
    if full:   # full nb x nb operator of diagonal blocks: nb terms per output element
        A = dj.JopDBlock(lambda i, j: dj.JopDiagonal(pool[(i + j) % 4]), dims=(nb, nb), pids=[0], dist=[1, 1],
                         shapes=([(bl,)] * nb, [(bl,)] * nb), dtype=dtype)
    else:      # block-diagonal
        A = dj.JopDBlock(lambda i, j: dj.JopDiagonal(pool[i % 4]) if i == j else dj.JopZeroBlock(dtype, bl, bl),
                         dims=(nb, nb), pids=[0], dist=[1, 1], isdiag=True, shapes=([(bl,)] * nb, [(bl,)] * nb), dtype=dtype)
    x = A.domain().rand(rng)
    y = A.range().zeros()
    nbytes = A.algorithmic_bytes()
    for _ in range(3):
        A.mul(y, x)
    ctx.sync()
    ctx.timer_start()
    for _ in range(10):
        A.mul(y, x)
    ms = ctx.timer_stop() / 10
    got = y.getblock(2)[:5]
    want = sum(pool[(2 + j) % 4][:5].astype(np.float64) * x.getblock(j)[:5] for j in range(1, nb + 1)) if full \
        else pool[2 % 4][:5].astype(np.float64) * x.getblock(2)[:5]
    print(json.dumps({"dtype": name, "blocks": nb, "block_len": bl, "full": full, "GiB": round(nbytes / 2**30, 2), "ms": round(ms, 3),
                      "GBps": round(nbytes / ms / 1e6), "relerr": float(np.max(np.abs(got - want) / np.abs(want)))}), flush=True)
    A.close()
    del A, x, y
