"""Where an apply of BASELINE config 1 (4x4 dense Float64 1000x1000 = 128 MB) spends its time on one GPU: timings of the
operator held by one rank and by four ranks (grid [2,2]) under a few settings, then the per-CTA time stamps (DJETS_TRACE)
of a few eager applies.   python scripts/trace_config1.py"""
import json
import os
import struct
import sys

import numpy as np

PREFIX = "/tmp/djets_trace_c1"
os.environ["DJETS_TRACE"] = PREFIX
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
rng = np.random.default_rng(0)
payload = [[np.asfortranarray(rng.random((1000, 1000))) for _ in range(4)] for _ in range(4)]


def timed(ctx, fn, reps=300):
    for _ in range(30):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps * 1e3


def build(world, params):
    ctx = dj.Context(world=world, devices=[0] * world)
    dj.set_default_context(ctx)
    for k, v in params.items():
        ctx.set_param(k, v)
    blocks = [[dj.JopDense(payload[i][j]) for j in range(4)] for i in range(4)]
    A = dj.JopDBlock(blocks, pids=list(range(world)), dist=[2, 2] if world == 4 else [1, 1], ctx=ctx)
    return ctx, A


for world in (1, 4):
    for params in ({}, {"l2_keep": 0}, {"graphs": 0}, {"pdl": 0}, {"prefetch": 0}):
        ctx, A = build(world, params)
        x, y = A.domain().rand(rng), A.range().rand(rng)
        yo, xo = A.range().zeros(), A.domain().zeros()
        f = timed(ctx, lambda: A.mul(yo, x))
        a = timed(ctx, lambda: A.mul_adjoint(xo, y))

        def pair():
            A.mul(yo, x)
            A.mul_adjoint(xo, yo)

        p = timed(ctx, pair)
        print(json.dumps({"ranks": world, "params": params, "fwd_us": round(f, 2), "adj_us": round(a, 2),
                          "pair_us": round(p, 2), "tiles": [A.schedule_info(False)[0], A.schedule_info(True)[0]]}), flush=True)
        A.close()
        del A, x, y, yo, xo
        ctx.close()

# ---- trace: eager applies, forward x3 then adjoint x3 ---------------------------------------------------------------
for world in (1, 4):
    for f in os.listdir("/tmp"):
        if f.startswith("djets_trace_c1"):
            os.remove(os.path.join("/tmp", f))
    ctx, A = build(world, {"graphs": 0})
    x, y = A.domain().rand(rng), A.range().rand(rng)
    yo, xo = A.range().zeros(), A.domain().zeros()
    for _ in range(20):
        A.mul(yo, x)
    ctx.sync()
    ctx.reset_stats()
    for _ in range(3):
        A.mul(yo, x)
    for _ in range(3):
        A.mul_adjoint(xo, y)
    ctx.sync()
    A.close()
    del A, x, y, yo, xo
    ctx.close()
    rows = []
    for r in range(world):
        path = f"{PREFIX}.rank{r}.bin"
        if not os.path.exists(path):
            continue
        with open(path, "rb") as fh:
            n = struct.unpack("i", fh.read(4))[0]
            meta = struct.unpack(f"{2 * n}i", fh.read(8 * n))
            for l in range(n):
                ctas = meta[2 * l + 1]
                a = np.frombuffer(fh.read(32 * ctas), dtype=np.uint64).reshape(ctas, 4).astype(np.int64)
                rows.append((r, l, meta[2 * l], a))
    t0 = min(int(a[:, 0].min()) for _, _, _, a in rows)
    print(f"--- {world} rank(s): rank launch kind ctas | first start, last start | released max | first end, median end, last end (us)")
    for r, l, kind, a in sorted(rows, key=lambda q: int(q[3][:, 0].min())):
        us = (a - t0) / 1e3
        print(f"{r} {l:2d} {'fwd' if kind == 0 else 'adj'} {len(a):4d} | {us[:, 0].min():7.1f} {us[:, 0].max():7.1f} | {us[:, 1].max():7.1f} | "
              f"{us[:, 3].min():7.1f} {np.median(us[:, 3]):7.1f} {us[:, 3].max():7.1f}")
