"""Throughput across block shapes (one GPU): looks for performance cliffs away from the BASELINE shapes.
    python scripts/bench_shapes.py
Full nb x nb operators of dense blocks cycling through a pool of 4 random payloads."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ctx = dj.init(world=1, devices=[0])
only = None
for kv in sys.argv[1:]:                                   # context parameters, e.g. persistent=1; cases=3,12 picks cases
    k, v = kv.split("=")
    if k == "cases":
        only = {int(c) for c in v.split(",")}
    else:
        ctx.set_param(k, int(v))
CASES = [  # dtype, m, n, nb (nb x nb blocks)
    ("f64", 1000, 1000, 32), ("f32", 1023, 1025, 45), ("f32", 256, 256, 128), ("f64", 100, 100, 300),
    ("f32", 4096, 512, 20), ("f32", 512, 4096, 20), ("f64", 5000, 64, 40), ("f64", 64, 5000, 40),
    ("f32", 8192, 8192, 4), ("f64", 333, 777, 60), ("c128", 2048, 2048, 12), ("c64", 2048, 2048, 16), ("c128", 300, 257, 60),
    ("f32", 128, 128, 300), ("f64", 64, 64, 400), ("f64", 100, 100, 120), ("f32", 600, 600, 60),
]
for case, (name, m, n, nb) in enumerate(CASES):
    if only is not None and case not in only:
        continue
    dtype = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}[name]
    rng = np.random.default_rng(1)
    if name.startswith("c"):
        rt = np.float32 if name == "c64" else np.float64
        pool = [np.asfortranarray((rng.random((m, n), dtype=rt) + 1j * rng.random((m, n), dtype=rt)).astype(dtype)) for _ in range(4)]
    else:
        pool = [np.asfortranarray(rng.random((m, n), dtype=dtype)) for _ in range(4)]
    A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[(i + 3 * j) % 4]), dims=(nb, nb), pids=[0], dist=[1, 1],
                     shapes=([(m,)] * nb, [(n,)] * nb), dtype=dtype)
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

    tf, ta = timed(lambda: A.mul(yo, x)), timed(lambda: A.mul_adjoint(xo, y))
    lhs, rhs = complex(yo.dot(y)), complex(x.dot(xo))
    print(json.dumps({"dtype": name, "block": [m, n], "blocks": [nb, nb], "GiB": round(nbytes / 2**30, 2),
                      "fwd_ms": round(tf, 3), "adj_ms": round(ta, 3), "fwd_GBps": round(nbytes / tf / 1e6),
                      "adj_GBps": round(nbytes / ta / 1e6), "dot_test": abs(lhs - rhs) / abs(lhs),
                      "tiles": [A.schedule_info(False)[0], A.schedule_info(True)[0]]}), flush=True)
    A.close()
    del A, x, y, yo, xo
