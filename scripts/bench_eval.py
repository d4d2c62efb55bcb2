"""BASELINE config 4 on one GPU: the elementwise program (djets_vec_eval) against the specialised lincomb kernels and
the copy roofline, plus host-returning and device-resident reductions.   python scripts/bench_eval.py [log2n] [f64|f32]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
es = np.dtype(dtype).itemsize
ctx = dj.init(world=1, devices=[0])
nb = 1024
bl = (1 << log2n) // nb
n = nb * bl
R = dj.JetDSpace([[(bl,)] * nb], [0], dtype, ctx)
x, y, z, out = R.zeros().fill(1.5), R.zeros().fill(-2.0), R.zeros().fill(0.25), R.zeros()
L = dj.lazy
T = np.dtype(dtype).type


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


s = dj.Scalar(0.0, dtype, ctx)
cases = [
    ("fill", lambda: out.fill(1.0), 1),
    ("copy (dest .= x)", lambda: out.assign(x), 2),
    ("scale lincomb (dest .= 2x, specialised kernel)", lambda: out.lincomb_([2.0], [x]), 2),
    ("scale eval (dest .= 2x)", lambda: out.eval_(T(2.0) * L(x), wide=False), 2),
    ("axpy lincomb", lambda: y.lincomb_([0.5, 1.0], [x, y]), 3),
    ("axpy eval", lambda: y.eval_(T(0.5) * L(x) + L(y), wide=False), 3),
    ("3-term lincomb", lambda: out.lincomb_([0.1, 0.2, 0.3], [x, y, z]), 4),
    ("3-term eval (benchmarks.jl:43)", lambda: out.eval_(T(0.1) * L(x) + T(0.2) * L(y) - T(0.3) * L(z), wide=False), 4),
    ("abs.(d)", lambda: out.eval_(abs(L(x)), wide=False), 2),
    ("d.^2", lambda: out.eval_(L(x) ** 2, wide=False), 2),
    ("d .+ 1.0", lambda: out.eval_(L(x) + T(1.0), wide=False), 2),
    ("max.(d, 0)", lambda: out.eval_(dj.bc_max(L(x), T(0)), wide=False), 2),
    ("x .* y ./ (abs.(z) .+ 1)", lambda: out.eval_(L(x) * L(y) / (abs(L(z)) + T(1.0)), wide=False), 4),
    ("sqrt.(abs.(x)) in place", lambda: z.eval_(dj.bc_sqrt(abs(L(z))), wide=False), 2),
    ("dot (host value)", lambda: x.dot(y), 2),
    ("norm (host value)", lambda: x.norm(), 1),
    ("dot_s (device value)", lambda: x.dot_s(y, s), 2),
    ("norm_s (device value)", lambda: x.norm_s(2, s), 1),
]
for name, fn, streams in cases:
    ms = timed(fn)
    print(json.dumps({"op": name, "n": n, "dtype": np.dtype(dtype).name, "ms": round(ms, 4),
                      "GBps": round(streams * n * es / ms / 1e6, 1)}), flush=True)
