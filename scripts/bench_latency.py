"""Latency-bound shapes: BASELINE config 1 (4x4 dense Float64 1000x1000, grid [2,2]) and the reference's own benchmark
shape (benchmark/benchmarks.jl:20: 50x10 diagonal blocks of 100 Float64 on a [3,1] grid), eager applies, graph-replayed
applies and a captured CGLS iteration.   python scripts/bench_latency.py [ngpus]"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def timed(ctx, fn, reps=200):
    for _ in range(20):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    dev = ctx.timer_stop() / reps
    ctx.sync()
    return dev * 1e3, (time.perf_counter() - t0) / reps * 1e6


def report(tag, **kw):
    print(json.dumps({"case": tag, **kw}), flush=True)


rng = np.random.default_rng(0)
# ---- config 1 ------------------------------------------------------------------------------------------------------
for world, devices, label in ((4, [0] * 4, "4 ranks on one GPU"),) + (((4, [k % ngpu for k in range(4)], f"4 ranks on {min(ngpu, 4)} GPUs"),) if ngpu > 1 else ()):
    ctx = dj.Context(world=world, devices=devices)
    dj.set_default_context(ctx)
    blocks = [[dj.JopDense(np.asfortranarray(rng.random((1000, 1000)))) for _ in range(4)] for _ in range(4)]
    A = dj.JopDBlock(blocks, pids=[0, 1, 2, 3], dist=[2, 2], ctx=ctx)
    x, y = A.domain().rand(rng), A.range().rand(rng)
    yo, xo = A.range().zeros(), A.domain().zeros()
    for graphs in (1, 0):
        ctx.set_param("graphs", graphs)
        f_us, f_wall = timed(ctx, lambda: A.mul(yo, x))
        a_us, a_wall = timed(ctx, lambda: A.mul_adjoint(xo, y))
        report("config1", where=label, graphs=graphs, fwd_us=round(f_us, 2), adj_us=round(a_us, 2),
               fwd_applies_per_s=round(1e6 / f_us), adj_applies_per_s=round(1e6 / a_us), fwd_wall_us=round(f_wall, 2))
    ctx.set_param("graphs", 1)
    with ctx.capture() as g:
        A.mul(yo, x)
        A.mul_adjoint(xo, yo)
    p_us, p_wall = timed(ctx, g.launch)
    report("config1 forward+adjoint captured as one sequence", where=label, pair_us=round(p_us, 2),
           applies_per_s=round(2e6 / p_us), wall_us=round(p_wall, 2))
    # one CGLS iteration in steady state: eager (dots return to the host, which forms alpha / beta) against captured
    # (device scalars, one graph launch per iteration)
    SC = sys.modules["djets_b200._lib"]
    L = dj.lazy
    b = A @ x
    xs, r, q = A.domain().zeros(), b.copy(), A.range().zeros()
    s = A.T @ r
    p = s.copy()
    T = A.dtype

    def eager_iter(state={"gamma": 1.0}):
        A.mul(q, p)
        alpha = state["gamma"] / max(float(q.dot(q)), 1e-300)
        xs.lincomb_([1.0, alpha], [xs, p])
        r.lincomb_([1.0, -alpha], [r, q])
        A.mul_adjoint(s, r)
        gnew = float(s.dot(s))
        float(r.norm())
        p.lincomb_([1.0, gnew / max(state["gamma"], 1e-300)], [s, p])
        state["gamma"] = 1.0                                      # keep the arithmetic finite: this is a timing loop

    e_us, e_wall = timed(ctx, eager_iter, reps=100)
    gamma, gnew, delta, alpha, beta, rnorm = (dj.Scalar(1.0, T, ctx) for _ in range(6))
    with ctx.capture() as step:
        A.mul(q, p)
        q.dot_s(q, delta)
        alpha.assign(SC.SC_DIV, gamma, delta)
        xs.assign(L(xs) + alpha * L(p))
        r.assign(L(r) - alpha * L(q))
        A.mul_adjoint(s, r)
        s.dot_s(s, gnew)
        beta.assign(SC.SC_DIV, gnew, gamma)
        p.assign(L(s) + beta * L(p))
        r.norm_s(2, rnorm)
    c_us, c_wall = timed(ctx, step.launch, reps=100)
    report("config1 CGLS iteration (mul!, adjoint, 2 dots, norm, 3 updates)", where=label, eager_stream_us=round(e_us, 1),
           eager_wall_us=round(e_wall, 1), captured_stream_us=round(c_us, 1), captured_wall_us=round(c_wall, 1))
    del step, xs, r, q, s, p
    A.close()
    del A, x, y, yo, xo, b, g
    ctx.close()

# ---- the reference's benchmark shape ---------------------------------------------------------------------------------
ctx = dj.Context(world=3, devices=[k % ngpu for k in range(3)])
dj.set_default_context(ctx)
blocks = [[dj.JopDiagonal(rng.random(100)) for _ in range(10)] for _ in range(50)]
A = dj.JopDBlock(blocks, pids=[0, 1, 2], dist=[3, 1], ctx=ctx)
x, y = A.domain().rand(rng), A.range().rand(rng)
yo, xo = A.range().zeros(), A.domain().zeros()
d1, d2, d3, d4 = (A.range().rand(rng) for _ in range(4))
for graphs in (1, 0):
    ctx.set_param("graphs", graphs)
    f_us, f_wall = timed(ctx, lambda: A.mul(yo, x))
    a_us, a_wall = timed(ctx, lambda: A.mul_adjoint(xo, y))
    report("benchmarks.jl:20 shape (50x10 diagonal blocks of 100 F64, grid [3,1])", gpus=min(ngpu, 3), graphs=graphs,
           fwd_us=round(f_us, 2), adj_us=round(a_us, 2), fwd_wall_us=round(f_wall, 2), adj_wall_us=round(a_wall, 2))
L = dj.lazy
for name, fn in (("norm", lambda: d1.norm()), ("dot", lambda: d1.dot(d2)), ("maximum", lambda: d1.maximum()),
                 ("broadcasting d4 .= a1*d1 .+ a2*d2 .- a3*d3", lambda: d4.assign(0.3 * L(d1) + 0.5 * L(d2) - 0.7 * L(d3)))):
    us, wall = timed(ctx, fn)
    report("benchmarks.jl DBArray group: " + name, gpus=min(ngpu, 3), stream_us=round(us, 2), wall_us=round(wall, 2))
