"""Secondary measurements: the BASELINE.json configs other than the one bench.py quotes, on ONE GPU
(ranks of a grid share the GPU, each on its own stream), with size-independent parity properties
checked at full size.  One JSON line per config on stdout; run under gpurun:

    python scripts/bench_configs.py [1] [3] [4] [5] [--small]

Payloads: U[0,1) from seeded generators; large operators cycle through a pool of distinct random
blocks (every block still has its own device copy, so HBM traffic is the real thing).
"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
PEAK = 6451.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SMALL = "--small" in sys.argv


def timed(ctx, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


def dot_test(A, rng):
    """<A x, y> == <x, A' y>: exact in exact arithmetic for any operator, any size."""
    x, y = A.domain().rand(rng), A.range().rand(rng)
    Ax, Aty = A @ x, A.T @ y
    lhs, rhs = float(Ax.dot(y)), float(x.dot(Aty))
    return abs(lhs - rhs) / max(abs(lhs), 1e-300)


def pool_blocks(rng, n, m, k, dtype):
    return [np.asfortranarray(rng.random((m, k), dtype=dtype)) for _ in range(n)]


def config1():
    """4x4 dense Float64 1000x1000 on a [2,2] grid of 4 ranks (the reference's test/benchmark shape).
    128 MB of blocks: L2-resident, latency-bound -> quoted in applies/s."""
    ndev = int(os.environ.get("DJETS_NDEV", "1"))          # 4: one GPU per rank (the Julia-master model)
    ctx = dj.init(world=4, devices=[r % ndev for r in range(4)])
    rng = np.random.default_rng(20241)
    blocks = {(i, j): np.asfortranarray(rng.random((1000, 1000))) for i in range(1, 5) for j in range(1, 5)}
    A = dj.JopDBlock(lambda i, j: dj.JopDense(blocks[(i, j)]), dims=(4, 4), pids=[0, 1, 2, 3], dist=[2, 2])
    x, y = A.domain().rand(rng), A.range().rand(rng)
    yo, xo = A.range().zeros(), A.domain().zeros()
    # full parity against float128-free ground truth (long double on the host)
    Ax = (A @ x).to_array()
    full = np.block([[blocks[(i, j)] for j in range(1, 5)] for i in range(1, 5)]).astype(np.longdouble)
    truth = full @ x.to_array().astype(np.longdouble)
    err = float(np.max([np.linalg.norm((Ax[k * 1000:(k + 1) * 1000] - truth[k * 1000:(k + 1) * 1000]).astype(np.float64)) /
                        np.linalg.norm(truth[k * 1000:(k + 1) * 1000].astype(np.float64)) for k in range(4)]))
    tf = timed(ctx, lambda: A.mul(yo, x), 200)
    ta = timed(ctx, lambda: A.mul_adjoint(xo, y), 200)
    nbytes = A.algorithmic_bytes()
    return {"config": 1, "workload": f"4x4 dense Float64 1000x1000, grid [2,2], 4 ranks on {ndev} GPU(s), one host process", "bytes_per_apply": nbytes,
            "forward_us": tf * 1e3, "adjoint_us": ta * 1e3, "applies_per_s": 2.0 / ((tf + ta) * 1e-3),
            "GBps_forward": nbytes / tf / 1e6, "GBps_adjoint": nbytes / ta / 1e6, "note": "L2-resident (128 MB < 126 MB L2 + reuse): latency case",
            "relerr_forward_vs_longdouble": err, "dot_test": dot_test(A, rng)}


def config3():
    """Full 32x32 dense Float64 2048x2048 (32 GiB) on one GPU, grid [1,1]; --small: 8x8."""
    nb = 8 if SMALL else 32
    ctx = dj.init(world=1, devices=[0])
    rng = np.random.default_rng(20243)
    pool = pool_blocks(rng, 8, 2048, 2048, np.float64)
    t0 = time.perf_counter()
    A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[(3 * i + 5 * j) % 8]), dims=(nb, nb), pids=[0], dist=[1, 1],
                     shapes=([(2048,)] * nb, [(2048,)] * nb), dtype=np.float64)
    build = time.perf_counter() - t0
    x, y = A.domain().rand(rng), A.range().rand(rng)
    yo, xo = A.range().zeros(), A.domain().zeros()
    # parity on block row 1 and block column 1 against long double
    Ax = (A @ x).getblock(1)
    xs = x.to_array()
    truth = sum(pool[(3 * 1 + 5 * j) % 8].astype(np.longdouble) @ xs[(j - 1) * 2048:j * 2048].astype(np.longdouble)
                for j in range(1, nb + 1))
    err_f = float(np.linalg.norm((Ax - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64)))
    Aty = (A.T @ y).getblock(1)
    ys = y.to_array()
    truth = sum(pool[(3 * i + 5 * 1) % 8].astype(np.longdouble).T @ ys[(i - 1) * 2048:i * 2048].astype(np.longdouble)
                for i in range(1, nb + 1))
    err_a = float(np.linalg.norm((Aty - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64)))
    tf = timed(ctx, lambda: A.mul(yo, x), 10)
    ta = timed(ctx, lambda: A.mul_adjoint(xo, y), 10)
    nbytes = A.algorithmic_bytes()
    return {"config": 3, "workload": f"full {nb}x{nb} dense Float64 2048x2048, one GPU", "bytes_per_apply": nbytes,
            "forward_ms": tf, "adjoint_ms": ta, "GBps_forward": nbytes / tf / 1e6, "GBps_adjoint": nbytes / ta / 1e6,
            "frac_of_measured_peak": [nbytes / tf / 1e6 / PEAK, nbytes / ta / 1e6 / PEAK],
            "relerr_blockrow1_vs_longdouble": err_f, "relerr_blockcol1_vs_longdouble": err_a, "dot_test": dot_test(A, rng),
            "build_s": build, "schedule": [A.schedule_info(False), A.schedule_info(True)]}


def config4():
    """DBArray vector algebra on a 2^30-element Float64 block vector (1024 blocks of 2^20); --small: 2^26."""
    nb, bl = (64, 1 << 20) if SMALL else (1024, 1 << 20)
    ctx = dj.init(world=1, devices=[0])
    n = nb * bl
    ramp = np.arange(bl, dtype=np.float64) / bl
    x = dj.DBArray(lambda i: ramp + i, (nb,), [0])          # block i = i + k/bl
    y = x.similar().fill(2.0)
    z = x.similar().fill(-1.0)
    out = x.similar()
    res = {"config": 4, "workload": f"DBArray algebra, {n} Float64 elements in {nb} blocks, one GPU", "n": n}
    # known answers (exact in double for these sizes)
    s_ramp = float(np.sum(ramp))
    want_sum = bl * nb * (nb + 1) / 2 + nb * s_ramp
    res["sum_relerr"] = abs(float(x.sum()) - want_sum) / want_sum
    res["dot_relerr"] = abs(float(x.dot(y)) - 2 * want_sum) / (2 * want_sum)
    res["norm_inf_ok"] = bool(float(x.norm(math.inf)) == float(nb + ramp[-1]) and float(y.norm(2)) == math.sqrt(4.0 * n))
    out.lincomb_([0.5, 1.0], [x, y])                          # y .= a.*x .+ y shape
    res["axpy_ok"] = bool(out[1] == 0.5 * 1.0 + 2.0 and out[n] == 0.5 * (nb + ramp[-1]) + 2.0)
    t = {}
    t["dot"] = (timed(ctx, lambda: x.dot(y), 10), 2 * n * 8)
    t["norm2"] = (timed(ctx, lambda: x.norm(2), 10), n * 8)
    t["axpy"] = (timed(ctx, lambda: y.lincomb_([0.5, 1.0], [x, y]), 10), 3 * n * 8)
    t["sum3"] = (timed(ctx, lambda: out.lincomb_([0.1, 0.2, 0.3], [x, y, z]), 10), 4 * n * 8)
    t["fill"] = (timed(ctx, lambda: out.fill(3.14), 10), n * 8)
    for k, (ms, nbytes) in t.items():
        res[k] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_measured_peak": nbytes / ms / 1e6 / PEAK}
    return res


def config5():
    """CGLS on a 128x128 mixed dense/diagonal Float32 operator, 1024x1024 blocks, diagonal where
    iseven(i) && isodd(j) (the reference's own mix, benchmark/benchmarks.jl:83-89); --small: 16x16."""
    nb = 16 if SMALL else 128
    bs = 1024
    ctx = dj.init(world=1, devices=[0])
    rng = np.random.default_rng(20245)
    pool = pool_blocks(rng, 16, bs, bs, np.float32)
    dpool = [rng.random(bs, dtype=np.float32) for _ in range(16)]
    isdiag = lambda i, j: (i % 2 == 0) and (j % 2 == 1)      # noqa: E731
    t0 = time.perf_counter()
    A = dj.JopDBlock(lambda i, j: dj.JopDiagonal(dpool[(i + 3 * j) % 16]) if isdiag(i, j) else dj.JopDense(pool[(5 * i + j) % 16]),
                     dims=(nb, nb), pids=[0], dist=[1, 1], shapes=([(bs,)] * nb, [(bs,)] * nb), dtype=np.float32)
    build = time.perf_counter() - t0
    nbytes = A.algorithmic_bytes()
    x_true = A.domain().rand(rng)
    b = A @ x_true
    # CGLS: min ||A x - b||
    x = A.domain().zeros()
    r = b.copy()
    s = A.T @ r
    p = s.copy()
    q = A.range().zeros()
    gamma = float(s.dot(s))
    r0 = float(r.norm())

    def iteration():
        nonlocal gamma
        A.mul(q, p)                                           # q = A p
        alpha = gamma / float(q.dot(q))
        x.lincomb_([1.0, alpha], [x, p])                      # x .+= alpha .* p
        r.lincomb_([1.0, -alpha], [r, q])                     # r .-= alpha .* q
        A.mul_adjoint(s, r)                                   # s = A' r
        g2 = float(s.dot(s))
        p.lincomb_([1.0, g2 / gamma], [s, p])                 # p .= s .+ beta .* p
        gamma = g2

    iters = 10
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        iteration()
    ctx.sync()
    per_iter = (time.perf_counter() - t0) / iters
    tf = timed(ctx, lambda: A.mul(q, p), 5)
    ta = timed(ctx, lambda: A.mul_adjoint(s, r), 5)
    return {"config": 5, "workload": f"CGLS on {nb}x{nb} mixed dense/diagonal Float32 {bs}x{bs}, one GPU",
            "bytes_per_apply": nbytes, "forward_ms": tf, "adjoint_ms": ta, "GBps_forward": nbytes / tf / 1e6,
            "GBps_adjoint": nbytes / ta / 1e6, "cgls_iteration_ms": per_iter * 1e3,
            "cgls_GBps": 2 * nbytes / per_iter / 1e9, "residual_drop_after_10_iters": float(r.norm()) / r0,
            "dot_test": dot_test(A, rng), "build_s": build}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a.isdigit()] or ["1", "3", "4", "5"]
    for w in which:
        res = {"1": config1, "3": config3, "4": config4, "5": config5}[w]()
        print(json.dumps(res), flush=True)
