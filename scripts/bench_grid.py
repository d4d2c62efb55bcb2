"""BASELINE config 3 across GPUs: the full 32x32 dense Float64 2048x2048 operator on an n1 x n2 grid,
block columns sharded across GPUs (cross-rank exchange: input parts to the other grid rows, reduced
partials to the owners of the range parts).

  one process per GPU over NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29540 \
        scripts/bench_grid.py --grid 2 P/2
  one host process driving P GPUs (in-place peer loads / stores, CUDA-graph replay):
    python scripts/bench_grid.py --grid 2 2 --single-process
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, nargs=2, default=[1, 1])
ap.add_argument("--nb", type=int, default=32)
ap.add_argument("--bs", type=int, default=2048)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--single-process", action="store_true")
ap.add_argument("--cgls", action="store_true",
                help="BASELINE config 5 instead: 128x128 mixed dense/diagonal Float32 1024x1024 blocks, timed as CGLS "
                     "iterations (mul!, adjoint, 2 dots, 3 fused updates per iteration)")
args = ap.parse_args()
n1, n2 = args.grid
world = n1 * n2
rank = int(os.environ.get("RANK", "0"))
dist = None
if args.single_process:
    ctx = dj.Context(world=world, devices=list(range(world)))
else:
    import torch
    import torch.distributed as dist
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    assert dist.get_world_size() == world, "launch one process per grid cell"
    box = [dj.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = dj.Context.spmd(rank, world, lr, box[0])
dj.set_default_context(ctx)
nb, bs = args.nb, args.bs
rng = np.random.default_rng(20243)
if args.cgls:
    nb, bs = (args.nb if args.nb != 32 else 128), 1024
    dt = np.float32
    pool = [np.asfortranarray(rng.random((bs, bs), dtype=dt) / bs) for _ in range(8)]
    dpool = [1.0 + rng.random(bs, dtype=dt) for _ in range(8)]
    isd = lambda i, j: (i % 2 == 0) and (j % 2 == 1)       # noqa: E731  benchmark/benchmarks.jl:83-89
    A = dj.JopDBlock(lambda i, j: dj.JopDiagonal(dpool[(i + 3 * j) % 8]) if isd(i, j) else dj.JopDense(pool[(3 * i + 5 * j) % 8]),
                     dims=(nb, nb), pids=list(range(world)), dist=[n1, n2], shapes=([(bs,)] * nb, [(bs,)] * nb), dtype=dt,
                     ctx=ctx)
else:
    dt = np.float64
    pool = [np.asfortranarray(rng.random((bs, bs))) for _ in range(8)]
    A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[(3 * i + 5 * j) % 8]), dims=(nb, nb), pids=list(range(world)),
                     dist=[n1, n2], shapes=([(bs,)] * nb, [(bs,)] * nb), dtype=dt, ctx=ctx)
xs, ys = rng.random(nb * bs).astype(dt), rng.random(nb * bs).astype(dt)          # same on every process
x, y = A.domain().zeros().scatter(xs), A.range().zeros().scatter(ys)
yo, xo = A.range().zeros(), A.domain().zeros()


def barrier():
    ctx.sync()
    if dist is not None:
        dist.barrier()


def reduce_max(v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def timed(fn):
    for _ in range(3):
        fn()
    barrier()
    ctx.timer_start()
    for _ in range(args.reps):
        fn()
    ms = ctx.timer_stop() / args.reps
    barrier()
    return reduce_max(ms)


tf = timed(lambda: A.mul(yo, x))
ta = timed(lambda: A.mul_adjoint(xo, y))
nbytes = reduce_sum(float(A.algorithmic_bytes()))
if args.cgls:
    import time
    b = A.range().zeros()
    A.mul(b, x)                                            # consistent right-hand side
    sol = A.domain().zeros()
    r = b.copy()
    s_ = A.domain().zeros()
    A.mul_adjoint(s_, r)
    p = s_.copy()
    q = A.range().zeros()
    gamma = float(s_.dot(s_))
    r0 = float(r.norm())
    iters = 10
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        A.mul(q, p)
        alpha = gamma / float(q.dot(q))
        sol.lincomb_([1.0, alpha], [sol, p])
        r.lincomb_([1.0, -alpha], [r, q])
        A.mul_adjoint(s_, r)
        g2 = float(s_.dot(s_))
        p.lincomb_([1.0, g2 / gamma], [s_, p])
        gamma = g2
    barrier()
    per_iter = reduce_max((time.perf_counter() - t0) / iters)
    drop = float(r.norm()) / r0
    # the same iteration with device-resident scalars: no host round trip (the dots are all-reduced in stream order)
    L = dj.lazy
    SC = sys.modules["djets_b200._lib"]
    g, g2s, dl, al, be = (dj.Scalar(1.0, dt, ctx) for _ in range(5))
    s_.dot_s(s_, g)

    def device_iter():
        A.mul(q, p)
        q.dot_s(q, dl)
        al.assign(SC.SC_DIV, g, dl)
        sol.assign(L(sol) + al * L(p))
        r.assign(L(r) - al * L(q))
        A.mul_adjoint(s_, r)
        s_.dot_s(s_, g2s)
        be.assign(SC.SC_DIV, g2s, g)
        p.assign(L(s_) + be * L(p))
        g.assign(SC.SC_COPY, g2s)
    for _ in range(3):
        device_iter()
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        device_iter()
    barrier()
    per_iter_dev = reduce_max((time.perf_counter() - t0) / iters)
    if rank == 0:
        print(json.dumps({"config": 5, "grid": [n1, n2], "n_gpus": world, "mode": "one process" if dist is None else "spmd+nccl",
                          "blocks": [nb, nb], "block": [bs, bs], "bytes_per_apply": nbytes, "forward_ms": tf, "adjoint_ms": ta,
                          "GBps_forward": nbytes / tf / 1e6, "GBps_adjoint": nbytes / ta / 1e6,
                          "cgls_iteration_ms": per_iter * 1e3, "cgls_GBps": 2 * nbytes / per_iter / 1e9,
                          "cgls_iteration_device_scalars_ms": per_iter_dev * 1e3,
                          "cgls_device_scalars_GBps": 2 * nbytes / per_iter_dev / 1e9,
                          "residual_drop_after_10_iters": drop}), flush=True)
    if dist is not None:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    sys.exit(0)
# parity on block row 1 / block column 1 (owned by rank 0) against long double
err_f = err_a = None
if rank == 0:
    got = yo.getblock(1)
    truth = sum(pool[(3 * 1 + 5 * j) % 8].astype(np.longdouble) @ xs[(j - 1) * bs:j * bs].astype(np.longdouble)
                for j in range(1, nb + 1))
    err_f = float(np.linalg.norm((got - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64)))
    got = xo.getblock(1)
    truth = sum(pool[(3 * i + 5 * 1) % 8].astype(np.longdouble).T @ ys[(i - 1) * bs:i * bs].astype(np.longdouble)
                for i in range(1, nb + 1))
    err_a = float(np.linalg.norm((got - truth).astype(np.float64)) / np.linalg.norm(truth.astype(np.float64)))
    print(json.dumps({"config": 3, "grid": [n1, n2], "n_gpus": world, "mode": "one process" if dist is None else "spmd+nccl",
                      "blocks": [nb, nb], "block": [bs, bs], "bytes_per_apply": nbytes, "forward_ms": tf, "adjoint_ms": ta,
                      "GBps_forward": nbytes / tf / 1e6, "GBps_adjoint": nbytes / ta / 1e6,
                      "GBps_per_gpu": [nbytes / tf / 1e6 / world, nbytes / ta / 1e6 / world],
                      "relerr_blockrow1": err_f, "relerr_blockcol1": err_a}), flush=True)
if dist is not None:
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
