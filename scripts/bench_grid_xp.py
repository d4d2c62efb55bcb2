"""A/B of the cross-process exchange mechanisms on BASELINE configs[2] inside ONE job (boxes differ by a few per cent,
so settings are alternated over several rounds in the same processes):
    torchrun --nproc-per-node 8 scripts/bench_grid_xp.py [n1 n2]
Operators: one finalised with the peer-memory exchange (ipc=1), one with ncclSend/ncclRecv (ipc=0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
box = [dj.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx = dj.Context.spmd(rank, world, lr, box[0])
dj.set_default_context(ctx)
grids = [[int(sys.argv[1]), int(sys.argv[2])]] if len(sys.argv) > 2 else [[2, world // 2], [1, world]]
nb, bs = 32, 2048
rng = np.random.default_rng(1)
pool = [np.asfortranarray(rng.random((bs, bs))) for _ in range(4)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ctx.sync(); dist.barrier()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop() / reps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for grid in grids:
    ops = {}
    for ipc in (1, 0):
        ctx.set_param("ipc", ipc)
        ops[ipc] = dj.JopDBlock(lambda i, j: dj.JopDense(pool[(i + 3 * j) % 4]), dims=(nb, nb), pids=list(range(world)), dist=grid,
                                shapes=([(bs,)] * nb, [(bs,)] * nb), dtype=np.float64, ctx=ctx)
    A = ops[1]
    x, y = A.domain().rand(np.random.default_rng(2)), A.range().rand(np.random.default_rng(3))
    yo, xo = A.range().zeros(), A.domain().zeros()
    nbytes = torch.tensor([float(A.algorithmic_bytes())], dtype=torch.float64, device="cuda")
    dist.all_reduce(nbytes)
    nbytes = float(nbytes.item())
    settings = [("nccl", 0, 1, 0, 1), ("ipc push kernel", 1, 1, 0, 1), ("ipc push kernel, side stream", 1, 1, 1, 1),
                ("ipc peer copies", 1, 1, 0, 0), ("ipc push kernel, wait kernels", 1, 0, 0, 1)]
    best = {}
    for rnd in range(3):
        for name, ipc, kwait, side, pushk in settings:
            ctx.set_param("xp_kwait", kwait); ctx.set_param("xp_side", side); ctx.set_param("xp_push", pushk)
            op = ops[ipc]
            f = timed(lambda: op.mul(yo, x))
            a = timed(lambda: op.mul_adjoint(xo, y))
            b = best.setdefault(name, [1e9, 1e9])
            b[0], b[1] = min(b[0], f), min(b[1], a)
    if rank == 0:
        for name, (f, a) in best.items():
            print(json.dumps({"grid": grid, "exchange": name, "fwd_ms": round(f, 4), "adj_ms": round(a, 4),
                              "fwd_GBps_per_gpu": round(nbytes / f / 1e6 / world), "adj_GBps_per_gpu": round(nbytes / a / 1e6 / world)}), flush=True)
    for op in ops.values():
        op.close()
    del ops, A, x, y, yo, xo
dist.barrier()
ctx.close()
dist.destroy_process_group()
