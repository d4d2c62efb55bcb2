"""BASELINE config 1 (4x4 dense Float64 1000x1000, grid [2,2]) with one PROCESS per GPU: the deployment model, where
every rank enqueues only its own cell.  Peer-memory exchange (ipc=1) against ncclSend/ncclRecv (ipc=0).
    torchrun --nproc-per-node 4 scripts/bench_latency_spmd.py"""
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
rng = np.random.default_rng(0)
payload = [[np.asfortranarray(rng.random((1000, 1000))) for _ in range(4)] for _ in range(4)]


def timed(fn, reps=300):
    for _ in range(30):
        fn()
    ctx.sync(); dist.barrier()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    us = ctx.timer_stop() / reps * 1e3
    t = torch.tensor([us], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for ipc in (1, 0):
    ctx.set_param("ipc", ipc)
    A = dj.JopDBlock([[dj.JopDense(payload[i][j]) for j in range(4)] for i in range(4)], pids=list(range(world)),
                     dist=[2, world // 2], ctx=ctx)
    x, y = A.domain().rand(np.random.default_rng(2)), A.range().rand(np.random.default_rng(3))
    yo, xo = A.range().zeros(), A.domain().zeros()
    f = timed(lambda: A.mul(yo, x))
    a = timed(lambda: A.mul_adjoint(xo, y))

    def pair():
        A.mul(yo, x)
        A.mul_adjoint(xo, yo)

    p = timed(pair)
    lhs, rhs = float(A.mul(A.range().zeros(), x).dot(y)), float(x.dot(A.mul_adjoint(A.domain().zeros(), y)))
    if rank == 0:
        print(json.dumps({"case": "config1, one process per GPU", "gpus": world, "grid": [2, world // 2],
                          "exchange": "peer memory (CUDA IPC)" if ipc else "ncclSend/ncclRecv", "fwd_us": round(f, 2),
                          "adj_us": round(a, 2), "pair_us": round(p, 2), "fwd_applies_per_s": round(1e6 / f),
                          "adj_applies_per_s": round(1e6 / a), "dot_test": abs(lhs - rhs) / abs(lhs)}), flush=True)
    A.close()
    del A, x, y, yo, xo
ctx.close()
dist.destroy_process_group()
