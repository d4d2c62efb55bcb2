"""BASELINE config 4 across GPUs: DBArray vector algebra on a 2^30-element Float64 block vector
(1024 blocks of 2^20), one process per GPU over NCCL (scalars are all-reduced).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/bench_vec.py
"""
import json
import math
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
nb, bl = 1024, 1 << 20
n = nb * bl
ramp = np.arange(bl, dtype=np.float64) / bl
x = dj.DBArray(lambda i: ramp + i, (nb,), list(range(world)))
y = x.similar().fill(2.0)
z = x.similar().fill(-1.0)
out = x.similar()


def timed(fn, reps=10):
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


want_sum = bl * nb * (nb + 1) / 2 + nb * float(np.sum(ramp))
res = {"config": 4, "n_gpus": world, "n": n,
       "sum_relerr": abs(float(x.sum()) - want_sum) / want_sum,
       "dot_relerr": abs(float(x.dot(y)) - 2 * want_sum) / (2 * want_sum),
       "norm_ok": bool(float(x.norm(math.inf)) == float(nb + ramp[-1]) and float(y.norm(2)) == math.sqrt(4.0 * n))}
sc = dj.Scalar(0.0, np.float64, ctx)
L = dj.lazy
for name, fn, nbytes in (("dot", lambda: x.dot(y), 2 * n * 8), ("norm2", lambda: x.norm(2), n * 8),
                         ("norm2 again", lambda: x.norm(2), n * 8), ("dot again", lambda: x.dot(y), 2 * n * 8),
                         ("dot_s (device-resident, all-reduced in stream order)", lambda: x.dot_s(y, sc), 2 * n * 8),
                         ("norm_s", lambda: x.norm_s(2, sc), n * 8),
                         ("eval abs.(x) .* y .+ 1", lambda: out.assign(abs(L(x)) * L(y) + 1.0), 3 * n * 8),
                         ("axpy", lambda: y.lincomb_([0.5, 1.0], [x, y]), 3 * n * 8),
                         ("sum3", lambda: out.lincomb_([0.1, 0.2, 0.3], [x, y, z]), 4 * n * 8)):
    ms = timed(fn)
    res[name] = {"ms": ms, "GBps": nbytes / ms / 1e6}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
