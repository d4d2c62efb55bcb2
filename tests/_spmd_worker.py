"""One process per GPU (torchrun): the operator's cross-rank exchange over NCCL, checked against the
oracle.  Launched by tests/test_gpu_spmd.py when the box has >= 2 GPUs, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/_spmd_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import djets_loader  # noqa: E402
from oracle import djets_oracle as O  # noqa: E402
from helpers import build_pair  # noqa: E402

dj = djets_loader.load()


def gather(vec, length):
    """Assemble a DBArray whose parts live in different processes (test-side only)."""
    flat = np.zeros(length, dtype=vec.dtype)
    vec.to_array(flat)                                     # fills the locally owned parts
    t = torch.from_numpy(flat).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def log(*a):
    print(f"[rank {os.environ.get('RANK')}]", *a, file=sys.stderr, flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(90, exit=True)       # a hung exchange must not hold the GPUs
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    box = [dj.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    log("uid exchanged")
    ctx = dj.Context.spmd(rank, world, lr, box[0])
    dj.set_default_context(ctx)
    log("context up")
    grids = [(world, 1), (1, world)] + ([(2, world // 2)] if world >= 4 else [])
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for n1, n2 in grids:
            rng = np.random.default_rng(17)
            kind = lambda i, j: "diag" if (i == 2 and j == 3) else ("zero" if (i, j) == (1, 1) else "dense")  # noqa
            rs, cs = [300, 200, 257, 64, 129], [200, 128, 200, 300, 31, 77]
            rs = (rs * world)[:max(len(rs), world)]             # at least one block row / column per rank
            cs = (cs * world)[:max(len(cs), world)]
            oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, dtype, list(range(world)), [n1, n2], ctx=ctx)
            xo = oop.dom.rand(rng)
            xg = gop.domain().zeros()
            ib = 1
            for part in xo.parts:
                for a in part.arrays:
                    xg.setblock(ib, a)                     # no-op on processes that do not own the block
                    ib += 1
            log("grid", n1, n2, np.dtype(dtype).name, "operator built")
            y = gop.mul(gop.range().zeros(), xg)
            ctx.sync()
            log("forward done")
            got = gather(y, sum(rs))
            err = O.blockwise_relerr(got, O.ground_truth_mul(oblocks, xo.to_array()), rs)
            assert err <= tol, (n1, n2, "fwd", err)
            do = oop.rng.rand(rng)
            dg = gop.range().zeros()
            ib = 1
            for part in do.parts:
                for a in part.arrays:
                    dg.setblock(ib, a)
                    ib += 1
            m = gop.mul_adjoint(gop.domain().zeros(), dg)
            ctx.sync()
            log("adjoint done")
            got = gather(m, sum(cs))
            err = O.blockwise_relerr(got, O.ground_truth_mul(oblocks, do.to_array(), adjoint=True), cs)
            assert err <= tol, (n1, n2, "adj", err)
            # scalars: per-rank partials + NCCL all-reduce, identical on every process
            want = float(np.dot(do.to_array().astype(np.float64), do.to_array().astype(np.float64)))
            assert abs(float(dg.dot(dg)) - want) <= 1e-5 * want
            assert abs(float(dg.norm()) - want ** 0.5) <= 1e-5 * want ** 0.5
            assert dg.maximum() == do.extrema()[1]
            log("scalars done")
            if rank != gop.range().procs()[0]:
                try:
                    dg.getblock(1)
                    raise AssertionError("getblock of a remote block must raise NotLocal")
                except dj.NotLocal:
                    pass
    dist.barrier()
    if rank == 0:
        print("spmd ok", world, flush=True)
    ctx.close()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
