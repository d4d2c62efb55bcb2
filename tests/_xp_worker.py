"""One process per rank WITHOUT NCCL: the control plane is a host all-gather (gloo), block data moves through
CUDA-IPC-mapped peer buffers guarded by epoch flags (csrc/djets_op.cu: apply_enqueue_xp).  One GPU per rank.  Launched by tests/test_gpu_xp.py:
    RANK=r WORLD_SIZE=4 MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/_xp_worker.py [ipc]
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
    flat = np.zeros(length, dtype=vec.dtype)
    vec.to_array(flat)                                     # fills the locally owned parts
    t = torch.from_numpy(flat)
    dist.all_reduce(t)
    return t.numpy()


def load(gvec, ovec):
    ib = 1
    for part in ovec.parts:
        for a in part.arrays:
            gvec.setblock(ib, a)                           # no-op on processes that do not own the block
            ib += 1
    return gvec


def main():
    import faulthandler
    faulthandler.dump_traceback_later(120, exit=True)      # a hung exchange must not hold the GPU
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    dist.init_process_group("gloo")
    assert ngpu >= world, "one GPU per rank"
    ctx = dj.Context.spmd(rank, world, rank, nccl_uid=None, allgather=dj.torch_allgather())
    dj.set_default_context(ctx)
    for name in ("xp_push", "xp_side", "xp_kwait"):                # A/B of the exchange mechanisms (scripts, debugging)
        if os.environ.get("DJETS_" + name.upper()):
            ctx.set_param(name, int(os.environ["DJETS_" + name.upper()]))
    grids = [(world, 1), (1, world)] + ([(2, world // 2)] if world >= 4 else [])
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for n1, n2 in grids:
            rng = np.random.default_rng(23)
            kind = lambda i, j: "diag" if (i == 2 and j == 3) else ("zero" if (i, j) == (1, 1) else "dense")  # noqa
            rs, cs = [300, 200, 257, 64, 129], [200, 128, 200, 300, 31, 77]
            rs = (rs * world)[:max(len(rs), world)]
            cs = (cs * world)[:max(len(cs), world)]
            oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, dtype, list(range(world)), [n1, n2], ctx=ctx)
            xo, do = oop.dom.rand(rng), oop.rng.rand(rng)
            xg, dg = load(gop.domain().zeros(), xo), load(gop.range().zeros(), do)
            y, m = gop.range().zeros(), gop.domain().zeros()
            truth_f = O.ground_truth_mul(oblocks, xo.to_array())
            truth_a = O.ground_truth_mul(oblocks, do.to_array(), adjoint=True)
            # many applies in flight without a host synchronisation: producers run ahead of consumers, the epoch flags
            # and the double-buffered receive planes must keep every apply's data apart
            for rep in range(7):
                gop.mul(y, xg)
                gop.mul_adjoint(m, dg)
            ctx.sync()
            assert O.blockwise_relerr(gather(y, sum(rs)), truth_f, rs) <= tol, (n1, n2, "fwd")
            assert O.blockwise_relerr(gather(m, sum(cs)), truth_a, cs) <= tol, (n1, n2, "adj")
            # a chain whose applies depend on each other across the processes: y = A x; m = A' y; y2 = A m
            gop.mul(y, xg)
            gop.mul_adjoint(m, y)
            y2 = gop.mul(gop.range().zeros(), m)
            ctx.sync()
            want = O.ground_truth_mul(oblocks, O.ground_truth_mul(oblocks, truth_f.astype(dtype), adjoint=True).astype(dtype))
            got = gather(y2, sum(rs))
            assert np.linalg.norm(got - want) <= 50 * tol * np.linalg.norm(want), (n1, n2, "chain")
            # scalars over the host control plane, host-returning and device-resident
            wd = do.to_array().astype(np.float64)
            want = float(np.dot(wd, wd))
            assert abs(float(dg.dot(dg)) - want) <= 1e-5 * want
            assert abs(float(dg.norm()) - want ** 0.5) <= 1e-5 * want ** 0.5
            assert dg.maximum() == do.extrema()[1]
            assert dg.dot_s(dg).get() == dg.dot(dg)
            gop.close()
    dist.barrier()
    if rank == 0:
        print("xp ok", world, flush=True)
    ctx.close()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
