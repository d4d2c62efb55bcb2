"""Worker of tests/test_multirank_gloo.py: two CPU processes execute the library's exchange plan
(djets_plan_exchange) with torch.distributed/gloo point-to-point messages, computing each grid cell's
partial with the oracle's block arithmetic, and check the assembled result against the oracle's
distributed operator.  This covers the host logic of the N>1 path: ownership, message list, receive
slots, owner-side summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import djets_loader  # noqa: E402
from oracle import djets_oracle as O  # noqa: E402

dj = djets_loader.load()


def run_case(rank, world, n1, n2, rs, cs, adjoint, seed):
    rng = np.random.default_rng(seed)                      # same stream on every process
    blocks = [[O.ODense(np.asfortranarray(rng.random((m, n)))) for n in cs] for m in rs]
    vin = rng.random(sum(rs) if adjoint else sum(cs))
    rcuts, ccuts = dj.even_cuts(len(rs), n1), dj.even_cuts(len(cs), n2)
    grid = dj.grid_of(list(range(n1 * n2)), n1, n2)        # cell -> rank
    proc_of = lambda r, c: grid[r][c] % world              # noqa: E731  rank -> process
    roff, coff = np.cumsum([0] + rs), np.cumsum([0] + cs)
    in_cuts, in_off = (rcuts, roff) if adjoint else (ccuts, coff)
    out_cuts, out_off = (ccuts, coff) if adjoint else (rcuts, roff)

    def part_slice(cuts, off, p):
        return slice(off[cuts[p]], off[cuts[p + 1]])

    # phase 0: input parts travel from their owner cell to the cells that need them
    have = {}
    for r in range(n1):
        for c in range(n2):
            owner = (c == 0) if adjoint else (r == 0)
            if owner and proc_of(r, c) == rank:
                p = r if adjoint else c
                have[(r, c)] = vin[part_slice(in_cuts, in_off, p)].copy()
    for sr, sc, dr, dc, part, _ in dj.plan_exchange(n1, n2, False, adjoint, 0):
        ps, pd = proc_of(sr, sc), proc_of(dr, dc)
        n = in_off[in_cuts[part + 1]] - in_off[in_cuts[part]]
        if ps == rank and pd == rank:
            have[(dr, dc)] = have[(sr, sc)].copy()
        elif ps == rank:
            dist.send(torch.from_numpy(have[(sr, sc)].copy()), dst=pd)
        elif pd == rank:
            buf = torch.empty(int(n), dtype=torch.float64)
            dist.recv(buf, src=ps)
            have[(dr, dc)] = buf.numpy()

    # local compute: every cell's partial of its output part
    partial = {}
    for r in range(n1):
        for c in range(n2):
            if proc_of(r, c) != rank:
                continue
            x = have[(r, c)]
            op_part = c if adjoint else r
            out = np.zeros(out_off[out_cuts[op_part + 1]] - out_off[out_cuts[op_part]])
            for i in range(rcuts[r], rcuts[r + 1]):
                for j in range(ccuts[c], ccuts[c + 1]):
                    if adjoint:
                        xo = roff[i] - roff[rcuts[r]]
                        oo = coff[j] - coff[ccuts[c]]
                        out[oo:oo + cs[j]] += blocks[i][j].mul_adjoint(x[xo:xo + rs[i]])
                    else:
                        xo = coff[j] - coff[ccuts[c]]
                        oo = roff[i] - roff[rcuts[r]]
                        out[oo:oo + rs[i]] += blocks[i][j].mul(x[xo:xo + cs[j]])
            partial[(r, c)] = out

    # phase 1: partials to the owners, summed in slot order
    slots = {}
    for sr, sc, dr, dc, part, slot in dj.plan_exchange(n1, n2, False, adjoint, 1):
        ps, pd = proc_of(sr, sc), proc_of(dr, dc)
        if ps == rank and pd == rank:
            slots[(dr, dc, slot)] = partial[(sr, sc)]
        elif ps == rank:
            dist.send(torch.from_numpy(partial[(sr, sc)].copy()), dst=pd)
        elif pd == rank:
            buf = torch.empty(partial[(dr, dc)].size, dtype=torch.float64)
            dist.recv(buf, src=ps)
            slots[(dr, dc, slot)] = buf.numpy()
    result = np.zeros(sum(cs) if adjoint else sum(rs))
    nslots = (n1 if adjoint else n2) - 1
    for r in range(n1):
        for c in range(n2):
            owner = (r == 0) if adjoint else (c == 0)
            if not owner or proc_of(r, c) != rank:
                continue
            acc = partial[(r, c)].copy()
            for q in range(nslots):
                acc += slots[(r, c, q)]
            p = c if adjoint else r
            result[part_slice(out_cuts, out_off, p)] = acc
    t = torch.from_numpy(result)
    dist.all_reduce(t)                                     # parts are disjoint: the sum assembles them
    # the oracle's own distributed operator on the same grid
    oop = O.OJopDBlock(blocks, list(range(n1 * n2)), [n1, n2])
    if adjoint:
        d = oop.rng.zeros()
        for ib in range(len(rs)):
            d.setblock(ib + 1, vin[roff[ib]:roff[ib + 1]])
        want = oop.mul_adjoint(oop.dom.zeros(), d).to_array()
    else:
        m = oop.dom.zeros()
        for ib in range(len(cs)):
            m.setblock(ib + 1, vin[coff[ib]:coff[ib + 1]])
        want = oop.mul(oop.rng.zeros(), m).to_array()
    np.testing.assert_allclose(t.numpy(), want, rtol=1e-13, atol=0)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank,
                            world_size=world)
    seed = 100
    for n1, n2 in ((2, 1), (1, 2), (2, 2), (2, 3), (4, 2)):
        for adjoint in (False, True):
            run_case(rank, world, n1, n2, [5, 7, 3, 6, 4], [9, 2, 6, 8, 3, 5], adjoint, seed)
            seed += 1
    # strong-scaling shard of the benchmark's block-diagonal operator: disjoint cover of the 64 blocks
    mine = dj.even_ranges(64, world)[rank]
    cover = torch.zeros(64)
    cover[mine.start - 1:mine.stop - 1] = 1
    dist.all_reduce(cover)
    assert bool((cover == 1).all())
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
