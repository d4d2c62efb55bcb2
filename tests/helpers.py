"""Shared builders: the same seeded blocks handed to the oracle and to the CUDA backend."""
import numpy as np


def build_pair(dj, O, rng, row_sizes, col_sizes, kind_of, dtype, pids, dist=None, isdiag=False, ctx=None,
               row_ranges=None, col_ranges=None):
    """kind_of(i, j) (1-based) -> 'dense' | 'diag' | 'zero'.  Returns (oracle op, gpu op, oracle blocks)."""
    N, M = len(row_sizes), len(col_sizes)
    oblocks, gblocks = [], []
    for i in range(1, N + 1):
        orow, grow = [], []
        for j in range(1, M + 1):
            k = kind_of(i, j)
            m, n = row_sizes[i - 1], col_sizes[j - 1]
            if k == "dense":
                A = np.asfortranarray(rng.random((m, n), dtype=dtype))
                orow.append(O.ODense(A))
                grow.append(dj.JopDense(A))
            elif k == "diag":
                assert m == n
                d = rng.random(n, dtype=dtype)
                orow.append(O.ODiag(d))
                grow.append(dj.JopDiagonal(d))
            else:
                orow.append(O.OZero(dtype, (m,), (n,)))
                grow.append(dj.JopZeroBlock(dtype, (m,), (n,)))
        oblocks.append(orow)
        gblocks.append(grow)
    oop = O.OJopDBlock(oblocks, pids, dist, isdiag=isdiag, row_ranges=row_ranges, col_ranges=col_ranges)
    gop = dj.JopDBlock(gblocks, pids=pids, dist=dist, isdiag=isdiag, row_ranges=row_ranges, col_ranges=col_ranges,
                       ctx=ctx)
    return oop, gop, oblocks


def load_like(gvec, ovec):
    """Copy an oracle ODBArray into a GPU DBArray of the same layout, block by block (setblock!)."""
    ib = 1
    for part in ovec.parts:
        for a in part.arrays:
            gvec.setblock(ib, a)
            ib += 1
    return gvec


def tol_for(dtype):
    # BASELINE.json north_star: <= 1e-12 relative in Float64, <= 1e-5 in Float32
    return 1e-12 if np.dtype(dtype) == np.float64 else 1e-5
