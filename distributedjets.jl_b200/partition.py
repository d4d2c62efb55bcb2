"""Partition rules of the reference, host-side and GPU-free.

The reference inherits these from DistributedArrays 0.6 (``defaultdist``, ``chunk_idxs``) and extends
them with caller-prescribed ranges (src/DistributedJets.jl:8-43).  Indices here are 1-based and
inclusive like Julia's; a Julia ``a:b`` is the Python ``range(a, b + 1)``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

from . import _lib


def urange(a: int, b: int) -> range:
    """Julia ``a:b``."""
    return range(int(a), int(b) + 1)


def even_cuts(n: int, nparts: int) -> List[int]:
    """0-based starts of an even split (first ``n % nparts`` parts get one extra); computed by the
    library so the Julia shim, this mirror and the scheduler agree (include/djets.h)."""
    out = (C.c_int64 * (nparts + 1))()
    _lib.call("djets_even_cuts", n, nparts, out)
    return list(out)


def even_ranges(n: int, nparts: int) -> List[range]:
    """1-based inclusive block ranges of an even split, e.g. 7 over 4 -> 1:2, 3:4, 5:6, 7:7
    (test/runtests.jl:172-221)."""
    cuts = even_cuts(n, nparts)
    return [urange(cuts[k] + 1, cuts[k + 1]) for k in range(nparts)]


def _prime_factors_desc(n: int) -> List[int]:
    out, p = [], 2
    while p * p <= n:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 1
    if n > 1:
        out.append(n)
    return sorted(out, reverse=True)


def default_grid(dims: Sequence[int], npids: int) -> List[int]:
    """``DistributedArrays.defaultdist(dims, pids)``: number of chunks per dimension when the caller
    gives no ``dist`` -- prime factors of the worker count, largest first, each handed to the
    currently largest remaining dimension (ties go to the last dimension)."""
    rem = [int(d) for d in dims]
    chunks = [1] * len(rem)
    for fac in _prime_factors_desc(int(npids)):
        best = max(rem)
        where = max(k for k, v in enumerate(rem) if v == best)
        if rem[where] >= fac:
            rem[where] //= fac
            chunks[where] *= fac
    return chunks


def cuts_from_ranges(rngs: Sequence[range]) -> List[int]:
    """0-based cuts of caller-prescribed 1-based ranges (src/DistributedJets.jl:30-38)."""
    cuts = [r[0] - 1 for r in rngs]
    cuts.append(rngs[-1][-1])
    for k, r in enumerate(rngs):
        if len(r) == 0 or r[0] - 1 != cuts[k] or r[-1] != cuts[k + 1]:
            raise ValueError("prescribed ranges must be non-empty, consecutive and cover the dimension")
    return cuts


def ranges_from_cuts(cuts: Sequence[int]) -> List[range]:
    return [urange(cuts[k] + 1, cuts[k + 1]) for k in range(len(cuts) - 1)]


def grid_of(pids: Sequence[int], n1: int, n2: int) -> List[List[int]]:
    """First n1*n2 pids reshaped column-major to n1 x n2 (src/DistributedJets.jl:12,483; pinned by
    test/runtests.jl:543)."""
    if len(pids) < n1 * n2:
        raise ValueError(f"need {n1 * n2} workers for a {n1}x{n2} grid, have {len(pids)}")
    return [[pids[r + c * n1] for c in range(n2)] for r in range(n1)]


def plan_exchange(n1: int, n2: int, isdiag: bool, adjoint: bool, phase: int) -> List[Tuple[int, int, int, int, int, int]]:
    """The messages of one apply as (src_r, src_c, dst_r, dst_c, part, slot), 0-based grid
    coordinates: exactly the list the library executes (include/djets.h djets_plan_exchange)."""
    n = C.c_int32()
    _lib.call("djets_plan_exchange", n1, n2, int(isdiag), int(adjoint), phase, None, 0, C.byref(n))
    buf = (C.c_int32 * max(6 * n.value, 1))()
    _lib.call("djets_plan_exchange", n1, n2, int(isdiag), int(adjoint), phase, buf, n.value, C.byref(n))
    return [tuple(buf[6 * k:6 * k + 6]) for k in range(n.value)]
