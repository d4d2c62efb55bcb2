"""CPU oracle for the DistributedJets.jl hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the algorithm of the reference
(ChevronETC/DistributedJets.jl v1.2.1, one file: src/DistributedJets.jl, cited
below as ``src:LINE``; its tests test/runtests.jl as ``test:LINE``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and only as the checker.  The product (``distributedjets.jl_b200``)
never imports it and has no CPU fallback.

Pinning status
--------------
* Index/layout semantics (partition rules, block/element ranges, blockmaps,
  owner lookup, ``collect`` re-basing) are PINNED: ``tests/test_oracle_reference.py``
  checks them against every literal (``==``) assertion the reference test-suite
  holds for this path (test:40-52,59-62,68-77,83-92,113-119,130-133,152-162,
  193-200,381-385,411-414,533-551,568-571) and against the reference's
  known-answer values (pi round trips test:138-140,167-169; ``i*pi*ones``
  test:173-220; alpha sums / 3.14 fills test:323-333,356-361).
* Floating-point results of mul!/adjoint/dot/norm are **parity unpinned**
  beyond what the reference itself checks: the reference holds no golden
  vectors, seeds its inputs with unseeded ``rand`` and compares only with
  ``isapprox`` (rtol sqrt(eps)) against the serial Jets block operator over the
  same blocks (test:418-428,460-469,485-493,509-511,523-526,574-583).  Julia is
  not installed in this image, and the arithmetic lives in un-vendored packages
  (Jets ^1.2, DistributedArrays 0.6, DistributedOperations 1; Project.toml:14-19,
  no Manifest), so the reference cannot be run to generate vectors.  The oracle
  therefore reproduces the reference's *own* test (distributed result vs serial
  block operator, same tolerance) and in addition provides a higher-precision
  ground truth against which both the oracle's reference-ordered result and
  the CUDA result are measured.

Index convention: everything here is 1-based and inclusive like Julia; a Julia
``a:b`` UnitRange is represented by the Python ``range(a, b+1)`` (so ``in``,
``len``, ``==`` and iteration behave like the Julia object).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def ur(a: int, b: int) -> range:
    """Julia ``a:b`` (inclusive, 1-based values)."""
    return range(int(a), int(b) + 1)


# --------------------------------------------------------------------------
# DistributedArrays partition rules (upstream DistributedArrays 0.6, recalled;
# verified against every split the reference tests assert literally:
# test:157,384-385,536-539 and the 7-blocks-on-4-workers case test:172-221).
# --------------------------------------------------------------------------
def defaultdist_cuts(sz: int, nc: int) -> List[int]:
    """``DistributedArrays.defaultdist(sz::Int, nc::Int)``: cut points of an
    even split -- the first ``rem`` chunks get one extra item."""
    if sz >= nc:
        q, r = divmod(sz, nc)
        grid = []
        for i in range(1, nc + 2):
            g = (i - 1) * q + 1
            g += (i - 1) if i <= r else r
            grid.append(g)
        return grid
    return list(range(1, sz + 2)) + [0] * (nc - sz)


def _factor_keys_desc(n: int) -> List[int]:
    f, p, m = [], 2, n
    while p * p <= m:
        if m % p == 0:
            f.append(p)
            while m % p == 0:
                m //= p
        p += 1
    if m > 1:
        f.append(m)
    return sorted(f, reverse=True)


def defaultdist_chunks(dims: Sequence[int], npids: int) -> List[int]:
    """``DistributedArrays.defaultdist(dims, pids)``: how many chunks along each
    dimension -- repeatedly give the largest prime factor of ``np`` to the
    largest dimension (ties to the highest dimension)."""
    dims = list(dims)
    chunks = [1] * len(dims)
    np_ = npids
    f = _factor_keys_desc(np_)
    k = 0
    while np_ > 1:
        if np_ % f[k] != 0:
            k += 1
            if k >= len(f):
                break
        fac = f[k]
        d = max(dims)
        dno = max(i for i, v in enumerate(dims) if v == d)
        if dims[dno] >= fac:
            dims[dno] //= fac
            chunks[dno] *= fac
        np_ //= fac
    return chunks


def indices_from_cuts(cuts: Sequence[int]) -> List[range]:
    """``Jets.indices(cuts)`` src:40."""
    return [ur(cuts[i], cuts[i + 1] - 1) for i in range(len(cuts) - 1)]


def prescribedist(rngs: Sequence[range]) -> List[int]:
    """src:30-38: cut points of caller-prescribed ranges."""
    cuts = [r[0] for r in rngs]
    cuts.append(rngs[-1][-1] + 1)
    return cuts


def chunk_idxs(dims: Sequence[int], chunks: Sequence[int]):
    """``DistributedArrays.chunk_idxs(dims, chunks)`` -> (idxs, cuts); idxs is a
    column-major nested structure indexed ``idxs[i1][i2]...`` flattened to a dict
    keyed by 1-based chunk coordinates."""
    cuts = [defaultdist_cuts(d, c) for d, c in zip(dims, chunks)]
    return _idxs_from_cuts(cuts, chunks), cuts


def chunk_idxs_prescribed(*rngs: Sequence[range]):
    """src:19-28: the irregular variant, ranges prescribed per dimension."""
    chunks = [len(r) for r in rngs]
    cuts = [prescribedist(r) for r in rngs]
    return _idxs_from_cuts(cuts, chunks), cuts


def _idxs_from_cuts(cuts, chunks):
    idxs = {}
    nd = len(chunks)

    def rec(prefix):
        d = len(prefix)
        if d == nd:
            idxs[tuple(prefix)] = tuple(
                ur(cuts[k][prefix[k] - 1], cuts[k][prefix[k]] - 1) for k in range(nd))
            return
        for c in range(1, chunks[d] + 1):
            rec(prefix + [c])

    rec([])
    return idxs


def grid_pids(pids: Sequence[int], n1: int, n2: int) -> List[List[int]]:
    """First n1*n2 pids reshaped COLUMN-major to n1 x n2 (src:12, src:483;
    pinned by test:543 ``[w1 w3; w2 w4]``)."""
    assert len(pids) >= n1 * n2
    return [[pids[r + c * n1] for c in range(n2)] for r in range(n1)]


# --------------------------------------------------------------------------
# Jets.BlockArray (upstream Jets ^1.2, recalled): a vector of dense arrays,
# each flattened column-major, plus the element range of every block.
# --------------------------------------------------------------------------
class BlockArray:
    def __init__(self, arrays: List[np.ndarray]):
        self.arrays = arrays
        self.indices: List[range] = []
        ibeg = 1
        for a in arrays:                      # src:119-124
            iend = ibeg + a.size - 1
            self.indices.append(ur(ibeg, iend))
            ibeg = iend + 1

    def __len__(self):
        return self.indices[-1].stop - 1 if self.indices else 0      # indices[end][end]; an empty last block is k+1:k

    @property
    def dtype(self):
        return self.arrays[0].dtype

    def flat(self) -> np.ndarray:
        if not self.arrays:
            return np.zeros(0)
        return np.concatenate([a.reshape(-1, order="F") for a in self.arrays])

    def getindex(self, j: int):
        for a, rng in zip(self.arrays, self.indices):
            if j in rng:
                return a.reshape(-1, order="F")[j - rng[0]]
        raise IndexError(j)

    def similar(self, dtype=None) -> "BlockArray":
        return BlockArray([np.zeros(a.shape, dtype=dtype or a.dtype, order="F") for a in self.arrays])


def _float_real(dtype) -> np.dtype:
    """Julia ``float(real(T))``."""
    dt = np.dtype(dtype)
    if dt.kind == "c":
        return np.dtype(np.float32 if dt == np.complex64 else np.float64)
    if dt.kind == "f":
        return dt
    return np.dtype(np.float64)


def _array_norm(a: np.ndarray, p) -> float:
    """``LinearAlgebra.norm(::Array, p)`` in the array's float type."""
    ft = _float_real(a.dtype)
    v = np.abs(a.reshape(-1)).astype(ft, copy=False)
    if v.size == 0:
        return ft.type(0)
    if p == math.inf:
        return v.max()
    if p == -math.inf:
        return v.min()
    if p == 0:
        return ft.type(np.count_nonzero(v))
    if p == 1:
        return v.sum(dtype=ft)
    if p == 2:
        return ft.type(np.sqrt(np.dot(v, v)))
    _p = ft.type(p)
    return ft.type(np.sum(v ** _p, dtype=ft) ** (ft.type(1) / _p))


def blockarray_norm(x: BlockArray, p=2):
    """Jets ``norm(::BlockArray, p)``: per-block norms folded with the same
    rule the distributed fold uses (src:200-209)."""
    ft = _float_real(x.dtype)
    z = [_array_norm(a, p) for a in x.arrays]
    return _fold_norm(z, p, ft)


def _fold_norm(z, p, ft):
    """src:200-209."""
    if p == math.inf:
        return ft.type(max(z))
    if p == -math.inf:
        return ft.type(min(z))
    if p == 0 or p == 1:
        s = ft.type(0)
        for v in z:
            s = ft.type(s + v)
        return s
    _p = ft.type(p)
    s = ft.type(0)
    for v in z:
        s = ft.type(s + ft.type(v) ** _p)
    return ft.type(s ** (ft.type(1) / _p))


def blockarray_dot(x: BlockArray, y: BlockArray):
    t = x.dtype.type
    s = t(0)
    for a, b in zip(x.arrays, y.arrays):
        s = t(s + np.vdot(a.reshape(-1, order="F"), b.reshape(-1, order="F")))   # Julia's dot conjugates its first argument
    return s


# --------------------------------------------------------------------------
# JetDSpace (src:48-107) and DBArray (src:109-412)
# --------------------------------------------------------------------------
class OSpace:
    """One worker's ``JetBSpace``: the shapes of its blocks."""

    def __init__(self, dtype, shapes: Sequence[Tuple[int, ...]]):
        self.dtype = np.dtype(dtype)
        self.shapes = [tuple(int(v) for v in (s if isinstance(s, (tuple, list)) else (s,))) for s in shapes]

    def nblocks(self):
        return len(self.shapes)

    def __len__(self):
        return sum(int(np.prod(s)) for s in self.shapes)

    def __eq__(self, other):
        return self.dtype == other.dtype and self.shapes == other.shapes


class OJetDSpace:
    def __init__(self, blkspaces: List[OSpace], pids: List[int]):
        """src:54-67."""
        self.blkspaces = blkspaces
        self.pids = list(pids)
        self.blkindices: List[range] = []
        self.indices: List[range] = []
        blk0, i0 = 1, 1
        for s in blkspaces:
            self.blkindices.append(ur(blk0, blk0 + s.nblocks() - 1))
            self.indices.append(ur(i0, i0 + len(s) - 1))
            blk0 = self.blkindices[-1].stop
            i0 = self.indices[-1].stop

    def __eq__(self, other):                                     # src:69
        return (self.blkspaces == other.blkspaces and self.blkindices == other.blkindices
                and self.indices == other.indices)

    def size(self):                                              # src:78
        return (self.indices[-1].stop - 1,)

    def __len__(self):
        return self.indices[-1].stop - 1

    @property
    def dtype(self):
        return self.blkspaces[0].dtype

    def nblocks(self):                                           # src:89
        return self.blkindices[-1].stop - 1

    def blockmap(self):                                          # src:104
        return self.blkindices

    def procs(self):                                             # src:106
        return self.pids

    def nprocs(self):
        return len(self.pids)

    def localindices(self, pid):                                 # src:91
        return self.indices[self.pids.index(pid)]

    def localblockindices(self, pid):                            # src:97
        return self.blkindices[self.pids.index(pid)]

    def _make(self, fn: Callable[[Tuple[int, ...], np.dtype], np.ndarray]) -> "ODBArray":
        parts = [BlockArray([fn(shape, s.dtype) for shape in s.shapes]) for s in self.blkspaces]
        return ODBArray(parts, self.pids, self.blkindices)

    def zeros(self):                                             # src:379-381
        return self._make(lambda sh, dt: np.zeros(sh, dtype=dt, order="F"))

    def ones(self):
        return self._make(lambda sh, dt: np.ones(sh, dtype=dt, order="F"))

    def array(self):
        return self._make(lambda sh, dt: np.empty(sh, dtype=dt, order="F"))

    def rand(self, rng: np.random.Generator):
        return self._make(lambda sh, dt: np.asfortranarray(rng.random(sh, dtype=dt)))


class ODBArray:
    """src:109-113: one BlockArray per worker + element/block ranges per worker."""

    def __init__(self, parts: List[BlockArray], pids: List[int], blkindices: List[range]):
        self.parts = parts
        self.pids = list(pids)
        self.blkindices = list(blkindices)
        self.indices: List[range] = []
        ibeg = 1
        for p in parts:                                          # src:141-147
            iend = ibeg + len(p) - 1
            self.indices.append(ur(ibeg, iend))
            ibeg = iend + 1

    # -- constructors -------------------------------------------------------
    @staticmethod
    def from_function(f: Callable[[int], np.ndarray], blkidxs: List[range], pids: List[int]) -> "ODBArray":
        """src:135-151: every worker builds its blocks with ``f(iblock)``."""
        parts = [BlockArray([np.asfortranarray(f(i)) for i in blkidxs[k]]) for k in range(len(pids))]
        return ODBArray(parts, pids, blkidxs)

    @staticmethod
    def from_nblocks(f, nblks: int, pids: List[int], dist: Optional[List[int]] = None) -> "ODBArray":
        """src:153-159."""
        if dist is None:
            dist = defaultdist_chunks([nblks], len(pids))
        idxs, _ = chunk_idxs([nblks], dist)
        blkidxs = [idxs[(k,)][0] for k in range(1, dist[0] + 1)]
        return ODBArray.from_function(f, blkidxs, list(pids)[:dist[0]])

    # -- array interface ----------------------------------------------------
    @property
    def dtype(self):
        return self.parts[0].dtype

    def size(self):                                              # src:168
        return (self.indices[-1].stop - 1,)

    def __len__(self):
        return self.indices[-1].stop - 1

    def nblocks(self):                                           # src:297
        return self.blkindices[-1].stop - 1

    def blockmap(self):                                          # src:304
        return self.blkindices

    def procs(self):
        return self.pids

    def localindices(self, pid):                                 # src:302
        return self.indices[self.pids.index(pid)]

    def localblockindices(self, pid):                            # src:303
        return self.blkindices[self.pids.index(pid)]

    def getindex(self, i: int):                                  # src:174-179
        ipid = next(k for k, rng in enumerate(self.indices) if i in rng)
        return self.parts[ipid].getindex(i - self.indices[ipid][0] + 1)

    def similar(self, dtype=None, n: Optional[int] = None):      # src:181-187
        dtype = np.dtype(dtype or self.dtype)
        if n is not None and n != len(self):
            return np.empty(n, dtype=dtype)
        return ODBArray([p.similar(dtype) for p in self.parts], self.pids, self.blkindices)

    def space(self) -> OJetDSpace:                               # src:161-164
        return OJetDSpace([OSpace(self.dtype, [a.shape for a in p.arrays]) for p in self.parts], self.pids)

    # -- block access (src:383-412) ------------------------------------------
    def _owner(self, iblock: int) -> Tuple[int, int]:
        ipid = next(k for k, rng in enumerate(self.blkindices) if iblock in rng)    # src:389
        return ipid, iblock - self.blkindices[ipid][0] + 1                          # src:390

    def getblock(self, iblock: int) -> np.ndarray:
        ipid, d = self._owner(iblock)
        return self.parts[ipid].arrays[d - 1].copy(order="F")    # remote fetch = copy (docs:126-132)

    def getblock_into(self, iblock: int, out: np.ndarray):       # src:396-399
        out[...] = self.getblock(iblock)
        return out

    def setblock(self, iblock: int, blk):                        # src:409-412
        ipid, d = self._owner(iblock)
        self.parts[ipid].arrays[d - 1][...] = blk

    # -- reductions (per-worker partial, master fold) -------------------------
    def norm(self, p=2):                                         # src:189-210
        ft = _float_real(self.dtype)
        z = [blockarray_norm(part, p) for part in self.parts]
        return _fold_norm(z, p, ft)

    def dot(self, other: "ODBArray"):                            # src:212-223
        t = self.dtype.type
        s = t(0)
        for a, b in zip(self.parts, other.parts):
            s = t(s + blockarray_dot(a, b))
        return s

    def extrema(self):                                           # src:225-237
        mn = [min(a.min() for a in p.arrays) for p in self.parts]
        mx = [max(a.max() for a in p.arrays) for p in self.parts]
        return min(mn), max(mx)

    def mapreduce(self, f, op):                                  # src:239-250 (re-applies f to partials)
        t = self.dtype.type
        part = []
        for p in self.parts:
            v = f(p.flat())
            part.append(t(op.reduce(v)))
        return t(op.reduce(f(np.array(part, dtype=self.dtype))))

    def mean(self):                                              # src:252-263
        t = self.dtype.type
        s = t(0)
        for p in self.parts:
            s = t(s + p.flat().sum(dtype=self.dtype))
        return t(s / t(len(self)))

    def var(self, corrected=True, mean=None):                    # src:265-283
        t = self.dtype.type
        mu = self.mean() if mean is None else mean
        s = t(0)
        for p in self.parts:
            s = t(s + np.sum((p.flat() - mu) ** 2))
        return s / (len(self) - 1) if corrected else s / len(self)

    # -- collect / convert / fill --------------------------------------------
    def collect(self) -> BlockArray:                             # src:310-323
        arrays: List[np.ndarray] = []
        for p in self.parts:
            arrays.extend(a.copy(order="F") for a in p.arrays)
        return BlockArray(arrays)                                 # ranges re-based by the running total

    def to_array(self) -> np.ndarray:                            # src:329-330
        return self.collect().flat()

    def fill(self, a):                                           # src:332-339
        for p in self.parts:
            for arr in p.arrays:
                arr[...] = a
        return self

    # -- broadcasting (src:348-376): per worker, per block ---------------------
    def broadcast_assign(self, fn: Callable[..., np.ndarray], *args):
        """``dest .= fn.(args...)``: DBArray args are mapped to their local parts
        (src:363) and evaluated block by block in dest's element type."""
        for k, part in enumerate(self.parts):
            for b, arr in enumerate(part.arrays):
                local = [a.parts[k].arrays[b] if isinstance(a, ODBArray) else a for a in args]
                arr[...] = fn(*local)
        return self


# --------------------------------------------------------------------------
# Block kinds.  The reference has no block operators of its own (they are
# Jets closures); these restate the fixtures of its test-suite.
# --------------------------------------------------------------------------
class ODense:
    """``JopBaz(A)`` test:30-36: ``d .= A*m``; adjoint ``m .= A'*d``."""
    kind = "dense"

    def __init__(self, A: np.ndarray):
        self.A = np.asfortranarray(A)
        self.m, self.n = self.A.shape
        self.dtype = self.A.dtype
        self.rng_shape, self.dom_shape = (self.m,), (self.n,)

    def mul(self, x):
        return self.A @ x.reshape(-1, order="F")

    def mul_adjoint(self, y):
        return self.A.conj().T @ y.reshape(-1, order="F")


class ODiag:
    """``JopFoo(diag)`` test:8-12 / bench:7-11: ``d .= diagonal .* m`` (self-adjoint
    for real element types; N-d blocks keep their shape, flattened column-major)."""
    kind = "diag"

    def __init__(self, d: np.ndarray):
        self.d = np.asfortranarray(d)
        self.m = self.n = self.d.size
        self.dtype = self.d.dtype
        self.rng_shape = self.dom_shape = self.d.shape

    def mul(self, x):
        return (self.d.reshape(-1, order="F") * x.reshape(-1, order="F"))

    def mul_adjoint(self, y):
        return (np.conj(self.d).reshape(-1, order="F") * y.reshape(-1, order="F"))


class OZero:
    """``JopZeroBlock(dom, rng)``: ``iszero`` is true only for this kind (Jets)."""
    kind = "zero"

    def __init__(self, dtype, rng_shape, dom_shape):
        self.dtype = np.dtype(dtype)
        self.rng_shape = tuple(rng_shape) if isinstance(rng_shape, (tuple, list)) else (rng_shape,)
        self.dom_shape = tuple(dom_shape) if isinstance(dom_shape, (tuple, list)) else (dom_shape,)
        self.m, self.n = int(np.prod(self.rng_shape)), int(np.prod(self.dom_shape))


def iszero(op) -> bool:
    return op.kind == "zero"


def tree_reduce(bufs: List[np.ndarray]) -> np.ndarray:
    """``DistributedOperations.reduce!`` (upstream, recalled): binary tree of
    in-place ``+=`` ending in the first (caller's) entry."""
    bufs = list(bufs)
    stride = 1
    n = len(bufs)
    while stride < n:
        for i in range(0, n - stride, 2 * stride):
            bufs[i] += bufs[i + stride]
        stride *= 2
    return bufs[0]


def worker_rows_diag_forward(blocks, m_blocks, d_blocks) -> None:
    """One worker's share of a block-diagonal forward apply, ``JetDBlock_local_rows_diag_df!`` src:656-661:
    ``for i in local rows: mul!(getblock(d,i), A[i,i], getblock(m,i))``.  On the owner ``getblock`` aliases
    the stored block (docs/src/index.md:126-132), so the product lands in place."""
    for A, m, d in zip(blocks, m_blocks, d_blocks):
        d[...] = A.mul(m)


def worker_cols_diag_adjoint(blocks, m_blocks, d_blocks) -> None:
    """``JetDBlock_local_cols_diag_df′!`` src:703-708: ``mul!(getblock(m,j), A[j,j]', getblock(d,j))``."""
    for A, m, d in zip(blocks, m_blocks, d_blocks):
        m[...] = A.mul_adjoint(d)


class OJopDBlock:
    """Distributed block operator, src:414-508 (build) and src:510-721 (apply).

    ``blocks[i][j]`` (0-based Python nesting, i block row, j block column) are
    ODense / ODiag / OZero.  ``pids`` and ``dist=[n1,n2]`` are the DArray
    arguments; ``row_ranges``/``col_ranges`` optionally prescribe the chunking
    (src:8-17) instead of the even split.
    """

    def __init__(self, blocks, pids: Sequence[int], dist: Optional[Sequence[int]] = None, isdiag: bool = False,
                 row_ranges: Optional[List[range]] = None, col_ranges: Optional[List[range]] = None):
        self.blocks = blocks
        self.N, self.M = len(blocks), len(blocks[0])
        if row_ranges is not None:
            idxs, cuts = chunk_idxs_prescribed(row_ranges, col_ranges)
            n1, n2 = len(row_ranges), len(col_ranges)
        else:
            if dist is None:
                dist = defaultdist_chunks([self.N, self.M], len(pids))
            n1, n2 = int(dist[0]), int(dist[1])
            idxs, cuts = chunk_idxs([self.N, self.M], [n1, n2])
        if isdiag and n2 != 1:                                   # src:481
            raise ValueError("block diagonal jets must define distribution along rows.")
        self.n1, self.n2, self.isdiag = n1, n2, isdiag
        self.pids = grid_pids(list(pids), n1, n2)                # src:483, column-major
        # blockmap[r][c] = (rowrange, colrange)  (src:507 ``indices(ops)``)
        self.blockmap = [[idxs[(r + 1, c + 1)] for c in range(n2)] for r in range(n1)]
        self.dtype = blocks[0][0].dtype

        # range: one JetBSpace per grid row, on pids[:,1] (src:485)
        rng_spaces = []
        for r in range(n1):
            rows = self.blockmap[r][0][0]
            rng_spaces.append(OSpace(self.dtype, [self._row_shape(i) for i in rows]))
        self.rng = OJetDSpace(rng_spaces, [self.pids[r][0] for r in range(n1)])
        if isdiag:
            # src:464-474: exactly one non-zero block per local row; the domain
            # block of row i is taken from that row's COLUMN-1 block (src:471-473).
            dom_spaces = []
            for r in range(n1):
                rows = self.blockmap[r][0][0]
                for i in rows:
                    s = sum(0 if iszero(self.blocks[i - 1][j]) else 1 for j in range(self.M))
                    if s != 1:
                        raise ValueError("Expecting one non-zero column block per row in block diagonal "
                                         f"construction, got {s}.")
                dom_spaces.append(OSpace(self.dtype, [self.blocks[i - 1][0].dom_shape for i in rows]))
            self.dom = OJetDSpace(dom_spaces, [self.pids[r][0] for r in range(n1)])
        else:
            dom_spaces = []
            for c in range(n2):                                  # src:493: on pids[1,:]
                cols = self.blockmap[0][c][1]
                dom_spaces.append(OSpace(self.dtype, [self._col_shape(j) for j in cols]))
            self.dom = OJetDSpace(dom_spaces, [self.pids[0][c] for c in range(n2)])

    def _row_shape(self, i):
        return self.blocks[i - 1][0].rng_shape

    def _col_shape(self, j):
        return self.blocks[0][j - 1].dom_shape

    # -- accessors (src:805-852) ---------------------------------------------
    def nblocks(self):
        return (self.N, self.M)

    def procs(self):
        return self.pids

    def nprocs(self):
        return self.n1 * self.n2

    def localblockindices(self, pid, dim: Optional[int] = None):   # src:814-825
        for c in range(self.n2):
            for r in range(self.n1):
                if self.pids[r][c] == pid:
                    bm = self.blockmap[r][c]
                    return bm if dim is None else bm[dim - 1]
        bm = (ur(0, 0), ur(0, 0))
        return bm if dim is None else bm[dim - 1]

    def getblock(self, i: int, j: int):                            # src:838-848
        for r in range(self.n1):
            for c in range(self.n2):
                if i in self.blockmap[r][c][0] and j in self.blockmap[r][c][1]:
                    return self.blocks[i - 1][j - 1]
        raise IndexError((i, j))

    # -- apply, distributed domain (src:582-721) ------------------------------
    def mul(self, d: ODBArray, m: ODBArray) -> ODBArray:
        """``mul!(d, A, m)`` -> ``JetDBlock_df!`` src:663-674 (linear blocks:
        ``f!`` src:616-627 is identical)."""
        for r in range(self.n1):
            rows = self.blockmap[r][0][0]
            if self.isdiag:                                       # src:656-661
                for i in rows:
                    d.setblock(i, self.blocks[i - 1][i - 1].mul(m.getblock(i)).reshape(self._row_shape(i), order="F"))
                continue
            for i in rows:                                        # src:641-654, sequential over block rows
                partials = []
                for c in range(self.n2):                          # one partial per grid column (src:647-650)
                    p = np.zeros(int(np.prod(self._row_shape(i))), dtype=self.dtype)
                    for j in self.blockmap[r][c][1]:              # src:632-637
                        op = self.blocks[i - 1][j - 1]
                        if not iszero(op):
                            p += op.mul(m.getblock(j))
                    partials.append(p)
                d.setblock(i, tree_reduce(partials).reshape(self._row_shape(i), order="F"))    # src:651
        return d

    def mul_adjoint(self, m: ODBArray, d: ODBArray) -> ODBArray:
        """``mul!(m, A', d)`` -> ``JetDBlock_df'!`` src:710-721."""
        if self.isdiag:                                           # src:703-708
            for r in range(self.n1):
                for j in self.blockmap[r][0][0]:
                    m.setblock(j, self.blocks[j - 1][j - 1].mul_adjoint(d.getblock(j)).reshape(
                        self.blocks[j - 1][0].dom_shape, order="F"))
            return m
        for c in range(self.n2):
            for j in self.blockmap[0][c][1]:                      # src:688-701
                partials = []
                for r in range(self.n1):
                    p = np.zeros(int(np.prod(self._col_shape(j))), dtype=self.dtype)
                    for i in self.blockmap[r][c][0]:              # src:679-684
                        op = self.blocks[i - 1][j - 1]
                        if not iszero(op):
                            p += op.mul_adjoint(d.getblock(i))
                    partials.append(p)
                m.setblock(j, tree_reduce(partials).reshape(self._col_shape(j), order="F"))    # src:698
        return m

    # -- apply, replicated domain (grid n x 1, m a plain array; src:510-580) ----
    def mul_replicated(self, d: ODBArray, m: np.ndarray) -> ODBArray:
        """src:544-557: broadcast m, every worker applies its serial local block op."""
        assert self.n2 == 1
        offs = self._dom_offsets()
        for r in range(self.n1):
            for i in self.blockmap[r][0][0]:
                acc = np.zeros(int(np.prod(self._row_shape(i))), dtype=self.dtype)
                for j in range(1, self.M + 1):
                    op = self.blocks[i - 1][j - 1]
                    if not iszero(op):
                        acc += op.mul(m[offs[j - 1]:offs[j]])
                d.setblock(i, acc.reshape(self._row_shape(i), order="F"))
        return d

    def mul_adjoint_replicated(self, m: np.ndarray, d: ODBArray) -> np.ndarray:
        """src:565-580: every worker fills a full-size partial of m; tree sum to master."""
        assert self.n2 == 1
        offs = self._dom_offsets()
        partials = [np.zeros(offs[-1], dtype=self.dtype)]         # master's m .= 0 (src:566)
        for r in range(self.n1):
            p = np.zeros(offs[-1], dtype=self.dtype)
            for j in range(1, self.M + 1):
                for i in self.blockmap[r][0][0]:
                    op = self.blocks[i - 1][j - 1]
                    if not iszero(op):
                        p[offs[j - 1]:offs[j]] += op.mul_adjoint(d.getblock(i))
            partials.append(p)
        m[...] = tree_reduce(partials)
        return m

    def _dom_offsets(self):
        offs = [0]
        for j in range(1, self.M + 1):
            offs.append(offs[-1] + int(np.prod(self._col_shape(j))))
        return offs


# --------------------------------------------------------------------------
# The reference's own oracle: the SERIAL Jets block operator over the same
# blocks (test:399-401 ``G = @blockop _G``): loop rows, then columns,
# accumulating through a temporary (Jets JetBlock_df!, recalled).
# --------------------------------------------------------------------------
def serial_blockop_mul(blocks, x: np.ndarray, dtype=None) -> np.ndarray:
    dtype = dtype or blocks[0][0].dtype
    xo = np.cumsum([0] + [b.n for b in blocks[0]])
    out = []
    for row in blocks:
        acc = np.zeros(row[0].m, dtype=dtype)
        for j, op in enumerate(row):
            if not iszero(op):
                acc += op.mul(x[xo[j]:xo[j + 1]])
        out.append(acc)
    return np.concatenate(out)


def serial_blockop_mul_adjoint(blocks, y: np.ndarray, dtype=None) -> np.ndarray:
    dtype = dtype or blocks[0][0].dtype
    yo = np.cumsum([0] + [row[0].m for row in blocks])
    out = []
    for j in range(len(blocks[0])):
        acc = np.zeros(blocks[0][j].n, dtype=dtype)
        for i in range(len(blocks)):
            op = blocks[i][j]
            if not iszero(op):
                acc += op.mul_adjoint(y[yo[i]:yo[i + 1]])
        out.append(acc)
    return np.concatenate(out)


# --------------------------------------------------------------------------
# Higher-precision ground truth (not in the reference): Float32 problems are
# evaluated in Float64, Float64 problems in x87 long double.
# --------------------------------------------------------------------------
def _wide(dtype):
    return np.float64 if np.dtype(dtype) == np.float32 else np.longdouble


def ground_truth_mul(blocks, x: np.ndarray, adjoint: bool = False) -> np.ndarray:
    wt = _wide(blocks[0][0].dtype)
    xw = x.astype(wt)
    if not adjoint:
        xo = np.cumsum([0] + [b.n for b in blocks[0]])
        out = []
        for row in blocks:
            acc = np.zeros(row[0].m, dtype=wt)
            for j, op in enumerate(row):
                seg = xw[xo[j]:xo[j + 1]]
                if op.kind == "dense":
                    acc += op.A.astype(wt) @ seg
                elif op.kind == "diag":
                    acc += op.d.reshape(-1, order="F").astype(wt) * seg
            out.append(acc)
        return np.concatenate(out)
    yo = np.cumsum([0] + [row[0].m for row in blocks])
    out = []
    for j in range(len(blocks[0])):
        acc = np.zeros(blocks[0][j].n, dtype=wt)
        for i in range(len(blocks)):
            op = blocks[i][j]
            seg = xw[yo[i]:yo[i + 1]]
            if op.kind == "dense":
                acc += op.A.astype(wt).T @ seg
            elif op.kind == "diag":
                acc += op.d.reshape(-1, order="F").astype(wt) * seg
        out.append(acc)
    return np.concatenate(out)


def blockwise_relerr(got: np.ndarray, truth: np.ndarray, block_lens: Sequence[int]) -> float:
    """Norm-wise relative error per block, worst block (BASELINE.md parity bar)."""
    worst, o = 0.0, 0
    for n in block_lens:
        t = truth[o:o + n].astype(np.longdouble)
        g = got[o:o + n].astype(np.longdouble)
        den = float(np.sqrt(np.sum(t * t)))
        num = float(np.sqrt(np.sum((g - t) ** 2)))
        worst = max(worst, num / den if den > 0 else num)
        o += n
    return worst


# --------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench: U[0,1) in
# the working dtype from a stated seed; blocks drawn in row-major block order.
# --------------------------------------------------------------------------
def make_blocks(rng: np.random.Generator, nbr: int, nbc: int, m: int, n: int, dtype,
                kind_of: Optional[Callable[[int, int], str]] = None):
    """``kind_of(i, j)`` (1-based) -> 'dense' | 'diag' | 'zero'; default all dense."""
    blocks = []
    for i in range(1, nbr + 1):
        row = []
        for j in range(1, nbc + 1):
            k = kind_of(i, j) if kind_of else "dense"
            if k == "dense":
                row.append(ODense(np.asfortranarray(rng.random((m, n), dtype=dtype))))
            elif k == "diag":
                assert m == n
                row.append(ODiag(rng.random(n, dtype=dtype)))
            else:
                row.append(OZero(dtype, (m,), (n,)))
        blocks.append(row)
    return blocks
