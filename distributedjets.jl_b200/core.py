"""Host-side mirror of the reference's public interface for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow src/DistributedJets.jl (cited as ``src:LINE``):
``DBArray``, ``JetDSpace``, ``blockmap``, ``localblockindices`` (the reference's exports, src:863),
``getblock`` / ``getblock_into`` (``getblock!``) / ``setblock`` (``setblock!``), ``mul`` (``mul!``),
``dot`` / ``norm`` / ``extrema`` / ``mapreduce`` / ``mean`` / ``var`` / ``fill``, broadcasting for the
expression class the reference's tests use, and ``blockop`` (``@blockop`` over a ``DArray`` of blocks).
Block and element indices are 1-based and inclusive like Julia's; "pids" are rank numbers
(one rank = one GPU = one former Distributed.jl worker).

Everything that touches data calls libdjets_b200.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DimensionMismatch, DJetsError
from .partition import cuts_from_ranges, default_grid, even_cuts, even_ranges, grid_of, urange

_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.complex64): _lib.C64,
       np.dtype(np.complex128): _lib.C128}


def _dt(dtype) -> int:
    try:
        return _DT[np.dtype(dtype)]
    except KeyError:
        raise DJetsError(_lib.ERR_UNSUPPORTED, f"element type {dtype} (Float32, Float64 and their Complex types only)") from None


# ======================================================================================================
# context: the set of workers
# ======================================================================================================
class Context:
    """The ranks of the job.  ``Context(world=4)`` drives 4 ranks from this process (the Julia-master
    model; with fewer GPUs than ranks several ranks share a GPU, each on its own stream);
    ``Context.spmd(rank, world, device, uid)`` drives one rank per process (torchrun)."""

    def __init__(self, world: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                 local_ranks: Optional[Sequence[int]] = None, nccl_uid: Optional[bytes] = None):
        ndev = C.c_int32()
        _lib.call("djets_device_count", C.byref(ndev))
        if world is None:
            world = max(int(ndev.value), 1)
        if local_ranks is None:
            local_ranks = list(range(world))
        if devices is None:
            devices = [r % max(int(ndev.value), 1) for r in local_ranks]
        uid = None
        if nccl_uid is not None:
            assert len(nccl_uid) == 128
            uid = (C.c_uint8 * 128).from_buffer_copy(nccl_uid)
        h = C.c_void_p()
        _lib.call("djets_ctx_create", world, len(local_ranks), _lib.arr_i32(local_ranks), _lib.arr_i32(devices),
                  uid, C.byref(h))
        self.h = h
        self.world = int(world)
        self.local_ranks = [int(r) for r in local_ranks]
        self.devices = [int(d) for d in devices]
        self.myrank = self.local_ranks[0] if len(self.local_ranks) == 1 and world > 1 else None

    @staticmethod
    def spmd(rank: int, world: int, device: int, nccl_uid: Optional[bytes] = None,
             allgather: Optional[Callable[[bytes], List[bytes]]] = None) -> "Context":
        """One rank per process.  ``nccl_uid``: exchange over NCCL where needed; ``allgather``: a host-side collective
        ``bytes -> [bytes of every process]`` (e.g. ``torch_allgather()``) that carries the control plane, so the job
        needs no NCCL at all: block data moves through CUDA-IPC-mapped peer memory."""
        ctx = Context(world=world, devices=[device], local_ranks=[rank], nccl_uid=nccl_uid)
        if allgather is not None:
            ctx.set_host_allgather(allgather)
        return ctx

    def set_host_allgather(self, fn: Callable[[bytes], List[bytes]]) -> None:
        def thunk(_user, send, recv, nbytes):
            try:
                records = fn(C.string_at(send, nbytes))
                for k, rec in enumerate(records):
                    assert len(rec) == nbytes
                    C.memmove(recv + k * nbytes, rec, nbytes)
                return 0
            except Exception:                                # never let an exception cross the C ABI
                import traceback
                traceback.print_exc()
                return 1
        self._allgather_cb = _lib.ALLGATHER_FN(thunk)        # kept alive as long as the context
        _lib.call("djets_ctx_set_host_allgather", self.h, C.cast(self._allgather_cb, C.c_void_p), None)

    def workers(self) -> List[int]:
        return list(range(self.world))

    def is_local(self, rank: int) -> bool:
        return int(rank) in self.local_ranks

    def sync(self) -> None:
        _lib.call("djets_ctx_sync", self.h)

    def set_param(self, name: str, value: int) -> None:
        _lib.call("djets_ctx_set_param", self.h, name.encode(), int(value))

    # timing ------------------------------------------------------------------------------------------
    def timer_start(self) -> None:
        _lib.call("djets_timer_start", self.h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        _lib.call("djets_timer_stop", self.h, C.byref(ms))
        return float(ms.value)

    def kernel_timing(self, enable: bool) -> None:
        _lib.call("djets_ctx_kernel_timing", self.h, 1 if enable else 0)

    def kernel_stats(self) -> dict:
        out = {}
        for k, name in _lib.KERNEL_NAMES.items():
            n, ms = C.c_int64(), C.c_double()
            _lib.call("djets_ctx_kernel_stats", self.h, k, C.byref(n), C.byref(ms))
            out[name] = {"launches": int(n.value), "ms": float(ms.value)}
        return out

    def reset_stats(self) -> None:
        _lib.call("djets_ctx_reset_stats", self.h)

    # stream marks and captured sequences -------------------------------------------------------------
    def mark(self, slot: int = 0) -> None:
        """Record mark ``slot`` (0..7) on every local rank's stream."""
        _lib.call("djets_ctx_mark", self.h, int(slot))

    def wait_mark(self, slot: int = 0) -> None:
        """Block until everything enqueued before mark ``slot`` has finished (later work keeps running)."""
        _lib.call("djets_ctx_wait_mark", self.h, int(slot))

    def capture(self) -> "_Capture":
        """``with ctx.capture() as g: ...`` records the stream-ordered calls made inside (mul!, adjoint,
        ``dot_s`` / ``norm_s``, scalar arithmetic, broadcasts) instead of running them; ``g.launch()``
        replays the whole sequence with one launch -- a solver iteration without host round trips."""
        return _Capture(self)

    def close(self) -> None:
        if self.h:
            _lib.call("djets_ctx_destroy", self.h)
            self.h = None


class Graph:
    """A captured sequence of calls (djets_graph)."""

    def __init__(self, ctx: "Context"):
        self.ctx, self.h = ctx, None

    def launch(self) -> None:
        _lib.call("djets_graph_launch", self.h)

    def close(self) -> None:
        h, self.h = self.h, None
        if h and getattr(self.ctx, "h", None):
            _lib.lib.djets_graph_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Capture:
    def __init__(self, ctx: "Context"):
        self.ctx, self.graph = ctx, Graph(ctx)

    def __enter__(self) -> Graph:
        _lib.call("djets_ctx_capture_begin", self.ctx.h)
        return self.graph

    def __exit__(self, exc_type, exc, tb):
        h = C.c_void_p()
        status = _lib.lib.djets_ctx_capture_end(self.ctx.h, C.byref(h))
        if exc_type is None:
            _lib.check(status)
            self.graph.h = h
        return False


class Scalar:
    """A scalar resident on the devices (one replica per rank): the result of ``dot_s`` / ``norm_s``, a solver
    coefficient.  Arithmetic between Scalars runs on the devices, rounded to ``dtype`` like the reference's
    element-type scalars; ``float(s)`` synchronises."""

    def __init__(self, value: Optional[float] = None, dtype=np.float64, ctx: Optional["Context"] = None):
        self.ctx = ctx or context()
        self.dtype = np.dtype(dtype)
        h = C.c_void_p()
        _lib.call("djets_scalar_create", self.ctx.h, C.byref(h))
        self.h = h
        if value is not None:
            self.set(value)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h and getattr(self.ctx, "h", None):
            try:
                _lib.lib.djets_scalar_destroy(h)
            except Exception:
                pass

    def set(self, v: float) -> "Scalar":
        _lib.call("djets_scalar_set", self.h, float(self.dtype.type(v)))
        return self

    def get(self):
        out = C.c_double()
        _lib.call("djets_scalar_get", self.h, C.byref(out))
        return self.dtype.type(out.value)

    def __float__(self):
        return float(self.get())

    def assign(self, op: int, a: "Scalar", b: Optional["Scalar"] = None) -> "Scalar":
        """``self = a op b`` on the devices (``_lib.SC_*``), in place."""
        _lib.call("djets_scalar_op", self.h, int(op), a.h, b.h if b is not None else None, _dt(self.dtype))
        return self

    def _new(self, op: int, other=None) -> "Scalar":
        if isinstance(other, (Expr, DBArray, LinExpr)):
            return NotImplemented                          # scalar (op) lazy vector expression: the Expr side lowers it
        if other is not None and not isinstance(other, Scalar):
            other = Scalar(other, self.dtype, self.ctx)
        return Scalar(None, self.dtype, self.ctx).assign(op, self, other)

    def __add__(self, o): return self._new(_lib.SC_ADD, o)
    def __sub__(self, o): return self._new(_lib.SC_SUB, o)
    def __mul__(self, o): return self._new(_lib.SC_MUL, o)
    def __truediv__(self, o): return self._new(_lib.SC_DIV, o)
    def __neg__(self): return self._new(_lib.SC_NEG)
    def sqrt(self): return self._new(_lib.SC_SQRT)
    def copy(self): return self._new(_lib.SC_COPY)


_default_ctx: Optional[Context] = None


def init(world: Optional[int] = None, devices: Optional[Sequence[int]] = None, **kw) -> Context:
    """``addprocs``: make ``world`` workers the default context."""
    global _default_ctx
    _default_ctx = Context(world, devices, **kw)
    return _default_ctx


def set_default_context(ctx: Context) -> None:
    global _default_ctx
    _default_ctx = ctx


def context() -> Context:
    if _default_ctx is None:
        raise DJetsError(_lib.ERR_INVALID, "no workers: call distributedjets_b200.init(world) first")
    return _default_ctx


def workers() -> List[int]:
    return context().workers()


def nworkers() -> int:
    return context().world


def torch_allgather(group=None) -> Callable[[bytes], List[bytes]]:
    """Host all-gather over an initialised ``torch.distributed`` process group (gloo or nccl)."""
    import torch.distributed as dist

    def gather(send: bytes) -> List[bytes]:
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, send, group=group)
        return out
    return gather


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _lib.call("djets_nccl_unique_id", buf)
    return bytes(buf)


class PinnedBuffer:
    """Page-locked host array for setblock!/getblock/collect traffic."""

    def __init__(self, n: int, dtype):
        self.dtype = np.dtype(dtype)
        p = C.c_void_p()
        _lib.call("djets_host_alloc", int(n) * self.dtype.itemsize, C.byref(p))
        self.ptr = p
        self.array = np.ctypeslib.as_array((C.c_byte * (int(n) * self.dtype.itemsize)).from_address(p.value)).view(
            self.dtype)

    def free(self):
        if self.ptr:
            self.array = None
            _lib.call("djets_host_free", self.ptr)
            self.ptr = None


# ======================================================================================================
# JetDSpace (src:48-107)
# ======================================================================================================
def _shape(s) -> Tuple[int, ...]:
    return tuple(int(v) for v in s) if isinstance(s, (tuple, list)) else (int(s),)


def _size(shape) -> int:
    n = 1
    for v in shape:
        n *= v
    return n


class JetDSpace:
    """Distributed block vector space: one list of block shapes per worker (the worker's JetBSpace),
    block ranges and element ranges per worker (src:48-67)."""

    def __init__(self, blkshapes: Sequence[Sequence], pids: Sequence[int], dtype, ctx: Optional[Context] = None):
        self.ctx = ctx or context()
        self.dtype = np.dtype(dtype)
        self.blkshapes: List[List[Tuple[int, ...]]] = [[_shape(s) for s in part] for part in blkshapes]
        self.pids = [int(p) for p in pids]
        if len(self.pids) != len(self.blkshapes):
            raise ValueError("one list of block shapes per worker")
        self.blkindices: List[range] = []
        self.indices: List[range] = []
        b0, e0 = 1, 1
        for part in self.blkshapes:                      # src:55-64
            nb, ne = len(part), sum(_size(s) for s in part)
            self.blkindices.append(urange(b0, b0 + nb - 1))
            self.indices.append(urange(e0, e0 + ne - 1))
            b0, e0 = b0 + nb, e0 + ne

    def __eq__(self, other):                             # src:69
        return (isinstance(other, JetDSpace) and self.blkshapes == other.blkshapes and self.dtype == other.dtype
                and self.blkindices == other.blkindices and self.indices == other.indices)

    def __len__(self):
        return self.indices[-1].stop - 1 if self.indices else 0      # Julia's indices[end][end], empty ranges included

    def size(self):                                      # src:78
        return (len(self),)

    @property
    def eltype(self):
        return self.dtype

    def nblocks(self) -> int:                            # src:89
        return self.blkindices[-1].stop - 1

    def blockmap(self) -> List[range]:                   # src:104
        return self.blkindices

    def procs(self) -> List[int]:                        # src:106
        return list(self.pids)

    def nprocs(self) -> int:
        return len(self.pids)

    def localindices(self, pid: Optional[int] = None) -> range:      # src:91-95
        return self.indices[self.pids.index(_mypid(self.ctx, pid))]

    def localblockindices(self, pid: Optional[int] = None) -> range:  # src:97-99
        return self.blkindices[self.pids.index(_mypid(self.ctx, pid))]

    def block_shapes(self) -> List[Tuple[int, ...]]:
        return [s for part in self.blkshapes for s in part]

    # Array / zeros / ones / rand (src:378-381) ---------------------------------------------------------
    def Array(self) -> "DBArray":
        return DBArray._new(self.ctx, self.dtype, self.pids, self.blkshapes)

    def zeros(self) -> "DBArray":
        return self.Array()                              # device storage is zero-initialised

    def ones(self) -> "DBArray":
        return self.Array().fill(1.0)

    def rand(self, rng: Optional[np.random.Generator] = None) -> "DBArray":
        rng = rng or np.random.default_rng()
        x = self.Array()
        flat = np.empty(len(self), dtype=self.dtype)
        o = 0
        for shape in self.block_shapes():                # block by block: drawn on every process, identical streams
            n = _size(shape)
            if self.dtype.kind == "c":                   # Julia's rand(Complex{T}): both parts U[0,1)
                rt = np.zeros(0, self.dtype).real.dtype
                blk = rng.random(shape, dtype=rt) + 1j * rng.random(shape, dtype=rt)
            else:
                blk = rng.random(shape, dtype=self.dtype)
            flat[o:o + n] = blk.reshape(-1, order="F")
            o += n
        return x.scatter(flat)                           # one DMA per rank


def _mypid(ctx: Context, pid: Optional[int]) -> int:
    if pid is not None:
        return int(pid)
    if ctx.myrank is None:
        raise DJetsError(_lib.ERR_INVALID, "this process drives several ranks: pass pid=")
    return ctx.myrank


# ======================================================================================================
# DBArray (src:109-412)
# ======================================================================================================
class DBArray:
    """Distributed block vector: per-worker runs of blocks resident in that worker's GPU memory."""

    __array_ufunc__ = None          # NumPy scalars defer to __rmul__ instead of coercing the DBArray

    def __init__(self, f: Callable[[int], np.ndarray], blkidxs, pids: Optional[Sequence[int]] = None,
                 dist: Optional[Sequence[int]] = None, ctx: Optional[Context] = None, dtype=None):
        """``DBArray(f, blkidxs::Vector{UnitRange}, pids)`` (src:135-151) or
        ``DBArray(f, (nblks,), pids[, dist])`` (src:153-159).  ``f(iblock)`` builds block ``iblock``."""
        ctx = ctx or context()
        if isinstance(blkidxs, tuple) or isinstance(blkidxs, int):
            nblks = int(blkidxs[0] if isinstance(blkidxs, tuple) else blkidxs)
            if pids is None:
                pids = ctx.workers()[:min(ctx.world, nblks)]                     # src:159
            nchunks = int(dist[0]) if dist is not None else default_grid([nblks], len(pids))[0]
            blkidxs = even_ranges(nblks, nchunks)
            pids = list(pids)[:nchunks]
        else:
            blkidxs = [r for r in blkidxs]
            pids = list(pids if pids is not None else ctx.workers())[:len(blkidxs)]
        # block shapes must be known everywhere; payloads are only materialised for local owners
        blocks = {}
        shapes: List[List[Tuple[int, ...]]] = []
        for pid, rng in zip(pids, blkidxs):
            part = []
            for ib in rng:
                b = np.asarray(f(ib))
                part.append(b.shape)
                if ctx.is_local(pid):
                    blocks[ib] = b
                dtype = dtype or b.dtype
            shapes.append(part)
        self._init(ctx, np.dtype(dtype), pids, shapes)
        for ib, b in blocks.items():
            self.setblock(ib, b)

    @classmethod
    def _new(cls, ctx, dtype, pids, blkshapes) -> "DBArray":
        self = cls.__new__(cls)
        self._init(ctx, np.dtype(dtype), pids, blkshapes)
        return self

    @classmethod
    def _wrap(cls, ctx, handle, dtype, pids, blkshapes) -> "DBArray":
        self = cls.__new__(cls)
        self._init(ctx, np.dtype(dtype), pids, blkshapes, handle=handle)
        return self

    def _init(self, ctx, dtype, pids, blkshapes, handle=None):
        self.ctx = ctx
        self.dtype = dtype
        self.pids = [int(p) for p in pids]
        self.blkshapes = [[_shape(s) for s in part] for part in blkshapes]
        self._shapes = [s for part in self.blkshapes for s in part]
        self._lens = [_size(s) for s in self._shapes]
        self.blkindices: List[range] = []
        self.indices: List[range] = []
        b0, e0 = 1, 1
        for part in self.blkshapes:                      # src:141-147
            nb, ne = len(part), sum(_size(s) for s in part)
            self.blkindices.append(urange(b0, b0 + nb - 1))
            self.indices.append(urange(e0, e0 + ne - 1))
            b0, e0 = b0 + nb, e0 + ne
        if handle is None:
            handle = C.c_void_p()
            _lib.call("djets_vec_create", ctx.h, _dt(dtype), len(self.pids), _lib.arr_i32(self.pids),
                      _lib.arr_i64([len(p) for p in self.blkshapes]), _lib.arr_i64(self._lens), C.byref(handle))
        self.h = handle

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h and getattr(self.ctx, "h", None):
            try:
                _lib.lib.djets_vec_destroy(h)
            except Exception:
                pass

    # array interface (src:166-187) --------------------------------------------------------------------
    def __len__(self):
        return self.indices[-1].stop - 1 if self.indices else 0      # Julia's indices[end][end], empty ranges included

    def size(self):
        return (len(self),)

    @property
    def eltype(self):
        return self.dtype

    def __getitem__(self, i):
        """``x[i]``, 1-based (src:174-179); ``x[range(a, b+1)]`` is Julia's ``x[a:b]`` (test:138, 167)."""
        if isinstance(i, range):
            return np.array([self[k] for k in i], dtype=self.dtype)
        if not 1 <= int(i) <= len(self):
            raise IndexError(f"BoundsError: attempt to access {len(self)}-element DBArray at index [{i}]")
        if self.dtype.kind == "c":
            z = (C.c_double * 2)()
            _lib.call("djets_vec_getindex_c", self.h, int(i) - 1, z)
            return self.dtype.type(complex(z[0], z[1]))
        out = C.c_double()
        _lib.call("djets_vec_getindex", self.h, int(i) - 1, C.byref(out))
        return self.dtype.type(out.value)

    def similar(self, dtype=None, n=None):
        """``similar(x[, T][, n])`` (src:181-187): a plain array when the length differs."""
        dtype = np.dtype(dtype or self.dtype)
        if n is not None:
            n = int(n[0]) if isinstance(n, tuple) else int(n)
            if n != len(self):
                return np.empty(n, dtype=dtype)
        h = C.c_void_p()
        _lib.call("djets_vec_similar", self.h, _dt(dtype), C.byref(h))
        return DBArray._wrap(self.ctx, h, dtype, self.pids, self.blkshapes)

    def space(self) -> JetDSpace:                        # src:161-164
        return JetDSpace(self.blkshapes, self.pids, self.dtype, self.ctx)

    def reshape(self, R: JetDSpace) -> "DBArray":        # src:342-345
        if len(self) != len(R):
            raise DJetsError(_lib.ERR_MISMATCH, "dimension mismatch, unable to reshap distributed block arrays")
        return self

    # metadata (src:285-304) -----------------------------------------------------------------------------
    def procs(self) -> List[int]:
        return list(self.pids)

    def nprocs(self) -> int:
        return len(self.pids)

    def nblocks(self) -> int:
        return self.blkindices[-1].stop - 1

    def blockmap(self) -> List[range]:
        return self.blkindices

    def localindices(self, pid: Optional[int] = None) -> range:
        return self.indices[self.pids.index(_mypid(self.ctx, pid))]

    def localblockindices(self, pid: Optional[int] = None) -> range:
        return self.blkindices[self.pids.index(_mypid(self.ctx, pid))]

    def block_shape(self, iblock: int) -> Tuple[int, ...]:
        return self._shapes[int(iblock) - 1]

    # block access (src:383-412) ---------------------------------------------------------------------------
    def _check_block(self, iblock: int) -> int:
        ib = int(iblock)
        if not 1 <= ib <= len(self._shapes):
            raise IndexError(f"block {ib} outside 1:{len(self._shapes)}")  # findfirst -> nothing in the reference
        return ib

    def getblock(self, iblock: int) -> np.ndarray:
        """``getblock(x, i)``: a copy of block ``i`` in its own shape (column-major)."""
        ib = self._check_block(iblock)
        out = np.empty(self._shapes[ib - 1], dtype=self.dtype, order="F")
        _lib.call("djets_vec_getblock", self.h, ib - 1, out.ctypes.data_as(C.c_void_p))
        return out

    def getblock_into(self, iblock: int, xblock: np.ndarray) -> np.ndarray:
        """``getblock!(x, i, xblock)`` (src:396-399)."""
        xblock[...] = self.getblock(iblock)
        return xblock

    def setblock(self, iblock: int, xblock) -> None:
        """``setblock!(x, i, xblock)`` (src:409-412): ``x.arrays[i] .= xblock`` on the owner."""
        ib = self._check_block(iblock)
        shape = self._shapes[ib - 1]
        src = np.asarray(xblock)
        if src.ndim == 0:
            src = np.full(shape, src, dtype=self.dtype)
        if _size(src.shape) != _size(shape):
            raise DimensionMismatch(_lib.ERR_MISMATCH,
                                    f"DimensionMismatch: block {ib} has shape {shape}, got {src.shape}")
        buf = np.asfortranarray(src.reshape(shape, order="F") if src.shape != shape else src, dtype=self.dtype)
        _lib.call("djets_vec_setblock", self.h, ib - 1, buf.ctypes.data_as(C.c_void_p))

    def getblock_raw(self, iblock: int, host_ptr: int) -> None:
        _lib.call("djets_vec_getblock", self.h, int(iblock) - 1, C.c_void_p(host_ptr))

    # collect / convert (src:310-330) ------------------------------------------------------------------------
    def collect(self) -> List[np.ndarray]:
        """``collect(x)``: the blocks in order, each in its own shape (a Jets.BlockArray's ``arrays``)."""
        flat = self.to_array()
        out, o = [], 0
        for shape, n in zip(self._shapes, self._lens):
            out.append(flat[o:o + n].reshape(shape, order="F"))
            o += n
        return out

    def to_array(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``convert(Array, x)``."""
        if out is None:
            out = np.zeros(len(self), dtype=self.dtype)
        assert out.dtype == self.dtype and out.size == len(self) and out.flags.c_contiguous
        _lib.call("djets_vec_collect", self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def scatter(self, flat: np.ndarray) -> "DBArray":
        """Inverse of ``to_array``: load every locally owned part from its slice of ``flat``."""
        assert flat.dtype == self.dtype and flat.size == len(self) and flat.flags.c_contiguous
        _lib.call("djets_vec_scatter", self.h, flat.ctypes.data_as(C.c_void_p))
        return self

    def scatter_async(self, flat: np.ndarray) -> "DBArray":
        """Stream-ordered ``scatter``: returns at once; ``flat`` (page-locked) must stay unmodified until the
        next synchronising call."""
        assert flat.dtype == self.dtype and flat.size == len(self) and flat.flags.c_contiguous
        _lib.call("djets_vec_scatter_async", self.h, flat.ctypes.data_as(C.c_void_p))
        return self

    def to_array_async(self, out: np.ndarray) -> np.ndarray:
        """Stream-ordered ``to_array`` into a page-locked buffer: valid after ``ctx.sync()`` / ``wait_mark``."""
        assert out.dtype == self.dtype and out.size == len(self) and out.flags.c_contiguous
        _lib.call("djets_vec_collect_async", self.h, out.ctypes.data_as(C.c_void_p))
        return out

    # fill! and broadcasting (src:332-376) ---------------------------------------------------------------------
    def fill(self, a) -> "DBArray":
        if isinstance(a, (complex, np.complexfloating)):
            _lib.call("djets_vec_fill_c", self.h, float(a.real), float(a.imag))
        else:
            _lib.call("djets_vec_fill", self.h, float(a))
        return self

    def lincomb_(self, coefs: Sequence[float], xs: Sequence["DBArray"], wide: Optional[bool] = None) -> "DBArray":
        """``self .= c0.*x0 .+ c1.*x1 .+ ...`` fused in one pass.  ``wide``: evaluate in Float64 (Julia
        semantics when the scalars are Float64 and the arrays Float32); default: only if any operand is
        Float64."""
        if wide is None:
            wide = any(x.dtype in (np.float64, np.complex128) for x in xs)
        if any(isinstance(c, (complex, np.complexfloating)) and c.imag != 0 for c in coefs):
            raise DJetsError(_lib.ERR_UNSUPPORTED, "complex coefficients in a broadcast are outside the closed set")
        coefs = [c.real if isinstance(c, (complex, np.complexfloating)) else c for c in coefs]
        arr = (C.c_void_p * len(xs))(*[x.h for x in xs])
        _lib.call("djets_vec_lincomb", self.h, len(xs), (C.c_double * len(xs))(*[float(c) for c in coefs]), arr,
                  1 if wide else 0)
        return self

    def assign(self, src) -> "DBArray":
        """``self .= src``: a scalar (fill!, test/runtests.jl:329), a DBArray (element-type conversion,
        test/runtests.jl:237-238) or a lazy linear expression."""
        if isinstance(src, DBArray):
            return self.lincomb_([1.0], [src])
        if isinstance(src, LinExpr):
            return self.lincomb_(src.coefs, src.vecs, wide=True if src.float64_scalars else None)
        if isinstance(src, Expr):
            return self.eval_(src)
        if isinstance(src, Scalar):
            return self.eval_(Expr("sc", src))
        return self.fill(src)

    def eval_(self, expr: "Expr", wide: Optional[bool] = None) -> "DBArray":
        """``self .= expr`` for any expression over the closed operator set (src:363-375): lowered to a
        postfix program and evaluated by one streaming kernel.  ``wide``: compute in Float64 although every
        array is Float32 (default: when a Float64 scalar takes part, as Julia's promotion does)."""
        prog, vecs, scal, dscal, any64 = _lower(expr)
        if wide is None:
            wide = any64
        n = len(prog) // 2
        pbuf = (C.c_uint8 * len(prog))(*prog)
        varr = (C.c_void_p * max(len(vecs), 1))(*[v.h for v in vecs])
        sarr = (C.c_double * max(len(scal), 1))(*scal)
        darr = (C.c_void_p * max(len(dscal), 1))(*[d.h if d is not None else None for d in dscal])
        _lib.call("djets_vec_eval", self.h, pbuf, n, varr, len(vecs), sarr, darr, len(scal), 1 if wide else 0)
        return self

    def mul_(self, a: "DBArray", b: "DBArray") -> "DBArray":
        """``self .= a .* b``."""
        _lib.call("djets_vec_binary", self.h, 0, a.h, b.h)
        return self

    def div_(self, a: "DBArray", b: "DBArray") -> "DBArray":
        """``self .= a ./ b``."""
        _lib.call("djets_vec_binary", self.h, 1, a.h, b.h)
        return self

    # non-dotted arithmetic allocates like Julia's (similar(bc) ignores the element type: src:353-356)
    def __rmul__(self, a):
        if isinstance(a, (DBArray, Expr, Scalar)):
            return self.similar().eval_(lazy(self) * a)
        return self.similar().lincomb_([a], [self], wide=True)

    __mul__ = __rmul__

    def __add__(self, other):
        if not isinstance(other, DBArray):
            return self.similar().eval_(lazy(self) + other)
        return self.similar().lincomb_([1.0, 1.0], [self, other])

    def __radd__(self, other):
        return self.similar().eval_(other + lazy(self))

    def __sub__(self, other):
        if not isinstance(other, DBArray):
            return self.similar().eval_(lazy(self) - other)
        return self.similar().lincomb_([1.0, -1.0], [self, other])

    def __rsub__(self, other):
        return self.similar().eval_(other - lazy(self))

    def __truediv__(self, other):
        return self.similar().eval_(lazy(self) / other)

    def __rtruediv__(self, other):
        return self.similar().eval_(other / lazy(self))

    def __pow__(self, p):
        return self.similar().eval_(lazy(self) ** p)

    def __abs__(self):
        return self.similar().eval_(abs(lazy(self)))

    def __neg__(self):
        return self.similar().lincomb_([-1.0], [self])

    def copy(self) -> "DBArray":
        return self.similar().assign(self)

    # reductions (src:189-283) ------------------------------------------------------------------------------------
    def _reduce(self, kind: int, param: float = 0.0) -> float:
        out = C.c_double()
        _lib.call("djets_vec_reduce", self.h, kind, float(param), C.byref(out))
        return out.value

    def reduce_parts(self, kind: int, param: float = 0.0) -> np.ndarray:
        out = (C.c_double * len(self.pids))()
        _lib.call("djets_vec_reduce_parts", self.h, kind, float(param), out)
        return np.array(out, dtype=np.float64)

    def norm(self, p: float = 2):
        out = C.c_double()
        _lib.call("djets_vec_norm", self.h, float(p), C.byref(out))
        return np.zeros(0, self.dtype).real.dtype.type(out.value)      # float(real(T))

    def dot(self, other: "DBArray"):
        """``dot(x, y)`` (conjugates ``x`` for complex element types, like Julia's)."""
        if self.dtype.kind == "c":
            z = (C.c_double * 2)()
            _lib.call("djets_vec_dot_c", self.h, other.h, z)
            return self.dtype.type(complex(z[0], z[1]))
        out = C.c_double()
        _lib.call("djets_vec_dot", self.h, other.h, C.byref(out))
        return self.dtype.type(out.value)

    def dot_s(self, other: "DBArray", out: Optional[Scalar] = None) -> Scalar:
        """``dot(x, y)`` left on the devices (no host synchronisation)."""
        out = out or Scalar(None, self.dtype, self.ctx)
        _lib.call("djets_vec_dot_s", self.h, other.h, out.h)
        return out

    def norm_s(self, p: float = 2, out: Optional[Scalar] = None) -> Scalar:
        """``norm(x, p)`` left on the devices."""
        out = out or Scalar(None, self.dtype, self.ctx)
        _lib.call("djets_vec_norm_s", self.h, float(p), out.h)
        return out

    def reduce_s(self, kind: int, param: float = 0.0, out: Optional[Scalar] = None) -> Scalar:
        out = out or Scalar(None, self.dtype, self.ctx)
        _lib.call("djets_vec_reduce_s", self.h, int(kind), float(param), out.h)
        return out

    def sum(self):
        return self.dtype.type(self._reduce(_lib.RED_SUM))

    def minimum(self, f=None):
        return self.dtype.type(self._reduce(_lib.RED_ABSMIN if f is abs else _lib.RED_MIN))

    def maximum(self, f=None):
        return self.dtype.type(self._reduce(_lib.RED_ABSMAX if f is abs else _lib.RED_MAX))

    def extrema(self):                                    # src:227-237
        return self.minimum(), self.maximum()

    def mapreduce(self, f, op):
        """``mapreduce(f, op, x)`` for the closed set the reference exercises: f in {abs, identity},
        op in {+, max, min} (test/runtests.jl:242-263; bench)."""
        table = {(abs, "+"): _lib.RED_ABSSUM, (abs, "max"): _lib.RED_ABSMAX, (abs, "min"): _lib.RED_ABSMIN,
                 (None, "+"): _lib.RED_SUM, (None, "max"): _lib.RED_MAX, (None, "min"): _lib.RED_MIN}
        name = {max: "max", min: "min", np.maximum: "max", np.minimum: "min", np.add: "+"}.get(op, op)
        if name not in ("+", "max", "min") or (f is not None and f is not abs):
            raise DJetsError(_lib.ERR_UNSUPPORTED, "mapreduce: f in {abs, None}, op in {'+', max, min}")
        return self.dtype.type(self._reduce(table[(f, name)]))

    def mean(self):                                       # src:254-263
        return self.dtype.type(self.sum() / self.dtype.type(len(self)))

    def var(self, corrected: bool = True, mean=None):     # src:267-283
        mu = self.mean() if mean is None else mean
        s = self.dtype.type(self._reduce(_lib.RED_SUMSQ_SHIFT, float(mu)))
        return s / (len(self) - 1) if corrected else s / len(self)


class LinExpr:
    """Lazy ``c0.*x0 .+ c1.*x1 ...`` for the dotted (fused) broadcast forms."""

    def __init__(self, coefs, vecs, float64_scalars=True):
        self.coefs, self.vecs, self.float64_scalars = list(coefs), list(vecs), float64_scalars

    def __add__(self, other):
        o = other if isinstance(other, LinExpr) else LinExpr([1.0], [other])
        return LinExpr(self.coefs + o.coefs, self.vecs + o.vecs)

    def __sub__(self, other):
        o = other if isinstance(other, LinExpr) else LinExpr([1.0], [other])
        return LinExpr(self.coefs + [-c for c in o.coefs], self.vecs + o.vecs)

    def __rmul__(self, a):
        return LinExpr([a * c for c in self.coefs], self.vecs)


def dotted(a: float, x: DBArray) -> LinExpr:
    """``a .* x`` as a lazy term: ``d.assign(dotted(a1, d1) + dotted(a2, d2))``."""
    return LinExpr([a], [x])


# ------------------------------------------------------------------------------------------------------
# general lazy broadcast expressions (the dotted forms of src:363-375) and their lowering to the closed
# elementwise program of djets_vec_eval
# ------------------------------------------------------------------------------------------------------
class Expr:
    """Node of a lazy elementwise expression: ``lazy(x) * 2.0 + abs(lazy(y)) ** 2`` ...; evaluated by
    ``dest.assign(expr)`` in one pass over the operands."""
    __array_ufunc__ = None
    __slots__ = ("kind", "a")

    def __init__(self, kind: str, *a):
        self.kind, self.a = kind, a

    @staticmethod
    def wrap(v) -> "Expr":
        if isinstance(v, Expr):
            return v
        if isinstance(v, DBArray):
            return Expr("in", v)
        if isinstance(v, LinExpr):
            e = None
            for c, x in zip(v.coefs, v.vecs):
                t = Expr("bin", "*", Expr("sc", c), Expr("in", x))
                e = t if e is None else Expr("bin", "+", e, t)
            return e
        if isinstance(v, (Scalar, int, float, np.integer, np.floating)):
            return Expr("sc", v)
        raise DJetsError(_lib.ERR_UNSUPPORTED, f"broadcast operand of type {type(v).__name__} is outside the closed set")

    def _bin(self, op, other, swap=False):
        o = Expr.wrap(other)
        return Expr("bin", op, o, self) if swap else Expr("bin", op, self, o)

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __pow__(self, o): return self._bin("^", o)
    def __neg__(self): return Expr("un", "neg", self)
    def __abs__(self): return Expr("un", "abs", self)


def lazy(x) -> Expr:
    """Mark ``x`` (a DBArray, a number, a device ``Scalar``) as a lazy broadcast operand."""
    return Expr.wrap(x)


def bc_abs(x) -> Expr: return Expr("un", "abs", Expr.wrap(x))
def bc_sqrt(x) -> Expr: return Expr("un", "sqrt", Expr.wrap(x))
def bc_exp(x) -> Expr: return Expr("un", "exp", Expr.wrap(x))
def bc_log(x) -> Expr: return Expr("un", "log", Expr.wrap(x))
def bc_max(a, b) -> Expr: return Expr("bin", "max", Expr.wrap(a), Expr.wrap(b))
def bc_min(a, b) -> Expr: return Expr("bin", "min", Expr.wrap(a), Expr.wrap(b))


_BIN = {"+": _lib.EV_ADD, "-": _lib.EV_SUB, "*": _lib.EV_MUL, "/": _lib.EV_DIV, "^": _lib.EV_POW, "min": _lib.EV_MIN,
        "max": _lib.EV_MAX}
_BIN_S = {"+": _lib.EV_ADDS, "-": _lib.EV_SUBS, "*": _lib.EV_MULS, "/": _lib.EV_DIVS, "^": _lib.EV_POWS,
          "min": _lib.EV_MINS, "max": _lib.EV_MAXS}
_BIN_RS = {"+": _lib.EV_ADDS, "-": _lib.EV_RSUBS, "*": _lib.EV_RMULS, "/": _lib.EV_RDIVS, "min": _lib.EV_MINS,
           "max": _lib.EV_MAXS}
_UN = {"neg": _lib.EV_NEG, "abs": _lib.EV_ABS, "sqrt": _lib.EV_SQRT, "exp": _lib.EV_EXP, "log": _lib.EV_LOG}


def _is_f64_scalar(v) -> bool:
    if isinstance(v, Scalar):
        return v.dtype == np.float64
    return isinstance(v, (float, np.float64)) and not isinstance(v, np.float32)


def _lower(expr: Expr):
    """Expression tree -> (program bytes, input vectors, host scalars, device scalars, any Float64 operand)."""
    prog: List[int] = []
    vecs: List[DBArray] = []
    scal: List[float] = []
    dscal: List[Optional[Scalar]] = []
    state = {"f64": False}

    def vec_index(v: DBArray) -> int:
        for k, w in enumerate(vecs):
            if w is v:
                return k
        vecs.append(v)
        if v.dtype == np.float64:
            state["f64"] = True
        return len(vecs) - 1

    def scalar_index(v) -> int:
        if _is_f64_scalar(v):
            state["f64"] = True
        if isinstance(v, Scalar):
            for k, d in enumerate(dscal):
                if d is v:
                    return k
            scal.append(0.0)
            dscal.append(v)
        else:
            scal.append(float(v))
            dscal.append(None)
        return len(scal) - 1

    def is_sc(e: Expr) -> bool:
        return e.kind == "sc"

    def fold(op, x, y):
        return {"+": x + y, "-": x - y, "*": x * y, "/": x / y, "^": x ** y, "min": min(x, y), "max": max(x, y)}[op]

    def emit(e: Expr) -> None:
        if e.kind == "in":
            prog.extend((_lib.EV_IN, vec_index(e.a[0])))
        elif e.kind == "sc":
            prog.extend((_lib.EV_SC, scalar_index(e.a[0])))
        elif e.kind == "un":
            emit(e.a[1])
            prog.extend((_UN[e.a[0]], 0))
        else:
            op, l, r = e.a
            if is_sc(l) and is_sc(r) and not isinstance(l.a[0], Scalar) and not isinstance(r.a[0], Scalar):
                emit(Expr("sc", fold(op, l.a[0], r.a[0])))          # scalar (op) scalar: folded like Julia does per element
            elif is_sc(r) and not is_sc(l):
                emit(l)
                v = r.a[0]
                if op == "^" and not isinstance(v, Scalar) and float(v) == 2.0:
                    prog.extend((_lib.EV_SQUARE, 0))                # Julia's literal_pow: x^2 = x*x
                elif op == "^" and not isinstance(v, Scalar) and float(v) == 3.0:
                    prog.extend((_lib.EV_CUBE, 0))                  # x^3 = x*x*x
                else:
                    prog.extend((_BIN_S[op], scalar_index(v)))
            elif is_sc(l) and not is_sc(r) and op in _BIN_RS:
                emit(r)
                prog.extend((_BIN_RS[op], scalar_index(l.a[0])))
            else:
                emit(l)
                emit(r)
                prog.extend((_BIN[op], 0))

    emit(expr)
    return prog, vecs, scal, dscal, state["f64"]


# free-function spellings used by the reference's tests ---------------------------------------------------------
def zeros(R: JetDSpace) -> DBArray:
    return R.zeros()


def ones(R: JetDSpace) -> DBArray:
    return R.ones()


def rand(R: JetDSpace, rng=None) -> DBArray:
    return R.rand(rng)


def Array(R: JetDSpace) -> DBArray:
    return R.Array()


def getblock(x, *idx):
    return x.getblock(*idx)


def getblock_into(x: DBArray, i: int, out: np.ndarray):
    return x.getblock_into(i, out)


def setblock(x: DBArray, i: int, b) -> None:
    x.setblock(i, b)


def collect(x: DBArray):
    return x.collect()


def dot(x: DBArray, y: DBArray):
    return x.dot(y)


def norm(x: DBArray, p: float = 2):
    return x.norm(p)


def blockmap(x):
    return x.blockmap()


def localblockindices(x, *a, **k):
    return x.localblockindices(*a, **k)


def procs(x):
    return x.procs()


def nprocs(x):
    return x.nprocs()


def nblocks(x):
    return x.nblocks()


# ======================================================================================================
# block kinds: what a CUDA library can recognise in place of arbitrary Jets closures
# ======================================================================================================
class JopDense:
    """Dense block ``d .= A*m`` / ``m .= A'*d`` (the reference's dense fixture JopBaz, test/runtests.jl:30-36)."""
    kind = _lib.BLOCK_DENSE

    def __init__(self, A: np.ndarray):
        if A.ndim != 2:
            raise ValueError("JopDense needs a matrix")
        self.A = A
        self.dtype = A.dtype
        self.rng_shape, self.dom_shape = (A.shape[0],), (A.shape[1],)


class JopDiagonal:
    """Diagonal block ``d .= diag .* m`` (JopFoo test/runtests.jl:8-12; JetPack's JopDiagonal, docs:19).
    N-d blocks keep their shape."""
    kind = _lib.BLOCK_DIAG

    def __init__(self, diag: np.ndarray):
        self.diag = diag
        self.dtype = diag.dtype
        self.rng_shape = self.dom_shape = tuple(diag.shape)


class JopZeroBlock:
    """``JopZeroBlock(dom, rng)``: the only kind for which ``iszero`` holds (skipped: src:587,634,681)."""
    kind = _lib.BLOCK_ZERO

    def __init__(self, dtype, rng_shape, dom_shape):
        self.dtype = np.dtype(dtype)
        self.rng_shape, self.dom_shape = _shape(rng_shape), _shape(dom_shape)


def iszero(op) -> bool:
    return op.kind == _lib.BLOCK_ZERO


# ======================================================================================================
# the distributed block operator (src:414-721, 805-861)
# ======================================================================================================
class JopDBlock:
    """``@blockop DArray(I->[blk(i,j) ...], (N,M), pids, [n1,n2]) isdiag=...`` for linear blocks.

    ``blocks`` is either a nested list ``blocks[i-1][j-1]`` or a callable ``(i, j) -> block`` with
    ``dims=(N, M)`` (1-based); like the reference's DArray constructor the callable is what builds the
    payload, and only the owner's payload is uploaded.  ``shapes=(row_shapes, col_shapes)`` lets
    non-owners skip building payloads altogether."""

    def __init__(self, blocks, dims: Optional[Tuple[int, int]] = None, pids: Optional[Sequence[int]] = None,
                 dist: Optional[Sequence[int]] = None, isdiag: bool = False,
                 row_ranges: Optional[Sequence[range]] = None, col_ranges: Optional[Sequence[range]] = None,
                 shapes=None, dtype=None, ctx: Optional[Context] = None, perfstatfile: str = ""):
        ctx = ctx or context()
        self.ctx = ctx
        self.perfstatfile = perfstatfile
        if perfstatfile and os.path.exists(perfstatfile):
            os.remove(perfstatfile)                       # src:505
        if callable(blocks):
            if dims is None:
                raise ValueError("dims=(N, M) is needed when blocks is a callable")
            N, M = int(dims[0]), int(dims[1])
            get = blocks
        else:
            N, M = len(blocks), len(blocks[0])
            get = lambda i, j: blocks[i - 1][j - 1]      # noqa: E731
        pids = list(pids if pids is not None else ctx.workers())
        if row_ranges is not None:                       # irregular chunks (src:8-17)
            rcuts, ccuts = cuts_from_ranges(row_ranges), cuts_from_ranges(col_ranges)
            n1, n2 = len(row_ranges), len(col_ranges)
        else:
            if dist is None:
                dist = default_grid([N, M], len(pids))
            n1, n2 = int(dist[0]), int(dist[1])
            rcuts, ccuts = even_cuts(N, n1), even_cuts(M, n2)
        if isdiag and n2 != 1:
            raise DJetsError(_lib.ERR_INVALID, "block diagonal jets must define distribution along rows.")  # src:481
        self.N, self.M, self.n1, self.n2, self.isdiag = N, M, n1, n2, bool(isdiag)
        self.pids = grid_of(pids, n1, n2)                # src:483
        self._rcuts, self._ccuts = rcuts, ccuts
        self._blockmap = [[(urange(rcuts[r] + 1, rcuts[r + 1]), urange(ccuts[c] + 1, ccuts[c + 1]))
                           for c in range(n2)] for r in range(n1)]          # src:507 indices(ops)

        def owner(i, j):
            r = max(k for k in range(n1) if rcuts[k] < i)
            c = max(k for k in range(n2) if ccuts[k] < j)
            return self.pids[r][c]

        # shapes of the block rows (from column 1) and block columns (from row 1): src:485-493.  With
        # isdiag the domain block of row i comes from that row's column-1 block (src:471-473).
        built = {}

        def blk(i, j):
            if (i, j) not in built:
                built[(i, j)] = get(i, j)
            return built[(i, j)]

        if shapes is not None:
            self.row_shapes = [_shape(s) for s in shapes[0]]
            self.col_shapes = [_shape(s) for s in shapes[1]]
        else:
            self.row_shapes = [_shape(blk(i, 1).rng_shape) for i in range(1, N + 1)]
            if isdiag:
                self.col_shapes = [_shape(blk(i, 1).dom_shape) for i in range(1, N + 1)]
            else:
                self.col_shapes = [_shape(blk(1, j).dom_shape) for j in range(1, M + 1)]
            dtype = dtype or blk(1, 1).dtype
        self.dtype = np.dtype(dtype)
        if isdiag and N != M:
            raise DJetsError(_lib.ERR_INVALID, "block diagonal operator must be square in blocks")

        grid_flat = [self.pids[r][c] for c in range(n2) for r in range(n1)]
        h = C.c_void_p()
        _lib.call("djets_op_create", ctx.h, _dt(self.dtype), N, M, _lib.arr_i64([_size(s) for s in self.row_shapes]),
                  _lib.arr_i64([_size(s) for s in self.col_shapes]), n1, n2, _lib.arr_i32(grid_flat),
                  _lib.arr_i64(rcuts), _lib.arr_i64(ccuts), 1 if isdiag else 0, C.byref(h))
        self.h = h
        self._kinds = {}
        for i in range(1, N + 1):
            cols = [i] if isdiag else range(1, M + 1)
            if isdiag and shapes is None:                # src:466-470
                nz = sum(0 if iszero(blk(i, j)) else 1 for j in range(1, M + 1))
                if nz != 1:
                    raise DJetsError(_lib.ERR_INVALID,
                                     f"Expecting one non-zero column block per row in block diagonal construction, got {nz}.")
            for j in cols:
                if not ctx.is_local(owner(i, j)):
                    continue
                b = blk(i, j)
                self._kinds[(i, j)] = b.kind
                if b.kind == _lib.BLOCK_DENSE:
                    A = b.A
                    if A.dtype != self.dtype:
                        raise DimensionMismatch(_lib.ERR_MISMATCH, "block element type differs from the operator's")
                    if not A.flags.f_contiguous:
                        A = np.asfortranarray(A)
                    _lib.call("djets_op_set_dense", h, i - 1, j - 1, A.ctypes.data_as(C.c_void_p), max(A.shape[0], 1))
                elif b.kind == _lib.BLOCK_DIAG:
                    d = np.asfortranarray(b.diag, dtype=self.dtype).reshape(-1, order="F")
                    _lib.call("djets_op_set_diag", h, i - 1, j - 1, d.ctypes.data_as(C.c_void_p))
                built.pop((i, j), None)
        built.clear()
        _lib.call("djets_op_finalize", h)
        if perfstatfile:
            _lib.call("djets_op_perfstat_enable", h, 1)
        self._range = self._space(True)
        self._domain = self._space(False)

    def _space(self, is_range: bool) -> JetDSpace:
        if is_range or self.isdiag:
            cuts, pids = self._rcuts, [self.pids[r][0] for r in range(self.n1)]       # src:485, 491
        else:
            cuts, pids = self._ccuts, [self.pids[0][c] for c in range(self.n2)]       # src:493
        shapes = self.row_shapes if is_range else self.col_shapes
        return JetDSpace([shapes[cuts[k]:cuts[k + 1]] for k in range(len(cuts) - 1)], pids, self.dtype, self.ctx)

    def __del__(self):
        self.close()

    def close(self) -> None:                              # src:854-861
        h, self.h = getattr(self, "h", None), None
        if h and getattr(self.ctx, "h", None):
            try:
                _lib.lib.djets_op_destroy(h)
            except Exception:
                pass

    # accessors (src:805-852) ----------------------------------------------------------------------------------
    def range(self) -> JetDSpace:
        return self._range

    def domain(self) -> JetDSpace:
        return self._domain

    @property
    def shape(self):
        return (len(self._range), len(self._domain))

    def size(self):
        return self.shape

    def nblocks(self, dim: Optional[int] = None):
        """``nblocks(A)`` / ``nblocks(A, dim)`` (test/runtests.jl:381-383)."""
        return (self.N, self.M) if dim is None else (self.N, self.M)[dim - 1]

    def isblockop(self) -> bool:                           # src:827
        return True

    def procs(self) -> List[List[int]]:
        return self.pids

    def nprocs(self) -> int:
        return self.n1 * self.n2

    def blockmap(self):
        return self._blockmap

    def localblockindices(self, pid: Optional[int] = None, dim: Optional[int] = None):   # src:814-825
        pid = _mypid(self.ctx, pid)
        for c in range(self.n2):
            for r in range(self.n1):
                if self.pids[r][c] == pid:
                    bm = self._blockmap[r][c]
                    return bm if dim is None else bm[dim - 1]
        bm = (urange(0, 0), urange(0, 0))
        return bm if dim is None else bm[dim - 1]

    def block_owner(self, i: int, j: int) -> Tuple[int, int, int]:
        r, c, rank = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.call("djets_op_block_owner", self.h, i - 1, j - 1, C.byref(r), C.byref(c), C.byref(rank))
        return r.value + 1, c.value + 1, rank.value

    def getblock(self, i: int, j: int):
        """``getblock(A, i, j)`` (src:836-852): the block operator, payload copied back from its owner."""
        if not (1 <= i <= self.N and 1 <= j <= self.M):
            raise IndexError((i, j))
        kind = C.c_int32()
        _lib.call("djets_op_getblock", self.h, i - 1, j - 1, C.byref(kind), None, 0)
        rs, cs = self.row_shapes[i - 1], self.col_shapes[j - 1]
        if kind.value == _lib.BLOCK_DENSE:
            A = np.empty((_size(rs), _size(cs)), dtype=self.dtype, order="F")
            _lib.call("djets_op_getblock", self.h, i - 1, j - 1, C.byref(kind), A.ctypes.data_as(C.c_void_p),
                      max(A.shape[0], 1))
            return JopDense(A)
        if kind.value == _lib.BLOCK_DIAG:
            d = np.empty(rs, dtype=self.dtype, order="F")
            _lib.call("djets_op_getblock", self.h, i - 1, j - 1, C.byref(kind), d.ctypes.data_as(C.c_void_p), 0)
            return JopDiagonal(d)
        return JopZeroBlock(self.dtype, rs, cs)

    def algorithmic_bytes(self) -> int:
        out = C.c_int64()
        _lib.call("djets_op_algorithmic_bytes", self.h, C.byref(out))
        return int(out.value)

    def replan(self) -> None:
        """Rebuild the tile schedules after ``ctx.set_param`` (tuning only)."""
        _lib.call("djets_op_replan", self.h)

    def schedule_info(self, adjoint: bool = False) -> Tuple[int, int]:
        nt, ni = C.c_int64(), C.c_int64()
        _lib.call("djets_op_schedule_info", self.h, 1 if adjoint else 0, C.byref(nt), C.byref(ni))
        return int(nt.value), int(ni.value)

    # apply (src:582-721) ------------------------------------------------------------------------------------------
    def _replicated_domain(self) -> DBArray:
        """Device-side home of a replicated (plain-array) model: the tall-and-skinny case n2 == 1, where
        the reference's domain is a local JetBSpace on the master and ``m`` is broadcast to every worker
        (src:510-557).  Here it is the single domain part on grid cell (1,1); phase 0 of the exchange
        plan forwards it to the other grid rows."""
        if self.n2 != 1 or self.isdiag:
            raise DimensionMismatch(_lib.ERR_MISMATCH, "a plain-array model needs a tall-and-skinny operator "
                                                       "(grid n x 1, not block diagonal): src:523-557")
        if getattr(self, "_mrep", None) is None:
            self._mrep = self._domain.zeros()
        return self._mrep

    def _perfstat(self, operation: str) -> None:
        """``Jets.perfstat(ops, operation, perfstatfile)`` (src:723-745): append one step to the JSON file,
        one entry per worker with ``id``, ``hostname``, ``operation`` and ``localblock`` -- here the
        milliseconds of that worker's share of the apply, split over its local blocks (column-major, like
        the reference's ``[...][:]``) in proportion to their algorithmic bytes."""
        if not self.perfstatfile:
            return
        import json
        import socket
        es = self.dtype.itemsize
        entries = []
        for c in range(self.n2):                          # procs(ops) order: column-major grid
            for r in range(self.n1):
                if not self.ctx.is_local(self.pids[r][c]):
                    continue
                ms = C.c_double()
                _lib.call("djets_op_perfstat", self.h, r, c, C.byref(ms))
                rows, cols = self._blockmap[r][c]
                w = []
                for j in cols:
                    for i in rows:
                        k = self._kinds.get((i, j), _lib.BLOCK_ZERO)
                        m_, n_ = _size(self.row_shapes[i - 1]), _size(self.col_shapes[j - 1])
                        w.append(float(m_ * n_ * es) if k == _lib.BLOCK_DENSE else float(n_ * es) if k == _lib.BLOCK_DIAG else 0.0)
                tot = sum(w) or 1.0
                entries.append({"id": self.pids[r][c], "hostname": socket.gethostname(), "operation": operation,
                                "localblock": [ms.value * v / tot for v in w]})
        if os.path.isfile(self.perfstatfile):
            with open(self.perfstatfile) as f:
                doc = json.load(f)
            doc["step"].append({"pid": entries})
        else:
            doc = {"step": [{"pid": entries}]}
        with open(self.perfstatfile, "w") as f:
            json.dump(doc, f, indent=1)

    def mul(self, d: DBArray, m) -> DBArray:
        """``mul!(d, A, m)``; ``m`` may be a plain array when the domain is replicated (src:544-557)."""
        if isinstance(m, np.ndarray):
            m = self._replicated_domain().scatter(np.ascontiguousarray(m.reshape(-1, order="F"), dtype=self.dtype))
        _lib.call("djets_op_mul", self.h, d.h, m.h)
        self._perfstat("df")
        return d

    def mul_adjoint(self, m, d: DBArray):
        """``mul!(m, A', d)``; a plain-array ``m`` receives the sum over workers (src:565-580)."""
        if isinstance(m, np.ndarray):
            rep = self._replicated_domain()
            _lib.call("djets_op_mul_adjoint", self.h, rep.h, d.h)
            self._perfstat("df\u2032")
            flat = rep.to_array()
            m[...] = flat.reshape(m.shape, order="F")
            return m
        _lib.call("djets_op_mul_adjoint", self.h, m.h, d.h)
        self._perfstat("df\u2032")
        return m

    def __matmul__(self, m) -> DBArray:
        """``A*m = mul!(zeros(range(A)), A, m)``."""
        return self.mul(self._range.zeros(), m)

    @property
    def T(self) -> "JopAdjoint":
        return JopAdjoint(self)

    def adjoint(self) -> "JopAdjoint":
        return JopAdjoint(self)


class JopAdjoint:
    """``A'``."""

    def __init__(self, op: JopDBlock):
        self.op = op

    def range(self):
        return self.op.domain()

    def domain(self):
        return self.op.range()

    @property
    def shape(self):
        return self.op.shape[::-1]

    def mul(self, m: DBArray, d: DBArray) -> DBArray:
        return self.op.mul_adjoint(m, d)

    def __matmul__(self, d: DBArray) -> DBArray:
        return self.op.mul_adjoint(self.op.domain().zeros(), d)

    @property
    def T(self):
        return self.op

    def adjoint(self):
        return self.op


def blockop(blocks, *args, **kw) -> JopDBlock:
    """``@blockop`` over a distributed array of linear blocks (src:414-415)."""
    return JopDBlock(blocks, *args, **kw)


def mul(out: DBArray, A, x: DBArray) -> DBArray:
    """``mul!(out, A, x)`` for ``A`` a JopDBlock or its adjoint."""
    return A.mul(out, x)
