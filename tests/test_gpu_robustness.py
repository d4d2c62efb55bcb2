"""Randomised layouts, error behaviour and resource hygiene of the C-ABI path (GPU)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_random_dbarray_layouts_against_oracle(dj, O, ctx4):
    """Thirty random DBArrays (1-4 workers, ragged blocks incl. empty and N-d ones, both element types):
    block access / collect / scatter bit-exact, reductions within tolerance, broadcasts bit-exact."""
    master = np.random.default_rng(99)
    for case in range(30):
        nw = int(master.integers(1, 5))
        nblk = int(master.integers(nw, nw + 6))
        dtype = np.float64 if master.integers(2) else np.float32
        shapes = []
        for _ in range(nblk):
            k = master.integers(5)
            shapes.append((0,) if k == 0 else (int(master.integers(1, 40)), int(master.integers(1, 6))) if k == 1
                          else (int(master.integers(1, 3000)),))
        cuts = sorted(master.choice(np.arange(1, nblk), size=nw - 1, replace=False).tolist()) if nw > 1 else []
        bounds = [0] + cuts + [nblk]
        blkidx = [range(bounds[k] + 1, bounds[k + 1] + 1) for k in range(nw)]
        rng = np.random.default_rng(case)
        data = {i + 1: np.asfortranarray((rng.random(s) - 0.5).astype(dtype)) for i, s in enumerate(shapes)}
        x = dj.DBArray(lambda i: data[i], blkidx, list(range(nw)))
        ox = O.ODBArray.from_function(lambda i: data[i], blkidx, list(range(nw)))
        assert x.indices == ox.indices and x.blkindices == ox.blkindices and len(x) == len(ox)
        flat = ox.to_array()
        assert np.array_equal(x.to_array(), flat)
        for ib in range(1, nblk + 1):
            g = x.getblock(ib)
            assert g.shape == shapes[ib - 1] and np.array_equal(g, data[ib])
        if len(x) == 0:
            continue
        j = int(master.integers(1, len(x) + 1))
        assert x[j] == flat[j - 1]
        rt = 1e-12 if dtype == np.float64 else 3e-6
        wide = flat.astype(np.float64)
        np.testing.assert_allclose(float(x.sum()), wide.sum(), rtol=rt, atol=rt * np.abs(wide).sum())
        np.testing.assert_allclose(float(x.dot(x)), (wide * wide).sum(), rtol=rt)
        np.testing.assert_allclose(float(x.norm()), math.sqrt((wide * wide).sum()), rtol=rt)
        np.testing.assert_allclose(float(x.norm(1)), np.abs(wide).sum(), rtol=rt)
        if any(len(r) == 0 for r in x.indices):                 # a worker with an empty local part: Julia's extrema / maximum throw
            with pytest.raises(dj.DJetsError, match="empty collection"):
                x.extrema()
        else:
            assert x.extrema() == (flat.min(), flat.max()) and x.maximum(abs) == np.abs(flat).max()
        y = x.similar().scatter(np.ascontiguousarray(flat[::-1]))
        out = x.similar().lincomb_([0.25, -3.0], [x, y], wide=True)
        want = (0.25 * flat.astype(np.float64) + (-3.0) * flat[::-1].astype(np.float64)).astype(dtype)
        assert np.array_equal(out.to_array(), want)
        assert np.array_equal(x.similar(np.float32).assign(x).to_array(), flat.astype(np.float32))


def test_error_behaviour(dj, ctx4, ctx1):
    x = dj.DBArray(lambda i: np.ones(4), (2,), [0, 1], ctx=ctx4)
    y = dj.DBArray(lambda i: np.ones(4), (2,), [0], ctx=ctx1)
    with pytest.raises(dj.DimensionMismatch):
        x.dot(y)                                            # different layout (and context)
    with pytest.raises(dj.DimensionMismatch):
        x.dot(x.similar(np.float32))                        # dot(::DBArray{T}, ::DBArray{T}) only (src:214)
    with pytest.raises(dj.DimensionMismatch):
        x.setblock(1, np.ones(5))
    with pytest.raises(IndexError):
        x.getblock(0)
    with pytest.raises(dj.DJetsError):
        x.lincomb_([1.0] * 5, [x] * 5)                      # more than four terms
    with pytest.raises(dj.DJetsError):
        ctx4.set_param("fwd_tx", 48)
    with pytest.raises(dj.DJetsError):
        ctx4.set_param("no_such_knob", 1)
    with pytest.raises(dj.DJetsError):
        dj.JetDSpace([[(4,)]], [7], np.float64, ctx=ctx4).zeros()       # rank outside the job
    with pytest.raises(dj.DJetsError):
        dj.JopDBlock([[dj.JopDense(np.ones((3, 3)))]], pids=[0, 1, 2, 3], dist=[2, 1], ctx=ctx4)   # grid larger than the blocks
    with pytest.raises(dj.DJetsError):
        dj.JopDBlock([[dj.JopDiagonal(np.ones(3)), dj.JopDense(np.ones((3, 4)))],
                      [dj.JopDense(np.ones((5, 3))), dj.JopDiagonal(np.ones(4))]], pids=[0], dist=[1, 1], ctx=ctx4)  # 5x4 "diagonal"
    A = dj.JopDBlock([[dj.JopDense(np.ones((3, 3)))]], pids=[0], dist=[1, 1], ctx=ctx4)
    import ctypes as C
    import djets_b200._lib as L
    buf = np.ones((3, 3), order="F")
    assert L.lib.djets_op_set_dense(A.h, 0, 0, buf.ctypes.data_as(C.c_void_p), 3) == L.ERR_INVALID   # after finalize
    assert L.lib.djets_op_set_dense(A.h, 5, 0, buf.ctypes.data_as(C.c_void_p), 3) == L.ERR_INVALID   # outside the operator
    assert L.lib.djets_op_mul(A.h, None, None) == L.ERR_INVALID
    assert b"null" in L.lib.djets_last_error()


def test_no_device_memory_leak_over_many_operators(dj, ctx1):
    """Two hundred operator + vector lifetimes leave device memory where it started (free lists included)."""
    import torch
    rng = np.random.default_rng(0)
    blocks = [np.asfortranarray(rng.random((256, 192))) for _ in range(4)]

    def lifetime():
        A = dj.JopDBlock(lambda i, j: dj.JopDense(blocks[(i + j) % 4]) if (i + j) % 3 else dj.JopZeroBlock(np.float64, 256, 192),
                         dims=(3, 4), pids=[0], dist=[1, 1], ctx=ctx1)
        x = A.domain().rand(rng)
        y = A @ x
        z = A.T @ y
        s = float(z.dot(z))
        A.close()
        return s

    first = lifetime()
    ctx1.sync()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(200):
        assert lifetime() > 0
    ctx1.sync()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 64 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB"
    assert first > 0
