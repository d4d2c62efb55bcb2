"""Pins the oracle (oracle/djets_oracle.py) against every literal assertion and known-answer value the
reference's own test-suite holds for the hot path (test/runtests.jl, cited per test).  CPU only.

The reference's floating-point checks are self-consistency checks under isapprox (distributed result
vs the serial Jets block operator over the same unseeded random blocks); they are reproduced here in
the same form, plus the higher-precision ground truth at BASELINE.json's tolerances."""
import math

import numpy as np
import pytest

from oracle import djets_oracle as O

PI = math.pi
ur = O.ur
W = [2, 3, 4, 5]          # Julia worker pids after addprocs(4) (test:4)


def diag_op(shape, nblk, pids, dist=None, seed=0):
    """@blockop DArray(I->[JopFoo(rand(shape)) ...], (nblk,1), pids[, dist]) (test:8-12)."""
    rng = np.random.default_rng(seed)
    blocks = [[O.ODiag(rng.random(shape))] for _ in range(nblk)]
    return O.OJopDBlock(blocks, pids, dist)


def test_irregular_darray_construction():
    """test:38-53: DArray(init, pids, [1:2,3:10], [1:2])."""
    idxs, cuts = O.chunk_idxs_prescribed([ur(1, 2), ur(3, 10)], [ur(1, 2)])
    assert idxs[(1, 1)] == (ur(1, 2), ur(1, 2))
    assert idxs[(2, 1)] == (ur(3, 10), ur(1, 2))
    assert O.indices_from_cuts(cuts[0]) == [ur(1, 2), ur(3, 10)]
    assert O.indices_from_cuts(cuts[1]) == [ur(1, 2)]
    assert (cuts[0][-1] - 1, cuts[1][-1] - 1) == (10, 2)          # size(A) == (10,2)
    idxs, cuts = O.chunk_idxs_prescribed([ur(1, 2), ur(3, 10)])
    assert idxs[(1,)] == (ur(1, 2),) and idxs[(2,)] == (ur(3, 10),)


def test_jetdspace_from_jetbspaces():
    """test:55-63."""
    R = O.OJetDSpace([O.OSpace(np.float64, [2, 2]), O.OSpace(np.float64, [2, 2])], W[:2])
    assert len(R) == 8 and R.nblocks() == 4
    assert R.blockmap() == [ur(1, 2), ur(3, 4)]
    assert R.indices == [ur(1, 4), ur(5, 8)]


@pytest.mark.parametrize("shape,n,idx", [((2,), 4, [ur(1, 2), ur(3, 4)]), ((2, 3), 12, [ur(1, 6), ur(7, 12)])])
def test_jetdspace_construction(shape, n, idx):
    """test:65-103."""
    A = diag_op(shape, 2, W[:2])
    R = A.rng
    assert R.size() == (n,) and len(R) == n and R.dtype == np.float64
    assert R.indices == idx and R.procs() == W[:2] and R.nprocs() == 2 and A.nprocs() == 2 and R.nblocks() == 2
    assert R == R.rand(np.random.default_rng(0)).space()


@pytest.mark.parametrize("shape,nblk", [((2,), 2), ((2, 3), 4)])
def test_jetdspace_operations(shape, nblk):
    """test:105-170."""
    A = diag_op(shape, nblk, W[:2])
    R = A.rng
    n = nblk * int(np.prod(shape))
    assert np.array_equal(R.zeros().to_array(), np.zeros(n)) and np.array_equal(R.ones().to_array(), np.ones(n))
    d = R.array()
    assert d.size() == (n,)
    assert d.indices == [ur(1, n // 2), ur(n // 2 + 1, n)]           # == drand((n,), 2 workers).indices
    assert d.blockmap() == [ur(1, nblk // 2), ur(nblk // 2 + 1, nblk)]
    assert d.similar().dtype == np.float64 and d.similar(np.float32).dtype == np.float32
    assert isinstance(d.similar(np.float32, len(d)), O.ODBArray)
    assert R.localindices(W[0]) == ur(1, n // 2) and R.localindices(W[1]) == ur(n // 2 + 1, n)
    assert R.localblockindices(W[0]) == ur(1, nblk // 2) and R.localblockindices(W[1]) == ur(nblk // 2 + 1, nblk)
    x = d.getblock(1)
    x[...] = PI
    d.setblock(1, x)
    m = int(np.prod(shape))
    assert [d.getindex(i) for i in range(1, m + 1)] == [PI] * m
    x[...] = 0
    assert np.array_equal(d.getblock_into(1, x), np.full(shape, PI)) and x.shape == shape


def test_dbarray_construction_known_answers():
    """test:172-221."""
    foo = lambda i: i * PI * np.ones(2 if i % 2 == 0 else 3)
    A = O.ODBArray.from_nblocks(foo, 7, W, [4])
    B = O.ODBArray.from_nblocks(foo, 7, W)
    D = O.ODBArray.from_function(foo, A.blockmap(), W)
    E = O.ODBArray.from_function(foo, [ur(1, 1), ur(2, 2), ur(3, 3), ur(4, 7)], W)
    assert A.blkindices == [ur(1, 2), ur(3, 4), ur(5, 6), ur(7, 7)]
    assert A.blkindices == B.blkindices == D.blkindices != E.blkindices
    assert A.indices == B.indices == D.indices != E.indices
    assert len(A) == 18
    for i in range(1, 19):
        assert A.getindex(i) == E.getindex(i)
    for ib in range(1, 8):
        m = 2 if ib % 2 == 0 else 3
        for v in (A, B, D, E):
            assert np.array_equal(v.getblock(ib), ib * PI * np.ones(m))


def test_default_dist_matches_asserted_splits():
    """Even-split rule against every split the reference asserts literally (test:157,384-385,536-539)."""
    assert O.indices_from_cuts(O.defaultdist_cuts(3, 2)) == [ur(1, 2), ur(3, 3)]
    assert O.indices_from_cuts(O.defaultdist_cuts(5, 2)) == [ur(1, 3), ur(4, 5)]
    assert O.indices_from_cuts(O.defaultdist_cuts(7, 4)) == [ur(1, 2), ur(3, 4), ur(5, 6), ur(7, 7)]
    assert O.indices_from_cuts(O.defaultdist_cuts(10, 4)) == [ur(1, 3), ur(4, 6), ur(7, 8), ur(9, 10)]
    assert O.indices_from_cuts(O.defaultdist_cuts(50, 3)) == [ur(1, 17), ur(18, 34), ur(35, 50)]
    assert O.grid_pids(W, 2, 2) == [[2, 4], [3, 5]]                   # test:543 [w1 w3; w2 w4]


def test_reductions_against_flat_array():
    """test:223-304, same form: f(d) ~ f(convert(Array, d))."""
    A = diag_op((2, 3), 4, W)
    rng = np.random.default_rng(5)
    d, d2 = A.rng.rand(rng), A.rng.rand(rng)
    f, f2 = d.to_array(), d2.to_array()
    np.testing.assert_allclose(d.dot(d2), np.dot(f, f2), rtol=1.5e-8)
    np.testing.assert_allclose(d.norm(), np.linalg.norm(f), rtol=1.5e-8)
    assert d.norm(0) == np.count_nonzero(f) and d.norm(math.inf) == np.abs(f).max()
    d32 = d.similar(np.float32).broadcast_assign(lambda a: a, d)      # _d .= d (test:237-238)
    assert type(d32.norm(0.5)) == np.float32
    np.testing.assert_allclose(d.mapreduce(np.abs, np.add), np.abs(f).sum(), rtol=1.5e-8)
    assert d.mapreduce(np.abs, np.maximum) == np.abs(f).max()
    assert d.extrema() == (f.min(), f.max())
    np.testing.assert_allclose(d.mean(), f.mean(), rtol=1.5e-8)
    np.testing.assert_allclose(d.var(), f.var(ddof=1), rtol=1.5e-8)
    np.testing.assert_allclose(d.var(mean=100.0), np.sum((f - 100.0) ** 2) / (f.size - 1), rtol=1.5e-8)
    np.testing.assert_allclose(d.var(corrected=False), f.var(), rtol=1.5e-8)


@pytest.mark.parametrize("shape", [(2,), (2, 3)])
def test_broadcasting(shape):
    """test:306-363: 7 blocks, 2 workers, dist [2,1]."""
    A = diag_op(shape, 7, W[:2], [2, 1])
    R = A.rng
    assert R.blkindices == [ur(1, 4), ur(5, 7)]
    d1, d2, d3, d = R.ones(), R.ones(), R.ones(), R.ones()
    a1, a2, a3 = np.random.default_rng(1).random(3)
    d.broadcast_assign(lambda x, y, z: a1 * x + a2 * y + a3 * z, d1, d2, d3)
    np.testing.assert_allclose(d.to_array(), a1 + a2 + a3, rtol=1.5e-8)
    d1.fill(3.14)
    assert all(d1.getindex(i) == 3.14 for i in range(1, len(d1) + 1))


def hetero_blocks(rng, N, M, dtype=np.float64):
    """A linear stand-in for the reference's heterogeneous chooser myblocks(i,j) (test:365-375): its
    nonlinear JopBar blocks become diagonal blocks (their Jacobian IS diagonal: 2 .* m0 .* dm, test:23),
    the zero block stays, the rest are dense JopBaz blocks."""
    blocks = []
    for i in range(1, N + 1):
        row = []
        for j in range(1, M + 1):
            if i in (1, 2, 3) and j in (1, 2, 3):
                row.append(O.ODiag(2.0 * rng.random(10).astype(dtype)))
            elif (i, j) == (3, 4):
                row.append(O.OZero(dtype, (10,), (10,)))
            else:
                row.append(O.ODense(np.asfortranarray(rng.random((10, 10)).astype(dtype))))
        blocks.append(row)
    return blocks


def check_against_serial(op, blocks, rng, rtol=1.5e-8):
    """collect(J*dm) ~ _J*dm and J'*dd ~ _J'*collect(dd) (test:418-428, 574-583), then BASELINE tolerance."""
    x = op.dom.rand(rng)
    y = op.mul(op.rng.zeros(), x)
    np.testing.assert_allclose(y.to_array(), O.serial_blockop_mul(blocks, x.to_array()), rtol=rtol)
    lens = [b[0].m for b in blocks]
    tol = 1e-12 if op.dtype == np.float64 else 1e-5
    assert O.blockwise_relerr(y.to_array(), O.ground_truth_mul(blocks, x.to_array()), lens) <= tol
    d = op.rng.rand(rng)
    m = op.mul_adjoint(op.dom.zeros(), d)
    np.testing.assert_allclose(m.to_array(), O.serial_blockop_mul_adjoint(blocks, d.to_array()), rtol=rtol)
    clens = [b.n for b in blocks[0]]
    assert O.blockwise_relerr(m.to_array(), O.ground_truth_mul(blocks, d.to_array(), adjoint=True), clens) <= tol


def test_operator_tall_and_skinny():
    """test:377-436: 3x4 blocks of size 10 on grid [2,1]."""
    rng = np.random.default_rng(2)
    blocks = hetero_blocks(rng, 3, 4)
    F = O.OJopDBlock(blocks, W[:2], [2, 1])
    assert F.nblocks() == (3, 4)
    assert F.blockmap[0][0] == (ur(1, 2), ur(1, 4)) and F.blockmap[1][0] == (ur(3, 3), ur(1, 4))
    assert F.localblockindices(W[0]) == (ur(1, 2), ur(1, 4)) and F.localblockindices(W[1]) == (ur(3, 3), ur(1, 4))
    assert F.localblockindices(W[0], 1) == ur(1, 2) and F.localblockindices(W[0], 2) == ur(1, 4)
    assert F.rng.indices == [ur(1, 20), ur(21, 30)]                   # cuts == [1,21,31]
    assert len(F.dom) == 40 and len(F.rng) == 30
    check_against_serial(F, blocks, rng)
    # replicated-domain path (m a plain array on the master: src:510-580)
    m = rng.random(40)
    y = F.mul_replicated(F.rng.zeros(), m)
    np.testing.assert_allclose(y.to_array(), O.serial_blockop_mul(blocks, m), rtol=1.5e-8)
    dd = F.rng.rand(rng)
    mm = F.mul_adjoint_replicated(np.zeros(40), dd)
    np.testing.assert_allclose(mm, O.serial_blockop_mul_adjoint(blocks, dd.to_array()), rtol=1.5e-8)
    assert F.getblock(3, 2) is blocks[2][1] and F.getblock(1, 4) is blocks[0][3]


def test_operator_block_diagonal_dense_10x5():
    """test:496-512: 4x4, dense 10x5 JopBaz on the diagonal, zero blocks elsewhere, 4 workers, isdiag."""
    rng = np.random.default_rng(3)
    blocks = [[O.ODense(np.asfortranarray(rng.random((10, 5)))) if i == j else O.OZero(np.float64, (10,), (5,))
               for j in range(4)] for i in range(4)]
    A = O.OJopDBlock(blocks, W, [4, 1], isdiag=True)
    assert A.rng.indices == [ur(1, 10), ur(11, 20), ur(21, 30), ur(31, 40)]
    assert A.dom.indices == [ur(1, 5), ur(6, 10), ur(11, 15), ur(16, 20)]       # test:505-506
    check_against_serial(A, blocks, rng)
    with pytest.raises(ValueError):
        O.OJopDBlock(blocks, W, [2, 2], isdiag=True)                              # src:481


def test_operator_block_diagonal_10_blocks_on_4_workers():
    """test:514-527: 3,3,2,2 split."""
    rng = np.random.default_rng(4)
    blocks = [[O.ODiag(rng.random(10)) if i == j else O.OZero(np.float64, (10,), (10,)) for j in range(10)]
              for i in range(10)]
    A = O.OJopDBlock(blocks, W, [4, 1], isdiag=True)
    assert A.rng.blkindices == [ur(1, 3), ur(4, 6), ur(7, 8), ur(9, 10)]
    check_against_serial(A, blocks, rng)


def test_operator_full_2d_grid():
    """test:529-591: 3x5 on grid [2,2]."""
    rng = np.random.default_rng(6)
    blocks = hetero_blocks(rng, 3, 5)
    F = O.OJopDBlock(blocks, W, [2, 2])
    bm = F.blockmap
    assert bm[0][0] == (ur(1, 2), ur(1, 3)) and bm[1][0] == (ur(3, 3), ur(1, 3))            # test:536-539
    assert bm[0][1] == (ur(1, 2), ur(4, 5)) and bm[1][1] == (ur(3, 3), ur(4, 5))
    assert F.procs() == [[W[0], W[2]], [W[1], W[3]]]                                          # test:543
    assert F.rng.indices == [ur(1, 20), ur(21, 30)] and F.dom.indices == [ur(1, 30), ur(31, 50)]
    assert F.localblockindices(W[3]) == (ur(3, 3), ur(4, 5))
    check_against_serial(F, blocks, rng)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_meets_baseline_tolerance_at_scale(dtype):
    """Reference-ordered arithmetic vs ground truth at a mid size (BASELINE tolerances)."""
    rng = np.random.default_rng(8)
    blocks = O.make_blocks(rng, 4, 4, 300, 300, dtype)
    F = O.OJopDBlock(blocks, W, [2, 2])
    check_against_serial(F, blocks, rng, rtol=1.5e-8 if dtype == np.float64 else 1e-4)
