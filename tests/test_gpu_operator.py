"""GPU parity of mul! / adjoint against the oracle.

Every case feeds identical seeded blocks and vectors to the oracle (reference-ordered arithmetic,
src:582-721) and to the CUDA backend through the C ABI, and measures both against a higher-precision
ground truth with the tolerance BASELINE.json states (1e-12 Float64, 1e-5 Float32, norm-wise per block).
"""
import numpy as np
import pytest

from helpers import build_pair, load_like, tol_for

pytestmark = pytest.mark.gpu


def _check(dj, O, oop, gop, oblocks, rng, dtype):
    tol = tol_for(dtype)
    # forward
    xo = oop.dom.rand(rng)
    xg = load_like(gop.domain().zeros(), xo)
    yo = oop.mul(oop.rng.zeros(), xo)
    yg = gop.mul(gop.range().zeros(), xg)
    truth = O.ground_truth_mul(oblocks, xo.to_array()) if not oop.isdiag else None
    if oop.isdiag:
        dblocks = [[oblocks[i][j] if i == j else O.OZero(dtype, oblocks[i][0].rng_shape, oblocks[j][j].dom_shape)
                    for j in range(len(oblocks))] for i in range(len(oblocks))]
        truth = O.ground_truth_mul(dblocks, xo.to_array())
    lens = [int(np.prod(s)) for s in gop.row_shapes]
    err_g = O.blockwise_relerr(yg.to_array(), truth, lens)
    err_o = O.blockwise_relerr(yo.to_array(), truth, lens)
    assert err_o <= tol, f"oracle itself off: {err_o}"
    assert err_g <= tol, f"forward: gpu {err_g} (oracle {err_o}) > {tol}"
    # the reference's own check: distributed == serial block operator under isapprox (test:418-428)
    np.testing.assert_allclose(yg.to_array(), yo.to_array(), rtol=1.5e-8 if dtype == np.float64 else 1e-4)
    # adjoint
    do = oop.rng.rand(rng)
    dg = load_like(gop.range().zeros(), do)
    mo = oop.mul_adjoint(oop.dom.zeros(), do)
    mg = gop.mul_adjoint(gop.domain().zeros(), dg)
    if oop.isdiag:
        truth_t = O.ground_truth_mul(dblocks, do.to_array(), adjoint=True)
    else:
        truth_t = O.ground_truth_mul(oblocks, do.to_array(), adjoint=True)
    clens = [int(np.prod(s)) for s in gop.col_shapes]
    err_g = O.blockwise_relerr(mg.to_array(), truth_t, clens)
    err_o = O.blockwise_relerr(mo.to_array(), truth_t, clens)
    assert err_o <= tol
    assert err_g <= tol, f"adjoint: gpu {err_g} (oracle {err_o}) > {tol}"
    # results land in the caller's vectors (mul! returns its first argument)
    assert yg.nblocks() == oop.rng.nblocks() and mg.nblocks() == oop.dom.nblocks()


CASES = [
    # name, row sizes, col sizes, kind, pids, dist, isdiag
    ("3x4 mixed on [2,1] (test:377-436 shape)", [10] * 3, [10] * 4,
     lambda i, j: "diag" if (i + j) % 3 == 0 else ("zero" if (i, j) == (2, 2) else "dense"), [0, 1], [2, 1], False),
    ("3x5 mixed on [2,2] (test:529-591 shape)", [10] * 3, [10] * 5,
     lambda i, j: "diag" if (i % 2 == 0 and j % 2 == 1) else "dense", [0, 1, 2, 3], [2, 2], False),
    ("4x4 block-diagonal dense 10x5 on 4 workers (test:496-512)", [10] * 4, [5] * 4,
     lambda i, j: "dense" if i == j else "zero", [0, 1, 2, 3], [4, 1], True),
    ("10x10 block-diagonal on 4 workers, 3,3,2,2 split (test:514-527)", [10] * 10, [10] * 10,
     lambda i, j: "diag" if (i == j and i % 2) else ("dense" if i == j else "zero"), [0, 1, 2, 3], [4, 1], True),
    ("misaligned sizes 5/7/3 x 9/1/6/2 on [2,2]", [5, 7, 3], [9, 1, 6, 2],
     lambda i, j: "dense", [0, 1, 2, 3], [2, 2], False),
    ("all-zero block row and column on [1,2] (extension F6)", [6, 4, 8], [4, 6, 5, 8],
     lambda i, j: "zero" if (i == 2 or j == 3) else ("diag" if (i, j) in ((1, 2), (3, 4)) else "dense"),
     [0, 1], [1, 2], False),
    ("single worker 2x3 (extension F6)", [33, 65], [17, 129, 64], lambda i, j: "dense", [0], [1, 1], False),
    ("tall blocks spanning several row strips / column panels on [2,2]", [300, 1100], [1300, 70, 2500],
     lambda i, j: "dense", [0, 1, 2, 3], [2, 2], False),
    ("4x1 grid column vector of blocks", [40, 50, 60, 70, 80], [30], lambda i, j: "dense", [0, 1, 2, 3], [4, 1], False),
    ("1x4 grid", [50], [40, 50, 60, 70, 80], lambda i, j: "dense", [0, 1, 2, 3], [1, 4], False),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_mul_and_adjoint_match_oracle(dj, O, ctx4, case, dtype):
    name, rs, cs, kind, pids, dist, isdiag = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, dtype, pids, dist, isdiag, ctx=ctx4)
    _check(dj, O, oop, gop, oblocks, rng, dtype)


def test_config1_shape_4x4_dense_f64_1000(dj, O, ctx4):
    """BASELINE config 1: 4x4 dense Float64 1000x1000 blocks on a [2,2] grid of 4 workers."""
    rng = np.random.default_rng(20241)
    oop, gop, oblocks = build_pair(dj, O, rng, [1000] * 4, [1000] * 4, lambda i, j: "dense", np.float64,
                                   [0, 1, 2, 3], [2, 2], ctx=ctx4)
    _check(dj, O, oop, gop, oblocks, rng, np.float64)


def test_irregular_chunks(dj, O, ctx4):
    """Caller-prescribed block ranges per grid row / column (src:8-38, test:39-52)."""
    rng = np.random.default_rng(7)
    rr, cr = [range(1, 2), range(2, 5)], [range(1, 4), range(4, 5)]
    oop, gop, oblocks = build_pair(dj, O, rng, [12, 8, 9, 20], [7, 16, 5, 11], lambda i, j: "dense", np.float64,
                                   [0, 1, 2, 3], None, ctx=ctx4, row_ranges=rr, col_ranges=cr)
    assert gop.blockmap()[1][0] == (range(2, 5), range(1, 4))
    _check(dj, O, oop, gop, oblocks, rng, np.float64)


def test_operator_metadata_and_getblock(dj, O, ctx4):
    """blockmap / procs / localblockindices / getblock(A,i,j) (src:805-852; test:381-385, 533-551)."""
    rng = np.random.default_rng(3)
    kind = lambda i, j: "diag" if (i % 2 == 0 and j % 2 == 1) else ("zero" if (i, j) == (1, 5) else "dense")
    oop, gop, oblocks = build_pair(dj, O, rng, [10] * 3, [10] * 5, kind, np.float64, [0, 1, 2, 3], [2, 2], ctx=ctx4)
    assert gop.nblocks() == (3, 5) == oop.nblocks() and gop.nblocks(1) == 3 and gop.nblocks(2) == 5 and gop.isblockop()
    assert gop.procs() == oop.procs() == [[0, 2], [1, 3]]                  # column-major grid, test:543
    assert gop.blockmap() == oop.blockmap
    assert gop.blockmap()[0][0] == (range(1, 3), range(1, 4))               # test:536-539
    assert gop.blockmap()[1][1] == (range(3, 4), range(4, 6))
    for pid in range(4):
        assert gop.localblockindices(pid) == oop.localblockindices(pid)
        assert gop.localblockindices(pid, 1) == oop.localblockindices(pid, 1)
    assert gop.range().indices == oop.rng.indices and gop.domain().indices == oop.dom.indices
    assert gop.domain().indices == [range(1, 31), range(31, 51)]            # test:568-571
    for i in range(1, 4):
        for j in range(1, 6):
            g, o = gop.getblock(i, j), oop.getblock(i, j)
            assert {"dense": dj.BLOCK_DENSE, "diag": dj.BLOCK_DIAG, "zero": dj.BLOCK_ZERO}[o.kind] == g.kind
            if o.kind == "dense":
                assert np.array_equal(g.A, o.A)                             # bit-exact round trip
            elif o.kind == "diag":
                assert np.array_equal(g.diag, o.d)


def test_mismatched_vectors_raise(dj, O, ctx4):
    rng = np.random.default_rng(5)
    _, gop, _ = build_pair(dj, O, rng, [8, 8], [6, 6, 6], lambda i, j: "dense", np.float64, [0, 1, 2, 3], [2, 2],
                           ctx=ctx4)
    with pytest.raises(dj.DimensionMismatch):
        gop.mul(gop.domain().zeros(), gop.domain().zeros())
    with pytest.raises(dj.DimensionMismatch):
        gop.mul_adjoint(gop.domain().zeros(), gop.domain().zeros())
    x32 = gop.domain().zeros().similar(np.float32)
    with pytest.raises(dj.DimensionMismatch):
        gop.mul(gop.range().zeros(), x32)
    with pytest.raises(dj.DJetsError):
        dj.JopDBlock([[dj.JopZeroBlock(np.float64, 3, 3)] * 2] * 2, pids=[0, 1, 2, 3], dist=[2, 2], isdiag=True,
                     ctx=ctx4)                                              # src:481


def test_deterministic_run_to_run(dj, O, ctx4):
    rng = np.random.default_rng(11)
    _, gop, _ = build_pair(dj, O, rng, [700, 900], [1500, 600], lambda i, j: "dense", np.float32, [0, 1, 2, 3],
                           [2, 2], ctx=ctx4)
    x = gop.domain().rand(np.random.default_rng(1))
    a = (gop @ x).to_array()
    for _ in range(3):
        assert np.array_equal(a, (gop @ x).to_array())
    d = gop.range().rand(np.random.default_rng(2))
    b = (gop.T @ d).to_array()
    for _ in range(3):
        assert np.array_equal(b, (gop.T @ d).to_array())


def test_replicated_domain_tall_and_skinny(dj, O, ctx4):
    """The reference's most used shape (test:377-470, docs:200-208): grid [n,1], the model a plain array
    on the master, broadcast forward (src:544-557), tree-summed adjoint (src:565-580)."""
    rng = np.random.default_rng(21)
    kind = lambda i, j: "diag" if (i, j) in ((1, 1), (3, 2)) else ("zero" if (i, j) == (2, 3) else "dense")
    oop, gop, oblocks = build_pair(dj, O, rng, [10, 10, 10], [10] * 4, kind, np.float64, [0, 1], [2, 1], ctx=ctx4)
    assert gop.blockmap()[0][0] == (range(1, 3), range(1, 5)) and gop.blockmap()[1][0] == (range(3, 4), range(1, 5))
    assert gop.range().indices == [range(1, 21), range(21, 31)]           # cuts [1,21,31], test:411-414
    m = rng.random(40)
    d = gop @ m
    want = oop.mul_replicated(oop.rng.zeros(), m).to_array()
    truth = O.ground_truth_mul(oblocks, m)
    assert O.blockwise_relerr(d.to_array(), truth, [10] * 3) <= 1e-12
    np.testing.assert_allclose(d.to_array(), want, rtol=1.5e-8)
    dd = oop.rng.rand(rng)
    dg = load_like(gop.range().zeros(), dd)
    got = gop.mul_adjoint(np.zeros(40), dg)
    want = oop.mul_adjoint_replicated(np.zeros(40), dd)
    assert O.blockwise_relerr(got, O.ground_truth_mul(oblocks, dd.to_array(), adjoint=True), [10] * 4) <= 1e-12
    np.testing.assert_allclose(got, want, rtol=1.5e-8)
    _, g22, _ = build_pair(dj, O, rng, [4, 4], [4, 4], lambda i, j: "dense", np.float64, [0, 1, 2, 3], [2, 2], ctx=ctx4)
    with pytest.raises(dj.DimensionMismatch):
        g22 @ np.zeros(8)


@pytest.mark.parametrize("dtype,m,n", [(np.float64, 9001, 70), (np.float32, 17003, 45), (np.float64, 6, 5000)])
def test_blocks_taller_than_one_stage(dj, O, ctx4, dtype, m, n):
    """Blocks taller than the 32 KB y stage of an adjoint tile (several row panels -> several planes per
    output strip) and wider than one x stage of a forward tile."""
    rng = np.random.default_rng(m)
    oop, gop, oblocks = build_pair(dj, O, rng, [m, 33], [n, 17], lambda i, j: "dense", dtype, [0, 1], [1, 2], ctx=ctx4)
    _check(dj, O, oop, gop, oblocks, rng, dtype)


def test_empty_blocks_and_scalar_blocks(dj, O, ctx4):
    """Zero-length and one-element blocks (ragged inputs)."""
    rng = np.random.default_rng(5)
    oop, gop, oblocks = build_pair(dj, O, rng, [1, 7, 1], [1, 3], lambda i, j: "dense", np.float64, [0, 1, 2, 3], [2, 2],
                                   ctx=ctx4)
    _check(dj, O, oop, gop, oblocks, rng, np.float64)
    x = dj.DBArray(lambda i: np.zeros(0) if i == 2 else np.full(3, float(i)), (3,), [0, 1, 2])
    assert len(x) == 6 and x.indices == [range(1, 4), range(4, 4), range(4, 7)]
    assert dj.getblock(x, 2).shape == (0,) and x.sum() == 12.0 and np.array_equal(x.to_array(), [1, 1, 1, 3, 3, 3])


def test_adjoint_dot_test_full_size_config2_shard(dj, ctx1):
    """Size-independent property at the benchmark's block shape: <A x, y> == <x, A' y> on 8 dense Float32
    4096x4096 diagonal blocks (one GPU's shard of BASELINE config 2 at N=8), plus a rank-one known answer."""
    rng = np.random.default_rng(8)
    u = [rng.random(4096, dtype=np.float32) for _ in range(8)]
    v = [rng.random(4096, dtype=np.float32) for _ in range(8)]
    A = dj.JopDBlock(lambda i, j: dj.JopDense(np.asfortranarray(np.outer(u[i - 1], v[i - 1])))
                     if i == j else dj.JopZeroBlock(np.float32, 4096, 4096), dims=(8, 8), pids=[0], dist=[1, 1],
                     isdiag=True, shapes=([(4096,)] * 8, [(4096,)] * 8), dtype=np.float32, ctx=ctx1)
    x = dj.DBArray(lambda i: rng.random(4096, dtype=np.float32), (8,), [0], ctx=ctx1)
    y = dj.DBArray(lambda i: rng.random(4096, dtype=np.float32), (8,), [0], ctx=ctx1)
    Ax, Aty = A @ x, A.T @ y
    lhs, rhs = float(Ax.dot(y)), float(x.dot(Aty))
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    for ib in range(1, 9):                                 # (u v') x = u (v . x)
        want = np.outer(u[ib - 1], v[ib - 1]).astype(np.float64) @ x.getblock(ib).astype(np.float64)
        got = Ax.getblock(ib).astype(np.float64)
        assert np.linalg.norm(got - want) <= 1e-5 * np.linalg.norm(want)


def test_graph_replay_matches_direct_launches(dj, O, ctx4):
    """Whole applies replayed as CUDA graphs (second sighting of a (vectors, direction) pair) give the
    same bits as direct launches, read fresh input data on every replay, and count their launches."""
    rng = np.random.default_rng(31)
    kind = lambda i, j: "diag" if (i, j) == (2, 2) else "dense"
    oop, gop, oblocks = build_pair(dj, O, rng, [300, 300, 120], [300, 300, 80, 64], kind, np.float64, [0, 1, 2, 3],
                                   [2, 2], ctx=ctx4)
    x, y = gop.domain().zeros(), gop.range().zeros()
    d, m = gop.range().zeros(), gop.domain().zeros()
    results = {}
    for graphs, p2p in ((0, 0), (0, 1), (1, 1), (1, 0)):
        ctx4.set_param("graphs", graphs)
        ctx4.set_param("p2p", p2p)                           # 0: staged copies; 1: in-place peer loads / stores
        gop.replan()
        ctx4.reset_stats()
        out = []
        for it in range(5):                                  # same vectors every time -> captured at it == 1
            x.scatter(np.random.default_rng(100 + it).random(len(x)))
            d.scatter(np.random.default_rng(200 + it).random(len(d)))
            gop.mul(y, x)
            gop.mul_adjoint(m, d)
            out.append((y.to_array().copy(), m.to_array().copy()))
        results[(graphs, p2p)] = out
        st = ctx4.kernel_stats()
        assert st["gemv_n"]["launches"] == 5 * 4 and st["gemv_t"]["launches"] == 5 * 4, st
    ctx4.set_param("graphs", 1)
    ctx4.set_param("p2p", 1)
    for key in ((0, 1), (1, 1), (1, 0)):
        for a, b in zip(results[(0, 0)], results[key]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), key
    xo = oop.dom.zeros()
    o = 0
    flat = np.random.default_rng(104).random(len(x))
    for ib, n in enumerate([300, 300, 80, 64], start=1):
        xo.setblock(ib, flat[o:o + n]); o += n
    truth = O.ground_truth_mul(oblocks, flat)
    assert O.blockwise_relerr(results[(1, 1)][-1][0], truth, [300, 300, 120]) <= 1e-12


def test_random_operators_match_oracle(dj, O, ctx4):
    """Forty seeded random operators: random block counts and sizes (including misaligned and
    one-element blocks), random grids over 1-4 ranks, random mix of dense / diagonal / zero blocks,
    random element type; forward and adjoint against ground truth and the oracle."""
    master = np.random.default_rng(424242)
    grids = [(1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (1, 4), (3, 1), (1, 3)]
    for case in range(40):
        n1, n2 = grids[master.integers(len(grids))]
        N, M = int(master.integers(n1, n1 + 4)), int(master.integers(n2, n2 + 4))
        sizes = [1, 2, 3, 5, 8, 17, 33, 64, 100, 129, 257, 600]
        rs = [int(sizes[master.integers(len(sizes))]) for _ in range(N)]
        cs = [int(sizes[master.integers(len(sizes))]) for _ in range(M)]
        dtype = np.float64 if master.integers(2) else np.float32
        pick = master.random((N, M))

        def kind(i, j, rs=rs, cs=cs, pick=pick):
            p = pick[i - 1, j - 1]
            if p < 0.2:
                return "zero"
            if p < 0.45 and rs[i - 1] == cs[j - 1]:
                return "diag"
            return "dense"

        rng = np.random.default_rng(1000 + case)
        oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, dtype, list(range(n1 * n2)), [n1, n2], ctx=ctx4)
        _check(dj, O, oop, gop, oblocks, rng, dtype)


@pytest.mark.parametrize("heights", ["two warp-loads", "one warp-load"])
@pytest.mark.parametrize("chain", [3, 16])
def test_chained_tiles_match_oracle(dj, O, ctx4, chain, heights):
    """Chains of small blocks (one CTA multiplies several blocks of a block column / block row and writes one
    partial): forced on for small operators, ragged block sizes up to the two-warp-load limit (and up to one
    warp-load: the adjoint then loads two blocks per batch), zero / diagonal blocks breaking the chains, forward
    chains cut where the staged columns would overflow, every grid shape; against ground truth and the oracle."""
    master = np.random.default_rng(77 + chain + len(heights))
    grids = [(1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (1, 4)]
    ctx4.set_param("chain", chain)
    try:
        fewer = 0
        for case in range(12):
            n1, n2 = grids[master.integers(len(grids))]
            N, M = int(master.integers(max(n1, 5), 12)), int(master.integers(max(n2, 4), 11))
            dtype = np.float64 if master.integers(2) else np.float32
            if heights == "one warp-load":
                rsizes = [1, 2, 3, 5, 8, 17, 33, 63, 64] + ([100, 127, 128] if dtype == np.float32 else [])
            else:
                rsizes = [1, 3, 8, 17, 64, 65, 100, 127, 128] + ([129, 255, 256] if dtype == np.float32 else [])
            csizes = [1, 2, 5, 17, 64, 100, 129, 256, 400]
            rs = [int(rsizes[master.integers(len(rsizes))]) for _ in range(N)]
            cs = [int(csizes[master.integers(len(csizes))]) for _ in range(M)]
            pick = master.random((N, M))

            def kind(i, j, rs=rs, cs=cs, pick=pick):
                p = pick[i - 1, j - 1]
                if p < 0.1:
                    return "zero"
                if p < 0.25 and rs[i - 1] == cs[j - 1]:
                    return "diag"
                return "dense"

            rng = np.random.default_rng(5000 + case)
            oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, dtype, list(range(n1 * n2)), [n1, n2], ctx=ctx4)
            _check(dj, O, oop, gop, oblocks, rng, dtype)
            ndense = sum(kind(i, j) == "dense" for i in range(1, N + 1) for j in range(1, M + 1))
            fewer += gop.schedule_info(adjoint=True)[0] < ndense and gop.schedule_info(adjoint=False)[0] < ndense
        assert fewer >= 10, "the chained schedules were not used"
    finally:
        ctx4.set_param("chain", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("nb,blen", [(1, 20000), (3, 9001), (5, 12345), (6, 4096)])
def test_long_diagonal_blocks_all_term_counts(dj, ctx1, dtype, nb, blen):
    """Grids of long diagonal blocks: every batching shape of the diagonal-term sum (one term; two to three terms;
    four and more; the vector tails and the scalar tail of a block length that is no multiple of anything), both
    directions, against NumPy in Float64 with the products rounded in the working type like Julia's broadcast."""
    rng = np.random.default_rng(31 * nb + blen)
    diags = [[rng.standard_normal(blen).astype(dtype) for _ in range(nb)] for _ in range(nb)]
    A = dj.JopDBlock([[dj.JopDiagonal(diags[i][j]) for j in range(nb)] for i in range(nb)], pids=[0], dist=[1, 1], ctx=ctx1)
    x = rng.standard_normal(nb * blen).astype(dtype)
    xg = A.domain().zeros()
    for j in range(nb):
        xg.setblock(j + 1, x[j * blen:(j + 1) * blen])
    y = A.mul(A.range().zeros(), xg).to_array()
    xt = A.mul_adjoint(A.domain().zeros(), xg).to_array()
    for i in range(nb):
        want = sum((diags[i][j] * x[j * blen:(j + 1) * blen]).astype(np.float64) for j in range(nb))
        want_t = sum((diags[j][i] * x[j * blen:(j + 1) * blen]).astype(np.float64) for j in range(nb))
        np.testing.assert_array_equal(y[i * blen:(i + 1) * blen], want.astype(dtype))
        np.testing.assert_array_equal(xt[i * blen:(i + 1) * blen], want_t.astype(dtype))


def test_perfstatfile_schema(dj, O, ctx4, tmp_path):
    """perfstatfile (src:723-745; schema pinned by test:618-651): one step per apply, one entry per worker
    with id / hostname / operation / localblock (8 and 4 local blocks for a 3x4 operator on grid [2,1])."""
    import json
    path = str(tmp_path / "stats.json")
    open(path, "w").write("stale")                       # removed at construction (src:505)
    rng = np.random.default_rng(1)
    F = dj.JopDBlock(lambda i, j: dj.JopDense(np.asfortranarray(rng.random((10, 10)))), dims=(3, 4), pids=[0, 1],
                     dist=[2, 1], perfstatfile=path, ctx=ctx4)
    assert not (tmp_path / "stats.json").exists()
    m = F.domain().rand(rng)
    d = F @ m
    s = json.load(open(path))
    assert len(s["step"]) == 1 and len(s["step"][0]["pid"]) == 2
    assert len(s["step"][0]["pid"][0]["localblock"]) == 8 and len(s["step"][0]["pid"][1]["localblock"]) == 4
    assert s["step"][0]["pid"][0]["operation"] == "df" and s["step"][0]["pid"][0]["id"] == 0
    assert all(v > 0 for p in s["step"][0]["pid"] for v in p["localblock"])
    F.T @ d
    F @ m
    s = json.load(open(path))
    assert [st["pid"][1]["operation"] for st in s["step"]] == ["df", "df′", "df"]
    assert isinstance(s["step"][1]["pid"][0]["hostname"], str)


def test_cgls_solver_loop_converges_to_lstsq(dj, O, ctx4):
    """The solver loop of BASELINE config 5 (mul!, adjoint, dot, norm, axpy per iteration) on a mixed
    dense/diagonal operator, against numpy's least-squares solution of the assembled matrix."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "cgls_example", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "cgls.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(12)
    nbr, nbc, bs = 5, 4, 24                                # over-determined: 120 x 96
    rs, cs = [bs] * nbr, [bs] * nbc
    kind = lambda i, j: "diag" if (i % 2 == 0 and j % 2 == 1) else "dense"
    oop, gop, oblocks = build_pair(dj, O, rng, rs, cs, kind, np.float64, [0, 1, 2, 3], [2, 2], ctx=ctx4)
    dense = np.zeros((nbr * bs, nbc * bs))
    for i in range(nbr):
        for j in range(nbc):
            b = oblocks[i][j]
            dense[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = b.A if b.kind == "dense" else np.diag(b.d)
    rhs = rng.random(nbr * bs)
    b = gop.range().zeros().scatter(rhs)
    x, hist = mod.cgls(gop, b, iters=400, tol=1e-13)
    want = np.linalg.lstsq(dense, rhs, rcond=None)[0]
    assert np.linalg.norm(x.to_array() - want) <= 1e-6 * np.linalg.norm(want)
    assert hist[-1] < hist[0]


def test_nan_follows_the_reference_through_zero_blocks(dj, O, ctx4):
    """A NaN in one domain block reaches exactly the block rows that have a non-zero block in that
    column: zero blocks are skipped (iszero, src:587,634), not multiplied."""
    rng = np.random.default_rng(2)
    kind = lambda i, j: "zero" if (i, j) in ((1, 2), (3, 2)) else ("diag" if (i, j) == (2, 2) else "dense")
    oop, gop, oblocks = build_pair(dj, O, rng, [7, 5, 9], [6, 5, 4], kind, np.float64, [0, 1, 2, 3], [2, 2], ctx=ctx4)
    x = gop.domain().rand(rng)
    blk = x.getblock(2)
    blk[3] = np.nan
    x.setblock(2, blk)
    y = gop @ x
    assert np.isfinite(y.getblock(1)).all() and np.isfinite(y.getblock(3)).all()
    yb = y.getblock(2)
    assert np.isnan(yb[3]) and np.isfinite(np.delete(yb, 3)).all()        # the diagonal block touches one row only
    assert np.array_equal(x[range(6 + 4, 6 + 5)], [np.nan], equal_nan=True)  # x[10:10]


def test_dense_block_with_more_than_2_pow_31_elements(dj, ctx1):
    """Maximum sizes: one dense Float32 block of 46400 x 46400 (2.15e9 elements, 8.6 GB) -- every tile
    offset must be 64-bit.  A = ones except one marked row and one marked column, so the answers are known."""
    import psutil
    if psutil.virtual_memory().available < 24 << 30:
        pytest.skip("needs 9 GB of host memory with headroom")
    n = 46400
    A = np.ones((n, n), dtype=np.float32, order="F")
    A[n - 1, :] = 2.0                                       # last row
    A[:, n - 1] = 3.0                                       # last column (corner = 3)
    op = dj.JopDBlock([[dj.JopDense(A)]], pids=[0], dist=[1, 1], ctx=ctx1)
    del A
    x = op.domain().zeros()
    xs = np.zeros(n, dtype=np.float32)
    xs[0], xs[n - 1] = 1.0, 0.5
    x.scatter(xs)
    y = (op @ x).to_array()
    assert y[0] == 1.0 + 1.5 and y[n - 2] == 2.5 and y[n - 1] == 2.0 + 1.5
    assert np.all(y[:n - 1] == 2.5)
    d = op.range().zeros()
    ds = np.zeros(n, dtype=np.float32)
    ds[n - 1], ds[5] = 1.0, 0.25
    d.scatter(ds)
    m = (op.T @ d).to_array()
    assert m[0] == 2.0 + 0.25 and m[n - 1] == 3.0 + 0.75 and np.all(m[:n - 1] == 2.25)
