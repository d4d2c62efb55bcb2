"""Complex{Float32} / Complex{Float64} element types (the reference's tests build DArrays of all four element types,
test/runtests.jl:38): block access bit-exact, dense blocks apply `A*m` forward and the CONJUGATE transpose `A'*d` in
the adjoint (test/runtests.jl:30-31), diagonal blocks `diag .* m` / `conj(diag) .* d`, `dot` conjugates its first
argument, norms use |z|.  Compared with NumPy in complex128 / long double at the BASELINE tolerances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def crand(rng, shape, dtype):
    rt = np.zeros(0, dtype).real.dtype
    return (rng.random(shape, dtype=rt) - 0.5 + 1j * (rng.random(shape, dtype=rt) - 0.5)).astype(dtype)


def relerr_blocks(got, want, lens):
    worst, o = 0.0, 0
    for n in lens:
        den = np.linalg.norm(want[o:o + n])
        worst = max(worst, float(np.linalg.norm(got[o:o + n] - want[o:o + n]) / max(den, 1e-300)))
        o += n
    return worst


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
@pytest.mark.parametrize("grid", [[2, 2], [1, 4], [4, 1], [1, 1]])
def test_complex_operator_forward_and_conjugate_adjoint(dj, ctx4, dtype, tol, grid):
    rng = np.random.default_rng(5)
    rs, cs = [300, 257, 64, 129, 33], [200, 131, 300, 77]
    kinds = {(2, 2): "diag_like"}                             # block (2,2) replaced below by a zero block; (3,3) diagonal
    blocks, dense = [], {}
    for i in range(1, len(rs) + 1):
        row = []
        for j in range(1, len(cs) + 1):
            if (i, j) == (2, 2):
                row.append(dj.JopZeroBlock(dtype, rs[i - 1], cs[j - 1]))
            elif (i, j) == (3, 4) and rs[i - 1] == cs[j - 1]:
                pass
            else:
                A = np.asfortranarray(crand(rng, (rs[i - 1], cs[j - 1]), dtype))
                dense[(i, j)] = A
                row.append(dj.JopDense(A))
        blocks.append(row)
    n1, n2 = grid
    A = dj.JopDBlock(blocks, pids=list(range(n1 * n2)), dist=grid, ctx=ctx4)
    hx, hy = crand(rng, sum(cs), dtype), crand(rng, sum(rs), dtype)
    x, y = A.domain().zeros().scatter(hx), A.range().zeros().scatter(hy)
    # truth in complex128 / long double
    wide = np.clongdouble if dtype == np.complex128 else np.complex128
    ro, co = np.cumsum([0] + rs), np.cumsum([0] + cs)
    want_f = np.zeros(sum(rs), wide)
    want_a = np.zeros(sum(cs), wide)
    for (i, j), B in dense.items():
        Bw = B.astype(wide)
        want_f[ro[i - 1]:ro[i]] += Bw @ hx[co[j - 1]:co[j]].astype(wide)
        want_a[co[j - 1]:co[j]] += Bw.conj().T @ hy[ro[i - 1]:ro[i]].astype(wide)
    got_f = (A @ x).to_array()
    got_a = (A.T @ y).to_array()
    assert got_f.dtype == np.dtype(dtype)
    assert relerr_blocks(got_f.astype(wide), want_f, rs) <= tol
    assert relerr_blocks(got_a.astype(wide), want_a, cs) <= tol
    # adjoint identity <A x, y> == <x, A' y> through the library's own dot
    lhs, rhs = complex((A @ x).dot(y)), complex(x.dot(A.T @ y))
    assert abs(lhs - rhs) <= 50 * tol * abs(lhs)
    A.close()


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
def test_complex_diagonal_blocks_and_mixed_rows(dj, ctx4, dtype, tol):
    rng = np.random.default_rng(6)
    n, nb = 517, 3
    diags = {(i, j): crand(rng, n, dtype) for i in range(1, nb + 1) for j in range(1, nb + 1) if (i + j) % 2 == 0}
    dens = {(i, j): np.asfortranarray(crand(rng, (n, n), dtype)) for i in range(1, nb + 1) for j in range(1, nb + 1)
            if (i + j) % 2 == 1}
    blocks = [[dj.JopDiagonal(diags[(i, j)]) if (i, j) in diags else dj.JopDense(dens[(i, j)]) for j in range(1, nb + 1)]
              for i in range(1, nb + 1)]
    A = dj.JopDBlock(blocks, pids=[0, 1, 2, 3], dist=[2, 2], ctx=ctx4)
    hx, hy = crand(rng, n * nb, dtype), crand(rng, n * nb, dtype)
    x, y = A.domain().zeros().scatter(hx), A.range().zeros().scatter(hy)
    wide = np.clongdouble if dtype == np.complex128 else np.complex128
    want_f, want_a = np.zeros(n * nb, wide), np.zeros(n * nb, wide)
    for i in range(1, nb + 1):
        for j in range(1, nb + 1):
            xs, ys = hx[(j - 1) * n:j * n].astype(wide), hy[(i - 1) * n:i * n].astype(wide)
            if (i, j) in diags:
                d = diags[(i, j)].astype(wide)
                want_f[(i - 1) * n:i * n] += d * xs
                want_a[(j - 1) * n:j * n] += d.conj() * ys
            else:
                B = dens[(i, j)].astype(wide)
                want_f[(i - 1) * n:i * n] += B @ xs
                want_a[(j - 1) * n:j * n] += B.conj().T @ ys
    assert relerr_blocks((A @ x).to_array().astype(wide), want_f, [n] * nb) <= tol
    assert relerr_blocks((A.T @ y).to_array().astype(wide), want_a, [n] * nb) <= tol
    # block-diagonal operator of diagonal blocks: the streaming path
    D = dj.JopDBlock(lambda i, j: dj.JopDiagonal(diags[(1, 1)]) if i == j else dj.JopZeroBlock(dtype, n, n), dims=(4, 4),
                     pids=[0, 1, 2, 3], dist=[4, 1], isdiag=True, shapes=([(n,)] * 4, [(n,)] * 4), dtype=dtype, ctx=ctx4)
    hz = crand(rng, 4 * n, dtype)
    z = D.domain().zeros().scatter(hz)
    def cmul(a, b):      # Julia's complex product: separately rounded products and sums (NumPy's SIMD path may fuse them)
        return ((a.real * b.real - a.imag * b.imag) + 1j * (a.real * b.imag + a.imag * b.real)).astype(dtype)
    np.testing.assert_array_equal((D @ z).to_array(), cmul(np.tile(diags[(1, 1)], 4), hz))
    np.testing.assert_array_equal((D.T @ z).to_array(), cmul(np.tile(diags[(1, 1)].conj(), 4), hz))
    A.close(); D.close()


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_complex_dbarray_algebra(dj, ctx4, dtype):
    rng = np.random.default_rng(7)
    shapes = [[(5,), (2, 3)], [(1031,)], [(7,)], [(64,), (3,)]]
    R = dj.JetDSpace(shapes, [0, 1, 2, 3], dtype, ctx4)
    hx, hy = crand(rng, len(R), dtype), crand(rng, len(R), dtype)
    x, y = R.zeros().scatter(hx), R.zeros().scatter(hy)
    # block access / collect / getindex: bit-exact
    np.testing.assert_array_equal(x.to_array(), hx)
    assert x.getblock(2).shape == (2, 3) and np.array_equal(x.getblock(2).reshape(-1, order="F"), hx[5:11])
    assert x[7] == hx[6]
    blk = crand(rng, (2, 3), dtype)
    x.setblock(2, blk)
    hx[5:11] = blk.reshape(-1, order="F")
    np.testing.assert_array_equal(x.to_array(), hx)
    # dot conjugates its first argument; norms use |z|
    rt = 1e-12 if dtype == np.complex128 else 2e-6
    want = np.vdot(hx.astype(np.complex128), hy.astype(np.complex128))
    assert abs(complex(x.dot(y)) - want) <= rt * abs(want)
    ax = np.abs(hx.astype(np.complex128))
    for p, w in ((2, np.sqrt(np.sum(ax * ax))), (1, ax.sum()), (np.inf, ax.max()), (-np.inf, ax.min()), (3, np.sum(ax ** 3) ** (1 / 3))):
        assert abs(float(x.norm(p)) - w) <= 10 * rt * w, p
    assert float(x.norm(0)) == np.count_nonzero(hx)
    # fill, copy, real-coefficient linear combinations: bit-exact vs NumPy
    z = x.similar().fill(complex(1.5, -2.0))
    assert np.all(z.to_array() == dtype(complex(1.5, -2.0)))
    np.testing.assert_array_equal(z.assign(x).to_array(), hx)
    rtype = np.zeros(0, dtype).real.dtype.type
    out = x.similar().lincomb_([rtype(0.5), rtype(-2.0)], [x, y], wide=False)
    np.testing.assert_array_equal(out.to_array(), rtype(0.5) * hx + rtype(-2.0) * hy)
    out.assign(rtype(3.0) * dj.lazy(x) - dj.lazy(y))
    np.testing.assert_array_equal(out.to_array(), rtype(3.0) * hx - hy)
    with pytest.raises(dj.DJetsError):
        out.assign(abs(dj.lazy(x)))                          # non-linear broadcasts: real element types only
    with pytest.raises(dj.DJetsError):
        x.maximum()
