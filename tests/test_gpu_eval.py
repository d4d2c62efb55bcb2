"""The closed elementwise program (djets_vec_eval), device-resident scalars, captured sequences and the
stream-ordered host transfers, through the C ABI on a B200.  Broadcast results are compared BIT-EXACTLY with
NumPy evaluating the same unfused expression (reference: src/DistributedJets.jl:363-375 evaluates the
Broadcasted tree per worker; test/runtests.jl:306-363, benchmark/benchmarks.jl:43 are the forms it exercises)."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAYOUTS = [  # blocks per part (4 ranks), block shapes cycle through these
    ([2, 2, 2, 1], [(2,), (2, 3), (5,), (1,)]),
    ([1, 1, 1, 1], [(1031,), (7, 11), (4099,), (3,)]),
    ([3, 0, 2, 1], [(64,), (129,), (2, 2, 2)]),
]


def make(dj, ctx, layout, dtype, rng, lo=-1.0, hi=1.0):
    counts, shapes = layout
    part_shapes, k = [], 0
    for c in counts:
        part_shapes.append([shapes[(k + j) % len(shapes)] for j in range(c)])
        k += c
    R = dj.JetDSpace(part_shapes, [0, 1, 2, 3], dtype, ctx)
    x = R.zeros()
    flat = (lo + (hi - lo) * rng.random(len(R))).astype(dtype)
    x.scatter(flat)
    return x, flat


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_eval_forms_are_bit_exact_vs_numpy(dj, ctx4, dtype, layout):
    rng = np.random.default_rng(7)
    T = np.dtype(dtype).type
    d1, a1 = make(dj, ctx4, layout, dtype, rng)
    d2, a2 = make(dj, ctx4, layout, dtype, rng)
    d3, a3 = make(dj, ctx4, layout, dtype, rng)
    out = d1.similar()
    c1, c2, c3 = T(0.37), T(-1.91), T(2.5)
    L = dj.lazy
    cases = [
        # benchmark/benchmarks.jl:43  d4 .= a1*d1 .+ a2*d2 .- a3*d3
        (c1 * L(d1) + c2 * L(d2) - c3 * L(d3), (c1 * a1 + c2 * a2) - c3 * a3),
        (abs(L(d1)), np.abs(a1)),                                   # abs.(d)
        (L(d1) ** 2, a1 * a1),                                      # d.^2 (literal_pow: x*x)
        (L(d1) ** 3, (a1 * a1) * a1),
        (L(d1) + T(1.0), a1 + T(1.0)),                              # d .+ 1.0
        (dj.bc_max(L(d1), T(0)), np.maximum(a1, T(0))),             # max.(d, 0)
        (dj.bc_min(L(d1), L(d2)), np.minimum(a1, a2)),
        (dj.bc_sqrt(abs(L(d1))), np.sqrt(np.abs(a1))),
        (L(d1) / (abs(L(d2)) + T(0.5)), a1 / (np.abs(a2) + T(0.5))),
        (T(2.0) / (L(d3) * L(d3) + T(1.0)), T(2.0) / (a3 * a3 + T(1.0))),
        (T(1.5) - L(d2), T(1.5) - a2),
        (-(L(d1) * L(d2)) + L(d3), -(a1 * a2) + a3),
        (((L(d1) + L(d2)) * (L(d2) - L(d3))) / ((abs(L(d1)) + T(1.0)) * (abs(L(d3)) + T(2.0))),
         ((a1 + a2) * (a2 - a3)) / ((np.abs(a1) + T(1.0)) * (np.abs(a3) + T(2.0)))),   # four operands deep
    ]
    for expr, want in cases:
        got = out.eval_(expr, wide=False).to_array()
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)
    # in place: dest aliases an input
    d1.eval_(L(d1) * c2 + L(d2), wide=False)
    np.testing.assert_array_equal(d1.to_array(), a1 * c2 + a2)


def test_eval_float64_scalars_promote_like_julia(dj, ctx4):
    """Float32 arrays with Float64 scalars are evaluated in Float64 and rounded once (Julia's promotion)."""
    rng = np.random.default_rng(3)
    x, ax = make(dj, ctx4, LAYOUTS[1], np.float32, rng)
    y, ay = make(dj, ctx4, LAYOUTS[1], np.float32, rng)
    out = x.similar()
    a, b = 0.1234567890123, -3.3333333333
    got = out.assign(a * dj.lazy(x) + b * dj.lazy(y)).to_array()
    want = (a * ax.astype(np.float64) + b * ay.astype(np.float64)).astype(np.float32)
    np.testing.assert_array_equal(got, want)
    # Float32 scalars stay in Float32
    got = out.assign(np.float32(a) * dj.lazy(x) + np.float32(b) * dj.lazy(y)).to_array()
    np.testing.assert_array_equal(got, np.float32(a) * ax + np.float32(b) * ay)


def test_eval_converts_element_types(dj, ctx4):
    rng = np.random.default_rng(4)
    x64, a64 = make(dj, ctx4, LAYOUTS[0], np.float64, rng)
    x32 = x64.similar(np.float32)
    x32.assign(dj.lazy(x64) * 2.0 + 1.0)
    np.testing.assert_array_equal(x32.to_array(), (a64 * 2.0 + 1.0).astype(np.float32))
    back = x64.similar()
    back.assign(abs(dj.lazy(x32)))
    np.testing.assert_array_equal(back.to_array(), np.abs((a64 * 2.0 + 1.0).astype(np.float32)).astype(np.float64))


def test_eval_general_power_and_transcendentals(dj, ctx4):
    rng = np.random.default_rng(5)
    x, ax = make(dj, ctx4, LAYOUTS[1], np.float64, rng, 0.1, 2.0)
    out = x.similar()
    np.testing.assert_allclose(out.assign(dj.lazy(x) ** 1.7).to_array(), ax ** 1.7, rtol=4e-16)
    np.testing.assert_allclose(out.assign(dj.bc_exp(dj.lazy(x))).to_array(), np.exp(ax), rtol=4e-16)
    np.testing.assert_allclose(out.assign(dj.bc_log(dj.lazy(x))).to_array(), np.log(ax), rtol=4e-16, atol=1e-16)


def test_eval_min_max_propagate_nan(dj, ctx4):
    rng = np.random.default_rng(6)
    x, ax = make(dj, ctx4, LAYOUTS[0], np.float64, rng)
    ax[3] = np.nan
    x.scatter(ax)
    out = x.similar()
    got = out.assign(dj.bc_max(dj.lazy(x), 0.0)).to_array()
    assert np.isnan(got[3]) and np.array_equal(np.delete(got, 3), np.maximum(np.delete(ax, 3), 0.0))


def test_non_dotted_scalar_arithmetic_allocates(dj, ctx4):
    rng = np.random.default_rng(8)
    x, ax = make(dj, ctx4, LAYOUTS[0], np.float64, rng)
    np.testing.assert_array_equal((x + 1.0).to_array(), ax + 1.0)
    np.testing.assert_array_equal((2.0 - x).to_array(), 2.0 - ax)
    np.testing.assert_array_equal((x / 4.0).to_array(), ax / 4.0)
    np.testing.assert_array_equal(abs(x).to_array(), np.abs(ax))
    np.testing.assert_array_equal((x ** 2).to_array(), ax * ax)


def test_eval_refuses_what_it_cannot_do(dj, ctx4):
    rng = np.random.default_rng(9)
    x, _ = make(dj, ctx4, LAYOUTS[0], np.float64, rng)
    y, _ = make(dj, ctx4, LAYOUTS[2], np.float64, rng)
    with pytest.raises(dj.DimensionMismatch):
        x.assign(dj.lazy(x) + dj.lazy(y))
    deep = dj.lazy(x)
    for _ in range(60):
        deep = deep * dj.lazy(x) + dj.lazy(x)
    with pytest.raises(dj.DJetsError):
        x.assign(deep)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_scalars_match_the_host_returning_reductions(dj, ctx4, dtype):
    """dot_s / norm_s fold the per-rank partials on the device with the master-side rules of src:200-222:
    the value must be the very one dot / norm return to the host."""
    rng = np.random.default_rng(11)
    for layout in LAYOUTS:
        x, ax = make(dj, ctx4, layout, dtype, rng)
        y, ay = make(dj, ctx4, layout, dtype, rng)
        assert x.dot_s(y).get() == x.dot(y)
        for p in (2, 1, np.inf, -np.inf, 0):
            assert x.norm_s(p).get() == x.norm(p), p
        np.testing.assert_allclose(float(x.norm_s(3).get()), float(x.norm(3)), rtol=1e-6 if dtype == np.float32 else 1e-14)
        from djets_b200 import _lib as L
        assert x.reduce_s(L.RED_SUM).get() == x.sum()
        if 0 not in layout[0]:
            assert x.reduce_s(L.RED_MAX).get() == x.maximum()
            assert x.reduce_s(L.RED_ABSMIN).get() == x.minimum(abs)
        else:                                                    # norm(empty local part) = 0 takes part in the fold (src:202-203)
            assert x.norm(-np.inf) == 0


def test_scalar_arithmetic_and_scalar_coefficients(dj, ctx4):
    rng = np.random.default_rng(12)
    x, ax = make(dj, ctx4, LAYOUTS[1], np.float64, rng)
    y, ay = make(dj, ctx4, LAYOUTS[1], np.float64, rng)
    g, d = x.dot_s(x), y.dot_s(y)
    alpha = g / d
    assert float(alpha) == float(x.dot(x)) / float(y.dot(y))
    assert float(-alpha) == -float(alpha) and float((g * d).sqrt()) == np.sqrt(float(g) * float(d))
    out = x.similar()
    out.assign(dj.lazy(x) - alpha * dj.lazy(y))
    np.testing.assert_array_equal(out.to_array(), ax - float(alpha) * ay)
    s32 = dj.Scalar(1.0, np.float32) / dj.Scalar(3.0, np.float32)
    assert s32.get() == np.float32(1.0) / np.float32(3.0)


def test_reductions_propagate_nan_and_refuse_empty_parts(dj, ctx4):
    rng = np.random.default_rng(13)
    x, ax = make(dj, ctx4, LAYOUTS[0], np.float64, rng)
    ax[5] = np.nan
    x.scatter(ax)
    assert np.isnan(x.maximum()) and np.isnan(x.minimum()) and np.isnan(x.norm(np.inf)) and np.isnan(x.norm(-np.inf))
    assert np.isnan(x.norm_s(np.inf).get())
    e, _ = make(dj, ctx4, LAYOUTS[2], np.float64, rng)        # part 2 of 4 is empty
    with pytest.raises(dj.DJetsError, match="empty collection"):
        e.maximum()
    assert np.isfinite(e.norm(2)) and np.isfinite(e.sum())


def test_async_transfers_and_marks(dj, ctx4):
    rng = np.random.default_rng(14)
    x, _ = make(dj, ctx4, LAYOUTS[1], np.float32, rng)
    n = len(x)
    hin = [dj.PinnedBuffer(n, np.float32) for _ in range(2)]
    hout = [dj.PinnedBuffer(n, np.float32) for _ in range(2)]
    y = x.similar()
    want = []
    for step in range(6):                                        # double-buffered: the host reads step k-1 while step k runs
        b = step % 2
        if step >= 2:
            ctx4.wait_mark(b)
            np.testing.assert_array_equal(hout[b].array, want[step - 2])
        hin[b].array[:] = rng.random(n, dtype=np.float32)
        want.append(hin[b].array * np.float32(3.0))
        x.scatter_async(hin[b].array)
        y.assign(np.float32(3.0) * dj.lazy(x))
        y.to_array_async(hout[b].array)
        ctx4.mark(b)
    ctx4.sync()
    np.testing.assert_array_equal(hout[0].array, want[4])
    np.testing.assert_array_equal(hout[1].array, want[5])
    for h in hin + hout:
        h.free()


def _load_cgls():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "cgls.py")
    spec = importlib.util.spec_from_file_location("cgls_example2", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_captured_cgls_iteration_equals_the_eager_loop(dj, ctx4, dtype):
    """One CGLS iteration captured as a graph (device scalars, no host synchronisation inside) replayed k times
    gives the iterate of k eager iterations, bit for bit."""
    dj.set_default_context(ctx4)
    mod = _load_cgls()
    rng = np.random.default_rng(21)
    nb, bs = 6, 96
    blocks = [[dj.JopDiagonal((1.0 + rng.random(bs)).astype(dtype)) if (i % 2 == 0 and j % 2 == 1) else
               dj.JopDense(np.asfortranarray((rng.random((bs, bs)) / bs + (np.eye(bs) if i == j else 0)).astype(dtype)))
               for j in range(1, nb + 1)] for i in range(1, nb + 1)]
    A = dj.JopDBlock(blocks, pids=[0, 1, 2, 3], dist=[2, 2], ctx=ctx4)
    x_true = A.domain().rand(rng)
    b = A @ x_true
    xe, he = mod.cgls(A, b, iters=12, tol=0.0)
    xg, hg = mod.cgls_graph(A, b, iters=12, check_every=4, tol=0.0)
    np.testing.assert_array_equal(xg.to_array(), xe.to_array())
    assert hg[-1] == he[-1]
    with pytest.raises(dj.DJetsError, match="synchronises"):
        with ctx4.capture():
            b.norm()
    A.close()


def test_set_param_drops_graphs_captured_under_old_settings(dj, ctx4, O):
    from helpers import build_pair, load_like
    rng = np.random.default_rng(31)
    oop, gop, oblocks = build_pair(dj, O, rng, [40, 50], [30, 20, 10], lambda i, j: "dense", np.float64, [0, 1, 2, 3],
                                   [2, 2], ctx=ctx4)
    xo = oop.dom.rand(rng)
    xg = load_like(gop.domain().zeros(), xo)
    y = gop.range().zeros()
    ref = None
    for p2p in (1, 0, 1):
        ctx4.set_param("p2p", p2p)
        for _ in range(3):                                       # third apply replays a graph captured under this setting
            gop.mul(y, xg)
        got = y.to_array()
        ref = got if ref is None else ref
        np.testing.assert_array_equal(got, ref)
    ctx4.set_param("p2p", 1)
    gop.close()
