"""When the reference itself has been run (tests/golden/make_golden.jl -> operator_cases_julia.json) its outputs pin
the oracle (CPU) and the CUDA path (GPU) at the BASELINE tolerances: <= 1e-12 in Float64, <= 1e-5 in Float32,
norm-wise per block.  Without that file the tests skip and DESIGN.md keeps saying "parity unpinned"."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
JULIA = os.path.join(HERE, "golden", "operator_cases_julia.json")
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

needs_julia = pytest.mark.skipif(not os.path.exists(JULIA),
                                 reason="tests/golden/operator_cases_julia.json absent: the reference has not been run "
                                        "(no Julia in this image); run `julia tests/golden/make_golden.jl`")
DATA = np.load(os.path.join(HERE, "golden", "operator_cases.npz"))


def blockwise_close(got, want, lens, tol, what):
    o = 0
    for n in lens:
        den = np.linalg.norm(np.asarray(want[o:o + n], dtype=np.float64))
        err = np.linalg.norm(np.asarray(got[o:o + n], dtype=np.float64) - np.asarray(want[o:o + n], dtype=np.float64))
        assert err <= tol * max(den, 1e-300), (what, o, err / max(den, 1e-300))
        o += n


@needs_julia
@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_matches_the_reference(name):
    ref = json.load(open(JULIA))["cases"][name]
    rs, cs, grid, isdiag, dtype, rule = G.CASES[name]
    tol = 1e-12 if dtype == "float64" else 1e-5
    blockwise_close(DATA[f"{name}/Ax"], ref["Ax"], rs, tol, "Ax")
    blockwise_close(DATA[f"{name}/Aty"], ref["Aty"], cs, tol, "Aty")
    x, y = DATA[f"{name}/x"], DATA[f"{name}/y"]
    assert abs(float(np.dot(x.astype(np.float64), x.astype(np.float64))) - ref["dot_x_x"]) <= 10 * tol * abs(ref["dot_x_x"])
    assert abs(float(np.linalg.norm(y.astype(np.float64))) - ref["norm_y"]) <= 10 * tol * ref["norm_y"]
    assert ref["nblocks"] == [len(rs), len(cs)]


@needs_julia
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_matches_the_reference(dj, ctx4, name):
    ref = json.load(open(JULIA))["cases"][name]
    rs, cs, grid, isdiag, dtype, rule = G.CASES[name]
    dtype = np.dtype(dtype)
    tol = 1e-12 if dtype == np.float64 else 1e-5

    def blk(i, j):
        k = G.kind(rule, i, j)
        if k == "dense":
            return dj.JopDense(DATA[f"{name}/A_{i}_{j}"])
        if k == "diag":
            return dj.JopDiagonal(DATA[f"{name}/D_{i}_{j}"])
        return dj.JopZeroBlock(dtype, rs[i - 1], cs[j - 1])

    A = dj.JopDBlock(blk, dims=(len(rs), len(cs)), pids=list(range(grid[0] * grid[1])), dist=grid, isdiag=isdiag, ctx=ctx4)
    x = A.domain().zeros().scatter(np.ascontiguousarray(DATA[f"{name}/x"]))
    y = A.range().zeros().scatter(np.ascontiguousarray(DATA[f"{name}/y"]))
    blockwise_close((A @ x).to_array(), ref["Ax"], rs, tol, "Ax")
    blockwise_close((A.T @ y).to_array(), ref["Aty"], cs, tol, "Aty")
    assert abs(float(x.dot(x)) - ref["dot_x_x"]) <= 10 * tol * abs(ref["dot_x_x"])
    assert abs(float(y.norm()) - ref["norm_y"]) <= 10 * tol * ref["norm_y"]
    assert abs(float(y.norm(1)) - ref["norm1_y"]) <= 10 * tol * ref["norm1_y"]
    assert float(y.norm(np.inf)) == np.dtype(dtype).type(ref["normInf_y"])
    bm = [[list(r), list(c)] for col in zip(*A.blockmap()) for (r, c) in col]      # vec(blockmap(F)): column-major grid
    assert bm == ref["blockmap"]
