"""Committed golden vectors (tests/golden/operator_cases.npz, made by tests/golden/make_golden.py):
the oracle must keep reproducing them bit for bit (CPU), the CUDA path must match them within the
BASELINE tolerance (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

DATA = np.load(os.path.join(HERE, "golden", "operator_cases.npz"))


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_golden(name):
    fresh = G.build(name)
    for key, val in fresh.items():
        assert np.array_equal(DATA[key], val), key


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_matches_golden(dj, ctx4, name):
    rs, cs, grid, isdiag, dtype, rule = G.CASES[name]
    dtype = np.dtype(dtype)

    def blk(i, j):
        k = G.kind(rule, i, j)
        if k == "dense":
            return dj.JopDense(DATA[f"{name}/A_{i}_{j}"])
        if k == "diag":
            return dj.JopDiagonal(DATA[f"{name}/D_{i}_{j}"])
        return dj.JopZeroBlock(dtype, rs[i - 1], cs[j - 1])

    A = dj.JopDBlock(blk, dims=(len(rs), len(cs)), pids=list(range(grid[0] * grid[1])), dist=grid, isdiag=isdiag,
                     ctx=ctx4)
    x = A.domain().zeros().scatter(np.ascontiguousarray(DATA[f"{name}/x"]))
    y = A.range().zeros().scatter(np.ascontiguousarray(DATA[f"{name}/y"]))
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for got, want, lens in (((A @ x).to_array(), DATA[f"{name}/Ax"], rs), ((A.T @ y).to_array(), DATA[f"{name}/Aty"], cs)):
        o = 0
        for n in lens:
            den = np.linalg.norm(want[o:o + n].astype(np.float64))
            err = np.linalg.norm(got[o:o + n].astype(np.float64) - want[o:o + n].astype(np.float64))
            assert err <= tol * max(den, 1e-300), (name, o, err / max(den, 1e-300))
            o += n
