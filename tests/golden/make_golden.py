"""Generates tests/golden/operator_cases.npz: seeded inputs and the oracle's outputs for small
operator cases.  The reference itself cannot be executed in this image (no Julia), so
these vectors come from the oracle restatement (oracle/djets_oracle.py) -- "parity unpinned" beyond
the reference's own isapprox checks -- and freeze it: any later change of the oracle or of the CUDA
path that moves a result shows up against these files.

    python tests/golden/make_golden.py

It also writes operator_cases_inputs.json: the same seeded INPUTS (blocks, x, y) in a form the Julia side can read
without extra packages.  tests/golden/make_golden.jl feeds them to the unmodified reference and writes
operator_cases_julia.json; once that file exists tests/test_golden_julia.py pins the oracle and the CUDA path to
the reference's own outputs (and "parity unpinned" can be lifted).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import djets_oracle as O  # noqa: E402

CASES = {
    # name: (row sizes, col sizes, grid, isdiag, dtype, kind rule)
    "hetero_3x4_grid21_f64": ([10] * 3, [10] * 4, [2, 1], False, "float64", "hetero"),
    "hetero_3x5_grid22_f64": ([10] * 3, [10] * 5, [2, 2], False, "float64", "hetero"),
    "diag_4x4_dense10x5_f64": ([10] * 4, [5] * 4, [4, 1], True, "float64", "blockdiag"),
    "dense_3x2_odd_f32": ([5, 7, 3], [9, 6], [2, 2], False, "float32", "dense"),
    "dense_2x2_66_f64": ([66, 24], [36, 50], [2, 2], False, "float64", "dense"),
}


def kind(rule, i, j):
    if rule == "hetero":
        if i in (1, 2, 3) and j in (1, 2, 3):
            return "diag"
        return "zero" if (i, j) == (3, 4) else "dense"
    if rule == "blockdiag":
        return "dense" if i == j else "zero"
    return "dense"


def build(name):
    rs, cs, grid, isdiag, dtype, rule = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    dtype = np.dtype(dtype)
    blocks, payload = [], {}
    for i in range(1, len(rs) + 1):
        row = []
        for j in range(1, len(cs) + 1):
            k = kind(rule, i, j)
            if k == "dense":
                A = np.asfortranarray(rng.random((rs[i - 1], cs[j - 1])).astype(dtype))
                payload[f"{name}/A_{i}_{j}"] = A
                row.append(O.ODense(A))
            elif k == "diag":
                d = rng.random(cs[j - 1]).astype(dtype)
                payload[f"{name}/D_{i}_{j}"] = d
                row.append(O.ODiag(d))
            else:
                row.append(O.OZero(dtype, (rs[i - 1],), (cs[j - 1],)))
        blocks.append(row)
    op = O.OJopDBlock(blocks, list(range(grid[0] * grid[1])), grid, isdiag=isdiag)
    x = rng.random(sum(cs)).astype(dtype)
    y = rng.random(sum(rs)).astype(dtype)
    xo, yo = op.dom.zeros(), op.rng.zeros()
    o = 0
    for ib, n in enumerate(cs, start=1):
        xo.setblock(ib, x[o:o + n]); o += n
    o = 0
    for ib, n in enumerate(rs, start=1):
        yo.setblock(ib, y[o:o + n]); o += n
    payload[f"{name}/x"] = x
    payload[f"{name}/y"] = y
    payload[f"{name}/Ax"] = op.mul(op.rng.zeros(), xo).to_array()
    payload[f"{name}/Aty"] = op.mul_adjoint(op.dom.zeros(), yo).to_array()
    return payload


def export_inputs(payload) -> dict:
    """Inputs of every case as plain JSON (floats printed with repr round-trip exactly)."""
    doc = {}
    for name, (rs, cs, grid, isdiag, dtype, rule) in CASES.items():
        blocks = {}
        for key, val in payload.items():
            case, _, item = key.partition("/")
            if case != name or item.count("_") != 2 or item[0] not in "AD":
                continue
            kind, i, j = item.split("_")
            blocks[f"{i}_{j}"] = {"kind": "dense" if kind == "A" else "diag", "shape": list(val.shape),
                                  "data": [float(v) for v in np.asarray(val).reshape(-1, order="F")]}
        doc[name] = {"dtype": dtype, "row_sizes": rs, "col_sizes": cs, "grid": grid, "isdiag": isdiag, "blocks": blocks,
                     "x": [float(v) for v in payload[f"{name}/x"]], "y": [float(v) for v in payload[f"{name}/y"]]}
    return doc


if __name__ == "__main__":
    import json
    out = {}
    for name in CASES:
        out.update(build(name))
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "operator_cases.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
    jpath = os.path.join(here, "operator_cases_inputs.json")
    with open(jpath, "w") as f:
        json.dump(export_inputs(out), f)
    print(jpath, os.path.getsize(jpath), "bytes")
