"""CPU-only tests of the product's host logic: partition rules against the oracle, the C-ABI library
(loads, exports every symbol include/djets.h declares, host-only planning entry points), and the loud
failure without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "djets.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(djets_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(dj):
    lib = C.CDLL(dj.LIB_PATH)
    names = header_functions()
    assert len(names) >= 45
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/djets.h but not exported"


def test_python_binding_covers_the_header(dj):
    import djets_b200._lib as L
    bound = set(L.SIGNATURES) | {"djets_version", "djets_last_error"}
    assert bound == set(header_functions())


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "djets.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                      # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_version_and_error_string(dj):
    import djets_b200._lib as L
    assert L.lib.djets_version() >= 100
    rc = L.lib.djets_even_cuts(-1, 2, (C.c_int64 * 3)())
    assert rc == L.ERR_INVALID and b"even_cuts" in L.lib.djets_last_error()


def test_even_cuts_match_oracle(dj, O):
    for n in range(0, 40):
        for k in range(1, 9):
            cuts = dj.even_cuts(n, k)
            if n >= k:                                                      # DistributedArrays needs sz >= nc
                assert [c + 1 for c in cuts] == O.defaultdist_cuts(n, k), (n, k)
            assert cuts[0] == 0 and cuts[-1] == n and all(b >= a for a, b in zip(cuts, cuts[1:]))
    assert dj.even_ranges(7, 4) == [range(1, 3), range(3, 5), range(5, 7), range(7, 8)]     # test:172-221
    assert dj.even_ranges(10, 4) == [range(1, 4), range(4, 7), range(7, 9), range(9, 11)]   # test:514-527


def test_default_grid_matches_oracle(dj, O):
    for dims in ([3, 4], [3, 5], [7, 1], [1, 7], [50, 10], [20, 20], [2, 2], [64, 64], [128, 128], [5], [7]):
        for npids in range(1, 13):
            assert dj.default_grid(dims, npids) == O.defaultdist_chunks(dims, npids), (dims, npids)
    assert dj.default_grid([7], 4) == [4] and dj.default_grid([3, 5], 4) == [2, 2]


def test_grid_and_prescribed_ranges(dj, O):
    assert dj.grid_of([2, 3, 4, 5], 2, 2) == [[2, 4], [3, 5]] == O.grid_pids([2, 3, 4, 5], 2, 2)   # test:543
    assert dj.cuts_from_ranges([range(1, 3), range(3, 11)]) == [0, 2, 10]                         # test:39-52
    assert [c - 1 for c in O.prescribedist([range(1, 3), range(3, 11)])] == [0, 2, 10]
    assert dj.ranges_from_cuts([0, 2, 10]) == [range(1, 3), range(3, 11)]
    with pytest.raises(ValueError):
        dj.cuts_from_ranges([range(1, 3), range(4, 11)])
    import djets_b200._lib as L
    r = C.c_int32()
    L.call("djets_grid_rank", 2, 2, 1, 0, L.arr_i32([7, 8, 9, 10]), C.byref(r))
    assert r.value == 8
    L.call("djets_grid_rank", 2, 2, 0, 1, None, C.byref(r))
    assert r.value == 2


def test_plan_tiles_shapes(dj):
    import djets_b200._lib as L

    def plan(dtype, m, n, adj):
        tr, tc, nt = C.c_int64(), C.c_int64(), C.c_int64()
        L.call("djets_plan_tiles", dtype, m, n, adj, C.byref(tr), C.byref(tc), C.byref(nt))
        return tr.value, tc.value, nt.value

    tr, tc, nt = plan(L.F32, 4096, 4096, 0)
    assert tr % 4 == 0 and tc <= 1024 and nt == -(-4096 // tr) * -(-4096 // tc)
    tr, tc, nt = plan(L.F64, 1000, 1000, 0)
    assert tr % 2 == 0 and nt == -(-1000 // tr) * -(-1000 // tc)
    tr, tc, nt = plan(L.F64, 10000, 64, 1)              # taller than one y stage: several row panels
    assert tr * 8 <= 32768 and tr % 2 == 0 and nt == -(-10000 // tr) * -(-64 // tc)
    assert plan(L.F32, 0, 5, 0)[2] == 0


def test_chain_rule_for_small_blocks(dj):
    """Blocks per chained tile (include/djets.h djets_plan_chains): 0.5-1 MB per chain, at least four waves of CTAs
    left, nothing for block-diagonal operators, complex types, tall blocks or wide forward blocks."""
    import djets_b200._lib as L

    def rule(dtype, adjoint, nblocks, longest, rows, cols, es, wave=296, chain=0):
        k, pair = C.c_int32(), C.c_int32()
        L.call("djets_plan_chains", dtype, adjoint, nblocks, longest, rows, cols, rows * cols * es, wave, chain,
               C.byref(k), C.byref(pair))
        return k.value, pair.value

    assert rule(L.F64, 1, 90000, 300, 100, 100, 8) == (13, 0)       # 80 KB blocks: thirteen make ~1 MB; 100 rows = two warp-loads
    assert rule(L.F64, 1, 160000, 400, 64, 64, 8) == (16, 1)        # 32 KB blocks: the cap; one warp-load tall: pairs
    assert rule(L.F32, 0, 90000, 300, 128, 128, 4) == (16, 1)
    assert rule(L.F32, 1, 16384, 128, 256, 256, 4) == (2, 0)        # 256 KB blocks, 16384 of them: pairs
    assert rule(L.F64, 1, 14400, 120, 100, 100, 8) == (6, 0)        # too few blocks for 1 MB chains: six make ~512 KB
    assert rule(L.F64, 1, 2400, 49, 100, 100, 8) == (2, 0)          # small cell: keep four waves of CTAs
    assert rule(L.F64, 1, 16, 4, 100, 100, 8)[0] == 1               # tiny operator: one block per CTA
    assert rule(L.F64, 1, 16, 4, 100, 100, 8, chain=4) == (4, 0)    # ... unless forced
    assert rule(L.F64, 1, 16, 4, 100, 100, 8, chain=16)[0] == 4     # never longer than the blocks there are
    assert rule(L.F64, 1, 100000, 1, 64, 64, 8)[0] == 1             # block-diagonal: nothing to chain
    assert rule(L.F64, 1, 90000, 300, 129, 100, 8)[0] == 1          # taller than two warp-loads
    assert rule(L.F32, 1, 90000, 300, 256, 100, 4)[0] > 1           # Float32 warp-loads are 128 rows
    assert rule(L.F64, 0, 90000, 300, 100, 600, 8)[0] == 1          # forward: columns must fit the staging buffer
    assert rule(L.C128, 1, 90000, 300, 64, 64, 16)[0] == 1          # complex blocks are not chained
    assert rule(L.F64, 1, 90000, 300, 100, 100, 8, chain=1)[0] == 1  # switched off
    assert rule(L.F64, 1, 90000, 300, 100, 100, 8, wave=0)[0] == 1  # no device: no rule


def test_exchange_plan_is_complete(dj):
    """Every cell gets each input part it needs exactly once, every non-owner sends exactly one partial
    to the right owner slot, block-diagonal operators exchange nothing."""
    for n1 in range(1, 5):
        for n2 in range(1, 5):
            for adj in (False, True):
                a = dj.plan_exchange(n1, n2, False, adj, 0)
                c = dj.plan_exchange(n1, n2, False, adj, 1)
                need = {(r, cc) for r in range(n1) for cc in range(n2) if (cc != 0 if adj else r != 0)}
                assert {(m[2], m[3]) for m in a} == need and len(a) == len(need)
                for sr, sc, dr, dc, part, slot in a:
                    assert (sr, sc) == ((dr, 0) if adj else (0, dc)) and part == (dr if adj else dc) and slot == -1
                senders = {(r, cc) for r in range(n1) for cc in range(n2) if (r != 0 if adj else cc != 0)}
                assert {(m[0], m[1]) for m in c} == senders and len(c) == len(senders)
                slots = {}
                for sr, sc, dr, dc, part, slot in c:
                    assert (dr, dc) == ((0, sc) if adj else (sr, 0)) and part == (sc if adj else sr)
                    slots.setdefault((dr, dc), []).append(slot)
                for owner, s in slots.items():
                    assert sorted(s) == list(range((n1 if adj else n2) - 1))
            assert dj.plan_exchange(n1, 1, True, False, 0) == [] == dj.plan_exchange(n1, 1, True, True, 1)


def test_no_cpu_fallback(dj):
    """Without a CUDA device every compute entry point refuses: the product never runs on the CPU."""
    n = C.c_int32(0)
    import djets_b200._lib as L
    rc = L.lib.djets_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dj.DJetsError) as e:
        dj.Context(world=1)
    assert e.value.code == L.ERR_CUDA
    import subprocess, sys
    # the package import itself must fail loudly when the shared library is absent
    probe = ("import sys, importlib.util, os, tempfile, shutil\n"
             "src = %r\n"
             "tmp = tempfile.mkdtemp()\n"
             "dst = os.path.join(tmp, 'pkg'); shutil.copytree(src, dst, ignore=shutil.ignore_patterns('*.so', 'csrc'))\n"
             "spec = importlib.util.spec_from_file_location('pkgx', os.path.join(dst, '__init__.py'), submodule_search_locations=[dst])\n"
             "m = importlib.util.module_from_spec(spec); sys.modules['pkgx'] = m\n"
             "try:\n    spec.loader.exec_module(m)\nexcept ImportError as e:\n    print('IMPORTERROR', e)\n") % os.path.join(
                 ROOT, "distributedjets.jl_b200")
    out = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True)
    assert "IMPORTERROR" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distributedjets.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "djets_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def _build_c_client(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "c_abi_smoke")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-ldl", "-lm"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_is_plain_c_and_refuses_without_gpu(dj, tmp_path):
    """include/djets.h compiles as C99 (-pedantic -Werror); the plain-C client loads the library and,
    with no GPU, sees djets_ctx_create refuse with DJETS_ERR_CUDA."""
    import subprocess
    exe = _build_c_client(tmp_path)
    out = subprocess.run([exe, dj.LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "version" in out.stdout


@pytest.mark.gpu
def test_plain_c_client_on_gpu(dj, tmp_path):
    import subprocess
    exe = _build_c_client(tmp_path)
    out = subprocess.run([exe, dj.LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0 and "c abi ok" in out.stdout, out.stdout + out.stderr


def test_broadcast_lowering_without_a_device(dj):
    """The Python mirror lowers lazy expressions to the postfix program of djets_vec_eval; the library validates it and
    runs its peephole pass on the host (djets_plan_eval): no GPU involved.  Vectors are stand-ins (only identity and
    dtype matter to the lowering)."""
    import ctypes as C
    import numpy as np
    from djets_b200 import _lib as L, core

    class FakeVec:
        def __init__(self, dtype):
            self.dtype = np.dtype(dtype)
            self.h = None
    x, y, z = (core.Expr("in", FakeVec(np.float64)) for _ in range(3))

    def plan(expr):
        prog, vecs, scal, dscal, f64 = core._lower(expr)
        buf = (C.c_uint8 * len(prog))(*prog)
        depth, fused = C.c_int32(), C.c_int32()
        L.call("djets_plan_eval", buf, len(prog) // 2, len(vecs), len(scal), C.byref(depth), C.byref(fused))
        return prog, len(vecs), scal, depth.value, fused.value

    # benchmark/benchmarks.jl:43  d4 .= a1*d1 .+ a2*d2 .- a3*d3: 8 operations, 4 after the operands are folded in
    prog, nin, scal, depth, fused = plan(0.3 * x + 0.5 * y - 0.7 * z)
    assert nin == 3 and scal == [0.3, 0.5, 0.7] and len(prog) // 2 == 8 and fused == 4 and depth == 1
    assert prog[:4] == [L.EV_IN, 0, L.EV_RMULS, 0] and prog[-2] == L.EV_SUB
    # x.^2, x.^3 are products (Julia's literal_pow); other powers call pow
    assert plan(x ** 2)[0] == [L.EV_IN, 0, L.EV_SQUARE, 0] and plan(x ** 3)[0][2] == L.EV_CUBE
    assert plan(x ** 1.5)[0][2] == L.EV_POWS
    # scalar on the left of a non-commutative operator
    assert plan(2.0 - x)[0] == [L.EV_IN, 0, L.EV_RSUBS, 0] and plan(2.0 / x)[0][2] == L.EV_RDIVS
    # the same vector is one input however often it appears; scalars fold like Julia's per-element arithmetic
    prog, nin, scal, depth, fused = plan((x * x + x) * (2.0 * 3.0))
    assert nin == 1 and scal == [6.0]
    # depth grows with parenthesised right operands, never past the limit silently
    deep = x
    for _ in range(3):
        deep = y * (z + deep)
    assert plan(deep)[3] >= 2
    import pytest
    too_long = x
    for _ in range(30):
        too_long = too_long * y + z
    with pytest.raises(dj.DJetsError):
        plan(too_long)
    with pytest.raises(dj.DJetsError):                      # malformed programs are refused
        buf = (C.c_uint8 * 2)(L.EV_ADD, 0)
        L.call("djets_plan_eval", buf, 1, 0, 0, None, None)
