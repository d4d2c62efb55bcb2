"""Static checks of the Julia-side pinning recipes (baseline/ref_bench.jl, tests/golden/make_golden.jl): Julia is
absent from this image, so they cannot be executed here; what can be checked is that every DistributedJets / Jets /
DistributedArrays name they call is one the reference itself exports, defines or calls (its source, tests and
benchmark under /root/reference), that the fixtures are the reference's own, and that the exported inputs agree
with the committed golden vectors."""
import json
import os
import re
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
RECIPES = [os.path.join(ROOT, "baseline", "ref_bench.jl"), os.path.join(ROOT, "tests", "golden", "make_golden.jl")]

# names of the reference's API (and of the packages it builds on) the recipes rely on
API = ["DBArray", "blockmap", "@blockop", "DArray", "JopLn", "JetSpace", "JopZeroBlock", "mul!", "domain", "range",
       "setblock!", "nblocks", "workers", "addprocs", "dot", "norm", "convert", "isdiag=true"]


def strip_comments(src: str) -> str:
    return "\n".join(line.split("#", 1)[0] for line in src.splitlines())


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_recipes_use_only_names_the_reference_knows():
    corpus = ""
    for rel in ("src/DistributedJets.jl", "test/runtests.jl", "benchmark/benchmarks.jl", "docs/src/index.md"):
        corpus += open(os.path.join(REF, rel)).read()
    exports = re.search(r"^export (.*)$", open(os.path.join(REF, "src/DistributedJets.jl")).read(), flags=re.M).group(1)
    assert {"DBArray", "JetDSpace", "blockmap", "localblockindices"} == {e.strip() for e in exports.split(",")}  # src:863
    for path in RECIPES:
        code = strip_comments(open(path).read())
        for name in API:
            if name in code:
                assert name in corpus, f"{os.path.basename(path)} uses {name}, which the reference never mentions"
        # every identifier followed by '(' that is not defined in the recipe itself must occur in the reference or be Base
        defined = set(re.findall(r"(?:function\s+|^|\s)([A-Za-z_][\w′!]*)\s*\([^)]*\)\s*=", code, flags=re.M))
        defined |= set(re.findall(r"function\s+([A-Za-z_][\w′!]*)", code))
        base = {"println", "parse", "findfirst", "error", "min", "max", "rand", "zeros", "string", "Dict", "open", "joinpath",
                "reshape", "enumerate", "length", "size", "eltype", "haskey", "split", "collect", "first", "last", "vec",
                "Float64", "Float32", "Int", "T", "rmprocs", "nworkers", "sizeof", "JSON", "isdiag", "Tuple", "typeof", "print",
                "parsefile", "json", "Any", "Dict{String,Any}", "if", "for", "elseif"}
        for call in set(re.findall(r"([A-Za-z_][\w′!]*)\s*\(", code)):
            if call in defined or call in base or call.startswith("JSON"):
                continue
            assert call in corpus, f"{os.path.basename(path)} calls {call}(...), unknown to the reference and to Base"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_recipe_fixtures_are_the_references_own():
    """JopBaz / JopFoo in the recipes are the fixtures of test/runtests.jl:8-12, 30-36 (typed by the payload)."""
    ref_test = open(os.path.join(REF, "test/runtests.jl")).read()
    for path in RECIPES:
        code = open(path).read()
        for line in ("JopBaz_df!(d,m;A,kwargs...) = d .= A*m", "JopBaz_df′!(m,d;A,kwargs...) = m .= A'*d"):
            assert line in code and line in ref_test
    gold = open(RECIPES[1]).read()
    assert "JopFoo_df!(d,m;diagonal,kwargs...) = d .= diagonal .* m" in gold
    assert "JopFoo_df!(d,m;diagonal,kwargs...) = d .= diagonal .* m" in ref_test


def test_exported_inputs_match_the_golden_npz():
    data = np.load(os.path.join(ROOT, "tests", "golden", "operator_cases.npz"))
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "operator_cases_inputs.json")))
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as G
    assert set(doc) == set(G.CASES)
    for name, case in doc.items():
        dt = np.dtype(case["dtype"])
        assert np.array_equal(np.array(case["x"], dtype=dt), data[f"{name}/x"])
        assert np.array_equal(np.array(case["y"], dtype=dt), data[f"{name}/y"])
        for key, blk in case["blocks"].items():
            tag = ("A_" if blk["kind"] == "dense" else "D_") + key
            arr = np.array(blk["data"], dtype=dt).reshape(blk["shape"], order="F")
            assert np.array_equal(arr, data[f"{name}/{tag}"])


def test_bench_prefers_julia_when_present():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'shutil.which("julia")' in src and "baseline/ref_bench.jl" in src.replace('", "', "/").replace("os.path.join(ROOT, ", "")
    assert '"kind": kind' in src and 'kind = "julia"' in src.replace("cores, kind = float", "x = float") or "\"julia\"" in src
    if shutil.which("julia") is None:
        pytest.skip("no julia on PATH: the reference arm falls back to the NumPy/OpenBLAS restatement")


# ---- block structure of the Julia files (they cannot be parsed by Julia here) ---------------------------------------
_OPENERS = {"type", "function", "if", "for", "while", "let", "begin", "do", "struct", "module", "try", "quote", "macro",
            "baremodule"}


def _strip_julia_literals(s: str) -> str:
    """Drops comments (`#`, `#= =#`) and replaces string / character literals by empty ones."""
    out, i, n = [], 0, len(s)
    while i < n:
        c = s[i]
        if s.startswith("#=", i):
            j = s.find("=#", i + 2)
            i = n if j < 0 else j + 2
        elif c == "#":
            j = s.find("\n", i)
            i = n if j < 0 else j
        elif s.startswith('"""', i):
            j = s.find('"""', i + 3)
            i = n if j < 0 else j + 3
            out.append('""')
        elif c == '"':
            j = i + 1
            while j < n and s[j] != '"':
                j += 2 if s[j] == "\\" else 1
            i = j + 1
            out.append('""')
        elif c == "'" and i + 2 < n and (s[i + 2] == "'" or (s[i + 1] == "\\" and i + 3 < n and s[i + 3] == "'")):
            i += 3 if s[i + 2] == "'" else 4
            out.append("' '")
        else:
            out.append(c)
            i += 1
    return "".join(out)


def julia_blocks_balanced(path: str):
    """(ok, message): every block opener outside brackets has its `end`, every bracket closes.  `for` / `if` inside
    brackets are comprehensions / generators, `end` inside brackets is the last index."""
    stack, line = [], 1
    for m in re.finditer(r"\n|[\[\]\(\)\{\}]|(?<![\w:.!])(?:[A-Za-z_]\w*)", _strip_julia_literals(open(path).read())):
        tok = m.group(0)
        if tok == "\n":
            line += 1
        elif tok in "[({":
            stack.append((tok, line))
        elif tok in "])}":
            if not stack or stack[-1][0] not in "[({":
                return False, f"unbalanced bracket at line {line}"
            stack.pop()
        else:
            in_brackets = bool(stack) and stack[-1][0] in "[({"
            if tok in _OPENERS and not (in_brackets and tok in ("for", "if")):
                stack.append((tok, line))
            elif tok == "end" and not in_brackets:
                if not stack:
                    return False, f"`end` without an opener at line {line}"
                stack.pop()
    return (not stack), (f"unclosed {stack[-3:]}" if stack else "balanced")


JULIA_FILES = RECIPES + [os.path.join(ROOT, "distributedjets.jl_b200", "julia", "DistributedJetsB200.jl")]


@pytest.mark.parametrize("path", JULIA_FILES, ids=[os.path.basename(p) for p in JULIA_FILES])
def test_julia_files_have_balanced_blocks(path):
    ok, msg = julia_blocks_balanced(path)
    assert ok, f"{os.path.basename(path)}: {msg}"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_block_checker_accepts_the_references_own_files(tmp_path):
    """The checker is a heuristic: it must accept Julia that is known to parse, and notice a missing `end`."""
    for rel in ("src/DistributedJets.jl", "test/runtests.jl", "benchmark/benchmarks.jl"):
        ok, msg = julia_blocks_balanced(os.path.join(REF, rel))
        assert ok, f"{rel}: {msg}"
    broken = tmp_path / "broken.jl"
    broken.write_text("function f(x)\n    if x > 0\n        x[end]\n    return [y for y in x if y > 0]\nend\n")
    assert not julia_blocks_balanced(str(broken))[0]
