"""Static checks of the Julia ccall shim (it cannot be executed in this image: no Julia).  Every
symbol it binds must be declared in include/djets.h with the same number of arguments, and every
argument type must be the Julia spelling of the C type at that position."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "distributedjets.jl_b200", "julia", "DistributedJetsB200.jl")
HEADER = os.path.join(ROOT, "include", "djets.h")

C2JL = {
    "int32_t": {"Int32"}, "int64_t": {"Int64"}, "double": {"Float64"}, "float": {"Float32"},
    "int32_t*": {"Ptr{Int32}"}, "int64_t*": {"Ptr{Int64}"}, "double*": {"Ptr{Float64}"}, "float*": {"Ptr{Float32}"},
    "uint8_t*": {"Ptr{UInt8}"}, "void*": {"Ptr{Cvoid}"}, "void**": {"Ptr{Ptr{Cvoid}}"}, "char*": {"Cstring", "Ptr{UInt8}"},
    "djets_ctx": {"Ptr{Cvoid}"}, "djets_vec": {"Ptr{Cvoid}"}, "djets_op": {"Ptr{Cvoid}"},
    "djets_ctx*": {"Ptr{Ptr{Cvoid}}"}, "djets_vec*": {"Ptr{Ptr{Cvoid}}"}, "djets_op*": {"Ptr{Ptr{Cvoid}}"},
    "djets_scalar": {"Ptr{Cvoid}"}, "djets_graph": {"Ptr{Cvoid}"}, "djets_scalar*": {"Ptr{Ptr{Cvoid}}"},
    "djets_graph*": {"Ptr{Ptr{Cvoid}}"}, "djets_allgather_fn": {"Ptr{Cvoid}"},
}


def header_signatures():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(?:int32_t|const char\*)\s+(djets_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.replace("const", "").strip()
                stars = a.count("*")
                base = re.sub(r"[\*\s]+[A-Za-z_0-9]*$", "", a).strip() if stars == 0 else a.split("*")[0].strip()
                if stars == 0:
                    base = " ".join(a.split()[:-1]) if len(a.split()) > 1 else a
                types.append(base + "*" * stars)
        sigs[name] = types
    return sigs


def split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def shim_bindings():
    src = open(SHIM).read()
    found = []
    for m in re.finditer(r"@dj\s+(djets_[a-z0-9_]+)\s+\(((?:[^()]|\([^()]*\))*)\)", src):
        found.append((m.group(1), split_top_level(m.group(2))))
    for m in re.finditer(r"ccall\(\(:(djets_[a-z0-9_]+),\s*libdjets\),\s*(\w+),\s*\(((?:[^()]|\([^()]*\))*)\)", src):
        found.append((m.group(1), split_top_level(m.group(3))))
    return found


def test_shim_symbols_and_arities_match_the_header():
    sigs = header_signatures()
    assert len(sigs) >= 50
    bindings = shim_bindings()
    assert len(bindings) >= 30
    for name, jl_types in bindings:
        assert name in sigs, f"{name} is bound in the Julia shim but not declared in include/djets.h"
        c_types = sigs[name]
        assert len(jl_types) == len(c_types), f"{name}: {len(jl_types)} Julia argument types vs {len(c_types)} in the header"
        for k, (jt, ct) in enumerate(zip(jl_types, c_types)):
            assert jt in C2JL[ct], f"{name} argument {k + 1}: Julia {jt} vs C {ct}"


def test_shim_covers_the_hot_path_entry_points():
    bound = {name for name, _ in shim_bindings()}
    for needed in ("djets_ctx_create", "djets_vec_create", "djets_vec_similar", "djets_vec_destroy", "djets_vec_getblock",
                   "djets_vec_setblock", "djets_vec_getindex", "djets_vec_collect", "djets_vec_scatter", "djets_vec_fill",
                   "djets_vec_lincomb", "djets_vec_eval", "djets_vec_dot", "djets_vec_norm", "djets_vec_reduce",
                   "djets_vec_dot_s", "djets_vec_norm_s", "djets_scalar_create", "djets_scalar_op", "djets_scalar_get",
                   "djets_vec_scatter_async", "djets_vec_collect_async", "djets_ctx_mark", "djets_ctx_wait_mark",
                   "djets_ctx_capture_begin", "djets_ctx_capture_end", "djets_graph_launch",
                   "djets_op_create", "djets_op_set_dense", "djets_op_set_diag", "djets_op_finalize", "djets_op_destroy",
                   "djets_op_mul", "djets_op_mul_adjoint", "djets_op_getblock", "djets_op_perfstat"):
        assert needed in bound, needed


def test_shim_lowers_broadcasts_like_the_python_mirror():
    """The two hosts must agree on the byte-code: same operation numbering as include/djets.h, `dest .= scalar` is a
    fill (test/runtests.jl:329), and the Float64 flag follows the scalar / array element types."""
    src = open(SHIM).read()
    hdr = open(HEADER).read()
    import djets_loader
    djets_loader.load()
    import djets_b200._lib as L
    names = ["IN", "SC", "ADD", "SUB", "MUL", "DIV", "POW", "MIN", "MAX", "NEG", "ABS", "SQRT", "SQUARE", "CUBE", "EXP", "LOG",
             "ADDS", "SUBS", "RSUBS", "MULS", "RMULS", "DIVS", "RDIVS", "POWS", "MINS", "MAXS"]
    for k, n in enumerate(names, start=1):
        assert re.search(rf"DJETS_EV_{n} = {k}\b", hdr), n
        assert getattr(L, "EV_" + n) == k
    jl = re.search(r"const (EV_IN,.*?) = UInt8\.\(1:26\)", src, flags=re.S).group(1)
    assert [t.strip() for t in jl.replace("\n", " ").split(",")] == ["EV_" + n for n in names]
    assert "return fill!(dest, unref(bc.args[1]))" in src
    assert "(a isa Float64 || a isa Complex{Float64}) && (p.wide = true)" in src and "defaultgrid(nblks, length(pids))" in src


def test_shim_keeps_the_reference_exports():
    src = open(SHIM).read()
    assert re.search(r"^export DBArray, JetDSpace, blockmap, localblockindices$", src, flags=re.M)   # src:863


def test_python_ctypes_signatures_match_the_header():
    """The ctypes mirror binds every function with the header's arity and argument types."""
    import ctypes as C
    import djets_loader
    djets_loader.load()
    import djets_b200._lib as L
    c2ct = {
        "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double, "float": C.c_float,
        "int32_t*": C.POINTER(C.c_int32), "int64_t*": C.POINTER(C.c_int64), "double*": C.POINTER(C.c_double),
        "float*": C.POINTER(C.c_float), "uint8_t*": C.POINTER(C.c_uint8), "void*": C.c_void_p,
        "void**": C.POINTER(C.c_void_p), "char*": C.c_char_p, "djets_ctx": C.c_void_p, "djets_vec": C.c_void_p,
        "djets_op": C.c_void_p, "djets_ctx*": C.POINTER(C.c_void_p), "djets_vec*": C.POINTER(C.c_void_p),
        "djets_op*": C.POINTER(C.c_void_p), "djets_scalar": C.c_void_p, "djets_graph": C.c_void_p,
        "djets_scalar*": C.POINTER(C.c_void_p), "djets_graph*": C.POINTER(C.c_void_p), "djets_allgather_fn": C.c_void_p,
    }
    sigs = header_signatures()
    for name, argtypes in L.SIGNATURES.items():
        want = [c2ct[t] for t in sigs[name]]
        assert len(argtypes) == len(want), name
        for k, (a, w) in enumerate(zip(argtypes, want)):
            assert a is w or a == w, f"{name} argument {k + 1}: {a} vs {w}"


def test_every_dj_call_passes_as_many_arguments_as_it_declares():
    """`@dj f (T1, T2, ...) a1 a2 ...` expands to `ccall(f, Int32, (T1, T2, ...), a1, a2, ...)`: a call whose argument
    list is shorter or longer than its type tuple would only fail when Julia first runs it.  Arguments are the
    space-separated expressions up to the end of the statement (newline, `;` or the bracket that encloses the macro
    call); an argument with a top-level space (`a + b`) would be miscounted here and is also ambiguous to Julia's
    macro-call parser, so none is allowed."""
    from test_julia_recipes import _strip_julia_literals
    src = _strip_julia_literals(open(SHIM).read())
    calls = 0
    for m in re.finditer(r"@dj\s+(djets_[a-z0-9_]+)\s+\(", src):
        i = j = m.end() - 1
        depth = 0
        while True:                                        # the type tuple
            if src[j] in "({[":
                depth += 1
            elif src[j] in ")}]":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        types = split_top_level(src[i + 1:j])
        args, cur, depth, k = [], "", 0, j + 1
        while k < len(src):
            c = src[k]
            if c in "({[":
                depth += 1
            elif c in ")}]":
                if depth == 0:
                    break
                depth -= 1
            if depth == 0 and c in "\n;":
                break
            if depth == 0 and c in " \t":
                if cur:
                    args.append(cur)
                cur = ""
            else:
                cur += c
            k += 1
        if cur:
            args.append(cur)
        calls += 1
        assert len(args) == len(types), f"@dj {m.group(1)}: {len(types)} argument types, {len(args)} arguments {args}"
    assert calls >= 40


def test_shim_defines_no_function_under_a_name_jets_owns():
    """Inside `module DistributedJetsB200` every name Jets exports is in scope through `using Jets`; a method meant to
    EXTEND one must be written `Jets.name(...)`, and a helper of the shim must not reuse such a name unqualified (it
    would shadow, or fail to define once the imported binding was used).  The names are the ones the reference itself
    extends with a `Jets.` prefix (src:161-164, 383-412, 414-415, 730, 805-827) plus the accessors its tests call."""
    from test_julia_recipes import _strip_julia_literals
    src = _strip_julia_literals(open(SHIM).read())
    jets_owned = {"getblock", "getblock!", "setblock!", "indices", "isblockop", "nblocks", "perfstat", "point!", "space",
                  "state", "state!", "jet", "domain", "shape", "jacobian", "jacobian!", "JopBlock", "JetBlock", "close"}
    defined = set(re.findall(r"^\s*(?:function\s+)?([A-Za-z_][\w!′]*)\s*(?:\{[^}\n]*\})?\(", src, flags=re.M))
    clashes = defined & jets_owned
    assert not clashes, f"unqualified definitions of names Jets owns: {sorted(clashes)}"
    for name in ("getblock", "getblock!", "setblock!", "indices", "nblocks", "space", "isblockop", "JopBlock"):
        assert re.search(rf"\bJets\.{re.escape(name)}\(", src), f"the shim should extend Jets.{name}"
