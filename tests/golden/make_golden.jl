# Pins the oracle's floating-point results to the REFERENCE ITSELF: loads the seeded inputs of
# tests/golden/operator_cases.npz (exported as JSON by tests/golden/make_golden.py --export-inputs), runs the
# UNMODIFIED ChevronETC/DistributedJets.jl on them -- `@blockop` over a DArray of the reference's own test fixtures
# (dense JopBaz test/runtests.jl:30-36, diagonal JopFoo test/runtests.jl:8-12, JopZeroBlock), `mul!` forward and adjoint
# (src/DistributedJets.jl:663-721), `dot` and `norm` on the DBArrays (src:189-223) -- and writes the reference's outputs to
# tests/golden/operator_cases_julia.json.  tests/test_golden_julia.py then holds the oracle and the CUDA path to them.
#
#     julia tests/golden/make_golden.jl          (needs the reference's packages; 4 local workers)
#
# Julia is absent from the image this backend was built in: this file has not been executed there
# (tests/test_julia_recipes.py checks the names it uses against the reference's sources).
using Distributed
addprocs(4)
@everywhere using DistributedArrays, DistributedJets, Jets, LinearAlgebra
using JSON

@everywhere JopFoo_df!(d,m;diagonal,kwargs...) = d .= diagonal .* m
@everywhere function JopFoo(diag)
    spc = JetSpace(eltype(diag), size(diag))
    JopLn(;df! = JopFoo_df!, df′! = JopFoo_df!, dom = spc, rng = spc, s = (diagonal=diag,))
end
@everywhere JopBaz_df!(d,m;A,kwargs...) = d .= A*m
@everywhere JopBaz_df′!(m,d;A,kwargs...) = m .= A'*d
@everywhere function JopBaz(A)
    dom = JetSpace(eltype(A), size(A,2))
    rng = JetSpace(eltype(A), size(A,1))
    JopLn(;df! = JopBaz_df!, df′! = JopBaz_df′!, dom = dom, rng = rng, s = (A=A,))
end

const HERE = @__DIR__
inputs = JSON.parsefile(joinpath(HERE, "operator_cases_inputs.json"))
eltypeof(s) = s == "float32" ? Float32 : Float64

function setblocks!(x, flat, lens)
    o = 0
    for (ib, n) in enumerate(lens)
        setblock!(x, ib, flat[o+1:o+n])
        o += n
    end
    x
end

out = Dict{String,Any}()
for (name, case) in inputs
    T = eltypeof(case["dtype"])
    rs, cs, grid, isdiag = Int.(case["row_sizes"]), Int.(case["col_sizes"]), Int.(case["grid"]), case["isdiag"]
    N, M = length(rs), length(cs)
    # payloads travel to the workers inside the closure, like any DArray initialiser of the reference's tests
    payload = Dict{Tuple{Int,Int},Any}()
    for (key, blk) in case["blocks"]
        i, j = parse.(Int, split(key, "_"))
        if blk["kind"] == "dense"
            A = reshape(T.(blk["data"]), blk["shape"][1], blk["shape"][2])      # column-major, as exported
            payload[(i,j)] = (:dense, A)
        else
            payload[(i,j)] = (:diag, T.(blk["data"]))
        end
    end
    function blk(i, j)
        if haskey(payload, (i,j))
            kind, data = payload[(i,j)]
            return kind == :dense ? JopBaz(data) : JopFoo(data)
        end
        JopZeroBlock(JetSpace(T, cs[j]), JetSpace(T, rs[i]))
    end
    np = grid[1]*grid[2]
    _F = DArray(I->[blk(i,j) for i in I[1], j in I[2]], (N,M), workers()[1:np], grid)
    F = isdiag ? (@blockop _F isdiag=true) : (@blockop _F)
    x = setblocks!(zeros(domain(F)), T.(case["x"]), isdiag ? cs[1:N] : cs)
    y = setblocks!(zeros(range(F)), T.(case["y"]), rs)
    Ax = mul!(zeros(range(F)), F, x)
    Aty = mul!(zeros(domain(F)), F', y)
    out[name] = Dict(
        "Ax" => Float64.(convert(Array, Ax)), "Aty" => Float64.(convert(Array, Aty)),
        "dot_x_x" => Float64(dot(x, x)), "norm_y" => Float64(norm(y)), "norm1_y" => Float64(norm(y, 1)),
        "normInf_y" => Float64(norm(y, Inf)),
        "blockmap" => [[collect(first(c)), collect(last(c))] for c in vec(blockmap(F))],
        "range_blockmap" => collect.(blockmap(range(F))), "nblocks" => collect(nblocks(F)))
end
open(joinpath(HERE, "operator_cases_julia.json"), "w") do io
    JSON.print(io, Dict("julia" => string(VERSION), "cases" => out))
end
println("wrote ", joinpath(HERE, "operator_cases_julia.json"))
rmprocs(workers())
