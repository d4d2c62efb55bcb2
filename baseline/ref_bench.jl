# Times the UNMODIFIED reference (ChevronETC/DistributedJets.jl) on the host cores for the BASELINE configs,
# with the reference's own public API: addprocs(W) workers, `@blockop` over a DArray of Jets JopLn blocks,
# `mul!` forward and adjoint on DBArrays.  Prints ONE JSON line (the same metric bench.py reports).
#
#     julia baseline/ref_bench.jl --config 2 --steps 20 --warmup 5 --workers 8
#
# bench.py --impl reference runs this file when a `julia` with the reference's packages (Jets ^1.2,
# DistributedArrays 0.6, DistributedOperations 1, JSON 0.21; /root/reference/Project.toml:14-19) is on PATH and
# falls back to the NumPy/OpenBLAS restatement otherwise.  Julia is absent from the image this backend was built in,
# so this file has not been executed there; tests/test_julia_recipes.py checks every DistributedJets / Jets name it
# uses against the reference's sources.
using Distributed

function argval(name, default)
    k = findfirst(==(name), ARGS)
    k === nothing ? default : parse(Int, ARGS[k+1])
end
const CONFIG = argval("--config", 2)
const STEPS = argval("--steps", 20)
const WARMUP = argval("--warmup", 5)
const W = argval("--workers", Sys.CPU_THREADS)

addprocs(W)
@everywhere using DistributedArrays, DistributedJets, Jets, LinearAlgebra
using JSON

# the reference's dense fixture (test/runtests.jl:30-36), typed by the payload
@everywhere JopBaz_df!(d,m;A,kwargs...) = d .= A*m
@everywhere JopBaz_df′!(m,d;A,kwargs...) = m .= A'*d
@everywhere function JopBaz(A)
    dom = JetSpace(eltype(A), size(A,2))
    rng = JetSpace(eltype(A), size(A,1))
    JopLn(;df! = JopBaz_df!, df′! = JopBaz_df′!, dom = dom, rng = rng, s = (A=A,))
end

function build(config)
    if config == 1        # 4x4 dense Float64 1000x1000 blocks, grid [2,2] on 4 workers (test/benchmark shape)
        T, nb, m, n, grid, isdiag = Float64, 4, 1000, 1000, [2,2], false
    elseif config == 2    # block-diagonal, 64 dense Float32 4096x4096 blocks, grid [P,1]
        T, nb, m, n, grid, isdiag = Float32, 64, 4096, 4096, [min(W,64),1], true
    elseif config == 3    # full 32x32 dense Float64 2048x2048 blocks, grid [2,P/2]
        T, nb, m, n, grid, isdiag = Float64, 32, 2048, 2048, [2,max(1,min(W,32)÷2)], false
    else
        error("--config 1, 2 or 3")
    end
    np = grid[1]*grid[2]
    np <= nworkers() || error("config $config needs $np workers")
    blk = isdiag ?
        ((i,j)->(i == j ? JopBaz(rand(T,m,n)) : JopZeroBlock(JetSpace(T,n), JetSpace(T,m)))) :
        ((i,j)->JopBaz(rand(T,m,n)))
    _F = DArray(I->[blk(i,j) for i in I[1], j in I[2]], (nb,nb), workers()[1:np], grid)
    F = isdiag ? (@blockop _F isdiag=true) : (@blockop _F)
    nblocks_nz = isdiag ? nb : nb*nb
    bytes = nblocks_nz*m*n*sizeof(T) + (nb*m + nb*n)*sizeof(T)
    F, bytes, (T=T, nb=nb, m=m, n=n, grid=grid, isdiag=isdiag)
end

F, bytes_per_apply, shape = build(CONFIG)
m = rand(domain(F))
d = zeros(range(F))
m2 = zeros(domain(F))
step!() = (mul!(d, F, m); mul!(m2, F', d))
for _ in 1:WARMUP
    step!()
end
t = @elapsed for _ in 1:STEPS
    step!()
end
ms_per_step = 1e3*t/STEPS
println(JSON.json(Dict(
    "impl" => "reference", "kind" => "julia", "metric" => "JopDBlock mul!+adjoint effective HBM GB/s",
    "value" => 2*bytes_per_apply/(ms_per_step*1e-3)/1e9, "unit" => "GB/s", "ms_per_step" => ms_per_step,
    "steps" => STEPS, "warmup" => WARMUP, "workers" => nworkers(), "host_cores" => Sys.CPU_THREADS,
    "config" => CONFIG, "eltype" => string(shape.T), "blocks" => shape.nb, "block_shape" => [shape.m, shape.n],
    "grid" => shape.grid, "isdiag" => shape.isdiag, "julia" => string(VERSION))))
rmprocs(workers())
