# DistributedJetsB200.jl -- the Julia side of the B200 backend: a thin `ccall` layer over
# libdjets_b200.so (include/djets.h) that keeps the public names of ChevronETC/DistributedJets.jl
# for the hot path (reference: src/DistributedJets.jl, cited as src:LINE):
#
#   DBArray, JetDSpace, blockmap, localblockindices              (the reference's exports, src:863)
#   getblock / getblock! / setblock!, collect, convert(Array,.)  (src:310-330, 383-412)
#   dot, norm, extrema, mapreduce, mean, var, fill!, broadcasting (src:189-376)
#   @blockop over a matrix of native blocks -> a Jets JopLn whose df!/df′! run on the GPUs
#                                                                 (src:414-721)
#
# One GPU stands in for one former Distributed.jl worker ("rank"); the Julia process is the master
# and drives every rank through one context.  Where the reference ships closures to workers, this
# layer ships block *payloads* once (dense / diagonal / zero) and handles afterwards.
#
# NOTE: Julia is not installed in the image this backend was developed in, so this file has not
# been executed there; tests/ exercise the identical C ABI through the ctypes mirror
# (distributedjets.jl_b200/_lib.py, core.py), which is kept 1:1 with this file.
module DistributedJetsB200

using Jets, LinearAlgebra, Statistics, JSON
import Libdl

export DBArray, JetDSpace, blockmap, localblockindices

# a constant: `ccall((:symbol, lib), ...)` needs the library as a compile-time constant
const libdjets = get(ENV, "DJETS_B200_LIB", joinpath(@__DIR__, "..", "libdjets_b200.so"))

const DJETS_F32, DJETS_F64, DJETS_C64, DJETS_C128 = Int32(0), Int32(1), Int32(2), Int32(3)
const BLOCK_ZERO, BLOCK_DENSE, BLOCK_DIAG = Int32(0), Int32(1), Int32(2)
const RED_SUM, RED_MIN, RED_MAX, RED_ABSSUM, RED_ABSMAX, RED_ABSMIN, RED_SUMSQ, RED_NNZ, RED_SUMPOW, RED_SUMSQ_SHIFT = Int32.(0:9)

# operation codes of djets_vec_eval (DJETS_EV_*) and djets_scalar_op (DJETS_SC_*)
const EV_IN, EV_SC, EV_ADD, EV_SUB, EV_MUL, EV_DIV, EV_POW, EV_MIN, EV_MAX, EV_NEG, EV_ABS, EV_SQRT, EV_SQUARE, EV_CUBE, EV_EXP, EV_LOG,
      EV_ADDS, EV_SUBS, EV_RSUBS, EV_MULS, EV_RMULS, EV_DIVS, EV_RDIVS, EV_POWS, EV_MINS, EV_MAXS = UInt8.(1:26)
const SC_ADD, SC_SUB, SC_MUL, SC_DIV, SC_NEG, SC_SQRT, SC_COPY, SC_ABS, SC_MAX, SC_MIN = Int32.(0:9)

dtypecode(::Type{Float32}) = DJETS_F32
dtypecode(::Type{Float64}) = DJETS_F64
dtypecode(::Type{Complex{Float32}}) = DJETS_C64
dtypecode(::Type{Complex{Float64}}) = DJETS_C128
dtypecode(::Type{T}) where {T} = error("DistributedJetsB200: element type $T is not supported (Float32, Float64 and their Complex types)")

lasterror() = unsafe_string(ccall((:djets_last_error, libdjets), Cstring, ()))
function check(status::Int32)
    status == 0 && return nothing
    msg = lasterror()
    status == 4 && throw(DimensionMismatch(msg))
    error("libdjets_b200 (status $status): $msg")
end
macro dj(f, argtypes, args...)
    quote
        check(ccall(($(QuoteNode(f)), libdjets), Int32, $(esc(argtypes)), $(map(esc, args)...)))
    end
end

# ---- context: the workers ----------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    world::Int
end
const CTX = Ref{Union{Nothing,Context}}(nothing)

"""
    init(nranks=ngpus; devices=0:nranks-1)
`addprocs` for GPUs: rank `r` (0-based, the stand-in for a worker pid) runs on `devices[r+1]`.
"""
function init(nranks::Integer=devicecount(); devices=collect(0:nranks-1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ranks = Int32.(0:nranks-1)
    @dj djets_ctx_create (Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}, Ptr{Ptr{Cvoid}}) Int32(nranks) Int32(nranks) ranks Int32.(devices) C_NULL h
    CTX[] = Context(h[], nranks)
end
function devicecount()
    n = Ref{Int32}(0)
    @dj djets_device_count (Ptr{Int32},) n
    Int(n[])
end
context() = CTX[] === nothing ? error("DistributedJetsB200.init() first") : CTX[]
workers() = collect(0:context().world-1)
synchronize() = @dj djets_ctx_sync (Ptr{Cvoid},) context().h

function evencuts(n::Integer, nparts::Integer)
    cuts = Vector{Int64}(undef, nparts + 1)
    @dj djets_even_cuts (Int64, Int32, Ptr{Int64}) Int64(n) Int32(nparts) cuts
    cuts                                   # 0-based starts; block range k is cuts[k]+1:cuts[k+1]
end

# DistributedArrays.defaultdist(dims, pids): prime factors of the worker count, largest first, each handed to the
# currently largest remaining dimension (ties go to the last dimension)
function defaultgrid(dims, npids::Integer)
    facs, n, p = Int[], Int(npids), 2
    while p * p <= n
        while n % p == 0; push!(facs, p); n ÷= p; end
        p += 1
    end
    n > 1 && push!(facs, n)
    sort!(facs, rev=true)
    rem, chunks = collect(Int, dims), ones(Int, length(dims))
    for fac in facs
        k = findlast(==(maximum(rem)), rem)
        if rem[k] >= fac
            rem[k] ÷= fac; chunks[k] *= fac
        end
    end
    chunks
end

# ---- JetDSpace (src:48-107) ------------------------------------------------------------------------
struct JetDSpace{T} <: Jets.JetAbstractSpace{T,1}
    blkshapes::Vector{Vector{Dims}}        # one JetBSpace worth of block shapes per rank
    pids::Vector{Int}
    blkindices::Vector{UnitRange{Int}}
    indices::Vector{UnitRange{Int}}
end
function JetDSpace(::Type{T}, blkshapes::Vector{Vector{Dims}}, pids) where {T}
    blkindices, indices = UnitRange{Int}[], UnitRange{Int}[]
    b0, i0 = 1, 1
    for part in blkshapes                                                    # src:55-64
        nb, ne = length(part), sum(prod, part; init=0)
        push!(blkindices, b0:b0+nb-1); push!(indices, i0:i0+ne-1)
        b0 += nb; i0 += ne
    end
    JetDSpace{T}(blkshapes, collect(Int, pids), blkindices, indices)
end
Base.:(==)(R::JetDSpace, S::JetDSpace) = R.blkshapes == S.blkshapes && R.blkindices == S.blkindices && R.indices == S.indices
Base.similar(::JetDSpace{T}, dims::NTuple{1,Int}) where {T} = JetSpace(T, dims)                       # src:71-72
Base.similar(R::JetDSpace, dims::Int...) = similar(R, dims)
Base.size(R::JetDSpace) = (R.indices[end][end],)
Base.eltype(::JetDSpace{T}) where {T} = T
Base.eltype(::Type{JetDSpace{T}}) where {T} = T                                                         # src:83-84
Base.vec(R::JetDSpace) = R                                                                               # src:85
Jets.space(R::JetDSpace{T}, ipart::Integer) where {T} = JetBSpace([JetSpace(T, s) for s in R.blkshapes[ipart]])  # src:88 (indexes by worker: reference quirk)
Jets.indices(R::JetDSpace) = R.indices
Jets.nblocks(R::JetDSpace) = R.blkindices[end][end]
blockmap(R::JetDSpace) = R.blkindices
procs(R::JetDSpace) = R.pids
nprocs(R::JetDSpace) = length(R.pids)
localindices(R::JetDSpace, pid::Integer) = R.indices[findfirst(==(pid), R.pids)]        # src:91-95 (pid explicit: one master)
localblockindices(R::JetDSpace, pid::Integer) = R.blkindices[findfirst(==(pid), R.pids)]
blockshapes(R::JetDSpace) = reduce(vcat, R.blkshapes)

# ---- DBArray (src:109-412) ----------------------------------------------------------------------------
mutable struct DBArray{T} <: AbstractArray{T,1}
    h::Ptr{Cvoid}
    space::JetDSpace{T}
    function DBArray{T}(h, space) where {T}
        x = new{T}(h, space)
        finalizer(x -> (x.h == C_NULL || ccall((:djets_vec_destroy, libdjets), Int32, (Ptr{Cvoid},), x.h); x.h = C_NULL), x)
    end
end
function Base.Array(R::JetDSpace{T}) where {T}                                             # src:379-381
    h = Ref{Ptr{Cvoid}}(C_NULL)
    lens = Int64[prod(s) for s in blockshapes(R)]
    @dj djets_vec_create (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Cvoid}}) context().h dtypecode(T) Int32(length(R.pids)) Int32.(R.pids) Int64.(length.(R.blkshapes)) lens h
    DBArray{T}(h[], R)
end
Base.zeros(R::JetDSpace) = Array(R)                                                        # storage is zero-initialised
Base.ones(R::JetDSpace{T}) where {T} = fill!(Array(R), one(T))
function Base.rand(R::JetDSpace{T}) where {T}
    x = Array(R)
    for (i, s) in enumerate(blockshapes(R)); setblock!(x, i, rand(T, s)); end
    x
end
"""
    DBArray(f, blkidxs::Vector{UnitRange{Int}}, pids=workers())     (src:135-151)
    DBArray(f, (nblks,), pids[, dist])                               (src:153-159)
"""
function DBArray(f::Function, blkidxs::Vector{UnitRange{Int}}, pids=workers())
    blocks = [[f(i) for i in rng] for rng in blkidxs]
    T = eltype(first(first(blocks)))
    x = Array(JetDSpace(T, [Dims[size(b) for b in part] for part in blocks], pids[1:length(blkidxs)]))
    for (rng, part) in zip(blkidxs, blocks), (i, b) in zip(rng, part); setblock!(x, i, b); end
    x
end
function DBArray(f::Function, nblks::Tuple, pids::AbstractArray=workers()[1:min(length(workers()), nblks[1])], dist::AbstractArray=defaultgrid(nblks, length(pids)))   # src:153-159
    cuts = evencuts(nblks[1], dist[1])
    DBArray(f, [cuts[k]+1:cuts[k+1] for k in 1:dist[1]], pids)
end

Jets.space(x::DBArray) = x.space
Base.IndexStyle(::Type{<:DBArray}) = IndexLinear()
Base.size(x::DBArray) = size(x.space)
Jets.indices(x::DBArray) = x.space.indices
Jets.nblocks(x::DBArray) = nblocks(x.space)
blockmap(x::DBArray) = x.space.blkindices
procs(x::DBArray) = x.space.pids
nprocs(x::DBArray) = length(x.space.pids)
localindices(x::DBArray, pid::Integer) = localindices(x.space, pid)
localblockindices(x::DBArray, pid::Integer) = localblockindices(x.space, pid)

function Base.getindex(x::DBArray{T}, i::Int) where {T}                                     # src:174-179
    checkbounds(x, i)
    v = Ref{Float64}(0)
    @dj djets_vec_getindex (Ptr{Cvoid}, Int64, Ptr{Float64}) x.h Int64(i - 1) v
    T(v[])
end
function Base.getindex(x::DBArray{T}, i::Int) where {T<:Complex}
    checkbounds(x, i)
    z = zeros(Float64, 2)
    @dj djets_vec_getindex_c (Ptr{Cvoid}, Int64, Ptr{Float64}) x.h Int64(i - 1) z
    T(z[1], z[2])
end
function Base.similar(x::DBArray, ::Type{T}) where {T}                                      # src:181-185
    h = Ref{Ptr{Cvoid}}(C_NULL)
    @dj djets_vec_similar (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}) x.h dtypecode(T) h
    R = x.space
    DBArray{T}(h[], JetDSpace{T}(R.blkshapes, R.pids, R.blkindices, R.indices))
end
Base.similar(x::DBArray{T}) where {T} = similar(x, T)
Base.similar(x::DBArray, ::Type{T}, n::Integer) where {T} = length(x) == n ? similar(x, T) : Array{T}(undef, n)   # src:186
Base.similar(x::DBArray, ::Type{T}, dims::Tuple{Int}) where {T} = similar(x, T, dims[1])
function Base.reshape(x::DBArray, R::JetDSpace)                                             # src:342-345
    length(x) == length(R) || error("dimension mismatch, unable to reshap distributed block arrays")
    x
end

# block access (src:383-412) ---------------------------------------------------------------------------
function Jets.getblock(x::DBArray{T}, iblock::Integer) where {T}
    b = Array{T}(undef, blockshapes(x.space)[iblock])
    GC.@preserve b @dj djets_vec_getblock (Ptr{Cvoid}, Int64, Ptr{Cvoid}) x.h Int64(iblock - 1) pointer(b)
    b
end
Jets.getblock!(x::DBArray, iblock::Integer, xblock::AbstractArray) = xblock .= getblock(x, iblock)
function Jets.setblock!(x::DBArray{T}, iblock::Integer, xblock) where {T}
    shp = blockshapes(x.space)[iblock]
    b = xblock isa Number ? fill(T(xblock), shp) : convert(Array{T}, xblock)
    length(b) == prod(shp) || throw(DimensionMismatch("block $iblock has shape $shp, got $(size(b))"))
    GC.@preserve b @dj djets_vec_setblock (Ptr{Cvoid}, Int64, Ptr{Cvoid}) x.h Int64(iblock - 1) pointer(b)
    nothing
end

# collect / convert / fill! (src:310-339) -----------------------------------------------------------------
function Base.convert(::Type{Array}, x::DBArray{T}) where {T}
    a = Vector{T}(undef, length(x))
    GC.@preserve a @dj djets_vec_collect (Ptr{Cvoid}, Ptr{Cvoid}) x.h pointer(a)
    a
end
function Base.collect(x::DBArray{T}) where {T}
    a = convert(Array, x)
    shapes = blockshapes(x.space)
    arrays, idx, o = Array{T}[], UnitRange{Int}[], 0
    for s in shapes
        n = prod(s)
        push!(arrays, reshape(a[o+1:o+n], s)); push!(idx, o+1:o+n); o += n
    end
    Jets.BlockArray(arrays, idx)
end
Base.fill!(x::DBArray, a::Real) = (@dj djets_vec_fill (Ptr{Cvoid}, Float64) x.h Float64(a); x)
Base.fill!(x::DBArray, a::Complex) = (@dj djets_vec_fill_c (Ptr{Cvoid}, Float64, Float64) x.h Float64(real(a)) Float64(imag(a)); x)
# stream-ordered forms for page-locked buffers: `a` must stay alive and untouched until `synchronize()` / `waitmark`
scatter_async!(x::DBArray{T}, a::Vector{T}) where {T} = (@dj djets_vec_scatter_async (Ptr{Cvoid}, Ptr{Cvoid}) x.h pointer(a); x)
collect_async!(a::Vector{T}, x::DBArray{T}) where {T} = (@dj djets_vec_collect_async (Ptr{Cvoid}, Ptr{Cvoid}) x.h pointer(a); a)
mark(slot::Integer=0) = @dj djets_ctx_mark (Ptr{Cvoid}, Int32) context().h Int32(slot)
waitmark(slot::Integer=0) = @dj djets_ctx_wait_mark (Ptr{Cvoid}, Int32) context().h Int32(slot)

# ---- device-resident scalars: dot / norm results that never visit the host, and arithmetic on them -----------------
mutable struct DScalar{T}
    h::Ptr{Cvoid}
    function DScalar{T}() where {T}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        @dj djets_scalar_create (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) context().h h
        s = new{T}(h[])
        finalizer(s -> (s.h == C_NULL || ccall((:djets_scalar_destroy, libdjets), Int32, (Ptr{Cvoid},), s.h); s.h = C_NULL), s)
    end
end
DScalar{T}(v::Real) where {T} = (s = DScalar{T}(); @dj djets_scalar_set (Ptr{Cvoid}, Float64) s.h Float64(T(v)); s)
function Base.getindex(s::DScalar{T}) where {T}                       # s[] synchronises
    v = Ref{Float64}(0)
    @dj djets_scalar_get (Ptr{Cvoid}, Ptr{Float64}) s.h v
    T(v[])
end
function scalarop!(out::DScalar{T}, op::Int32, a::DScalar{T}, b::Union{Nothing,DScalar{T}}=nothing) where {T}
    @dj djets_scalar_op (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int32) out.h op a.h (b === nothing ? C_NULL : b.h) dtypecode(T)
    out
end
Base.:+(a::DScalar{T}, b::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_ADD, a, b)
Base.:-(a::DScalar{T}, b::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_SUB, a, b)
Base.:*(a::DScalar{T}, b::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_MUL, a, b)
Base.:/(a::DScalar{T}, b::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_DIV, a, b)
Base.:-(a::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_NEG, a)
Base.sqrt(a::DScalar{T}) where {T} = scalarop!(DScalar{T}(), SC_SQRT, a)
Base.broadcastable(s::DScalar) = Ref(s)                              # `x .+ alpha .* p` with a device scalar
dot!(out::DScalar{T}, x::DBArray{T}, y::DBArray{T}) where {T} = (@dj djets_vec_dot_s (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}) x.h y.h out.h; out)
norm!(out::DScalar, x::DBArray, p::Real=2) = (@dj djets_vec_norm_s (Ptr{Cvoid}, Float64, Ptr{Cvoid}) x.h Float64(p) out.h; out)

# captured sequences: `g = capture() do ... end; launch(g)` replays the stream-ordered calls made in the block
mutable struct Graph; h::Ptr{Cvoid}; end
function capture(f::Function)
    @dj djets_ctx_capture_begin (Ptr{Cvoid},) context().h
    h = Ref{Ptr{Cvoid}}(C_NULL)
    try
        f()
    finally
        status = ccall((:djets_ctx_capture_end, libdjets), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), context().h, h)
        check(status)
    end
    g = Graph(h[])
    finalizer(g -> (g.h == C_NULL || ccall((:djets_graph_destroy, libdjets), Int32, (Ptr{Cvoid},), g.h); g.h = C_NULL), g)
end
launch(g::Graph) = @dj djets_graph_launch (Ptr{Cvoid},) g.h

# reductions (src:189-283) ------------------------------------------------------------------------------------
function reduce_kind(x::DBArray{T}, kind::Int32, param=0.0) where {T}
    v = Ref{Float64}(0)
    @dj djets_vec_reduce (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}) x.h kind Float64(param) v
    T(v[])
end
function LinearAlgebra.norm(x::DBArray{T}, p::Real=2) where {T}
    v = Ref{Float64}(0)
    @dj djets_vec_norm (Ptr{Cvoid}, Float64, Ptr{Float64}) x.h Float64(p) v
    float(real(T))(v[])
end
function LinearAlgebra.dot(x::DBArray{T}, y::DBArray{T}) where {T}
    v = Ref{Float64}(0)
    @dj djets_vec_dot (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}) x.h y.h v
    T(v[])
end
function LinearAlgebra.dot(x::DBArray{T}, y::DBArray{T}) where {T<:Complex}                  # conjugates x, like Julia's dot
    z = zeros(Float64, 2)
    @dj djets_vec_dot_c (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}) x.h y.h z
    T(z[1], z[2])
end
Base.extrema(x::DBArray) = (reduce_kind(x, RED_MIN), reduce_kind(x, RED_MAX))
Base.maximum(x::DBArray) = reduce_kind(x, RED_MAX)
Base.minimum(x::DBArray) = reduce_kind(x, RED_MIN)
Base.maximum(::typeof(abs), x::DBArray) = reduce_kind(x, RED_ABSMAX)
Base.minimum(::typeof(abs), x::DBArray) = reduce_kind(x, RED_ABSMIN)
Base.sum(x::DBArray) = reduce_kind(x, RED_SUM)
Base.mapreduce(::typeof(abs), ::typeof(+), x::DBArray) = reduce_kind(x, RED_ABSSUM)
Base.mapreduce(::typeof(abs), ::typeof(max), x::DBArray) = reduce_kind(x, RED_ABSMAX)
Base.mapreduce(::typeof(abs), ::typeof(min), x::DBArray) = reduce_kind(x, RED_ABSMIN)
Base.mapreduce(::typeof(identity), ::typeof(+), x::DBArray) = reduce_kind(x, RED_SUM)
Base.mapreduce(f, op, x::DBArray) = error("DistributedJetsB200: mapreduce($f, $op, ::DBArray) is outside the closed set (abs|identity) x (+|max|min)")
Statistics.mean(x::DBArray) = sum(x) / length(x)
function Statistics.var(x::DBArray{T}; corrected=true, mean=nothing) where {T}
    μ = mean === nothing ? Statistics.mean(x) : mean
    σ² = reduce_kind(x, RED_SUMSQ_SHIFT, μ)
    corrected ? σ² / (length(x) - 1) : σ² / length(x)
end

# broadcasting (src:348-376).  The reference ships the Broadcasted tree to the workers and evaluates it over the local
# blocks (src:363-375); here the tree is lowered to the closed elementwise program of djets_vec_eval -- operators
# + - * / ^ abs sqrt exp log min max, unary minus, scalars (host numbers or device DScalars) -- and ONE streaming kernel
# per GPU evaluates it.  Anything else (an arbitrary Julia closure) is refused.
struct DBArrayStyle <: Broadcast.AbstractArrayStyle{1} end
Base.BroadcastStyle(::Type{<:DBArray}) = DBArrayStyle()
DBArrayStyle(::Val{1}) = DBArrayStyle()
find_dbarray(bc::Broadcast.Broadcasted) = find_dbarray(bc.args)
find_dbarray(args::Tuple) = find_dbarray(find_dbarray(args[1]), Base.tail(args))
find_dbarray(x) = x
find_dbarray(a::DBArray, rest) = a
find_dbarray(::Any, rest) = find_dbarray(rest)
Base.similar(bc::Broadcast.Broadcasted{DBArrayStyle}, ::Type{T}) where {T} = similar(find_dbarray(bc))      # src:353-356 (ignores T)

mutable struct Program
    code::Vector{UInt8}                    # (op, arg) pairs, postfix
    vecs::Vector{DBArray}
    scal::Vector{Float64}
    dscal::Vector{Ptr{Cvoid}}              # C_NULL: host scalar
    wide::Bool                             # a Float64 scalar or array takes part: evaluate in Float64 (Julia's promotion)
end
Program() = Program(UInt8[], DBArray[], Float64[], Ptr{Cvoid}[], false)
isscalar(x) = x isa Number || x isa DScalar || (x isa Base.RefValue && (x[] isa Number || x[] isa DScalar))
unref(x) = x isa Base.RefValue ? x[] : x
function vecindex!(p::Program, x::DBArray)
    k = findfirst(v -> v === x, p.vecs)
    k === nothing && (push!(p.vecs, x); k = length(p.vecs))
    eltype(x) === Float64 && (p.wide = true)
    UInt8(k - 1)
end
function scalarindex!(p::Program, a)
    a = unref(a)
    if a isa DScalar
        push!(p.scal, 0.0); push!(p.dscal, a.h)
        a isa DScalar{Float64} && (p.wide = true)
    else
        a isa Complex && !iszero(imag(a)) && error("DistributedJetsB200: complex coefficients in a broadcast are outside the closed set")
        push!(p.scal, Float64(real(a))); push!(p.dscal, C_NULL)
        (a isa Float64 || a isa Complex{Float64}) && (p.wide = true)                              # Int and Float32 scalars do not promote Float32 arrays
    end
    UInt8(length(p.scal) - 1)
end
emit!(p::Program, op::UInt8, arg::UInt8=0x00) = (push!(p.code, op); push!(p.code, arg); p)
const BIN = Dict{Any,Tuple{UInt8,UInt8,Union{UInt8,Nothing}}}(     # f => (stack form, vector op scalar, scalar op vector)
    (+) => (EV_ADD, EV_ADDS, EV_ADDS), (-) => (EV_SUB, EV_SUBS, EV_RSUBS), (*) => (EV_MUL, EV_MULS, EV_RMULS),
    (/) => (EV_DIV, EV_DIVS, EV_RDIVS), (^) => (EV_POW, EV_POWS, nothing), min => (EV_MIN, EV_MINS, EV_MINS),
    max => (EV_MAX, EV_MAXS, EV_MAXS))
const UN = Dict{Any,UInt8}((-) => EV_NEG, abs => EV_ABS, sqrt => EV_SQRT, exp => EV_EXP, log => EV_LOG)
refuse(f) = error("DistributedJetsB200: broadcast of $f over a DBArray is outside the closed operator set (+ - * / ^ abs sqrt exp log min max)")

lower!(p::Program, x::DBArray) = emit!(p, EV_IN, vecindex!(p, x))
lower!(p::Program, x) = isscalar(x) ? emit!(p, EV_SC, scalarindex!(p, x)) : refuse(typeof(x))
function lower!(p::Program, bc::Broadcast.Broadcasted)
    f, a = bc.f, bc.args
    if f === identity && length(a) == 1
        return lower!(p, a[1])
    elseif f === Base.literal_pow && length(a) == 3                   # x .^ 2 lowers to literal_pow(^, x, Val(2)): x*x, like Julia
        pw = typeof(unref(a[3])).parameters[1]
        lower!(p, a[2])
        return pw == 2 ? emit!(p, EV_SQUARE) : pw == 3 ? emit!(p, EV_CUBE) : emit!(p, EV_POWS, scalarindex!(p, Float64(pw)))
    elseif length(a) == 1 && haskey(UN, f)
        lower!(p, a[1])
        return emit!(p, UN[f])
    elseif length(a) >= 2 && haskey(BIN, f)
        (length(a) == 2 || f === (+) || f === (*)) || refuse(f)
        stack, vs, sv = BIN[f]
        # n-ary + and * fold from the left, as Base's +(a, b, c...) does
        if isscalar(a[1]) && !isscalar(a[2]) && sv !== nothing
            lower!(p, a[2]); emit!(p, sv, scalarindex!(p, a[1]))
        elseif !isscalar(a[1]) && isscalar(a[2])
            lower!(p, a[1]); emit!(p, vs, scalarindex!(p, a[2]))
        else
            lower!(p, a[1]); lower!(p, a[2]); emit!(p, stack)
        end
        for k in 3:length(a)
            isscalar(a[k]) ? emit!(p, vs, scalarindex!(p, a[k])) : (lower!(p, a[k]); emit!(p, stack))
        end
        return p
    end
    refuse(f)
end

function Base.copyto!(dest::DBArray, bc::Broadcast.Broadcasted{DBArrayStyle})                 # src:370-375
    if bc.f === identity && length(bc.args) == 1 && unref(bc.args[1]) isa Number               # dest .= a  (test:329)
        return fill!(dest, unref(bc.args[1]))
    end
    p = lower!(Program(), bc)
    length(p.code) <= 96 || error("DistributedJetsB200: broadcast expression longer than 48 operations; split it")
    GC.@preserve p @dj djets_vec_eval (Ptr{Cvoid}, Ptr{UInt8}, Int32, Ptr{Ptr{Cvoid}}, Int32, Ptr{Float64}, Ptr{Ptr{Cvoid}}, Int32, Int32) dest.h p.code Int32(length(p.code) ÷ 2) Ptr{Cvoid}[v.h for v in p.vecs] Int32(length(p.vecs)) p.scal p.dscal Int32(length(p.scal)) Int32(p.wide && eltype(dest) === Float32 ? 1 : 0)
    dest
end
function Base.copyto!(dest::DBArray, src::DBArray)                                            # dest .= x, also converts (test:237-238)
    coef, hs = Float64[1.0], Ptr{Cvoid}[src.h]
    @dj djets_vec_lincomb (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Ptr{Cvoid}}, Int32) dest.h Int32(1) coef hs Int32(0)
    dest
end

# ---- native block kinds (SURVEY F5: what replaces arbitrary JopLn closures) -------------------------------
abstract type B200Block{T} end
struct DenseBlock{T} <: B200Block{T}; A::Matrix{T}; end              # d .= A*m ; m .= A'*d   (JopBaz, test:30-36)
struct DiagBlock{T,N} <: B200Block{T}; d::Array{T,N}; end            # d .= diag .* m         (JopFoo, test:8-12)
struct ZeroBlock{T} <: B200Block{T}; rng::Dims; dom::Dims; end       # JopZeroBlock
rngshape(b::DenseBlock) = (size(b.A, 1),); domshape(b::DenseBlock) = (size(b.A, 2),)
rngshape(b::DiagBlock) = size(b.d);        domshape(b::DiagBlock) = size(b.d)
rngshape(b::ZeroBlock) = b.rng;            domshape(b::ZeroBlock) = b.dom
Base.iszero(b::B200Block) = b isa ZeroBlock

mutable struct DBlockOp{T}
    h::Ptr{Cvoid}
    pids::Matrix{Int}
    blockmap::Matrix{Tuple{UnitRange{Int},UnitRange{Int}}}
    isdiag::Bool
    mrep::Any                              # DBArray home of a replicated (plain-array) model, made on demand
    function DBlockOp{T}(h, pids, bm, isdiag) where {T}
        op = new{T}(h, pids, bm, isdiag, nothing)
        finalizer(o -> (o.h == C_NULL || ccall((:djets_op_destroy, libdjets), Int32, (Ptr{Cvoid},), o.h); o.h = C_NULL), op)
    end
end

function JetDBlock_df!(d::DBArray, m::DBArray; op, perfstatfile="", kwargs...)                                      # src:663-674
    @dj djets_op_mul (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}) op.h d.h m.h
    write_perfstat(op, "df", perfstatfile)
    d
end
function JetDBlock_df′!(m::DBArray, d::DBArray; op, perfstatfile="", kwargs...)                                     # src:710-721
    @dj djets_op_mul_adjoint (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}) op.h m.h d.h
    write_perfstat(op, "df′", perfstatfile)
    m
end

# perfstatfile (src:723-745): one step per apply, one entry per worker; "localblock" holds the worker's
# elapsed milliseconds, one value per local block (the JSON layout the reference's test:618-651 pins)
function write_perfstat(op, operation, perfstatfile)   # not `perfstat`: Jets exports that name (the reference extends Jets.perfstat, src:730)
    perfstatfile == "" && return nothing
    entries = Dict{String,Any}[]
    for c in 1:size(op.pids, 2), r in 1:size(op.pids, 1)
        ms = Ref{Float64}(0)
        @dj djets_op_perfstat (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}) op.h Int32(r - 1) Int32(c - 1) ms
        nloc = length(op.blockmap[r, c][1]) * length(op.blockmap[r, c][2])
        push!(entries, Dict("id" => op.pids[r, c], "hostname" => gethostname(), "operation" => operation,
                            "localblock" => fill(ms[] / nloc, nloc)))
    end
    doc = isfile(perfstatfile) ? JSON.parse(read(perfstatfile, String)) : Dict("step" => Any[])
    push!(doc["step"], Dict("pid" => entries))
    write(perfstatfile, JSON.json(doc, 1))
    nothing
end

# replicated domain (tall and skinny, grid n x 1): `m` is a plain array on the master, broadcast forward
# (src:544-557) and tree-summed in the adjoint (src:565-580).  Its device-side home is the single
# domain part on grid cell (1,1); phase 0 of the exchange plan forwards it to the other grid rows.
function replicated!(op, dom::JetDSpace)
    op.mrep === nothing && (op.mrep = zeros(dom))
    op.mrep
end
function JetDBlock_df!(d::DBArray, m::Array{T}; op, dom, kwargs...) where {T}
    x = replicated!(op, dom)
    GC.@preserve m @dj djets_vec_scatter (Ptr{Cvoid}, Ptr{Cvoid}) x.h pointer(m)
    JetDBlock_df!(d, x; op=op)
end
function JetDBlock_df′!(m::Array{T}, d::DBArray; op, dom, kwargs...) where {T}
    x = replicated!(op, dom)
    JetDBlock_df′!(x, d; op=op)
    GC.@preserve m @dj djets_vec_collect (Ptr{Cvoid}, Ptr{Cvoid}) x.h pointer(m)
    m
end

"""
    JopBlock(blocks::Matrix{<:B200Block}; pids=workers(), dist=[n1,n2], isdiag=false)
The `@blockop DArray(...) isdiag=...` of the reference (src:414-415, 476-508) for native blocks: the
blocks are laid out on an `n1 x n2` grid of ranks (column-major, src:483), payloads are uploaded to
the owning GPU, and the result is a Jets `JopLn` whose `df!`/`df′!` are `djets_op_mul`/`_adjoint`.
"""
function Jets.JopBlock(blocks::Matrix{<:B200Block{T}}; pids=workers(), dist=nothing, isdiag=false, perfstatfile="") where {T}
    N, M = size(blocks)
    dist === nothing && (dist = [min(N, length(pids)), 1])
    n1, n2 = dist
    (isdiag && n2 != 1) && error("block diagonal jets must define distribution along rows.")                      # src:481
    grid = reshape(collect(Int, pids[1:n1*n2]), n1, n2)
    rcuts, ccuts = evencuts(N, n1), evencuts(M, n2)
    rowshapes = Dims[rngshape(blocks[i, 1]) for i in 1:N]
    colshapes = isdiag ? Dims[domshape(blocks[i, 1]) for i in 1:N] : Dims[domshape(blocks[1, j]) for j in 1:M]     # src:471-473, 493
    if isdiag
        for i in 1:N
            s = count(!iszero, @view blocks[i, :])
            s == 1 || error("Expecting one non-zero column block per row in block diagonal construction, got $s.")   # src:469
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    @dj djets_op_create (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Ptr{Cvoid}}) context().h dtypecode(T) Int64(N) Int64(M) Int64.(prod.(rowshapes)) Int64.(prod.(colshapes)) Int32(n1) Int32(n2) Int32.(vec(grid)) rcuts ccuts Int32(isdiag) h
    for i in 1:N, j in (isdiag ? (i:i) : (1:M))
        b = blocks[i, j]
        if b isa DenseBlock
            GC.@preserve b @dj djets_op_set_dense (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64) h[] Int64(i - 1) Int64(j - 1) pointer(b.A) Int64(max(size(b.A, 1), 1))
        elseif b isa DiagBlock
            GC.@preserve b @dj djets_op_set_diag (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}) h[] Int64(i - 1) Int64(j - 1) pointer(b.d)
        end
    end
    @dj djets_op_finalize (Ptr{Cvoid},) h[]
    if perfstatfile != ""                                                                                          # src:505
        rm(perfstatfile, force=true)
        @dj djets_op_perfstat_enable (Ptr{Cvoid}, Int32) h[] Int32(1)
    end
    bm = [(rcuts[r]+1:rcuts[r+1], ccuts[c]+1:ccuts[c+1]) for r in 1:n1, c in 1:n2]                                 # src:507 indices(ops)
    rng = JetDSpace(T, [rowshapes[rcuts[r]+1:rcuts[r+1]] for r in 1:n1], grid[:, 1])                               # src:485
    dom = isdiag ? JetDSpace(T, [colshapes[rcuts[r]+1:rcuts[r+1]] for r in 1:n1], grid[:, 1]) :
                   JetDSpace(T, [colshapes[ccuts[c]+1:ccuts[c+1]] for c in 1:n2], grid[1, :])                       # src:491-493
    op = DBlockOp{T}(h[], grid, bm, isdiag)
    # like the reference (src:459-462) a single domain worker means a plain local space: with n2 == 1 the
    # operator also accepts `m::Array`; the DBArray form of the domain stays available as `state(A).dom`.
    JopLn(dom = dom, rng = rng, df! = JetDBlock_df!, df′! = JetDBlock_df′!, s = (op=op, dom=dom, blockmap=bm, isdiag=isdiag, perfstatfile=perfstatfile))
end

# accessors (src:805-852) ------------------------------------------------------------------------------------------
const JetDBlockB200 = Jet{D,R,typeof(JetDBlock_df!)} where {D,R}
Jets.isblockop(::JetDBlockB200) = true
procs(jet::JetDBlockB200) = state(jet).op.pids
nprocs(jet::JetDBlockB200) = length(procs(jet))
blockmap(jet::JetDBlockB200) = state(jet).blockmap
procs(A::Jop) = procs(jet(A)); nprocs(A::Jop) = nprocs(jet(A)); blockmap(A::Jop) = blockmap(jet(A))
function localblockindices(jet::JetDBlockB200, pid::Integer)
    k = findfirst(==(pid), procs(jet))
    k === nothing ? (0:0, 0:0) : blockmap(jet)[k]
end
localblockindices(A::Jop, pid::Integer) = localblockindices(jet(A), pid)
localblockindices(A::Jop, pid::Integer, i::Integer) = localblockindices(A, pid)[i]
function Jets.getblock(jet::JetDBlockB200, i::Integer, j::Integer)
    op = state(jet).op
    T = eltype(domain(jet))
    kind = Ref{Int32}(0)
    @dj djets_op_getblock (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Cvoid}, Int64) op.h Int64(i - 1) Int64(j - 1) kind C_NULL Int64(0)
    rshp = blockshapes(range(jet))[i]
    dshp = blockshapes(domain(jet))[op.isdiag ? i : j]
    if kind[] == BLOCK_DENSE
        A = Matrix{T}(undef, prod(rshp), prod(dshp))
        GC.@preserve A @dj djets_op_getblock (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Cvoid}, Int64) op.h Int64(i - 1) Int64(j - 1) kind pointer(A) Int64(max(size(A, 1), 1))
        return DenseBlock{T}(A)
    elseif kind[] == BLOCK_DIAG
        d = Array{T}(undef, rshp)
        GC.@preserve d @dj djets_op_getblock (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}, Ptr{Cvoid}, Int64) op.h Int64(i - 1) Int64(j - 1) kind pointer(d) Int64(0)
        return DiagBlock{T,length(rshp)}(d)
    end
    ZeroBlock{T}(rshp, dshp)
end
Base.close(jet::JetDBlockB200) = finalize(state(jet).op)                                                             # src:854-861

end # module
