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

const DJETS_F32, DJETS_F64 = Int32(0), Int32(1)
const BLOCK_ZERO, BLOCK_DENSE, BLOCK_DIAG = Int32(0), Int32(1), Int32(2)
const RED_SUM, RED_MIN, RED_MAX, RED_ABSSUM, RED_ABSMAX, RED_ABSMIN, RED_SUMSQ, RED_NNZ, RED_SUMPOW, RED_SUMSQ_SHIFT = Int32.(0:9)

dtypecode(::Type{Float32}) = DJETS_F32
dtypecode(::Type{Float64}) = DJETS_F64
dtypecode(::Type{T}) where {T} = error("DistributedJetsB200: element type $T is not supported (Float32, Float64)")

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
function DBArray(f::Function, nblks::Tuple, pids::AbstractArray=workers()[1:min(length(workers()), nblks[1])], dist::AbstractArray=[min(length(pids), nblks[1])])
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
Base.fill!(x::DBArray, a) = (@dj djets_vec_fill (Ptr{Cvoid}, Float64) x.h Float64(a); x)

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

# broadcasting (src:348-376): the closed expression class  dest .= Σ cₖ .* xₖ ,  dest .= x ,  dest .= a ,
# dest .= x .* y ,  dest .= x ./ y ;  anything else is refused (a CUDA library cannot run a Julia closure).
struct DBArrayStyle <: Broadcast.AbstractArrayStyle{1} end
Base.BroadcastStyle(::Type{<:DBArray}) = DBArrayStyle()
DBArrayStyle(::Val{1}) = DBArrayStyle()
find_dbarray(bc::Broadcast.Broadcasted) = find_dbarray(bc.args)
find_dbarray(args::Tuple) = find_dbarray(find_dbarray(args[1]), Base.tail(args))
find_dbarray(x) = x
find_dbarray(a::DBArray, rest) = a
find_dbarray(::Any, rest) = find_dbarray(rest)
Base.similar(bc::Broadcast.Broadcasted{DBArrayStyle}, ::Type{T}) where {T} = similar(find_dbarray(bc))      # src:353-356 (ignores T)

# linear form of a broadcast tree: Vector of (coefficient, DBArray), or nothing
linterms(x::DBArray, c=1.0) = [(Float64(c), x)]
function linterms(bc::Broadcast.Broadcasted, c=1.0)
    f, a = bc.f, bc.args
    if f === (+)
        ts = Tuple{Float64,DBArray}[]
        for arg in a
            t = linterms(arg, c); t === nothing && return nothing; append!(ts, t)
        end
        return ts
    elseif f === (-) && length(a) == 2
        t1, t2 = linterms(a[1], c), linterms(a[2], -c)
        return (t1 === nothing || t2 === nothing) ? nothing : vcat(t1, t2)
    elseif f === (-) && length(a) == 1
        return linterms(a[1], -c)
    elseif f === (*) && length(a) == 2 && a[1] isa Number
        return linterms(a[2], c * a[1])
    elseif f === (*) && length(a) == 2 && a[2] isa Number
        return linterms(a[1], c * a[2])
    elseif f === (/) && length(a) == 2 && a[2] isa Number
        return linterms(a[1], c / a[2])
    elseif f === identity && length(a) == 1
        return linterms(a[1], c)
    end
    nothing
end
linterms(::Any, c=1.0) = nothing

function lincomb!(dest::DBArray, ts)
    length(ts) <= 4 || error("DistributedJetsB200: at most 4 terms per fused broadcast; split the expression")
    coef = Float64[t[1] for t in ts]
    hs = Ptr{Cvoid}[t[2].h for t in ts]
    GC.@preserve ts @dj djets_vec_lincomb (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Ptr{Cvoid}}, Int32) dest.h Int32(length(ts)) coef hs Int32(1)
    dest
end
function Base.copyto!(dest::DBArray, bc::Broadcast.Broadcasted{DBArrayStyle})                 # src:370-375
    ts = linterms(bc)
    ts === nothing || return lincomb!(dest, ts)
    if (bc.f === (*) || bc.f === (/)) && length(bc.args) == 2 && all(a -> a isa DBArray, bc.args)
        @dj djets_vec_binary (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}) dest.h Int32(bc.f === (*) ? 0 : 1) bc.args[1].h bc.args[2].h
        return dest
    end
    error("DistributedJetsB200: broadcast expression outside the closed set (linear combinations, .*, ./, fill, convert)")
end
Base.copyto!(dest::DBArray, src::DBArray) = lincomb!(dest, [(1.0, src)])

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
    perfstat(op, "df", perfstatfile)
    d
end
function JetDBlock_df′!(m::DBArray, d::DBArray; op, perfstatfile="", kwargs...)                                     # src:710-721
    @dj djets_op_mul_adjoint (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}) op.h m.h d.h
    perfstat(op, "df′", perfstatfile)
    m
end

# perfstatfile (src:723-745): one step per apply, one entry per worker; "localblock" holds the worker's
# elapsed milliseconds, one value per local block (the JSON layout the reference's test:618-651 pins)
function perfstat(op, operation, perfstatfile)
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
