/*
 * djets.h -- C ABI of libdjets_b200.so, the B200 (sm_100a) backend for the hot path of
 * ChevronETC/DistributedJets.jl: a distributed block operator (`@blockop` over a DArray of Jets
 * JopLn blocks) applied to distributed block vectors (DBArray), forward and adjoint, plus the
 * DBArray algebra around it.
 *
 * The reference has NO foreign-function interface of its own (it is one pure-Julia file,
 * src/DistributedJets.jl, cited below as `src:LINE`); each entry point names the reference
 * method(s) it replaces.  A Julia shim binds these with `ccall` (see INTEGRATION.md and
 * distributedjets.jl_b200/julia/DistributedJetsB200.jl); tests bind the same symbols with ctypes.
 *
 * Conventions
 *  - Every function returns an int32 status (0 = DJETS_OK).  `djets_last_error()` returns a
 *    thread-local message for the last failure.  No C++ exception or abort crosses this ABI.
 *  - Handles are opaque.  Host pointers are borrowed for the duration of the call only.
 *  - All indices are 0-based; ranges are (first, count).  The Julia/Python shims convert from
 *    the reference's 1-based inclusive ranges.
 *  - Dense block payloads are column-major with an explicit leading dimension (Julia layout).
 *  - A *rank* stands in for one Distributed.jl worker and owns one GPU.  A context either owns
 *    all ranks of the job (one host process driving P GPUs: the Julia-master model) or a subset
 *    (one process per GPU under torchrun/SPMD: every process makes the same sequence of calls).
 *  - Calls are stream-ordered per rank and return before completion, except those that return
 *    host data (getblock, getindex, collect, dot, norm, reduce, timer_stop), which synchronise.
 *  - Results are run-to-run deterministic: no floating-point atomics anywhere.
 */
#ifndef DJETS_H
#define DJETS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct djets_ctx_s* djets_ctx;
typedef struct djets_vec_s* djets_vec;
typedef struct djets_op_s* djets_op;
typedef struct djets_scalar_s* djets_scalar; /* a scalar resident on the devices (result of dot / norm, solver coefficients) */
typedef struct djets_graph_s* djets_graph;   /* a captured sequence of calls, replayed with one launch */

/* element types; the complex ones are interleaved (re, im) pairs: Julia's Complex{Float32} / Complex{Float64} */
enum { DJETS_F32 = 0, DJETS_F64 = 1, DJETS_C64 = 2, DJETS_C128 = 3 };
enum { DJETS_BLOCK_ZERO = 0, DJETS_BLOCK_DENSE = 1, DJETS_BLOCK_DIAG = 2 };
enum {
  DJETS_OK = 0,
  DJETS_ERR_INVALID = 1,     /* bad argument / handle */
  DJETS_ERR_CUDA = 2,        /* CUDA runtime failure (no device, launch error, OOM ...) */
  DJETS_ERR_NCCL = 3,        /* NCCL missing or failed */
  DJETS_ERR_MISMATCH = 4,    /* dimension / layout / dtype mismatch (reference: DimensionMismatch) */
  DJETS_ERR_UNSUPPORTED = 5, /* legal in the reference but outside this backend's closed set */
  DJETS_ERR_NOT_LOCAL = 6    /* the block is owned by a rank this process does not drive */
};
/* kinds for djets_vec_reduce */
enum {
  DJETS_RED_SUM = 0,        /* sum(x)                 src:252 (mean) */
  DJETS_RED_MIN = 1,        /* minimum(x)             src:225-237 */
  DJETS_RED_MAX = 2,        /* maximum(x)             src:225-237 */
  DJETS_RED_ABSSUM = 3,     /* mapreduce(abs,+,x)     src:239-250, norm(x,1) */
  DJETS_RED_ABSMAX = 4,     /* maximum(abs,x), norm(x,Inf) */
  DJETS_RED_ABSMIN = 5,     /* minimum(abs,x), norm(x,-Inf) */
  DJETS_RED_SUMSQ = 6,      /* sum(abs2,x)            norm(x,2) partial */
  DJETS_RED_NNZ = 7,        /* count(!iszero,x)       norm(x,0) */
  DJETS_RED_SUMPOW = 8,     /* sum(abs(x)^param)      norm(x,p) partial */
  DJETS_RED_SUMSQ_SHIFT = 9 /* sum((x-param)^2)       src:265 (var) */
};
/* Operation codes of the closed elementwise program of djets_vec_eval: a postfix program of (op, arg) byte pairs.
 * IN / SC push input vector `arg` / scalar `arg`; binary operators pop b, pop a and push `a op b`; unary operators
 * replace the top; the *S forms combine the top with scalar `arg` without a push (x .+ a, a .* x, max.(x, a) ...;
 * the R* forms put the scalar on the left: a .- x, a ./ x). */
enum {
  DJETS_EV_IN = 1, DJETS_EV_SC = 2,
  DJETS_EV_ADD = 3, DJETS_EV_SUB = 4, DJETS_EV_MUL = 5, DJETS_EV_DIV = 6, DJETS_EV_POW = 7, DJETS_EV_MIN = 8, DJETS_EV_MAX = 9,
  DJETS_EV_NEG = 10, DJETS_EV_ABS = 11, DJETS_EV_SQRT = 12, DJETS_EV_SQUARE = 13, DJETS_EV_CUBE = 14, DJETS_EV_EXP = 15,
  DJETS_EV_LOG = 16,
  DJETS_EV_ADDS = 17, DJETS_EV_SUBS = 18, DJETS_EV_RSUBS = 19, DJETS_EV_MULS = 20, DJETS_EV_RMULS = 21, DJETS_EV_DIVS = 22,
  DJETS_EV_RDIVS = 23, DJETS_EV_POWS = 24, DJETS_EV_MINS = 25, DJETS_EV_MAXS = 26
};
/* operations of djets_scalar_op */
enum {
  DJETS_SC_ADD = 0, DJETS_SC_SUB = 1, DJETS_SC_MUL = 2, DJETS_SC_DIV = 3, DJETS_SC_NEG = 4, DJETS_SC_SQRT = 5, DJETS_SC_COPY = 6,
  DJETS_SC_ABS = 7, DJETS_SC_MAX = 8, DJETS_SC_MIN = 9
};
/* kernel kinds for djets_ctx_kernel_stats */
enum {
  DJETS_K_GEMV_N = 0,  /* dense tiles, forward  */
  DJETS_K_GEMV_T = 1,  /* dense tiles, adjoint  */
  DJETS_K_FINISH = 2,  /* segmented block-row / block-column sum + diagonal blocks */
  DJETS_K_REDUCE = 3,  /* dot / norm / reductions */
  DJETS_K_ELTWISE = 4, /* fill / linear combination / convert / eval */
  DJETS_K_COUNT = 5
};

/* ---- status ------------------------------------------------------------------------------ */
int32_t djets_version(void);
const char* djets_last_error(void);

/* ---- planning: pure host arithmetic, callable without a GPU ------------------------------- */
/* DistributedArrays.defaultdist(sz, nc) (even split; the first `sz % nparts` parts get one
 * extra) as 0-based starts: cuts[0]=0 ... cuts[nparts]=n.  Replaces chunk_idxs/defaultdist as
 * used at src:154,158-159,483. */
int32_t djets_even_cuts(int64_t n, int32_t nparts, int64_t* cuts);
/* Rank of grid cell (r,c): pids reshaped column-major to n1 x n2 (src:12, src:483). */
int32_t djets_grid_rank(int32_t n1, int32_t n2, int32_t r, int32_t c, const int32_t* grid_ranks, int32_t* rank);
/* Number of dense tiles the forward / adjoint schedule would use for an m x n block (for tests
 * of the scheduler without a device). */
int32_t djets_plan_tiles(int32_t dtype, int64_t m, int64_t n, int32_t adjoint, int64_t* tile_rows,
                         int64_t* tile_cols, int64_t* ntiles);

/* Chained tiles (small blocks): how many dense blocks of one block row (forward) / block column (adjoint) one CTA
 * multiplies, by the rule the planner applies to a grid cell (pure function; no device).  `nblocks` dense blocks in the
 * cell, at most `longest` of them feeding one output block, the largest `max_rows` x `max_cols` and `max_bytes` bytes;
 * `wave` = resident CTAs of the chain kernel on the device; `chain` = the context parameter (0: sized by the rule, 1: off,
 * k: forced).  Out: blocks per chain (1 = no chains) and whether the adjoint may load two blocks per batch (every block
 * at most one warp-load tall). */
int32_t djets_plan_chains(int32_t dtype, int32_t adjoint, int64_t nblocks, int64_t longest, int64_t max_rows,
                          int64_t max_cols, int64_t max_bytes, int32_t wave, int32_t chain, int32_t* blocks_per_chain,
                          int32_t* pair);

/* The point-to-point messages of one apply on an n1 x n2 grid (0-based grid coordinates), six
 * int32 per message: src_r, src_c, dst_r, dst_c, part, slot.  phase 0: input parts from their owner
 * (grid row 0 for mul!, grid column 0 for the adjoint) to the other cells of the grid column / row
 * (replaces the remote getblock of src:588,635,682); phase 1: reduced partials to the owner of the
 * output part, receive slot `slot` (replaces ArrayFutures + reduce! of src:600-604,647-651,694-698).
 * out6 may be NULL to query the count.  This is the list djets_op_mul / djets_op_mul_adjoint execute. */
int32_t djets_plan_exchange(int32_t n1, int32_t n2, int32_t isdiag, int32_t adjoint, int32_t phase, int32_t* out6,
                            int32_t capacity, int32_t* count);

/* Validates an elementwise program for djets_vec_eval without a device: stack depth it needs, and the number of
 * operations left after the peephole pass that folds pushed operands into their consumers. */
int32_t djets_plan_eval(const uint8_t* program, int32_t nops, int32_t ninputs, int32_t nscalars, int32_t* depth,
                        int32_t* nops_fused);

/* ---- context ------------------------------------------------------------------------------ */
int32_t djets_device_count(int32_t* n);
/* 128-byte NCCL unique id, to be generated on one process and handed to every process'
 * djets_ctx_create (multi-process only). */
int32_t djets_nccl_unique_id(uint8_t* out128);
/* world: number of ranks in the job.  nlocal/local_ranks/device_ids: the ranks this process
 * drives and their CUDA devices.  nccl_uid: NULL when nlocal == world (generated internally when
 * a cross-rank exchange is first needed) or when no cross-rank exchange will be used. */
int32_t djets_ctx_create(int32_t world, int32_t nlocal, const int32_t* local_ranks, const int32_t* device_ids,
                         const uint8_t* nccl_uid, djets_ctx* out);
int32_t djets_ctx_destroy(djets_ctx ctx);
/* Host-side all-gather among the processes of the job, provided by the caller (Julia: Distributed; Python:
 * torch.distributed / MPI): every process passes one record of `bytes` bytes; `recv` has room for `world` records and
 * the callback stores the record of every process of the job there, contiguously from the start, in any order (each
 * record names its sender).  Returns 0 on success.  With it a multi-process job needs no NCCL at all: it carries the
 * control plane (exchange of CUDA IPC handles when an operator is finalised, the P scalar partials of dot / norm),
 * while block data moves between GPUs through peer memory.  Every process must make the same sequence of calls. */
typedef int32_t (*djets_allgather_fn)(void* user, const void* send, void* recv, int64_t bytes);
int32_t djets_ctx_set_host_allgather(djets_ctx ctx, djets_allgather_fn fn, void* user);
int32_t djets_ctx_sync(djets_ctx ctx);
/* nprocs() of the job, how many ranks this process drives, whether cross-process exchange is up. */
int32_t djets_ctx_info(djets_ctx ctx, int32_t* world, int32_t* nlocal, int32_t* has_nccl);
/* myid()-style query: does this process drive `rank`? (src:302-303 localindices/localblockindices) */
int32_t djets_ctx_is_local(djets_ctx ctx, int32_t rank, int32_t* is_local);
/* Page-locked host buffers for block payloads moved with setblock!/getblock/collect/scatter
 * (the reference moves them by Serialization over TCP; here they cross PCIe by DMA). */
int32_t djets_host_alloc(int64_t bytes, void** out);
int32_t djets_host_free(void* p);
/* CUDA-event timer on every local rank's stream; stop synchronises and returns the max over
 * local ranks in milliseconds. */
int32_t djets_timer_start(djets_ctx ctx);
int32_t djets_timer_stop(djets_ctx ctx, float* ms);
/* Per-kernel accounting.  Launch counts are always kept; with timing enabled every launch is
 * bracketed by CUDA events on its stream (small overhead: use a separate pass). */
int32_t djets_ctx_kernel_timing(djets_ctx ctx, int32_t enable);
int32_t djets_ctx_kernel_stats(djets_ctx ctx, int32_t kind, int64_t* launches, double* total_ms);
int32_t djets_ctx_reset_stats(djets_ctx ctx);
/* Tuning knob.  Schedule shapes (affect operators finalised or re-planned afterwards): "fwd_tx", "fwd_tc", "adj_ru",
 * "adj_tc", "ld_policy", "auto_tile", "adj_light", "chain", "wave_fit", "l2_keep" (MB of a cell's last tiles handed to the next
 * apply through L2; 0 = off).  Launch mechanisms (take effect on the next apply; applies captured under older settings
 * are dropped): "graphs", "p2p", "pdl", "prefetch", "persistent", "ipc" (before the operator is finalised), "xp_kwait",
 * "xp_side", "xp_push". */
int32_t djets_ctx_set_param(djets_ctx ctx, const char* name, int64_t value);
/* Stream marks: djets_ctx_mark records slot (0..7) on every local rank's stream; djets_ctx_wait_mark blocks the host
 * until everything enqueued before that mark has finished, without draining what was enqueued after it (double
 * buffering of host transfers: scatter_async / collect_async). */
int32_t djets_ctx_mark(djets_ctx ctx, int32_t slot);
int32_t djets_ctx_wait_mark(djets_ctx ctx, int32_t slot);
/* Captured sequences: every stream-ordered call between begin and end (mul!, adjoint, *_s reductions, scalar_op,
 * eval / lincomb / binary / fill) is recorded instead of executed; djets_graph_launch replays the whole sequence
 * with one launch (a solver iteration without host round trips).  Calls that return host data or allocate are
 * refused while a capture is open.  Needs every rank of the job in this process. */
int32_t djets_ctx_capture_begin(djets_ctx ctx);
int32_t djets_ctx_capture_end(djets_ctx ctx, djets_graph* out);
int32_t djets_graph_launch(djets_graph g);
int32_t djets_graph_destroy(djets_graph g);

/* ---- device-resident scalars ---------------------------------------------------------------------- */
/* The reference returns dot / norm to the master and the solver forms alpha = gamma / delta there (src:189-223);
 * here the value can stay on the devices (a replica per rank), feed djets_vec_eval as a coefficient and be combined
 * by djets_scalar_op, so a CG-type iteration never waits for the host. */
int32_t djets_scalar_create(djets_ctx ctx, djets_scalar* out);
int32_t djets_scalar_destroy(djets_scalar s);
int32_t djets_scalar_set(djets_scalar s, double v);
int32_t djets_scalar_get(djets_scalar s, double* out); /* synchronises */
/* out = a op b (DJETS_SC_*; b may be NULL for unary operations), rounded to the element type `dtype`. */
int32_t djets_scalar_op(djets_scalar out, int32_t op, djets_scalar a, djets_scalar b, int32_t dtype);

/* ---- DBArray: distributed block vector (src:109-412) ------------------------------------------ */
/* DBArray(f, blkidxs, pids) / zeros(R) (src:135-159, 378-381): `nparts` parts, part p owned by
 * rank part_rank[p] and holding part_nblocks[p] consecutive blocks; block_len lists every block
 * in global order.  Storage is zero-initialised. */
int32_t djets_vec_create(djets_ctx ctx, int32_t dtype, int32_t nparts, const int32_t* part_rank,
                         const int64_t* part_nblocks, const int64_t* block_len, djets_vec* out);
/* similar(x, T) (src:181-185). */
int32_t djets_vec_similar(djets_vec x, int32_t dtype, djets_vec* out);
int32_t djets_vec_destroy(djets_vec x);
/* size / nblocks / nprocs / eltype (src:168, 296-297). */
int32_t djets_vec_info(djets_vec x, int32_t* dtype, int64_t* length, int64_t* nblocks, int32_t* nparts);
/* procs / indices / blockmap per part (src:291, 302-304). */
int32_t djets_vec_part_info(djets_vec x, int32_t part, int32_t* rank, int64_t* blk_first, int64_t* blk_count,
                            int64_t* elem_first, int64_t* elem_count);
/* owner lookup of src:389-390: part, global element offset and length of block `iblock`. */
int32_t djets_vec_block_info(djets_vec x, int64_t iblock, int32_t* part, int64_t* elem_first, int64_t* len);
/* setblock!(x, i, b) (src:409-412); a no-op on processes that do not drive the owner. */
int32_t djets_vec_setblock(djets_vec x, int64_t iblock, const void* host_src);
/* getblock(x, i) / getblock!(x, i, b) (src:388-399): copies the block out. */
int32_t djets_vec_getblock(djets_vec x, int64_t iblock, void* host_dst);
/* x[i] (src:174-179). */
int32_t djets_vec_getindex(djets_vec x, int64_t i, double* out);
/* x[i] of a complex DBArray: out2 = {re, im}. */
int32_t djets_vec_getindex_c(djets_vec x, int64_t i, double* out2);
/* collect(x) / convert(Array, x) (src:310-330): parts concatenated in procs order.  Parts owned
 * by ranks of other processes are left untouched in host_dst. */
int32_t djets_vec_collect(djets_vec x, void* host_dst);
/* inverse of collect: every locally driven part is loaded from its slice of host_src (one H2D
 * per rank; the block-by-block form is setblock!, docs:211-227). */
int32_t djets_vec_scatter(djets_vec x, const void* host_src);
/* Asynchronous forms of scatter / collect for page-locked buffers (djets_host_alloc): they return at once and run
 * on the owner's copy stream -- after everything enqueued so far, overlapping whatever is enqueued afterwards that
 * does not touch x; the next call that touches x waits for the copy.  host_src must stay unmodified / host_dst is
 * valid only after a later synchronising call (djets_ctx_sync, djets_ctx_wait_mark of a mark made after the copy). */
int32_t djets_vec_scatter_async(djets_vec x, const void* host_src);
int32_t djets_vec_collect_async(djets_vec x, void* host_dst);
/* fill!(x, a) (src:332-339; also `x .= a`, test:329). */
int32_t djets_vec_fill(djets_vec x, double a);
/* fill!(x, re + im*i) for complex DBArrays. */
int32_t djets_vec_fill_c(djets_vec x, double re, double im);
/* Broadcast `dest .= c0*x0 .+ c1*x1 .+ ...` (src:348-376 for the expression class the
 * reference's tests and benchmark use: test:322, bench:43; also `dest .= x` with element-type
 * conversion, test:237-238).  1 <= nterms <= 4; dest may alias any x_k; all layouts must match.
 * flags bit 0: evaluate in Float64 even when every vector is Float32 (Julia semantics with
 * Float64 scalars). */
int32_t djets_vec_lincomb(djets_vec dest, int32_t nterms, const double* coef, const djets_vec* xs, int32_t flags);
/* Broadcast `dest .= a .* b` (op 0) or `dest .= a ./ b` (op 1). */
int32_t djets_vec_binary(djets_vec dest, int32_t op, djets_vec a, djets_vec b);
/* General broadcast `dest .= f(x_1 .. x_k, scalars)` (src:363-375: the reference evaluates any Broadcasted tree per
 * worker) for the closed operator set DJETS_EV_*: `program` holds nops (op, arg) byte pairs in postfix order,
 * at most 48 operations, 8 inputs, 16 scalars, 8 operands deep.  Scalar k is scalars[k] unless dscalars != NULL and
 * dscalars[k] != NULL (then the device-resident value is read by the kernel).  dest may alias any input; layouts
 * must match; element types may differ (conversion).  flags bit 0: evaluate in Float64 although every vector is
 * Float32 (Julia semantics with Float64 scalars).  One streaming kernel per part: every input read once. */
int32_t djets_vec_eval(djets_vec dest, const uint8_t* program, int32_t nops, const djets_vec* inputs, int32_t ninputs,
                       const double* scalars, const djets_scalar* dscalars, int32_t nscalars, int32_t flags);
/* dot / norm / reductions whose result stays on the devices (no host synchronisation): the per-rank partials are
 * folded by a one-thread kernel with the reference's rules (src:200-209, 222, 236) into `out`. */
int32_t djets_vec_dot_s(djets_vec x, djets_vec y, djets_scalar out);
int32_t djets_vec_norm_s(djets_vec x, double p, djets_scalar out);
int32_t djets_vec_reduce_s(djets_vec x, int32_t kind, double param, djets_scalar out);
/* dot(x, y) (src:212-223): per-rank partial, folded over ranks in rank order in the element type. */
int32_t djets_vec_dot(djets_vec x, djets_vec y, double* out);
/* dot(x, y) = sum conj(x) .* y of complex DBArrays: out2 = {re, im}.  norm(x, p) of a complex DBArray goes through
 * djets_vec_norm (|z| per element); lincomb / eval accept complex vectors with real coefficients (linear forms);
 * the other reductions and the non-linear broadcast operators are defined for real element types only. */
int32_t djets_vec_dot_c(djets_vec x, djets_vec y, double* out2);
/* norm(x, p) (src:189-210) with the reference's fold rules; p may be +-INFINITY, 0, 1, 2, other. */
int32_t djets_vec_norm(djets_vec x, double p, double* out);
/* Per-rank partial reductions folded over ranks (src:225-283): see DJETS_RED_*. */
int32_t djets_vec_reduce(djets_vec x, int32_t kind, double param, double* out);
/* Per-part partials of the last fold (length nparts), for callers that want the reference's
 * `z[ipid]` vector itself. */
int32_t djets_vec_reduce_parts(djets_vec x, int32_t kind, double param, double* out_parts);

/* ---- distributed block operator (src:414-721, 805-861) ---------------------------------------- */
/* JopBlock(::DArray) / JetBlock(ops; isdiag) (src:414-508).  The operator has nbrow x nbcol
 * blocks; block row i has row_len[i] elements, block column j has col_len[j].  The blocks are
 * distributed over an n1 x n2 grid of ranks (grid_ranks, column-major, src:483): grid cell (r,c)
 * owns block rows [row_cuts[r], row_cuts[r+1]) x block columns [col_cuts[c], col_cuts[c+1]).
 * Range vectors live on grid column 0, domain vectors on grid row 0 (src:485,493); with isdiag
 * (requires n2 == 1, src:481) both live on grid column 0 and only blocks (i,i) are applied
 * (src:609-614, 656-661, 703-708).  Blocks never set are zero blocks (JopZeroBlock). */
int32_t djets_op_create(djets_ctx ctx, int32_t dtype, int64_t nbrow, int64_t nbcol, const int64_t* row_len,
                        const int64_t* col_len, int32_t n1, int32_t n2, const int32_t* grid_ranks,
                        const int64_t* row_cuts, const int64_t* col_cuts, int32_t isdiag, djets_op* out);
/* Dense block A (row_len[i] x col_len[j], column-major, leading dimension ld): `d .= A*m`,
 * `m .= A'*d` (the reference's dense fixture JopBaz, test:30-36; `A'` is the conjugate transpose for complex operators).  No-op on processes that do
 * not drive the owner (host may be NULL there). */
int32_t djets_op_set_dense(djets_op op, int64_t i, int64_t j, const void* host, int64_t ld);
/* Diagonal block `d .= diag .* m` (JopFoo test:8-12, JetPack JopDiagonal docs:19); adjoint `m .= conj(diag) .* d`. */
int32_t djets_op_set_diag(djets_op op, int64_t i, int64_t j, const void* host);
/* Build the tile schedules (after the last set_*; implicit on first apply). */
int32_t djets_op_finalize(djets_op op);
/* Rebuild the tile schedules with the context's current tuning parameters (payloads stay put). */
int32_t djets_op_replan(djets_op op);
int32_t djets_op_destroy(djets_op op);
/* zeros(range(A)) / zeros(domain(A)) (src:378-381 over the spaces of src:502-503). which: 0 = domain, 1 = range. */
int32_t djets_op_new_vec(djets_op op, int32_t which, djets_vec* out);
/* mul!(y, A, x): JetDBlock_df!/f! (src:616-627, 663-674). */
int32_t djets_op_mul(djets_op op, djets_vec y, djets_vec x);
/* mul!(x, A', y): JetDBlock_df'! (src:710-721). */
int32_t djets_op_mul_adjoint(djets_op op, djets_vec x, djets_vec y);
/* blockmap(A)[r,c] (src:833): block-row and block-column range of grid cell (r,c). */
int32_t djets_op_blockmap(djets_op op, int32_t r, int32_t c, int64_t* row_first, int64_t* row_count,
                          int64_t* col_first, int64_t* col_count);
/* Grid cell and rank that own block (i,j): the search of src:838-847. */
int32_t djets_op_block_owner(djets_op op, int64_t i, int64_t j, int32_t* r, int32_t* c, int32_t* rank);
/* getblock(A, i, j) (src:836-852): kind, and (when host_dst != NULL) the payload copied out
 * column-major with leading dimension ld (dense) or contiguous (diag). */
int32_t djets_op_getblock(djets_op op, int64_t i, int64_t j, int32_t* kind, void* host_dst, int64_t ld);
/* Algorithmic bytes of one apply on the ranks this process drives: non-zero block payloads +
 * domain + range elements. */
int32_t djets_op_algorithmic_bytes(djets_op op, int64_t* bytes);
/* perfstatfile support (src:723-745: per-worker statistics gathered after every apply).  With it
 * enabled every grid cell's share of each apply is bracketed by CUDA events on its stream (graph replay
 * is then bypassed); djets_op_perfstat returns the elapsed milliseconds of cell (r,c) in the last apply. */
int32_t djets_op_perfstat_enable(djets_op op, int32_t enable);
int32_t djets_op_perfstat(djets_op op, int32_t r, int32_t c, double* ms);
/* Size of the finalised schedule on the ranks this process drives: dense tiles (one CTA each) and
 * segmented-sum items left to the separate finish kernel (the others are fused into the GEMV kernel). */
int32_t djets_op_schedule_info(djets_op op, int32_t adjoint, int64_t* ntiles, int64_t* nitems);

#ifdef __cplusplus
}
#endif
#endif /* DJETS_H */
