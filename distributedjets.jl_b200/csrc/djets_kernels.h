// Device-side work descriptors and launcher prototypes shared by the host scheduler
// (djets_ctx.cu, djets_vec.cu, djets_op.cu) and the sm_100a kernels (djets_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace djets {

// One dense tile of one block: a rows x cols window of a column-major block.
//   forward (gemv_n):  partial[r] = sum_c A[r,c] * in[vin + c]      r < rows
//   adjoint (gemv_t):  partial[c] = sum_r A[r,c] * in[vin + r]      c < cols
// The partial goes to workspace plane `wout` or, when this tile is the only contribution to its
// output strip (`direct`), straight into the output vector at element offset `wout`.
struct DenseTile {
  const void* A;      // tile origin (16-byte aligned: row0 % VEC == 0, ld % VEC == 0)
  long long ld;       // leading dimension in elements (padded to a 16-byte multiple, pad rows are 0)
  long long vin;      // element offset of the tile's slice of the input vector
  long long wout;     // element offset into the workspace (or the output vector when direct)
  int rows, cols;     // tile extents
  int direct;
  int fin;            // >= 0: finish item this tile's strip belongs to (fused segmented sum), else -1
  int keep;           // 1: one of the last tiles of the launch -- loaded with an L2 evict-last hint, because the next
                      // apply of the operator (the other direction) starts with exactly these tiles
  int chain;          // adjoint tiles over small blocks: number of further blocks (ChainLink entries from `link0` on)
  int link0;          // the same CTA multiplies and accumulates into this tile's partial; 0 everywhere else
  int pad_;
};

// A further block of a chained tile.  Adjoint: same columns as the head tile, its own rows and input slice,
//   partial[c] = sum_r A_head[r,c] in[vin_head + r] + sum_links sum_r A_l[r,c] in[vin_l + r];
// forward: same rows as the head tile, its own columns and input slice,
//   partial[r] = sum_c A_head[r,c] in[vin_head + c] + sum_links sum_c A_l[r,c] in[vin_l + c].
// Small blocks (100x100 Float64 = 80 KB) leave a CTA more time in its fixed costs (descriptor, staging, ticket)
// than in its loads; a chain of K blocks of one block column / row pays them once and writes one partial instead of K.
struct ChainLink {
  const void* A;
  long long ld;
  long long vin;
  int n;  // rows of the block (adjoint) / columns of the block (forward)
  int pad_;
};
constexpr int kChainMax = 16;  // blocks per chained tile, head included

// One term of a diagonal block inside a finish item: out[e] += d[e] * in[vin + e].
struct DiagTerm {
  const void* d;      // diagonal, already offset to the item's first row
  long long vin;      // element offset into the input vector, already offset likewise
};

// Segmented sum of one output chunk: out[out+e] = sum_p W[w0 + p*wstride + e]
//                                               + sum_q R[r0 + q*rstride + e]     (planes received from peers)
//                                               + sum_t d_t[e] * in[vin_t + e]    (diagonal blocks)
// An item is either "loose" (run by djets_segsum_kernel) or "fused": the last of the `nplanes` tile CTAs
// that contribute to the strip (counted with an atomic ticket) performs the sum inside the GEMV
// kernel, always in plane order, so the result does not depend on which CTA came last.
struct FinishItem {
  long long out;
  long long w0, wstride;
  long long r0, rstride;
  int nplanes, nremote;
  int diag_begin, diag_end;
  int len;
  int flags;  // bit 0: complex (elements are (re, im) pairs, len counts reals), bit 1: conjugate the diagonal (adjoint)
};

// ---- launchers (all asynchronous on `stream`) ------------------------------------------------
// dtype: 0 = float, 1 = double.  fwd_tx in {32,64,128,256}; adj_ru in {2,4,8}.
struct FusedFinish {
  const FinishItem* items;
  const DiagTerm* diags;
  unsigned* tickets;  // one per item, zero between launches
};
// L2 prefetch ahead of the dependency wait: the first `ctas` CTAs of a launch each pull `bytes` of their
// tile (beyond what their first batch of loads covers) towards L2 while the predecessor kernel drains.
struct Prefetch {
  int ctas;
  int bytes;
  int pdl;  // launch attribute: programmatic dependent launch on / off (a context parameter)
  unsigned long long* trace;  // debugging (DJETS_TRACE): 4 globaltimer stamps per CTA of this launch, or nullptr
  // Cross-process exchange: epoch flags (in this GPU's memory, raised by peers) every CTA waits for before it
  // touches the input slice or stores a partial -- the kernel is resident and its L2 prefetch under way while
  // the input is still arriving over NVLink.  nwait == 0: nothing to wait for.
  unsigned* wait_flag[2];
  unsigned wait_value[2];
  int nwait;
  unsigned* wait_err;
};
// `grid` CTAs (<= ntiles) walk the tile table: CTA b takes tile b, further tiles are drawn from `counter`
// (one unsigned, zero between launches; rearmed by the launch itself).  grid == ntiles: one tile per CTA.
cudaError_t launch_gemv_n(int dtype, int fwd_tx, int ld_policy, const DenseTile* tiles, int ntiles, int grid,
                          unsigned* counter, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                          cudaStream_t stream);
cudaError_t launch_gemv_t(int dtype, int adj_ru, int adj_cw, int ld_policy, const DenseTile* tiles, int ntiles,
                          int grid, unsigned* counter, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                          cudaStream_t stream);
// Chained tiles (blocks of at most 64 * VEC rows), one tile per CTA.  Adjoint: `pair` = every block is at most
// 32 * VEC rows tall (one warp-load), two blocks are loaded per batch.  Forward: fwd_tx in {32, 64}, the columns of a
// chain add up to at most kFwdMaxTC.
cudaError_t launch_gemv_t_chain(int dtype, int pair, int ld_policy, const DenseTile* tiles, const ChainLink* links,
                                int ntiles, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                cudaStream_t stream);
cudaError_t launch_gemv_n_chain(int dtype, int fwd_tx, int ld_policy, const DenseTile* tiles, const ChainLink* links,
                                int ntiles, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                cudaStream_t stream);
int gemv_chain_resident_ctas(int dtype, int adjoint, int ld_policy);
inline int chain_max_rows(int dtype) { return 64 * (dtype == 0 ? 4 : 2); }
// CTAs of the tile kernel variant that fit on the current device at once (occupancy x SM count).
int gemv_resident_ctas(int dtype, int adjoint, int shape /* fwd_tx | adj_ru */, int adj_cw, int ld_policy);
cudaError_t launch_finish(int dtype, const FinishItem* items, int nitems, const DiagTerm* diags, const void* in,
                          const void* W, const void* R, void* out, cudaStream_t stream);

// Elementwise: dest[i] = sum_k coef[k] * x_k[i]; dtypes per operand; compute in double if wide.
struct LincombArgs {
  const void* x[4];
  double coef[4];
  int xdtype[4];
  int nterms;
};
cudaError_t launch_lincomb(int ddtype, void* dest, const LincombArgs& a, long long n, int wide, cudaStream_t stream);
cudaError_t launch_fill(int dtype, void* dest, double a, long long n, cudaStream_t stream);
cudaError_t launch_fill_complex(int dtype, void* dest, double re, double im, long long n, cudaStream_t stream);
cudaError_t launch_binary(int dtype, int op, void* dest, const void* a, const void* b, long long n,
                          cudaStream_t stream);

// Epoch flags of the cross-process exchange over peer memory (CUDA IPC): a producer stores its data into the
// consumer's buffer (NVLink peer stores / copies) and then raises the consumer's flag to the apply's epoch; the
// consumer's stream spins on its own (local) flag before the kernel that reads the data.  Comparisons are
// wrap-safe; a wait that is not satisfied within ~20 s raises *err instead of hanging the GPU.
struct FlagList {
  unsigned* p[16];
  unsigned v[16];
  int n;
};
// One kernel for the whole input push of the peer-memory exchange: wait for the consumers' *_consumed flags, copy
// `bytes` from src into every destination (peer stores: all NVLink destinations are written in parallel), then the
// last CTA raises the consumers' in_ready flags.  `ticket`: one unsigned, zero between launches.
struct PushList {
  char* dst[16];
  int n;
};
cudaError_t launch_xp_push(const char* src, size_t bytes, const PushList& dst, const FlagList& waits, const FlagList& signals,
                           unsigned* ticket, unsigned* err, cudaStream_t stream);
cudaError_t launch_flags_wait(const FlagList& f, unsigned* err, cudaStream_t stream);
cudaError_t launch_flags_signal(const FlagList& f, cudaStream_t stream);

// Closed elementwise program (djets_vec_eval): postfix byte-code, see DJETS_EV_* in include/djets.h.
enum {
  EV_IN = 1, EV_SC = 2,
  EV_ADD = 3, EV_SUB = 4, EV_MUL = 5, EV_DIV = 6, EV_POW = 7, EV_MIN = 8, EV_MAX = 9,           // pop b, pop a, push a op b
  EV_NEG = 10, EV_ABS = 11, EV_SQRT = 12, EV_SQUARE = 13, EV_CUBE = 14, EV_EXP = 15, EV_LOG = 16,  // top = f(top)
  EV_ADDS = 17, EV_SUBS = 18, EV_RSUBS = 19, EV_MULS = 20, EV_RMULS = 21, EV_DIVS = 22, EV_RDIVS = 23,
  EV_POWS = 24, EV_MINS = 25, EV_MAXS = 26,                                                        // top = top op scalar[arg]
  // internal (never cross the ABI): operand folded into the operator by the peephole pass; arg = input | scalar << 3
  EV_ADD_IN = 32, EV_SUB_IN = 33, EV_RSUB_IN = 34, EV_MUL_IN = 35, EV_DIV_IN = 36, EV_RDIV_IN = 37, EV_MIN_IN = 38,
  EV_MAX_IN = 39, EV_AXPY_IN = 40, EV_AXMY_IN = 41
};
constexpr int kEvalMaxOps = 48, kEvalMaxIn = 8, kEvalMaxScalars = 16, kEvalMaxDepth = 8;
struct EvalArgs {
  unsigned char op[kEvalMaxOps], arg[kEvalMaxOps];
  int nops;
  const void* in[kEvalMaxIn];
  int in_dtype[kEvalMaxIn];
  int nin;
  double sc[kEvalMaxScalars];          // host value of scalar k ...
  const double* dsc[kEvalMaxScalars];  // ... unless this is non-null: device-resident scalar read by the kernel
  int nsc;
};
// Validates a program; returns its stack depth (>= 1) or -1 with *why set.
int eval_check(const EvalArgs& a, const char** why);
// The program after the peephole pass (operands folded into their consumers); returns its length.
int eval_peephole(const EvalArgs& a, unsigned char* op, unsigned char* arg);
cudaError_t launch_eval(int ddtype, void* dest, const EvalArgs& a, long long n, int wide, cudaStream_t stream);

// Reductions in one launch.  `partials` holds at least reduce_max_ctas() doubles, `ticket` one unsigned that is
// zero between launches, `result` one double (device).  kind: DJETS_RED_* or RED_DOT (100).
// complex kinds: |z|-based maps over (re, im) pairs (n counts complex elements) and the imaginary part of dot(x, y) = sum conj(x) y
enum { RED_DOT = 100, RED_CABSSUM = 101, RED_CABSMAX = 102, RED_CABSMIN = 103, RED_CNNZ = 104, RED_CSUMPOW = 105, RED_CDOT_IM = 106 };
int reduce_max_ctas();
cudaError_t launch_reduce(int dtype, int kind, const void* x, const void* y, long long n, double param,
                          double* partials, unsigned* ticket, double* result, cudaStream_t stream);

// Master-side fold of per-part partials on the device (the reference folds them on the master, src:200-209,
// 222, 236, 262, 276), leaving the scalar in device memory: out[0] on this rank and, when `peers` is given, in
// the replicas other local ranks hold (peer stores).
enum { FOLD_SUM = 0, FOLD_MIN = 1, FOLD_MAX = 2, FOLD_NORM2 = 3, FOLD_NORMP = 4 };
struct ScalarPeers {
  double* p[16];
  int n;
};
cudaError_t launch_fold(int dtype, int mode, double p, const double* parts, int nparts, double* out, ScalarPeers peers,
                        cudaStream_t stream);
// Scalar arithmetic on device-resident scalars: out = a op b (b ignored by unary ops), rounded to `dtype`.
enum { SC_ADD = 0, SC_SUB = 1, SC_MUL = 2, SC_DIV = 3, SC_NEG = 4, SC_SQRT = 5, SC_COPY = 6, SC_ABS = 7, SC_MAX = 8, SC_MIN = 9 };
cudaError_t launch_scalar_op(int dtype, int op, const double* a, const double* b, double* out, cudaStream_t stream);
cudaError_t launch_scalar_set(double v, double* out, cudaStream_t stream);

// geometry helpers shared with the scheduler
// dtype codes: 0 Float32, 1 Float64, 2 Complex{Float32}, 3 Complex{Float64} (interleaved re, im).  Complex data is
// processed by the real-typed kernels over twice as many real elements; the GEMV / diagonal / |z| kernels take a
// CPLX switch for the places where the pairing matters.
inline bool is_complex(int dtype) { return dtype >= 2; }
inline int real_dtype(int dtype) { return dtype & 1; }
inline int elem_size(int dtype) { return dtype == 0 ? 4 : (dtype == 3 ? 16 : 8); }
inline int vec_elems(int dtype) { return 16 / elem_size(dtype); }
constexpr int kThreads = 256;
constexpr int kFwdMaxTC = 1024;   // columns of x staged in shared memory by a forward tile
constexpr int kAdjMaxTRBytes = 32768;  // bytes of y staged in shared memory by an adjoint tile
constexpr int kAdjCW = 4;         // fewest columns a warp handles at once in the adjoint kernel (tall columns)
constexpr int kFinishChunk = 4096;        // elements of one loose segmented-sum item (one CTA), at least ...
constexpr int kFinishChunkBytes = 32768;  // ... this many bytes per stream: Float32 items are 8192 elements
inline int adj_max_rows(int dtype) { return kAdjMaxTRBytes / elem_size(dtype); }

}  // namespace djets
