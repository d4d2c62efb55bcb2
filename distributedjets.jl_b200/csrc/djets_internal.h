// Internal declarations shared by the host-side translation units of libdjets_b200
// (djets_ctx.cu: contexts, timers, tuning; djets_vec.cu: DBArray; djets_op.cu: the distributed block
// operator).  Nothing here is part of the ABI: include/djets.h is.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/djets.h"
#include "djets_kernels.h"
#include "djets_nccl.h"

namespace djets {

// ---- errors ---------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
// One lock for the whole ABI: the contract is one calling thread per context, but *_destroy may
// arrive from a finalizer thread (Julia GC) while the master thread is inside another call.
extern std::recursive_mutex g_api_mutex;

struct Err {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Err{code, msg}; }

#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      ::djets::fail(DJETS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define NCK(api, expr)                                                                    \
  do {                                                                                    \
    ncclResult_t r__ = (expr);                                                            \
    if (r__ != ncclSuccess) ::djets::fail(DJETS_ERR_NCCL, std::string(#expr) + ": " + (api)->GetErrorString(r__)); \
  } while (0)

#define REQUIRE(cond, code, msg)          \
  do {                                    \
    if (!(cond)) ::djets::fail(code, msg); \
  } while (0)

// Every extern "C" entry point runs inside guard(): no C++ exception crosses the ABI.
template <class F>
int32_t guard(F&& f) {
  std::lock_guard<std::recursive_mutex> api_lock(g_api_mutex);
  try {
    f();
    return DJETS_OK;
  } catch (const Err& e) {
    g_last_error = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = std::string("internal: ") + e.what();
    return DJETS_ERR_INVALID;
  } catch (...) {
    g_last_error = "internal: unknown exception";
    return DJETS_ERR_INVALID;
  }
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

constexpr int kMaxParts = 1024;
constexpr int kMaxScalars = 4096;
constexpr int kTraceLaunches = 16, kTraceCtas = 16384;

}  // namespace djets

// ---- context --------------------------------------------------------------------------------------
struct KernelEvent {
  int kind;
  cudaEvent_t a, b;
};

struct RankCtx {
  int rank = -1, device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;  // scratch event for cross-stream ordering
  cudaEvent_t ev_fork = nullptr;  // fork / join of a captured apply
  cudaStream_t cstream = nullptr;  // copy stream: scatter_async / collect_async overlap the kernels of the main stream
  cudaEvent_t cmarks[8] = {};      // djets_ctx_mark on the copy stream
  cudaEvent_t ev_c0 = nullptr;     // main -> copy stream ordering
  cudaStream_t xstream = nullptr;  // side stream: pushes of an input part to its consumers in other processes
  cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  double* d_partials = nullptr;  // per-CTA partials of the reduction in flight
  unsigned* d_ticket = nullptr;  // last-CTA ticket of the reduction kernel (zero between launches)
  double* d_result = nullptr;    // 2*kMaxParts doubles: per-part results, then the all-reduce buffer
  double* h_result = nullptr;    // pinned mirror, also mapped into the device's address space ...
  double* h_result_dev = nullptr;  // ... so a reduction kernel can store its result straight into host memory
  double* d_gather = nullptr;    // kMaxParts doubles: per-part partials gathered on the root rank of a device-side fold
  double* d_scalars = nullptr;   // kMaxScalars doubles: this rank's replica of every device-resident scalar
  unsigned* d_xp_err = nullptr;  // raised by a flag wait that timed out (cross-process peer exchange)
  unsigned* d_xp_ticket = nullptr;  // completion ticket of the push kernel
  cudaEvent_t marks[8] = {};     // djets_ctx_mark / djets_ctx_wait_mark
  ncclComm_t comm = nullptr;
  int64_t launches[DJETS_K_COUNT] = {0, 0, 0, 0, 0};
  int64_t launches_mark[DJETS_K_COUNT] = {0, 0, 0, 0, 0};  // counters at djets_ctx_capture_begin
  double ms[DJETS_K_COUNT] = {0, 0, 0, 0, 0};
  std::vector<KernelEvent> pending;
  std::vector<cudaEvent_t> pool;
  // Stream-ordered cache of vector-part allocations: a freed part is handed to the next DBArray of the
  // same size on this rank without a device synchronisation (the reference's non-dotted arithmetic
  // allocates a fresh DBArray per operation, src:353-356, so allocation has to be cheap).
  // DJETS_TRACE=<file prefix>: per-CTA time stamps of the first kTraceLaunches tile-kernel launches after
  // the last djets_ctx_reset_stats, written to <prefix>.rank<r>.bin when the context is destroyed
  unsigned long long* d_trace = nullptr;
  int trace_launches = 0;
  std::vector<int> trace_meta;  // per traced launch: kind, CTAs
  std::multimap<size_t, char*> vec_cache;
  size_t vec_cache_bytes = 0;
};

struct djets_ctx_s {
  int world = 0;
  std::vector<RankCtx> ranks;
  std::vector<int> local_of_rank;
  bool has_nccl = false;
  bool timing = false;
  int64_t fwd_tx = 64, fwd_tc = 512, adj_ru = 4, adj_tc = 64, ld_policy = 1;
  int64_t adj_light = 1;  // adjoint of blocks with very short columns: (RU,CW) = (1,8) variant at three CTAs per SM
  int64_t chain = 0;  // small blocks: blocks per chained tile (0 = sized by the planner, 1 = off, k = forced)
  int64_t l2_keep = 96;   // MB of a cell's last tiles loaded evict-last, for the next apply (which walks the tiles the other way round); 0: off
  int64_t wave_fit = 1;   // auto_tile: grow the tiles of a cell that would need between one and two resident waves into one wave
  int64_t auto_tile = 1;  // shrink tiles of small operators so a launch still fills the SMs
  int64_t graphs = 1;  // replay multi-launch applies as CUDA graphs
  int64_t p2p = 1;     // kernels of one rank read / write another local rank's buffers in place
  int64_t pdl = 1;     // programmatic dependent launch of the tile kernels
  // 1: the tile kernels run one resident wave of CTAs that draw tiles from a device-side queue; 0: one CTA
  // per tile, handed out by the hardware scheduler.  Measured on B200 (profiles/r02_persistent_ab.txt): the
  // queue is 2-3 % slower on 0.5 GiB shards and 0.5 % slower on 4 GiB, so it is off by default.
  int64_t persistent = 0;
  int64_t prefetch = 32768;  // bytes each first-wave CTA prefetches to L2 ahead of the dependency wait
  uint64_t epoch = 0;  // bumped by every set_param: captured graphs of an older epoch are dropped
  int64_t ipc = 1;     // one rank per process: exchange through CUDA-IPC-mapped peer buffers and epoch flags instead of NCCL
  int64_t xp_kwait = 1;  // peer-memory exchange: epoch-flag waits inside the tile kernels' prologue (0: separate wait kernels)
  int64_t xp_side = 0;   // peer-memory exchange: input pushes on a side stream (0: on the rank's main stream)
  int64_t xp_push = 1;   // peer-memory exchange: one push kernel (wait + copy to all consumers + signal); 0: wait kernel, one peer copy per consumer, signal kernel
  std::vector<void*> graveyard;  // exchange buffers other processes may still have mapped: freed with the context
  djets_allgather_fn host_allgather = nullptr;  // caller-provided control-plane collective (no NCCL needed)
  void* host_allgather_user = nullptr;
  bool capturing = false;  // between djets_ctx_capture_begin and _end: every call is recorded, nothing may synchronise
  std::vector<int> free_scalars;  // slots of d_scalars not in use
  int next_scalar = 0;
  std::vector<char> peer;  // peer[a * ndev + b]: device a can address device b's memory
  int ndev = 0;
  bool can_address(const RankCtx& a, const RankCtx& b) const {
    return a.device == b.device || peer[(size_t)a.device * ndev + b.device];
  }

  RankCtx* local(int rank) {
    if (rank < 0 || rank >= world) return nullptr;
    const int l = local_of_rank[rank];
    return l < 0 ? nullptr : &ranks[l];
  }
};

namespace djets {

void use(RankCtx& rc);
cudaEvent_t pool_event(RankCtx& rc);
void drain_events(RankCtx& rc);
void stream_after(RankCtx& src, RankCtx& dst);  // dst stream waits for everything enqueued so far on src stream
unsigned long long* trace_slot(RankCtx& rc, int kind, int ctas);  // nullptr unless DJETS_TRACE is set
// cudaMalloc that gives the per-rank caches of freed vector parts back to the driver and retries once on OOM
cudaError_t device_alloc(djets_ctx ctx, void** p, size_t bytes);
void flush_vec_caches(djets_ctx ctx);
// All-gather of one fixed-size record per process among the processes of the job: the host callback when the caller
// provided one, else NCCL on rank-local device buffers.  `recv` holds world records, indexed by rank (records of
// ranks that are not a process' first local rank are zero).
void control_allgather(djets_ctx ctx, const void* send, void* recv, size_t bytes);
inline bool has_control_plane(djets_ctx ctx) { return ctx->host_allgather != nullptr || ctx->has_nccl; }
inline void no_sync_in_capture(djets_ctx ctx, const char* what) {
  if (ctx->capturing) fail(DJETS_ERR_INVALID, std::string(what) + " synchronises with the host: not allowed between djets_ctx_capture_begin and _end");
}

template <class F>
void launch_counted(djets_ctx ctx, RankCtx& rc, int kind, F&& f) {
  rc.launches[kind] += 1;
  if (ctx->timing) {
    if (rc.pending.size() > 200000) drain_events(rc);
    KernelEvent k{kind, pool_event(rc), pool_event(rc)};
    CK(cudaEventRecord(k.a, rc.stream));
    CK(f());
    CK(cudaEventRecord(k.b, rc.stream));
    rc.pending.push_back(k);
  } else {
    CK(f());
  }
}

}  // namespace djets

// A scalar that lives on the devices: one slot of every local rank's d_scalars (replicas hold the same value).
struct djets_scalar_s {
  djets_ctx ctx = nullptr;
  int slot = 0;
};

// A captured sequence of calls (djets_ctx_capture_begin / _end) replayed with one launch.
struct djets_graph_s {
  djets_ctx ctx = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches[DJETS_K_COUNT] = {0, 0, 0, 0, 0};
};

// ---- DBArray --------------------------------------------------------------------------------------
struct VecPart {
  int rank = 0;
  int64_t blk_first = 0, blk_count = 0, elem_first = 0, elem_count = 0;
  char* d = nullptr;  // device storage on the owner (nullptr when the owner is another process)
  size_t cap = 0;     // allocated bytes
  // an asynchronous host copy of this part is (or may still be) running on the owner's copy stream: the next call that
  // touches the vector makes the main stream wait for `pending` first
  cudaEvent_t pending = nullptr;
  bool has_pending = false;
};

struct djets_vec_s {
  djets_ctx ctx = nullptr;
  int dtype = 0;
  std::vector<VecPart> parts;
  std::vector<int64_t> block_len, block_off;  // block_off: global element offset, nblocks+1 entries
  std::vector<int> block_part;
  int64_t length = 0;
};

namespace djets {

djets_vec make_vec(djets_ctx ctx, int dtype, int nparts, const int32_t* part_rank, const int64_t* part_nblocks,
                   const int64_t* block_len);
// Orders the main streams after any asynchronous host copy still running on this vector's parts.  Called by every
// entry point for each vector it touches (a host-side flag test unless a copy is pending).
inline void settle(djets_vec v) {
  if (!v) return;
  for (VecPart& p : v->parts)
    if (p.has_pending) {
      if (RankCtx* rc = v->ctx->local(p.rank)) CK(cudaStreamWaitEvent(rc->stream, p.pending, 0));
      p.has_pending = false;
    }
}

}  // namespace djets
