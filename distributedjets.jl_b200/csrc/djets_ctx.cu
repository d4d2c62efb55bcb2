// Contexts of libdjets_b200: one rank = one GPU = one former Distributed.jl worker; streams, events,
// timers, per-kernel accounting, tuning parameters, and the host-only planning entry points.
#include "djets_internal.h"

using namespace djets;

namespace djets {

thread_local std::string g_last_error;
std::recursive_mutex g_api_mutex;


void use(RankCtx& rc) { CK(cudaSetDevice(rc.device)); }

cudaEvent_t pool_event(RankCtx& rc) {
  if (!rc.pool.empty()) {
    cudaEvent_t e = rc.pool.back();
    rc.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  return e;
}

void drain_events(RankCtx& rc) {
  if (rc.pending.empty()) return;
  use(rc);
  CK(cudaStreamSynchronize(rc.stream));
  for (auto& k : rc.pending) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, k.a, k.b));
    rc.ms[k.kind] += ms;
    rc.pool.push_back(k.a);
    rc.pool.push_back(k.b);
  }
  rc.pending.clear();
}

void flush_vec_caches(djets_ctx ctx) {
  for (RankCtx& rc : ctx->ranks) {
    if (rc.vec_cache.empty()) continue;
    cudaSetDevice(rc.device);
    cudaStreamSynchronize(rc.stream);
    for (auto& kv : rc.vec_cache) cudaFree(kv.second);
    rc.vec_cache.clear();
    rc.vec_cache_bytes = 0;
  }
}

cudaError_t device_alloc(djets_ctx ctx, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaErrorMemoryAllocation) return e;
  (void)cudaGetLastError();
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ctx) return cudaErrorMemoryAllocation;
  flush_vec_caches(ctx);  // freed DBArray parts held for reuse, on every rank of this process
  cudaSetDevice(dev);
  return cudaMalloc(p, bytes);
}

void control_allgather(djets_ctx ctx, const void* send, void* recv, size_t bytes) {
  REQUIRE(bytes > 0 && bytes <= (size_t)1 << 24, DJETS_ERR_INVALID, "control all-gather: record size out of range");
  RankCtx& rc = ctx->ranks.front();
  if (ctx->host_allgather) {
    // the callback orders records by process; a process stands for its first local rank
    std::vector<char> tmp(bytes * (size_t)ctx->world, 0), mine(bytes + sizeof(int32_t));
    const int32_t myrank = rc.rank;
    memcpy(mine.data(), &myrank, sizeof(int32_t));
    memcpy(mine.data() + sizeof(int32_t), send, bytes);
    std::vector<char> all((bytes + sizeof(int32_t)) * (size_t)ctx->world, 0);
    // the number of processes is not known to the library: the callback fills one record per process and leaves the
    // rest of `all` untouched (rank field -1 marks unused slots)
    for (size_t k = 0; k < (size_t)ctx->world; ++k) {
      const int32_t none = -1;
      memcpy(all.data() + k * (bytes + sizeof(int32_t)), &none, sizeof(int32_t));
    }
    const int32_t st = ctx->host_allgather(ctx->host_allgather_user, mine.data(), all.data(), (int64_t)(bytes + sizeof(int32_t)));
    REQUIRE(st == 0, DJETS_ERR_NCCL, "the host all-gather callback failed");
    memset(recv, 0, bytes * (size_t)ctx->world);
    for (size_t k = 0; k < (size_t)ctx->world; ++k) {
      int32_t r;
      memcpy(&r, all.data() + k * (bytes + sizeof(int32_t)), sizeof(int32_t));
      if (r >= 0 && r < ctx->world)
        memcpy(reinterpret_cast<char*>(recv) + (size_t)r * bytes, all.data() + k * (bytes + sizeof(int32_t)) + sizeof(int32_t), bytes);
    }
    return;
  }
  REQUIRE(ctx->has_nccl, DJETS_ERR_NCCL, "a multi-process job needs a host all-gather callback or an NCCL unique id");
  const char* why = nullptr;
  const NcclApi* api = nccl_api(&why);
  REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
  use(rc);
  char *d_send = nullptr, *d_recv = nullptr;
  CK(cudaMalloc(&d_send, bytes));
  CK(cudaMalloc(&d_recv, bytes * (size_t)ctx->world));
  CK(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, rc.stream));
  NCK(api, api->GroupStart());
  for (RankCtx& each : ctx->ranks) {  // every local rank takes part in the collective; only the first carries data
    use(each);
    if (&each == &rc) {
      NCK(api, api->AllGather(d_send, d_recv, bytes, ncclChar, each.comm, each.stream));
    } else {
      char *s2 = nullptr, *r2 = nullptr;
      CK(cudaMalloc(&s2, bytes));
      CK(cudaMalloc(&r2, bytes * (size_t)ctx->world));
      CK(cudaMemsetAsync(s2, 0, bytes, each.stream));
      NCK(api, api->AllGather(s2, r2, bytes, ncclChar, each.comm, each.stream));
      ctx->graveyard.push_back(s2);
      ctx->graveyard.push_back(r2);
    }
  }
  NCK(api, api->GroupEnd());
  use(rc);
  CK(cudaMemcpyAsync(recv, d_recv, bytes * (size_t)ctx->world, cudaMemcpyDeviceToHost, rc.stream));
  CK(cudaStreamSynchronize(rc.stream));
  cudaFree(d_send);
  cudaFree(d_recv);
}

unsigned long long* trace_slot(RankCtx& rc, int kind, int ctas) {
  static const char* prefix = getenv("DJETS_TRACE");
  if (!prefix || rc.trace_launches >= kTraceLaunches || ctas > kTraceCtas) return nullptr;
  if (!rc.d_trace) {
    CK(cudaMalloc(&rc.d_trace, sizeof(unsigned long long) * 4 * kTraceCtas * kTraceLaunches));
    CK(cudaMemset(rc.d_trace, 0, sizeof(unsigned long long) * 4 * kTraceCtas * kTraceLaunches));
  }
  rc.trace_meta.push_back(kind);
  rc.trace_meta.push_back(ctas);
  return rc.d_trace + (size_t)4 * kTraceCtas * rc.trace_launches++;
}

static void trace_dump(RankCtx& rc) {
  const char* prefix = getenv("DJETS_TRACE");
  if (!prefix || !rc.d_trace || rc.trace_launches == 0) return;
  const std::string path = std::string(prefix) + ".rank" + std::to_string(rc.rank) + ".bin";
  if (FILE* f = fopen(path.c_str(), "wb")) {
    std::vector<unsigned long long> h((size_t)4 * kTraceCtas);
    int n = rc.trace_launches;
    fwrite(&n, sizeof(int), 1, f);
    fwrite(rc.trace_meta.data(), sizeof(int), rc.trace_meta.size(), f);
    for (int l = 0; l < n; ++l) {
      const int ctas = rc.trace_meta[2 * l + 1];
      cudaMemcpy(h.data(), rc.d_trace + (size_t)4 * kTraceCtas * l, sizeof(unsigned long long) * 4 * ctas,
                 cudaMemcpyDeviceToHost);
      fwrite(h.data(), sizeof(unsigned long long), (size_t)4 * ctas, f);
    }
    fclose(f);
  }
}

// dst stream waits for everything enqueued so far on src stream
void stream_after(RankCtx& src, RankCtx& dst) {
  if (&src == &dst) return;
  // no cudaSetDevice: event record / stream wait only need event and stream to agree, not the current device
  CK(cudaEventRecord(src.ev, src.stream));
  CK(cudaStreamWaitEvent(dst.stream, src.ev, 0));
}


}  // namespace djets

extern "C" {

int32_t djets_version(void) { return 100; }
const char* djets_last_error(void) { return g_last_error.c_str(); }

int32_t djets_even_cuts(int64_t n, int32_t nparts, int64_t* cuts) {
  return guard([&] {
    REQUIRE(n >= 0 && nparts >= 1 && cuts, DJETS_ERR_INVALID, "even_cuts: bad arguments");
    const int64_t q = n / nparts, rem = n % nparts;
    cuts[0] = 0;
    for (int k = 0; k < nparts; ++k) cuts[k + 1] = cuts[k] + q + (k < rem ? 1 : 0);
  });
}

int32_t djets_grid_rank(int32_t n1, int32_t n2, int32_t r, int32_t c, const int32_t* grid_ranks, int32_t* rank) {
  return guard([&] {
    REQUIRE(n1 >= 1 && n2 >= 1 && r >= 0 && r < n1 && c >= 0 && c < n2 && rank, DJETS_ERR_INVALID,
            "grid_rank: bad arguments");
    const int idx = r + c * n1;
    *rank = grid_ranks ? grid_ranks[idx] : idx;
  });
}

int32_t djets_device_count(int32_t* n) {
  return guard([&] {
    int c = 0;
    CK(cudaGetDeviceCount(&c));
    *n = c;
  });
}

int32_t djets_nccl_unique_id(uint8_t* out128) {
  return guard([&] {
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCK(api, api->GetUniqueId(&id));
    memcpy(out128, &id, 128);
  });
}

int32_t djets_host_alloc(int64_t bytes, void** out) {
  return guard([&] {
    REQUIRE(bytes >= 0 && out, DJETS_ERR_INVALID, "host_alloc: bad arguments");
    CK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable));
  });
}
int32_t djets_host_free(void* p) {
  return guard([&] {
    if (p) CK(cudaFreeHost(p));
  });
}

int32_t djets_ctx_create(int32_t world, int32_t nlocal, const int32_t* local_ranks, const int32_t* device_ids,
                         const uint8_t* nccl_uid, djets_ctx* out) {
  return guard([&] {
    REQUIRE(out, DJETS_ERR_INVALID, "ctx_create: null out");
    REQUIRE(world >= 1 && nlocal >= 1 && nlocal <= world, DJETS_ERR_INVALID, "ctx_create: bad world / nlocal");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQUIRE(ndev > 0, DJETS_ERR_CUDA, "no CUDA device: libdjets_b200 has no CPU path");
    auto ctx = std::make_unique<djets_ctx_s>();
    ctx->world = world;
    ctx->local_of_rank.assign(world, -1);
    ctx->ranks.resize(nlocal);
    for (int l = 0; l < nlocal; ++l) {
      RankCtx& rc = ctx->ranks[l];
      rc.rank = local_ranks ? local_ranks[l] : l;
      rc.device = device_ids ? device_ids[l] : (l % ndev);
      REQUIRE(rc.rank >= 0 && rc.rank < world && ctx->local_of_rank[rc.rank] < 0, DJETS_ERR_INVALID,
              "ctx_create: bad or repeated local rank");
      REQUIRE(rc.device >= 0 && rc.device < ndev, DJETS_ERR_INVALID, "ctx_create: device id out of range");
      ctx->local_of_rank[rc.rank] = l;
      use(rc);
      cudaDeviceProp prop;
      CK(cudaGetDeviceProperties(&prop, rc.device));
      REQUIRE(prop.major >= 10, DJETS_ERR_CUDA,
              std::string("device ") + prop.name + " is not sm_100-class: libdjets_b200 carries sm_100a code only");
      CK(cudaStreamCreateWithFlags(&rc.stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&rc.ev, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&rc.ev_fork, cudaEventDisableTiming));
      CK(cudaStreamCreateWithFlags(&rc.cstream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&rc.ev_c0, cudaEventDisableTiming));
      for (cudaEvent_t& e : rc.cmarks) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CK(cudaStreamCreateWithFlags(&rc.xstream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&rc.ev_x0, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&rc.ev_x1, cudaEventDisableTiming));
      CK(cudaEventCreate(&rc.t0));
      CK(cudaEventCreate(&rc.t1));
      CK(cudaMalloc(&rc.d_partials, sizeof(double) * reduce_max_ctas()));
      CK(cudaMalloc(&rc.d_ticket, sizeof(unsigned)));
      CK(cudaMemset(rc.d_ticket, 0, sizeof(unsigned)));
      CK(cudaMalloc(&rc.d_result, sizeof(double) * 2 * kMaxParts));
      CK(cudaMalloc(&rc.d_gather, sizeof(double) * kMaxParts));
      CK(cudaMalloc(&rc.d_xp_err, sizeof(unsigned)));
      CK(cudaMemset(rc.d_xp_err, 0, sizeof(unsigned)));
      CK(cudaMalloc(&rc.d_xp_ticket, sizeof(unsigned)));
      CK(cudaMemset(rc.d_xp_ticket, 0, sizeof(unsigned)));
      CK(cudaMalloc(&rc.d_scalars, sizeof(double) * kMaxScalars));
      CK(cudaMemset(rc.d_scalars, 0, sizeof(double) * kMaxScalars));
      for (cudaEvent_t& e : rc.marks) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CK(cudaHostAlloc(&rc.h_result, sizeof(double) * 2 * kMaxParts, cudaHostAllocPortable | cudaHostAllocMapped));
      CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&rc.h_result_dev), rc.h_result, 0));
    }
    // peer access between the GPUs this process drives: local->local exchange is in-place peer
    // loads / stores from the kernels (or a peer copy when p2p is off)
    ctx->ndev = ndev;
    ctx->peer.assign((size_t)ndev * ndev, 0);
    for (RankCtx& a : ctx->ranks)
      for (RankCtx& b : ctx->ranks) {
        if (a.device == b.device) continue;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, a.device, b.device));
        if (can) {
          use(a);
          cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
          (void)cudaGetLastError();
          ctx->peer[(size_t)a.device * ndev + b.device] = 1;
        }
      }
    if (nccl_uid) {
      const char* why = nullptr;
      const NcclApi* api = nccl_api(&why);
      REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
      ncclUniqueId id;
      memcpy(&id, nccl_uid, 128);
      NCK(api, api->GroupStart());
      for (RankCtx& rc : ctx->ranks) {
        use(rc);
        NCK(api, api->CommInitRank(&rc.comm, world, id, rc.rank));
      }
      NCK(api, api->GroupEnd());
      ctx->has_nccl = true;
    }
    *out = ctx.release();
  });
}

int32_t djets_ctx_destroy(djets_ctx ctx) {
  return guard([&] {
    if (!ctx) return;
    const NcclApi* api = ctx->has_nccl ? nccl_api(nullptr) : nullptr;
    for (RankCtx& rc : ctx->ranks) {
      cudaSetDevice(rc.device);
      cudaStreamSynchronize(rc.stream);
      trace_dump(rc);
      cudaFree(rc.d_trace);
      if (rc.comm && api) api->CommDestroy(rc.comm);
      for (auto& k : rc.pending) {
        cudaEventDestroy(k.a);
        cudaEventDestroy(k.b);
      }
      for (auto e : rc.pool) cudaEventDestroy(e);
      for (auto& kv : rc.vec_cache) cudaFree(kv.second);
      cudaFree(rc.d_partials);
      cudaFree(rc.d_ticket);
      cudaFree(rc.d_gather);
      cudaFree(rc.d_xp_err);
      cudaFree(rc.d_xp_ticket);
      cudaFree(rc.d_scalars);
      for (cudaEvent_t e : rc.marks)
        if (e) cudaEventDestroy(e);
      cudaFree(rc.d_result);
      cudaFreeHost(rc.h_result);
      cudaEventDestroy(rc.ev);
      cudaEventDestroy(rc.ev_fork);
      cudaStreamSynchronize(rc.xstream);
      cudaStreamDestroy(rc.xstream);
      cudaStreamSynchronize(rc.cstream);
      cudaStreamDestroy(rc.cstream);
      cudaEventDestroy(rc.ev_c0);
      for (cudaEvent_t e : rc.cmarks)
        if (e) cudaEventDestroy(e);
      cudaEventDestroy(rc.ev_x0);
      cudaEventDestroy(rc.ev_x1);
      cudaEventDestroy(rc.t0);
      cudaEventDestroy(rc.t1);
      cudaStreamDestroy(rc.stream);
    }
    for (void* p : ctx->graveyard) cudaFree(p);
    delete ctx;
  });
}

int32_t djets_ctx_set_host_allgather(djets_ctx ctx, djets_allgather_fn fn, void* user) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    ctx->host_allgather = fn;
    ctx->host_allgather_user = user;
  });
}

int32_t djets_ctx_sync(djets_ctx ctx) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    no_sync_in_capture(ctx, "djets_ctx_sync");
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      CK(cudaStreamSynchronize(rc.stream));
      CK(cudaStreamSynchronize(rc.cstream));
    }
    if ((int)ctx->ranks.size() < ctx->world)
      for (RankCtx& rc : ctx->ranks) {
        unsigned err = 0;
        use(rc);
        CK(cudaMemcpy(&err, rc.d_xp_err, sizeof(unsigned), cudaMemcpyDeviceToHost));
        REQUIRE(err == 0, DJETS_ERR_NCCL,
                "cross-process exchange: a peer did not raise its flag within 20 s (a process died or the call "
                "sequences of the processes differ)");
      }
  });
}

int32_t djets_ctx_info(djets_ctx ctx, int32_t* world, int32_t* nlocal, int32_t* has_nccl) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    if (world) *world = ctx->world;
    if (nlocal) *nlocal = (int32_t)ctx->ranks.size();
    if (has_nccl) *has_nccl = ctx->has_nccl ? 1 : 0;
  });
}

int32_t djets_ctx_is_local(djets_ctx ctx, int32_t rank, int32_t* is_local) {
  return guard([&] {
    REQUIRE(ctx && is_local, DJETS_ERR_INVALID, "null argument");
    *is_local = ctx->local(rank) ? 1 : 0;
  });
}

int32_t djets_timer_start(djets_ctx ctx) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      CK(cudaEventRecord(rc.t0, rc.stream));
    }
  });
}

int32_t djets_timer_stop(djets_ctx ctx, float* ms) {
  return guard([&] {
    REQUIRE(ctx && ms, DJETS_ERR_INVALID, "null argument");
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      CK(cudaEventRecord(rc.t1, rc.stream));
    }
    float worst = 0.f;
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      CK(cudaEventSynchronize(rc.t1));
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, rc.t0, rc.t1));
      worst = std::max(worst, t);
    }
    *ms = worst;
  });
}

int32_t djets_ctx_kernel_timing(djets_ctx ctx, int32_t enable) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    if (!enable)
      for (RankCtx& rc : ctx->ranks) drain_events(rc);
    ctx->timing = enable != 0;
  });
}

int32_t djets_ctx_kernel_stats(djets_ctx ctx, int32_t kind, int64_t* launches, double* total_ms) {
  return guard([&] {
    REQUIRE(ctx && kind >= 0 && kind < DJETS_K_COUNT, DJETS_ERR_INVALID, "kernel_stats: bad arguments");
    int64_t n = 0;
    double ms = 0.0;
    for (RankCtx& rc : ctx->ranks) {
      drain_events(rc);
      n += rc.launches[kind];
      ms = std::max(ms, rc.ms[kind]);
    }
    if (launches) *launches = n;
    if (total_ms) *total_ms = ms;
  });
}

int32_t djets_ctx_reset_stats(djets_ctx ctx) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    for (RankCtx& rc : ctx->ranks) {
      drain_events(rc);
      rc.trace_launches = 0;
      rc.trace_meta.clear();
      for (int k = 0; k < DJETS_K_COUNT; ++k) {
        rc.launches[k] = 0;
        rc.ms[k] = 0.0;
      }
    }
  });
}


// ---- stream marks: wait for a point of the stream without draining what was enqueued after it ----------
int32_t djets_ctx_mark(djets_ctx ctx, int32_t slot) {
  return guard([&] {
    REQUIRE(ctx && slot >= 0 && slot < 8, DJETS_ERR_INVALID, "mark: slot in 0..7");
    REQUIRE(!ctx->capturing, DJETS_ERR_INVALID, "mark inside a captured sequence");
    for (RankCtx& rc : ctx->ranks) {
      CK(cudaEventRecord(rc.marks[slot], rc.stream));
      CK(cudaEventRecord(rc.cmarks[slot], rc.cstream));
    }
  });
}

int32_t djets_ctx_wait_mark(djets_ctx ctx, int32_t slot) {
  return guard([&] {
    REQUIRE(ctx && slot >= 0 && slot < 8, DJETS_ERR_INVALID, "wait_mark: slot in 0..7");
    no_sync_in_capture(ctx, "djets_ctx_wait_mark");
    for (RankCtx& rc : ctx->ranks) {
      CK(cudaEventSynchronize(rc.marks[slot]));
      CK(cudaEventSynchronize(rc.cmarks[slot]));
    }
  });
}

// ---- captured sequences -----------------------------------------------------------------------------------
int32_t djets_ctx_capture_begin(djets_ctx ctx) {
  return guard([&] {
    REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
    REQUIRE(!ctx->capturing, DJETS_ERR_INVALID, "capture_begin: a capture is already open");
    REQUIRE((int)ctx->ranks.size() == ctx->world, DJETS_ERR_UNSUPPORTED,
            "capture needs every rank of the job in this process (NCCL exchanges stay outside graphs)");
    REQUIRE(!ctx->timing, DJETS_ERR_INVALID, "capture_begin: per-kernel timing is on");
    RankCtx& origin = ctx->ranks.front();
    use(origin);
    for (size_t k = 1; k < ctx->ranks.size(); ++k) stream_after(ctx->ranks[k], origin);  // earlier work of the other ranks
    CK(cudaStreamBeginCapture(origin.stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    try {
      CK(cudaEventRecord(origin.ev_fork, origin.stream));
      for (size_t k = 1; k < ctx->ranks.size(); ++k) CK(cudaStreamWaitEvent(ctx->ranks[k].stream, origin.ev_fork, 0));
    } catch (...) {
      cudaGraph_t g = nullptr;
      cudaStreamEndCapture(origin.stream, &g);
      if (g) cudaGraphDestroy(g);
      (void)cudaGetLastError();
      ctx->capturing = false;
      throw;
    }
    for (RankCtx& rc : ctx->ranks)
      for (int k = 0; k < DJETS_K_COUNT; ++k) rc.launches_mark[k] = rc.launches[k];
  });
}

int32_t djets_ctx_capture_end(djets_ctx ctx, djets_graph* out) {
  return guard([&] {
    REQUIRE(ctx && out, DJETS_ERR_INVALID, "capture_end: null argument");
    REQUIRE(ctx->capturing, DJETS_ERR_INVALID, "capture_end without capture_begin");
    RankCtx& origin = ctx->ranks.front();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaSuccess;
    for (size_t k = 1; k < ctx->ranks.size() && e == cudaSuccess; ++k) {
      e = cudaEventRecord(ctx->ranks[k].ev_fork, ctx->ranks[k].stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(origin.stream, ctx->ranks[k].ev_fork, 0);
    }
    use(origin);
    cudaError_t e2 = cudaStreamEndCapture(origin.stream, &graph);
    ctx->capturing = false;
    auto g = std::make_unique<djets_graph_s>();
    g->ctx = ctx;
    for (RankCtx& rc : ctx->ranks)  // the capture launched nothing: the counters go to the graph, to be added per replay
      for (int k = 0; k < DJETS_K_COUNT; ++k) {
        g->launches[k] += rc.launches[k] - rc.launches_mark[k];
        rc.launches[k] = rc.launches_mark[k];
      }
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      CK(e != cudaSuccess ? e : e2);
    }
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    CK(e);
    *out = g.release();
  });
}

int32_t djets_graph_launch(djets_graph g) {
  return guard([&] {
    REQUIRE(g && g->exec, DJETS_ERR_INVALID, "null graph");
    djets_ctx ctx = g->ctx;
    REQUIRE(!ctx->capturing, DJETS_ERR_INVALID, "graph_launch inside a capture");
    RankCtx& origin = ctx->ranks.front();
    for (size_t k = 1; k < ctx->ranks.size(); ++k) stream_after(ctx->ranks[k], origin);
    use(origin);
    CK(cudaGraphLaunch(g->exec, origin.stream));
    if (ctx->ranks.size() > 1) {
      CK(cudaEventRecord(origin.ev_fork, origin.stream));
      for (size_t k = 1; k < ctx->ranks.size(); ++k) CK(cudaStreamWaitEvent(ctx->ranks[k].stream, origin.ev_fork, 0));
    }
    for (int k = 0; k < DJETS_K_COUNT; ++k) origin.launches[k] += g->launches[k];
  });
}

int32_t djets_graph_destroy(djets_graph g) {
  return guard([&] {
    if (!g) return;
    if (g->exec) {
      cudaSetDevice(g->ctx->ranks.front().device);
      cudaStreamSynchronize(g->ctx->ranks.front().stream);
      cudaGraphExecDestroy(g->exec);
    }
    delete g;
  });
}

// ---- device-resident scalars ------------------------------------------------------------------------------
int32_t djets_scalar_create(djets_ctx ctx, djets_scalar* out) {
  return guard([&] {
    REQUIRE(ctx && out, DJETS_ERR_INVALID, "scalar_create: null argument");
    int slot;
    if (!ctx->free_scalars.empty()) {
      slot = ctx->free_scalars.back();
      ctx->free_scalars.pop_back();
    } else {
      REQUIRE(ctx->next_scalar < kMaxScalars, DJETS_ERR_INVALID, "scalar_create: too many live scalars");
      slot = ctx->next_scalar++;
    }
    auto s = std::make_unique<djets_scalar_s>();
    s->ctx = ctx;
    s->slot = slot;
    *out = s.release();
  });
}

int32_t djets_scalar_destroy(djets_scalar s) {
  return guard([&] {
    if (!s) return;
    s->ctx->free_scalars.push_back(s->slot);
    delete s;
  });
}

int32_t djets_scalar_set(djets_scalar s, double v) {
  return guard([&] {
    REQUIRE(s, DJETS_ERR_INVALID, "null scalar");
    for (RankCtx& rc : s->ctx->ranks) {
      use(rc);
      launch_counted(s->ctx, rc, DJETS_K_REDUCE, [&] { return launch_scalar_set(v, rc.d_scalars + s->slot, rc.stream); });
    }
  });
}

int32_t djets_scalar_get(djets_scalar s, double* out) {
  return guard([&] {
    REQUIRE(s && out, DJETS_ERR_INVALID, "scalar_get: null argument");
    no_sync_in_capture(s->ctx, "djets_scalar_get");
    RankCtx& rc = s->ctx->ranks.front();
    use(rc);
    CK(cudaMemcpyAsync(out, rc.d_scalars + s->slot, sizeof(double), cudaMemcpyDeviceToHost, rc.stream));
    CK(cudaStreamSynchronize(rc.stream));
  });
}

int32_t djets_scalar_op(djets_scalar out, int32_t op, djets_scalar a, djets_scalar b, int32_t dtype) {
  return guard([&] {
    REQUIRE(out && a, DJETS_ERR_INVALID, "scalar_op: null argument");
    REQUIRE(op >= SC_ADD && op <= SC_MIN, DJETS_ERR_INVALID, "scalar_op: unknown operation");
    REQUIRE(dtype == DJETS_F32 || dtype == DJETS_F64, DJETS_ERR_INVALID, "scalar_op: dtype");
    REQUIRE(a->ctx == out->ctx && (!b || b->ctx == out->ctx), DJETS_ERR_INVALID, "scalar_op: scalars of different contexts");
    const bool binary = op <= SC_DIV || op >= SC_MAX;
    REQUIRE(!binary || b, DJETS_ERR_INVALID, "scalar_op: binary operation needs two operands");
    for (RankCtx& rc : out->ctx->ranks) {  // every replica repeats the same arithmetic: no exchange
      use(rc);
      launch_counted(out->ctx, rc, DJETS_K_REDUCE, [&] {
        return launch_scalar_op(dtype, op, rc.d_scalars + a->slot, b ? rc.d_scalars + b->slot : nullptr,
                                rc.d_scalars + out->slot, rc.stream);
      });
    }
  });
}

int32_t djets_ctx_set_param(djets_ctx ctx, const char* name, int64_t value) {
  return guard([&] {
    REQUIRE(ctx && name, DJETS_ERR_INVALID, "null argument");
    const std::string n(name);
    if (n == "fwd_tx") {
      REQUIRE(value == 32 || value == 64 || value == 128 || value == 256, DJETS_ERR_INVALID, "fwd_tx in {32,64,128,256}");
      ctx->fwd_tx = value;
    } else if (n == "fwd_tc") {
      REQUIRE(value >= 1 && value <= kFwdMaxTC, DJETS_ERR_INVALID, "fwd_tc in 1..1024");
      ctx->fwd_tc = value;
    } else if (n == "adj_ru") {
      REQUIRE(value == 1 || value == 2 || value == 4, DJETS_ERR_INVALID, "adj_ru in {1,2,4}");
      ctx->adj_ru = value;
    } else if (n == "adj_tc") {
      REQUIRE(value >= kAdjCW && value <= 4096, DJETS_ERR_INVALID, "adj_tc in 4..4096");
      ctx->adj_tc = value;
    } else if (n == "ld_policy") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "ld_policy in {0,1}");
      ctx->ld_policy = value;
    } else if (n == "pdl") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "pdl in {0,1}");
      ctx->pdl = value;
    } else if (n == "persistent") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "persistent in {0,1}");
      ctx->persistent = value;
    } else if (n == "ipc") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "ipc in {0,1}");
      ctx->ipc = value;
    } else if (n == "xp_kwait" || n == "xp_side" || n == "xp_push") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "xp_kwait / xp_side / xp_push in {0,1}");
      (n == "xp_kwait" ? ctx->xp_kwait : (n == "xp_side" ? ctx->xp_side : ctx->xp_push)) = value;
    } else if (n == "prefetch") {
      REQUIRE(value >= 0 && value <= (1 << 20), DJETS_ERR_INVALID, "prefetch in 0..1048576 bytes");
      ctx->prefetch = value;
    } else if (n == "l2_keep") {
      REQUIRE(value >= 0 && value <= 120, DJETS_ERR_INVALID, "l2_keep in 0..120 MB");
      ctx->l2_keep = value;
    } else if (n == "wave_fit") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "wave_fit in {0,1}");
      ctx->wave_fit = value;
    } else if (n == "chain") {
      REQUIRE(value >= 0 && value <= kChainMax, DJETS_ERR_INVALID, "chain in 0..16");
      ctx->chain = value;
    } else if (n == "adj_light") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "adj_light in {0,1}");
      ctx->adj_light = value;
    } else if (n == "auto_tile") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "auto_tile in {0,1}");
      ctx->auto_tile = value;
    } else if (n == "p2p") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "p2p in {0,1}");
      ctx->p2p = value;
    } else if (n == "graphs") {
      REQUIRE(value == 0 || value == 1, DJETS_ERR_INVALID, "graphs in {0,1}");
      ctx->graphs = value;
    } else {
      fail(DJETS_ERR_INVALID, "unknown parameter " + n);
    }
    ctx->epoch += 1;  // applies captured under the old parameters are re-captured (djets_op.cu: apply)
  });
}

}  // extern "C"
