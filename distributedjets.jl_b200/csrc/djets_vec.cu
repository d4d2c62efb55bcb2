// DBArray side of libdjets_b200 (reference: src/DistributedJets.jl:109-412): device-resident block
// vectors, block access, reductions with the reference's master-side folds, closed-set broadcasting.
#include "djets_internal.h"

using namespace djets;

namespace {


constexpr size_t kVecCacheLimit = size_t(32) << 30;  // bytes kept per rank before freed parts go back to the driver

char* part_alloc(djets_ctx ctx, RankCtx& rc, size_t bytes) {
  auto it = rc.vec_cache.find(bytes);
  if (it != rc.vec_cache.end()) {
    char* p = it->second;
    rc.vec_cache.erase(it);
    rc.vec_cache_bytes -= bytes;
    return p;  // every earlier use was ordered before the free on this rank's stream
  }
  REQUIRE(!ctx->capturing, DJETS_ERR_INVALID, "allocating a DBArray inside a captured sequence: create the vectors first");
  void* p = nullptr;
  CK(device_alloc(ctx, &p, bytes));  // on OOM every local rank's cache goes back to the driver first
  return reinterpret_cast<char*>(p);
}

void part_free(RankCtx& rc, char* p, size_t bytes) {
  if (rc.vec_cache_bytes + bytes > kVecCacheLimit) {
    cudaStreamSynchronize(rc.stream);
    cudaFree(p);
    return;
  }
  rc.vec_cache.emplace(bytes, p);
  rc.vec_cache_bytes += bytes;
}

}  // namespace

namespace djets {


djets_vec make_vec(djets_ctx ctx, int dtype, int nparts, const int32_t* part_rank, const int64_t* part_nblocks,
                   const int64_t* block_len) {
  REQUIRE(ctx, DJETS_ERR_INVALID, "null context");
  REQUIRE(dtype >= DJETS_F32 && dtype <= DJETS_C128, DJETS_ERR_INVALID, "dtype must be DJETS_F32, F64, C64 or C128");
  REQUIRE(nparts >= 1 && nparts <= kMaxParts, DJETS_ERR_INVALID, "nparts out of range");
  auto v = std::make_unique<djets_vec_s>();
  v->ctx = ctx;
  v->dtype = dtype;
  int64_t nblocks = 0;
  for (int p = 0; p < nparts; ++p) {
    REQUIRE(part_nblocks[p] >= 0, DJETS_ERR_INVALID, "negative block count");
    REQUIRE(part_rank[p] >= 0 && part_rank[p] < ctx->world, DJETS_ERR_INVALID, "part rank outside the job");
    nblocks += part_nblocks[p];
  }
  v->block_len.assign(block_len, block_len + nblocks);
  v->block_off.resize(nblocks + 1);
  v->block_part.resize(nblocks);
  v->block_off[0] = 0;
  for (int64_t b = 0; b < nblocks; ++b) {
    REQUIRE(block_len[b] >= 0, DJETS_ERR_INVALID, "negative block length");
    v->block_off[b + 1] = v->block_off[b] + block_len[b];
  }
  v->length = v->block_off[nblocks];
  v->parts.resize(nparts);
  int64_t b0 = 0;
  const int es = elem_size(dtype);
  for (int p = 0; p < nparts; ++p) {
    VecPart& part = v->parts[p];
    part.rank = part_rank[p];
    part.blk_first = b0;
    part.blk_count = part_nblocks[p];
    part.elem_first = v->block_off[b0];
    part.elem_count = v->block_off[b0 + part.blk_count] - part.elem_first;
    for (int64_t b = b0; b < b0 + part.blk_count; ++b) v->block_part[b] = p;
    b0 += part.blk_count;
    if (RankCtx* rc = ctx->local(part.rank)) {
      try {
        use(*rc);
        const size_t bytes = (size_t)round_up(part.elem_count * es + 256, 512);  // slack for 16-byte over-reads
        part.d = part_alloc(ctx, *rc, bytes);
        part.cap = bytes;
        CK(cudaMemsetAsync(part.d, 0, bytes, rc->stream));
      } catch (...) {  // give back what the earlier parts took
        for (int q = 0; q <= p; ++q)
          if (v->parts[q].d)
            if (RankCtx* rq = ctx->local(v->parts[q].rank)) {
              cudaSetDevice(rq->device);
              part_free(*rq, v->parts[q].d, v->parts[q].cap);
            }
        throw;
      }
    }
  }
  return v.release();
}

}  // namespace djets

namespace {

void free_vec(djets_vec v) {
  if (!v) return;
  for (VecPart& p : v->parts) {
    RankCtx* rc = v->ctx->local(p.rank);
    if (rc && p.has_pending) cudaStreamWaitEvent(rc->stream, p.pending, 0);  // the part is reused in main-stream order
    if (p.pending) cudaEventDestroy(p.pending);
    if (!p.d || !rc) continue;
    cudaSetDevice(rc->device);
    part_free(*rc, p.d, p.cap);
  }
  delete v;
}

bool same_layout(djets_vec a, djets_vec b) {
  if (a->ctx != b->ctx || a->parts.size() != b->parts.size() || a->block_len != b->block_len) return false;
  for (size_t p = 0; p < a->parts.size(); ++p)
    if (a->parts[p].rank != b->parts[p].rank || a->parts[p].blk_count != b->parts[p].blk_count) return false;
  return true;
}

double round_to(int dtype, double v) { return real_dtype(dtype) == DJETS_F32 ? (double)(float)v : v; }

// Per-part partial reductions: one result per part in out_parts (all parts, after the cross-process
// exchange when some parts live elsewhere).
bool kind_is_min(int kind) { return kind == DJETS_RED_MIN || kind == DJETS_RED_ABSMIN; }
bool kind_is_max(int kind) { return kind == DJETS_RED_MAX || kind == DJETS_RED_ABSMAX; }

// minimum / maximum over a worker's empty local part is an error in the reference (Julia: "reducing over an
// empty collection is not allowed"), not +-Inf
void check_extrema_parts(djets_vec x, int kind) {
  if (!kind_is_min(kind) && !kind_is_max(kind)) return;
  for (const VecPart& part : x->parts)
    REQUIRE(part.elem_count > 0, DJETS_ERR_INVALID,
            "ArgumentError: reducing over an empty collection is not allowed (a worker's local part is empty)");
}

// `as_norm`: the partial of an empty local part is norm(empty) = 0 (Julia), whatever p; otherwise minimum /
// maximum over an empty part is an error.
void reduce_parts(djets_vec x, djets_vec y, int kind, double param, double* out_parts, bool as_norm = false) {
  djets_ctx ctx = x->ctx;
  no_sync_in_capture(ctx, "a reduction returning its value to the host");
  settle(x);
  settle(y);
  if (!as_norm) check_extrema_parts(x, kind);
  const int np = (int)x->parts.size();
  // Every process of a multi-process job makes the same calls, so the decision to exchange must not
  // depend on which parts happen to be local: with NCCL up, always all-reduce.
  const bool multi_process = (int)ctx->ranks.size() < ctx->world;
  const bool exchange = multi_process && ctx->has_nccl;
  const bool host_exchange = multi_process && !ctx->has_nccl && ctx->host_allgather != nullptr;
  // ranks that take part: the owners of x's parts; with an exchange every rank this process drives
  std::vector<RankCtx*> involved;
  if (exchange) {
    for (RankCtx& rc : ctx->ranks) involved.push_back(&rc);
  } else {
    for (const VecPart& part : x->parts)
      if (RankCtx* rc = ctx->local(part.rank))
        if (std::find(involved.begin(), involved.end(), rc) == involved.end()) involved.push_back(rc);
  }
  if (exchange)
    for (RankCtx* rc : involved) {  // zeros in the slots of parts owned elsewhere: the sum is then exact
      use(*rc);
      CK(cudaMemsetAsync(rc->d_result, 0, sizeof(double) * np, rc->stream));
    }
  bool any_remote = false;
  for (int p = 0; p < np; ++p) {
    VecPart& part = x->parts[p];
    RankCtx* rc = ctx->local(part.rank);
    if (!rc) {
      any_remote = true;
      continue;
    }
    use(*rc);
    const void* yp = y ? y->parts[p].d : nullptr;
    // without an exchange the kernel stores the part's result straight into mapped host memory: no copy to wait for
    double* res = exchange ? rc->d_result + p : rc->h_result_dev + p;
    launch_counted(ctx, *rc, DJETS_K_REDUCE, [&] {
      return launch_reduce(real_dtype(x->dtype), kind, part.d, yp, part.elem_count * (is_complex(x->dtype) ? 2 : 1), param,
                           rc->d_partials, rc->d_ticket, res, rc->stream);
    });
  }
  if (exchange) {
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
    NCK(api, api->GroupStart());
    for (RankCtx* rc : involved) {
      use(*rc);
      NCK(api, api->AllReduce(rc->d_result, rc->d_result, np, ncclDouble, ncclSum, rc->comm, rc->stream));
    }
    NCK(api, api->GroupEnd());
  } else {
    REQUIRE(!any_remote || host_exchange, DJETS_ERR_NCCL,
            "vector has parts in other processes but the context has neither NCCL nor a host all-gather");
  }
  if (exchange)
    for (RankCtx* rc : involved) {
      use(*rc);
      CK(cudaMemcpyAsync(rc->h_result, rc->d_result, sizeof(double) * np, cudaMemcpyDeviceToHost, rc->stream));
    }
  for (RankCtx* rc : involved) {
    use(*rc);
    CK(cudaStreamSynchronize(rc->stream));
  }
  if (host_exchange) {  // the P partials travel through the caller's control plane: one contributor per slot
    std::vector<double> mine((size_t)np, 0.0), all((size_t)np * (size_t)ctx->world, 0.0);
    for (int p = 0; p < np; ++p)
      if (RankCtx* rc = ctx->local(x->parts[p].rank)) mine[(size_t)p] = rc->h_result[p];
    control_allgather(ctx, mine.data(), all.data(), sizeof(double) * (size_t)np);
    for (int p = 0; p < np; ++p) {
      double v = 0.0;
      for (int r = 0; r < ctx->world; ++r) v += all[(size_t)r * np + p];
      out_parts[p] = v;
      if (as_norm && x->parts[p].elem_count == 0) out_parts[p] = 0.0;
    }
    return;
  }
  for (int p = 0; p < np; ++p) {
    RankCtx* rc = ctx->local(x->parts[p].rank);
    out_parts[p] = rc ? rc->h_result[p] : involved.front()->h_result[p];
    if (as_norm && x->parts[p].elem_count == 0) out_parts[p] = 0.0;
  }
}

// master-side fold over the per-worker partials in the element type (src:222, 236, 249, 262, 276)
double fold_parts(int dtype, int kind, const double* z, int np) {
  if (kind_is_min(kind)) {
    double m = INFINITY;
    for (int p = 0; p < np; ++p) {
      const double v = round_to(dtype, z[p]);
      m = (m != m || v != v) ? NAN : fmin(m, v);  // Julia's minimum / maximum propagate NaN
    }
    return m;
  }
  if (kind_is_max(kind)) {
    double m = -INFINITY;
    for (int p = 0; p < np; ++p) {
      const double v = round_to(dtype, z[p]);
      m = (m != m || v != v) ? NAN : fmax(m, v);
    }
    return m;
  }
  double s = 0.0;
  for (int p = 0; p < np; ++p) s = round_to(dtype, s + round_to(dtype, z[p]));
  return s;
}


// Per-part partials folded ON THE DEVICE into a device-resident scalar (no host round trip): the partials of
// all parts are gathered on the rank that owns part 0 (peer stores straight from the reduction kernels, or the
// all-reduce buffer when parts live in other processes), one thread folds them with the reference's rules and
// stores the result into every local rank's replica of the scalar.
void reduce_to_scalar(djets_vec x, djets_vec y, int kind, double param, int fold_mode, double fold_p, djets_scalar out,
                      bool as_norm = false) {
  djets_ctx ctx = x->ctx;
  REQUIRE(!is_complex(x->dtype), DJETS_ERR_UNSUPPORTED, "device-resident reductions: real element types only");
  REQUIRE(out && out->ctx == ctx, DJETS_ERR_INVALID, "reduction into a scalar of another context");
  settle(x);
  settle(y);
  if (!as_norm) check_extrema_parts(x, kind);
  // partial of one part into `res` on its owner's stream: the reduction kernel, or norm(empty) = 0
  auto partial = [&](RankCtx& rc, int p, double* res) {
    const void* yp = y ? y->parts[p].d : nullptr;
    launch_counted(ctx, rc, DJETS_K_REDUCE, [&] {
      if (as_norm && x->parts[p].elem_count == 0) return launch_scalar_set(0.0, res, rc.stream);
      return launch_reduce(real_dtype(x->dtype), kind, x->parts[p].d, yp,
                           x->parts[p].elem_count * (is_complex(x->dtype) ? 2 : 1), param, rc.d_partials, rc.d_ticket, res,
                           rc.stream);
    });
  };
  const int np = (int)x->parts.size();
  const bool multi_process = (int)ctx->ranks.size() < ctx->world;
  if (multi_process && !ctx->has_nccl) {
    // no NCCL: the partials go through the caller's control plane (this synchronises), the fold still runs on the device
    REQUIRE(ctx->host_allgather != nullptr, DJETS_ERR_NCCL,
            "vector has parts in other processes but the context has neither NCCL nor a host all-gather");
    no_sync_in_capture(ctx, "a reduction over parts in other processes without NCCL");
    std::vector<double> z((size_t)np, 0.0);
    reduce_parts(x, y, kind, param, z.data(), as_norm);
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      memcpy(rc.h_result + kMaxParts, z.data(), sizeof(double) * (size_t)np);
      CK(cudaMemcpyAsync(rc.d_result, rc.h_result + kMaxParts, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, rc.stream));
      launch_counted(ctx, rc, DJETS_K_REDUCE, [&] {
        return launch_fold(x->dtype, fold_mode, fold_p, rc.d_result, np, rc.d_scalars + out->slot, ScalarPeers{{}, 0},
                           rc.stream);
      });
      CK(cudaStreamSynchronize(rc.stream));  // h_result is reused by the next reduction
    }
    return;
  }
  if (multi_process) {
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      CK(cudaMemsetAsync(rc.d_result, 0, sizeof(double) * np, rc.stream));
    }
    for (int p = 0; p < np; ++p) {
      RankCtx* rc = ctx->local(x->parts[p].rank);
      if (!rc) continue;
      use(*rc);
      partial(*rc, p, rc->d_result + p);
    }
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
    NCK(api, api->GroupStart());
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      NCK(api, api->AllReduce(rc.d_result, rc.d_result, np, ncclDouble, ncclSum, rc.comm, rc.stream));
    }
    NCK(api, api->GroupEnd());
    for (RankCtx& rc : ctx->ranks) {
      use(rc);
      launch_counted(ctx, rc, DJETS_K_REDUCE, [&] {
        return launch_fold(x->dtype, fold_mode, fold_p, rc.d_result, np, rc.d_scalars + out->slot, ScalarPeers{{}, 0},
                           rc.stream);
      });
    }
    return;
  }
  RankCtx* root = ctx->local(x->parts[0].rank);
  REQUIRE(root, DJETS_ERR_INVALID, "reduce: part 0 is not local");
  // gather: each owner's reduction kernel stores its partial into the root's gather buffer (in place over NVLink
  // when the devices can address each other, a peer copy otherwise)
  for (int p = 0; p < np; ++p) {
    RankCtx* rc = ctx->local(x->parts[p].rank);
    const bool direct = ctx->can_address(*rc, *root);
    if (rc != root) stream_after(*root, *rc);  // the root's previous fold has read the gather buffer
    use(*rc);
    double* res = direct ? root->d_gather + p : rc->d_result + p;
    partial(*rc, p, res);
    if (!direct)
      CK(cudaMemcpyPeerAsync(root->d_gather + p, root->device, res, rc->device, sizeof(double), rc->stream));
    if (rc != root) stream_after(*rc, *root);
  }
  // fold on the root, replicas on the other local ranks
  ScalarPeers peers{{}, 0};
  std::vector<RankCtx*> later;
  for (RankCtx& rc : ctx->ranks) {
    if (&rc == root) continue;
    stream_after(rc, *root);  // earlier readers of the replica are done
    if (ctx->can_address(*root, rc) && peers.n < 16)
      peers.p[peers.n++] = rc.d_scalars + out->slot;
    else
      later.push_back(&rc);
  }
  use(*root);
  launch_counted(ctx, *root, DJETS_K_REDUCE, [&] {
    return launch_fold(x->dtype, fold_mode, fold_p, root->d_gather, np, root->d_scalars + out->slot, peers, root->stream);
  });
  for (RankCtx* rc : later)
    CK(cudaMemcpyPeerAsync(rc->d_scalars + out->slot, rc->device, root->d_scalars + out->slot, root->device,
                           sizeof(double), root->stream));
  for (RankCtx& rc : ctx->ranks)
    if (&rc != root) stream_after(*root, rc);
}

// scatter_async / collect_async run on the owner's copy stream: they start once everything enqueued so far on the
// main stream has finished (the vector's earlier readers and writers) and overlap whatever is enqueued afterwards
// that does not touch this vector; the next call that does touch it waits for the copy (VecPart::pending).
void bulk_copy_async(djets_vec x, void* host, bool to_device) {
  REQUIRE(x, DJETS_ERR_INVALID, "null vector");
  REQUIRE(host || x->length == 0, DJETS_ERR_INVALID, "null host pointer");
  REQUIRE(!x->ctx->capturing, DJETS_ERR_INVALID, "host transfers cannot be part of a captured sequence");
  settle(x);
  const int es = elem_size(x->dtype);
  for (VecPart& p : x->parts) {
    RankCtx* rc = x->ctx->local(p.rank);
    if (!rc || p.elem_count == 0) continue;
    use(*rc);
    CK(cudaEventRecord(rc->ev_c0, rc->stream));
    CK(cudaStreamWaitEvent(rc->cstream, rc->ev_c0, 0));
    char* h = reinterpret_cast<char*>(host) + (size_t)p.elem_first * es;
    if (to_device)
      CK(cudaMemcpyAsync(p.d, h, (size_t)p.elem_count * es, cudaMemcpyHostToDevice, rc->cstream));
    else
      CK(cudaMemcpyAsync(h, p.d, (size_t)p.elem_count * es, cudaMemcpyDeviceToHost, rc->cstream));
    if (!p.pending) CK(cudaEventCreateWithFlags(&p.pending, cudaEventDisableTiming));
    CK(cudaEventRecord(p.pending, rc->cstream));
    p.has_pending = true;
  }
}

}  // namespace

extern "C" {

int32_t djets_vec_create(djets_ctx ctx, int32_t dtype, int32_t nparts, const int32_t* part_rank,
                         const int64_t* part_nblocks, const int64_t* block_len, djets_vec* out) {
  return guard([&] {
    REQUIRE(out && part_rank && part_nblocks && block_len, DJETS_ERR_INVALID, "vec_create: null argument");
    *out = make_vec(ctx, dtype, nparts, part_rank, part_nblocks, block_len);
  });
}

int32_t djets_vec_similar(djets_vec x, int32_t dtype, djets_vec* out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "vec_similar: null argument");
    std::vector<int32_t> ranks;
    std::vector<int64_t> nb;
    for (VecPart& p : x->parts) {
      ranks.push_back(p.rank);
      nb.push_back(p.blk_count);
    }
    *out = make_vec(x->ctx, dtype, (int)x->parts.size(), ranks.data(), nb.data(), x->block_len.data());
  });
}

int32_t djets_vec_destroy(djets_vec x) {
  return guard([&] {
    if (!x) return;
    free_vec(x);
  });
}

int32_t djets_vec_info(djets_vec x, int32_t* dtype, int64_t* length, int64_t* nblocks, int32_t* nparts) {
  return guard([&] {
    REQUIRE(x, DJETS_ERR_INVALID, "null vector");
    if (dtype) *dtype = x->dtype;
    if (length) *length = x->length;
    if (nblocks) *nblocks = (int64_t)x->block_len.size();
    if (nparts) *nparts = (int32_t)x->parts.size();
  });
}

int32_t djets_vec_part_info(djets_vec x, int32_t part, int32_t* rank, int64_t* blk_first, int64_t* blk_count,
                            int64_t* elem_first, int64_t* elem_count) {
  return guard([&] {
    REQUIRE(x && part >= 0 && part < (int)x->parts.size(), DJETS_ERR_INVALID, "part_info: bad part");
    const VecPart& p = x->parts[part];
    if (rank) *rank = p.rank;
    if (blk_first) *blk_first = p.blk_first;
    if (blk_count) *blk_count = p.blk_count;
    if (elem_first) *elem_first = p.elem_first;
    if (elem_count) *elem_count = p.elem_count;
  });
}

int32_t djets_vec_block_info(djets_vec x, int64_t iblock, int32_t* part, int64_t* elem_first, int64_t* len) {
  return guard([&] {
    REQUIRE(x && iblock >= 0 && iblock < (int64_t)x->block_len.size(), DJETS_ERR_INVALID, "block index out of range");
    if (part) *part = x->block_part[iblock];
    if (elem_first) *elem_first = x->block_off[iblock];
    if (len) *len = x->block_len[iblock];
  });
}

static void block_copy(djets_vec x, int64_t iblock, void* host, bool to_device) {
  REQUIRE(x && iblock >= 0 && iblock < (int64_t)x->block_len.size(), DJETS_ERR_INVALID, "block index out of range");
  no_sync_in_capture(x->ctx, "getblock / setblock!");
  settle(x);
  VecPart& p = x->parts[x->block_part[iblock]];
  RankCtx* rc = x->ctx->local(p.rank);
  if (!rc) {
    if (to_device) return;  // setblock! on a process that does not drive the owner is a no-op
    fail(DJETS_ERR_NOT_LOCAL, "getblock: block is owned by a rank of another process");
  }
  const int es = elem_size(x->dtype);
  const size_t bytes = (size_t)x->block_len[iblock] * es;
  if (bytes == 0) return;
  REQUIRE(host, DJETS_ERR_INVALID, "null host pointer");
  use(*rc);
  char* d = p.d + (size_t)(x->block_off[iblock] - p.elem_first) * es;
  if (to_device)
    CK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, rc->stream));
  else
    CK(cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, rc->stream));
  CK(cudaStreamSynchronize(rc->stream));
}

int32_t djets_vec_setblock(djets_vec x, int64_t iblock, const void* host_src) {
  return guard([&] { block_copy(x, iblock, const_cast<void*>(host_src), true); });
}
int32_t djets_vec_getblock(djets_vec x, int64_t iblock, void* host_dst) {
  return guard([&] { block_copy(x, iblock, host_dst, false); });
}

static void getindex2(djets_vec x, int64_t i, double* out2);

int32_t djets_vec_getindex_c(djets_vec x, int64_t i, double* out2) {
  return guard([&] {
    REQUIRE(x && out2, DJETS_ERR_INVALID, "null argument");
    getindex2(x, i, out2);
  });
}

int32_t djets_vec_getindex(djets_vec x, int64_t i, double* out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "null argument");
    if (is_complex(x->dtype)) {
      double z[2];
      getindex2(x, i, z);
      REQUIRE(z[1] == 0.0, DJETS_ERR_UNSUPPORTED, "getindex: complex element (use djets_vec_getindex_c)");
      *out = z[0];
      return;
    }
    REQUIRE(i >= 0 && i < x->length, DJETS_ERR_INVALID, "BoundsError: index outside the DBArray");
    settle(x);
    for (VecPart& p : x->parts) {
      if (i < p.elem_first || i >= p.elem_first + p.elem_count) continue;  // findfirst(rng->i in rng) src:175
      RankCtx* rc = x->ctx->local(p.rank);
      REQUIRE(rc, DJETS_ERR_NOT_LOCAL, "getindex: element is owned by a rank of another process");
      use(*rc);
      const int es = elem_size(x->dtype);
      double buf = 0.0;
      CK(cudaMemcpyAsync(&buf, p.d + (size_t)(i - p.elem_first) * es, es, cudaMemcpyDeviceToHost, rc->stream));
      CK(cudaStreamSynchronize(rc->stream));
      if (x->dtype == DJETS_F32) {
        float f;
        memcpy(&f, &buf, 4);
        *out = (double)f;
      } else {
        *out = buf;
      }
      return;
    }
    fail(DJETS_ERR_INVALID, "getindex: no part holds the index");
  });
}

static void bulk_copy(djets_vec x, void* host, bool to_device) {
  REQUIRE(x, DJETS_ERR_INVALID, "null vector");
  no_sync_in_capture(x->ctx, "collect / scatter");
  settle(x);
  REQUIRE(host || x->length == 0, DJETS_ERR_INVALID, "null host pointer");
  const int es = elem_size(x->dtype);
  for (VecPart& p : x->parts) {
    RankCtx* rc = x->ctx->local(p.rank);
    if (!rc || p.elem_count == 0) continue;
    use(*rc);
    char* h = reinterpret_cast<char*>(host) + (size_t)p.elem_first * es;
    if (to_device)
      CK(cudaMemcpyAsync(p.d, h, (size_t)p.elem_count * es, cudaMemcpyHostToDevice, rc->stream));
    else
      CK(cudaMemcpyAsync(h, p.d, (size_t)p.elem_count * es, cudaMemcpyDeviceToHost, rc->stream));
  }
  for (VecPart& p : x->parts)
    if (RankCtx* rc = x->ctx->local(p.rank)) {
      use(*rc);
      CK(cudaStreamSynchronize(rc->stream));
    }
}

static void getindex2(djets_vec x, int64_t i, double* out2) {
  REQUIRE(i >= 0 && i < x->length, DJETS_ERR_INVALID, "BoundsError: index outside the DBArray");
  no_sync_in_capture(x->ctx, "getindex");
  settle(x);
  for (VecPart& p : x->parts) {
    if (i < p.elem_first || i >= p.elem_first + p.elem_count) continue;
    RankCtx* rc = x->ctx->local(p.rank);
    REQUIRE(rc, DJETS_ERR_NOT_LOCAL, "getindex: element is owned by a rank of another process");
    use(*rc);
    const int es = elem_size(x->dtype);
    double buf[2] = {0.0, 0.0};
    CK(cudaMemcpyAsync(buf, p.d + (size_t)(i - p.elem_first) * es, es, cudaMemcpyDeviceToHost, rc->stream));
    CK(cudaStreamSynchronize(rc->stream));
    if (real_dtype(x->dtype) == DJETS_F32) {
      float f[2];
      memcpy(f, buf, 8);
      out2[0] = f[0];
      out2[1] = is_complex(x->dtype) ? f[1] : 0.0;
    } else {
      out2[0] = buf[0];
      out2[1] = is_complex(x->dtype) ? buf[1] : 0.0;
    }
    return;
  }
  fail(DJETS_ERR_INVALID, "getindex: no part holds the index");
}

int32_t djets_vec_collect(djets_vec x, void* host_dst) {
  return guard([&] { bulk_copy(x, host_dst, false); });
}
int32_t djets_vec_scatter(djets_vec x, const void* host_src) {
  return guard([&] { bulk_copy(x, const_cast<void*>(host_src), true); });
}

int32_t djets_vec_fill(djets_vec x, double a) {
  return guard([&] {
    REQUIRE(x, DJETS_ERR_INVALID, "null vector");
    settle(x);
    for (VecPart& p : x->parts) {
      RankCtx* rc = x->ctx->local(p.rank);
      if (!rc) continue;
      use(*rc);
      launch_counted(x->ctx, *rc, DJETS_K_ELTWISE, [&] {
        return is_complex(x->dtype) ? launch_fill_complex(x->dtype, p.d, a, 0.0, p.elem_count, rc->stream)
                                    : launch_fill(x->dtype, p.d, a, p.elem_count, rc->stream);
      });
    }
  });
}

int32_t djets_vec_fill_c(djets_vec x, double re, double im) {
  return guard([&] {
    REQUIRE(x, DJETS_ERR_INVALID, "null vector");
    REQUIRE(is_complex(x->dtype) || im == 0.0, DJETS_ERR_MISMATCH, "InexactError: complex value into a real DBArray");
    settle(x);
    for (VecPart& p : x->parts) {
      RankCtx* rc = x->ctx->local(p.rank);
      if (!rc) continue;
      use(*rc);
      launch_counted(x->ctx, *rc, DJETS_K_ELTWISE, [&] {
        return is_complex(x->dtype) ? launch_fill_complex(x->dtype, p.d, re, im, p.elem_count, rc->stream)
                                    : launch_fill(x->dtype, p.d, re, p.elem_count, rc->stream);
      });
    }
  });
}

int32_t djets_vec_lincomb(djets_vec dest, int32_t nterms, const double* coef, const djets_vec* xs, int32_t flags) {
  return guard([&] {
    REQUIRE(dest && coef && xs, DJETS_ERR_INVALID, "lincomb: null argument");
    REQUIRE(nterms >= 1 && nterms <= 4, DJETS_ERR_UNSUPPORTED, "lincomb: 1 to 4 terms");
    for (int k = 0; k < nterms; ++k) {
      REQUIRE(xs[k], DJETS_ERR_INVALID, "lincomb: null operand");
      REQUIRE(same_layout(dest, xs[k]), DJETS_ERR_MISMATCH,
              "DimensionMismatch: broadcast operands have different block layouts");
      settle(xs[k]);
      REQUIRE(is_complex(xs[k]->dtype) == is_complex(dest->dtype), DJETS_ERR_UNSUPPORTED,
              "broadcast between real and complex DBArrays is outside the closed set");
    }
    settle(dest);
    // complex vectors with real coefficients: the same kernels over twice as many reals
    const int cs = is_complex(dest->dtype) ? 2 : 1;
    for (size_t p = 0; p < dest->parts.size(); ++p) {
      VecPart& dp = dest->parts[p];
      RankCtx* rc = dest->ctx->local(dp.rank);
      if (!rc) continue;
      use(*rc);
      LincombArgs a;
      a.nterms = nterms;
      for (int k = 0; k < nterms; ++k) {
        a.x[k] = xs[k]->parts[p].d;
        a.coef[k] = coef[k];
        a.xdtype[k] = real_dtype(xs[k]->dtype);
      }
      launch_counted(dest->ctx, *rc, DJETS_K_ELTWISE, [&] {
        return launch_lincomb(real_dtype(dest->dtype), dp.d, a, dp.elem_count * cs, flags & 1, rc->stream);
      });
    }
  });
}

int32_t djets_vec_binary(djets_vec dest, int32_t op, djets_vec a, djets_vec b) {
  return guard([&] {
    REQUIRE(dest && a && b, DJETS_ERR_INVALID, "binary: null argument");
    REQUIRE(op == 0 || op == 1, DJETS_ERR_UNSUPPORTED, "binary: op 0 (.*) or 1 (./)");
    REQUIRE(same_layout(dest, a) && same_layout(dest, b), DJETS_ERR_MISMATCH,
            "DimensionMismatch: broadcast operands have different block layouts");
    REQUIRE(dest->dtype == a->dtype && dest->dtype == b->dtype, DJETS_ERR_UNSUPPORTED,
            "binary: operands must share the element type");
    REQUIRE(!is_complex(dest->dtype), DJETS_ERR_UNSUPPORTED, "binary: real element types only");
    settle(dest);
    settle(a);
    settle(b);
    for (size_t p = 0; p < dest->parts.size(); ++p) {
      VecPart& dp = dest->parts[p];
      RankCtx* rc = dest->ctx->local(dp.rank);
      if (!rc) continue;
      use(*rc);
      launch_counted(dest->ctx, *rc, DJETS_K_ELTWISE, [&] {
        return launch_binary(dest->dtype, op, dp.d, a->parts[p].d, b->parts[p].d, dp.elem_count, rc->stream);
      });
    }
  });
}

int32_t djets_vec_reduce_parts(djets_vec x, int32_t kind, double param, double* out_parts) {
  return guard([&] {
    REQUIRE(x && out_parts, DJETS_ERR_INVALID, "reduce: null argument");
    REQUIRE(kind >= 0 && kind <= DJETS_RED_SUMSQ_SHIFT, DJETS_ERR_INVALID, "reduce: unknown kind");
    REQUIRE(!is_complex(x->dtype), DJETS_ERR_UNSUPPORTED, "reduce: real element types only (complex: dot, norm)");
    reduce_parts(x, nullptr, kind, param, out_parts);
  });
}

int32_t djets_vec_reduce(djets_vec x, int32_t kind, double param, double* out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "reduce: null argument");
    REQUIRE(kind >= 0 && kind <= DJETS_RED_SUMSQ_SHIFT, DJETS_ERR_INVALID, "reduce: unknown kind");
    REQUIRE(!is_complex(x->dtype), DJETS_ERR_UNSUPPORTED, "reduce: real element types only (complex: dot, norm)");
    std::vector<double> z(x->parts.size());
    reduce_parts(x, nullptr, kind, param, z.data());
    *out = fold_parts(x->dtype, kind, z.data(), (int)z.size());
  });
}

int32_t djets_vec_dot_c(djets_vec x, djets_vec y, double* out2) {
  return guard([&] {
    REQUIRE(x && y && out2, DJETS_ERR_INVALID, "dot: null argument");
    REQUIRE(same_layout(x, y), DJETS_ERR_MISMATCH, "DimensionMismatch: dot of DBArrays with different block layouts");
    REQUIRE(x->dtype == y->dtype, DJETS_ERR_MISMATCH, "dot: element types differ (reference: MethodError, src:214)");
    std::vector<double> z(x->parts.size());
    reduce_parts(x, y, RED_DOT, 0.0, z.data());  // over the (re, im) reals: Re sum conj(x) y
    out2[0] = fold_parts(x->dtype, DJETS_RED_SUM, z.data(), (int)z.size());
    out2[1] = 0.0;
    if (is_complex(x->dtype)) {
      reduce_parts(x, y, RED_CDOT_IM, 0.0, z.data());
      out2[1] = fold_parts(x->dtype, DJETS_RED_SUM, z.data(), (int)z.size());
    }
  });
}

int32_t djets_vec_dot(djets_vec x, djets_vec y, double* out) {
  return guard([&] {
    REQUIRE(x && y && out, DJETS_ERR_INVALID, "dot: null argument");
    REQUIRE(!is_complex(x->dtype), DJETS_ERR_UNSUPPORTED, "dot: complex DBArrays return a complex value (djets_vec_dot_c)");
    REQUIRE(same_layout(x, y), DJETS_ERR_MISMATCH, "DimensionMismatch: dot of DBArrays with different block layouts");
    REQUIRE(x->dtype == y->dtype, DJETS_ERR_MISMATCH, "dot: element types differ (reference: MethodError, src:214)");
    std::vector<double> z(x->parts.size());
    reduce_parts(x, y, RED_DOT, 0.0, z.data());
    *out = fold_parts(x->dtype, DJETS_RED_SUM, z.data(), (int)z.size());
  });
}

int32_t djets_vec_norm(djets_vec x, double p, double* out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "norm: null argument");
    const int np = (int)x->parts.size();
    const int dt = x->dtype;
    std::vector<double> z(np);
    if (is_complex(dt)) {  // |z| per element (Julia's norm of a complex array); same folds
      if (p == INFINITY) {
        reduce_parts(x, nullptr, RED_CABSMAX, 0.0, z.data(), true);
        *out = fold_parts(dt, DJETS_RED_MAX, z.data(), np);
      } else if (p == -INFINITY) {
        reduce_parts(x, nullptr, RED_CABSMIN, 0.0, z.data(), true);
        *out = fold_parts(dt, DJETS_RED_MIN, z.data(), np);
      } else if (p == 0.0 || p == 1.0) {
        reduce_parts(x, nullptr, p == 0.0 ? RED_CNNZ : RED_CABSSUM, 0.0, z.data(), true);
        *out = fold_parts(dt, DJETS_RED_SUM, z.data(), np);
      } else {
        const double pp = round_to(dt, p);
        reduce_parts(x, nullptr, p == 2.0 ? (int)DJETS_RED_SUMSQ : (int)RED_CSUMPOW, pp, z.data(), true);
        double s = 0.0;
        for (int k = 0; k < np; ++k) {
          const double zk = round_to(dt, p == 2.0 ? sqrt(z[k]) : pow(z[k], 1.0 / pp));
          s = round_to(dt, s + round_to(dt, p == 2.0 ? zk * zk : pow(zk, pp)));
        }
        *out = round_to(dt, p == 2.0 ? sqrt(s) : pow(s, round_to(dt, 1.0 / pp)));
      }
      return;
    }
    if (p == INFINITY) {  // src:200-201
      reduce_parts(x, nullptr, DJETS_RED_ABSMAX, 0.0, z.data(), true);
      *out = fold_parts(dt, DJETS_RED_MAX, z.data(), np);
    } else if (p == -INFINITY) {  // src:202-203
      reduce_parts(x, nullptr, DJETS_RED_ABSMIN, 0.0, z.data(), true);
      *out = fold_parts(dt, DJETS_RED_MIN, z.data(), np);
    } else if (p == 0.0 || p == 1.0) {  // src:204-205
      reduce_parts(x, nullptr, p == 0.0 ? DJETS_RED_NNZ : DJETS_RED_ABSSUM, 0.0, z.data());
      *out = fold_parts(dt, DJETS_RED_SUM, z.data(), np);
    } else {  // src:206-209: per-worker norms z, then (sum z^p)^(1/p) in float(real(T))
      const double pp = round_to(dt, p);
      reduce_parts(x, nullptr, p == 2.0 ? DJETS_RED_SUMSQ : DJETS_RED_SUMPOW, pp, z.data());
      double s = 0.0;
      for (int k = 0; k < np; ++k) {
        const double zk = round_to(dt, p == 2.0 ? sqrt(z[k]) : pow(z[k], 1.0 / pp));
        s = round_to(dt, s + round_to(dt, p == 2.0 ? zk * zk : pow(zk, pp)));
      }
      *out = round_to(dt, p == 2.0 ? sqrt(s) : pow(s, round_to(dt, 1.0 / pp)));
    }
  });
}


// ---- stream-ordered forms: nothing here waits for the device -------------------------------------------------
int32_t djets_vec_scatter_async(djets_vec x, const void* host_src) {
  return guard([&] { bulk_copy_async(x, const_cast<void*>(host_src), true); });
}
int32_t djets_vec_collect_async(djets_vec x, void* host_dst) {
  return guard([&] { bulk_copy_async(x, host_dst, false); });
}

int32_t djets_vec_dot_s(djets_vec x, djets_vec y, djets_scalar out) {
  return guard([&] {
    REQUIRE(x && y && out, DJETS_ERR_INVALID, "dot: null argument");
    REQUIRE(same_layout(x, y), DJETS_ERR_MISMATCH, "DimensionMismatch: dot of DBArrays with different block layouts");
    REQUIRE(x->dtype == y->dtype, DJETS_ERR_MISMATCH, "dot: element types differ (reference: MethodError, src:214)");
    reduce_to_scalar(x, y, RED_DOT, 0.0, FOLD_SUM, 0.0, out);
  });
}

int32_t djets_vec_norm_s(djets_vec x, double p, djets_scalar out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "norm: null argument");
    if (p == INFINITY)
      reduce_to_scalar(x, nullptr, DJETS_RED_ABSMAX, 0.0, FOLD_MAX, 0.0, out, true);
    else if (p == -INFINITY)
      reduce_to_scalar(x, nullptr, DJETS_RED_ABSMIN, 0.0, FOLD_MIN, 0.0, out, true);
    else if (p == 0.0 || p == 1.0)
      reduce_to_scalar(x, nullptr, p == 0.0 ? DJETS_RED_NNZ : DJETS_RED_ABSSUM, 0.0, FOLD_SUM, 0.0, out);
    else if (p == 2.0)
      reduce_to_scalar(x, nullptr, DJETS_RED_SUMSQ, 0.0, FOLD_NORM2, 2.0, out);
    else
      reduce_to_scalar(x, nullptr, DJETS_RED_SUMPOW, round_to(x->dtype, p), FOLD_NORMP, p, out);
  });
}

int32_t djets_vec_reduce_s(djets_vec x, int32_t kind, double param, djets_scalar out) {
  return guard([&] {
    REQUIRE(x && out, DJETS_ERR_INVALID, "reduce: null argument");
    REQUIRE(kind >= 0 && kind <= DJETS_RED_SUMSQ_SHIFT, DJETS_ERR_INVALID, "reduce: unknown kind");
    reduce_to_scalar(x, nullptr, kind, param, kind_is_min(kind) ? FOLD_MIN : (kind_is_max(kind) ? FOLD_MAX : FOLD_SUM),
                     0.0, out);
  });
}

int32_t djets_plan_eval(const uint8_t* program, int32_t nops, int32_t ninputs, int32_t nscalars, int32_t* depth,
                        int32_t* nops_fused) {
  return guard([&] {
    REQUIRE(program && nops >= 1 && nops <= kEvalMaxOps, DJETS_ERR_UNSUPPORTED, "eval: 1 to 48 operations per program");
    REQUIRE(ninputs >= 0 && ninputs <= kEvalMaxIn && nscalars >= 0 && nscalars <= kEvalMaxScalars, DJETS_ERR_UNSUPPORTED,
            "eval: at most 8 input vectors and 16 scalars");
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.nops = nops;
    a.nin = ninputs;
    a.nsc = nscalars;
    for (int k = 0; k < nops; ++k) {
      a.op[k] = program[2 * k];
      a.arg[k] = program[2 * k + 1];
    }
    const char* why = nullptr;
    const int d0 = eval_check(a, &why);
    REQUIRE(d0 >= 0, DJETS_ERR_UNSUPPORTED, why ? why : "eval: malformed program");
    EvalArgs b = a;
    b.nops = eval_peephole(a, b.op, b.arg);
    const int d1 = eval_check(b, &why);
    REQUIRE(d1 >= 0, DJETS_ERR_INVALID, "eval: the peephole pass produced a malformed program");
    if (depth) *depth = d1;
    if (nops_fused) *nops_fused = b.nops;
  });
}

// ---- the closed elementwise program (src:363-375) --------------------------------------------------------------
int32_t djets_vec_eval(djets_vec dest, const uint8_t* program, int32_t nops, const djets_vec* inputs, int32_t ninputs,
                       const double* scalars, const djets_scalar* dscalars, int32_t nscalars, int32_t flags) {
  return guard([&] {
    REQUIRE(dest && program, DJETS_ERR_INVALID, "eval: null argument");
    REQUIRE(nops >= 1 && nops <= kEvalMaxOps, DJETS_ERR_UNSUPPORTED, "eval: 1 to 48 operations per program");
    REQUIRE(ninputs >= 0 && ninputs <= kEvalMaxIn, DJETS_ERR_UNSUPPORTED, "eval: at most 8 input vectors");
    REQUIRE(nscalars >= 0 && nscalars <= kEvalMaxScalars, DJETS_ERR_UNSUPPORTED, "eval: at most 16 scalars");
    REQUIRE(ninputs == 0 || inputs, DJETS_ERR_INVALID, "eval: null input list");
    REQUIRE(nscalars == 0 || scalars || dscalars, DJETS_ERR_INVALID, "eval: null scalar list");
    for (int k = 0; k < ninputs; ++k) {
      REQUIRE(inputs[k], DJETS_ERR_INVALID, "eval: null operand");
      REQUIRE(same_layout(dest, inputs[k]), DJETS_ERR_MISMATCH,
              "DimensionMismatch: broadcast operands have different block layouts");
      settle(inputs[k]);
      REQUIRE(is_complex(inputs[k]->dtype) == is_complex(dest->dtype), DJETS_ERR_UNSUPPORTED,
              "broadcast between real and complex DBArrays is outside the closed set");
    }
    settle(dest);
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.nops = nops;
    for (int k = 0; k < nops; ++k) {
      a.op[k] = program[2 * k];
      a.arg[k] = program[2 * k + 1];
    }
    a.nin = ninputs;
    for (int k = 0; k < ninputs; ++k) a.in_dtype[k] = inputs[k]->dtype;
    a.nsc = nscalars;
    for (int k = 0; k < nscalars; ++k) {
      a.sc[k] = scalars ? scalars[k] : 0.0;
      REQUIRE(!(dscalars && dscalars[k]) || dscalars[k]->ctx == dest->ctx, DJETS_ERR_INVALID, "eval: scalar of another context");
    }
    const char* why = nullptr;
    REQUIRE(eval_check(a, &why) >= 0, DJETS_ERR_UNSUPPORTED, why ? why : "eval: malformed program");
    // Linear combinations `c0*x0 + c1*x1 - c2*x2 ...` of up to four vectors with host scalars (the forms the reference's
    // tests and benchmark use: test:322, bench:43) go to the specialised kernel: same products, same sums, same rounding.
    {
      double coef[4];
      int which[4], nt = 0, pc = 0;
      bool linear = true;
      while (linear && pc < nops) {
        if (a.op[pc] != EV_IN || nt == 4) {
          linear = false;
          break;
        }
        which[nt] = a.arg[pc++];
        coef[nt] = 1.0;
        if (pc < nops && (a.op[pc] == EV_MULS || a.op[pc] == EV_RMULS)) {
          const int j = a.arg[pc++];
          if (dscalars && dscalars[j]) linear = false;
          coef[nt] = a.sc[j];
        }
        if (nt > 0) {
          if (pc < nops && a.op[pc] == EV_ADD) {
            ++pc;
          } else if (pc < nops && a.op[pc] == EV_SUB) {
            coef[nt] = -coef[nt];
            ++pc;
          } else {
            linear = false;
          }
        }
        ++nt;
      }
      if (linear && nt >= 1) {
        for (size_t p = 0; p < dest->parts.size(); ++p) {
          VecPart& dp = dest->parts[p];
          RankCtx* rc = dest->ctx->local(dp.rank);
          if (!rc) continue;
          use(*rc);
          LincombArgs la;
          la.nterms = nt;
          for (int k = 0; k < nt; ++k) {
            la.x[k] = inputs[which[k]]->parts[p].d;
            la.coef[k] = coef[k];
            la.xdtype[k] = real_dtype(inputs[which[k]]->dtype);
          }
          const int cs = is_complex(dest->dtype) ? 2 : 1;
          launch_counted(dest->ctx, *rc, DJETS_K_ELTWISE, [&] {
            return launch_lincomb(real_dtype(dest->dtype), dp.d, la, dp.elem_count * cs, flags & 1, rc->stream);
          });
        }
        return;
      }
    }
    REQUIRE(!is_complex(dest->dtype), DJETS_ERR_UNSUPPORTED,
            "eval: complex DBArrays support linear combinations with real coefficients only");
    for (size_t p = 0; p < dest->parts.size(); ++p) {
      VecPart& dp = dest->parts[p];
      RankCtx* rc = dest->ctx->local(dp.rank);
      if (!rc) continue;
      use(*rc);
      for (int k = 0; k < ninputs; ++k) a.in[k] = inputs[k]->parts[p].d;
      for (int k = 0; k < nscalars; ++k) a.dsc[k] = (dscalars && dscalars[k]) ? rc->d_scalars + dscalars[k]->slot : nullptr;
      launch_counted(dest->ctx, *rc, DJETS_K_ELTWISE,
                     [&] { return launch_eval(dest->dtype, dp.d, a, dp.elem_count, flags & 1, rc->stream); });
    }
  });
}

}  // extern "C"
