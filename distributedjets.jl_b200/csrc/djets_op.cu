// The distributed block operator of libdjets_b200 (reference: src/DistributedJets.jl:414-721, 805-861):
// block payloads, tile schedules, the cross-rank exchange, CUDA-graph replay, and its C ABI.
#include "djets_internal.h"

using namespace djets;

// ---- distributed block operator -------------------------------------------------------------------
struct Block {
  int kind = DJETS_BLOCK_ZERO;
  char* d = nullptr;
  int64_t ld = 0;
};

struct DirPlan {
  DenseTile* d_tiles = nullptr;      // tiles in ascending block order ...
  DenseTile* d_tiles_rev = nullptr;  // ... and the same tiles in descending order (L2 hand-over, see build_plan)
  ChainLink* d_links = nullptr;      // further blocks of chained adjoint tiles (small blocks); nullptr: no chains
  int ntiles = 0;
  FinishItem* d_items = nullptr;  // loose items first (finish_kernel), then the fused ones
  int nitems = 0, nloose = 0;
  unsigned* d_tickets = nullptr;  // one per item
  unsigned* d_counter = nullptr;  // tile queue of the persistent kernels (zero between launches)
  unsigned* flags = nullptr;      // epoch flags of the cross-process exchange (see XP_* below)
  int64_t in_len = 0, part_len = 0;  // elements of the input part this cell reads / of the output part it contributes to
  int nremote = 0;
  DiagTerm* d_diags = nullptr;
  char* W = nullptr;
  char* stage_in = nullptr;   // input slice received from its owner
  char* stage_out = nullptr;  // reduced partial pushed to the owner
  char* R = nullptr;          // planes received from the other cells of the grid row / column
  int adj_ru = 4, adj_cw = 4;  // adjoint kernel shape chosen for this cell (rows unrolled, columns per warp)
  int fwd_tx = 64;             // forward kernel shape chosen for this cell (threads along the rows)
  int wave = 0;                // CTAs of this cell's tile kernel that are resident at once (first wave)
  bool chained = false;        // tiles are chains of small blocks (djets_blockmul_fwd/adj_chain_kernel)
  bool chain_pair = false;     // ... none taller than one warp-load: the adjoint loads two blocks per batch
};

struct Cell {
  int r = 0, c = 0, rank = 0;
  DirPlan fwd, adj;
  bool at_end = false;  // the cell's previous tile launch (either direction) ended on its LAST blocks: the next walks backwards
  cudaEvent_t pt0 = nullptr, pt1 = nullptr;  // perfstat: bracket of this cell's work in the last apply
  bool ptimed = false;
};

struct djets_op_s {
  djets_ctx ctx = nullptr;
  int dtype = 0;
  int64_t nbrow = 0, nbcol = 0;
  std::vector<int64_t> row_len, col_len;
  int n1 = 1, n2 = 1;
  std::vector<int> grid;  // column-major n1 x n2
  std::vector<int64_t> row_cuts, col_cuts;
  bool isdiag = false;
  std::vector<Block> blocks;  // nbrow x nbcol, row-major; payloads only for blocks of local cells
  std::vector<Cell> cells;    // column-major n1 x n2
  std::vector<int64_t> row_off, col_off;         // element offset of block row / column inside its part
  std::vector<int64_t> rowpart_len, colpart_len;  // elements per range / domain part
  bool finalized = false;
  int64_t p_fwd_tx = 64, p_adj_ru = 4, p_ld_policy = 1;
  // CUDA-graph replay of whole applies (latency-bound operators: many small launches and copies)
  struct GraphEntry {
    std::vector<const void*> key;  // device pointers of the input and output parts
    cudaGraphExec_t exec = nullptr;
    int64_t launches[DJETS_K_COUNT] = {0, 0, 0, 0, 0};
    uint64_t last_use = 0;
  };
  std::vector<GraphEntry> graphs[2];
  std::vector<std::vector<const void*>> seen[2];
  uint64_t graph_clock = 0;
  uint64_t graph_epoch = 0;  // context parameter epoch the cached graphs were captured under
  bool graphs_broken = false;
  // Cross-process exchange over peer memory (one rank per process, CUDA IPC): what this process can address of
  // every cell's exchange buffers, per direction; the epoch of the last apply per direction.
  struct PeerBuf {
    char* stage_in = nullptr;
    char* R = nullptr;
    unsigned* flags = nullptr;
  };
  bool xp = false, xp_tried = false;
  std::vector<PeerBuf> xpeer[2];
  std::vector<void*> xp_opened;
  unsigned xp_epoch[2] = {0u, 0u};
  bool perfstat = false;  // time every cell's share of each apply (replaces Jets.perfstat, src:723-745)
  int enqueue_ops[2] = {0, 0};  // launches + copies of one apply, counted on the first direct run

  int rank_at(int r, int c) const { return grid[r + c * n1]; }
  Cell& cell(int r, int c) { return cells[r + c * n1]; }
  Block& block(int64_t i, int64_t j) { return blocks[i * nbcol + j]; }
  int row_cell(int64_t i) const {
    return (int)(std::upper_bound(row_cuts.begin(), row_cuts.end(), i) - row_cuts.begin()) - 1;
  }
  int col_cell(int64_t j) const {
    return (int)(std::upper_bound(col_cuts.begin(), col_cuts.end(), j) - col_cuts.begin()) - 1;
  }
  int ndom_parts() const { return isdiag ? n1 : n2; }
};

namespace {

struct Transfer {
  int src_rank;
  const char* src;
  int dst_rank;
  char* dst;
  size_t bytes;
};

// One exchange phase: local->local pairs are device copies ordered with events; pairs with one end
// in another process go through ncclSend/ncclRecv inside one group.
void run_transfers(djets_ctx ctx, const std::vector<Transfer>& ts) {
  if (ts.empty()) return;
  const NcclApi* api = nullptr;
  bool grouped = false;
  for (const Transfer& t : ts) {
    if (t.bytes == 0) continue;
    RankCtx* s = ctx->local(t.src_rank);
    RankCtx* d = ctx->local(t.dst_rank);
    if (s && d) {
      stream_after(*d, *s);  // destination buffer is free once d's earlier work is done
      use(*s);
      if (s->device == d->device)
        CK(cudaMemcpyAsync(t.dst, t.src, t.bytes, cudaMemcpyDeviceToDevice, s->stream));
      else
        CK(cudaMemcpyPeerAsync(t.dst, d->device, t.src, s->device, t.bytes, s->stream));
      stream_after(*s, *d);
    } else if (s || d) {
      REQUIRE(ctx->has_nccl, DJETS_ERR_NCCL,
              "exchange with a rank of another process needs a context created with an NCCL unique id");
      if (!api) {
        const char* why = nullptr;
        api = nccl_api(&why);
        REQUIRE(api, DJETS_ERR_NCCL, why ? why : "NCCL unavailable");
      }
      if (!grouped) {
        NCK(api, api->GroupStart());
        grouped = true;
      }
      if (s) {
        use(*s);
        NCK(api, api->Send(t.src, t.bytes, ncclChar, t.dst_rank, s->comm, s->stream));
      } else {
        use(*d);
        NCK(api, api->Recv(t.dst, t.bytes, ncclChar, t.src_rank, d->comm, d->stream));
      }
    }
  }
  if (grouped) NCK(api, api->GroupEnd());
}


// `keep`: exchange buffers that were published to other processes (CUDA IPC) outlive the operator: a slower
// peer may still hold them mapped; they are freed with the context.
void free_plan(DirPlan& p, std::vector<void*>* keep = nullptr) {
  cudaFree(p.d_tiles);
  cudaFree(p.d_tiles_rev);
  cudaFree(p.d_links);
  cudaFree(p.d_items);
  cudaFree(p.d_tickets);
  cudaFree(p.d_counter);
  cudaFree(p.d_diags);
  cudaFree(p.W);
  cudaFree(p.stage_out);
  for (void* q : {(void*)p.flags, (void*)p.stage_in, (void*)p.R}) {
    if (!q) continue;
    if (keep)
      keep->push_back(q);
    else
      cudaFree(q);
  }
  p = DirPlan();
}

djets_ctx g_alloc_ctx = nullptr;  // context whose vector caches an allocation may reclaim (set by build_plan)

template <class T>
T* upload(const std::vector<T>& v) {
  if (v.empty()) return nullptr;
  T* d = nullptr;
  CK(device_alloc(g_alloc_ctx, reinterpret_cast<void**>(&d), v.size() * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

char* dalloc(size_t bytes) {
  if (bytes == 0) return nullptr;
  char* d = nullptr;
  CK(device_alloc(g_alloc_ctx, reinterpret_cast<void**>(&d), bytes + 256));
  CK(cudaMemset(d, 0, bytes + 256));
  return d;
}

struct TileShape {
  int64_t tile_rows, tile_cols, ntiles, nplanes;
};

// Forward: row strips of fwd_tx*VEC rows, column panels of at most fwd_tc (<= kFwdMaxTC) columns,
// balanced.  Adjoint: row panels of at most adj_max_rows rows, column groups of adj_tc columns.
TileShape plan_shape(int dtype, int64_t m, int64_t n, bool adjoint, int64_t fwd_tx, int64_t fwd_tc, int64_t adj_tc) {
  TileShape s{0, 0, 0, 0};
  if (m <= 0 || n <= 0) return s;
  if (!adjoint) {
    const int64_t tr = fwd_tx * vec_elems(dtype);
    const int64_t tcmax = std::max<int64_t>(1, std::min<int64_t>(fwd_tc, kFwdMaxTC / (is_complex(dtype) ? 2 : 1)));
    const int64_t npan = ceil_div(n, tcmax);
    s.tile_rows = tr;
    s.tile_cols = ceil_div(n, npan);
    s.nplanes = npan;
    s.ntiles = ceil_div(m, tr) * npan;
  } else {
    const int64_t trmax = adj_max_rows(dtype);
    const int64_t npan = ceil_div(m, trmax);
    s.tile_rows = round_up(ceil_div(m, npan), vec_elems(dtype));
    s.tile_cols = std::max<int64_t>(kAdjCW, round_up(adj_tc, kAdjCW));
    s.nplanes = ceil_div(m, s.tile_rows);
    s.ntiles = s.nplanes * ceil_div(n, s.tile_cols);
  }
  return s;
}

// Blocks per chained tile of a cell (1 = no chains) and whether its adjoint may load two blocks per batch: see
// build_plan and djets_plan_chains.  Real dtypes only; every dense block at most two warp-loads tall; forward chains
// also need their columns to fit the staging buffer.
struct ChainRule {
  int64_t k = 1;
  bool pair = false;
};
ChainRule chain_rule(int dtype, bool adjoint, int64_t nblocks, int64_t longest, int64_t max_rows, int64_t max_cols,
                     int64_t max_bytes, int wave, int64_t chain_param, bool auto_tile) {
  ChainRule r;
  if (is_complex(dtype) || chain_param == 1 || nblocks <= 0) return r;
  if (max_rows > chain_max_rows(dtype) || (!adjoint && max_cols > kFwdMaxTC / 2)) return r;
  if (chain_param > 1)
    r.k = std::min<int64_t>(chain_param, kChainMax);
  else if (auto_tile && wave > 0) {
    // ~1 MB per chain where that still leaves sixteen waves of CTAs (90,000 blocks of 100x100 Float64, forward / adjoint
    // TB/s: 256 KB 6.75 / 6.12, 512 KB 7.23 / 6.70, 1 MB 7.28 / 6.82), otherwise ~512 KB and at least four waves (14,400
    // blocks: 6.90 / 6.42 against 6.72 / 6.42 with longer chains and eight waves): a short last wave costs a whole chain.
    const int64_t bytes = std::max<int64_t>(max_bytes, 1);
    const int64_t big = (1024 * 1024) / bytes;
    if (nblocks / (16 * (int64_t)wave) >= big)
      r.k = std::min<int64_t>(kChainMax, big);
    else
      r.k = std::min<int64_t>({(int64_t)kChainMax, (512 * 1024) / bytes, nblocks / (4 * (int64_t)wave)});
  }
  r.k = std::max<int64_t>(1, std::min(r.k, longest));  // block-diagonal operators have nothing to chain
  r.pair = r.k > 1 && max_rows <= chain_max_rows(dtype) / 2;
  return r;
}

// Builds one direction of one cell.  `outs` are the block indices along the output dimension that
// the cell owns (block rows for forward, block columns for adjoint), `ins` along the input one.
void build_plan(djets_op op, Cell& cell, bool adjoint) {
  djets_ctx ctx = op->ctx;
  RankCtx* rc = ctx->local(cell.rank);
  if (!rc) return;
  use(*rc);
  g_alloc_ctx = ctx;
  DirPlan& plan = adjoint ? cell.adj : cell.fwd;
  free_plan(plan, op->xp_tried ? &ctx->graveyard : nullptr);
  const int es = elem_size(op->dtype);
  // Complex operators: the kernels are real-typed and see (re, im) pairs, so every extent and offset that counts
  // ELEMENTS of a vector, of the workspace or down a matrix column is written into the descriptors in reals (x cs);
  // tile column counts stay in (complex) columns.  Sizes in bytes use `es` and need no scaling.
  const int64_t cs = is_complex(op->dtype) ? 2 : 1;
  const int64_t r0 = op->row_cuts[cell.r], r1 = op->row_cuts[cell.r + 1];
  const int64_t c0 = op->isdiag ? r0 : op->col_cuts[cell.c], c1 = op->isdiag ? r1 : op->col_cuts[cell.c + 1];
  const int64_t o0 = adjoint ? c0 : r0, o1 = adjoint ? c1 : r1;  // output blocks
  const int64_t i0 = adjoint ? r0 : c0, i1 = adjoint ? r1 : c1;  // input blocks
  const std::vector<int64_t>& out_len = adjoint ? op->col_len : op->row_len;
  const std::vector<int64_t>& out_off = adjoint ? op->col_off : op->row_off;
  const std::vector<int64_t>& in_off = adjoint ? op->row_off : op->col_off;
  // the cell that owns the output part is on grid column 0 (forward) / grid row 0 (adjoint)
  const bool owner = adjoint ? (op->isdiag || cell.r == 0) : (cell.c == 0);
  const int nremote = owner ? ((adjoint ? (op->isdiag ? 1 : op->n1) : op->n2) - 1) : 0;
  const int64_t part_len = adjoint ? (op->isdiag ? op->colpart_len[cell.r] : op->colpart_len[cell.c])
                                   : op->rowpart_len[cell.r];

  // Tile extents of this cell and direction.  The tuned maxima (fwd_tc / adj_tc) apply to large
  // operators; small ones get smaller tiles so that one launch still fills 148 SMs several times over
  // (auto_tile).  One value per cell: every block of a block row / column must cut the same strips.
  int64_t eff_fwd_tc = ctx->fwd_tc, eff_adj_tc = ctx->adj_tc;
  int adj_ru = (int)op->p_adj_ru, adj_cw = kAdjCW;
  int64_t fwd_tx = op->p_fwd_tx;
  if (ctx->auto_tile) {
    int64_t dense_bytes = 0, max_rows = 1, max_cols = 1;
    double weighted_rows = 0.0, weighted_cols = 0.0;
    for (int64_t i = r0; i < r1; ++i)
      for (int64_t j = c0; j < c1; ++j) {
        if (op->isdiag && i != j) continue;
        if (op->block(i, j).kind != DJETS_BLOCK_DENSE) continue;
        const int64_t b = op->row_len[i] * op->col_len[j] * es;
        dense_bytes += b;
        weighted_rows += (double)b * (double)op->row_len[i];
        weighted_cols += (double)b * (double)op->col_len[j];
        max_rows = std::max(max_rows, op->row_len[i]);
        max_cols = std::max(max_cols, op->col_len[j]);
      }
    // ranks of this process that share the device run their cells side by side: size the tiles for the sum of them
    int share = 0;
    for (const Cell& other : op->cells)
      if (RankCtx* orc = ctx->local(other.rank)) share += orc->device == rc->device ? 1 : 0;
    const int64_t tile_bytes = share > 1 ? std::max<int64_t>(64 * 1024, dense_bytes * share / (148 * 4))
                                         : std::max<int64_t>(64 * 1024, dense_bytes / (148 * 6));
    // forward: narrow blocks (few columns) give a tile its bytes through more rows -- the CTA goes wider
    // along the rows and needs fewer (or no) column groups to sum
    const int64_t typical_cols = dense_bytes > 0 ? (int64_t)(weighted_cols / (double)dense_bytes) : 1;
    const int64_t typical_rows_f = dense_bytes > 0 ? (int64_t)(weighted_rows / (double)dense_bytes) : 1;
    int64_t wide = typical_cols <= 128 ? 256 : (typical_cols <= 256 ? 128 : fwd_tx);
    while (wide > fwd_tx && wide * vec_elems(op->dtype) > typical_rows_f) wide /= 2;  // never wider than the blocks are tall
    fwd_tx = std::max(fwd_tx, wide);
    const int64_t tr = fwd_tx * vec_elems(op->dtype);
    eff_fwd_tc = std::max<int64_t>(64, std::min<int64_t>(ctx->fwd_tc, round_up(std::max<int64_t>(tile_bytes / (tr * es), 1), 64)));
    // adjoint: short columns (bytes-weighted mean height) fan a warp out over more columns, and get wider
    // tiles so that a tile is still >= 256 KB
    const int64_t typical_rows = dense_bytes > 0 ? (int64_t)(weighted_rows / (double)dense_bytes) : 1;
    const int64_t chunks = ceil_div(std::min<int64_t>(typical_rows, adj_max_rows(op->dtype)), 32 * vec_elems(op->dtype));
    // blocks too small to make a 128 KB tile (100x100 Float64 ...): fewer loads per lane, three CTAs per SM instead of
    // two (4.62 -> 4.99 TB/s on 90,000 blocks of 100x100 F64; no gain once tiles can be widened)
    // (columns of 5 or 6 warp-loads were tried on the 2 x 8 shape as well: 333x777 Float64 6.79 -> 5.95 TB/s, 600x600
    // Float32 6.87 -> 6.20 -- the 4 x 4 shape with its paired remainder stays)
    if (chunks <= 2 && ctx->adj_light && max_rows * max_cols * es < 128 * 1024) {
      adj_ru = 1;
      adj_cw = 8;
    } else if (chunks <= 2) {
      adj_ru = 1;
      adj_cw = 16;
    } else if (chunks == 3) {
      adj_ru = 2;
      adj_cw = 8;
    }
    const int64_t rows = std::min<int64_t>(max_rows, adj_max_rows(op->dtype));
    if (adj_cw == kAdjCW) {
      // tall columns: adj_tc columns (tuned at 16 KB columns = 1 MiB tiles), more of them when the
      // columns are shorter so that a tile stays about 1 MiB; smaller only for small operators
      const int64_t cap = std::min<int64_t>(1024, std::max<int64_t>(ctx->adj_tc, round_up(ctx->adj_tc * 16384 / (rows * es), 32)));
      eff_adj_tc = std::max<int64_t>(32, std::min<int64_t>(cap, round_up(std::max<int64_t>(tile_bytes / (rows * es), 1), 32)));
    } else {
      const int64_t pass = 8 * adj_cw;  // columns one CTA covers per pass over its warps
      const int64_t want = std::min<int64_t>(std::max<int64_t>(tile_bytes, 256 * 1024), 1 << 20) / (rows * es);
      eff_adj_tc = std::min<int64_t>(4096, std::max<int64_t>(pass, round_up(std::max<int64_t>(want, 1), pass)));
    }
  }
  if (cs == 2) {  // the complex adjoint kernel comes in two shapes: 4 row chunks x 4 columns, 1 x 8 for short columns
    adj_ru = adj_ru == 1 ? 1 : 4;
    fwd_tx = std::max<int64_t>(fwd_tx, 64);
  }
  if (adjoint) {
    plan.adj_ru = adj_ru;
    plan.adj_cw = adj_cw;
  } else {
    plan.fwd_tx = (int)fwd_tx;
  }
  plan.wave = gemv_resident_ctas(op->dtype, adjoint ? 1 : 0, adjoint ? adj_ru : (int)fwd_tx, adj_cw, (int)op->p_ld_policy);
  // Chained tiles: when every dense block of the cell is at most two warp-loads tall (128 Float64 / 256 Float32 rows)
  // a CTA multiplies up to chain_k blocks of one block column (adjoint) / block row (forward) and writes one partial --
  // enough blocks to make a tile of 0.5-1 MB, but never so many that the cell is left with fewer than four waves of
  // CTAs ("chain": 0 = this rule, 1 = off, k > 1 = k blocks whatever the operator's size).  90,000 blocks of 100x100
  // Float64: adjoint 5.0 -> 6.7 TB/s (profiles/r02_block_shapes_1gpu.jsonl).
  int64_t chain_k = 1;
  bool chain_pair = false;
  if (cs == 1 && ctx->chain != 1) {
    int64_t nblk = 0, max_rows = 0, max_cols = 0, max_bytes = 1;
    std::vector<int64_t> per_out((size_t)(adjoint ? c1 - c0 : r1 - r0), 0);  // dense blocks feeding each output block
    for (int64_t i = r0; i < r1; ++i)
      for (int64_t j = c0; j < c1; ++j) {
        if (op->isdiag && i != j) continue;
        if (op->block(i, j).kind != DJETS_BLOCK_DENSE || op->row_len[i] == 0 || op->col_len[j] == 0) continue;
        ++nblk;
        ++per_out[(size_t)(adjoint ? j - c0 : i - r0)];
        max_rows = std::max(max_rows, op->row_len[i]);
        max_cols = std::max(max_cols, op->col_len[j]);
        max_bytes = std::max(max_bytes, op->row_len[i] * op->col_len[j] * es);
      }
    const int chain_wave = gemv_chain_resident_ctas(op->dtype, adjoint ? 1 : 0, (int)op->p_ld_policy);
    const int64_t longest = per_out.empty() ? 0 : *std::max_element(per_out.begin(), per_out.end());
    const ChainRule rule = chain_rule(op->dtype, adjoint, nblk, longest, max_rows, max_cols, max_bytes, chain_wave,
                                      ctx->chain, ctx->auto_tile != 0);
    chain_k = rule.k;
    chain_pair = rule.pair;
    if (chain_k > 1) {
      plan.wave = chain_wave;
      if (!adjoint) plan.fwd_tx = chain_pair ? 32 : 64;  // one row strip covers the tallest block
    }
  }
  // Wave fit: a tile count between one and two resident waves leaves the second wave part empty and its CTAs finish
  // one by one (profiles/r02_trace_boundary_8blocks.txt).  If slightly larger tiles bring the whole cell into ONE wave,
  // take them -- only when the second wave would be less than 60 % full; the N=8 shard of the bench operator sits at 1.73
  // waves (profiles/r02_gemv_shard8_ncu_summary.txt) and keeps its tiles.
  if (ctx->auto_tile && ctx->wave_fit && plan.wave > 0 && chain_k == 1) {
    auto count_tiles = [&](int64_t ftc, int64_t atc) {
      int64_t nt = 0;
      for (int64_t i = r0; i < r1; ++i)
        for (int64_t j = c0; j < c1; ++j) {
          if (op->isdiag && i != j) continue;
          if (op->block(i, j).kind != DJETS_BLOCK_DENSE) continue;
          nt += plan_shape(op->dtype, op->row_len[i], op->col_len[j], adjoint, fwd_tx, ftc, atc).ntiles;
        }
      return nt;
    };
    const int64_t nt0 = count_tiles(eff_fwd_tc, eff_adj_tc);
    if (nt0 > plan.wave && 10 * nt0 <= 16 * (int64_t)plan.wave) {  // second wave less than 60 % full
      if (!adjoint) {
        for (int64_t tc = eff_fwd_tc + 64; tc <= kFwdMaxTC / cs; tc += 64)
          if (count_tiles(tc, eff_adj_tc) <= plan.wave) {
            eff_fwd_tc = tc;
            break;
          }
      } else {
        const int64_t step = adj_cw == kAdjCW ? 16 : 8 * adj_cw;
        for (int64_t tc = eff_adj_tc + step; tc <= 2 * eff_adj_tc; tc += step)
          if (count_tiles(eff_fwd_tc, tc) <= plan.wave) {
            eff_adj_tc = tc;
            break;
          }
      }
    }
  }
  std::vector<DenseTile> tiles;
  std::vector<ChainLink> links;
  std::vector<int64_t> tile_key;  // block (i, j) of every tile, row-major: both directions walk the blocks in this order
  std::vector<FinishItem> loose, fused;
  std::vector<int64_t> loose_key;  // offset of the item's chunk inside its block: the launch order of the loose items
  std::vector<DiagTerm> diags;
  int64_t wtotal = 0;
  // A cell that sums planes received from peers finishes in a separate kernel after the receive;
  // everywhere else the last tile CTA of a strip does the sum inside the GEMV kernel.
  const bool can_fuse = (nremote == 0);

  for (int64_t o = o0; o < o1; ++o) {
    const int64_t olen = out_len[o];
    if (olen == 0) continue;
    struct Contribution {
      int64_t in;
      Block* b;
    };
    std::vector<Contribution> dense, diag;
    int64_t nplanes = 0;
    for (int64_t in = i0; in < i1; ++in) {
      if (op->isdiag && in != o) continue;
      const int64_t i = adjoint ? in : o, j = adjoint ? o : in;
      Block& b = op->block(i, j);
      if (b.kind == DJETS_BLOCK_DENSE && op->row_len[i] > 0 && op->col_len[j] > 0) {
        dense.push_back({in, &b});
        if (chain_k == 1)
          nplanes += plan_shape(op->dtype, op->row_len[i], op->col_len[j], adjoint, fwd_tx, eff_fwd_tc, eff_adj_tc)
                         .nplanes;
      } else if (b.kind == DJETS_BLOCK_DIAG) {
        diag.push_back({in, &b});
      }
    }
    // chains: runs of up to chain_k blocks (forward: as long as their columns fit the staging buffer), one plane each
    auto extent = [&](const Contribution& cb) { return adjoint ? op->row_len[cb.in] : op->col_len[cb.in]; };
    std::vector<std::pair<size_t, size_t>> chains;
    for (size_t first = 0; chain_k > 1 && first < dense.size();) {
      size_t end = first + 1;
      int64_t span = extent(dense[first]);
      while (end < dense.size() && (int64_t)(end - first) < chain_k && (adjoint || span + extent(dense[end]) <= kFwdMaxTC))
        span += extent(dense[end++]);
      chains.push_back({first, end});
      first = end;
    }
    if (chain_k > 1) nplanes = (int64_t)chains.size();
    const bool direct = (nplanes == 1 && diag.empty() && nremote == 0);
    const bool fuse = can_fuse && !direct && !dense.empty();
    const int64_t wbase = wtotal;
    if (!direct) wtotal += nplanes * olen;

    // segmented-sum items of this output block, in chunks of `chunk` elements
    auto emit_items = [&](std::vector<FinishItem>& dst, int64_t chunk) {
      const int first = (int)dst.size();
      for (int64_t e0 = 0; e0 < olen; e0 += chunk) {
        FinishItem it;
        it.out = cs * (out_off[o] + e0);
        it.w0 = cs * (wbase + e0);
        it.wstride = cs * olen;
        it.r0 = cs * (out_off[o] + e0);
        it.rstride = cs * part_len;
        it.nplanes = (int)nplanes;
        it.nremote = nremote;
        it.diag_begin = (int)diags.size();
        for (const Contribution& cb : diag) {
          DiagTerm d;
          d.d = cb.b->d + e0 * es;
          d.vin = cs * (in_off[cb.in] + e0);
          diags.push_back(d);
        }
        it.diag_end = (int)diags.size();
        it.len = (int)(cs * std::min<int64_t>(chunk, olen - e0));
        it.flags = cs == 2 ? (adjoint ? 3 : 1) : 0;
        dst.push_back(it);
        if (&dst == &loose) loose_key.push_back(e0);
      }
      return first;
    };

    int fused_first = -1;
    int64_t strip = 0;  // strip length along the output dimension (identical for every block of the row/column)
    if (fuse) {
      const Contribution& cb = dense.front();
      const int64_t i = adjoint ? cb.in : o, j = adjoint ? o : cb.in;
      const TileShape sh = plan_shape(op->dtype, op->row_len[i], op->col_len[j], adjoint, fwd_tx, eff_fwd_tc,
                                      eff_adj_tc);
      strip = chain_k > 1 ? olen : (adjoint ? sh.tile_cols : sh.tile_rows);
      fused_first = emit_items(fused, strip);
    } else if (!direct) {
      emit_items(loose, std::max<int64_t>(kFinishChunk, kFinishChunkBytes / es));
    }

    int64_t plane = 0;
    for (const auto& [first, end] : chains) {
      const Contribution& head = dense[first];
      const int64_t i = adjoint ? head.in : o, j = adjoint ? o : head.in;
      DenseTile t;
      t.A = head.b->d;
      t.ld = head.b->ld;
      t.vin = in_off[head.in];
      t.rows = (int)op->row_len[i];
      t.cols = (int)op->col_len[j];
      t.direct = direct ? 1 : 0;
      t.wout = direct ? out_off[o] : wbase + plane * olen;
      t.fin = fuse ? fused_first : -1;
      t.chain = (int)(end - first - 1);
      t.link0 = (int)links.size();
      for (size_t k = first + 1; k < end; ++k)
        links.push_back({dense[k].b->d, dense[k].b->ld, in_off[dense[k].in], (int)extent(dense[k]), 0});
      tiles.push_back(t);
      tile_key.push_back(i * op->nbcol + j);
      ++plane;
    }
    if (chain_k > 1) continue;
    for (const Contribution& cb : dense) {
      const int64_t i = adjoint ? cb.in : o, j = adjoint ? o : cb.in;
      const int64_t m = op->row_len[i], n = op->col_len[j];
      const TileShape sh = plan_shape(op->dtype, m, n, adjoint, fwd_tx, eff_fwd_tc, eff_adj_tc);
      if (!adjoint) {
        for (int64_t col0 = 0, p = 0; col0 < n; col0 += sh.tile_cols, ++p) {
          for (int64_t row0 = 0; row0 < m; row0 += sh.tile_rows) {
            DenseTile t;
            t.A = cb.b->d + (row0 + col0 * cb.b->ld) * es;
            t.ld = cs * cb.b->ld;
            t.vin = cs * (in_off[j] + col0);
            t.rows = (int)(cs * std::min<int64_t>(sh.tile_rows, m - row0));
            t.cols = (int)std::min<int64_t>(sh.tile_cols, n - col0);
            t.direct = direct ? 1 : 0;
            t.wout = cs * (direct ? out_off[o] + row0 : wbase + (plane + p) * olen + row0);
            t.fin = fuse ? fused_first + (int)(row0 / strip) : -1;
            t.chain = t.link0 = 0;
            tiles.push_back(t);
            tile_key.push_back(i * op->nbcol + j);
          }
        }
      } else {
        for (int64_t row0 = 0, p = 0; row0 < m; row0 += sh.tile_rows, ++p) {
          for (int64_t col0 = 0; col0 < n; col0 += sh.tile_cols) {
            DenseTile t;
            t.A = cb.b->d + (row0 + col0 * cb.b->ld) * es;
            t.ld = cs * cb.b->ld;
            t.vin = cs * (in_off[i] + row0);
            t.rows = (int)(cs * std::min<int64_t>(sh.tile_rows, m - row0));
            t.cols = (int)std::min<int64_t>(sh.tile_cols, n - col0);
            t.direct = direct ? 1 : 0;
            t.wout = cs * (direct ? out_off[o] + col0 : wbase + (plane + p) * olen + col0);
            t.fin = fuse ? fused_first + (int)(col0 / strip) : -1;
            t.chain = t.link0 = 0;
            tiles.push_back(t);
            tile_key.push_back(i * op->nbcol + j);
          }
        }
      }
      plane += sh.nplanes;
    }
  }
  // Loose items whose block rows share input blocks are launched chunk-major (same offset inside every block, block
  // after block): the diagonal terms of neighbouring CTAs then read the same slices of the input blocks, which stay in
  // L2 -- a full grid of diagonal blocks reads every input block once from HBM instead of once per block row
  // (24x24 diagonal blocks of 2^20 F64: 5.04 -> 5.67 TB/s).  Block-diagonal operators keep the block-major order
  // (contiguous streams: 6.86 TB/s against 6.58 chunk-major).
  bool shared_inputs = false;  // some item sums several diagonal terms: input slices are shared between block rows
  for (const FinishItem& it : loose) shared_inputs = shared_inputs || (it.diag_end - it.diag_begin >= 2);
  if (shared_inputs) {
    std::vector<size_t> order(loose.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return loose_key[a] < loose_key[b]; });
    std::vector<FinishItem> sorted;
    sorted.reserve(loose.size());
    for (size_t k : order) sorted.push_back(loose[k]);
    loose.swap(sorted);
  }
  // L2 hand-over between consecutive applies.  A solver alternates mul!(d, A, m) and mul!(m, A', d) (or repeats one of
  // them): both directions list their tiles in the same block order, every launch walks the table in the direction
  // opposite to the cell's previous launch -- it starts on the blocks that launch finished with -- and the last
  // `l2_keep` MB of a walk are loaded evict-last (everything else evict-first), so the next launch finds them in the
  // 126 MB L2 instead of HBM.
  {
    std::vector<size_t> order(tiles.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tile_key[a] < tile_key[b]; });
    std::vector<DenseTile> sorted;
    sorted.reserve(tiles.size());
    for (size_t k : order) sorted.push_back(tiles[k]);
    tiles.swap(sorted);
  }
  // Small operators (everything the GPU holds of it is within twice the L2) keep a fixed resident set instead: the
  // first `l2_keep` MB of the block order -- the same blocks in both directions and in every apply, shared evenly by
  // the ranks that share the GPU -- are always loaded evict-last and stay in L2 from apply to apply, whatever the
  // order CTAs run in (a cell that fits one wave has no walk order to speak of).
  int64_t cell_bytes = 0;
  for (const DenseTile& t : tiles) cell_bytes += (int64_t)t.rows * t.cols * (es / cs) * (1 + t.chain);
  int sharing = 0;
  for (const Cell& other : op->cells)
    if (RankCtx* orc = ctx->local(other.rank)) sharing += orc->device == rc->device ? 1 : 0;
  sharing = std::max(sharing, 1);
  const bool resident_set = ctx->l2_keep > 0 && cell_bytes * sharing <= (int64_t)252 << 20;
  if (resident_set) {
    int64_t budget = (ctx->l2_keep << 20) / sharing;
    for (DenseTile& t : tiles) {
      t.keep = budget > 0 ? 1 : 0;
      t.pad_ = 0;
      budget -= (int64_t)t.rows * t.cols * (es / cs) * (1 + t.chain);
    }
  }
  std::vector<DenseTile> tiles_rev(tiles.rbegin(), tiles.rend());
  for (std::vector<DenseTile>* table : {&tiles, &tiles_rev}) {
    if (resident_set) break;
    int64_t budget = ctx->l2_keep << 20;
    for (DenseTile& t : *table) {
      t.keep = 0;
      t.pad_ = 0;
    }
    for (size_t k = table->size(); k-- > 0 && budget > 0;) {
      (*table)[k].keep = 1;
      budget -= (int64_t)(*table)[k].rows * (*table)[k].cols * (es / cs) * (1 + (*table)[k].chain);
    }
  }
  // loose items first, fused items after them; tiles index the concatenated table
  const int nloose = (int)loose.size();
  for (std::vector<DenseTile>* table : {&tiles, &tiles_rev})
    for (DenseTile& t : *table)
      if (t.fin >= 0) t.fin += nloose;
  std::vector<FinishItem> items(loose);
  items.insert(items.end(), fused.begin(), fused.end());
  plan.nloose = nloose;
  plan.ntiles = (int)tiles.size();
  plan.nitems = (int)items.size();
  plan.d_tiles = upload(tiles);
  plan.d_tiles_rev = ctx->l2_keep > 0 ? upload(tiles_rev) : nullptr;
  plan.d_links = upload(links);
  plan.chained = chain_k > 1;
  plan.chain_pair = chain_pair;
  plan.d_items = upload(items);
  plan.d_tickets = reinterpret_cast<unsigned*>(dalloc(items.size() * sizeof(unsigned)));
  plan.d_counter = reinterpret_cast<unsigned*>(dalloc(sizeof(unsigned)));
  plan.d_diags = upload(diags);
  plan.W = dalloc((size_t)wtotal * es);
  // exchange buffers
  const int64_t in_len = adjoint ? op->rowpart_len[cell.r]
                                 : (op->isdiag ? op->colpart_len[cell.r] : op->colpart_len[cell.c]);
  const bool in_owner = adjoint ? (cell.c == 0) : (op->isdiag || cell.r == 0);
  // with ranks in other processes the receive buffers are double-buffered by the parity of the apply's epoch
  const int nbuf = (int)ctx->ranks.size() < ctx->world ? 2 : 1;
  if (!in_owner) plan.stage_in = dalloc((size_t)nbuf * round_up(in_len * es, 256));  // each buffer 256-byte aligned
  if (!owner) plan.stage_out = dalloc((size_t)part_len * es);
  if (nremote > 0) plan.R = dalloc((size_t)nbuf * nremote * part_len * es);
  plan.in_len = in_len;
  plan.part_len = part_len;
  plan.nremote = nremote;
  if (nbuf == 2 && !op->isdiag)
    plan.flags = reinterpret_cast<unsigned*>(dalloc(sizeof(unsigned) * (size_t)(2 + 2 * std::max(op->n1, op->n2))));
}

// Layout of a cell's flag array (per direction), nmax = max(n1, n2):
//   [0]              in_ready        raised by the owner of my input part once stage_in holds the apply's input
//   [1]              part_consumed   raised by the owner of my output part once it has summed my partial
//   [2 + slot]       part_ready      raised by the sender of receive slot `slot` (owners only)
//   [2 + nmax + idx] in_consumed     raised by consumer `idx` of my input part (input owners only)
inline int XP_IN_READY() { return 0; }
inline int XP_PART_CONSUMED() { return 1; }
inline int XP_PART_READY(int slot) { return 2 + slot; }
inline int XP_IN_CONSUMED(djets_op op, int idx) { return 2 + std::max(op->n1, op->n2) + idx; }

// Collective over the processes of the job (called from finalize, which every process reaches in the same
// order): all-gather the CUDA IPC handles of every cell's exchange buffers, map the remote ones, and agree
// (all-reduce) on whether every process succeeded; otherwise the exchange stays on ncclSend / ncclRecv.
void xp_setup(djets_op op) {
  djets_ctx ctx = op->ctx;
  if (op->xp_tried) return;
  op->xp_tried = true;
  const bool multi_process = (int)ctx->ranks.size() < ctx->world;
  if (!multi_process || !has_control_plane(ctx) || ctx->ranks.size() != 1 || op->isdiag || op->n1 * op->n2 < 2) return;
  RankCtx& rc = ctx->ranks.front();
  use(rc);
  struct Record {  // what one rank publishes
    cudaIpcMemHandle_t h[2][3];
    int present[2][3];
    int cell;  // index of the rank's cell in the grid, -1 when it is not part of this operator
    int ok;
    char gpu[16];  // UUID of the rank's GPU: ranks that share a GPU must not wait on each other's flags
  };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  Record mine;
  memset(&mine, 0, sizeof(mine));
  mine.cell = -1;
  const int ncell = op->n1 * op->n2;
  for (int k = 0; k < ncell; ++k)
    if (op->grid[k] == rc.rank) mine.cell = k;
  int ok = ctx->ipc ? 1 : 0;
  if (mine.cell >= 0 && ok) {
    Cell& cell = op->cells[mine.cell];
    for (int d = 0; d < 2; ++d) {
      DirPlan& plan = d ? cell.adj : cell.fwd;
      void* bufs[3] = {plan.stage_in, plan.R, plan.flags};
      for (int b = 0; b < 3; ++b) {
        if (!bufs[b]) continue;
        if (cudaIpcGetMemHandle(&mine.h[d][b], bufs[b]) != cudaSuccess) {
          (void)cudaGetLastError();
          ok = 0;
        } else {
          mine.present[d][b] = 1;
        }
      }
    }
  }
  mine.ok = ok;
  {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, rc.device));
    memcpy(mine.gpu, prop.uuid.bytes, 16);
  }
  std::vector<Record> all((size_t)ctx->world);
  control_allgather(ctx, &mine, all.data(), sizeof(Record));
  for (int r = 0; r < ctx->world; ++r)
    if (all[(size_t)r].cell >= 0 && !all[(size_t)r].ok) ok = 0;
  // Kernels of different ranks wait on one another's flags: that needs every rank on a GPU of its own (nothing
  // guarantees that kernels of several processes on ONE GPU run at the same time; the time-sliced spin can trip
  // the driver's context-switch watchdog).  Ranks sharing a GPU keep the stream-ordered paths.
  for (int a = 0; a < ctx->world && ok; ++a)
    for (int b = a + 1; b < ctx->world && ok; ++b)
      if (all[(size_t)a].cell >= 0 && all[(size_t)b].cell >= 0 && memcmp(all[(size_t)a].gpu, all[(size_t)b].gpu, 16) == 0) ok = 0;
  for (int d = 0; d < 2; ++d) op->xpeer[d].assign((size_t)ncell, djets_op_s::PeerBuf());
  for (int r = 0; r < ctx->world && ok; ++r) {
    const Record& rec = all[(size_t)r];
    if (rec.cell < 0 || rec.cell >= ncell) continue;
    for (int d = 0; d < 2 && ok; ++d) {
      djets_op_s::PeerBuf& pb = op->xpeer[d][(size_t)rec.cell];
      void* ptr[3] = {nullptr, nullptr, nullptr};
      if (r == rc.rank) {
        DirPlan& plan = d ? op->cells[(size_t)rec.cell].adj : op->cells[(size_t)rec.cell].fwd;
        ptr[0] = plan.stage_in;
        ptr[1] = plan.R;
        ptr[2] = plan.flags;
      } else {
        for (int b = 0; b < 3 && ok; ++b) {
          if (!rec.present[d][b]) continue;
          if (cudaIpcOpenMemHandle(&ptr[b], rec.h[d][b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            (void)cudaGetLastError();
            ok = 0;
          } else {
            op->xp_opened.push_back(ptr[b]);
          }
        }
      }
      pb.stage_in = reinterpret_cast<char*>(ptr[0]);
      pb.R = reinterpret_cast<char*>(ptr[1]);
      pb.flags = reinterpret_cast<unsigned*>(ptr[2]);
    }
  }
  // agree: the peer-memory path is used only if every process could map what it needs
  std::vector<int> oks((size_t)ctx->world, 0);
  control_allgather(ctx, &ok, oks.data(), sizeof(int));
  bool everyone = true;
  for (int r = 0; r < ctx->world; ++r)
    if (all[(size_t)r].cell >= 0 && !oks[(size_t)r]) everyone = false;
  op->xp = everyone;
  REQUIRE(op->xp || ctx->has_nccl, DJETS_ERR_NCCL,
          "the peer-memory exchange is unavailable (ranks of different processes share a GPU, or CUDA IPC could not map the "
          "buffers) and the context has no NCCL to fall back on");
}

void xp_release(djets_op op) {
  for (void* p : op->xp_opened) cudaIpcCloseMemHandle(p);
  op->xp_opened.clear();
  op->xp = false;
}

// The cross-rank traffic of one apply as a list of point-to-point messages between grid cells.
//   phase 0 (before the tile kernels): the owner of an input part -- grid row 0 for the domain
//     (forward), grid column 0 for the range (adjoint) -- sends it to the other cells of its grid
//     column / row (one message per apply instead of the reference's per-(i,j) remote getblock,
//     src:588,635,682).
//   phase 1 (after them): every cell that does not own its output part sends its reduced partial to
//     the owner on grid column 0 (forward) / grid row 0 (adjoint), which stores it in receive slot
//     `slot` (replaces ArrayFutures + reduce!, src:600-604, 647-651, 694-698).
// Block-diagonal operators exchange nothing (src:609-614, 656-661, 703-708).
struct Message {
  int src_r, src_c, dst_r, dst_c;
  int part;  // index of the vector part being moved (input part in phase 0, output part in phase 1)
  int slot;  // phase 1: receive plane on the owner; phase 0: -1
};

std::vector<Message> plan_exchange(int n1, int n2, bool isdiag, bool adjoint, int phase) {
  std::vector<Message> out;
  if (isdiag) return out;
  for (int c = 0; c < n2; ++c)
    for (int r = 0; r < n1; ++r) {
      if (phase == 0) {
        const bool in_owner = adjoint ? (c == 0) : (r == 0);
        if (in_owner) continue;
        out.push_back(adjoint ? Message{r, 0, r, c, r, -1} : Message{0, c, r, c, c, -1});
      } else {
        const bool owner = adjoint ? (r == 0) : (c == 0);
        if (owner) continue;
        out.push_back(adjoint ? Message{r, c, 0, c, c, r - 1} : Message{r, c, r, 0, r, c - 1});
      }
    }
  return out;
}

void check_vec(djets_op op, djets_vec v, bool is_range, const char* what) {
  REQUIRE(v && v->ctx == op->ctx, DJETS_ERR_INVALID, std::string(what) + ": vector of another context");
  REQUIRE(v->dtype == op->dtype, DJETS_ERR_MISMATCH, std::string(what) + ": element type differs from the operator's");
  const int np = is_range ? op->n1 : op->ndom_parts();
  const std::vector<int64_t>& len = is_range ? op->row_len : op->col_len;
  const std::vector<int64_t>& cuts = (is_range || op->isdiag) ? op->row_cuts : op->col_cuts;
  REQUIRE((int)v->parts.size() == np && v->block_len == len, DJETS_ERR_MISMATCH,
          std::string(what) + ": block layout differs from the operator's " + (is_range ? "range" : "domain"));
  for (int p = 0; p < np; ++p) {
    const int rank = (is_range || op->isdiag) ? op->rank_at(p, 0) : op->rank_at(0, p);
    REQUIRE(v->parts[p].rank == rank && v->parts[p].blk_first == cuts[p] &&
                v->parts[p].blk_count == cuts[p + 1] - cuts[p],
            DJETS_ERR_MISMATCH, std::string(what) + ": partition differs from the operator's");
  }
}

void finalize_op(djets_op op) {
  if (op->finalized) return;
  op->p_fwd_tx = op->ctx->fwd_tx;
  op->p_adj_ru = op->ctx->adj_ru;
  op->p_ld_policy = op->ctx->ld_policy;
  for (Cell& c : op->cells) {
    build_plan(op, c, false);
    build_plan(op, c, true);
  }
  xp_setup(op);
  op->finalized = true;
}

// perfstatfile: bracket every cell's share of an apply with events on its stream
void perfstat_mark(djets_op op, bool begin) {
  if (!op->perfstat) return;
  for (Cell& cell : op->cells) {
    RankCtx* rc = op->ctx->local(cell.rank);
    if (!rc) continue;
    use(*rc);
    if (!cell.pt0) {
      CK(cudaEventCreate(&cell.pt0));
      CK(cudaEventCreate(&cell.pt1));
    }
    CK(cudaEventRecord(begin ? cell.pt0 : cell.pt1, rc->stream));
    cell.ptimed = !begin;
  }
}

int apply_enqueue(djets_op op, djets_vec out, djets_vec in, bool adjoint);

// The tile table of a cell's next launch: opposite to the walk of its previous launch (L2 hand-over).
const DenseTile* walk(Cell& cell, const DirPlan& plan) {
  const bool backwards = cell.at_end && plan.d_tiles_rev;
  cell.at_end = !backwards && plan.d_tiles_rev != nullptr;
  return backwards ? plan.d_tiles_rev : plan.d_tiles;
}

// CTAs launched for a cell's tile kernel: one resident wave that walks the tile table (persistent), or one
// CTA per tile.
int tile_grid(djets_ctx ctx, const DirPlan& plan) {
  if (!ctx->persistent || plan.wave <= 0 || plan.chained) return plan.ntiles;
  return std::min(plan.ntiles, plan.wave);
}

int apply_enqueue_timed(djets_op op, djets_vec out, djets_vec in, bool adjoint) {
  perfstat_mark(op, true);
  const int n = apply_enqueue(op, out, in, adjoint);
  perfstat_mark(op, false);
  return n;
}

// One apply when every rank lives in its own process and the exchange buffers are mapped into each other's address
// space (CUDA IPC): nothing goes through NCCL.  The owner of an input part copies it into its consumers' stage_in
// (peer copies over NVLink) and raises their in_ready flags; a cell that does not own its output part lets its
// kernels store the reduced partial straight into the owner's receive plane (peer stores) and raises the owner's
// part_ready flag; each stream spins on its own local flags right before the kernel that needs the data.
// Buffers alternate with the parity of the apply's epoch, the *_consumed flags keep a producer at most two applies
// ahead of its consumer.  Replaces ArrayFutures + reduce! (src:600-604, 647-651, 694-698) and the remote getblock
// (src:588, 635, 682) like the NCCL path, with one-sided traffic and ~2 us flag kernels instead of two collectives.
int apply_enqueue_xp(djets_op op, djets_vec out, djets_vec in, bool adjoint) {
  djets_ctx ctx = op->ctx;
  RankCtx& rc = ctx->ranks.front();
  const int n1 = op->n1, n2 = op->n2, dir = adjoint ? 1 : 0;
  int mine = -1;
  for (int k = 0; k < n1 * n2; ++k)
    if (op->grid[k] == rc.rank) mine = k;
  const unsigned e = ++op->xp_epoch[dir];
  if (mine < 0) return 0;  // this rank holds no cell of the operator
  use(rc);
  const int r = mine % n1, c = mine / n1;
  Cell& cell = op->cells[mine];
  DirPlan& plan = adjoint ? cell.adj : cell.fwd;
  const int es = elem_size(op->dtype);
  const int par = (int)(e & 1u);
  const bool in_owner = adjoint ? (c == 0) : (r == 0);
  const bool owner = adjoint ? (r == 0) : (c == 0);
  const int ip = adjoint ? r : c, opart = adjoint ? c : r;
  std::vector<djets_op_s::PeerBuf>& peer = op->xpeer[dir];
  int nops = 0;
  auto wait = [&](const FlagList& f) {
    if (f.n == 0) return;
    ++nops;
    CK(launch_flags_wait(f, rc.d_xp_err, rc.stream));
  };
  auto signal = [&](const FlagList& f) {
    if (f.n == 0) return;
    ++nops;
    CK(launch_flags_signal(f, rc.stream));
  };
  auto push = [&](FlagList& f, unsigned* p, unsigned v) {
    REQUIRE(f.n < 16, DJETS_ERR_UNSUPPORTED, "peer-memory exchange: more than 16 peers along one grid dimension");
    f.p[f.n] = p;
    f.v[f.n++] = v;
  };
  // cells that consume my input part / that send partials to me
  const int nbcast = adjoint ? n2 : n1;   // cells along the dimension the input is broadcast over
  const int nred = adjoint ? n1 : n2;     // cells along the dimension the output is reduced over
  auto bcast_cell = [&](int idx) { return adjoint ? (r + idx * n1) : (idx + c * n1); };
  auto red_cell = [&](int idx) { return adjoint ? (idx + c * n1) : (r + idx * n1); };

  const char* inp = nullptr;
  bool pushed = false;
  Prefetch pf{plan.wave, (int)ctx->prefetch, (int)ctx->pdl, nullptr, {nullptr, nullptr}, {0u, 0u}, 0, rc.d_xp_err};
  auto kernel_wait = [&](unsigned* flag, unsigned v) {  // folded into the tile kernel's prologue when there is one
    pf.wait_flag[pf.nwait] = flag;
    pf.wait_value[pf.nwait++] = v;
  };
  if (in_owner) {
    inp = in->parts[ip].d;
    if (nbcast > 1) {
      // the pushes run on a side stream, so this cell's own tiles start at once
      cudaStream_t ps = ctx->xp_side ? rc.xstream : rc.stream;
      if (ctx->xp_side) {
        CK(cudaEventRecord(rc.ev_x0, rc.stream));
        CK(cudaStreamWaitEvent(rc.xstream, rc.ev_x0, 0));
      }
      FlagList w{}, s{};
      PushList pl{};
      for (int idx = 1; idx < nbcast; ++idx) push(w, plan.flags + XP_IN_CONSUMED(op, idx), e - 2u);
      const size_t bytes = (size_t)in->parts[ip].elem_count * es;
      for (int idx = 1; idx < nbcast; ++idx) {
        const djets_op_s::PeerBuf& pb = peer[(size_t)bcast_cell(idx)];
        pl.dst[pl.n++] = pb.stage_in + (size_t)par * round_up((int64_t)bytes, 256);
        push(s, pb.flags + XP_IN_READY(), e);
      }
      if (ctx->xp_push) {
        ++nops;
        CK(launch_xp_push(inp, bytes, pl, w, s, rc.d_xp_ticket, rc.d_xp_err, ps));
      } else {
        ++nops;
        CK(launch_flags_wait(w, rc.d_xp_err, ps));
        for (int k = 0; k < pl.n; ++k) {
          if (bytes) CK(cudaMemcpyAsync(pl.dst[k], inp, bytes, cudaMemcpyDeviceToDevice, ps));
          ++nops;
        }
        ++nops;
        CK(launch_flags_signal(s, ps));
      }
      if (ctx->xp_side) {
        CK(cudaEventRecord(rc.ev_x1, rc.xstream));
        pushed = true;
      }
    }
  } else {
    kernel_wait(plan.flags + XP_IN_READY(), e);
    inp = plan.stage_in + (size_t)par * round_up(plan.in_len * es, 256);
  }
  char* outp = nullptr;
  if (owner) {
    outp = out->parts[opart].d;
  } else {
    kernel_wait(plan.flags + XP_PART_CONSUMED(), e - 2u);
    const int slot = (adjoint ? r : c) - 1;
    outp = peer[(size_t)red_cell(0)].R + ((size_t)par * (nred - 1) + slot) * (size_t)plan.part_len * es;
  }
  if ((plan.ntiles == 0 || !ctx->xp_kwait) && pf.nwait > 0) {  // no tile kernel to carry the waits
    FlagList w{};
    for (int k = 0; k < pf.nwait; ++k) push(w, pf.wait_flag[k], pf.wait_value[k]);
    wait(w);
    pf.nwait = 0;
  }
  if (plan.ntiles > 0) {
    ++nops;
    pf.trace = trace_slot(rc, adjoint ? DJETS_K_GEMV_T : DJETS_K_GEMV_N, plan.ntiles);
    const FusedFinish ff{plan.d_items, plan.d_diags, plan.d_tickets};
    if (!adjoint)
      launch_counted(ctx, rc, DJETS_K_GEMV_N, [&] {
        if (plan.chained)
          return launch_gemv_n_chain(op->dtype, plan.fwd_tx, (int)op->p_ld_policy, walk(cell, plan), plan.d_links,
                                     plan.ntiles, inp, plan.W, outp, ff, pf, rc.stream);
        return launch_gemv_n(op->dtype, plan.fwd_tx, (int)op->p_ld_policy, walk(cell, plan), plan.ntiles, tile_grid(ctx, plan),
                             plan.d_counter, inp, plan.W, outp, ff, pf, rc.stream);
      });
    else
      launch_counted(ctx, rc, DJETS_K_GEMV_T, [&] {
        if (plan.chained)
          return launch_gemv_t_chain(op->dtype, plan.chain_pair ? 1 : 0, (int)op->p_ld_policy, walk(cell, plan), plan.d_links,
                                     plan.ntiles, inp, plan.W, outp, ff, pf, rc.stream);
        return launch_gemv_t(op->dtype, plan.adj_ru, plan.adj_cw, (int)op->p_ld_policy, walk(cell, plan), plan.ntiles,
                             tile_grid(ctx, plan), plan.d_counter, inp, plan.W, outp, ff, pf, rc.stream);
      });
  }
  if (!owner) {
    if (plan.nloose > 0) {
      ++nops;
      launch_counted(ctx, rc, DJETS_K_FINISH, [&] {
        return launch_finish(op->dtype, plan.d_items, plan.nloose, plan.d_diags, inp, plan.W, nullptr, outp, rc.stream);
      });
    }
    FlagList s{};
    const int slot = (adjoint ? r : c) - 1;
    push(s, peer[(size_t)red_cell(0)].flags + XP_PART_READY(slot), e);
    if (!in_owner) push(s, peer[(size_t)bcast_cell(0)].flags + XP_IN_CONSUMED(op, adjoint ? c : r), e);
    signal(s);
  } else {
    if (plan.nloose > 0) {
      FlagList w{};
      for (int idx = 1; idx < nred; ++idx) push(w, plan.flags + XP_PART_READY(idx - 1), e);
      wait(w);
      ++nops;
      const char* Rp = plan.R ? plan.R + (size_t)par * plan.nremote * plan.part_len * es : nullptr;
      launch_counted(ctx, rc, DJETS_K_FINISH, [&] {
        return launch_finish(op->dtype, plan.d_items, plan.nloose, plan.d_diags, inp, plan.W, Rp, outp, rc.stream);
      });
    }
    FlagList s{};
    for (int idx = 1; idx < nred; ++idx) push(s, peer[(size_t)red_cell(idx)].flags + XP_PART_CONSUMED(), e);
    if (!in_owner) push(s, peer[(size_t)bcast_cell(0)].flags + XP_IN_CONSUMED(op, adjoint ? c : r), e);
    signal(s);
  }
  if (pushed) CK(cudaStreamWaitEvent(rc.stream, rc.ev_x1, 0));  // later writers of the input part wait for the pushes
  return nops;
}

// Enqueues one apply on the ranks' streams: exchange phase 0, tile kernels (+ fused sums), exchange
// phase 1, owners' segmented sums.  Returns the number of enqueued operations.
int apply_enqueue(djets_op op, djets_vec out, djets_vec in, bool adjoint) {
  djets_ctx ctx = op->ctx;
  if (op->xp) return apply_enqueue_xp(op, out, in, adjoint);
  int nops = 0;
  const int es = elem_size(op->dtype);
  const int n1 = op->n1, n2 = op->isdiag ? 1 : op->n2;

  // When both ends of a message are ranks of this process whose GPUs can address each other, no
  // copy is made: the consumer's kernels read the owner's input part in place (peer loads over
  // NVLink), and a non-owner's kernels store their reduced partial straight into the owner's receive
  // plane (peer stores) -- the transfer is part of the compute kernel.  Streams are ordered with events.
  auto in_place = [&](int rank_a, int rank_b) {
    RankCtx* a = ctx->local(rank_a);
    RankCtx* b = ctx->local(rank_b);
    return ctx->p2p && a && b && ctx->can_address(*a, *b) && ctx->can_address(*b, *a);
  };
  const int ncell = n1 * n2;
  std::vector<const char*> in_direct(ncell, nullptr);  // input part read in place from its owner
  std::vector<char*> out_direct(ncell, nullptr);       // receive plane of the owner written in place
  std::vector<int> in_from(ncell, -1), out_to(ncell, -1);

  // phase A: every cell gets the input blocks of its block columns (forward) / block rows (adjoint)
  std::vector<Transfer> ts;
  for (const Message& msg : plan_exchange(n1, n2, op->isdiag, adjoint, 0)) {
    Cell& cell = op->cell(msg.dst_r, msg.dst_c);
    DirPlan& plan = adjoint ? cell.adj : cell.fwd;
    const VecPart& part = in->parts[msg.part];
    const int k = msg.dst_r + msg.dst_c * n1;
    if (in_place(part.rank, cell.rank)) {
      in_direct[k] = part.d;
      in_from[k] = part.rank;
      stream_after(*ctx->local(part.rank), *ctx->local(cell.rank));  // the part is complete before it is read
    } else {
      ts.push_back({part.rank, part.d, cell.rank, plan.stage_in, (size_t)part.elem_count * es});
    }
  }
  run_transfers(ctx, ts);
  nops += (int)ts.size();

  std::vector<Message> phase1 = plan_exchange(n1, n2, op->isdiag, adjoint, 1);
  for (const Message& msg : phase1) {
    Cell& cell = op->cell(msg.src_r, msg.src_c);
    Cell& own = op->cell(msg.dst_r, msg.dst_c);
    DirPlan& oplan = adjoint ? own.adj : own.fwd;
    if (in_place(cell.rank, own.rank)) {
      const int64_t part_len = adjoint ? op->colpart_len[msg.part] : op->rowpart_len[msg.part];
      const int k = msg.src_r + msg.src_c * n1;
      out_direct[k] = oplan.R + (size_t)msg.slot * part_len * es;
      out_to[k] = own.rank;
      stream_after(*ctx->local(own.rank), *ctx->local(cell.rank));  // the owner's previous sum has read the plane
    }
  }

  // phase B: tile kernels (with the fused segmented sums), loose sums on cells that do not own the output part
  for (int c = 0; c < n2; ++c)
    for (int r = 0; r < n1; ++r) {
      Cell& cell = op->cell(r, c);
      RankCtx* rc = ctx->local(cell.rank);
      if (!rc) continue;
      use(*rc);
      DirPlan& plan = adjoint ? cell.adj : cell.fwd;
      const int k = r + c * n1;
      const bool in_owner = op->isdiag || (adjoint ? (c == 0) : (r == 0));
      const bool owner = op->isdiag || (adjoint ? (r == 0) : (c == 0));
      const int ip = op->isdiag ? r : (adjoint ? r : c);
      const int opart = op->isdiag ? r : (adjoint ? c : r);
      const char* inp = in_owner ? in->parts[ip].d : (in_direct[k] ? in_direct[k] : plan.stage_in);
      char* outp = owner ? out->parts[opart].d : (out_direct[k] ? out_direct[k] : plan.stage_out);
      if (plan.ntiles > 0) {
        ++nops;
        if (!adjoint)
          launch_counted(ctx, *rc, DJETS_K_GEMV_N, [&] {
            const FusedFinish ff{plan.d_items, plan.d_diags, plan.d_tickets};
            const Prefetch pf{plan.wave, (int)ctx->prefetch, (int)ctx->pdl, trace_slot(*rc, DJETS_K_GEMV_N, plan.ntiles),
                              {nullptr, nullptr}, {0u, 0u}, 0, nullptr};
            if (plan.chained)
              return launch_gemv_n_chain(op->dtype, plan.fwd_tx, (int)op->p_ld_policy, walk(cell, plan), plan.d_links,
                                         plan.ntiles, inp, plan.W, outp, ff, pf, rc->stream);
            return launch_gemv_n(op->dtype, plan.fwd_tx, (int)op->p_ld_policy, walk(cell, plan), plan.ntiles,
                                 tile_grid(ctx, plan), plan.d_counter, inp, plan.W, outp, ff, pf, rc->stream);
          });
        else
          launch_counted(ctx, *rc, DJETS_K_GEMV_T, [&] {
            const FusedFinish ff{plan.d_items, plan.d_diags, plan.d_tickets};
            const Prefetch pf{plan.wave, (int)ctx->prefetch, (int)ctx->pdl, trace_slot(*rc, DJETS_K_GEMV_T, plan.ntiles),
                              {nullptr, nullptr}, {0u, 0u}, 0, nullptr};
            if (plan.chained)
              return launch_gemv_t_chain(op->dtype, plan.chain_pair ? 1 : 0, (int)op->p_ld_policy, walk(cell, plan),
                                         plan.d_links, plan.ntiles, inp, plan.W, outp, ff, pf, rc->stream);
            return launch_gemv_t(op->dtype, plan.adj_ru, plan.adj_cw, (int)op->p_ld_policy, walk(cell, plan), plan.ntiles,
                                 tile_grid(ctx, plan), plan.d_counter, inp, plan.W, outp, ff, pf, rc->stream);
          });
      }
      if (!owner && plan.nloose > 0) {
        ++nops;
        launch_counted(ctx, *rc, DJETS_K_FINISH, [&] {
          return launch_finish(op->dtype, plan.d_items, plan.nloose, plan.d_diags, inp, plan.W, nullptr, outp,
                               rc->stream);
        });
      }
      // later writes to the input part (on its owner's stream) wait for this cell's reads; the owner's
      // segmented sum waits for this cell's stores
      if (in_from[k] >= 0) stream_after(*rc, *ctx->local(in_from[k]));
      if (out_to[k] >= 0) stream_after(*rc, *ctx->local(out_to[k]));
    }

  // phase C: partials of the other cells of a grid row (forward) / grid column (adjoint) go to the owner
  ts.clear();
  for (const Message& msg : phase1) {
    if (out_direct[msg.src_r + msg.src_c * n1]) continue;
    Cell& cell = op->cell(msg.src_r, msg.src_c);
    Cell& own = op->cell(msg.dst_r, msg.dst_c);
    DirPlan& plan = adjoint ? cell.adj : cell.fwd;
    DirPlan& oplan = adjoint ? own.adj : own.fwd;
    const int64_t part_len = adjoint ? op->colpart_len[msg.part] : op->rowpart_len[msg.part];
    ts.push_back({cell.rank, plan.stage_out, own.rank,
                  oplan.R ? oplan.R + (size_t)msg.slot * part_len * es : nullptr, (size_t)part_len * es});
  }
  run_transfers(ctx, ts);
  nops += (int)ts.size();

  // phase D: owners sum their planes, the received planes and their diagonal blocks
  for (int c = 0; c < n2; ++c)
    for (int r = 0; r < n1; ++r) {
      const bool owner = op->isdiag || (adjoint ? (r == 0) : (c == 0));
      if (!owner) continue;
      Cell& cell = op->cell(r, c);
      RankCtx* rc = ctx->local(cell.rank);
      if (!rc) continue;
      use(*rc);
      DirPlan& plan = adjoint ? cell.adj : cell.fwd;
      if (plan.nloose == 0) continue;
      const int ip = op->isdiag ? r : (adjoint ? r : c);
      const int opart = op->isdiag ? r : (adjoint ? c : r);
      const bool in_owner = op->isdiag || (adjoint ? (c == 0) : (r == 0));
      const int k = r + c * n1;
      const char* inp = in_owner ? in->parts[ip].d : (in_direct[k] ? in_direct[k] : plan.stage_in);
      ++nops;
      launch_counted(ctx, *rc, DJETS_K_FINISH, [&] {
        return launch_finish(op->dtype, plan.d_items, plan.nloose, plan.d_diags, inp, plan.W, plan.R,
                             out->parts[opart].d, rc->stream);
      });
      if (in_from[k] >= 0) stream_after(*rc, *ctx->local(in_from[k]));  // diagonal terms read the input in place
    }
  return nops;
}

void clear_graphs(djets_op op) {
  for (int d = 0; d < 2; ++d) {
    for (auto& g : op->graphs[d])
      if (g.exec) cudaGraphExecDestroy(g.exec);
    op->graphs[d].clear();
    op->seen[d].clear();
  }
}

// The distinct local ranks an operator runs on, the first one being the capture origin.
std::vector<RankCtx*> op_ranks(djets_op op) {
  std::vector<RankCtx*> out;
  for (Cell& c : op->cells)
    if (RankCtx* rc = op->ctx->local(c.rank))
      if (std::find(out.begin(), out.end(), rc) == out.end()) out.push_back(rc);
  return out;
}

// mul!(out, A, in) (adjoint = false) or mul!(out, A', in) (adjoint = true)
void apply(djets_op op, djets_vec out, djets_vec in, bool adjoint) {
  djets_ctx ctx = op->ctx;
  check_vec(op, out, !adjoint, adjoint ? "mul!(m, A', d): m" : "mul!(d, A, m): d");
  check_vec(op, in, adjoint, adjoint ? "mul!(m, A', d): d" : "mul!(d, A, m): m");
  REQUIRE(out != in, DJETS_ERR_INVALID, "mul!: output aliases input");
  settle(out);
  settle(in);
  REQUIRE(!(ctx->capturing && (!op->finalized || op->perfstat)), DJETS_ERR_INVALID,
          "mul! inside a captured sequence needs a finalised operator without perfstat");
  finalize_op(op);
  if (op->graph_epoch != ctx->epoch) {  // a tuning parameter changed since the graphs were captured
    clear_graphs(op);
    op->graph_epoch = ctx->epoch;
  }
  const int dir = adjoint ? 1 : 0;
  // Graph replay pays when an apply is a chain of many small launches and copies (grids on few GPUs,
  // small blocks).  It needs every rank of the operator in this process (NCCL stays outside graphs)
  // and per-launch timing off.  A (vectors, direction) pair is captured the second time it is seen.
  const bool all_local = (int)ctx->ranks.size() == ctx->world;
  // ... and all of them on one device: replaying a graph that spans GPUs costs more host time than the
  // launches it replaces (measured: 8 GPUs, block-diagonal operator, 0.26 ms per step with graph replay
  // against 0.15 ms with direct launches).
  bool one_device = true;
  for (const Cell& c : op->cells)
    if (RankCtx* rc = ctx->local(c.rank)) one_device = one_device && rc->device == ctx->ranks.front().device;
  const bool eligible = ctx->graphs && !ctx->capturing && !ctx->timing && !op->perfstat && all_local && one_device &&
                        !op->graphs_broken && (op->enqueue_ops[dir] == 0 || op->enqueue_ops[dir] >= 3);
  if (!eligible) {
    op->enqueue_ops[dir] = apply_enqueue_timed(op, out, in, adjoint);
    return;
  }
  std::vector<const void*> key;
  for (const VecPart& p : in->parts) key.push_back(p.d);
  for (const VecPart& p : out->parts) key.push_back(p.d);
  std::vector<RankCtx*> ranks = op_ranks(op);
  RankCtx* origin = ranks.front();

  auto replay = [&](djets_op_s::GraphEntry& g) {
    for (size_t k = 1; k < ranks.size(); ++k) stream_after(*ranks[k], *origin);  // earlier work of the other ranks
    use(*origin);
    CK(cudaGraphLaunch(g.exec, origin->stream));
    if (ranks.size() > 1) {
      CK(cudaEventRecord(origin->ev_fork, origin->stream));
      for (size_t k = 1; k < ranks.size(); ++k) {
        use(*ranks[k]);
        CK(cudaStreamWaitEvent(ranks[k]->stream, origin->ev_fork, 0));
      }
    }
    for (int k = 0; k < DJETS_K_COUNT; ++k) origin->launches[k] += g.launches[k];
    g.last_use = ++op->graph_clock;
  };

  for (auto& g : op->graphs[dir])
    if (g.key == key) {
      replay(g);
      return;
    }
  auto& seen = op->seen[dir];
  if (std::find(seen.begin(), seen.end(), key) == seen.end()) {
    if (seen.size() >= 16) seen.erase(seen.begin());
    seen.push_back(key);
    op->enqueue_ops[dir] = apply_enqueue(op, out, in, adjoint);
    return;
  }

  // capture: fork every other rank's stream from the origin, enqueue, join back
  std::vector<std::vector<int64_t>> before;  // the capture launches nothing: counters are restored after it
  for (RankCtx* rc : ranks) before.emplace_back(rc->launches, rc->launches + DJETS_K_COUNT);
  auto restore_counters = [&](int64_t* captured) {
    for (size_t r = 0; r < ranks.size(); ++r)
      for (int k = 0; k < DJETS_K_COUNT; ++k) {
        if (captured) captured[k] += ranks[r]->launches[k] - before[r][k];
        ranks[r]->launches[k] = before[r][k];
      }
  };
  use(*origin);
  CK(cudaStreamBeginCapture(origin->stream, cudaStreamCaptureModeRelaxed));
  cudaGraph_t graph = nullptr;
  try {
    CK(cudaEventRecord(origin->ev_fork, origin->stream));
    for (size_t k = 1; k < ranks.size(); ++k) {
      use(*ranks[k]);
      CK(cudaStreamWaitEvent(ranks[k]->stream, origin->ev_fork, 0));
    }
    apply_enqueue(op, out, in, adjoint);
    for (size_t k = 1; k < ranks.size(); ++k) {
      use(*ranks[k]);
      CK(cudaEventRecord(ranks[k]->ev_fork, ranks[k]->stream));
      use(*origin);
      CK(cudaStreamWaitEvent(origin->stream, ranks[k]->ev_fork, 0));
    }
    use(*origin);
    CK(cudaStreamEndCapture(origin->stream, &graph));
  } catch (...) {
    use(*origin);
    cudaStreamEndCapture(origin->stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    op->graphs_broken = true;  // fall back to direct launches for this operator from now on
    restore_counters(nullptr);
    apply_enqueue(op, out, in, adjoint);
    return;
  }
  djets_op_s::GraphEntry g;
  g.key = key;
  restore_counters(g.launches);
  cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    op->graphs_broken = true;
    apply_enqueue(op, out, in, adjoint);
    return;
  }
  if (op->graphs[dir].size() >= 8) {  // evict the least recently used
    auto lru = std::min_element(op->graphs[dir].begin(), op->graphs[dir].end(),
                                [](const auto& a, const auto& b) { return a.last_use < b.last_use; });
    cudaGraphExecDestroy(lru->exec);
    op->graphs[dir].erase(lru);
  }
  op->graphs[dir].push_back(g);
  replay(op->graphs[dir].back());
}

}  // namespace

extern "C" {

int32_t djets_plan_tiles(int32_t dtype, int64_t m, int64_t n, int32_t adjoint, int64_t* tile_rows, int64_t* tile_cols,
                         int64_t* ntiles) {
  return guard([&] {
    REQUIRE(dtype >= DJETS_F32 && dtype <= DJETS_C128, DJETS_ERR_INVALID, "plan_tiles: bad dtype");
    REQUIRE(m >= 0 && n >= 0, DJETS_ERR_INVALID, "plan_tiles: negative size");
    djets_ctx_s defaults;
    const TileShape s = plan_shape(dtype, m, n, adjoint != 0, defaults.fwd_tx, defaults.fwd_tc, defaults.adj_tc);
    if (tile_rows) *tile_rows = s.tile_rows;
    if (tile_cols) *tile_cols = s.tile_cols;
    if (ntiles) *ntiles = s.ntiles;
  });
}

int32_t djets_plan_chains(int32_t dtype, int32_t adjoint, int64_t nblocks, int64_t longest, int64_t max_rows,
                          int64_t max_cols, int64_t max_bytes, int32_t wave, int32_t chain, int32_t* blocks_per_chain,
                          int32_t* pair) {
  return guard([&] {
    REQUIRE(dtype >= DJETS_F32 && dtype <= DJETS_C128, DJETS_ERR_INVALID, "plan_chains: bad dtype");
    REQUIRE(chain >= 0 && chain <= kChainMax, DJETS_ERR_INVALID, "plan_chains: chain in 0..16");
    const ChainRule r = chain_rule(dtype, adjoint != 0, nblocks, longest, max_rows, max_cols, max_bytes, wave, chain, true);
    if (blocks_per_chain) *blocks_per_chain = (int32_t)r.k;
    if (pair) *pair = r.pair ? 1 : 0;
  });
}

int32_t djets_plan_exchange(int32_t n1, int32_t n2, int32_t isdiag, int32_t adjoint, int32_t phase, int32_t* out6,
                            int32_t capacity, int32_t* count) {
  return guard([&] {
    REQUIRE(n1 >= 1 && n2 >= 1 && (phase == 0 || phase == 1) && count, DJETS_ERR_INVALID, "plan_exchange: bad arguments");
    const std::vector<Message> msgs = plan_exchange(n1, n2, isdiag != 0, adjoint != 0, phase);
    *count = (int32_t)msgs.size();
    if (!out6) return;
    REQUIRE(capacity >= (int32_t)msgs.size(), DJETS_ERR_INVALID, "plan_exchange: output too small");
    for (size_t k = 0; k < msgs.size(); ++k) {
      const Message& m = msgs[k];
      const int32_t v[6] = {m.src_r, m.src_c, m.dst_r, m.dst_c, m.part, m.slot};
      memcpy(out6 + 6 * k, v, sizeof(v));
    }
  });
}

int32_t djets_op_create(djets_ctx ctx, int32_t dtype, int64_t nbrow, int64_t nbcol, const int64_t* row_len,
                        const int64_t* col_len, int32_t n1, int32_t n2, const int32_t* grid_ranks,
                        const int64_t* row_cuts, const int64_t* col_cuts, int32_t isdiag, djets_op* out) {
  return guard([&] {
    REQUIRE(ctx && out && row_len && col_len, DJETS_ERR_INVALID, "op_create: null argument");
    REQUIRE(dtype >= DJETS_F32 && dtype <= DJETS_C128, DJETS_ERR_INVALID, "dtype must be DJETS_F32, F64, C64 or C128");
    REQUIRE(nbrow >= 1 && nbcol >= 1, DJETS_ERR_INVALID, "op_create: the operator needs at least one block");
    REQUIRE(n1 >= 1 && n2 >= 1 && n1 <= nbrow && n2 <= nbcol && n1 * n2 <= ctx->world, DJETS_ERR_INVALID,
            "op_create: process grid does not fit the blocks / the job");
    REQUIRE(n1 <= kMaxParts && n2 <= kMaxParts, DJETS_ERR_INVALID, "op_create: grid too large");
    if (isdiag) {
      REQUIRE(n2 == 1, DJETS_ERR_INVALID, "block diagonal jets must define distribution along rows.");  // src:481
      REQUIRE(nbrow == nbcol, DJETS_ERR_INVALID, "block diagonal operator must be square in blocks");
    }
    auto op = std::make_unique<djets_op_s>();
    op->ctx = ctx;
    op->dtype = dtype;
    op->nbrow = nbrow;
    op->nbcol = nbcol;
    op->row_len.assign(row_len, row_len + nbrow);
    op->col_len.assign(col_len, col_len + nbcol);
    for (int64_t v : op->row_len) REQUIRE(v >= 0 && v < (1LL << 31), DJETS_ERR_INVALID, "bad block row length");
    for (int64_t v : op->col_len) REQUIRE(v >= 0 && v < (1LL << 31), DJETS_ERR_INVALID, "bad block column length");
    op->n1 = n1;
    op->n2 = n2;
    op->isdiag = isdiag != 0;
    op->grid.resize((size_t)n1 * n2);
    std::vector<char> seen(ctx->world, 0);
    for (int k = 0; k < n1 * n2; ++k) {
      op->grid[k] = grid_ranks ? grid_ranks[k] : k;
      REQUIRE(op->grid[k] >= 0 && op->grid[k] < ctx->world && !seen[op->grid[k]], DJETS_ERR_INVALID,
              "op_create: bad or repeated rank in the process grid");
      seen[op->grid[k]] = 1;
    }
    op->row_cuts.resize(n1 + 1);
    op->col_cuts.resize(n2 + 1);
    if (row_cuts)
      std::copy(row_cuts, row_cuts + n1 + 1, op->row_cuts.begin());
    else
      djets_even_cuts(nbrow, n1, op->row_cuts.data());
    if (col_cuts)
      std::copy(col_cuts, col_cuts + n2 + 1, op->col_cuts.begin());
    else
      djets_even_cuts(nbcol, n2, op->col_cuts.data());
    REQUIRE(op->row_cuts[0] == 0 && op->row_cuts[n1] == nbrow && op->col_cuts[0] == 0 && op->col_cuts[n2] == nbcol,
            DJETS_ERR_INVALID, "op_create: cuts must start at 0 and end at the block count");
    for (int r = 0; r < n1; ++r) REQUIRE(op->row_cuts[r] <= op->row_cuts[r + 1], DJETS_ERR_INVALID, "row cuts decrease");
    for (int c = 0; c < n2; ++c) REQUIRE(op->col_cuts[c] <= op->col_cuts[c + 1], DJETS_ERR_INVALID, "col cuts decrease");
    op->blocks.resize((size_t)nbrow * nbcol);
    op->cells.resize((size_t)n1 * n2);
    for (int c = 0; c < n2; ++c)
      for (int r = 0; r < n1; ++r) {
        Cell& cell = op->cell(r, c);
        cell.r = r;
        cell.c = c;
        cell.rank = op->rank_at(r, c);
      }
    // element offsets of every block row inside its range part, of every block column inside its domain part
    op->row_off.resize(nbrow);
    op->rowpart_len.assign(n1, 0);
    for (int r = 0; r < n1; ++r) {
      int64_t o = 0;
      for (int64_t i = op->row_cuts[r]; i < op->row_cuts[r + 1]; ++i) {
        op->row_off[i] = o;
        o += op->row_len[i];
      }
      op->rowpart_len[r] = o;
    }
    op->col_off.resize(nbcol);
    const int ndp = op->ndom_parts();
    const std::vector<int64_t>& dcuts = op->isdiag ? op->row_cuts : op->col_cuts;
    op->colpart_len.assign(ndp, 0);
    for (int c = 0; c < ndp; ++c) {
      int64_t o = 0;
      for (int64_t j = dcuts[c]; j < dcuts[c + 1]; ++j) {
        op->col_off[j] = o;
        o += op->col_len[j];
      }
      op->colpart_len[c] = o;
    }
    *out = op.release();
  });
}

static RankCtx* block_owner_ctx(djets_op op, int64_t i, int64_t j, int* rr, int* cc) {
  REQUIRE(op, DJETS_ERR_INVALID, "null operator");
  REQUIRE(i >= 0 && i < op->nbrow && j >= 0 && j < op->nbcol, DJETS_ERR_INVALID, "block index outside the operator");
  const int r = op->row_cell(i), c = op->col_cell(j);
  if (rr) *rr = r;
  if (cc) *cc = c;
  return op->ctx->local(op->rank_at(r, c));
}

static void release_block(Block& b) {
  if (b.d) cudaFree(b.d);
  b = Block();
}

int32_t djets_op_set_dense(djets_op op, int64_t i, int64_t j, const void* host, int64_t ld) {
  return guard([&] {
    RankCtx* rc = block_owner_ctx(op, i, j, nullptr, nullptr);
    if (!rc) return;
    REQUIRE(!op->finalized, DJETS_ERR_INVALID, "op_set_dense after the operator was finalised");
    const int64_t m = op->row_len[i], n = op->col_len[j];
    REQUIRE(host || m * n == 0, DJETS_ERR_INVALID, "op_set_dense: null payload");
    REQUIRE(ld >= std::max<int64_t>(m, 1), DJETS_ERR_INVALID, "op_set_dense: leading dimension smaller than the block");
    use(*rc);
    Block& b = op->block(i, j);
    release_block(b);
    const int es = elem_size(op->dtype);
    const int64_t bld = round_up(std::max<int64_t>(m, 1), vec_elems(op->dtype));
    const size_t bytes = (size_t)bld * std::max<int64_t>(n, 1) * es;
    void* payload = nullptr;
    CK(device_alloc(op->ctx, &payload, bytes + 256));  // the block changes kind only once its storage exists
    b.kind = DJETS_BLOCK_DENSE;
    b.ld = bld;
    b.d = reinterpret_cast<char*>(payload);
    if (b.ld != m) CK(cudaMemsetAsync(b.d, 0, bytes + 256, rc->stream));
    if (m * n > 0) {
      if (ld == b.ld)  // contiguous on both sides: one linear DMA (much faster than a pitched copy from pageable memory)
        CK(cudaMemcpyAsync(b.d, host, (size_t)m * n * es, cudaMemcpyHostToDevice, rc->stream));
      else
        CK(cudaMemcpy2DAsync(b.d, (size_t)b.ld * es, host, (size_t)ld * es, (size_t)m * es, (size_t)n,
                             cudaMemcpyHostToDevice, rc->stream));
    }
    CK(cudaStreamSynchronize(rc->stream));
  });
}

int32_t djets_op_set_diag(djets_op op, int64_t i, int64_t j, const void* host) {
  return guard([&] {
    RankCtx* rc = block_owner_ctx(op, i, j, nullptr, nullptr);
    REQUIRE(op->row_len[i] == op->col_len[j], DJETS_ERR_MISMATCH, "op_set_diag: diagonal block must be square");
    if (!rc) return;
    REQUIRE(!op->finalized, DJETS_ERR_INVALID, "op_set_diag after the operator was finalised");
    const int64_t n = op->row_len[i];
    REQUIRE(host || n == 0, DJETS_ERR_INVALID, "op_set_diag: null payload");
    use(*rc);
    Block& b = op->block(i, j);
    release_block(b);
    const int es = elem_size(op->dtype);
    void* payload = nullptr;
    CK(device_alloc(op->ctx, &payload, (size_t)std::max<int64_t>(n, 1) * es + 256));
    b.kind = DJETS_BLOCK_DIAG;
    b.ld = n;
    b.d = reinterpret_cast<char*>(payload);
    if (n > 0) CK(cudaMemcpyAsync(b.d, host, (size_t)n * es, cudaMemcpyHostToDevice, rc->stream));
    CK(cudaStreamSynchronize(rc->stream));
  });
}

int32_t djets_op_finalize(djets_op op) {
  return guard([&] {
    REQUIRE(op, DJETS_ERR_INVALID, "null operator");
    finalize_op(op);
  });
}

int32_t djets_op_replan(djets_op op) {
  return guard([&] {
    REQUIRE(op, DJETS_ERR_INVALID, "null operator");
    for (Cell& c : op->cells)
      if (RankCtx* rc = op->ctx->local(c.rank)) {
        use(*rc);
        CK(cudaStreamSynchronize(rc->stream));
      }
    clear_graphs(op);
    op->enqueue_ops[0] = op->enqueue_ops[1] = 0;
    op->finalized = false;
    finalize_op(op);  // build_plan retires published exchange buffers to the context's graveyard
    if (op->xp_tried && (int)op->ctx->ranks.size() < op->ctx->world) {
      xp_release(op);  // collective, like finalize: every process re-publishes its new buffers
      op->xp_tried = false;
      op->xp_epoch[0] = op->xp_epoch[1] = 0;
      xp_setup(op);
    }
  });
}

int32_t djets_op_destroy(djets_op op) {
  return guard([&] {
    if (!op) return;
    for (Cell& c : op->cells) {
      RankCtx* rc = op->ctx->local(c.rank);
      if (!rc) continue;
      cudaSetDevice(rc->device);
      cudaStreamSynchronize(rc->stream);
    }
    clear_graphs(op);
    xp_release(op);
    for (Cell& c : op->cells) {
      RankCtx* rc = op->ctx->local(c.rank);
      if (!rc) continue;
      cudaSetDevice(rc->device);
      if (c.pt0) cudaEventDestroy(c.pt0);
      if (c.pt1) cudaEventDestroy(c.pt1);
      free_plan(c.fwd, op->xp_tried ? &op->ctx->graveyard : nullptr);
      free_plan(c.adj, op->xp_tried ? &op->ctx->graveyard : nullptr);
    }
    for (int64_t i = 0; i < op->nbrow; ++i)
      for (int64_t j = 0; j < op->nbcol; ++j) {
        Block& b = op->block(i, j);
        if (!b.d) continue;
        RankCtx* rc = op->ctx->local(op->rank_at(op->row_cell(i), op->col_cell(j)));
        if (rc) cudaSetDevice(rc->device);
        release_block(b);
      }
    delete op;
  });
}

int32_t djets_op_new_vec(djets_op op, int32_t which, djets_vec* out) {
  return guard([&] {
    REQUIRE(op && out, DJETS_ERR_INVALID, "op_new_vec: null argument");
    REQUIRE(which == 0 || which == 1, DJETS_ERR_INVALID, "op_new_vec: which is 0 (domain) or 1 (range)");
    const bool is_range = which == 1;
    const int np = is_range ? op->n1 : op->ndom_parts();
    const std::vector<int64_t>& cuts = (is_range || op->isdiag) ? op->row_cuts : op->col_cuts;
    std::vector<int32_t> ranks(np);
    std::vector<int64_t> nb(np);
    for (int p = 0; p < np; ++p) {
      ranks[p] = (is_range || op->isdiag) ? op->rank_at(p, 0) : op->rank_at(0, p);
      nb[p] = cuts[p + 1] - cuts[p];
    }
    *out = make_vec(op->ctx, op->dtype, np, ranks.data(), nb.data(),
                    is_range ? op->row_len.data() : op->col_len.data());
  });
}

int32_t djets_op_mul(djets_op op, djets_vec y, djets_vec x) {
  return guard([&] {
    REQUIRE(op && y && x, DJETS_ERR_INVALID, "op_mul: null argument");
    apply(op, y, x, false);
  });
}

int32_t djets_op_mul_adjoint(djets_op op, djets_vec x, djets_vec y) {
  return guard([&] {
    REQUIRE(op && y && x, DJETS_ERR_INVALID, "op_mul_adjoint: null argument");
    apply(op, x, y, true);
  });
}

int32_t djets_op_blockmap(djets_op op, int32_t r, int32_t c, int64_t* row_first, int64_t* row_count,
                          int64_t* col_first, int64_t* col_count) {
  return guard([&] {
    REQUIRE(op && r >= 0 && r < op->n1 && c >= 0 && c < op->n2, DJETS_ERR_INVALID, "blockmap: cell outside the grid");
    if (row_first) *row_first = op->row_cuts[r];
    if (row_count) *row_count = op->row_cuts[r + 1] - op->row_cuts[r];
    if (col_first) *col_first = op->col_cuts[c];
    if (col_count) *col_count = op->col_cuts[c + 1] - op->col_cuts[c];
  });
}

int32_t djets_op_block_owner(djets_op op, int64_t i, int64_t j, int32_t* r, int32_t* c, int32_t* rank) {
  return guard([&] {
    int rr = 0, cc = 0;
    block_owner_ctx(op, i, j, &rr, &cc);
    if (r) *r = rr;
    if (c) *c = cc;
    if (rank) *rank = op->rank_at(rr, cc);
  });
}

int32_t djets_op_getblock(djets_op op, int64_t i, int64_t j, int32_t* kind, void* host_dst, int64_t ld) {
  return guard([&] {
    RankCtx* rc = block_owner_ctx(op, i, j, nullptr, nullptr);
    REQUIRE(rc, DJETS_ERR_NOT_LOCAL, "getblock(A,i,j): block is owned by a rank of another process");
    Block& b = op->block(i, j);
    if (kind) *kind = b.kind;
    if (!host_dst || b.kind == DJETS_BLOCK_ZERO) return;
    use(*rc);
    const int es = elem_size(op->dtype);
    const int64_t m = op->row_len[i], n = op->col_len[j];
    if (b.kind == DJETS_BLOCK_DENSE) {
      REQUIRE(ld >= std::max<int64_t>(m, 1), DJETS_ERR_INVALID, "op_getblock: leading dimension smaller than the block");
      if (m * n > 0)
        CK(cudaMemcpy2DAsync(host_dst, (size_t)ld * es, b.d, (size_t)b.ld * es, (size_t)m * es, (size_t)n,
                             cudaMemcpyDeviceToHost, rc->stream));
    } else if (m > 0) {
      CK(cudaMemcpyAsync(host_dst, b.d, (size_t)m * es, cudaMemcpyDeviceToHost, rc->stream));
    }
    CK(cudaStreamSynchronize(rc->stream));
  });
}

int32_t djets_op_algorithmic_bytes(djets_op op, int64_t* bytes) {
  return guard([&] {
    REQUIRE(op && bytes, DJETS_ERR_INVALID, "null argument");
    const int es = elem_size(op->dtype);
    int64_t total = 0;
    for (int64_t i = 0; i < op->nbrow; ++i)
      for (int64_t j = 0; j < op->nbcol; ++j) {
        if (op->isdiag && i != j) continue;
        const Block& b = op->block(i, j);
        if (!b.d) continue;  // payloads exist only on the ranks this process drives
        if (b.kind == DJETS_BLOCK_DENSE) total += op->row_len[i] * op->col_len[j] * es;
        if (b.kind == DJETS_BLOCK_DIAG) total += op->row_len[i] * es;
      }
    for (int r = 0; r < op->n1; ++r)
      if (op->ctx->local(op->rank_at(r, 0))) total += op->rowpart_len[r] * es;
    for (int c = 0; c < op->ndom_parts(); ++c)
      if (op->ctx->local(op->isdiag ? op->rank_at(c, 0) : op->rank_at(0, c))) total += op->colpart_len[c] * es;
    *bytes = total;
  });
}

int32_t djets_op_perfstat_enable(djets_op op, int32_t enable) {
  return guard([&] {
    REQUIRE(op, DJETS_ERR_INVALID, "null operator");
    op->perfstat = enable != 0;
  });
}

int32_t djets_op_perfstat(djets_op op, int32_t r, int32_t c, double* ms) {
  return guard([&] {
    REQUIRE(op && ms && r >= 0 && r < op->n1 && c >= 0 && c < op->n2, DJETS_ERR_INVALID, "perfstat: bad arguments");
    Cell& cell = op->cell(r, c);
    RankCtx* rc = op->ctx->local(cell.rank);
    REQUIRE(rc, DJETS_ERR_NOT_LOCAL, "perfstat: the cell is driven by another process");
    REQUIRE(cell.ptimed, DJETS_ERR_INVALID, "perfstat: no timed apply yet (djets_op_perfstat_enable first)");
    use(*rc);
    CK(cudaEventSynchronize(cell.pt1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, cell.pt0, cell.pt1));
    *ms = t;
  });
}

int32_t djets_op_schedule_info(djets_op op, int32_t adjoint, int64_t* ntiles, int64_t* nitems) {
  return guard([&] {
    REQUIRE(op, DJETS_ERR_INVALID, "null operator");
    finalize_op(op);
    int64_t t = 0, it = 0;
    for (Cell& c : op->cells) {
      if (!op->ctx->local(c.rank)) continue;
      const DirPlan& p = adjoint ? c.adj : c.fwd;
      t += p.ntiles;
      it += p.nloose;
    }
    if (ntiles) *ntiles = t;
    if (nitems) *nitems = it;
  });
}

}  // extern "C"
