// sm_100a kernels of libdjets_b200: batched block GEMV (forward / adjoint) over a tile table,
// segmented block-row / block-column sum with fused diagonal blocks, deterministic reductions and
// streaming elementwise kernels.  Everything here is HBM-bound CUDA-core work: 128-bit streaming
// loads of the block matrices, the x / y slice of a tile staged in shared memory (1-D TMA bulk copy
// when 16-byte aligned), warp-shuffle reductions, no tensor cores, no floating-point atomics.
//
// Reference arithmetic being replaced (ChevronETC/DistributedJets.jl, src/DistributedJets.jl):
//   forward  d_i  = sum_j A_ij m_j     src:629-639 (inner loop), src:641-654 (row loop)
//   adjoint  m_j  = sum_i A_ij' d_i    src:676-686, src:688-701
//   diagonal d_i  = diag .* m_i        test/runtests.jl:8-12 (JopFoo), dense A*m test/runtests.jl:30-36 (JopBaz)
//   dot/norm/extrema/mapreduce/mean/var partials  src:189-283; fill!/broadcast src:332-376
#include "djets_kernels.h"

#include <math.h>

#include <set>
#include <utility>

namespace djets {

// ------------------------------------------------------------------------------------------------
// 128-bit streaming loads of read-only block payloads.  L1 is bypassed (no reuse inside an SM);
// POLICY 1 additionally tags the lines evict-first in L2 so the small vectors / workspace stay
// resident while the matrices stream through.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct VecT;
template <>
struct VecT<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <>
struct VecT<double> {
  using type = double2;
  static constexpr int N = 2;
};

__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ uint64_t make_evict_last_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// Streaming tiles are tagged evict-first; the last tiles of a launch evict-last, so that the next apply of the operator
// (which walks the tiles in the opposite order) finds them in L2.
__device__ __forceinline__ uint64_t tile_policy(int keep) { return keep ? make_evict_last_policy() : make_evict_first_policy(); }

template <int POLICY>
__device__ __forceinline__ float4 ld_stream(const float4* p, uint64_t pol) {
  float4 r;
  if (POLICY == 1) {
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p), "l"(pol));
  } else {
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
  }
  return r;
}
template <int POLICY>
__device__ __forceinline__ double2 ld_stream(const double2* p, uint64_t pol) {
  double2 r;
  if (POLICY == 1) {
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
        : "=d"(r.x), "=d"(r.y)
        : "l"(p), "l"(pol));
  } else {
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  }
  return r;
}

__device__ __forceinline__ void fma_vec(float (&acc)[4], const float4& a, float x) {
  acc[0] = fmaf(a.x, x, acc[0]);
  acc[1] = fmaf(a.y, x, acc[1]);
  acc[2] = fmaf(a.z, x, acc[2]);
  acc[3] = fmaf(a.w, x, acc[3]);
}
__device__ __forceinline__ void fma_vec(double (&acc)[2], const double2& a, double x) {
  acc[0] = fma(a.x, x, acc[0]);
  acc[1] = fma(a.y, x, acc[1]);
}
__device__ __forceinline__ void dot_vec(float& acc, const float4& a, const float4& y) {
  acc = fmaf(a.x, y.x, acc);
  acc = fmaf(a.y, y.y, acc);
  acc = fmaf(a.z, y.z, acc);
  acc = fmaf(a.w, y.w, acc);
}
__device__ __forceinline__ void dot_vec(double& acc, const double2& a, const double2& y) {
  acc = fma(a.x, y.x, acc);
  acc = fma(a.y, y.y, acc);
}
// complex (interleaved re, im): acc += a * x for the complex rows a thread owns, x = (xr, xi)
__device__ __forceinline__ void cfma_vec(float (&acc)[4], const float4& a, float xr, float xi) {
  acc[0] = fmaf(-a.y, xi, fmaf(a.x, xr, acc[0]));
  acc[1] = fmaf(a.x, xi, fmaf(a.y, xr, acc[1]));
  acc[2] = fmaf(-a.w, xi, fmaf(a.z, xr, acc[2]));
  acc[3] = fmaf(a.z, xi, fmaf(a.w, xr, acc[3]));
}
__device__ __forceinline__ void cfma_vec(double (&acc)[2], const double2& a, double xr, double xi) {
  acc[0] = fma(-a.y, xi, fma(a.x, xr, acc[0]));
  acc[1] = fma(a.x, xi, fma(a.y, xr, acc[1]));
}
// (re, im) += conj(a) * y over the complex rows of one 128-bit load
__device__ __forceinline__ void cdot_vec(float& re, float& im, const float4& a, const float4& y) {
  re = fmaf(a.y, y.y, fmaf(a.x, y.x, re));
  im = fmaf(-a.y, y.x, fmaf(a.x, y.y, im));
  re = fmaf(a.w, y.w, fmaf(a.z, y.z, re));
  im = fmaf(-a.w, y.z, fmaf(a.z, y.w, im));
}
__device__ __forceinline__ void cdot_vec(double& re, double& im, const double2& a, const double2& y) {
  re = fma(a.y, y.y, fma(a.x, y.x, re));
  im = fma(-a.y, y.x, fma(a.x, y.y, im));
}

// ------------------------------------------------------------------------------------------------
// Staging a contiguous slice of a vector into shared memory.  When source and length are 16-byte
// aligned the copy is one 1-D TMA bulk transfer (cp.async.bulk -> SASS UBLKCP) completing on an
// mbarrier; otherwise the CTA copies it with plain loads.  `pad_to` elements past `n` are zeroed.
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: a GEMV kernel lets the next kernel of the stream start as soon as its
// own last CTA is resident (launch_dependents), and itself touches nothing a predecessor may still
// be writing -- vectors, workspace, tickets -- before griddepcontrol.wait.  The block payloads and the
// tile tables are immutable, so the first batch of matrix loads is issued ahead of the wait: launch
// latency, the predecessor's tail and the x staging are hidden behind HBM traffic.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Spin of a whole CTA on the epoch flags of the cross-process exchange (see Prefetch::wait_flag): wrap-safe compare,
// 20 s timeout that raises *err instead of hanging the GPU.
__device__ __forceinline__ void spin_epoch_flag(const unsigned* flag, unsigned value, unsigned* err) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned cur;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(flag) : "memory");
    if ((int)(cur - value) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 20000000000ull) {
      atomicExch(err, 1u);
      break;
    }
    __nanosleep(32);
  }
}
__device__ __forceinline__ void wait_epoch_flags(const Prefetch& pf) {
  if (pf.nwait <= 0) return;
  if (threadIdx.x == 0) spin_epoch_flag(pf.wait_flag[0], pf.wait_value[0], pf.wait_err);
  if (threadIdx.x == 32 && pf.nwait > 1) spin_epoch_flag(pf.wait_flag[1], pf.wait_value[1], pf.wait_err);
  __syncthreads();
}

// DJETS_TRACE: per-CTA time stamps (start, dependency released, input staged, done) of one launch
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int tile, int slot) {
  if (trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)tile * 4 + slot] = t;
  }
}

// L2 prefetch of a contiguous run of an immutable block payload (TMA bulk prefetch, one instruction per
// run; address and size 16-byte aligned).  Issued by the CTAs of a launch's first wave ahead of
// griddepcontrol.wait: while the predecessor kernel drains, the idle HBM bandwidth already pulls this
// kernel's first tiles towards L2.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Asynchronous two-step form used by the persistent tile loops: stage_issue starts the copy of a tile's
// input slice (one elected thread arms the mbarrier and issues the bulk copy; pad elements are zeroed by
// all threads) and returns whether the bulk path applies (CTA-uniform); stage_wait completes it: the bulk
// path polls the mbarrier phase, the fallback copies with plain loads.  Both end in a CTA barrier.
__device__ __forceinline__ void mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ bool stage_bulk_ok(const T* src, int n) {
  const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(T);
  return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0) && bytes >= 512;
}

template <typename T>
__device__ __forceinline__ bool stage_issue(T* __restrict__ dst, const T* __restrict__ src, int n, int pad_to,
                                            uint64_t* bar) {
  const bool bulk = stage_bulk_ok(src, n);
  if (bulk) {
    if (threadIdx.x == 0) {
      const uint32_t b = smem_u32(bar);
      const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(T);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(dst)),
          "l"(src), "r"(bytes), "r"(b)
          : "memory");
    }
    for (int i = n + threadIdx.x; i < pad_to; i += blockDim.x) dst[i] = T(0);
  }
  return bulk;
}

template <typename T>
__device__ __forceinline__ void stage_wait(bool bulk, T* __restrict__ dst, const T* __restrict__ src, int n,
                                           int pad_to, uint64_t* bar, uint32_t parity) {
  if (bulk) {
    const uint32_t b = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(b), "r"(parity)
          : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < pad_to; i += blockDim.x) dst[i] = (i < n) ? src[i] : T(0);
  }
  __syncthreads();
}

// Synchronous form used by the one-tile-per-CTA kernels.
template <typename T>
__device__ __forceinline__ void stage_vector(T* __restrict__ dst, const T* __restrict__ src, int n, int pad_to,
                                             uint64_t* bar) {
  const int tid = threadIdx.x;
  const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(T);
  const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0) && bytes >= 512;
  if (bulk) {
    const uint32_t b = smem_u32(bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(dst)),
          "l"(src), "r"(bytes), "r"(b)
          : "memory");
    }
    for (int i = n + tid; i < pad_to; i += blockDim.x) dst[i] = T(0);
    // all threads wait for the bytes to land (phase 0)
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(b)
          : "memory");
    }
    __syncthreads();
  } else {
    for (int i = tid; i < pad_to; i += blockDim.x) dst[i] = (i < n) ? src[i] : T(0);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Segmented sum of the partial planes of one output chunk, with the planes pushed by peer ranks
// and the diagonal blocks of the block row (column) fused in.  Sums run in a fixed order, in
// double for Float32 vectors.  COHERENT: the planes were written by other CTAs of this launch, so
// they are read with ld.global.cg (L2), never from this SM's L1.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }

template <typename T, bool COHERENT>
__device__ __forceinline__ T ld_plane(const T* p) {
  return COHERENT ? __ldcg(p) : *p;
}

template <typename T>
__device__ __forceinline__ bool aligned16(const T* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// Items made of diagonal blocks only -- the streaming `d .= diag .* m` of a diagonal operator and its
// sums over a block row: 128-bit loads of the diagonals (streamed, L1 bypassed) and of the vector, when every slice is
// 16-byte aligned.  A thread works on VU vectors (256 threads apart) and U terms at a time, VU*U = 8 diagonal loads
// and as many vector loads issued back to back before the first product (loads into arrays: left to itself the
// compiler interleaves load - multiply - load and keeps ONE term in flight per thread; ncu of a 24x24 grid of diagonal
// blocks showed every DFMA waiting on its own pair of loads).  Terms are added in term order, as before.
// Loads that keep their program order (asm volatile): the whole batch is issued before the first use.
__device__ __forceinline__ float4 ld_ordered(const float4* p, bool streamed) {
  float4 r;
  if (streamed)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_ordered(const double2* p, bool streamed) {
  double2 r;
  if (streamed)
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  else
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

template <typename T, int VU, int U>
__device__ __forceinline__ void diag_vec_terms(double (&acc)[VU][VecT<T>::N], const DiagTerm* __restrict__ diags, int d,
                                               const T* __restrict__ in, int v) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  V dv[U][VU], xv[U][VU];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const V* dp = reinterpret_cast<const V*>(diags[d + u].d) + v;
#pragma unroll
    for (int j = 0; j < VU; ++j) dv[u][j] = ld_ordered(dp + j * kThreads, true);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const V* xp = reinterpret_cast<const V*>(in + diags[d + u].vin) + v;
#pragma unroll
    for (int j = 0; j < VU; ++j) xv[u][j] = ld_ordered(xp + j * kThreads, false);
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int j = 0; j < VU; ++j) {
      const T* de = reinterpret_cast<const T*>(&dv[u][j]);
      const T* xe = reinterpret_cast<const T*>(&xv[u][j]);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[j][k] += (double)rn_mul(de[k], xe[k]);  // product rounded in T, like the broadcast `diag .* m`
    }
}

template <typename T, int VU, int U>
__device__ __forceinline__ void diag_vec_span(const FinishItem& it, const DiagTerm* __restrict__ diags,
                                              const T* __restrict__ in, T* __restrict__ o, int v) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  double acc[VU][VEC];
#pragma unroll
  for (int j = 0; j < VU; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[j][k] = 0.0;
  int d = it.diag_begin;
  for (; d + U <= it.diag_end; d += U) diag_vec_terms<T, VU, U>(acc, diags, d, in, v);
  for (; d < it.diag_end; ++d) diag_vec_terms<T, VU, 1>(acc, diags, d, in, v);
#pragma unroll
  for (int j = 0; j < VU; ++j) {
    V r;
    T* re = reinterpret_cast<T*>(&r);
#pragma unroll
    for (int k = 0; k < VEC; ++k) re[k] = (T)acc[j][k];
    reinterpret_cast<V*>(o)[v + j * kThreads] = r;
  }
}

// One term: the streaming `d .= diag .* m`.  The product needs no wide accumulator; `+ 0` keeps the sign of zero the
// accumulating form produces (0 + -0 = +0).
template <typename T, int VU>
__device__ __forceinline__ void diag_vec_single(const DiagTerm& term, const T* __restrict__ in, T* __restrict__ o, int v) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  const V* dp = reinterpret_cast<const V*>(term.d) + v;
  const V* xp = reinterpret_cast<const V*>(in + term.vin) + v;
  V dv[VU], xv[VU];
#pragma unroll
  for (int j = 0; j < VU; ++j) dv[j] = ld_ordered(dp + j * kThreads, true);
#pragma unroll
  for (int j = 0; j < VU; ++j) xv[j] = ld_ordered(xp + j * kThreads, false);
#pragma unroll
  for (int j = 0; j < VU; ++j) {
    const T* de = reinterpret_cast<const T*>(&dv[j]);
    const T* xe = reinterpret_cast<const T*>(&xv[j]);
    V r;
    T* re = reinterpret_cast<T*>(&r);
#pragma unroll
    for (int k = 0; k < VEC; ++k) re[k] = rn_add(rn_mul(de[k], xe[k]), T(0));
    reinterpret_cast<V*>(o)[v + j * kThreads] = r;
  }
}

template <typename T>
__device__ __forceinline__ bool finish_item_diag_vec(const FinishItem& it, const DiagTerm* __restrict__ diags,
                                                     const T* __restrict__ in, T* __restrict__ out) {
  constexpr int VEC = VecT<T>::N;
  T* o = out + it.out;
  bool ok = aligned16(o);
  for (int d = it.diag_begin; d < it.diag_end && ok; ++d)
    ok = aligned16(reinterpret_cast<const T*>(diags[d].d)) && aligned16(in + diags[d].vin);
  if (!ok) return false;  // uniform over the CTA
  const int nv = it.len / VEC;
  int v = threadIdx.x;
  const int nt = it.diag_end - it.diag_begin;
  if (nt == 1) {
    const DiagTerm term = diags[it.diag_begin];
    for (; v + 3 * kThreads < nv; v += 4 * kThreads) diag_vec_single<T, 4>(term, in, o, v);
    for (; v < nv; v += kThreads) diag_vec_single<T, 1>(term, in, o, v);
  } else if (nt >= 4) {
    for (; v + kThreads < nv; v += 2 * kThreads) diag_vec_span<T, 2, 4>(it, diags, in, o, v);
  } else {
    for (; v + kThreads < nv; v += 2 * kThreads) diag_vec_span<T, 2, 2>(it, diags, in, o, v);
  }
  for (; v < nv; v += kThreads) diag_vec_span<T, 1, 4>(it, diags, in, o, v);
  for (int e = nv * VEC + threadIdx.x; e < it.len; e += kThreads) {
    double s = 0.0;
    for (int d = it.diag_begin; d < it.diag_end; ++d)
      s += (double)rn_mul(reinterpret_cast<const T*>(diags[d].d)[e], in[diags[d].vin + e]);
    o[e] = (T)s;
  }
  return true;
}

template <typename T, bool COHERENT, bool CPLX = false>
__device__ __forceinline__ void finish_item(const FinishItem& it, const DiagTerm* __restrict__ diags,
                                            const T* __restrict__ in, const T* W, const T* __restrict__ R,
                                            T* __restrict__ out) {
  if (CPLX) {  // complex: elements are (re, im) pairs of T; planes add component-wise, diagonal terms multiply as complex
    const bool conj = (it.flags & 2) != 0;
    for (int e = 2 * threadIdx.x; e < it.len; e += 2 * kThreads) {
      double sr = 0.0, si = 0.0;
      const T* w = W + it.w0 + e;
      for (int p = 0; p < it.nplanes; ++p) {
        sr += (double)ld_plane<T, COHERENT>(w);
        si += (double)ld_plane<T, COHERENT>(w + 1);
        w += it.wstride;
      }
      for (int q = 0; q < it.nremote; ++q) {
        sr += (double)R[it.r0 + q * it.rstride + e];
        si += (double)R[it.r0 + q * it.rstride + e + 1];
      }
      for (int d = it.diag_begin; d < it.diag_end; ++d) {
        const T* dd = reinterpret_cast<const T*>(diags[d].d);
        const T dr = dd[e], di = conj ? -dd[e + 1] : dd[e + 1];
        const T xr = in[diags[d].vin + e], xi = in[diags[d].vin + e + 1];
        // complex product with separately rounded products and sum, like Julia's and NumPy's (no FMA contraction)
        sr += (double)rn_sub(rn_mul(dr, xr), rn_mul(di, xi));
        si += (double)rn_add(rn_mul(dr, xi), rn_mul(di, xr));
      }
      out[it.out + e] = (T)sr;
      out[it.out + e + 1] = (T)si;
    }
    return;
  }
  if (it.nplanes == 0 && it.nremote == 0 && it.diag_end > it.diag_begin && finish_item_diag_vec<T>(it, diags, in, out))
    return;
  for (int e = threadIdx.x; e < it.len; e += kThreads) {
    double s = 0.0;
    const T* w = W + it.w0 + e;
    int p = 0;
    for (; p + 4 <= it.nplanes; p += 4) {
      const T a0 = ld_plane<T, COHERENT>(w), a1 = ld_plane<T, COHERENT>(w + it.wstride),
              a2 = ld_plane<T, COHERENT>(w + 2 * it.wstride), a3 = ld_plane<T, COHERENT>(w + 3 * it.wstride);
      s += (double)a0;
      s += (double)a1;
      s += (double)a2;
      s += (double)a3;
      w += 4 * it.wstride;
    }
    for (; p < it.nplanes; ++p) {
      s += (double)ld_plane<T, COHERENT>(w);
      w += it.wstride;
    }
    for (int q = 0; q < it.nremote; ++q) s += (double)R[it.r0 + q * it.rstride + e];
    for (int d = it.diag_begin; d < it.diag_end; ++d) {
      const T* dd = reinterpret_cast<const T*>(diags[d].d);
      s += (double)rn_mul(dd[e], in[diags[d].vin + e]);
    }
    out[it.out + e] = (T)s;
  }
}

// Ticket protocol of a fused strip: every contributing CTA publishes its plane (fence), takes a
// ticket, and the CTA holding the last ticket sums all planes and rearms the ticket for the next apply.
template <typename T, bool CPLX = false>
__device__ __forceinline__ void fused_finish(int fin, const FusedFinish& ff, const T* __restrict__ in,
                                             const T* W, T* __restrict__ out) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(ff.tickets + fin, 1u);
    s_last = (old + 1u == (unsigned)ff.items[fin].nplanes) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    finish_item<T, true, CPLX>(ff.items[fin], ff.diags, in, W, nullptr, out);
    if (threadIdx.x == 0) ff.tickets[fin] = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// One tile per CTA (the default): the hardware scheduler hands the tiles out.
// Forward dense tiles: partial[r] = sum_c A[r,c] x[c].  A is column-major, so a thread owns VEC
// consecutive rows (one 128-bit load per column) and walks the columns; the CTA is TX threads wide
// along the rows and TY = 256/TX column groups deep; the TY partial strips are summed in shared
// memory in a fixed order.  U independent 128-bit loads are in flight per thread.
// ------------------------------------------------------------------------------------------------
template <typename T, int TX, int POLICY, bool CPLX = false>
__global__ void __launch_bounds__(kThreads, 4) djets_blockmul_fwd_tile_kernel(const DenseTile* __restrict__ tiles,
                                                          const T* __restrict__ in, T* W, T* __restrict__ out,
                                                          FusedFinish ff, Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int TY = kThreads / TX;
  constexpr int TR = TX * VEC;
  constexpr int U = 8;
  __shared__ __align__(128) T xs[kFwdMaxTC];
  __shared__ __align__(16) T red[(TY > 1) ? TY * TR : 1];
  __shared__ __align__(8) uint64_t bar;

  griddep_launch_dependents();
  trace_stamp(pf.trace, blockIdx.x, 0);
  const DenseTile t = tiles[blockIdx.x];
  const int tid = threadIdx.x;
  const int rows = t.rows, cols = t.cols;

  uint64_t pol = 0;
  if (POLICY == 1) pol = tile_policy(t.keep);

  const int tx = tid % TX, ty = tid / TX;
  const int r0 = tx * VEC;
  T acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = T(0);

  const long long ld = t.ld;
  const T* a = reinterpret_cast<const T*>(t.A) + r0 + (long long)ty * ld;
  const long long cstep = (long long)TY * ld;
  int c = ty;
  // first batch of matrix loads goes out before the input vector is even looked at
  const bool pre = (r0 < rows) && (c + (U - 1) * TY < cols);
  V v0[U];
  if (pre) {
#pragma unroll
    for (int u = 0; u < U; ++u) v0[u] = ld_stream<POLICY>(reinterpret_cast<const V*>(a + u * cstep), pol);
  }
  if ((int)blockIdx.x < pf.ctas) {
    // columns after the pre-loaded batch, one run of `rows` elements per column
    const uint32_t run = ((uint32_t)rows * (uint32_t)sizeof(T)) & ~15u;
    const int ncol = run ? min(cols - U * TY, (int)(pf.bytes / run)) : 0;
    const T* base = reinterpret_cast<const T*>(t.A) + (long long)(U * TY) * ld;
    for (int cc = tid; cc < ncol; cc += kThreads) prefetch_l2_bulk(base + (long long)cc * ld, run);
  }
  griddep_wait();
  wait_epoch_flags(pf);
  trace_stamp(pf.trace, blockIdx.x, 1);
  stage_vector<T>(xs, in + t.vin, CPLX ? 2 * cols : cols, CPLX ? 2 * cols : cols, &bar);
  trace_stamp(pf.trace, blockIdx.x, 2);

  if (r0 < rows) {
    if (pre) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (CPLX)
          cfma_vec(acc, v0[u], xs[2 * (c + u * TY)], xs[2 * (c + u * TY) + 1]);
        else
          fma_vec(acc, v0[u], xs[c + u * TY]);
      }
      a += U * cstep;
      c += U * TY;
    }
    for (; c + (U - 1) * TY < cols; c += U * TY) {
      V v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_stream<POLICY>(reinterpret_cast<const V*>(a + u * cstep), pol);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (CPLX)
          cfma_vec(acc, v[u], xs[2 * (c + u * TY)], xs[2 * (c + u * TY) + 1]);
        else
          fma_vec(acc, v[u], xs[c + u * TY]);
      }
      a += U * cstep;
    }
    for (; c < cols; c += TY) {
      V v = ld_stream<POLICY>(reinterpret_cast<const V*>(a), pol);
      if (CPLX)
        cfma_vec(acc, v, xs[2 * c], xs[2 * c + 1]);
      else
        fma_vec(acc, v, xs[c]);
      a += cstep;
    }
  }

  T* dst = (t.direct ? out : W) + t.wout;
  if (TY > 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) red[ty * TR + r0 + k] = acc[k];
    __syncthreads();
    for (int r = tid; r < rows; r += kThreads) {  // rows <= TR
      T s = red[r];
#pragma unroll
      for (int y = 1; y < TY; ++y) s += red[y * TR + r];
      dst[r] = s;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (r0 + k < rows) dst[r0 + k] = acc[k];
  }
  if (t.fin >= 0) fused_finish<T, CPLX>(t.fin, ff, in, W, out);
  trace_stamp(pf.trace, blockIdx.x, 3);
}

// ------------------------------------------------------------------------------------------------
// Adjoint dense tiles: partial[c] = sum_r A[r,c] y[r].  A warp owns CW columns at a time and
// streams down them with 128-bit loads (each lane VEC rows, 32 lanes = 512 contiguous bytes per
// column), RU row chunks unrolled -> CW*RU = 16 independent loads in flight per lane whatever the
// column height: (RU,CW) = (4,4) for tall columns, (2,8) and (1,16) for short ones; the y slice is
// staged once per CTA in shared memory and each 128-bit LDS of y is reused for the CW columns; the
// per-lane sums are combined with a column-halving shuffle butterfly.
// ------------------------------------------------------------------------------------------------
template <typename T, int RU, int CW, int POLICY>
__global__ void __launch_bounds__(kThreads, (RU * CW <= 8) ? 3 : 2) djets_blockmul_adj_tile_kernel(const DenseTile* __restrict__ tiles,
                                                             const T* __restrict__ in, T* W, T* __restrict__ out,
                                                             FusedFinish ff, Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int S = 32 * VEC;  // rows one warp covers with one load per lane
  __shared__ __align__(128) T ys[kAdjMaxTRBytes / sizeof(T)];
  __shared__ __align__(8) uint64_t bar;

  griddep_launch_dependents();
  trace_stamp(pf.trace, blockIdx.x, 0);
  const DenseTile t = tiles[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = t.rows, cols = t.cols;
  const int rowsv = (rows + VEC - 1) / VEC * VEC;  // pad rows of A are zero; pad y with zeros too
  if ((int)blockIdx.x < pf.ctas) {
    // the columns of the first pass over the warps, below the first batch of loads: one run per column
    const int ncol = min(cols, (kThreads / 32) * CW);
    const int below = rowsv - RU * S;
    const uint32_t want = ncol ? (uint32_t)(pf.bytes / ncol) & ~15u : 0u;
    const uint32_t run = below > 0 ? min(want, ((uint32_t)below * (uint32_t)sizeof(T)) & ~15u) : 0u;
    const T* base = reinterpret_cast<const T*>(t.A) + RU * S;
    if (run)
      for (int cc = tid; cc < ncol; cc += kThreads) prefetch_l2_bulk(base + (long long)cc * t.ld, run);
  }
  griddep_wait();
  wait_epoch_flags(pf);
  trace_stamp(pf.trace, blockIdx.x, 1);
  stage_vector<T>(ys, in + t.vin, rows, rowsv, &bar);
  trace_stamp(pf.trace, blockIdx.x, 2);

  uint64_t pol = 0;
  if (POLICY == 1) pol = tile_policy(t.keep);

  const long long ld = t.ld;
  const T* A = reinterpret_cast<const T*>(t.A);
  T* dst = (t.direct ? out : W) + t.wout;
  const int ngroups = (cols + CW - 1) / CW;

  for (int g = warp; g < ngroups; g += kThreads / 32) {
    const int c0 = g * CW;
    const int last = cols - 1 - c0;  // columns past the tile edge re-read the last one and are not stored
    const T* base = A + (long long)c0 * ld + lane * VEC;
    T acc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) acc[k] = T(0);

    int rb = 0;
    for (; rb + RU * S <= rowsv; rb += RU * S) {
      V a[RU][CW];
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int k = 0; k < CW; ++k)
          a[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
        for (int k = 0; k < CW; ++k) dot_vec(acc[k], a[u][k], y);
      }
    }
    if (CW == 4) {
      // tall columns: the remainder is at most RU-1 chunks: pairs of chunks first, then a guarded last one
      if (RU > 2) {
        for (; rb + 2 * S <= rowsv; rb += 2 * S) {
          V a[2][CW];
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < CW; ++k)
              a[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
            for (int k = 0; k < CW; ++k) dot_vec(acc[k], a[u][k], y);
          }
        }
      }
      for (; rb < rowsv; rb += S) {
        if (rb + lane * VEC < rowsv) {
          V a[CW];
#pragma unroll
          for (int k = 0; k < CW; ++k)
            a[k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb), pol);
          const V y = *reinterpret_cast<const V*>(ys + rb + lane * VEC);
#pragma unroll
          for (int k = 0; k < CW; ++k) dot_vec(acc[k], a[k], y);
        }
      }
    } else if (rb < rowsv) {
      // short columns: the remainder IS most of the column -- one guarded batch, CW*RU loads in flight
      V a[RU][CW];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const bool ok = rb + u * S + lane * VEC < rowsv;
#pragma unroll
        for (int k = 0; k < CW; ++k) {
          if (ok)
            a[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (rb + u * S + lane * VEC < rowsv) {
          const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
          for (int k = 0; k < CW; ++k) dot_vec(acc[k], a[u][k], y);
        }
      }
    }
    // Warp reduction of CW columns at once: while more than one column is live, each butterfly step
    // also halves the columns a lane keeps (the lanes with the offset bit set keep the upper half), so
    // CW columns cost CW-1 + (5 - log2 CW) shuffles instead of 5*CW.  The pattern is fixed: deterministic.
    int width = CW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (width > 1) {
        width >>= 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
          if (k < width) {
            const T send = upper ? acc[k] : acc[k + width];
            const T keep = upper ? acc[k + width] : acc[k];
            acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
      }
    }
    // lane bits 16, 8, ... (one per halving step) spell the column a lane ended up with
    constexpr int kHalvings = (CW == 16) ? 4 : (CW == 8) ? 3 : (CW == 4) ? 2 : (CW == 2) ? 1 : 0;
    int mycol = 0;
#pragma unroll
    for (int h = 0; h < kHalvings; ++h) mycol |= ((lane >> (4 - h)) & 1) << (kHalvings - 1 - h);
    const bool writer = (lane & ((32 >> kHalvings) - 1)) == 0;
    if (writer && c0 + mycol < cols) dst[c0 + mycol] = acc[0];
  }
  if (t.fin >= 0) fused_finish<T>(t.fin, ff, in, W, out);
  trace_stamp(pf.trace, blockIdx.x, 3);
}

// ------------------------------------------------------------------------------------------------
// Adjoint over chains of small blocks (every block at most 2 warp-loads tall: 128 Float64 / 256 Float32 rows).
// A tile is a head block plus up to kChainMax-1 further blocks of the same block column (ChainLink); a warp owns CW
// columns and walks down the chain with them, 2*CW 128-bit loads in flight per lane -- the two row chunks of one
// block, or (PAIR: blocks of one warp-load) the single chunks of two blocks -- and reduces / stores once per chain.
// The first batch is issued before the dependency wait and the staging of the y slices (plain coalesced loads into
// zero-padded shared-memory rows, one per block), so the fixed costs of a CTA (descriptor, staging, ticket) are paid
// once per chain and overlap its first loads.
// ------------------------------------------------------------------------------------------------
template <typename T, int CW, int POLICY>
__device__ __forceinline__ void chain_load_slot(typename VecT<T>::type (&a)[CW], const T* base, long long ld, int last,
                                                bool ok, uint64_t pol) {
  using V = typename VecT<T>::type;
#pragma unroll
  for (int k = 0; k < CW; ++k) {
    if (ok)
      a[k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld), pol);
    else
      a[k] = V{};
  }
}

template <typename T, int CW, bool PAIR, int POLICY>
__global__ void __launch_bounds__(kThreads, 2) djets_blockmul_adj_chain_kernel(const DenseTile* __restrict__ tiles,
                                                                               const ChainLink* __restrict__ links,
                                                                               const T* __restrict__ in, T* W,
                                                                               T* __restrict__ out, FusedFinish ff,
                                                                               Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int S = 32 * VEC;
  constexpr int YS = PAIR ? S : 2 * S;  // one zero-padded row of shared memory per block of the chain
  __shared__ __align__(16) T ys[kChainMax * YS];
  __shared__ ChainLink sl[kChainMax];

  griddep_launch_dependents();
  trace_stamp(pf.trace, blockIdx.x, 0);
  const DenseTile t = tiles[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cols = t.cols, nblk = 1 + t.chain;
  const int ngroups = (cols + CW - 1) / CW;
  uint64_t pol = 0;
  if (POLICY == 1) pol = tile_policy(t.keep);

  // slot u of a batch: row chunk u of block l, or (PAIR) the only chunk of block l + u
  V a[2][CW];
  auto load_batch = [&](int c0, const void* A0, long long ld0, int n0, const void* A1, long long ld1, int n1) {
    const int last = cols - 1 - c0;
    const T* base = reinterpret_cast<const T*>(A0) + (long long)c0 * ld0 + lane * VEC;
    chain_load_slot<T, CW, POLICY>(a[0], base, ld0, last, lane * VEC < n0, pol);
    if (PAIR)  // n1 == 0: there is no second block
      chain_load_slot<T, CW, POLICY>(a[1], reinterpret_cast<const T*>(A1) + (long long)c0 * ld1 + lane * VEC, ld1, last,
                                     lane * VEC < n1, pol);
    else
      chain_load_slot<T, CW, POLICY>(a[1], base + S, ld0, last, S + lane * VEC < n0, pol);
  };
  // the first batch goes out before anything else is looked at (payloads and tables are immutable)
  const bool pre = warp < ngroups;
  if (pre) {
    const void* A1 = t.A;
    long long ld1 = t.ld;
    int n1 = 0;
    if (PAIR && t.chain > 0) {
      A1 = links[t.link0].A;
      ld1 = links[t.link0].ld;
      n1 = links[t.link0].n;
    }
    load_batch(warp * CW, t.A, t.ld, t.rows, A1, ld1, n1);
  }
  if (tid == 0) sl[0] = ChainLink{t.A, t.ld, t.vin, t.rows, 0};
  for (int l = tid; l < t.chain; l += kThreads) sl[1 + l] = links[t.link0 + l];
  griddep_wait();
  wait_epoch_flags(pf);
  __syncthreads();
  trace_stamp(pf.trace, blockIdx.x, 1);
  for (int idx = tid; idx < nblk * YS; idx += kThreads) {
    const int l = idx / YS, e = idx % YS;
    ys[idx] = e < sl[l].n ? in[sl[l].vin + e] : T(0);
  }
  __syncthreads();
  trace_stamp(pf.trace, blockIdx.x, 2);

  T* dst = (t.direct ? out : W) + t.wout;
  for (int g = warp; g < ngroups; g += kThreads / 32) {
    const int c0 = g * CW;
    T acc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) acc[k] = T(0);
    for (int l = 0; l < nblk; l += PAIR ? 2 : 1) {
      if (!(l == 0 && g == warp)) {
        const int l1 = (PAIR && l + 1 < nblk) ? l + 1 : l;
        load_batch(c0, sl[l].A, sl[l].ld, sl[l].n, sl[l1].A, sl[l1].ld, l1 != l ? sl[l1].n : 0);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (PAIR && l + u >= nblk) break;  // no such block: its shared-memory row was never written
        const V y = *reinterpret_cast<const V*>(ys + (PAIR ? (l + u) * YS : l * YS + u * S) + lane * VEC);
#pragma unroll
        for (int k = 0; k < CW; ++k) dot_vec(acc[k], a[u][k], y);
      }
    }
    // the column-halving butterfly of the tile kernel
    int width = CW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (width > 1) {
        width >>= 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
          if (k < width) {
            const T send = upper ? acc[k] : acc[k + width];
            const T keep = upper ? acc[k + width] : acc[k];
            acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
      }
    }
    constexpr int kHalvings = (CW == 16) ? 4 : (CW == 8) ? 3 : (CW == 4) ? 2 : (CW == 2) ? 1 : 0;
    int mycol = 0;
#pragma unroll
    for (int h = 0; h < kHalvings; ++h) mycol |= ((lane >> (4 - h)) & 1) << (kHalvings - 1 - h);
    const bool writer = (lane & ((32 >> kHalvings) - 1)) == 0;
    if (writer && c0 + mycol < cols) dst[c0 + mycol] = acc[0];
  }
  if (t.fin >= 0) fused_finish<T>(t.fin, ff, in, W, out);
  trace_stamp(pf.trace, blockIdx.x, 3);
}

// ------------------------------------------------------------------------------------------------
// Forward over chains of small blocks: a head block plus further blocks of the same block row.  The columns of
// the chain are laid end to end (at most kFwdMaxTC): their x slices are staged side by side in shared memory next
// to a table of column addresses, and the CTA walks the concatenated columns exactly like the tile kernel walks
// the columns of one tile (a thread owns VEC rows, TY column groups, U loads in flight), so block boundaries cost
// nothing.  One partial per chain.
// ------------------------------------------------------------------------------------------------
template <typename T, int TX, int POLICY>
__global__ void __launch_bounds__(kThreads, 3) djets_blockmul_fwd_chain_kernel(const DenseTile* __restrict__ tiles,
                                                                               const ChainLink* __restrict__ links,
                                                                               const T* __restrict__ in, T* W,
                                                                               T* __restrict__ out, FusedFinish ff,
                                                                               Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int TY = kThreads / TX;
  constexpr int TR = TX * VEC;
  constexpr int U = 8;
  __shared__ __align__(16) T xs[kFwdMaxTC];
  __shared__ const T* colp[kFwdMaxTC];  // address of row 0 of every column of the chain
  __shared__ __align__(16) T red[TY * TR];
  __shared__ ChainLink sl[kChainMax];
  __shared__ int first_col[kChainMax + 1];  // position of every block's first column in the chain

  griddep_launch_dependents();
  trace_stamp(pf.trace, blockIdx.x, 0);
  const DenseTile t = tiles[blockIdx.x];
  const int tid = threadIdx.x;
  const int rows = t.rows, nblk = 1 + t.chain;
  uint64_t pol = 0;
  if (POLICY == 1) pol = tile_policy(t.keep);

  const int tx = tid % TX, ty = tid / TX;
  const int r0 = tx * VEC;
  T acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = T(0);

  // first batch: columns of the head block, issued before anything else is looked at
  const bool pre = (r0 < rows) && (ty + (U - 1) * TY < t.cols);
  V v0[U];
  if (pre) {
    const T* a = reinterpret_cast<const T*>(t.A) + r0 + (long long)ty * t.ld;
#pragma unroll
    for (int u = 0; u < U; ++u) v0[u] = ld_stream<POLICY>(reinterpret_cast<const V*>(a + (long long)(u * TY) * t.ld), pol);
  }
  if (tid == 0) {
    sl[0] = ChainLink{t.A, t.ld, t.vin, t.cols, 0};
    first_col[0] = 0;
  }
  for (int l = tid; l < t.chain; l += kThreads) sl[1 + l] = links[t.link0 + l];
  griddep_wait();
  wait_epoch_flags(pf);
  __syncthreads();
  if (tid == 0) {
    int p = 0;
    for (int l = 0; l < nblk; ++l) {
      first_col[l] = p;
      p += sl[l].n;
    }
    first_col[nblk] = p;
  }
  __syncthreads();
  trace_stamp(pf.trace, blockIdx.x, 1);
  const int ncols = first_col[nblk];
  for (int f = tid; f < ncols; f += kThreads) {
    int l = 0;
    while (f >= first_col[l + 1]) ++l;
    const int c = f - first_col[l];
    xs[f] = in[sl[l].vin + c];
    colp[f] = reinterpret_cast<const T*>(sl[l].A) + (long long)c * sl[l].ld;
  }
  __syncthreads();
  trace_stamp(pf.trace, blockIdx.x, 2);

  if (r0 < rows) {
    int f = ty;
    if (pre) {
#pragma unroll
      for (int u = 0; u < U; ++u) fma_vec(acc, v0[u], xs[f + u * TY]);
      f += U * TY;
    }
    for (; f + (U - 1) * TY < ncols; f += U * TY) {
      V v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_stream<POLICY>(reinterpret_cast<const V*>(colp[f + u * TY] + r0), pol);
#pragma unroll
      for (int u = 0; u < U; ++u) fma_vec(acc, v[u], xs[f + u * TY]);
    }
    for (; f < ncols; f += TY) {
      const V v = ld_stream<POLICY>(reinterpret_cast<const V*>(colp[f] + r0), pol);
      fma_vec(acc, v, xs[f]);
    }
  }

  T* dst = (t.direct ? out : W) + t.wout;
#pragma unroll
  for (int k = 0; k < VEC; ++k) red[ty * TR + r0 + k] = acc[k];
  __syncthreads();
  for (int r = tid; r < rows; r += kThreads) {  // rows <= TR
    T s = red[r];
#pragma unroll
    for (int y = 1; y < TY; ++y) s += red[y * TR + r];
    dst[r] = s;
  }
  if (t.fin >= 0) fused_finish<T>(t.fin, ff, in, W, out);
  trace_stamp(pf.trace, blockIdx.x, 3);
}

// ------------------------------------------------------------------------------------------------
// Persistent tile loop shared by both GEMV kernels.  The grid is one resident wave of CTAs (or fewer);
// CTA b starts on tile b and then draws further tiles from an atomic counter, so tiles are handed out
// in table order -- largest first, the last wave's worth cut finer by the scheduler -- to whichever CTA
// frees up.  While a tile is being processed the next one is already known: its input slice is staged
// into the other shared-memory buffer by a bulk copy, and (forward) its first batch of matrix loads is
// issued before the current tile's epilogue.  Draws run two tiles ahead, so the atomic's round trip is never
// waited for; every CTA makes exactly two draws past the table, and the CTA that makes the last draw of
// the launch (ntiles + gridDim.x in total) rearms the counter for the next launch.
// ------------------------------------------------------------------------------------------------
struct TileQueue {
  int ntiles;
  unsigned* counter;  // zero between launches
};

// Forward dense tiles: partial[r] = sum_c A[r,c] x[c].  A is column-major, so a thread owns VEC
// consecutive rows (one 128-bit load per column) and walks the columns; the CTA is TX threads wide
// along the rows and TY = 256/TX column groups deep; the TY partial strips are summed in shared
// memory in a fixed order.  U independent 128-bit loads are in flight per thread.
template <typename T, int TX, int POLICY>
__global__ void __launch_bounds__(kThreads, 4) djets_blockmul_fwd_queue_kernel(const DenseTile* __restrict__ tiles, TileQueue q,
                                                             const T* __restrict__ in, T* W, T* __restrict__ out,
                                                             FusedFinish ff, Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int TY = kThreads / TX;
  constexpr int TR = TX * VEC;
  constexpr int U = 8;
  __shared__ __align__(128) T xs[2][kFwdMaxTC];
  __shared__ __align__(16) T red[(TY > 1) ? TY * TR : 1];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int s_next[2];

  griddep_launch_dependents();
  int cur = blockIdx.x;  // descriptors stay in memory (L1-resident table); only what the hot loop needs is in registers
  trace_stamp(pf.trace, cur, 0);
  const int tid = threadIdx.x;

  uint64_t pol = 0;
  if (POLICY == 1) pol = make_evict_first_policy();

  const int tx = tid % TX, ty = tid / TX;
  const int r0 = tx * VEC;

  // Ahead of the dependency wait the first-wave CTAs pull the head of their first tile towards L2 (the block
  // payloads are immutable, so this is legal before the predecessor kernel has finished): one run of
  // `rows` elements per column.
  if ((int)blockIdx.x < pf.ctas) {
    const DenseTile& t = tiles[cur];
    const uint32_t run = ((uint32_t)t.rows * (uint32_t)sizeof(T)) & ~15u;
    const int ncol = run ? min(t.cols, (int)(pf.bytes / run)) : 0;
    const T* base = reinterpret_cast<const T*>(t.A);
    for (int cc = tid; cc < ncol; cc += kThreads) prefetch_l2_bulk(base + (long long)cc * t.ld, run);
  }
  if (tid == 0) {
    mbar_init(&bar[0]);
    mbar_init(&bar[1]);
    mbar_fence_init();
  }
  griddep_wait();
  wait_epoch_flags(pf);
  trace_stamp(pf.trace, cur, 1);
  __syncthreads();  // barriers initialised
  // bit 0: staging buffer in use; bits 1,2: phase parity of bar[0], bar[1]; bit 3: current tile staged by bulk copy
  unsigned state = stage_issue<T>(xs[0], in + tiles[cur].vin, tiles[cur].cols, tiles[cur].cols, &bar[0]) ? 8u : 0u;
  // Two draws ahead: the tile after `cur` is known at the top of an iteration (its x slice is staged while
  // `cur` is processed); the draw issued at the top returns during the tile and is published at its end.
  if (tid == 0) s_next[0] = (int)gridDim.x + (int)atomicAdd(q.counter, 1u);
  int draw = 0;

  for (int it = 0;; ++it) {
    const int buf = state & 1u;
    if (tid == 0) draw = (int)gridDim.x + (int)atomicAdd(q.counter, 1u);
    {
      const DenseTile& t = tiles[cur];
      stage_wait<T>((state & 8u) != 0u, xs[buf], in + t.vin, t.cols, t.cols, &bar[buf], (state >> (1 + buf)) & 1u);
    }
    if (state & 8u) state ^= 2u << buf;
    trace_stamp(pf.trace, cur, 2);
    const int nxt = s_next[it & 1];
    const bool more = nxt < q.ntiles;
    state &= ~8u;
    if (more) {
      const DenseTile& g = tiles[nxt];
      if (stage_issue<T>(xs[buf ^ 1], in + g.vin, g.cols, g.cols, &bar[buf ^ 1])) state |= 8u;
    }

    T acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = T(0);
    {
      const int rows = tiles[cur].rows, cols = tiles[cur].cols;
      const long long cstep = (long long)TY * tiles[cur].ld;
      const T* x = xs[buf];
      if (r0 < rows) {
        const T* a = reinterpret_cast<const T*>(tiles[cur].A) + r0 + (long long)ty * tiles[cur].ld;
        int c = ty;
        for (; c + (U - 1) * TY < cols; c += U * TY) {
          V v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = ld_stream<POLICY>(reinterpret_cast<const V*>(a + u * cstep), pol);
#pragma unroll
          for (int u = 0; u < U; ++u) fma_vec(acc, v[u], x[c + u * TY]);
          a += U * cstep;
        }
        for (; c < cols; c += TY) {
          V v = ld_stream<POLICY>(reinterpret_cast<const V*>(a), pol);
          fma_vec(acc, v, x[c]);
          a += cstep;
        }
      }
    }
    {
      const DenseTile& t = tiles[cur];
      const int rows = t.rows;
      T* dst = (t.direct ? out : W) + t.wout;
      if (TY > 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) red[ty * TR + r0 + k] = acc[k];
        __syncthreads();
        for (int r = tid; r < rows; r += kThreads) {  // rows <= TR
          T sum = red[r];
#pragma unroll
          for (int y = 1; y < TY; ++y) sum += red[y * TR + r];
          dst[r] = sum;
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (r0 + k < rows) dst[r0 + k] = acc[k];
      }
      const int fin = t.fin;
      if (fin >= 0) fused_finish<T>(fin, ff, in, W, out);
    }
    trace_stamp(pf.trace, cur, 3);
    if (tid == 0) {
      s_next[(it + 1) & 1] = draw;
      // every CTA makes two draws past the table: ntiles + gridDim.x draws per launch, the last one rearms the queue
      if (draw == 2 * (int)gridDim.x + q.ntiles - 1) *q.counter = 0u;
    }
    if (!more) break;
    __syncthreads();  // `red` and the finished tile's x buffer are free again
    cur = nxt;
    state ^= 1u;
    trace_stamp(pf.trace, cur, 0);
    trace_stamp(pf.trace, cur, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Adjoint dense tiles: partial[c] = sum_r A[r,c] y[r].  A warp owns CW columns at a time and
// streams down them with 128-bit loads (each lane VEC rows, 32 lanes = 512 contiguous bytes per
// column), RU row chunks unrolled -> CW*RU = 16 independent loads in flight per lane whatever the
// column height: (RU,CW) = (4,4) for tall columns, (2,8) and (1,16) for short ones; the y slice is
// staged once per tile in shared memory (double-buffered across tiles) and each 128-bit LDS of y is
// reused for the CW columns; the per-lane sums are combined with a column-halving shuffle butterfly.
// ------------------------------------------------------------------------------------------------
// Complex adjoint tiles: partial[c] = sum_r conj(A[r,c]) y[r] on interleaved (re, im) data.  Tile extents: `rows` and
// `ld` count reals (two per complex row), `cols` counts complex columns, outputs are 2*cols reals.  A warp owns CWC
// complex columns at a time (2*CWC accumulators), RU row chunks unrolled; same staging and butterfly as the real kernel.
template <typename T, int RU, int CWC, int POLICY>
__global__ void __launch_bounds__(kThreads, 2) djets_blockmul_adjc_tile_kernel(const DenseTile* __restrict__ tiles,
                                                                   const T* __restrict__ in, T* W, T* __restrict__ out,
                                                                   FusedFinish ff, Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int S = 32 * VEC;
  constexpr int CW = 2 * CWC;  // real outputs per warp pass
  __shared__ __align__(128) T ys[kAdjMaxTRBytes / sizeof(T)];
  __shared__ __align__(8) uint64_t bar;

  griddep_launch_dependents();
  const DenseTile t = tiles[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = t.rows, cols = t.cols;
  const int rowsv = (rows + VEC - 1) / VEC * VEC;
  griddep_wait();
  wait_epoch_flags(pf);
  stage_vector<T>(ys, in + t.vin, rows, rowsv, &bar);

  uint64_t pol = 0;
  if (POLICY == 1) pol = tile_policy(t.keep);

  const long long ld = t.ld;
  const T* A = reinterpret_cast<const T*>(t.A);
  T* dst = (t.direct ? out : W) + t.wout;
  const int ngroups = (cols + CWC - 1) / CWC;

  for (int g = warp; g < ngroups; g += kThreads / 32) {
    const int c0 = g * CWC;
    const int last = cols - 1 - c0;
    const T* base = A + (long long)c0 * ld + lane * VEC;
    T acc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) acc[k] = T(0);

    int rb = 0;
    for (; rb + RU * S <= rowsv; rb += RU * S) {
      V a[RU][CWC];
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int k = 0; k < CWC; ++k)
          a[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
        for (int k = 0; k < CWC; ++k) cdot_vec(acc[2 * k], acc[2 * k + 1], a[u][k], y);
      }
    }
    if (rb < rowsv) {  // one guarded batch for the remainder
      V a[RU][CWC];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const bool ok = rb + u * S + lane * VEC < rowsv;
#pragma unroll
        for (int k = 0; k < CWC; ++k) {
          if (ok)
            a[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (rb + u * S + lane * VEC < rowsv) {
          const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
          for (int k = 0; k < CWC; ++k) cdot_vec(acc[2 * k], acc[2 * k + 1], a[u][k], y);
        }
      }
    }
    int width = CW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (width > 1) {
        width >>= 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
          if (k < width) {
            const T send = upper ? acc[k] : acc[k + width];
            const T keep = upper ? acc[k + width] : acc[k];
            acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
      }
    }
    constexpr int kHalvings = (CW == 16) ? 4 : (CW == 8) ? 3 : (CW == 4) ? 2 : (CW == 2) ? 1 : 0;
    int mycol = 0;
#pragma unroll
    for (int h = 0; h < kHalvings; ++h) mycol |= ((lane >> (4 - h)) & 1) << (kHalvings - 1 - h);
    const bool writer = (lane & ((32 >> kHalvings) - 1)) == 0;
    if (writer && 2 * c0 + mycol < 2 * cols) dst[2 * c0 + mycol] = acc[0];
  }
  if (t.fin >= 0) fused_finish<T, true>(t.fin, ff, in, W, out);
}

template <typename T, int RU, int CW, int POLICY>
__global__ void __launch_bounds__(kThreads, 2) djets_blockmul_adj_queue_kernel(const DenseTile* __restrict__ tiles, TileQueue q,
                                                             const T* __restrict__ in, T* W, T* __restrict__ out,
                                                             FusedFinish ff, Prefetch pf) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int S = 32 * VEC;  // rows one warp covers with one load per lane
  constexpr int YN = kAdjMaxTRBytes / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* ys0 = reinterpret_cast<T*>(smem_raw);
  T* ys1 = ys0 + YN;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ys1 + YN);
  int* s_next = reinterpret_cast<int*>(bar + 2);

  griddep_launch_dependents();
  int cur = blockIdx.x;
  trace_stamp(pf.trace, cur, 0);
  DenseTile t = tiles[cur];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x < pf.ctas) {
    // the columns of the first pass over the warps, below the first batch of loads: one run per column
    const int rowsv = (t.rows + VEC - 1) / VEC * VEC;
    const int ncol = min(t.cols, (kThreads / 32) * CW);
    const int below = rowsv - RU * S;
    const uint32_t want = ncol ? (uint32_t)(pf.bytes / ncol) & ~15u : 0u;
    const uint32_t run = below > 0 ? min(want, ((uint32_t)below * (uint32_t)sizeof(T)) & ~15u) : 0u;
    const T* base = reinterpret_cast<const T*>(t.A) + RU * S;
    if (run)
      for (int cc = tid; cc < ncol; cc += kThreads) prefetch_l2_bulk(base + (long long)cc * t.ld, run);
  }
  if (tid == 0) {
    mbar_init(&bar[0]);
    mbar_init(&bar[1]);
    mbar_fence_init();
  }
  griddep_wait();
  wait_epoch_flags(pf);
  trace_stamp(pf.trace, cur, 1);
  __syncthreads();  // barriers initialised

  uint64_t pol = 0;
  if (POLICY == 1) pol = make_evict_first_policy();

  int rowsv = (t.rows + VEC - 1) / VEC * VEC;  // pad rows of A are zero; pad y with zeros too
  bool bulk = stage_issue<T>(ys0, in + t.vin, t.rows, rowsv, &bar[0]);
  uint32_t par = 0u;  // bit b: phase parity of bar[b]
  int buf = 0;

  if (tid == 0) s_next[0] = (int)gridDim.x + (int)atomicAdd(q.counter, 1u);  // two draws ahead, as in the forward kernel
  int draw = 0;

  for (int it = 0;; ++it) {
    T* ys = buf ? ys1 : ys0;
    if (tid == 0) draw = (int)gridDim.x + (int)atomicAdd(q.counter, 1u);
    stage_wait<T>(bulk, ys, in + t.vin, t.rows, rowsv, &bar[buf], (par >> buf) & 1u);
    if (bulk) par ^= 1u << buf;
    trace_stamp(pf.trace, cur, 2);
    const int nxt = s_next[it & 1];
    const bool more = nxt < q.ntiles;
    bool bulk_n = false;
    if (more) {
      const DenseTile& g = tiles[nxt];
      bulk_n = stage_issue<T>(buf ? ys0 : ys1, in + g.vin, g.rows, (g.rows + VEC - 1) / VEC * VEC, &bar[buf ^ 1]);
    }

    const int cols = t.cols;
    const long long ld = t.ld;
    const T* A = reinterpret_cast<const T*>(t.A);
    T* dst = (t.direct ? out : W) + t.wout;
    const int ngroups = (cols + CW - 1) / CW;

    for (int g = warp; g < ngroups; g += kThreads / 32) {
      const int c0 = g * CW;
      const int last = cols - 1 - c0;  // columns past the tile edge re-read the last one and are not stored
      const T* base = A + (long long)c0 * ld + lane * VEC;
      T acc[CW];
#pragma unroll
      for (int k = 0; k < CW; ++k) acc[k] = T(0);

      int rb = 0;
      for (; rb + RU * S <= rowsv; rb += RU * S) {
        V av[RU][CW];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
          for (int k = 0; k < CW; ++k)
            av[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
          for (int k = 0; k < CW; ++k) dot_vec(acc[k], av[u][k], y);
        }
      }
      if (CW == 4) {
        // tall columns: the remainder is at most RU-1 chunks: pairs of chunks first, then a guarded last one
        if (RU > 2) {
          for (; rb + 2 * S <= rowsv; rb += 2 * S) {
            V av[2][CW];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int k = 0; k < CW; ++k)
                av[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
              for (int k = 0; k < CW; ++k) dot_vec(acc[k], av[u][k], y);
            }
          }
        }
        for (; rb < rowsv; rb += S) {
          if (rb + lane * VEC < rowsv) {
            V av[CW];
#pragma unroll
            for (int k = 0; k < CW; ++k)
              av[k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb), pol);
            const V y = *reinterpret_cast<const V*>(ys + rb + lane * VEC);
#pragma unroll
            for (int k = 0; k < CW; ++k) dot_vec(acc[k], av[k], y);
          }
        }
      } else if (rb < rowsv) {
        // short columns: the remainder IS most of the column -- one guarded batch, CW*RU loads in flight
        V av[RU][CW];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const bool ok = rb + u * S + lane * VEC < rowsv;
#pragma unroll
          for (int k = 0; k < CW; ++k) {
            if (ok)
              av[u][k] = ld_stream<POLICY>(reinterpret_cast<const V*>(base + (long long)min(k, last) * ld + rb + u * S), pol);
          }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          if (rb + u * S + lane * VEC < rowsv) {
            const V y = *reinterpret_cast<const V*>(ys + rb + u * S + lane * VEC);
#pragma unroll
            for (int k = 0; k < CW; ++k) dot_vec(acc[k], av[u][k], y);
          }
        }
      }
      // Warp reduction of CW columns at once: while more than one column is live, each butterfly step
      // also halves the columns a lane keeps (the lanes with the offset bit set keep the upper half), so
      // CW columns cost CW-1 + (5 - log2 CW) shuffles instead of 5*CW.  The pattern is fixed: deterministic.
      int width = CW;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        if (width > 1) {
          width >>= 1;
          const bool upper = (lane & o) != 0;
#pragma unroll
          for (int k = 0; k < CW / 2; ++k) {
            if (k < width) {
              const T send = upper ? acc[k] : acc[k + width];
              const T keep = upper ? acc[k + width] : acc[k];
              acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
        } else {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
        }
      }
      // lane bits 16, 8, ... (one per halving step) spell the column a lane ended up with
      constexpr int kHalvings = (CW == 16) ? 4 : (CW == 8) ? 3 : (CW == 4) ? 2 : (CW == 2) ? 1 : 0;
      int mycol = 0;
#pragma unroll
      for (int h = 0; h < kHalvings; ++h) mycol |= ((lane >> (4 - h)) & 1) << (kHalvings - 1 - h);
      const bool writer = (lane & ((32 >> kHalvings) - 1)) == 0;
      if (writer && c0 + mycol < cols) dst[c0 + mycol] = acc[0];
    }
    if (t.fin >= 0) fused_finish<T>(t.fin, ff, in, W, out);
    trace_stamp(pf.trace, cur, 3);
    if (tid == 0) {
      s_next[(it + 1) & 1] = draw;
      if (draw == 2 * (int)gridDim.x + q.ntiles - 1) *q.counter = 0u;  // the last draw of the launch rearms the queue
    }
    if (!more) break;
    t = tiles[nxt];
    bulk = bulk_n;
    rowsv = (t.rows + VEC - 1) / VEC * VEC;
    cur = nxt;
    buf ^= 1;
    trace_stamp(pf.trace, cur, 0);
    trace_stamp(pf.trace, cur, 1);
  }
}

// Loose items: block rows / columns with no dense tile on this rank (diagonal-only: the fused
// streaming `d .= diag .* m`), and every item of a rank that also sums planes received from peers.
template <typename T, bool CPLX>
__global__ void __launch_bounds__(kThreads, 3) djets_segsum_kernel(const FinishItem* __restrict__ items,
                                                                const DiagTerm* __restrict__ diags,
                                                                const T* __restrict__ in, const T* __restrict__ W,
                                                                const T* __restrict__ R, T* __restrict__ out) {
  finish_item<T, false, CPLX>(items[blockIdx.x], diags, in, W, R, out);
}

// ------------------------------------------------------------------------------------------------
// Reductions (dot, norms, extrema, sums) in one launch: one partial per CTA, folded by the last CTA to
// arrive.  Grid shape depends only on n, the combine order is fixed -> deterministic.
// ------------------------------------------------------------------------------------------------
enum { K_SUM = 0, K_MIN, K_MAX, K_ABSSUM, K_ABSMAX, K_ABSMIN, K_SUMSQ, K_NNZ, K_SUMPOW, K_SUMSQ_SHIFT, K_DOT,
       // complex data (interleaved re, im; n counts reals): |z|-based maps and Im(sum conj(x) y)
       K_CABSSUM, K_CABSMAX, K_CABSMIN, K_CNNZ, K_CSUMPOW, K_CDOT_IM };
template <int KIND>
struct KindInfo {
  static constexpr bool pairs = KIND >= K_CABSSUM;
  static constexpr bool two = KIND == K_DOT || KIND == K_CDOT_IM;
  // the kind whose identity / combine / map applies to |z|
  static constexpr int base = KIND == K_CABSSUM ? K_ABSSUM : KIND == K_CABSMAX ? K_ABSMAX : KIND == K_CABSMIN ? K_ABSMIN
                              : KIND == K_CNNZ ? K_NNZ : KIND == K_CSUMPOW ? K_SUMPOW : KIND == K_CDOT_IM ? K_SUM : KIND;
};

template <int KIND>
__device__ __forceinline__ double red_identity() {
  if (KIND == K_MIN || KIND == K_ABSMIN) return INFINITY;
  if (KIND == K_MAX) return -INFINITY;
  if (KIND == K_ABSMAX) return 0.0;
  return 0.0;
}
// minimum / maximum propagate NaN like Julia's (src:200-203, 236); fmin / fmax would drop it
template <int KIND>
__device__ __forceinline__ double red_combine(double a, double b) {
  if (KIND == K_MIN || KIND == K_ABSMIN) return (a != a || b != b) ? NAN : fmin(a, b);
  if (KIND == K_MAX || KIND == K_ABSMAX) return (a != a || b != b) ? NAN : fmax(a, b);
  return a + b;
}
template <int KIND>
__device__ __forceinline__ double red_map(double x, double y, double param) {
  switch (KIND) {
    case K_SUM:
    case K_MIN:
    case K_MAX: return x;
    case K_ABSSUM:
    case K_ABSMAX:
    case K_ABSMIN: return fabs(x);
    case K_SUMSQ: return x * x;
    case K_NNZ: return x != 0.0 ? 1.0 : 0.0;
    case K_SUMPOW: return pow(fabs(x), param);
    case K_SUMSQ_SHIFT: return (x - param) * (x - param);
    case K_DOT: return x * y;
  }
  return 0.0;
}

// one 128-bit vector (or a scalar tail) worth of elements into the running value
template <int KIND, int N, typename T>
__device__ __forceinline__ double red_accumulate(double acc, const T* ae, const T* be, double param) {
  constexpr int B = KindInfo<KIND>::base;
  if (KindInfo<KIND>::pairs) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
      const double v = KIND == K_CDOT_IM ? (double)ae[k] * (double)be[k + 1] - (double)ae[k + 1] * (double)be[k]
                                         : red_map<B>(hypot((double)ae[k], (double)ae[k + 1]), 0.0, param);
      acc = red_combine<B>(acc, KIND == K_CDOT_IM ? v : v);
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k)
      acc = red_combine<B>(acc, red_map<B>((double)ae[k], KindInfo<KIND>::two ? (double)be[k] : 0.0, param));
  }
  return acc;
}

template <int KIND>
__device__ __forceinline__ double block_fold(double v) {
  __shared__ double sh[kThreads / 32];
  __syncthreads();  // a previous fold of this CTA has been read
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red_identity<KIND>();
  if (threadIdx.x == 0) {
    r = sh[0];
    for (int w = 1; w < kThreads / 32; ++w) r = red_combine<KIND>(r, sh[w]);
  }
  return r;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads) djets_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                          long long n, double param, double* partials,
                                                          unsigned* ticket, double* __restrict__ result) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int U = 4;
  constexpr int B = KindInfo<KIND>::base;
  constexpr bool TWO = KindInfo<KIND>::two;
  const long long nvec = n / VEC;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long gstride = (long long)gridDim.x * kThreads;
  double acc = red_identity<B>();
  const V* xv = reinterpret_cast<const V*>(x);
  const V* yv = reinterpret_cast<const V*>(y);
  long long i = gtid;
  for (; i + (U - 1) * gstride < nvec; i += U * gstride) {
    V a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a[u] = ld_stream<0>(xv + i + u * gstride, 0);
      if (TWO) b[u] = ld_stream<0>(yv + i + u * gstride, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      acc = red_accumulate<KIND, VEC>(acc, reinterpret_cast<const T*>(&a[u]), reinterpret_cast<const T*>(&b[u]), param);
  }
  for (; i < nvec; i += gstride) {
    V a = ld_stream<0>(xv + i, 0), b;
    if (TWO) b = ld_stream<0>(yv + i, 0);
    acc = red_accumulate<KIND, VEC>(acc, reinterpret_cast<const T*>(&a), reinterpret_cast<const T*>(&b), param);
  }
  constexpr int STEP = KindInfo<KIND>::pairs ? 2 : 1;  // the tail keeps (re, im) pairs together
  for (long long j = nvec * VEC + STEP * gtid; j + STEP <= n; j += STEP * gstride)
    acc = red_accumulate<KIND, STEP>(acc, x + j, TWO ? y + j : x + j, param);
  const double r = block_fold<B>(acc);
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) result[0] = r;
    return;
  }
  // One launch: every CTA publishes its partial and takes a ticket; the CTA holding the last ticket folds the
  // partials in index order (the result does not depend on which CTA came last) and rearms the ticket.
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = r;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) + 1u == gridDim.x) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double fold = red_identity<B>();
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) fold = red_combine<B>(fold, __ldcg(partials + i));
  const double total = block_fold<B>(fold);
  if (threadIdx.x == 0) {
    result[0] = total;
    *ticket = 0u;
  }
}

int reduce_max_ctas() { return 148 * 8; }

template <typename T, int KIND>
static cudaError_t launch_reduce_k(const void* x, const void* y, long long n, double param, double* partials,
                                   unsigned* ticket, double* result, cudaStream_t stream) {
  constexpr int VEC = VecT<T>::N;
  long long want = (n / VEC + (long long)kThreads * 4 - 1) / ((long long)kThreads * 4);
  int grid = (int)(want < 1 ? 1 : (want > reduce_max_ctas() ? reduce_max_ctas() : want));
  djets_reduce_kernel<T, KIND><<<grid, kThreads, 0, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<const T*>(y),
                                                        n, param, partials, ticket, result);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_reduce_t(int kind, const void* x, const void* y, long long n, double param,
                                   double* partials, unsigned* ticket, double* result, cudaStream_t stream) {
  switch (kind) {
    case 0: return launch_reduce_k<T, K_SUM>(x, y, n, param, partials, ticket, result, stream);
    case 1: return launch_reduce_k<T, K_MIN>(x, y, n, param, partials, ticket, result, stream);
    case 2: return launch_reduce_k<T, K_MAX>(x, y, n, param, partials, ticket, result, stream);
    case 3: return launch_reduce_k<T, K_ABSSUM>(x, y, n, param, partials, ticket, result, stream);
    case 4: return launch_reduce_k<T, K_ABSMAX>(x, y, n, param, partials, ticket, result, stream);
    case 5: return launch_reduce_k<T, K_ABSMIN>(x, y, n, param, partials, ticket, result, stream);
    case 6: return launch_reduce_k<T, K_SUMSQ>(x, y, n, param, partials, ticket, result, stream);
    case 7: return launch_reduce_k<T, K_NNZ>(x, y, n, param, partials, ticket, result, stream);
    case 8: return launch_reduce_k<T, K_SUMPOW>(x, y, n, param, partials, ticket, result, stream);
    case 9: return launch_reduce_k<T, K_SUMSQ_SHIFT>(x, y, n, param, partials, ticket, result, stream);
    case RED_DOT: return launch_reduce_k<T, K_DOT>(x, y, n, param, partials, ticket, result, stream);
    case RED_CABSSUM: return launch_reduce_k<T, K_CABSSUM>(x, y, n, param, partials, ticket, result, stream);
    case RED_CABSMAX: return launch_reduce_k<T, K_CABSMAX>(x, y, n, param, partials, ticket, result, stream);
    case RED_CABSMIN: return launch_reduce_k<T, K_CABSMIN>(x, y, n, param, partials, ticket, result, stream);
    case RED_CNNZ: return launch_reduce_k<T, K_CNNZ>(x, y, n, param, partials, ticket, result, stream);
    case RED_CSUMPOW: return launch_reduce_k<T, K_CSUMPOW>(x, y, n, param, partials, ticket, result, stream);
    case RED_CDOT_IM: return launch_reduce_k<T, K_CDOT_IM>(x, y, n, param, partials, ticket, result, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_reduce(int dtype, int kind, const void* x, const void* y, long long n, double param,
                          double* partials, unsigned* ticket, double* result, cudaStream_t stream) {
  return dtype == 0 ? launch_reduce_t<float>(kind, x, y, n, param, partials, ticket, result, stream)
                    : launch_reduce_t<double>(kind, x, y, n, param, partials, ticket, result, stream);
}

// ------------------------------------------------------------------------------------------------
// Streaming elementwise kernels.  Products and sums use the non-contracting intrinsics so that
// `dest .= c0*x0 .+ c1*x1 ...` rounds exactly like the reference's unfused broadcast (src:363-368).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

struct LincombDev {
  const void* x[4];
  double coef[4];
  int xdtype[4];
  int nterms;
  int unit0;  // coef[0] == 1 and nterms == 1: plain copy / convert
};

template <typename ACC>
__device__ __forceinline__ ACC load_as(const void* p, int dtype, long long i) {
  return dtype == 0 ? (ACC) reinterpret_cast<const float*>(p)[i] : (ACC) reinterpret_cast<const double*>(p)[i];
}

// generic path: any mix of operand element types
template <typename TD, typename ACC>
__global__ void __launch_bounds__(kThreads) djets_lincomb_generic_kernel(TD* dest /* may alias an operand */, LincombDev a, long long n) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    ACC s;
    if (a.unit0) {
      s = load_as<ACC>(a.x[0], a.xdtype[0], i);
    } else {
      s = mul_rn((ACC)a.coef[0], load_as<ACC>(a.x[0], a.xdtype[0], i));
      for (int k = 1; k < a.nterms; ++k)
        s = add_rn(s, mul_rn((ACC)a.coef[k], load_as<ACC>(a.x[k], a.xdtype[k], i)));
    }
    dest[i] = (TD)s;
  }
}

// fast path: all operands and dest share one element type; 128-bit loads / stores.  A thread takes UV vectors (one
// grid stride apart) at a time: NT*UV = 6-8 loads issued before the first store (2^30 Float64: scale 6.14 -> 6.29,
// axpy 6.39 -> 6.49, three terms 6.68 -> 6.97 TB/s).
template <typename T, typename ACC, int NT>
__global__ void __launch_bounds__(kThreads, 2) djets_lincomb_kernel(T* dest /* may alias an operand */, LincombDev a, long long n) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  constexpr int UV = (NT == 1) ? 8 : (NT == 2) ? 4 : 2;
  const long long nvec = n / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
  ACC c[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) c[k] = (ACC)a.coef[k];
  auto combine = [&](const V (&v)[NT]) {
    V o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      ACC s = mul_rn(c[0], (ACC) reinterpret_cast<const T*>(&v[0])[e]);
#pragma unroll
      for (int k = 1; k < NT; ++k) s = add_rn(s, mul_rn(c[k], (ACC) reinterpret_cast<const T*>(&v[k])[e]));
      oe[e] = (T)s;
    }
    return o;
  };
  long long i = gtid;
  for (; i + (UV - 1) * stride < nvec; i += UV * stride) {
    V v[UV][NT];
#pragma unroll
    for (int u = 0; u < UV; ++u)
#pragma unroll
      for (int k = 0; k < NT; ++k) v[u][k] = reinterpret_cast<const V*>(a.x[k])[i + u * stride];
#pragma unroll
    for (int u = 0; u < UV; ++u) reinterpret_cast<V*>(dest)[i + u * stride] = combine(v[u]);
  }
  for (; i < nvec; i += stride) {
    V v[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) v[k] = reinterpret_cast<const V*>(a.x[k])[i];
    reinterpret_cast<V*>(dest)[i] = combine(v);
  }
  for (long long j = nvec * VEC + gtid; j < n; j += stride) {
    ACC s = mul_rn(c[0], (ACC) reinterpret_cast<const T*>(a.x[0])[j]);
#pragma unroll
    for (int k = 1; k < NT; ++k) s = add_rn(s, mul_rn(c[k], (ACC) reinterpret_cast<const T*>(a.x[k])[j]));
    dest[j] = (T)s;
  }
}

// Grid cap of the specialised streaming kernels in CTAs per SM (DJETS_STREAM_CTAS overrides).  More CTAs than are
// resident, each with a few batches, measured best: 16 -> 6.29 / 6.49 / 6.97 TB/s (1 / 2 / 3 terms), 2 -> 6.10 / 6.44 /
// 6.66, 4 -> 5.96 / 6.20 / 6.44.
static int stream_ctas_per_sm() {
  static const int v = getenv("DJETS_STREAM_CTAS") ? atoi(getenv("DJETS_STREAM_CTAS")) : 16;
  return v < 1 ? 1 : v;
}

static int stream_grid(long long nvec) {
  long long want = (nvec + kThreads - 1) / kThreads;
  const long long cap = 148LL * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <typename T, typename ACC>
static void launch_lincomb_vec(T* dest, const LincombDev& d, long long n, cudaStream_t stream) {
  const int grid = std::min(stream_grid(n / VecT<T>::N), 148 * stream_ctas_per_sm());
  switch (d.nterms) {
    case 1: djets_lincomb_kernel<T, ACC, 1><<<grid, kThreads, 0, stream>>>(dest, d, n); break;
    case 2: djets_lincomb_kernel<T, ACC, 2><<<grid, kThreads, 0, stream>>>(dest, d, n); break;
    case 3: djets_lincomb_kernel<T, ACC, 3><<<grid, kThreads, 0, stream>>>(dest, d, n); break;
    default: djets_lincomb_kernel<T, ACC, 4><<<grid, kThreads, 0, stream>>>(dest, d, n); break;
  }
}

cudaError_t launch_lincomb(int ddtype, void* dest, const LincombArgs& a, long long n, int wide, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  LincombDev d;
  bool same = true;
  for (int k = 0; k < 4; ++k) {
    d.x[k] = k < a.nterms ? a.x[k] : nullptr;
    d.coef[k] = k < a.nterms ? a.coef[k] : 0.0;
    d.xdtype[k] = k < a.nterms ? a.xdtype[k] : ddtype;
    if (k < a.nterms && a.xdtype[k] != ddtype) same = false;
  }
  d.nterms = a.nterms;
  d.unit0 = (a.nterms == 1 && a.coef[0] == 1.0) ? 1 : 0;
  if (same && d.unit0) {  // `dest .= x` in one element type: a device copy
    if (dest == a.x[0]) return cudaSuccess;
    return cudaMemcpyAsync(dest, a.x[0], (size_t)n * elem_size(ddtype), cudaMemcpyDeviceToDevice, stream);
  }
  if (same && !d.unit0) {
    if (ddtype == 1)
      launch_lincomb_vec<double, double>(reinterpret_cast<double*>(dest), d, n, stream);
    else if (wide)
      launch_lincomb_vec<float, double>(reinterpret_cast<float*>(dest), d, n, stream);
    else
      launch_lincomb_vec<float, float>(reinterpret_cast<float*>(dest), d, n, stream);
  } else {
    const int grid = stream_grid(n);
    if (ddtype == 1)
      djets_lincomb_generic_kernel<double, double><<<grid, kThreads, 0, stream>>>(reinterpret_cast<double*>(dest), d, n);
    else if (wide || !same)
      djets_lincomb_generic_kernel<float, double><<<grid, kThreads, 0, stream>>>(reinterpret_cast<float*>(dest), d, n);
    else
      djets_lincomb_generic_kernel<float, float><<<grid, kThreads, 0, stream>>>(reinterpret_cast<float*>(dest), d, n);
  }
  return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(kThreads) djets_fill_kernel(T* __restrict__ dest, T a, long long n) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  const long long nvec = n / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
  V v;
  T* ve = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int e = 0; e < VEC; ++e) ve[e] = a;
  for (long long i = gtid; i < nvec; i += stride) reinterpret_cast<V*>(dest)[i] = v;
  for (long long j = nvec * VEC + gtid; j < n; j += stride) dest[j] = a;
}

cudaError_t launch_fill(int dtype, void* dest, double a, long long n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (dtype == 0)
    djets_fill_kernel<float><<<stream_grid(n / 4), kThreads, 0, stream>>>(reinterpret_cast<float*>(dest), (float)a, n);
  else
    djets_fill_kernel<double><<<stream_grid(n / 2), kThreads, 0, stream>>>(reinterpret_cast<double*>(dest), a, n);
  return cudaGetLastError();
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads) djets_binary_kernel(T* dest /* may alias a or b */, const T* a, const T* b, long long n) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  const long long nvec = n / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
  for (long long i = gtid; i < nvec; i += stride) {
    V va = reinterpret_cast<const V*>(a)[i], vb = reinterpret_cast<const V*>(b)[i], o;
    const T* ae = reinterpret_cast<const T*>(&va);
    const T* be = reinterpret_cast<const T*>(&vb);
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) oe[e] = OP == 0 ? mul_rn(ae[e], be[e]) : ae[e] / be[e];
    reinterpret_cast<V*>(dest)[i] = o;
  }
  for (long long j = nvec * VEC + gtid; j < n; j += stride) dest[j] = OP == 0 ? mul_rn(a[j], b[j]) : a[j] / b[j];
}

template <typename T>
__global__ void __launch_bounds__(kThreads) djets_fill_complex_kernel(T* __restrict__ dest, T re, T im, long long n /* reals */) {
  using V = typename VecT<T>::type;
  constexpr int VEC = VecT<T>::N;
  const long long nvec = n / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
  V v;
  T* ve = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int e = 0; e < VEC; ++e) ve[e] = (e & 1) ? im : re;
  for (long long i = gtid; i < nvec; i += stride) reinterpret_cast<V*>(dest)[i] = v;
  for (long long j = nvec * VEC + gtid; j < n; j += stride) dest[j] = (j & 1) ? im : re;
}

cudaError_t launch_fill_complex(int dtype, void* dest, double re, double im, long long n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (real_dtype(dtype) == 0)
    djets_fill_complex_kernel<float><<<stream_grid(n / 2), kThreads, 0, stream>>>(reinterpret_cast<float*>(dest), (float)re, (float)im, 2 * n);
  else
    djets_fill_complex_kernel<double><<<stream_grid(n), kThreads, 0, stream>>>(reinterpret_cast<double*>(dest), re, im, 2 * n);
  return cudaGetLastError();
}

cudaError_t launch_binary(int dtype, int op, void* dest, const void* a, const void* b, long long n,
                          cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (op != 0 && op != 1) return cudaErrorInvalidValue;
#define DJ_BIN(T, OP)                                                                                              \
  djets_binary_kernel<T, OP><<<stream_grid(n / VecT<T>::N), kThreads, 0, stream>>>(                                      \
      reinterpret_cast<T*>(dest), reinterpret_cast<const T*>(a), reinterpret_cast<const T*>(b), n)
  if (dtype == 0) {
    if (op == 0) DJ_BIN(float, 0); else DJ_BIN(float, 1);
  } else {
    if (op == 0) DJ_BIN(double, 0); else DJ_BIN(double, 1);
  }
#undef DJ_BIN
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launch dispatch for the tile kernels
// ------------------------------------------------------------------------------------------------
constexpr size_t kAdjSmemBytes = 2 * kAdjMaxTRBytes + 2 * sizeof(uint64_t) + 2 * sizeof(int);

// Kernels that need more than 48 KB of dynamic shared memory opt in once per (kernel, device).
static cudaError_t allow_smem(const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (done.count({kernel, dev})) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.insert({kernel, dev});
  return e;
}

template <typename... Args>
static cudaError_t launch_tiles(void (*kernel)(Args...), int grid, size_t smem, int pdl, cudaStream_t stream,
                                Args... args) {
  cudaError_t e = allow_smem(reinterpret_cast<const void*>(kernel), smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_n_t(int fwd_tx, const DenseTile* tiles, int ntiles, int grid, unsigned* counter,
                                   const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                   cudaStream_t stream) {
  const TileQueue q{ntiles, counter};
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  switch (fwd_tx) {
    case 32:
      if (grid >= ntiles) return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 32, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
      return launch_tiles(djets_blockmul_fwd_queue_kernel<T, 32, POLICY>, grid, (size_t)0, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
    case 64:
      if (grid >= ntiles) return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 64, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
      return launch_tiles(djets_blockmul_fwd_queue_kernel<T, 64, POLICY>, grid, (size_t)0, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
    case 128:
      if (grid >= ntiles) return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 128, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
      return launch_tiles(djets_blockmul_fwd_queue_kernel<T, 128, POLICY>, grid, (size_t)0, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
    case 256:
      if (grid >= ntiles) return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 256, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
      return launch_tiles(djets_blockmul_fwd_queue_kernel<T, 256, POLICY>, grid, (size_t)0, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_nc_t(int fwd_tx, const DenseTile* tiles, int ntiles, const void* in, void* W, void* out,
                                    FusedFinish ff, Prefetch pf, cudaStream_t stream) {
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  switch (fwd_tx) {
    case 64: return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 64, POLICY, true>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    case 128: return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 128, POLICY, true>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    case 256: return launch_tiles(djets_blockmul_fwd_tile_kernel<T, 256, POLICY, true>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_gemv_n(int dtype, int fwd_tx, int ld_policy, const DenseTile* tiles, int ntiles, int grid,
                          unsigned* counter, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                          cudaStream_t stream) {
  if (ntiles <= 0) return cudaSuccess;
  if (is_complex(dtype)) {  // complex: one tile per CTA (the queue-driven variants are real only)
    if (real_dtype(dtype) == 0)
      return ld_policy ? launch_gemv_nc_t<float, 1>(fwd_tx, tiles, ntiles, in, W, out, ff, pf, stream)
                       : launch_gemv_nc_t<float, 0>(fwd_tx, tiles, ntiles, in, W, out, ff, pf, stream);
    return ld_policy ? launch_gemv_nc_t<double, 1>(fwd_tx, tiles, ntiles, in, W, out, ff, pf, stream)
                     : launch_gemv_nc_t<double, 0>(fwd_tx, tiles, ntiles, in, W, out, ff, pf, stream);
  }
  if (dtype == 0)
    return ld_policy ? launch_gemv_n_t<float, 1>(fwd_tx, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream)
                     : launch_gemv_n_t<float, 0>(fwd_tx, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream);
  return ld_policy ? launch_gemv_n_t<double, 1>(fwd_tx, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream)
                   : launch_gemv_n_t<double, 0>(fwd_tx, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream);
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_t_t(int adj_ru, int adj_cw, const DenseTile* tiles, int ntiles, int grid,
                                   unsigned* counter, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                   cudaStream_t stream) {
  const TileQueue q{ntiles, counter};
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  // (rows unrolled, columns per warp): tall columns stream down 4 columns, short ones fan out over 16
  if (adj_cw == 8 && adj_ru == 1) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 1, 8, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 1, 8, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  if (adj_cw == 16 && adj_ru == 1) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 1, 16, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 1, 16, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  if (adj_cw == 8 && adj_ru == 2) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 2, 8, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 2, 8, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  if (adj_cw == 4 && adj_ru == 4) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 4, 4, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 4, 4, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  if (adj_cw == 4 && adj_ru == 2) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 2, 4, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 2, 4, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  if (adj_cw == 4 && adj_ru == 1) {
    if (grid >= ntiles) return launch_tiles(djets_blockmul_adj_tile_kernel<T, 1, 4, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
    return launch_tiles(djets_blockmul_adj_queue_kernel<T, 1, 4, POLICY>, grid, kAdjSmemBytes, pf.pdl, stream, tiles, q, i, w, o, ff, pf);
  }
  return cudaErrorInvalidValue;
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_tc_t(int adj_ru, const DenseTile* tiles, int ntiles, const void* in, void* W, void* out,
                                    FusedFinish ff, Prefetch pf, cudaStream_t stream) {
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  if (adj_ru == 1) return launch_tiles(djets_blockmul_adjc_tile_kernel<T, 1, 8, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
  return launch_tiles(djets_blockmul_adjc_tile_kernel<T, 4, 4, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, i, w, o, ff, pf);
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_t_chain_t(int pair, const DenseTile* tiles, const ChainLink* links, int ntiles,
                                         const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                         cudaStream_t stream) {
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  if (pair) return launch_tiles(djets_blockmul_adj_chain_kernel<T, 8, true, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, links, i, w, o, ff, pf);
  return launch_tiles(djets_blockmul_adj_chain_kernel<T, 8, false, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, links, i, w, o, ff, pf);
}

cudaError_t launch_gemv_t_chain(int dtype, int pair, int ld_policy, const DenseTile* tiles, const ChainLink* links,
                                int ntiles, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                cudaStream_t stream) {
  if (ntiles <= 0) return cudaSuccess;
  if (is_complex(dtype)) return cudaErrorInvalidValue;
  if (dtype == 0)
    return ld_policy ? launch_gemv_t_chain_t<float, 1>(pair, tiles, links, ntiles, in, W, out, ff, pf, stream)
                     : launch_gemv_t_chain_t<float, 0>(pair, tiles, links, ntiles, in, W, out, ff, pf, stream);
  return ld_policy ? launch_gemv_t_chain_t<double, 1>(pair, tiles, links, ntiles, in, W, out, ff, pf, stream)
                   : launch_gemv_t_chain_t<double, 0>(pair, tiles, links, ntiles, in, W, out, ff, pf, stream);
}

template <typename T, int POLICY>
static cudaError_t launch_gemv_n_chain_t(int fwd_tx, const DenseTile* tiles, const ChainLink* links, int ntiles,
                                         const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                         cudaStream_t stream) {
  const T* i = reinterpret_cast<const T*>(in);
  T* w = reinterpret_cast<T*>(W);
  T* o = reinterpret_cast<T*>(out);
  if (fwd_tx == 32) return launch_tiles(djets_blockmul_fwd_chain_kernel<T, 32, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, links, i, w, o, ff, pf);
  if (fwd_tx == 64) return launch_tiles(djets_blockmul_fwd_chain_kernel<T, 64, POLICY>, ntiles, (size_t)0, pf.pdl, stream, tiles, links, i, w, o, ff, pf);
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemv_n_chain(int dtype, int fwd_tx, int ld_policy, const DenseTile* tiles, const ChainLink* links,
                                int ntiles, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                                cudaStream_t stream) {
  if (ntiles <= 0) return cudaSuccess;
  if (is_complex(dtype)) return cudaErrorInvalidValue;
  if (dtype == 0)
    return ld_policy ? launch_gemv_n_chain_t<float, 1>(fwd_tx, tiles, links, ntiles, in, W, out, ff, pf, stream)
                     : launch_gemv_n_chain_t<float, 0>(fwd_tx, tiles, links, ntiles, in, W, out, ff, pf, stream);
  return ld_policy ? launch_gemv_n_chain_t<double, 1>(fwd_tx, tiles, links, ntiles, in, W, out, ff, pf, stream)
                   : launch_gemv_n_chain_t<double, 0>(fwd_tx, tiles, links, ntiles, in, W, out, ff, pf, stream);
}

cudaError_t launch_gemv_t(int dtype, int adj_ru, int adj_cw, int ld_policy, const DenseTile* tiles, int ntiles,
                          int grid, unsigned* counter, const void* in, void* W, void* out, FusedFinish ff, Prefetch pf,
                          cudaStream_t stream) {
  if (ntiles <= 0) return cudaSuccess;
  if (is_complex(dtype)) {
    if (real_dtype(dtype) == 0)
      return ld_policy ? launch_gemv_tc_t<float, 1>(adj_ru, tiles, ntiles, in, W, out, ff, pf, stream)
                       : launch_gemv_tc_t<float, 0>(adj_ru, tiles, ntiles, in, W, out, ff, pf, stream);
    return ld_policy ? launch_gemv_tc_t<double, 1>(adj_ru, tiles, ntiles, in, W, out, ff, pf, stream)
                     : launch_gemv_tc_t<double, 0>(adj_ru, tiles, ntiles, in, W, out, ff, pf, stream);
  }
  if (dtype == 0)
    return ld_policy ? launch_gemv_t_t<float, 1>(adj_ru, adj_cw, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream)
                     : launch_gemv_t_t<float, 0>(adj_ru, adj_cw, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream);
  return ld_policy ? launch_gemv_t_t<double, 1>(adj_ru, adj_cw, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream)
                   : launch_gemv_t_t<double, 0>(adj_ru, adj_cw, tiles, ntiles, grid, counter, in, W, out, ff, pf, stream);
}

// ------------------------------------------------------------------------------------------------
// epoch flags of the cross-process peer-memory exchange
// ------------------------------------------------------------------------------------------------
__global__ void djets_flags_wait_kernel(FlagList f, unsigned* err) {
  const int k = threadIdx.x;
  if (k >= f.n) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned cur;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(f.p[k]) : "memory");
    if ((int)(cur - f.v[k]) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 20000000000ull) {  // 20 s: a peer died or the call sequences of the processes diverged
      atomicExch(err, 1u);
      break;
    }
    __nanosleep(64);
  }
}

__global__ void djets_flags_signal_kernel(FlagList f) {
  const int k = threadIdx.x;
  if (k >= f.n) return;
  __threadfence_system();  // everything this stream stored before (peer stores included) is visible first
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[k]), "r"(f.v[k]) : "memory");
}

__global__ void __launch_bounds__(kThreads) djets_xp_push_kernel(const char* __restrict__ src, size_t bytes, PushList dst,
                                                           FlagList waits, FlagList signals, unsigned* ticket,
                                                           unsigned* err) {
  const int tid = threadIdx.x;
  if (tid < waits.n) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned cur;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(waits.p[tid]) : "memory");
      if ((int)(cur - waits.v[tid]) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 20000000000ull) {
        atomicExch(err, 1u);
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const size_t nvec = bytes / 16;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + tid; i < nvec; i += stride) {
    const int4 v = reinterpret_cast<const int4*>(src)[i];
    for (int k = 0; k < dst.n; ++k) reinterpret_cast<int4*>(dst.dst[k])[i] = v;
  }
  if (blockIdx.x == 0)  // tail of a part whose size is not a multiple of 16 bytes (element sizes are 4 or 8)
    for (size_t b = nvec * 16 + (size_t)tid * 4; b < bytes; b += (size_t)kThreads * 4) {
      const int v = *reinterpret_cast<const int*>(src + b);
      for (int k = 0; k < dst.n; ++k) *reinterpret_cast<int*>(dst.dst[k] + b) = v;
    }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    if (t + 1u == gridDim.x) {
      *ticket = 0u;
      __threadfence_system();
      for (int k = 0; k < signals.n; ++k)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(signals.p[k]), "r"(signals.v[k]) : "memory");
    }
  }
}

cudaError_t launch_xp_push(const char* src, size_t bytes, const PushList& dst, const FlagList& waits, const FlagList& signals,
                           unsigned* ticket, unsigned* err, cudaStream_t stream) {
  const size_t nvec = bytes / 16;
  int grid = (int)((nvec + kThreads - 1) / kThreads);
  grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
  djets_xp_push_kernel<<<grid, kThreads, 0, stream>>>(src, bytes, dst, waits, signals, ticket, err);
  return cudaGetLastError();
}

cudaError_t launch_flags_wait(const FlagList& f, unsigned* err, cudaStream_t stream) {
  if (f.n <= 0) return cudaSuccess;
  djets_flags_wait_kernel<<<1, 32, 0, stream>>>(f, err);
  return cudaGetLastError();
}

cudaError_t launch_flags_signal(const FlagList& f, cudaStream_t stream) {
  if (f.n <= 0) return cudaSuccess;
  djets_flags_signal_kernel<<<1, 32, 0, stream>>>(f);
  return cudaGetLastError();
}

template <typename K>
static int resident_ctas(K kernel, size_t smem = 0) {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (allow_smem(reinterpret_cast<const void*>(kernel), smem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess) return 0;
  return sms * per_sm;
}

template <typename T, int POLICY>
static int resident_t(int adjoint, int shape, int adj_cw) {
  if (!adjoint) {
    switch (shape) {
      case 32: return resident_ctas(djets_blockmul_fwd_queue_kernel<T, 32, POLICY>);
      case 64: return resident_ctas(djets_blockmul_fwd_queue_kernel<T, 64, POLICY>);
      case 128: return resident_ctas(djets_blockmul_fwd_queue_kernel<T, 128, POLICY>);
      case 256: return resident_ctas(djets_blockmul_fwd_queue_kernel<T, 256, POLICY>);
    }
    return 0;
  }
  if (adj_cw == 8 && shape == 1) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 1, 8, POLICY>, kAdjSmemBytes);
  if (adj_cw == 16 && shape == 1) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 1, 16, POLICY>, kAdjSmemBytes);
  if (adj_cw == 8 && shape == 2) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 2, 8, POLICY>, kAdjSmemBytes);
  if (adj_cw == 4 && shape == 4) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 4, 4, POLICY>, kAdjSmemBytes);
  if (adj_cw == 4 && shape == 2) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 2, 4, POLICY>, kAdjSmemBytes);
  if (adj_cw == 4 && shape == 1) return resident_ctas(djets_blockmul_adj_queue_kernel<T, 1, 4, POLICY>, kAdjSmemBytes);
  return 0;
}

int gemv_chain_resident_ctas(int dtype, int adjoint, int ld_policy) {
  if (is_complex(dtype)) return 0;
  if (!adjoint) {
    if (dtype == 0)
      return ld_policy ? resident_ctas(djets_blockmul_fwd_chain_kernel<float, 64, 1>) : resident_ctas(djets_blockmul_fwd_chain_kernel<float, 64, 0>);
    return ld_policy ? resident_ctas(djets_blockmul_fwd_chain_kernel<double, 64, 1>) : resident_ctas(djets_blockmul_fwd_chain_kernel<double, 64, 0>);
  }
  if (dtype == 0)
    return ld_policy ? resident_ctas(djets_blockmul_adj_chain_kernel<float, 8, false, 1>) : resident_ctas(djets_blockmul_adj_chain_kernel<float, 8, false, 0>);
  return ld_policy ? resident_ctas(djets_blockmul_adj_chain_kernel<double, 8, false, 1>) : resident_ctas(djets_blockmul_adj_chain_kernel<double, 8, false, 0>);
}

int gemv_resident_ctas(int dtype, int adjoint, int shape, int adj_cw, int ld_policy) {
  if (is_complex(dtype)) {
    dtype = real_dtype(dtype);
    if (adjoint) {  // the complex adjoint has its own two shapes; the estimate only sizes the prefetch wave
      shape = 4;
      adj_cw = 4;
    }
  }
  if (dtype == 0) return ld_policy ? resident_t<float, 1>(adjoint, shape, adj_cw) : resident_t<float, 0>(adjoint, shape, adj_cw);
  return ld_policy ? resident_t<double, 1>(adjoint, shape, adj_cw) : resident_t<double, 0>(adjoint, shape, adj_cw);
}

cudaError_t launch_finish(int dtype, const FinishItem* items, int nitems, const DiagTerm* diags, const void* in,
                          const void* W, const void* R, void* out, cudaStream_t stream) {
  if (nitems <= 0) return cudaSuccess;
#define DJ_SEGSUM(T, C)                                                                                      \
  djets_segsum_kernel<T, C><<<nitems, kThreads, 0, stream>>>(items, diags, reinterpret_cast<const T*>(in), \
                                                             reinterpret_cast<const T*>(W),                \
                                                             reinterpret_cast<const T*>(R), reinterpret_cast<T*>(out))
  if (real_dtype(dtype) == 0) {
    if (is_complex(dtype)) DJ_SEGSUM(float, true); else DJ_SEGSUM(float, false);
  } else {
    if (is_complex(dtype)) DJ_SEGSUM(double, true); else DJ_SEGSUM(double, false);
  }
#undef DJ_SEGSUM
  return cudaGetLastError();
}

}  // namespace djets
