// The closed elementwise program of libdjets_b200 (djets_vec_eval): the reference evaluates any
// `Broadcasted` tree per worker over the local blocks (src/DistributedJets.jl:363-375); a CUDA library
// cannot run a Julia closure, so the host shims lower the tree to a postfix byte-code over a closed set
// of operators and ONE streaming kernel interprets it per 128-bit vector: every input is read once,
// the destination written once, whatever the expression.
//
// The interpreter keeps the expression stack in registers: `top` plus D-1 spill levels that are shifted
// by push / binary-pop with compile-time indices (no local memory); unary operators and the
// scalar-operand forms (x .+ a, a .* x, max.(x, a), x .^ p ...) touch `top` only.  All inputs of an
// iteration are loaded before the program runs (several 128-bit loads in flight per thread); products
// and sums use the non-contracting intrinsics so results round exactly like the unfused broadcast of
// the reference (bit-identical to NumPy / Julia for + - * / abs sqrt min max neg x^2 x^3).
#include "djets_kernels.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace djets {

namespace {

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int N = 2;
};

__device__ __forceinline__ float ev_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double ev_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float ev_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double ev_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float ev_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double ev_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float ev_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double ev_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float ev_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double ev_sqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float ev_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double ev_abs(double a) { return fabs(a); }
__device__ __forceinline__ float ev_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double ev_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float ev_exp(float a) { return expf(a); }
__device__ __forceinline__ double ev_exp(double a) { return exp(a); }
__device__ __forceinline__ float ev_log(float a) { return logf(a); }
__device__ __forceinline__ double ev_log(double a) { return log(a); }
// Julia's min / max propagate NaN (fmin / fmax drop it)
template <typename A>
__device__ __forceinline__ A ev_min(A a, A b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : (b < a ? b : (signbit(a) ? a : b))));
}
template <typename A>
__device__ __forceinline__ A ev_max(A a, A b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : (b > a ? b : (signbit(a) ? b : a))));
}

struct EvalDev {
  unsigned char op[kEvalMaxOps];
  unsigned char arg[kEvalMaxOps];
  const void* in[kEvalMaxIn];
  double sc[kEvalMaxScalars];
  const double* dsc[kEvalMaxScalars];
  int in_dtype[kEvalMaxIn];
  int nops, nin, nsc;
};

// The interpreter proper over E elements held in registers.  XIN(k, e) yields input k's element e.
template <typename ACC, int E, int D, typename XIN>
__device__ __forceinline__ void run_program(const EvalDev& a, const ACC* __restrict__ sc, XIN xin, ACC (&top)[E]) {
  ACC st[D > 1 ? D - 1 : 1][E];
#pragma unroll
  for (int k = 0; k < D - 1; ++k)
#pragma unroll
    for (int e = 0; e < E; ++e) st[k][e] = ACC(0);
#pragma unroll
  for (int e = 0; e < E; ++e) top[e] = ACC(0);

#define EV_PUSH()                                         \
  do {                                                    \
    _Pragma("unroll") for (int k = D - 2; k > 0; --k)    \
        _Pragma("unroll") for (int e = 0; e < E; ++e) st[k][e] = st[k - 1][e]; \
    _Pragma("unroll") for (int e = 0; e < E; ++e) st[0][e] = top[e];           \
  } while (0)
#define EV_BIN(expr)                                      \
  do {                                                    \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {       \
      const ACC x = st[0][e], y = top[e];                 \
      top[e] = (expr);                                    \
    }                                                     \
    _Pragma("unroll") for (int k = 0; k < D - 2; ++k)     \
        _Pragma("unroll") for (int e = 0; e < E; ++e) st[k][e] = st[k + 1][e]; \
  } while (0)
#define EV_UN(expr)                                       \
  do {                                                    \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {       \
      const ACC x = top[e];                               \
      top[e] = (expr);                                    \
    }                                                     \
  } while (0)
#define EV_SCL(expr)                                      \
  do {                                                    \
    const ACC s = sc[arg];                                \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {       \
      const ACC x = top[e];                               \
      top[e] = (expr);                                    \
    }                                                     \
  } while (0)

#define EV_OPD(expr)                                      \
  do {                                                    \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {       \
      const ACC x = top[e], v = xin(arg & 7, e);          \
      top[e] = (expr);                                    \
    }                                                     \
  } while (0)
#define EV_OPDS(expr)                                     \
  do {                                                    \
    const ACC s = sc[arg >> 3];                           \
    _Pragma("unroll") for (int e = 0; e < E; ++e) {       \
      const ACC x = top[e], v = xin(arg & 7, e);          \
      top[e] = (expr);                                    \
    }                                                     \
  } while (0)

  for (int pc = 0; pc < a.nops; ++pc) {
    const int op = a.op[pc], arg = a.arg[pc];
    switch (op) {
      case EV_IN:
        EV_PUSH();
#pragma unroll
        for (int e = 0; e < E; ++e) top[e] = xin(arg, e);
        break;
      case EV_SC: {
        EV_PUSH();
        const ACC s = sc[arg];
#pragma unroll
        for (int e = 0; e < E; ++e) top[e] = s;
      } break;
      case EV_ADD: EV_BIN(ev_add(x, y)); break;
      case EV_SUB: EV_BIN(ev_sub(x, y)); break;
      case EV_MUL: EV_BIN(ev_mul(x, y)); break;
      case EV_DIV: EV_BIN(ev_div(x, y)); break;
      case EV_POW: EV_BIN(ev_pow(x, y)); break;
      case EV_MIN: EV_BIN(ev_min(x, y)); break;
      case EV_MAX: EV_BIN(ev_max(x, y)); break;
      case EV_NEG: EV_UN(-x); break;
      case EV_ABS: EV_UN(ev_abs(x)); break;
      case EV_SQRT: EV_UN(ev_sqrt(x)); break;
      case EV_SQUARE: EV_UN(ev_mul(x, x)); break;
      case EV_CUBE: EV_UN(ev_mul(ev_mul(x, x), x)); break;
      case EV_EXP: EV_UN(ev_exp(x)); break;
      case EV_LOG: EV_UN(ev_log(x)); break;
      // fused operand forms (made by the host-side peephole pass): arg = input | scalar << 3
      case EV_ADD_IN: EV_OPD(ev_add(x, v)); break;
      case EV_SUB_IN: EV_OPD(ev_sub(x, v)); break;
      case EV_RSUB_IN: EV_OPD(ev_sub(v, x)); break;
      case EV_MUL_IN: EV_OPD(ev_mul(x, v)); break;
      case EV_DIV_IN: EV_OPD(ev_div(x, v)); break;
      case EV_RDIV_IN: EV_OPD(ev_div(v, x)); break;
      case EV_MIN_IN: EV_OPD(ev_min(x, v)); break;
      case EV_MAX_IN: EV_OPD(ev_max(x, v)); break;
      case EV_AXPY_IN: EV_OPDS(ev_add(x, ev_mul(v, s))); break;
      case EV_AXMY_IN: EV_OPDS(ev_sub(x, ev_mul(v, s))); break;
      case EV_ADDS: EV_SCL(ev_add(x, s)); break;
      case EV_SUBS: EV_SCL(ev_sub(x, s)); break;
      case EV_RSUBS: EV_SCL(ev_sub(s, x)); break;
      case EV_MULS: EV_SCL(ev_mul(x, s)); break;
      case EV_RMULS: EV_SCL(ev_mul(s, x)); break;
      case EV_DIVS: EV_SCL(ev_div(x, s)); break;
      case EV_RDIVS: EV_SCL(ev_div(s, x)); break;
      case EV_POWS: EV_SCL(ev_pow(x, s)); break;
      case EV_MINS: EV_SCL(ev_min(x, s)); break;
      case EV_MAXS: EV_SCL(ev_max(x, s)); break;
      default: break;
    }
  }
#undef EV_OPD
#undef EV_OPDS
#undef EV_PUSH
#undef EV_BIN
#undef EV_UN
#undef EV_SCL
}

template <typename ACC>
__device__ __forceinline__ void load_scalars(const EvalDev& a, ACC* sc) {
  if (threadIdx.x < kEvalMaxScalars) {
    const int k = threadIdx.x;
    double v = 0.0;
    if (k < a.nsc) v = a.dsc[k] ? *a.dsc[k] : a.sc[k];
    sc[k] = (ACC)v;
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Vector path: destination and every input share the element type T.  A CTA walks tiles of NV*256 vectors
// (16 bytes each); thread t owns vectors t, t+256, ... of the tile (a warp's accesses stay fully coalesced).
// ring[stage][input][v][thread] holds the operands; `stages` tiles are in flight per CTA.
template <typename T, typename ACC, int NV, int D, bool kStreamStores>
__global__ void __launch_bounds__(kThreads) djets_eval_kernel(T* dest /* may alias an input */, EvalDev a, long long n,
                                                            int stages) {
  using V = typename Vec16<T>::type;
  constexpr int VEC = Vec16<T>::N;
  constexpr int E = NV * VEC;
  extern __shared__ __align__(16) unsigned char ring_raw[];
  V* ring = reinterpret_cast<V*>(ring_raw);
  __shared__ ACC sc[kEvalMaxScalars];
  load_scalars<ACC>(a, sc);
  const int tid = threadIdx.x;
  const int nin = a.nin;
  const long long nvec = (n + VEC - 1) / VEC;
  const long long tile = (long long)NV * kThreads;
  const long long ntiles = (nvec + tile - 1) / tile;
  const int stage_vecs = nin * NV * kThreads;

  auto issue = [&](long long t, int stage) {  // this thread's slots of tile t
    if (t < ntiles) {
      V* dst = ring + (size_t)stage * stage_vecs + tid;
      for (int k = 0; k < nin; ++k) {
        const V* src = reinterpret_cast<const V*>(a.in[k]) + t * tile + tid;
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          const long long idx = t * tile + (long long)u * kThreads + tid;
          // parts carry >= 256 B of slack, so a partial last vector may be read whole; past the end nothing is read
          const bool live = idx < nvec;
          cp_async16(dst + (k * NV + u) * kThreads, live ? src + u * kThreads : reinterpret_cast<const V*>(a.in[k]), live ? 16 : 0);
        }
      }
    }
    cp_async_commit();
  };

  long long t = blockIdx.x;
  for (int s = 0; s < stages - 1; ++s) issue(t + (long long)s * gridDim.x, s);
  int stage = 0;
  for (; t < ntiles; t += gridDim.x) {
    {
      const int ahead = stage == 0 ? stages - 1 : stage - 1;
      issue(t + (long long)(stages - 1) * gridDim.x, ahead);
    }
    // all but the newest stages-1 groups are complete: tile t has landed
    switch (stages) {
      case 2: cp_async_wait<1>(); break;
      case 3: cp_async_wait<2>(); break;
      default: cp_async_wait<3>(); break;
    }
    const V* mine = ring + (size_t)stage * stage_vecs + tid;
    ACC top[E];
    run_program<ACC, E, D>(
        a, sc,
        [&](int k, int e) {
          const T* v = reinterpret_cast<const T*>(mine + (k * NV + e / VEC) * kThreads);
          return (ACC)v[e % VEC];
        },
        top);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const long long idx = t * tile + (long long)u * kThreads + tid;
      if (idx >= nvec) continue;
      V o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < VEC; ++e) oe[e] = (T)top[u * VEC + e];
      if ((idx + 1) * VEC <= n) {
        if (kStreamStores)
          __stcs(reinterpret_cast<V*>(dest) + idx, o);
        else
          reinterpret_cast<V*>(dest)[idx] = o;
      } else {
        for (int e = 0; e < VEC; ++e)
          if (idx * VEC + e < n) dest[idx * VEC + e] = oe[e];
      }
    }
    stage = stage + 1 == stages ? 0 : stage + 1;
  }
  cp_async_wait<0>();
}

// Generic path: any mix of element types (conversions); one element per thread iteration, in double.
template <typename TD>
__global__ void __launch_bounds__(kThreads) djets_eval_generic_kernel(TD* dest, EvalDev a, long long n) {
  __shared__ double sc[kEvalMaxScalars];
  load_scalars<double>(a, sc);
  const long long gstride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += gstride) {
    double top[1];
    run_program<double, 1, kEvalMaxDepth>(
        a, sc,
        [&](int k, int) {
          return a.in_dtype[k] == 0 ? (double) reinterpret_cast<const float*>(a.in[k])[i]
                                    : reinterpret_cast<const double*>(a.in[k])[i];
        },
        top);
    dest[i] = (TD)top[0];
  }
}

int eval_sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <typename T, typename ACC, int NV, int D>
cudaError_t launch_vec_k(T* dest, const EvalDev& d, long long n, cudaStream_t stream) {
  // ring size: about 48 KB per CTA, 2 to 4 stages -> 4 CTAs per SM keep ~100-150 KB in flight per SM
  const size_t stage_bytes = (size_t)std::max(d.nin, 1) * NV * kThreads * 16;
  static const int ring_kb = getenv("DJETS_EVAL_RING_KB") ? atoi(getenv("DJETS_EVAL_RING_KB")) : 32;
  int stages = (int)std::min<size_t>(4, std::max<size_t>(2, ((size_t)ring_kb * 1024) / stage_bytes));
  const size_t smem = stage_bytes * stages;
  static const bool cs = getenv("DJETS_EVAL_CS") && atoi(getenv("DJETS_EVAL_CS"));
  auto kernel = cs ? djets_eval_kernel<T, ACC, NV, D, true> : djets_eval_kernel<T, ACC, NV, D, false>;
  if (smem > 40 * 1024) {  // the kernel also holds a few hundred bytes of static shared memory
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  // Two resident CTAs per SM, not as many as fit: a read+write stream runs faster with few resident threads and the
  // ring keeping the bytes in flight (abs.(d) 5.89 -> 6.03 TB/s, max.(d,0) 5.73 -> 6.14, sqrt.(abs.(x)) in place 5.52 ->
  // 6.00 on 2^30 Float64; scripts/stream_peak.cu shows the same step at 512 resident threads).  DJETS_EVAL_CTAS overrides.
  static const int cap = getenv("DJETS_EVAL_CTAS") ? atoi(getenv("DJETS_EVAL_CTAS")) : 2;
  if (cap > 0) per_sm = std::min(per_sm, cap);
  const long long nvec = (n + Vec16<T>::N - 1) / Vec16<T>::N;
  const long long ntiles = (nvec + (long long)NV * kThreads - 1) / ((long long)NV * kThreads);
  const int grid = (int)std::min<long long>(ntiles, (long long)eval_sm_count() * per_sm);
  kernel<<<grid, kThreads, smem, stream>>>(dest, d, n, stages);
  return cudaGetLastError();
}

template <typename T, typename ACC>
cudaError_t launch_vec(T* dest, const EvalDev& d, long long n, int depth, cudaStream_t stream) {
  constexpr int NV = 4;  // 8 doubles / 16 floats per thread iteration (profiles/r02_eval_sweep.txt)
  static const int nv_env = getenv("DJETS_EVAL_NV") ? atoi(getenv("DJETS_EVAL_NV")) : NV;
  if (nv_env == 8 && depth <= 1) return launch_vec_k<T, ACC, 8, 1>(dest, d, n, stream);
  if (nv_env == 8 && depth <= 2) return launch_vec_k<T, ACC, 8, 2>(dest, d, n, stream);
  if (nv_env == 2 && depth <= 1) return launch_vec_k<T, ACC, 2, 1>(dest, d, n, stream);
  if (nv_env == 2 && depth <= 2) return launch_vec_k<T, ACC, 2, 2>(dest, d, n, stream);
  if (nv_env == 1 && depth <= 1) return launch_vec_k<T, ACC, 1, 1>(dest, d, n, stream);
  if (nv_env == 1 && depth <= 2) return launch_vec_k<T, ACC, 1, 2>(dest, d, n, stream);
  if (depth <= 1) return launch_vec_k<T, ACC, NV, 1>(dest, d, n, stream);
  if (depth <= 2) return launch_vec_k<T, ACC, NV, 2>(dest, d, n, stream);
  if (depth <= 4) return launch_vec_k<T, ACC, 2, 4>(dest, d, n, stream);
  return launch_vec_k<T, ACC, 1, kEvalMaxDepth>(dest, d, n, stream);
}

}  // namespace

int eval_check(const EvalArgs& a, const char** why) {
  static const char* msg = "";
  auto bad = [&](const char* m) {
    msg = m;
    if (why) *why = msg;
    return -1;
  };
  if (a.nops < 1 || a.nops > kEvalMaxOps) return bad("eval: 1 to 48 operations per program");
  if (a.nin < 0 || a.nin > kEvalMaxIn) return bad("eval: at most 8 input vectors");
  if (a.nsc < 0 || a.nsc > kEvalMaxScalars) return bad("eval: at most 16 scalars");
  int sp = 0, depth = 0;
  for (int pc = 0; pc < a.nops; ++pc) {
    const int op = a.op[pc], arg = a.arg[pc];
    if (op == EV_IN) {
      if (arg >= a.nin) return bad("eval: input index outside the input list");
      ++sp;
    } else if (op == EV_SC) {
      if (arg >= a.nsc) return bad("eval: scalar index outside the scalar list");
      ++sp;
    } else if (op >= EV_ADD && op <= EV_MAX) {
      if (sp < 2) return bad("eval: binary operator on fewer than two operands");
      --sp;
    } else if (op >= EV_NEG && op <= EV_LOG) {
      if (sp < 1) return bad("eval: unary operator on an empty stack");
    } else if (op >= EV_ADDS && op <= EV_MAXS) {
      if (sp < 1) return bad("eval: scalar-operand operator on an empty stack");
      if (arg >= a.nsc) return bad("eval: scalar index outside the scalar list");
    } else if (op >= EV_ADD_IN && op <= EV_AXMY_IN) {  // internal forms (peephole pass)
      if (sp < 1) return bad("eval: operand form on an empty stack");
      if ((arg & 7) >= a.nin) return bad("eval: input index outside the input list");
      if (op >= EV_AXPY_IN && (arg >> 3) >= a.nsc) return bad("eval: scalar index outside the scalar list");
    } else {
      return bad("eval: unknown operation code");
    }
    depth = sp > depth ? sp : depth;
    if (sp > kEvalMaxDepth) return bad("eval: expression deeper than 8 operands");
  }
  if (sp != 1) return bad("eval: the program must leave exactly one value");
  return depth;
}

// Peephole pass over a valid program: an operand that is pushed only to be consumed by the next binary operator is
// folded into that operator (`... IN k ADD` -> `ADD_IN k`, `... IN k MULS j ADD` -> `AXPY_IN k,j`), which saves the
// stack shift and a dispatch per operand.  Same arithmetic, same rounding: x*s == s*x and the fused forms round the
// product and the sum separately.
int eval_peephole(const EvalArgs& a, unsigned char* op, unsigned char* arg) {
  int n = 0;
  for (int pc = 0; pc < a.nops;) {
    const bool stacked = n > 0;  // something is already on the stack for the operand to combine with
    if (stacked && a.op[pc] == EV_IN && a.arg[pc] < 8) {
      const int k = a.arg[pc];
      if (pc + 2 < a.nops && (a.op[pc + 1] == EV_MULS || a.op[pc + 1] == EV_RMULS) && a.arg[pc + 1] < 16 &&
          (a.op[pc + 2] == EV_ADD || a.op[pc + 2] == EV_SUB)) {
        op[n] = a.op[pc + 2] == EV_ADD ? EV_AXPY_IN : EV_AXMY_IN;
        arg[n++] = (unsigned char)(k | (a.arg[pc + 1] << 3));
        pc += 3;
        continue;
      }
      if (pc + 1 < a.nops) {
        int f = 0;
        switch (a.op[pc + 1]) {
          case EV_ADD: f = EV_ADD_IN; break;
          case EV_SUB: f = EV_SUB_IN; break;
          case EV_MUL: f = EV_MUL_IN; break;
          case EV_DIV: f = EV_DIV_IN; break;
          case EV_MIN: f = EV_MIN_IN; break;
          case EV_MAX: f = EV_MAX_IN; break;
        }
        if (f) {
          op[n] = (unsigned char)f;
          arg[n++] = (unsigned char)k;
          pc += 2;
          continue;
        }
      }
    }
    op[n] = a.op[pc];
    arg[n++] = a.arg[pc];
    ++pc;
  }
  return n;
}

cudaError_t launch_eval(int ddtype, void* dest, const EvalArgs& a, long long n, int wide, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (eval_check(a, nullptr) < 0) return cudaErrorInvalidValue;
  EvalDev d;
  bool same = true;
  memset(&d, 0, sizeof(d));
  d.nops = eval_peephole(a, d.op, d.arg);
  for (int k = 0; k < kEvalMaxIn; ++k) {
    d.in[k] = k < a.nin ? a.in[k] : nullptr;
    d.in_dtype[k] = k < a.nin ? a.in_dtype[k] : ddtype;
    if (k < a.nin && a.in_dtype[k] != ddtype) same = false;
  }
  for (int k = 0; k < kEvalMaxScalars; ++k) {
    d.sc[k] = k < a.nsc ? a.sc[k] : 0.0;
    d.dsc[k] = k < a.nsc ? a.dsc[k] : nullptr;
  }
  d.nin = a.nin;
  d.nsc = a.nsc;
  // stack depth of the rewritten program
  EvalArgs b = a;
  b.nops = d.nops;
  memcpy(b.op, d.op, sizeof(b.op));
  memcpy(b.arg, d.arg, sizeof(b.arg));
  const int depth = eval_check(b, nullptr);
  if (depth < 0) return cudaErrorInvalidValue;
  if (same) {
    if (ddtype == 1) return launch_vec<double, double>(reinterpret_cast<double*>(dest), d, n, depth, stream);
    if (wide) return launch_vec<float, double>(reinterpret_cast<float*>(dest), d, n, depth, stream);
    return launch_vec<float, float>(reinterpret_cast<float*>(dest), d, n, depth, stream);
  }
  const int grid = (int)std::min<long long>((n + kThreads - 1) / kThreads, 148LL * 8);
  if (ddtype == 1)
    djets_eval_generic_kernel<double><<<grid, kThreads, 0, stream>>>(reinterpret_cast<double*>(dest), d, n);
  else
    djets_eval_generic_kernel<float><<<grid, kThreads, 0, stream>>>(reinterpret_cast<float*>(dest), d, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Device-resident scalars: the master-side folds of the reference's reductions and the scalar arithmetic
// of a solver step (alpha = gamma / delta ...) as one-thread kernels, so that a CG-type iteration needs
// no host round trip and can be captured in a CUDA graph.
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ double round_to(int dtype, double v) { return dtype == 0 ? (double)(float)v : v; }

__global__ void djets_fold_kernel(int dtype, int mode, double p, const double* __restrict__ parts, int nparts, double* out,
                            ScalarPeers peers) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double r;
  if (mode == FOLD_MIN || mode == FOLD_MAX) {
    r = mode == FOLD_MIN ? INFINITY : -INFINITY;
    for (int k = 0; k < nparts; ++k) {
      const double z = round_to(dtype, parts[k]);
      r = (r != r || z != z) ? NAN : (mode == FOLD_MIN ? fmin(r, z) : fmax(r, z));
    }
  } else if (mode == FOLD_SUM) {
    r = 0.0;
    for (int k = 0; k < nparts; ++k) r = round_to(dtype, r + round_to(dtype, parts[k]));
  } else {  // src:206-209: per-worker norms z, then (sum z^p)^(1/p) in float(real(T))
    const bool two = mode == FOLD_NORM2;
    const double pp = round_to(dtype, p);
    double sum = 0.0;
    for (int k = 0; k < nparts; ++k) {
      const double zk = round_to(dtype, two ? sqrt(parts[k]) : pow(parts[k], 1.0 / pp));
      sum = round_to(dtype, sum + round_to(dtype, two ? zk * zk : pow(zk, pp)));
    }
    r = round_to(dtype, two ? sqrt(sum) : pow(sum, round_to(dtype, 1.0 / pp)));
  }
  out[0] = r;
  for (int k = 0; k < peers.n; ++k) peers.p[k][0] = r;
}

__global__ void djets_scalar_op_kernel(int dtype, int op, const double* a, const double* b, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double x = a ? a[0] : 0.0, y = b ? b[0] : 0.0;
  double r = 0.0;
  switch (op) {
    case SC_ADD: r = x + y; break;
    case SC_SUB: r = x - y; break;
    case SC_MUL: r = x * y; break;
    case SC_DIV: r = x / y; break;
    case SC_NEG: r = -x; break;
    case SC_SQRT: r = sqrt(x); break;
    case SC_COPY: r = x; break;
    case SC_ABS: r = fabs(x); break;
    case SC_MAX: r = (x != x || y != y) ? NAN : fmax(x, y); break;
    case SC_MIN: r = (x != x || y != y) ? NAN : fmin(x, y); break;
  }
  out[0] = round_to(dtype, r);
}

__global__ void djets_scalar_set_kernel(double v, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = v;
}

}  // namespace

cudaError_t launch_fold(int dtype, int mode, double p, const double* parts, int nparts, double* out, ScalarPeers peers,
                        cudaStream_t stream) {
  djets_fold_kernel<<<1, 32, 0, stream>>>(dtype, mode, p, parts, nparts, out, peers);
  return cudaGetLastError();
}

cudaError_t launch_scalar_op(int dtype, int op, const double* a, const double* b, double* out, cudaStream_t stream) {
  if (op < SC_ADD || op > SC_MIN) return cudaErrorInvalidValue;
  djets_scalar_op_kernel<<<1, 32, 0, stream>>>(dtype, op, a, b, out);
  return cudaGetLastError();
}

cudaError_t launch_scalar_set(double v, double* out, cudaStream_t stream) {
  djets_scalar_set_kernel<<<1, 32, 0, stream>>>(v, out);
  return cudaGetLastError();
}

}  // namespace djets
