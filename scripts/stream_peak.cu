// Read-only and copy "stream peak" of one B200: the denominators next to which the roofline fractions of the block
// kernels are read (profiles/r02_stream_peak.txt).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/stream_peak scripts/stream_peak.cu
//   scripts/_bin/stream_peak [GiB=8]
// Kernels: 128-bit ld.global.nc.L1::no_allocate loads (the block kernels' load), U independent loads in flight per
// thread, grid = 148 SMs x CTAs per SM, grid-stride over the buffer; the sum is kept alive through one store per CTA.
// Every timing is CUDA events over `reps` back-to-back launches on a buffer far larger than the 126 MB L2.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_cs(uint4* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int U>
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ src, size_t n16, unsigned* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_nc(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) {
    const uint4 v = ld_nc(src + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9e3779b9u) sink[blockIdx.x] = acc;  // practically never taken: the loads cannot be dropped
}

template <int U>
__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_nc(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) st_cs(dst + i + u * stride, v[u]);
  }
  for (; i < n16; i += stride) st_cs(dst + i, ld_nc(src + i));
}

// copy with wider phases: U loads then U stores per thread, any block size
template <int U>
__global__ void __launch_bounds__(1024) copy_wide_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_nc(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) st_cs(dst + i + u * stride, v[u]);
  }
  for (; i < n16; i += stride) st_cs(dst + i, ld_nc(src + i));
}

__global__ void __launch_bounds__(256) fill_kernel(uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(1u, 2u, 3u, 4u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) st_cs(dst + i, v);
}

// ---- store variants of the fill -------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void st_mode(uint4* p, uint4 v) {
  if (MODE == 0) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 3) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 4) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
  }
}
// contiguous: every CTA owns one contiguous span of the buffer and walks it front to back (4 KB per step)
template <int MODE, bool CONTIG>
__global__ void __launch_bounds__(256) fill_mode_kernel(uint4* __restrict__ dst, size_t n16) {
  const uint4 v = make_uint4(1u, 2u, 3u, 4u);
  if (CONTIG) {
    const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
    const size_t lo = per * blockIdx.x, hi = lo + per < n16 ? lo + per : n16;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) st_mode<MODE>(dst + i, v);
  } else {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) st_mode<MODE>(dst + i, v);
  }
}

// ---- copy through shared memory with 1-D TMA bulk transfers (no registers on the data path) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(32) tma_copy_kernel(const char* __restrict__ src, char* __restrict__ dst, size_t nchunks) {
  extern __shared__ __align__(128) char buf[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const size_t first = blockIdx.x, step = gridDim.x;
  const size_t mine = first < nchunks ? (nchunks - first + step - 1) / step : 0;
  auto load = [&](size_t n) {
    const int s = (int)(n % STAGES);
    const uint32_t b = smem_u32(&bar[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(buf + (size_t)s * CHUNK)),
                 "l"(src + (first + n * step) * (size_t)CHUNK), "r"(CHUNK), "r"(b)
                 : "memory");
  };
  for (size_t n = 0; n < (size_t)STAGES && n < mine; ++n) load(n);
  for (size_t n = 0; n < mine; ++n) {
    const int s = (int)(n % STAGES);
    const uint32_t parity = (uint32_t)((n / STAGES) & 1);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar[s])), "r"(parity)
                   : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (first + n * step) * (size_t)CHUNK),
                 "r"(smem_u32(buf + (size_t)s * CHUNK)), "r"(CHUNK)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the previous store has read its buffer
    if (n >= 1 && n - 1 + STAGES < mine) load(n - 1 + STAGES);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static double time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int k = 0; k < 3; ++k) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int k = 0; k < reps; ++k) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 8.0;
  const size_t bytes = (size_t)(gib * (1ull << 30)) & ~(size_t)4095;
  const size_t n16 = bytes / 16;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint4 *src = nullptr, *dst = nullptr;
  unsigned* sink = nullptr;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&dst, bytes));
  CK(cudaMalloc(&sink, 1 << 20));
  CK(cudaMemset(src, 1, bytes));
  CK(cudaMemset(dst, 0, bytes));
  const int reps = 10;
  printf("{\"buffer_GiB\": %.2f, \"sms\": %d, \"reps\": %d}\n", bytes / 1073741824.0, sms, reps);
  double best_read = 0.0, best_copy = 0.0;
  for (int per_sm : {2, 4, 8}) {
    const int grid = sms * per_sm;
    auto report = [&](const char* what, int u, double ms, double traffic) {
      const double gbps = traffic / ms * 1e-6;
      printf("{\"kernel\": \"%s\", \"loads_in_flight\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", what, u, per_sm, ms, gbps);
      return gbps;
    };
    double r;
    r = report("read", 4, time_ms([&] { read_kernel<4><<<grid, 256>>>(src, n16, sink); }, reps), (double)bytes);
    best_read = r > best_read ? r : best_read;
    r = report("read", 8, time_ms([&] { read_kernel<8><<<grid, 256>>>(src, n16, sink); }, reps), (double)bytes);
    best_read = r > best_read ? r : best_read;
    r = report("read", 16, time_ms([&] { read_kernel<16><<<grid, 256>>>(src, n16, sink); }, reps), (double)bytes);
    best_read = r > best_read ? r : best_read;
    r = report("copy", 4, time_ms([&] { copy_kernel<4><<<grid, 256>>>(src, dst, n16); }, reps), 2.0 * bytes);
    best_copy = r > best_copy ? r : best_copy;
    r = report("copy", 8, time_ms([&] { copy_kernel<8><<<grid, 256>>>(src, dst, n16); }, reps), 2.0 * bytes);
    best_copy = r > best_copy ? r : best_copy;
  }
  {
    const char* names[5] = {"st", "st.cs", "st.cg", "st.wt", "st.L2::evict_first"};
    double ms[5][2];
    ms[0][0] = time_ms([&] { fill_mode_kernel<0, false><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[1][0] = time_ms([&] { fill_mode_kernel<1, false><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[2][0] = time_ms([&] { fill_mode_kernel<2, false><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[3][0] = time_ms([&] { fill_mode_kernel<3, false><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[4][0] = time_ms([&] { fill_mode_kernel<4, false><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[0][1] = time_ms([&] { fill_mode_kernel<0, true><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[1][1] = time_ms([&] { fill_mode_kernel<1, true><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[2][1] = time_ms([&] { fill_mode_kernel<2, true><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[3][1] = time_ms([&] { fill_mode_kernel<3, true><<<sms * 8, 256>>>(dst, n16); }, reps);
    ms[4][1] = time_ms([&] { fill_mode_kernel<4, true><<<sms * 8, 256>>>(dst, n16); }, reps);
    for (int m = 0; m < 5; ++m)
      printf("{\"kernel\": \"fill\", \"store\": \"%s\", \"grid_stride_GBps\": %.0f, \"contiguous_per_cta_GBps\": %.0f}\n", names[m],
             bytes / ms[m][0] * 1e-6, bytes / ms[m][1] * 1e-6);
  }
  {
    auto tma = [&](auto kernel, int chunk, int stages, int per_sm) {
      const size_t smem = (size_t)chunk * stages;
      CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const size_t nchunks = bytes / chunk;
      const double ms = time_ms([&] { kernel<<<sms * per_sm, 32, smem>>>((const char*)src, (char*)dst, nchunks); }, reps);
      printf("{\"kernel\": \"tma_copy\", \"chunk\": %d, \"stages\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", chunk, stages,
             per_sm, ms, 2.0 * bytes / ms * 1e-6);
    };
    tma(tma_copy_kernel<8192, 4>, 8192, 4, 4);
    tma(tma_copy_kernel<8192, 4>, 8192, 4, 6);
    tma(tma_copy_kernel<16384, 4>, 16384, 4, 2);
    tma(tma_copy_kernel<16384, 4>, 16384, 4, 3);
    tma(tma_copy_kernel<32768, 3>, 32768, 3, 2);
    tma(tma_copy_kernel<4096, 8>, 4096, 8, 6);
  }
  for (int threads : {256, 512, 1024})
    for (int per_sm : {1, 2, 3}) {
      if (threads * per_sm > 2048) continue;
      const int grid = sms * per_sm;
      const double m4 = time_ms([&] { copy_wide_kernel<4><<<grid, threads>>>(src, dst, n16); }, reps);
      const double m8 = time_ms([&] { copy_wide_kernel<8><<<grid, threads>>>(src, dst, n16); }, reps);
      const double m16 = threads <= 512 ? time_ms([&] { copy_wide_kernel<16><<<grid, threads>>>(src, dst, n16); }, reps) : 0.0;
      printf("{\"kernel\": \"copy_wide\", \"threads\": %d, \"ctas_per_sm\": %d, \"U4_GBps\": %.0f, \"U8_GBps\": %.0f, \"U16_GBps\": %.0f}\n", threads,
             per_sm, 2.0 * bytes / m4 * 1e-6, 2.0 * bytes / m8 * 1e-6, m16 > 0 ? 2.0 * bytes / m16 * 1e-6 : 0.0);
    }
  const double fill = (double)bytes / time_ms([&] { fill_kernel<<<sms * 8, 256>>>(dst, n16); }, reps) * 1e-6;
  const double memcpy_gbps =
      2.0 * bytes / time_ms([&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, 0)); }, reps) * 1e-6;
  const double memset_gbps = (double)bytes / time_ms([&] { CK(cudaMemsetAsync(dst, 0, bytes, 0)); }, reps) * 1e-6;
  printf("{\"stream_peak\": {\"read_GBps\": %.0f, \"copy_GBps\": %.0f, \"fill_GBps\": %.0f, \"cudaMemcpy_d2d_GBps\": %.0f, "
         "\"cudaMemset_GBps\": %.0f}}\n",
         best_read, best_copy, fill, memcpy_gbps, memset_gbps);
  return 0;
}
