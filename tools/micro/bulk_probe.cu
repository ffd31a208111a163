// bulk_probe.cu -- what does a cp.async.bulk (global -> shared, mbarrier complete_tx)
// stream cost on B200, per copy and in aggregate?  Written after the first version of
// the bulk-copy-fed SpMV kernel (csrc/sell_tma.cuh) measured ~10 us per slice per warp
// whatever the ring depth, the warp count or the bytes per copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bulk_probe bulk_probe.cu
//   ./bulk_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *b) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
template <int FENCE, int HINT>
__device__ __forceinline__ void fetch(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  const uint32_t b = s32(bar);
  if (FENCE)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  if (HINT)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(b), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
// WAITMODE 0: try_wait loop (hardware suspend), 1: test_wait spin
template <int WAITMODE>
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = s32(bar);
  uint32_t ok = 0;
  do {
    if (WAITMODE == 0)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}

// ---- A: latency of one copy, one warp on one SM ----------------------------------------------
__global__ void k_latency(const char *src, uint32_t bytes, int reps, long long *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) bar_init(&bar);
  fence_init();
  __syncwarp();
  long long tot = 0, mx = 0;
  uint32_t ph = 0;
  for (int r = 0; r < reps; r++) {
    long long t0 = clock64();
    if (threadIdx.x == 0)
      fetch<1, 0>(smem, src + (size_t)r * 65536, bytes, &bar, 0);
    wait<0>(&bar, ph);
    long long t1 = clock64();
    ph ^= 1;
    tot += t1 - t0;
    mx = t1 - t0 > mx ? t1 - t0 : mx;
    __syncwarp();
  }
  if (threadIdx.x == 0) out[0] = tot / reps, out[1] = mx;
}

// ---- B: streaming, W warps x S stages per SM, consumer reads one word per lane per stage -------
template <int FENCE, int HINT, int WAITMODE>
__global__ void k_stream(const char *src, size_t total, uint32_t bytes, int nstages, double *sink,
                         long long *prof) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[32 * 8];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  unsigned char *ring = smem + (size_t)warp * nstages * bytes;
  uint64_t *bar = bars + warp * 8;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (lane == 0)
    for (int i = 0; i < nstages; i++) bar_init(bar + i);
  fence_init();
  __syncwarp();
  const size_t nchunks = total / bytes, stride = (size_t)gridDim.x * W;
  const size_t v0 = (size_t)blockIdx.x * W + warp;
  if (lane == 0)
    for (int i = 0; i < nstages; i++)
      if (v0 + i * stride < nchunks)
        fetch<FENCE, HINT>(ring + (size_t)i * bytes, src + (v0 + i * stride) * bytes, bytes, bar + i, pol);
  double acc = 0;
  int stage = 0;
  uint32_t ph = 0;
  long long t_wait = 0, t_all0 = clock64();
  for (size_t v = v0; v < nchunks; v += stride) {
    long long t0 = clock64();
    wait<WAITMODE>(bar + stage, ph);
    t_wait += clock64() - t0;
    acc += reinterpret_cast<const double *>(ring + (size_t)stage * bytes)[lane];
    __syncwarp();
    const size_t vn = v + (size_t)nstages * stride;
    if (lane == 0 && vn < nchunks)
      fetch<FENCE, HINT>(ring + (size_t)stage * bytes, src + vn * bytes, bytes, bar + stage, pol);
    if (++stage == nstages) stage = 0, ph ^= 1;
  }
  if (acc == 1.2345) sink[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) prof[0] = t_wait, prof[1] = clock64() - t_all0;
}

// ---- B2: the same ring with the consumer of an SpMV grown feature by feature ---------------------
// FEAT bit 0: all 27 values of the stage through LDS + fma     bit 1: + one y store per lane
//      bit 2: + 27 gathers x[row + d_k] (27-point deltas of a 256^3 grid), on demand
//      bit 3: the gathers of the NEXT chunk are requested before this chunk is consumed
template <int FEAT>
__global__ void k_stream2(const char *src, size_t total, int nstages, const double *__restrict__ x,
                          double *__restrict__ y, size_t nx, double *sink, long long *prof) {
  constexpr uint32_t bytes = 6912;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[32 * 8];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  unsigned char *ring = smem + (size_t)warp * nstages * bytes;
  uint64_t *bar = bars + warp * 8;
  if (lane == 0)
    for (int i = 0; i < nstages; i++) bar_init(bar + i);
  fence_init();
  __syncwarp();
  const size_t nchunks = total / bytes, stride = (size_t)gridDim.x * W;
  const size_t v0 = (size_t)blockIdx.x * W + warp;
  if (lane == 0)
    for (int i = 0; i < nstages; i++)
      if (v0 + i * stride < nchunks)
        fetch<1, 1>(ring + (size_t)i * bytes, src + (v0 + i * stride) * bytes, bytes, bar + i, 0);
  const int N = 256;
  auto gather = [&](size_t v, double (&xv)[27]) {
    const long long row = (long long)((v * 32 + lane) % (nx - 2 * (N * N + N + 1))) + N * N + N + 1;
    int k = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; dz++)
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
          xv[k++] = __ldg(x + row + dz * N * N + dy * N + dx);
  };
  double acc = 0, xa[27], xb[27];
#pragma unroll
  for (int k = 0; k < 27; k++) xa[k] = 1.0, xb[k] = 1.0;
  if ((FEAT & 8) && v0 < nchunks) gather(v0, xa);
  int stage = 0;
  uint32_t ph = 0;
  long long t_wait = 0, t_all0 = clock64();
  auto step = [&](size_t v, double (&cur)[27], double (&nxt)[27]) {
    if ((FEAT & 8) && v + stride < nchunks) gather(v + stride, nxt);
    if ((FEAT & 4) && !(FEAT & 8)) gather(v, cur);
    long long t0 = clock64();
    wait<0>(bar + stage, ph);
    t_wait += clock64() - t0;
    const double *sv = reinterpret_cast<const double *>(ring + (size_t)stage * bytes) + lane;
    double sum = 0;
    if (FEAT & 1) {
#pragma unroll
      for (int k = 0; k < 27; k++) sum = fma(sv[k * 32], cur[k], sum);
    } else {
      sum = sv[0];
    }
    __syncwarp();
    const size_t vn = v + (size_t)nstages * stride;
    if (lane == 0 && vn < nchunks)
      fetch<1, 1>(ring + (size_t)stage * bytes, src + vn * bytes, bytes, bar + stage, 0);
    if (FEAT & 2) y[(v * 32 + lane) % nx] = sum;
    else acc += sum;
    if (++stage == nstages) stage = 0, ph ^= 1;
  };
  for (size_t v = v0; v < nchunks; v += 2 * stride) {
    step(v, xa, xb);
    if (v + stride < nchunks) step(v + stride, xb, xa);
  }
  if (acc == 1.2345) sink[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) prof[0] = t_wait, prof[1] = clock64() - t_all0;
}

template <int FEAT>
static void run_stream2(const char *d, size_t total, int W, int S, const double *x, double *y, size_t nx,
                        double *sink, long long *prof) {
  size_t smem = (size_t)W * S * 6912;
  auto k = k_stream2<FEAT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<<<148, W * 32, smem>>>(d, total, S, x, y, nx, sink, prof);
  cudaEventRecord(e0);
  k<<<148, W * 32, smem>>>(d, total, S, x, y, nx, sink, prof);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[2];
  cudaMemcpy(h, prof, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("consumer feat=%2d (%s%s%s%s)  W=%2d S=%d  %8.3f ms  %7.1f GB/s of values  warp0: %.0f%% in wait  %s\n", FEAT,
         FEAT & 1 ? "lds+fma " : "", FEAT & 2 ? "ystore " : "", FEAT & 4 ? "gathers " : "", FEAT & 8 ? "ahead" : "",
         W, S, ms, total / ms / 1e6, h[1] ? 100.0 * h[0] / h[1] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// ---- C: the same stream with plain 16-byte loads (what a copy kernel does), for scale ------------
__global__ void k_ldg(const double2 *src, size_t n16, double *sink) {
  double acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = __ldcs(src + i);
    acc += v.x + v.y;
  }
  if (acc == 1.2345) sink[0] = acc;
}

template <int FENCE, int HINT, int WAITMODE>
static void run_stream(const char *d, size_t total, int W, int S, uint32_t bytes, double *sink, long long *prof) {
  size_t smem = (size_t)W * S * bytes;
  if (smem > 220 * 1024) return;
  auto k = k_stream<FENCE, HINT, WAITMODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<<<148, W * 32, smem>>>(d, total, bytes, S, sink, prof);
  cudaEventRecord(e0);
  k<<<148, W * 32, smem>>>(d, total, bytes, S, sink, prof);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[2];
  cudaMemcpy(h, prof, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("stream fence=%d hint=%d wait=%s  W=%2d S=%d bytes=%5u  %8.3f ms  %7.1f GB/s  warp0: %.0f%% of its time in wait  %s\n",
         FENCE, HINT, WAITMODE ? "test" : "try ", W, S, bytes, ms, total / ms / 1e6,
         h[1] ? 100.0 * h[0] / h[1] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const size_t total = (size_t)4 << 30;
  char *d;
  double *sink;
  long long *prof;
  cudaMalloc(&d, total);
  cudaMemset(d, 1, total);
  cudaMalloc(&sink, 8);
  cudaMalloc(&prof, 64);
  for (uint32_t bytes : {256u, 2048u, 6912u, 16384u, 32768u}) {
    cudaFuncSetAttribute(k_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_latency<<<1, 32, 65536>>>(d, bytes, 200, prof);
    long long h[2];
    cudaMemcpy(h, prof, 16, cudaMemcpyDeviceToHost);
    printf("latency: one copy of %5u bytes, issue -> wait passed: mean %lld cycles, max %lld  %s\n", bytes, h[0], h[1],
           cudaGetErrorString(cudaGetLastError()));
  }
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k_ldg<<<148 * 8, 256>>>((const double2 *)d, total / 16, sink);
    cudaEventRecord(e0);
    k_ldg<<<148 * 8, 256>>>((const double2 *)d, total / 16, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.128 stream (2048 threads/SM): %.3f ms  %.1f GB/s\n", ms, total / ms / 1e6);
  }
  for (int W : {1, 4, 8, 16})
    for (int S : {2, 3, 4})
      run_stream<1, 1, 0>(d, total, W, S, 6912, sink, prof);
  run_stream<0, 1, 0>(d, total, 8, 3, 6912, sink, prof);
  run_stream<1, 0, 0>(d, total, 8, 3, 6912, sink, prof);
  run_stream<0, 0, 0>(d, total, 8, 3, 6912, sink, prof);
  run_stream<1, 1, 1>(d, total, 8, 3, 6912, sink, prof);
  run_stream<0, 0, 1>(d, total, 8, 3, 6912, sink, prof);
  for (uint32_t bytes : {2048u, 3456u, 13824u, 27648u})
    run_stream<0, 0, 0>(d, total, 8, 3, bytes, sink, prof), run_stream<0, 0, 0>(d, total, 4, 2, bytes, sink, prof);
  {
    const size_t nx = (size_t)256 * 256 * 256, tot2 = (size_t)1 << 30;
    double *x, *y;
    cudaMalloc(&x, nx * 8), cudaMalloc(&y, nx * 8);
    cudaMemset(x, 0, nx * 8);
    for (int W : {8, 12}) {
      run_stream2<0>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<1>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<3>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<5>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<7>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<13>(d, tot2, W, 2, x, y, nx, sink, prof);
      run_stream2<15>(d, tot2, W, 2, x, y, nx, sink, prof);
    }
  }
  return 0;
}
