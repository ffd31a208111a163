// gather_probe2.cu -- two questions behind the power-law SpMV (BASELINE.json config 5):
//  (1) a random 8-byte gather costs ~101 B of DRAM traffic (gather_probe.cu): is that the L2 fetch
//      granularity (cudaLimitMaxL2FetchGranularity, 32 / 64 / 128 B), and can it be turned down?
//  (2) how much of x stays resident in L2 when the gathers of one pass fall into ONE range of x while
//      the matrix streams past (12 B per gather, evict-first)?  A whole "far half" SpMV is emulated:
//      400/R passes, pass k gathers from range k only; reported: ms for all passes together.
//      With and without an L2 persisting window on the range.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_probe2 gather_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// every thread: `per` gathers from x[base, base+range) and, when STREAM, 12 B per gather of a
// coalesced evict-first stream (the matrix).  grid-stride over chunks of 8 gathers so that neighbouring
// threads stream neighbouring memory.
template <bool STREAM>
__global__ void __launch_bounds__(256) k_pass(const double *__restrict__ x, uint64_t base, uint64_t range,
                                              const uint4 *__restrict__ mat, uint64_t chunks, uint64_t salt,
                                              double *out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, T = (uint64_t)gridDim.x * blockDim.x;
  double s = 0;
  for (uint64_t c = t; c < chunks; c += T) {
    uint4 m[6];
    if (STREAM) {
#pragma unroll
      for (int j = 0; j < 6; j++)
        m[j] = __ldcs(mat + (uint64_t)j * chunks + c);     // 6 x 16 B = 96 B = 8 x 12 B
    }
    uint64_t idx[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      idx[j] = base + mix(salt + c * 8 + j) % range;
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      v[j] = __ldg(x + idx[j]);
#pragma unroll
    for (int j = 0; j < 8; j++)
      s += v[j];
    if (STREAM) {
#pragma unroll
      for (int j = 0; j < 6; j++)
        s += (double)(m[j].x ^ m[j].y ^ m[j].z ^ m[j].w);
    }
  }
  if (s == 123.456)
    out[0] = s;
}

static float far_half(const double *x, uint64_t n, uint64_t range, const uint4 *mat, uint64_t gathers,
                      double *out, bool stream, bool persist, cudaStream_t st) {
  uint64_t passes = (n + range - 1) / range;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a, st);
    for (uint64_t p = 0; p < passes; p++) {
      uint64_t lo = p * range, len = lo + range <= n ? range : n - lo;
      uint64_t chunks = gathers * len / n / 8;
      if (persist) {
        cudaStreamAttrValue v = {};
        v.accessPolicyWindow.base_ptr = (void *)(x + lo);
        v.accessPolicyWindow.num_bytes = len * 8;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
      }
      if (stream)
        k_pass<true><<<148 * 8, 256, 0, st>>>(x, lo, len, mat, chunks, p * 1000003ull + rep, out);
      else
        k_pass<false><<<148 * 8, 256, 0, st>>>(x, lo, len, mat, chunks, p * 1000003ull + rep, out);
    }
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best)
      best = ms;
  }
  if (persist) {
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaCtxResetPersistingL2Cache();
  }
  return best;
}

int main(int argc, char **argv) {
  uint64_t mb = argc > 1 ? atoll(argv[1]) : 400;
  int only = argc > 2 ? atoi(argv[2]) : 0;            // 1: granularity part only (for ncu)
  uint64_t n = mb * (1 << 20) / 8;
  const uint64_t gathers = 387000000ull;              // the far half of the 50 M-row operator
  double *x, *out;
  uint4 *mat;
  cudaMalloc(&x, n * 8);
  cudaMalloc(&out, 8);
  cudaMalloc(&mat, gathers * 12 + 4096);
  cudaMemset(x, 0, n * 8);
  cudaMemset(mat, 0, gathers * 12 + 4096);
  cudaStream_t st;
  cudaStreamCreate(&st);
  size_t g0 = 0;
  cudaDeviceGetLimit(&g0, cudaLimitMaxL2FetchGranularity);
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", g0);
  int grans[4] = {(int)g0, 32, 64, 128};
  for (int gi = 0; gi < 4; gi++) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, grans[gi]);
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    float ms = far_half(x, n, n, mat, gathers, out, false, false, st);
    float ms2 = far_half(x, n, n, mat, gathers, out, true, false, st);
    printf("granularity set %3d (%s) reads back %3zu: whole-x gathers %.3f ms (%.1f Ggather/s); with the 12 B/gather "
           "stream %.3f ms\n", grans[gi], cudaGetErrorString(e), g, ms, gathers / ms / 1e6, ms2);
  }
  if (only == 1)
    return 0;
  int gsel[2] = {(int)g0, 32};
  for (int gi = 0; gi < 2; gi++) {
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gsel[gi]);
    size_t maxp = 0;
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, pr.persistingL2CacheMaxSize);
    cudaDeviceGetLimit(&maxp, cudaLimitPersistingL2CacheSize);
    printf("-- granularity %d; L2 %d MB, persisting set-aside %zu MB (max %d MB), window max %d MB\n", gsel[gi],
           pr.l2CacheSize >> 20, maxp >> 20, pr.persistingL2CacheMaxSize >> 20, pr.accessPolicyMaxWindowSize >> 20);
    int ranges[] = {8, 16, 24, 32, 48, 64, 96, 128, 200, 400};
    for (int ri = 0; ri < 10; ri++) {
      uint64_t r = (uint64_t)ranges[ri] * (1 << 20) / 8;
      if (r > n)
        r = n;
      float a = far_half(x, n, r, mat, gathers, out, false, false, st);
      float b = far_half(x, n, r, mat, gathers, out, true, false, st);
      float c = far_half(x, n, r, mat, gathers, out, true, true, st);
      printf("range %3d MB (%3llu passes): gathers alone %.3f ms; + stream %.3f ms; + stream, persisting window %.3f ms\n",
             ranges[ri], (unsigned long long)((n + r - 1) / r), a, b, c);
    }
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  }
  printf("stream alone (4.6 GB): ");
  {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a, st);
    k_pass<true><<<148 * 8, 256, 0, st>>>(x, 0, 1, mat, gathers / 8, 1, out);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%.3f ms\n", ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
