// gather_probe.cu -- which load flavour makes a random 8-byte gather cheapest
// on B200?  (power-law SpMV: ncu shows ~3 sectors of L2/DRAM traffic per gathered
// x entry with LDG.E.64.CONSTANT.)   nvcc -arch=sm_100a -O3 -lineinfo
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <int MODE> __device__ __forceinline__ double ld(const double *p) {
  double v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = *(const volatile double *)p;   // plain LDG (volatile: no .nc)
  else if (MODE == 2) v = __ldcg(p);
  else if (MODE == 3) v = __ldcs(p);
  else if (MODE == 4) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 6) asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 7) asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else v = __ldlu(p);
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const double *__restrict__ x, uint64_t n,
                                                uint64_t per_thread, double *out) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  double s = 0;
  for (uint64_t i = 0; i < per_thread; i += 8) {
    uint64_t idx[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      idx[j] = mix(t * per_thread + i + j) % n;
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      v[j] = ld<MODE>(x + idx[j]);
#pragma unroll
    for (int j = 0; j < 8; j++)
      s += v[j];
  }
  if (s == 123.456)
    out[0] = s;
}

template <int MODE> void run(const double *x, uint64_t n, double *out, const char *name) {
  const int grid = 148 * 8, per = 256 * 4;  // 310 M gathers
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  k_gather<MODE><<<grid, 256>>>(x, n, per, out);
  cudaEventRecord(a);
  k_gather<MODE><<<grid, 256>>>(x, n, per, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  double g = (double)grid * 256 * per;
  printf("%-34s n=%llu MB  %.3f ms  %.1f Ggather/s  (32B-sector rate %.0f GB/s) %s\n", name,
         (unsigned long long)(n * 8 >> 20), ms, g / ms / 1e6, g * 32 / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
  uint64_t mb = argc > 1 ? atoll(argv[1]) : 400;
  uint64_t n = mb * (1 << 20) / 8;
  double *x, *out;
  cudaMalloc(&x, n * 8);
  cudaMalloc(&out, 8);
  cudaMemset(x, 0, n * 8);
  run<0>(x, n, out, "__ldg (LDG.CONSTANT)");
  run<1>(x, n, out, "plain ld.global (volatile)");
  run<2>(x, n, out, "__ldcg (ld.global.cg)");
  run<3>(x, n, out, "__ldcs (ld.global.cs)");
  run<4>(x, n, out, "ld.global.L1::no_allocate");
  run<5>(x, n, out, "ld.global.nc.L1::no_allocate");
  run<6>(x, n, out, "ld.global.nc.L1::evict_last");
  run<7>(x, n, out, "ld.global.cv");
  run<8>(x, n, out, "__ldlu");
  return 0;
}
