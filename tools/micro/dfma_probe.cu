// dfma_probe.cu -- DFMA dependent latency and per-SM throughput on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, double a, double b, int n) {
  double s = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) s = fma(s, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / n;
  if (s == 1.2345) out[1] = s;
}
__global__ void k_thr(double *out, double a, double b, int n) {
  double s0 = threadIdx.x, s1 = s0 + 1, s2 = s0 + 2, s3 = s0 + 3, s4 = s0 + 4, s5 = s0 + 5, s6 = s0 + 6, s7 = s0 + 7;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    s0 = fma(s0, a, b); s1 = fma(s1, a, b); s2 = fma(s2, a, b); s3 = fma(s3, a, b);
    s4 = fma(s4, a, b); s5 = fma(s5, a, b); s6 = fma(s6, a, b); s7 = fma(s7, a, b);
  }
  long long t1 = clock64();
  double s = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
  if (s == 1.2345) out[1] = s;
}
int main() {
  double *d, h[2];
  cudaMalloc(&d, 16);
  k_lat<<<1, 32>>>(d, 1.0000001, 1e-9, 100000);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", h[0]);
  for (int threads : {128, 256, 512, 1024}) {
    int n = 20000;
    k_thr<<<1, threads>>>(d, 1.0000001, 1e-9, n);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("DFMA throughput, 1 CTA of %4d threads: %.2f lane-FMA/clk/SM\n", threads,
           (double)threads * 8 * n / h[0]);
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_thr<<<148 * 2, 1024>>>(d, 1.0000001, 1e-9, 20000);
  cudaEventRecord(a);
  k_thr<<<148 * 2, 1024>>>(d, 1.0000001, 1e-9, 20000);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("whole GPU: %.2f TFLOP/s fp64 (FMA = 2 flop)\n", 148.0 * 2 * 1024 * 8 * 20000 * 2 / ms / 1e9);
  return 0;
}
