// spmv_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the bodies of the SELL SpMV
// kernels (lsbench_b200/csrc/sell_kernels.cuh: k_spmv_sell, k_spmv_sellc) for the
// host, with one-line stand-ins for the CUDA built-ins, and runs them thread by
// thread over a launch grid.  Without the fused dot product a thread talks to no other thread,
// so running the threads one after another is exactly what the GPU computes.
// The caller (tests/test_spmv_emul.py) builds the index-compressed SELL layout
// of DESIGN.md section 2 with numpy and compares y with a CSR product bit for bit.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

struct uint4 {
  unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) {
  return uint4{x, y, z, w};
}
struct Dim {
  unsigned x;
};
static Dim threadIdx, blockIdx, gridDim;
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define B2_SLICE 32
struct double2 {
  double x, y;
};
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T ld_stream(const T *p) { return *p; }
// (the row-major kernels of the same header reduce over a warp: they compile here
// and run on the SIMT emulator of pcg_emul.cpp, not thread by thread)
static inline double warp_sum(double v) { return v; }
// (k_spmv_sell_grp hands out its work through a shared counter: compiled here, run on the
// SIMT emulator of pcg_emul.cpp)
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
struct PcgState {
  int done;
  unsigned ticket[4];
};
struct XrArgs {};
template <int N> static double block_sum(double v, double *) { return v; }
template <int NV, int NW>
static void grid_sum_finish(const double (&)[NV], double *, unsigned, unsigned, unsigned,
                            unsigned *, double *, double *, const XrArgs &) {}

#include "sell_kernels.cuh"

// ---- the default kernels: k_spmv_sellc (index-compressed) and k_spmv_sell ----------
template <typename VT>
static void run_sellc(unsigned grid, const uint4 *meta, const uint32_t *cols, const int32_t *dcols,
                      const VT *vals, const uint32_t *perm, const double *x, double *y,
                      uint32_t b0, uint32_t e0, uint32_t b1, uint32_t e1, uint32_t n_rows) {
  gridDim.x = grid;
  for (unsigned b = 0; b < grid; b++)
    for (unsigned t = 0; t < SPMV_THREADS; t++) {
      blockIdx.x = b, threadIdx.x = t;
      k_spmv_sellc<false, VT>(meta, cols, dcols, vals, perm, x, y, b0, e0, b1, e1, n_rows, nullptr,
                              0, 0, nullptr, nullptr, XrArgs{});
    }
}

// the same kernel with chunks of 9 entries (fp64 values; what 27-wide rows run)
extern "C" int emul_sellc9(unsigned grid, const uint4 *meta, const uint32_t *cols, const int32_t *dcols,
                           const double *vals, const uint32_t *perm, const double *x, double *y,
                           uint32_t b0, uint32_t e0, uint32_t b1, uint32_t e1, uint32_t n_rows) {
  gridDim.x = grid;
  for (unsigned b = 0; b < grid; b++)
    for (unsigned t = 0; t < SPMV_THREADS; t++) {
      blockIdx.x = b, threadIdx.x = t;
      k_spmv_sellc<false, double, false, 9>(meta, cols, dcols, vals, perm, x, y, b0, e0, b1, e1, n_rows,
                                            nullptr, 0, 0, nullptr, nullptr, XrArgs{});
    }
  return 0;
}

extern "C" int emul_sellc(int f64, unsigned grid, const uint4 *meta, const uint32_t *cols,
                          const int32_t *dcols, const void *vals, const uint32_t *perm,
                          const double *x, double *y, uint32_t b0, uint32_t e0, uint32_t b1,
                          uint32_t e1, uint32_t n_rows) {
  if (f64)
    run_sellc<double>(grid, meta, cols, dcols, (const double *)vals, perm, x, y, b0, e0, b1, e1, n_rows);
  else
    run_sellc<float>(grid, meta, cols, dcols, (const float *)vals, perm, x, y, b0, e0, b1, e1, n_rows);
  return 0;
}

template <typename VT>
static void run_sell(unsigned grid, const uint32_t *sell_off, const uint32_t *cols, const VT *vals,
                     const uint32_t *perm, const double *x, double *y, uint32_t b0, uint32_t e0,
                     uint32_t b1, uint32_t e1, uint32_t n_rows) {
  gridDim.x = grid;
  for (unsigned b = 0; b < grid; b++)
    for (unsigned t = 0; t < SPMV_THREADS; t++) {
      blockIdx.x = b, threadIdx.x = t;
      k_spmv_sell<false, VT>(sell_off, cols, vals, perm, x, y, b0, e0, b1, e1, n_rows, nullptr, 0, 0,
                             nullptr, nullptr, XrArgs{});
    }
}

extern "C" int emul_sell(int f64, unsigned grid, const uint32_t *sell_off, const uint32_t *cols,
                         const void *vals, const uint32_t *perm, const double *x, double *y,
                         uint32_t b0, uint32_t e0, uint32_t b1, uint32_t e1, uint32_t n_rows) {
  if (f64)
    run_sell<double>(grid, sell_off, cols, (const double *)vals, perm, x, y, b0, e0, b1, e1, n_rows);
  else
    run_sell<float>(grid, sell_off, cols, (const float *)vals, perm, x, y, b0, e0, b1, e1, n_rows);
  return 0;
}
