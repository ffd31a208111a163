// sellc32p.cuh -- the software-pipelined fp32-value SELL kernel (included by
// spmv.cu).  Kept in a header of its own so that tests/ can also compile the
// kernel body for the host, with one-line shims for the CUDA built-ins, and run
// it thread by thread against a CSR product (tests/spmv_emul.cpp): the
// indexing and the order of the additions are checked without a GPU.
#pragma once

// Software-pipelined variant of k_spmv_sellc for fp32-stored values and slices
// no wider than WMAX (B200_SPMV_PIPE=1; NOT the default -- written after the
// round-1 measurement below and not yet timed on hardware).
//
// Why: with fp32 values the plain loop moves 42 % fewer bytes in 5 % MORE time
// (27-point 192^3: 0.292 -> 0.306 ms), i.e. the kernel is bound by one DRAM
// round trip per chunk per warp, not by bytes.  fp32 values are small enough to
// hold TWO whole slices in registers (2 x 27 floats): the values of the warp's
// next slice are requested before the gathers of the current one start, so a
// warp has a full slice of the value stream (w x 128 B) in flight all the time
// and pays no DRAM latency on its critical path.  Same fma chain per row, so
// the same bits.
//
// The same idea carries to fp64 values when the CTA is small enough to give a
// thread the registers (2 x 27 doubles): THREADS = 128, three CTAs per SM -- 12
// warps, each with 6.9 KB of values in flight, which is more than the 65 KB per
// SM that 6.4 TB/s x 1.5 us asks for (B200_SPMV_PIPE=2; same caveat).
template <typename VT, int WMAX> struct PipeCfg {
  static constexpr int threads = sizeof(VT) == 8 && WMAX > 16 ? 128 : 256;
  static constexpr int min_blocks =
      sizeof(VT) == 8 ? (WMAX > 16 ? 3 : (WMAX > 8 ? 2 : 3)) : (WMAX > 16 ? 2 : (WMAX > 8 ? 3 : 4));
};

template <bool DOT, int WMAX, typename VT>
__global__ void __launch_bounds__(PipeCfg<VT, WMAX>::threads, PipeCfg<VT, WMAX>::min_blocks)
k_spmv_sellc32p(const uint4 *__restrict__ meta, const uint32_t *__restrict__ cols,
                const int32_t *__restrict__ dcols, const VT *__restrict__ vals,
                const uint32_t *__restrict__ perm, const double *__restrict__ x,
                double *__restrict__ y, uint32_t b0, uint32_t e0, uint32_t b1,
                uint32_t e1, uint32_t n_rows, double *partials, unsigned slot_base,
                unsigned total_slots, PcgState *st, double *dot_out, const XrArgs xr) {
  if constexpr (DOT) {
    if (st->done)
      return;
  }
  constexpr int WARPS = PipeCfg<VT, WMAX>::threads / 32;
  __shared__ double red[WARPS];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = e0 - b0, nv = n0 + (e1 - b1);
  const uint32_t stride = gridDim.x * WARPS;
  double dot = 0.0;
  uint32_t v = blockIdx.x * WARPS + warp;
  VT a[WMAX];
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  if (v < nv) {
    m = __ldg(meta + (v < n0 ? b0 + v : b1 + (v - n0)));
    const VT *vp = vals + (size_t)m.x * B2_SLICE + lane;
    const uint32_t w = m.y & 0x7fffffffu;
#pragma unroll
    for (int k = 0; k < WMAX; k++)
      a[k] = k < w ? ld_stream(vp + (size_t)k * B2_SLICE) : VT(0);
  }
  while (v < nv) {
    // ---- request the next slice's values -------------------------------------------
    const uint32_t vn = v + stride;
    VT an[WMAX];
    uint4 mn = m;
    if (vn < nv) {
      mn = __ldg(meta + (vn < n0 ? b0 + vn : b1 + (vn - n0)));
      const VT *vp = vals + (size_t)mn.x * B2_SLICE + lane;
      const uint32_t wn = mn.y & 0x7fffffffu;
#pragma unroll
      for (int k = 0; k < WMAX; k++)
        an[k] = k < wn ? ld_stream(vp + (size_t)k * B2_SLICE) : VT(0);
    } else {
#pragma unroll
      for (int k = 0; k < WMAX; k++)
        an[k] = VT(0);
    }
    // ---- the current slice: gathers and the fma chain, 8 entries at a time ------------
    const uint32_t s = v < n0 ? b0 + v : b1 + (v - n0);
    const uint32_t w = m.y & 0x7fffffffu;
    const uint32_t pos = s * B2_SLICE + lane;
    const uint32_t row = perm ? __ldg(perm + pos) : pos;
    const bool uniform = m.y >> 31;
    const int32_t *dp = dcols + (uniform ? m.z : 0u);
    const uint32_t *cp = cols + (uniform ? 0 : (size_t)m.z * B2_SLICE) + lane;
    double sum = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < WMAX; k0 += 8) {
      if (k0 < w) {
        uint32_t c[8];
        double xv[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
          c[j] = (k0 + j < WMAX && k0 + j < w)
                     ? (uniform ? row + (uint32_t)__ldg(dp + k0 + j)
                                : ld_stream(cp + (size_t)(k0 + j) * B2_SLICE))
                     : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++)
          xv[j] = (k0 + j < WMAX && k0 + j < w) ? __ldg(x + c[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (k0 + j < WMAX && k0 + j < w)
            sum = fma((double)a[k0 + j < WMAX ? k0 + j : 0], xv[j], sum);
      }
    }
    if (row < n_rows) {
      y[row] = sum;
      if constexpr (DOT)
        dot = fma(sum, __ldg(x + row), dot);
    }
    v = vn, m = mn;
#pragma unroll
    for (int k = 0; k < WMAX; k++)
      a[k] = an[k];
  }
  if constexpr (DOT) {
    double b[1] = {block_sum<WARPS>(dot, red)};
    grid_sum_finish<1, WARPS>(b, partials, 0, slot_base + blockIdx.x,
                                   total_slots, &st->ticket[0], dot_out, red, xr);
  }
}

