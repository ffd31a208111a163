// ACC (column-blocked matrices, convert.cu): y += A x instead of y = A x, so the
// column blocks of one operator can be multiplied one after another.
//
// sell_kernels.cuh -- the SELL-32 SpMV kernels (included by spmv.cu, which holds
// the launch logic and the description of the layout).  In a header of their own
// so that tests/ can also compile the kernel bodies for the host, with one-line
// stand-ins for the CUDA built-ins, and run them thread by thread against a CSR
// product without a GPU (tests/spmv_emul.cpp, tests/test_spmv_emul.py).
#pragma once

#define SPMV_THREADS 256
#define SPMV_WARPS (SPMV_THREADS / 32)

// One k-chunk of up to 8 entries of a slice: columns -> values -> gathers ->
// fma in row order.  `colf(j)` yields the column of entry k + j.
template <bool FULL, int NCH = 8, typename VT, typename ColF>
__device__ __forceinline__ double sell_chunk_n(const VT *vp, const double *__restrict__ x,
                                               uint32_t k, uint32_t rem, double sum, ColF colf) {
  // A partial chunk (rem < 8 entries) clamps the ENTRY INDEX of the surplus loads to the
  // chunk's last entry and leaves them unused; it must not select on the loaded VALUE
  // (j < rem ? load : 0): that select consumes every load where it stands, and the loads
  // of a 3-entry tail go out one after the other, a round trip each (round 2, ncu source
  // page of the bulk-copy-fed kernel, where the same pattern cost 8 us per slice).
  constexpr int N = FULL ? NCH : NCH - 1;
  uint32_t c[NCH];
  VT a[NCH];
  double xv[NCH];
#pragma unroll
  for (int j = 0; j < N; j++)
    c[j] = colf(FULL || (uint32_t)j < rem ? j : (int)rem - 1);
#pragma unroll
  for (int j = 0; j < N; j++)
    a[j] = ld_stream(vp + (size_t)(k + (FULL || (uint32_t)j < rem ? (uint32_t)j : rem - 1u)) * B2_SLICE);
#pragma unroll
  for (int j = 0; j < N; j++)
    xv[j] = __ldg(x + c[j]);
#pragma unroll
  for (int j = 0; j < N; j++)
    if (FULL || (uint32_t)j < rem)
      sum = fma((double)a[j], xv[j], sum);
  return sum;
}
template <bool FULL, int NCH = 8, typename ColF>
__device__ __forceinline__ double sell_chunk(const double *vp, const double *__restrict__ x,
                                             uint32_t k, uint32_t rem, double sum, ColF colf) {
  return sell_chunk_n<FULL, NCH>(vp, x, k, rem, sum, colf);
}

// The same for fp32-stored values (B200_MAT_VALUES_F32): a chunk is 16 entries
// so that a warp keeps as many value bytes in flight as with fp64 (16 x 128 B);
// the gathers and the fma chain run in two halves of 8, in row order, on the
// values widened to fp64 -- when the stored fp32 equals the original fp64 the
// result has the same bits as the fp64 kernel's.
template <bool FULL, int NCH = 16, typename ColF>
__device__ __forceinline__ double sell_chunk(const float *vp, const double *__restrict__ x,
                                             uint32_t k, uint32_t rem, double sum, ColF colf) {
  if constexpr (NCH != 16)  // chunks of 9 (27-wide rows): the plain chunk, values widened at the fma
    return sell_chunk_n<FULL, NCH>(vp, x, k, rem, sum, colf);
  constexpr int N = FULL ? 16 : 15;
  float a[16];
  // (surplus loads of a partial chunk: entry index clamped, value never selected -- see above)
#pragma unroll
  for (int j = 0; j < N; j++)
    a[j] = ld_stream(vp + (size_t)(k + (FULL || (uint32_t)j < rem ? (uint32_t)j : rem - 1u)) * B2_SLICE);
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t c[8];
    double xv[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
      c[j] = colf(FULL || (uint32_t)(h * 8 + j) < rem ? h * 8 + j : (int)rem - 1);
#pragma unroll
    for (int j = 0; j < 8; j++)
      xv[j] = __ldg(x + c[j]);
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (h * 8 + j < N && (FULL || h * 8 + j < rem))
        sum = fma((double)a[h * 8 + j], xv[j], sum);
  }
  return sum;
}

// entries per full chunk of the value type
template <typename VT> struct ChunkOf { static constexpr uint32_t n = 8; };
template <> struct ChunkOf<float> { static constexpr uint32_t n = 16; };

template <bool DOT, typename VT, bool ACC = false>
__global__ void __launch_bounds__(SPMV_THREADS, 4)
k_spmv_sell(const uint32_t *__restrict__ sell_off,
            const uint32_t *__restrict__ cols, const VT *__restrict__ vals,
            const uint32_t *__restrict__ perm, const double *__restrict__ x,
            double *__restrict__ y, uint32_t b0, uint32_t e0, uint32_t b1,
            uint32_t e1, uint32_t n_rows, double *partials, unsigned slot_base,
            unsigned total_slots, PcgState *st, double *dot_out, const XrArgs xr) {
  if (DOT && st->done)
    return;
  __shared__ double red[SPMV_WARPS];
  constexpr uint32_t CH = ChunkOf<VT>::n;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = e0 - b0, nv = n0 + (e1 - b1);
  const uint32_t stride = gridDim.x * SPMV_WARPS;
  double dot = 0.0;
  for (uint32_t v = blockIdx.x * SPMV_WARPS + warp; v < nv; v += stride) {
    const uint32_t s = v < n0 ? b0 + v : b1 + (v - n0);
    const uint32_t o = __ldg(sell_off + s);
    const uint32_t w = __ldg(sell_off + s + 1) - o;
    if (ACC && w == 0)  // nothing to add: leave y alone (no read-modify-write traffic)
      continue;
    const size_t base = (size_t)o * B2_SLICE + lane;
    const uint32_t *cp = cols + base;
    const VT *vp = vals + base;
    double sum = 0.0;
    uint32_t k = 0;
    for (; k + CH <= w; k += CH)
      sum = sell_chunk<true>(vp, x, k, CH, sum, [&](int j) {
        return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
      });
    if (k < w)
      sum = sell_chunk<false>(vp, x, k, w - k, sum, [&](int j) {
        return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
      });
    const uint32_t pos = s * B2_SLICE + lane;
    const uint32_t row = perm ? __ldg(perm + pos) : pos;
    if (row < n_rows) {
      y[row] = ACC ? y[row] + sum : sum;
      if (DOT)
        dot = fma(sum, __ldg(x + row), dot);
    }
  }
  if (DOT) {
    double b[1] = {block_sum<SPMV_WARPS>(dot, red)};
    grid_sum_finish<1, SPMV_WARPS>(b, partials, 0, slot_base + blockIdx.x,
                                   total_slots, &st->ticket[0], dot_out, red, xr);
  }
}

// ---- SELL slices taken several at a time (the column ranges of a column-blocked operator) ----
// A column range of a power-law operator holds one or two entries in most of its rows.  A warp
// that takes one slice per trip pays the same dependent round trips for those 32-64 entries --
// offsets, then columns / values / row ids, then the gathers and y -- as for a chunk of 256, and
// the pass is bound by that latency (round 2: DRAM at 50 % with 49 % of the warp slots busy,
// profiles/r02_powerlaw_colblock_ncu.txt).  Here a warp takes FOUR consecutive slices per trip and,
// when they are narrow, issues the loads of all of them together: 4 slices x <= 2 entries or
// 2 slices x <= 4 entries = 8 loads of each stream in flight per lane, the same as a full chunk.
// Wider slices run the chunked loop of k_spmv_sell.  Surplus loads have their ADDRESS clamped
// (never a select on the loaded value, see sell_chunk_n); each row still adds its entries left to
// right with fma, so a row's result has the bits k_spmv_sell gives it.
template <int U, int W, bool ACC>
__device__ __forceinline__ void sell_narrow(const uint32_t *o, const uint32_t *w, uint32_t s0, uint32_t ns,
                                            uint32_t safe_o, uint32_t safe_s, uint32_t lane,
                                            const uint32_t *__restrict__ cols, const double *__restrict__ vals,
                                            const uint32_t *__restrict__ perm, const double *__restrict__ x,
                                            double *__restrict__ y, uint32_t n_rows) {
  // safe_o / safe_s: offset and index of a live slice of the group -- where the loads of an
  // empty slice (or of one past the end of the list) point; nothing of them is used
  uint32_t row[U], c[U][W];
  double a[U][W], xv[U][W], yv[U];
  size_t e[U][W];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t pos = (s0 + u < ns ? s0 + u : safe_s) * B2_SLICE + lane;
    row[u] = perm ? __ldg(perm + pos) : pos;
#pragma unroll
    for (int k = 0; k < W; k++)
      e[u][k] = (size_t)(w[u] ? o[u] + ((uint32_t)k < w[u] ? (uint32_t)k : w[u] - 1u) : safe_o) * B2_SLICE + lane;
  }
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < W; k++)
      c[u][k] = ld_stream(cols + e[u][k]);
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < W; k++)
      a[u][k] = ld_stream(vals + e[u][k]);
  bool live[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    live[u] = s0 + u < ns && row[u] < n_rows && (!ACC || w[u] != 0);
    if (ACC)
      yv[u] = y[live[u] ? row[u] : 0u];
  }
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < W; k++)
      xv[u][k] = __ldg(x + c[u][k]);
  // a scheduling fence: left to itself ptxas interleaves the fma chain of the first slices with
  // the loads of the last ones, and half the loads wait for a round trip of the other half
  __syncwarp();
#pragma unroll
  for (int u = 0; u < U; u++) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < W; k++)
      if ((uint32_t)k < w[u])
        sum = fma(a[u][k], xv[u][k], sum);
    if (live[u])
      y[row[u]] = ACC ? yv[u] + sum : sum;  // (y = A x: the rows of an empty slice are zero)
  }
}

#define SELL_GRP 4
// CTAs per SM: 4 (64 registers, 8 bytes spilled in the accumulate instantiation) against 3 (80
// registers, none): 5.62 against 5.87 ms per SpMV of the 50 M-row power-law operator -- once
// the traffic was down (22 GB per SpMV) the pass was bound by its round trips at 24 warps per
// SM (ncu: 36 % of the warp slots, 30 long-scoreboard stalls per issue), and 32 warps pay more
// than the spill costs.  (With work units of 256 slices, at 28 GB, the two were equal.)
#ifndef SELL_GRP_MINB
#define SELL_GRP_MINB 4
#endif
// groups per unit of work: 8 = one group per warp per unit, 32 slices, 1 024 rows.  Measured on
// the 50 M-row power-law operator (64 MB ranges, windows of 32 768 rows): 128 groups 7.11 ms,
// 64 6.60, 32 6.41, 16 6.17, 8 6.19, 4 7.53 -- the finer the units, the narrower the front of
// rows all CTAs work on, and the y they update and the x they gather near the diagonal stay in
// cache; below 8 the warps of a CTA run out of work between two draws.
#ifndef SELL_UNIT
#define SELL_UNIT 8
#endif
template <bool ACC>
__global__ void __launch_bounds__(SPMV_THREADS, SELL_GRP_MINB)
k_spmv_sell_grp(const uint32_t *__restrict__ sell_off, const uint32_t *__restrict__ cols,
                const double *__restrict__ vals, const uint32_t *__restrict__ perm,
                const double *__restrict__ x, double *__restrict__ y, uint32_t ns, uint32_t n_rows,
                unsigned *work /* {next unit, CTAs done}, both 0 between launches */,
                uint32_t long_rows, const uint32_t *__restrict__ long_ids, const uint64_t *__restrict__ long_off,
                uint32_t vec_rows, const uint32_t *__restrict__ vec_ids, const uint64_t *__restrict__ vec_off,
                const uint32_t *__restrict__ vl_cols, const double *__restrict__ vl_vals) {
  // Work is handed out in units of SELL_UNIT consecutive groups, first come first served.  A
  // static grid-stride walk resonates with the length-sort windows (a window runs from its
  // widest rows down to its empty ones; the stride is a fixed number of slices, so a warp lands
  // on the same few window phases every time and some warps only ever see wide slices):
  // measured 2 x slower than the same layout sorted as a whole.  No sum is
  // formed across rows here, so who multiplies which slice does not change any bit of y.
  //
  // The rows of the row-major bins (longer than 256 entries in this range) are units of the same
  // list, longest first: a CTA-per-row unit per "long" row, then units of one warp-per-row row
  // per warp, then the slices -- one launch per range instead of three, and the long rows do
  // not wait for a tail of their own (same loops as k_spmv_long / k_spmv_vec, same bits).
  __shared__ uint32_t unit_s;
  __shared__ double red[SPMV_WARPS];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ngroups = (ns + SELL_GRP - 1) / SELL_GRP;
  const uint32_t vec_units = (vec_rows + SPMV_WARPS - 1) / SPMV_WARPS;
  const uint32_t nunits = (ngroups + SELL_UNIT - 1) / SELL_UNIT + long_rows + vec_units;
  for (;;) {
    if (threadIdx.x == 0)
      unit_s = atomicAdd(&work[0], 1u);
    __syncthreads();
    uint32_t unit = unit_s;
    __syncthreads();
    if (unit >= nunits)
      break;
    if (unit < long_rows) {  // one CTA, one row
      const uint64_t s = __ldg(long_off + unit), e = __ldg(long_off + unit + 1);
      double sum = 0.0;
      for (uint64_t k = s + threadIdx.x * 4; k < e; k += SPMV_THREADS * 4) {
        const uint4 c = __ldcs(reinterpret_cast<const uint4 *>(vl_cols + k));
        const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(vl_vals + k));
        const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(vl_vals + k + 2));
        const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
        sum = fma(v01.x, x0, sum);
        sum = fma(v01.y, x1, sum);
        sum = fma(v23.x, x2, sum);
        sum = fma(v23.y, x3, sum);
      }
      sum = block_sum<SPMV_WARPS>(sum, red);
      if (threadIdx.x == 0) {
        const uint32_t row = __ldg(long_ids + unit);
        y[row] = ACC ? y[row] + sum : sum;
      }
      continue;
    }
    unit -= long_rows;
    if (unit < vec_units) {  // one warp, one row
      const uint32_t r = unit * SPMV_WARPS + warp;
      if (r < vec_rows) {
        const uint64_t s = __ldg(vec_off + r), e = __ldg(vec_off + r + 1);
        double sum = 0.0;
        for (uint64_t k = s + lane * 4; k < e; k += 128) {
          const uint4 c = __ldcs(reinterpret_cast<const uint4 *>(vl_cols + k));
          const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(vl_vals + k));
          const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(vl_vals + k + 2));
          const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
          sum = fma(v01.x, x0, sum);
          sum = fma(v01.y, x1, sum);
          sum = fma(v23.x, x2, sum);
          sum = fma(v23.y, x3, sum);
        }
        sum = warp_sum(sum);
        if (lane == 0) {
          const uint32_t row = __ldg(vec_ids + r);
          y[row] = ACC ? y[row] + sum : sum;
        }
      }
      continue;
    }
    unit -= vec_units;
    const uint32_t g_end = (unit + 1) * SELL_UNIT < ngroups ? (unit + 1) * SELL_UNIT : ngroups;
  for (uint32_t g = unit * SELL_UNIT + warp; g < g_end; g += SPMV_WARPS) {
    const uint32_t s0 = g * SELL_GRP;
    // offsets of the group's slices: one load, broadcast (slices past the end are empty)
    const uint32_t mine = __ldg(sell_off + (s0 + lane < ns ? s0 + lane : ns));
    uint32_t o[SELL_GRP + 1], w[SELL_GRP];
#pragma unroll
    for (int u = 0; u <= SELL_GRP; u++)
      o[u] = __shfl_sync(0xffffffffu, mine, u);
    uint32_t wmax = 0, safe_o = 0, safe_s = s0;
#pragma unroll
    for (int u = SELL_GRP - 1; u >= 0; u--) {
      w[u] = o[u + 1] - o[u];
      wmax = w[u] > wmax ? w[u] : wmax;
      if (w[u])
        safe_o = o[u], safe_s = s0 + u;
    }
    if (ACC && wmax == 0)  // nothing to add: leave y alone (no read-modify-write traffic)
      continue;
    if (wmax != 0 && wmax <= 2) {
      sell_narrow<SELL_GRP, 2, ACC>(o, w, s0, ns, safe_o, safe_s, lane, cols, vals, perm, x, y, n_rows);
      continue;
    }
    if (wmax != 0 && wmax <= 4) {
      sell_narrow<2, 4, ACC>(o, w, s0, ns, safe_o, safe_s, lane, cols, vals, perm, x, y, n_rows);
      sell_narrow<2, 4, ACC>(o + 2, w + 2, s0 + 2, ns, safe_o, safe_s, lane, cols, vals, perm, x, y, n_rows);
      continue;
    }
#pragma unroll 1
    for (int u = 0; u < SELL_GRP; u++) {
      if (s0 + u >= ns)
        break;
      const uint32_t ou = __shfl_sync(0xffffffffu, mine, u), wu = __shfl_sync(0xffffffffu, mine, u + 1) - ou;
      if (ACC && wu == 0)
        continue;
      const size_t base = (size_t)ou * B2_SLICE + lane;
      const uint32_t *cp = cols + base;
      const double *vp = vals + base;
      const uint32_t pos = (s0 + u) * B2_SLICE + lane;
      const uint32_t row = perm ? __ldg(perm + pos) : pos;
      double sum = 0.0;
      uint32_t k = 0;
      for (; k + 8 <= wu; k += 8)
        sum = sell_chunk<true>(vp, x, k, 8, sum, [&](int j) {
          return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
        });
      if (k < wu)
        sum = sell_chunk<false>(vp, x, k, wu - k, sum, [&](int j) {
          return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
        });
      if (row < n_rows)
        y[row] = ACC ? y[row] + sum : sum;
    }
  }
  }
  // the last CTA out rearms the counters (every CTA has drawn its final, empty unit by then)
  if (threadIdx.x == 0 && atomicAdd(&work[1], 1u) == gridDim.x - 1) {
    work[0] = 0u;
    work[1] = 0u;
  }
}

// Index-compressed SELL (convert.cu, step 6).  A slice whose 32 rows all have
// the slice's width and whose k-th column is `row + d_k` with the same d_k in
// every lane (the interior of any stencil or banded operator, in any numbering
// that keeps neighbouring rows together) stores its w deltas once -- 4 w bytes
// instead of 128 w -- and its gathers become one coalesced warp access each.
// Other slices keep explicit columns.  meta[s] = {o, w | uniform << 31, c, 0}:
// o as in sell_off; c = offset into dcols (entries) for a uniform slice, into
// cols (units of 32 entries) otherwise.  Values and the order of the additions
// are untouched, so the result is bit-identical to the uncompressed kernel.
//
// Measured alternatives that lost to this plain loop on 27-point 512^3 (5.99 ms
// in PCG): deltas prefetched one slice ahead and broadcast by shuffle (6.38),
// next-chunk values prefetched into registers (8.0-8.4), a bulk-copy
// (cp.async.bulk + mbarrier) ring of value tiles in shared memory (10.3), and
// balanced 9+9+9 chunks instead of 8+8+8+3 (6.25).
//
// Compiled for 5 CTAs per SM (48 registers, 40 warps): the uniform path needs no
// column registers, and the extra warps are worth 6 % (27-point 512^3: 6.07 ->
// 5.73 ms in PCG; 6 CTAs/SM spills and loses again).  The fp32-value
// instantiation holds 16 values per chunk and is compiled for 4 CTAs per SM
// (64 registers; at 48 it spills 112 bytes in the loop).
//
// CHD: entries per chunk with fp64 values, 8 or 9.  A 27-wide row is 8 + 8 + 8 + 3 with
// 8 (four dependent rounds of loads per slice) and 9 + 9 + 9 with 9 (three).
#define SELLC_MINB 5
template <bool DOT, typename VT, bool ACC = false, int CHD = 8>
__global__ void __launch_bounds__(SPMV_THREADS, (sizeof(VT) == 8 || CHD == 9) ? SELLC_MINB : 4)
k_spmv_sellc(const uint4 *__restrict__ meta, const uint32_t *__restrict__ cols,
             const int32_t *__restrict__ dcols, const VT *__restrict__ vals,
             const uint32_t *__restrict__ perm, const double *__restrict__ x,
             double *__restrict__ y, uint32_t b0, uint32_t e0, uint32_t b1,
             uint32_t e1, uint32_t n_rows, double *partials, unsigned slot_base,
             unsigned total_slots, PcgState *st, double *dot_out, const XrArgs xr) {
  if (DOT && st->done)
    return;
  __shared__ double red[SPMV_WARPS];
  constexpr uint32_t CH = CHD == 9 ? 9u : ChunkOf<VT>::n;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = e0 - b0, nv = n0 + (e1 - b1);
  const uint32_t stride = gridDim.x * SPMV_WARPS;
  double dot = 0.0;
  for (uint32_t v = blockIdx.x * SPMV_WARPS + warp; v < nv; v += stride) {
    const uint32_t s = v < n0 ? b0 + v : b1 + (v - n0);
    const uint4 m = __ldg(meta + s);
    const uint32_t o = m.x, w = m.y & 0x7fffffffu;
    if (ACC && w == 0)  // nothing to add: leave y alone
      continue;
    const VT *vp = vals + (size_t)o * B2_SLICE + lane;
    const uint32_t pos = s * B2_SLICE + lane;
    const uint32_t row = perm ? __ldg(perm + pos) : pos;
    double sum = 0.0;
    uint32_t k = 0;
    if (m.y >> 31) {
      const int32_t *dp = dcols + m.z;
      for (; k + CH <= w; k += CH)
        sum = sell_chunk<true, (int)CH>(vp, x, k, CH, sum, [&](int j) {
          return row + (uint32_t)__ldg(dp + k + j);
        });
      if (k < w)
        sum = sell_chunk<false, (int)CH>(vp, x, k, w - k, sum, [&](int j) {
          return row + (uint32_t)__ldg(dp + k + j);
        });
    } else {
      const uint32_t *cp = cols + (size_t)m.z * B2_SLICE + lane;
      for (; k + CH <= w; k += CH)
        sum = sell_chunk<true, (int)CH>(vp, x, k, CH, sum, [&](int j) {
          return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
        });
      if (k < w)
        sum = sell_chunk<false, (int)CH>(vp, x, k, w - k, sum, [&](int j) {
          return ld_stream(cp + (size_t)(k + j) * B2_SLICE);
        });
    }
    if (row < n_rows) {
      y[row] = ACC ? y[row] + sum : sum;
      if (DOT)
        dot = fma(sum, __ldg(x + row), dot);
    }
  }
  if (DOT) {
    double b[1] = {block_sum<SPMV_WARPS>(dot, red)};
    grid_sum_finish<1, SPMV_WARPS>(b, partials, 0, slot_base + blockIdx.x,
                                   total_slots, &st->ticket[0], dot_out, red, xr);
  }
}


// ---- the row-major bins: one warp per row, one CTA per row ----------------------------
template <bool DOT, bool ACC = false>
__global__ void __launch_bounds__(SPMV_THREADS, 4)
k_spmv_vec(uint32_t nrows, const uint32_t *__restrict__ ids,
           const uint64_t *__restrict__ off, const uint32_t *__restrict__ cols,
           const double *__restrict__ vals, const double *__restrict__ x,
           double *__restrict__ y, double *partials, unsigned slot_base,
           unsigned total_slots, PcgState *st, double *dot_out, const XrArgs xr) {
  if (DOT && st->done)
    return;
  __shared__ double red[SPMV_WARPS];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stride = gridDim.x * SPMV_WARPS;
  double dot = 0.0;
  for (uint32_t r = blockIdx.x * SPMV_WARPS + warp; r < nrows; r += stride) {
    const uint64_t s = __ldg(off + r), e = __ldg(off + r + 1);
    double sum = 0.0;
    for (uint64_t k = s + lane * 4; k < e; k += 128) {
      const uint4 c = __ldcs(reinterpret_cast<const uint4 *>(cols + k));
      const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(vals + k));
      const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(vals + k + 2));
      const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y),
                   x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
      sum = fma(v01.x, x0, sum);
      sum = fma(v01.y, x1, sum);
      sum = fma(v23.x, x2, sum);
      sum = fma(v23.y, x3, sum);
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      const uint32_t row = __ldg(ids + r);
      y[row] = ACC ? y[row] + sum : sum;
      if (DOT)
        dot = fma(sum, __ldg(x + row), dot);
    }
  }
  if (DOT) {
    double b[1] = {block_sum<SPMV_WARPS>(dot, red)};
    grid_sum_finish<1, SPMV_WARPS>(b, partials, 0, slot_base + blockIdx.x,
                                   total_slots, &st->ticket[0], dot_out, red, xr);
  }
}

template <bool DOT, bool ACC = false>
__global__ void __launch_bounds__(SPMV_THREADS, 4)
k_spmv_long(uint32_t nrows, const uint32_t *__restrict__ ids,
            const uint64_t *__restrict__ off, const uint32_t *__restrict__ cols,
            const double *__restrict__ vals, const double *__restrict__ x,
            double *__restrict__ y, double *partials, unsigned slot_base,
            unsigned total_slots, PcgState *st, double *dot_out, const XrArgs xr) {
  if (DOT && st->done)
    return;
  __shared__ double red[SPMV_WARPS];
  double dot = 0.0;
  for (uint32_t r = blockIdx.x; r < nrows; r += gridDim.x) {
    const uint64_t s = __ldg(off + r), e = __ldg(off + r + 1);
    double sum = 0.0;
    for (uint64_t k = s + threadIdx.x * 4; k < e; k += SPMV_THREADS * 4) {
      const uint4 c = __ldcs(reinterpret_cast<const uint4 *>(cols + k));
      const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(vals + k));
      const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(vals + k + 2));
      const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y),
                   x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
      sum = fma(v01.x, x0, sum);
      sum = fma(v01.y, x1, sum);
      sum = fma(v23.x, x2, sum);
      sum = fma(v23.y, x3, sum);
    }
    sum = block_sum<SPMV_WARPS>(sum, red);
    if (threadIdx.x == 0) {
      const uint32_t row = __ldg(ids + r);
      y[row] = ACC ? y[row] + sum : sum;
      if (DOT)
        dot = fma(sum, __ldg(x + row), dot);
    }
  }
  if (DOT) {
    // only thread 0 carries a value; block_sum keeps the protocol uniform
    double b[1] = {block_sum<SPMV_WARPS>(dot, red)};
    grid_sum_finish<1, SPMV_WARPS>(b, partials, 0, slot_base + blockIdx.x,
                                   total_slots, &st->ticket[0], dot_out, red, xr);
  }
}

