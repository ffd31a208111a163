// colblock_kernels.cuh -- the kernels that cut an operator into column ranges
// (included by convert.cu; in a header so that tests/ can run them on the host
// SIMT emulator, tests/pcg_emul.cpp).
#pragma once

// ---- column blocking (B200_MAT_COL_BLOCK) ------------------------------------------------
// Rows hold their columns in ascending GLOBAL order (every input path sorts them), so
// the entries of row i that fall into column range b are one contiguous piece.  On one
// rank a column id is its own position in that order.  On several ranks the columns have
// been renumbered (dist.cu): owned columns -> [0, n_own), remote ones -> n_own + slot with
// the slots in ascending global order, of which the first n_low lie below the rank's own
// rows.  col_ord gives the position in global order back: [slots below | owned | slots
// above]; the ranges are cut in that order (cuts[b] <= ord < cuts[b + 1]), never across
// one of the two seams, so every range is one contiguous piece of the extended x vector.
// cnt is block-major: cnt[b * (n + 1) + i]; entry n of every block is 0 (scan tail).
__host__ __device__ __forceinline__ uint64_t col_ord(uint32_t c, uint64_t n_own, uint64_t n_low) {
  return c < n_own ? n_low + c : (c - n_own < n_low ? c - n_own : (uint64_t)c);
}

__global__ void k_colblock_count(uint64_t n, const uint64_t *__restrict__ offs,
                                 const uint32_t *__restrict__ cols, const uint64_t *__restrict__ cuts,
                                 uint32_t nb, uint64_t n_own, uint64_t n_low,
                                 uint64_t *__restrict__ cnt) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > n)
    return;
  if (i == n) {
    for (uint32_t b = 0; b < nb; b++)
      cnt[(uint64_t)b * (n + 1) + n] = 0;
    return;
  }
  uint64_t e = offs[i];
  const uint64_t end = offs[i + 1];
  for (uint32_t b = 0; b < nb; b++) {
    // first entry at or after e whose column lies at or beyond the end of range b
    const uint64_t lim = cuts[b + 1];
    uint64_t lo = e, hi = end;
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo) / 2;
      if (col_ord(cols[mid], n_own, n_low) < lim)
        lo = mid + 1;
      else
        hi = mid;
    }
    cnt[(uint64_t)b * (n + 1) + i] = lo - e;
    e = lo;
  }
}

// one warp per row: the pieces of the row go to their blocks, entry order kept
__global__ void k_colblock_fill(uint64_t n, const uint64_t *__restrict__ offs,
                                const uint32_t *__restrict__ cols,
                                const double *__restrict__ vals, uint32_t nb,
                                const uint64_t *__restrict__ boffs /* scanned, block-major */,
                                uint32_t *const *__restrict__ ocols,
                                double *const *__restrict__ ovals) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i = warp; i < n; i += nwarps) {
    uint64_t src = offs[i];
    for (uint32_t b = 0; b < nb; b++) {
      const uint64_t dst = boffs[(uint64_t)b * (n + 1) + i];
      const uint64_t len = boffs[(uint64_t)b * (n + 1) + i + 1] - dst;
      for (uint64_t k = lane; k < len; k += 32)
        ocols[b][dst + k] = cols[src + k], ovals[b][dst + k] = vals[src + k];
      src += len;
    }
  }
}

// *unsorted is set when some row does not hold its columns in ascending (global) order
__global__ void k_rows_sorted(uint64_t n, const uint64_t *__restrict__ offs,
                              const uint32_t *__restrict__ cols, uint64_t n_own, uint64_t n_low,
                              unsigned *unsorted) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  bool bad = false;
  for (uint64_t e = offs[i] + 1; e < offs[i + 1]; e++)
    bad |= col_ord(cols[e], n_own, n_low) < col_ord(cols[e - 1], n_own, n_low);
  if (bad)
    *unsorted = 1u;
}

