// colblock_kernels.cuh -- the kernels that cut an operator into column ranges
// (included by convert.cu; in a header so that tests/ can run them on the host
// SIMT emulator, tests/pcg_emul.cpp).
#pragma once

// ---- column blocking (B200_MAT_COL_BLOCK) ------------------------------------------------
// Rows hold their columns in ascending order (every input path sorts them), so
// the entries of row i that fall into column range b are one contiguous piece.
// cnt is block-major: cnt[b * (n + 1) + i]; entry n of every block is 0 (scan tail).
__global__ void k_colblock_count(uint64_t n, const uint64_t *__restrict__ offs,
                                 const uint32_t *__restrict__ cols, uint64_t width, uint32_t nb,
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
    // first entry at or after e whose column is >= (b + 1) * width
    const uint64_t lim = (uint64_t)(b + 1) * width;
    uint64_t lo = e, hi = end;
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo) / 2;
      if ((uint64_t)cols[mid] < lim)
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

// *unsorted is set when some row does not hold its columns in ascending order
__global__ void k_rows_sorted(uint64_t n, const uint64_t *__restrict__ offs,
                              const uint32_t *__restrict__ cols, unsigned *unsorted) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  bool bad = false;
  for (uint64_t e = offs[i] + 1; e < offs[i + 1]; e++)
    bad |= cols[e] < cols[e - 1];
  if (bad)
    *unsorted = 1u;
}

