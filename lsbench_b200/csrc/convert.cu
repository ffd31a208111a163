// convert.cu -- north_star piece (1): host CSR (src/lsbench-impl.h:22-26, as
// built by src/lsbench-csr.c:79-86) -> device layout.
//
// Stands where a reference backend's csr_init stands (src/cusparse.c:47-125):
// index conversion, base handling (:59,:63) and H2D, but instead of handing
// the arrays to a library it rewrites them for the SpMV kernels:
//
//   SELL-32 bin   rows with len <= B2_SELL_MAX.  32 rows form a slice; a slice is
//                 stored column-major (entry k of lane l at off*32 + 32k + l),
//                 so every warp load of cols / vals is one aligned 128 B /
//                 256 B segment and each lane accumulates its own row left to
//                 right.  Slices are padded to their longest row; a sort of
//                 the rows by length inside windows of B2_SELL_SIGMA rows keeps
//                 that padding small on irregular matrices and is skipped
//                 (identity permutation) when the natural order already pads
//                 less than 3 %.
//   vector bin    B2_SELL_MAX < len < B2_LONG_MIN: row-major, each row padded to a
//                 multiple of 4 entries and 32 B aligned for 128-bit loads,
//                 one warp per row.
//   long bin      len >= B2_LONG_MIN: same storage, one CTA per row.
//
// Index compression (on unless B200_MAT_NO_COMPRESS): a SELL slice whose 32
// rows all have the slice width and the same column-minus-row pattern stores
// its w deltas once instead of 32 w columns (k_slice_uniform / k_compact_cols).
// Lossless -- b200_mat_export gives back the same CSR bit for bit -- and the
// values and summation order are untouched.
//
// With B200_MAT_SYM_UPPER the operator is first replaced by the one CHOLMOD
// factorises (src/cholmod-impl.h:5-21): entries with col >= row, mirrored.
#include "common.cuh"
#include <cub/cub.cuh>
#include <vector>

#define B2_SELL_MAX 256u
#define B2_LONG_MIN 8192u
#define B2_SELL_SIGMA 1024u
#define T256 256

static inline unsigned nblk(uint64_t n, unsigned t = T256) {
  return (unsigned)((n + t - 1) / t);
}

// ---------------------------------------------------------------------------
// scans (CUB is used for the set-up prefix sums only, never on the solve path)
template <typename T>
static int exclusive_scan(cudaStream_t s, const T *in, T *out, uint64_t n) {
  void *tmp = nullptr;
  size_t bytes = 0;
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
  CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, s));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaFree(tmp));
  return B200_OK;
}

int plain_free(PlainCsr *A) {
  if (A->offs) cudaFree(A->offs);
  if (A->cols) cudaFree(A->cols);
  if (A->vals) cudaFree(A->vals);
  *A = PlainCsr();
  return B200_OK;
}

// ---------------------------------------------------------------------------
// upload
__global__ void k_widen_offs(const uint32_t *in, uint64_t *out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[i];
}
__global__ void k_rebase_cols(uint32_t *cols, uint64_t nnz, uint32_t base) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < nnz)
    cols[i] -= base;  // src/cholmod-impl.h:16, src/amgx.c:41
}

static int upload_csr(b200_ctx *c, uint32_t nrows, uint32_t base,
                      const uint32_t *offs, const uint32_t *cols,
                      const double *vals, PlainCsr *A) {
  uint64_t nnz = offs[nrows];  // src/cusparse.c:55
  A->n = nrows, A->nnz = nnz;
  uint32_t *tmp = nullptr;
  CU_TRY(cudaMalloc(&tmp, (nrows + 1ull) * 4));
  CU_TRY(cudaMalloc(&A->offs, (nrows + 1ull) * 8));
  CU_TRY(cudaMalloc(&A->cols, (nnz ? nnz : 1) * 4));
  CU_TRY(cudaMalloc(&A->vals, (nnz ? nnz : 1) * 8));
  cudaStream_t s = c->stream;
  CU_TRY(cudaMemcpyAsync(tmp, offs, (nrows + 1ull) * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(A->cols, cols, nnz * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(A->vals, vals, nnz * 8, cudaMemcpyHostToDevice, s));
  k_widen_offs<<<nblk(nrows + 1ull), T256, 0, s>>>(tmp, A->offs, nrows + 1ull);
  if (base && nnz)
    k_rebase_cols<<<nblk(nnz), T256, 0, s>>>(A->cols, nnz, base);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaFree(tmp));
  return B200_OK;
}

// ---------------------------------------------------------------------------
// upper-triangle mirror (src/cholmod-impl.h:11-17 + triplet_to_sparse :21)
__global__ void k_sym_count(uint64_t n, const uint64_t *offs,
                            const uint32_t *cols, uint32_t *n_upper,
                            uint32_t *n_mirror) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  uint32_t u = 0;
  for (uint64_t k = offs[i]; k < offs[i + 1]; k++) {
    uint64_t j = cols[k];
    if (j >= i) {
      u++;
      if (j > i && j < n)
        atomicAdd(&n_mirror[j], 1u);  // integer count: order-independent
    }
  }
  n_upper[i] = u;
}

__global__ void k_add_u32_to_u64(const uint32_t *a, const uint32_t *b,
                                 uint64_t *out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = (uint64_t)a[i] + b[i];
  if (i == n)
    out[i] = 0;
}

__global__ void k_sym_fill(uint64_t n, const uint64_t *offs,
                           const uint32_t *cols, const double *vals,
                           const uint64_t *noffs, const uint32_t *n_mirror,
                           uint32_t *cursor, uint32_t *ncols, double *nvals) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  uint64_t up = noffs[i] + n_mirror[i];
  for (uint64_t k = offs[i]; k < offs[i + 1]; k++) {
    uint64_t j = cols[k];
    if (j < i)
      continue;  // lower-triangle file values are never used
    double v = vals[k];
    ncols[up] = (uint32_t)j, nvals[up] = v, up++;
    if (j > i && j < n) {
      uint64_t slot = noffs[j] + atomicAdd(&cursor[j], 1u);
      ncols[slot] = (uint32_t)i, nvals[slot] = v;
    }
  }
}

// The mirrored part of each row was filled in arrival order; put it in
// ascending column order so the result does not depend on scheduling.
__global__ void k_sym_sort(uint64_t n, const uint64_t *noffs,
                           const uint32_t *n_mirror, uint32_t *ncols,
                           double *nvals) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  uint64_t s = noffs[i];
  uint32_t m = n_mirror[i];
  for (uint32_t a = 1; a < m; a++) {
    uint32_t c = ncols[s + a];
    double v = nvals[s + a];
    uint32_t b = a;
    while (b > 0 && ncols[s + b - 1] > c) {
      ncols[s + b] = ncols[s + b - 1], nvals[s + b] = nvals[s + b - 1];
      b--;
    }
    ncols[s + b] = c, nvals[s + b] = v;
  }
}

__global__ void k_cols_differ(const uint32_t *a, const uint32_t *b, uint64_t n,
                              uint32_t *flag) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && a[i] != b[i])
    *flag = 1;
}

static int sym_upper(b200_ctx *c, PlainCsr *A, uint32_t *pattern_symmetric) {
  cudaStream_t s = c->stream;
  uint64_t n = A->n;
  uint32_t *n_upper, *n_mirror, *cursor, *flag;
  uint64_t *len64, *noffs;
  CU_TRY(cudaMalloc(&n_upper, (n + 1) * 4));
  CU_TRY(cudaMalloc(&n_mirror, (n + 1) * 4));
  CU_TRY(cudaMalloc(&cursor, (n + 1) * 4));
  CU_TRY(cudaMalloc(&flag, 4));
  CU_TRY(cudaMalloc(&len64, (n + 1) * 8));
  CU_TRY(cudaMalloc(&noffs, (n + 1) * 8));
  CU_TRY(cudaMemsetAsync(n_mirror, 0, (n + 1) * 4, s));
  CU_TRY(cudaMemsetAsync(cursor, 0, (n + 1) * 4, s));
  CU_TRY(cudaMemsetAsync(flag, 0, 4, s));
  k_sym_count<<<nblk(n), T256, 0, s>>>(n, A->offs, A->cols, n_upper, n_mirror);
  k_add_u32_to_u64<<<nblk(n + 1), T256, 0, s>>>(n_upper, n_mirror, len64, n);
  B_TRY(exclusive_scan<uint64_t>(s, len64, noffs, n + 1));
  uint64_t nnz = 0;
  CU_TRY(cudaMemcpy(&nnz, noffs + n, 8, cudaMemcpyDeviceToHost));
  uint32_t *ncols;
  double *nvals;
  CU_TRY(cudaMalloc(&ncols, (nnz ? nnz : 1) * 4));
  CU_TRY(cudaMalloc(&nvals, (nnz ? nnz : 1) * 8));
  k_sym_fill<<<nblk(n), T256, 0, s>>>(n, A->offs, A->cols, A->vals, noffs,
                                      n_mirror, cursor, ncols, nvals);
  k_sym_sort<<<nblk(n), T256, 0, s>>>(n, noffs, n_mirror, ncols, nvals);
  uint32_t differ = 1;
  if (nnz == A->nnz) {
    k_cols_differ<<<nblk(nnz ? nnz : 1), T256, 0, s>>>(A->cols, ncols, nnz, flag);
    CU_TRY(cudaMemcpyAsync(&differ, flag, 4, cudaMemcpyDeviceToHost, s));
  }
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  *pattern_symmetric = !differ;
  cudaFree(A->offs), cudaFree(A->cols), cudaFree(A->vals);
  A->offs = noffs, A->cols = ncols, A->vals = nvals, A->nnz = nnz;
  cudaFree(n_upper), cudaFree(n_mirror), cudaFree(cursor), cudaFree(flag);
  cudaFree(len64);
  return B200_OK;
}

// ---------------------------------------------------------------------------
// layout planning
__global__ void k_row_len(uint64_t n, const uint64_t *offs, uint32_t *len,
                          unsigned long long *hist, unsigned long long *maxlen) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  uint64_t l = offs[i + 1] - offs[i];
  len[i] = (uint32_t)l;
  int b = l <= 1 ? 0 : 64 - __clzll((unsigned long long)(l - 1));
  if (b >= B200_HIST_BINS)
    b = B200_HIST_BINS - 1;
  atomicAdd(&hist[b], 1ull);
  atomicMax(maxlen, (unsigned long long)l);
}

// bin: 0 SELL, 1 vector, 2 long
__device__ __forceinline__ int bin_of(uint32_t len, uint32_t sell_max,
                                      uint32_t long_min) {
  return len <= sell_max ? 0 : (len < long_min ? 1 : 2);
}

__global__ void k_bin_flags(uint64_t n, const uint32_t *len, uint32_t sell_max,
                            uint32_t long_min, uint32_t *f0, uint32_t *f1,
                            uint32_t *f2) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > n)
    return;
  int b = i < n ? bin_of(len[i], sell_max, long_min) : -1;
  f0[i] = b == 0, f1[i] = b == 1, f2[i] = b == 2;
}

__global__ void k_scatter_ids(uint64_t n, const uint32_t *len,
                              uint32_t sell_max, uint32_t long_min,
                              const uint32_t *p0, const uint32_t *p1,
                              const uint32_t *p2, uint32_t *l0, uint32_t *l1,
                              uint32_t *l2) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int b = bin_of(len[i], sell_max, long_min);
  if (b == 0)
    l0[p0[i]] = (uint32_t)i;
  else if (b == 1)
    l1[p1[i]] = (uint32_t)i;
  else
    l2[p2[i]] = (uint32_t)i;
}

__global__ void k_iota_u32(uint32_t *p, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    p[i] = (uint32_t)i;
}

__global__ void k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    p[i] = v;
}

// Sort key for the SELL row list: (window, descending length, position), so a
// device-wide radix sort orders every window of `sigma` list entries by
// decreasing row length and keeps ties in their original order.
__global__ void k_sort_keys(const uint32_t *list, uint64_t padded,
                            const uint32_t *len, uint64_t sigma, uint64_t *keys) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= padded)
    return;
  uint32_t row = list[i];
  uint64_t l = row == 0xffffffffu ? 0u : len[row];
  keys[i] = ((i / sigma) << 41) | ((B2_SELL_MAX - l) << 32) | (i % sigma);
}

static int window_sort(cudaStream_t s, uint32_t *list, uint64_t padded,
                       const uint32_t *len, uint64_t sigma) {
  uint64_t *k_in, *k_out;
  uint32_t *v_out;
  CU_TRY(cudaMalloc(&k_in, (padded + 1) * 8));
  CU_TRY(cudaMalloc(&k_out, (padded + 1) * 8));
  CU_TRY(cudaMalloc(&v_out, (padded + 1) * 4));
  k_sort_keys<<<nblk(padded), T256, 0, s>>>(list, padded, len, sigma, k_in);
  void *tmp = nullptr;
  size_t bytes = 0;
  CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, list, v_out,
                                         padded, 0, 64, s));
  CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
  CU_TRY(cub::DeviceRadixSort::SortPairs(tmp, bytes, k_in, k_out, list, v_out,
                                         padded, 0, 64, s));
  CU_TRY(cudaMemcpyAsync(list, v_out, padded * 4, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaStreamSynchronize(s));
  cudaFree(tmp), cudaFree(k_in), cudaFree(k_out), cudaFree(v_out);
  return B200_OK;
}

// one warp per slice: width = longest row in the slice
__global__ void k_slice_width(uint32_t nslices, const uint32_t *list,
                              uint64_t n, const uint32_t *len, uint32_t *width,
                              unsigned long long *entries_true) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  uint64_t pos = (uint64_t)s * B2_SLICE + lane;
  uint32_t row = list ? list[pos] : (pos < n ? (uint32_t)pos : 0xffffffffu);
  uint32_t l = row == 0xffffffffu ? 0u : len[row];
  uint32_t w = l, t = l;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (lane == 0) {
    width[s] = w;
    atomicAdd(entries_true, (unsigned long long)t);
  }
  if (s == nslices - 1 && lane == 0)
    width[nslices] = 0;
}

__global__ void k_sell_fill(uint32_t nslices, const uint32_t *list, uint64_t n,
                            const uint32_t *len, const uint32_t *sell_off,
                            const uint64_t *offs, const uint32_t *cols,
                            const double *vals, uint32_t *scols, double *svals,
                            uint32_t range_col) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  uint64_t pos = (uint64_t)s * B2_SLICE + lane;
  uint32_t row = list ? list[pos] : (pos < n ? (uint32_t)pos : 0xffffffffu);
  uint32_t l = 0;
  uint64_t src = 0;
  if (row != 0xffffffffu)
    l = len[row], src = offs[row];
  uint32_t o = sell_off[s], w = sell_off[s + 1] - o;
  uint64_t dst = (uint64_t)o * B2_SLICE + lane;
  // padding multiplies x[own row] (always a valid local index) by zero; in a column range of
  // a column-blocked operator x[own row] may lie outside the range the pass keeps in L2, so
  // there the padding repeats the row's last column (or the first column of the range)
  uint32_t padcol = row == 0xffffffffu ? 0u : row;
  if (range_col != 0xffffffffu)
    padcol = l ? cols[src + l - 1] : range_col;
  for (uint32_t k = 0; k < w; k++, dst += B2_SLICE) {
    bool live = k < l;
    scols[dst] = live ? cols[src + k] : padcol;
    svals[dst] = live ? vals[src + k] : 0.0;
  }
}

// ---- index compression of the SELL column stream -----------------------------
// One warp per slice.  A slice is "uniform" when every lane holds a real row
// of exactly the slice width and, for every k, col(lane, k) - row(lane) is the
// same in all 32 lanes.  Such a slice needs w deltas, not 32 w columns.
__global__ void k_slice_uniform(uint32_t nslices, const uint32_t *list, uint64_t n,
                                const uint32_t *len, const uint32_t *sell_off,
                                const uint32_t *scols, uint32_t *cnt_u,
                                uint32_t *cnt_e) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  uint64_t pos = (uint64_t)s * B2_SLICE + lane;
  uint32_t row = list ? list[pos] : (pos < n ? (uint32_t)pos : 0xffffffffu);
  uint32_t o = sell_off[s], w = sell_off[s + 1] - o;
  bool ok = row != 0xffffffffu && w > 0 && len[row] == w;
  ok = __all_sync(0xffffffffu, ok);
  if (ok) {
    uint64_t src = (uint64_t)o * B2_SLICE + lane;
    for (uint32_t k = 0; k < w; k++, src += B2_SLICE) {
      const uint32_t d = scols[src] - row;
      const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);  // every lane, every k
      ok = ok && d == d0;
    }
    ok = __all_sync(0xffffffffu, ok);
  }
  if (lane == 0) {
    cnt_u[s] = ok ? w : 0u, cnt_e[s] = ok ? 0u : w;
    if (s == nslices - 1)
      cnt_u[nslices] = 0u, cnt_e[nslices] = 0u;
  }
}

__global__ void k_compact_cols(uint32_t nslices, const uint32_t *list,
                               const uint32_t *sell_off, const uint32_t *scols,
                               const uint32_t *cnt_u, const uint32_t *off_u,
                               const uint32_t *off_e, int32_t *dcols,
                               uint32_t *ecols, uint4 *meta) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  uint32_t o = sell_off[s], w = sell_off[s + 1] - o;
  if (cnt_u[s]) {
    uint32_t row0 = list ? list[(uint64_t)s * B2_SLICE] : s * B2_SLICE;
    for (uint32_t k = lane; k < w; k += 32)
      dcols[off_u[s] + k] = (int32_t)(scols[(uint64_t)(o + k) * B2_SLICE] - row0);
    if (lane == 0)
      meta[s] = make_uint4(o, w | 0x80000000u, off_u[s], 0u);
  } else {
    uint64_t src = (uint64_t)o * B2_SLICE + lane, dst = (uint64_t)off_e[s] * B2_SLICE + lane;
    for (uint32_t k = 0; k < w; k++, src += B2_SLICE, dst += B2_SLICE)
      ecols[dst] = scols[src];
    if (lane == 0)
      meta[s] = make_uint4(o, w, off_e[s], 0u);
  }
}

__global__ void k_pad4_len(uint32_t nrows, const uint32_t *ids,
                           const uint32_t *len, uint64_t *out,
                           unsigned long long *true_nnz) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) {
    out[i] = (len[ids[i]] + 3u) & ~3ull;
    atomicAdd(true_nnz, (unsigned long long)len[ids[i]]);
  }
  if (i == nrows)
    out[i] = 0;
}

__global__ void k_add_const_u64(uint64_t *p, uint64_t n, uint64_t v) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    p[i] += v;
}

// one warp per row of the vector / long bins
__global__ void k_vl_fill(uint32_t nrows, const uint32_t *ids,
                          const uint32_t *len, const uint64_t *voff,
                          const uint64_t *offs, const uint32_t *cols,
                          const double *vals, uint32_t *vcols, double *vvals) {
  uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= nrows)
    return;
  uint32_t row = ids[r], l = len[row];
  uint64_t src = offs[row], dst = voff[r], cap = voff[r + 1] - dst;
  for (uint64_t k = lane; k < cap; k += 32) {
    bool live = k < l;
    vcols[dst + k] = live ? cols[src + k] : row;
    vvals[dst + k] = live ? vals[src + k] : 0.0;
  }
}

__global__ void k_inv_diag(uint64_t n, const uint64_t *offs,
                           const uint32_t *cols, const double *vals,
                           double *dinv) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double d = 0.0;
  for (uint64_t k = offs[i]; k < offs[i + 1]; k++)
    if (cols[k] == i)
      d = vals[k];
  dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
}

// B200_MAT_VALUES_F32: the SELL value stream rounded to fp32; *inexact is set
// when some value did not survive the rounding (NaN counts as not surviving)
__global__ void k_vals_to_f32(const double *__restrict__ v, float *__restrict__ o,
                              uint64_t n, unsigned *inexact) {
  bool bad = false;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = v[i];
    const float f = (float)d;
    o[i] = f;
    bad |= !((double)f == d);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
    atomicOr(inexact, 1u);
}

int build_layout(b200_ctx *c, PlainCsr *A, uint64_t n_global,
                 uint64_t row_begin, uint32_t flags, b200_mat **out) {
  // NOTE: *out may already carry halo / partition fields; create if null.
  b200_mat *M = *out ? *out : new b200_mat();
  *out = M;
  cudaStream_t s = c->stream;
  uint64_t n = A->n;
  M->ctx = c, M->n_global = n_global, M->row_begin = row_begin;
  M->n_local = n, M->nnz = A->nnz, M->flags = flags;
  if (n >= 0xffffffffull)
    B_FAIL(B200_ERANGE, "b200: %llu local rows exceed 32-bit row ids",
           (unsigned long long)n);

  // ---- row lengths, histogram ------------------------------------------------
  B_TRY(dev_alloc(M, (void **)&M->row_len, (n + 1) * 4));
  unsigned long long *d_hist;
  CU_TRY(cudaMalloc(&d_hist, (B200_HIST_BINS + 2) * 8));
  CU_TRY(cudaMemsetAsync(d_hist, 0, (B200_HIST_BINS + 2) * 8, s));
  if (n)
    k_row_len<<<nblk(n), T256, 0, s>>>(n, A->offs, M->row_len, d_hist,
                                       d_hist + B200_HIST_BINS);
  unsigned long long h_hist[B200_HIST_BINS + 2];
  CU_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof h_hist, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  for (int b = 0; b < B200_HIST_BINS; b++)
    M->hist[b] = h_hist[b];
  M->max_row_len = h_hist[B200_HIST_BINS];

  // ---- kernel selection by row-length histogram -----------------------------
  uint32_t sell_max = B2_SELL_MAX, long_min = B2_LONG_MIN;
  if (flags & B200_MAT_FORCE_VECTOR)
    sell_max = 0, long_min = 0xffffffffu;  // every row: one warp
  if (flags & B200_MAT_FORCE_SELL) {
    if (M->max_row_len > B2_SELL_MAX)
      B_FAIL(B200_EINVAL, "FORCE_SELL: longest row %llu > %u",
             (unsigned long long)M->max_row_len, B2_SELL_MAX);
    long_min = 0xffffffffu;
  }

  uint32_t *f[3], *p[3], *ids[3] = {nullptr, nullptr, nullptr};
  uint32_t cnt[3] = {0, 0, 0};
  bool all_sell = M->max_row_len <= sell_max;
  if (all_sell) {
    cnt[0] = (uint32_t)n;
  } else {
    for (int b = 0; b < 3; b++) {
      CU_TRY(cudaMalloc(&f[b], (n + 1) * 4));
      CU_TRY(cudaMalloc(&p[b], (n + 1) * 4));
    }
    k_bin_flags<<<nblk(n + 1), T256, 0, s>>>(n, M->row_len, sell_max, long_min,
                                             f[0], f[1], f[2]);
    for (int b = 0; b < 3; b++) {
      B_TRY(exclusive_scan<uint32_t>(s, f[b], p[b], n + 1));
      CU_TRY(cudaMemcpy(&cnt[b], p[b] + n, 4, cudaMemcpyDeviceToHost));
    }
  }
  uint64_t sell_padded_rows = ((uint64_t)cnt[0] + B2_SLICE - 1) / B2_SLICE * B2_SLICE;
  if (!all_sell) {
    CU_TRY(cudaMalloc(&ids[0], (sell_padded_rows + 1) * 4));
    CU_TRY(cudaMalloc(&ids[1], (cnt[1] + 1ull) * 4));
    CU_TRY(cudaMalloc(&ids[2], (cnt[2] + 1ull) * 4));
    k_fill_u32<<<nblk(sell_padded_rows + 1), T256, 0, s>>>(ids[0], sell_padded_rows + 1, 0xffffffffu);
    k_scatter_ids<<<nblk(n), T256, 0, s>>>(n, M->row_len, sell_max, long_min,
                                           p[0], p[1], p[2], ids[0], ids[1], ids[2]);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < 3; b++)
      cudaFree(f[b]), cudaFree(p[b]);
  }

  // ---- SELL bin ----------------------------------------------------------------
  M->sell_rows = cnt[0];
  M->sell_slices = (uint32_t)(sell_padded_rows / B2_SLICE);
  if (M->sell_slices) {
    uint32_t ns = M->sell_slices;
    uint32_t *width;
    CU_TRY(cudaMalloc(&width, (ns + 1ull) * 4));
    unsigned long long *d_true = d_hist;  // reuse
    auto widths = [&](const uint32_t *list, uint64_t *padded, uint64_t *truth) -> int {
      CU_TRY(cudaMemsetAsync(d_true, 0, 8, s));
      k_slice_width<<<nblk((uint64_t)ns * 32), T256, 0, s>>>(ns, list, n, M->row_len, width, d_true);
      B_TRY(exclusive_scan<uint32_t>(s, width, M->sell_off, ns + 1ull));
      uint32_t tot = 0;
      unsigned long long tr = 0;
      CU_TRY(cudaMemcpy(&tot, M->sell_off + ns, 4, cudaMemcpyDeviceToHost));
      CU_TRY(cudaMemcpy(&tr, d_true, 8, cudaMemcpyDeviceToHost));
      *padded = (uint64_t)tot * B2_SLICE, *truth = tr;
      return B200_OK;
    };
    B_TRY(dev_alloc(M, (void **)&M->sell_off, (ns + 1ull) * 4));
    if (M->grouped_slices) {
      B_TRY(dev_alloc(M, (void **)&M->grp_work, 8));
      CU_TRY(cudaMemsetAsync(M->grp_work, 0, 8, s));
    }
    uint64_t padded = 0, truth = 0;
    B_TRY(widths(ids[0], &padded, &truth));
    bool sort = !(flags & B200_MAT_NO_SORT) && truth > 0 &&
                (double)(padded - truth) > 0.03 * (double)truth;
    if (sort) {
      if (!ids[0]) {  // identity list so far: materialise it
        CU_TRY(cudaMalloc(&ids[0], (sell_padded_rows + 1) * 4));
        k_fill_u32<<<nblk(sell_padded_rows + 1), T256, 0, s>>>(ids[0], sell_padded_rows + 1, 0xffffffffu);
        k_iota_u32<<<nblk(n), T256, 0, s>>>(ids[0], n);
        CU_TRY(cudaGetLastError());
      }
      // widen the window until the padding is small: 1024 rows, 32768, all
      // (a column range of a column-blocked operator keeps its windows small -- sell_sigma_cap:
      // rows that are multiplied together stay neighbours, so the y they update and the
      // near-diagonal x they gather stay in cache; measured, power-law 50 M rows, 64 MB ranges,
      // work units of 8 groups: 65 536 rows 5.93 ms, 131 072 5.87, 262 144 5.89, 1 048 576 6.4,
      // the whole list 7.2; with units of 64 groups 32 768 was the optimum, at 6.60)
      const uint64_t cap = M->sell_sigma_cap;
      const uint64_t sig[3] = {B2_SELL_SIGMA, cap ? cap : 32 * B2_SELL_SIGMA, sell_padded_rows};
      for (int t = 0; t < (cap ? 2 : 3); t++) {
        uint64_t sg = sig[t] < sell_padded_rows ? sig[t] : sell_padded_rows;
        if (t > 0 && sg <= sig[t - 1])
          break;
        B_TRY(window_sort(s, ids[0], sell_padded_rows, M->row_len, sg));
        B_TRY(widths(ids[0], &padded, &truth));
        M->sell_sigma = (uint32_t)sg;
        if ((double)(padded - truth) <= 0.03 * (double)truth || sg == sell_padded_rows)
          break;
      }
    }
    M->sell_entries = padded;
    B_TRY(dev_alloc(M, (void **)&M->sell_cols, (padded ? padded : 1) * 4));
    B_TRY(dev_alloc(M, (void **)&M->sell_vals, (padded ? padded : 1) * 8));
    k_sell_fill<<<nblk((uint64_t)ns * 32), T256, 0, s>>>(
        ns, ids[0], n, M->row_len, M->sell_off, A->offs, A->cols, A->vals,
        M->sell_cols, M->sell_vals, M->pad_col);
    CU_TRY(cudaGetLastError());
    uint32_t wmax = 0;
    {
      void *tmp = nullptr;
      size_t bytes = 0;
      uint32_t *d_max;
      CU_TRY(cudaMalloc(&d_max, 4));
      CU_TRY(cub::DeviceReduce::Max(nullptr, bytes, width, d_max, (int)ns, s));
      CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
      CU_TRY(cub::DeviceReduce::Max(tmp, bytes, width, d_max, (int)ns, s));
      CU_TRY(cudaMemcpyAsync(&wmax, d_max, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      cudaFree(tmp), cudaFree(d_max);
    }
    M->sell_max_width = wmax;
    cudaFree(width);
    // ---- index compression: uniform slices keep w deltas instead of 32 w columns
    M->sell_col_entries = padded;
    if (!(flags & B200_MAT_NO_COMPRESS) && padded && !M->grouped_slices) {
      uint32_t *cnt_u, *cnt_e, *off_u, *off_e;
      CU_TRY(cudaMalloc(&cnt_u, (ns + 1ull) * 4));
      CU_TRY(cudaMalloc(&cnt_e, (ns + 1ull) * 4));
      CU_TRY(cudaMalloc(&off_u, (ns + 1ull) * 4));
      CU_TRY(cudaMalloc(&off_e, (ns + 1ull) * 4));
      k_slice_uniform<<<nblk((uint64_t)ns * 32), T256, 0, s>>>(
          ns, ids[0], n, M->row_len, M->sell_off, M->sell_cols, cnt_u, cnt_e);
      CU_TRY(cudaGetLastError());
      B_TRY(exclusive_scan<uint32_t>(s, cnt_u, off_u, ns + 1ull));
      B_TRY(exclusive_scan<uint32_t>(s, cnt_e, off_e, ns + 1ull));
      uint32_t tot_u = 0, tot_e = 0;
      CU_TRY(cudaMemcpy(&tot_u, off_u + ns, 4, cudaMemcpyDeviceToHost));
      CU_TRY(cudaMemcpy(&tot_e, off_e + ns, 4, cudaMemcpyDeviceToHost));
      // worth it when at least a tenth of the column stream goes away; both
      // offsets must leave bit 31 free
      if ((uint64_t)tot_u * 10 >= (uint64_t)tot_u + tot_e && tot_u < 0x80000000u &&
          tot_e < 0x80000000u) {
        uint32_t *ecols = nullptr;
        int32_t *dcols = nullptr;
        uint4 *meta = nullptr;
        CU_TRY(cudaMalloc(&ecols, ((uint64_t)tot_e * B2_SLICE + 1) * 4));
        CU_TRY(cudaMalloc(&dcols, ((uint64_t)tot_u + 40) * 4));
        CU_TRY(cudaMalloc(&meta, (ns + 1ull) * 16));
        CU_TRY(cudaMemsetAsync(dcols + tot_u, 0, 40 * 4, s));
        CU_TRY(cudaMemsetAsync(meta + ns, 0, 16, s));
        k_compact_cols<<<nblk((uint64_t)ns * 32), T256, 0, s>>>(
            ns, ids[0], M->sell_off, M->sell_cols, cnt_u, off_u, off_e, dcols, ecols, meta);
        CU_TRY(cudaGetLastError());
        // how many slices went uniform (for the info block)
        {
          std::vector<uint32_t> h(ns);
          CU_TRY(cudaMemcpy(h.data(), cnt_u, (size_t)ns * 4, cudaMemcpyDeviceToHost));
          uint64_t nu = 0;
          for (uint32_t v : h)
            nu += v != 0;
          M->sell_uniform_slices = nu;
        }
        CU_TRY(cudaStreamSynchronize(s));
        cudaFree(M->sell_cols);
        M->device_bytes -= (padded ? padded : 1) * 4;
        M->sell_cols = ecols, M->sell_dcols = dcols, M->sell_meta = (uint32_t *)meta;
        M->sell_col_entries = (uint64_t)tot_e * B2_SLICE, M->sell_delta_entries = tot_u;
        M->device_bytes += ((uint64_t)tot_e * B2_SLICE + 1) * 4 + ((uint64_t)tot_u + 40) * 4 +
                           (ns + 1ull) * 16;
      }
      cudaFree(cnt_u), cudaFree(cnt_e), cudaFree(off_u), cudaFree(off_e);
    }
    if (ids[0]) {
      M->sell_perm = ids[0];
      M->device_bytes += (sell_padded_rows + 1) * 4;
      ids[0] = nullptr;
    }
    // ---- fp32 value stream ------------------------------------
    if ((flags & B200_MAT_VALUES_F32) && padded) {
      unsigned *d_bad, h_bad = 0;
      CU_TRY(cudaMalloc(&d_bad, 4));
      CU_TRY(cudaMemsetAsync(d_bad, 0, 4, s));
      B_TRY(dev_alloc(M, (void **)&M->sell_vals32, padded * 4));
      uint64_t g = (padded + T256 - 1) / T256;
      if (g > (uint64_t)c->sm_count * 16)
        g = (uint64_t)c->sm_count * 16;
      k_vals_to_f32<<<(unsigned)g, T256, 0, s>>>(M->sell_vals, M->sell_vals32, padded, d_bad);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      cudaFree(d_bad);
      M->vals32_exact = h_bad == 0;
      if (M->vals32_exact) {  // lossless: the fp64 copy is not needed any more
        cudaFree(M->sell_vals);
        M->sell_vals = nullptr;
        M->device_bytes -= padded * 8;
      }
    }
  }

  // ---- vector and long bins: one shared row-major store ------------------------
  M->vec_rows = cnt[1], M->long_rows = cnt[2];
  if (cnt[1] + cnt[2]) {
    uint64_t *len4[2] = {nullptr, nullptr};
    uint64_t tot[2] = {0, 0};
    uint64_t **offp[2] = {&M->vec_off, &M->long_off};
    for (int b = 0; b < 2; b++) {
      uint32_t nr = cnt[1 + b];
      B_TRY(dev_alloc(M, (void **)offp[b], (nr + 1ull) * 8));
      CU_TRY(cudaMalloc(&len4[b], (nr + 1ull) * 8));
      CU_TRY(cudaMemsetAsync(d_hist, 0, 8, s));
      k_pad4_len<<<nblk(nr + 1ull), T256, 0, s>>>(nr, ids[1 + b], M->row_len, len4[b], d_hist);
      unsigned long long tn = 0;
      CU_TRY(cudaMemcpyAsync(&tn, d_hist, 8, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      (b == 0 ? M->vec_nnz : M->long_nnz) = tn;
      B_TRY(exclusive_scan<uint64_t>(s, len4[b], *offp[b], nr + 1ull));
      CU_TRY(cudaMemcpy(&tot[b], *offp[b] + nr, 8, cudaMemcpyDeviceToHost));
      cudaFree(len4[b]);
    }
    // long rows live after the vector rows
    if (cnt[2])
      k_add_const_u64<<<nblk(cnt[2] + 1ull), T256, 0, s>>>(M->long_off, cnt[2] + 1ull, tot[0]);
    M->vl_entries = tot[0] + tot[1];
    B_TRY(dev_alloc(M, (void **)&M->vl_cols, (M->vl_entries + 4) * 4));
    B_TRY(dev_alloc(M, (void **)&M->vl_vals, (M->vl_entries + 4) * 8));
    if (cnt[1])
      k_vl_fill<<<nblk((uint64_t)cnt[1] * 32), T256, 0, s>>>(
          cnt[1], ids[1], M->row_len, M->vec_off, A->offs, A->cols, A->vals,
          M->vl_cols, M->vl_vals);
    if (cnt[2])
      k_vl_fill<<<nblk((uint64_t)cnt[2] * 32), T256, 0, s>>>(
          cnt[2], ids[2], M->row_len, M->long_off, A->offs, A->cols, A->vals,
          M->vl_cols, M->vl_vals);
    CU_TRY(cudaGetLastError());
    M->vec_row_ids = ids[1], M->long_row_ids = ids[2];
    M->device_bytes += (cnt[1] + cnt[2] + 2ull) * 4;
    ids[1] = ids[2] = nullptr;
  }

  // ---- Jacobi preconditioner ------------------------------------------------------
  // (columns are local ids here; the diagonal of local row i is column i)
  B_TRY(dev_alloc(M, (void **)&M->dinv, (n + 1) * 8));
  if (n)
    k_inv_diag<<<nblk(n), T256, 0, s>>>(n, A->offs, A->cols, A->vals, M->dinv);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  cudaFree(d_hist);
  for (int b = 0; b < 3; b++)
    if (ids[b]) cudaFree(ids[b]);
  if (!M->halo.n_halo)
    M->interior_begin = 0, M->interior_end = n;
  return B200_OK;
}

#include "colblock_kernels.cuh"

// What build_layout is to one matrix, for a column-blocked one: M keeps the row
// lengths, the histogram and D^-1 of the whole operator and no entries; every
// column range becomes a child matrix over all rows.  Falls back to
// build_layout when blocking does not apply.
int build_layout_or_blocks(b200_ctx *c, PlainCsr *A, uint64_t n_global, uint64_t row_begin,
                           uint32_t flags, b200_mat **out) {
  uint64_t mb = 64;
  if (const char *v = getenv("B200_COL_BLOCK_MB"))
    if (atoll(v) > 0)
      mb = (uint64_t)atoll(v);
  const uint64_t width = (mb << 20) / 8 / 32 * 32;
  // The columns in global order: [remote below the own rows | owned | remote above] (one rank:
  // owned only).  Each of the three segments is cut into equal pieces of at most `width`
  // columns, so that a range never crosses a seam and is one contiguous piece of the
  // extended x vector.
  const uint64_t n_own = A->n;
  const uint64_t n_halo = *out ? (*out)->halo.n_halo : 0, n_low = *out ? (*out)->halo.n_low : 0;
  std::vector<uint64_t> cuts(1, 0);
  std::vector<int> local_piece;  // per range: 1 when it holds owned columns only
  {
    const uint64_t seg[3] = {n_low, n_own, n_halo - n_low};
    uint64_t at = 0;
    for (int g = 0; g < 3 && width; g++) {
      const uint64_t pieces = (seg[g] + width - 1) / width;
      for (uint64_t k = 0; k < pieces; k++) {
        uint64_t hi = k + 1 == pieces ? seg[g] : ((seg[g] * (k + 1) / pieces) + 31) / 32 * 32;
        hi = hi < seg[g] ? hi : seg[g];
        if (at + hi > cuts.back())
          cuts.push_back(at + hi), local_piece.push_back(g == 1);
      }
      at += seg[g];
    }
  }
  const uint64_t nb64 = cuts.size() - 1;
  if (!(flags & B200_MAT_COL_BLOCK) || nb64 < 2 || A->nnz == 0 ||
      (c->nranks == 1 && A->n != n_global))
    return build_layout(c, A, n_global, row_begin, flags & ~(uint32_t)B200_MAT_COL_BLOCK, out);
  if (nb64 > 64)
    B_FAIL(B200_EINVAL, "B200_MAT_COL_BLOCK: %llu column blocks of %llu MB (at most 64)",
           (unsigned long long)nb64, (unsigned long long)mb);
  const uint32_t nb = (uint32_t)nb64;
  cudaStream_t s = c->stream;
  const uint64_t n = A->n;
  b200_mat *M = *out ? *out : new b200_mat();
  *out = M;
  M->ctx = c, M->n_global = n_global, M->row_begin = row_begin;
  M->n_local = n, M->nnz = A->nnz, M->flags = flags, M->col_block_width = width;
  if (!n_halo)
    M->interior_begin = 0, M->interior_end = n;
  // ---- what describes the whole operator: row lengths, histogram, D^-1 ---------------------
  B_TRY(dev_alloc(M, (void **)&M->row_len, (n + 1) * 4));
  B_TRY(dev_alloc(M, (void **)&M->dinv, (n + 1) * 8));
  {
    unsigned long long *d_hist, h_hist[B200_HIST_BINS + 2];
    CU_TRY(cudaMalloc(&d_hist, (B200_HIST_BINS + 2) * 8));
    CU_TRY(cudaMemsetAsync(d_hist, 0, (B200_HIST_BINS + 2) * 8, s));
    k_row_len<<<nblk(n), T256, 0, s>>>(n, A->offs, M->row_len, d_hist, d_hist + B200_HIST_BINS);
    k_inv_diag<<<nblk(n), T256, 0, s>>>(n, A->offs, A->cols, A->vals, M->dinv);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof h_hist, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(d_hist);
    for (int b = 0; b < B200_HIST_BINS; b++)
      M->hist[b] = h_hist[b];
    M->max_row_len = h_hist[B200_HIST_BINS];
  }
  // ---- cut the rows ------------------------------------------------------------------------
  {
    unsigned *d_bad, h_bad = 0;
    CU_TRY(cudaMalloc(&d_bad, 4));
    CU_TRY(cudaMemsetAsync(d_bad, 0, 4, s));
    k_rows_sorted<<<nblk(n), T256, 0, s>>>(n, A->offs, A->cols, n_own, n_low, d_bad);
    CU_TRY(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(d_bad);
    if (h_bad)
      B_FAIL(B200_EINVAL, "B200_MAT_COL_BLOCK needs rows with ascending columns");
  }
  uint64_t *cnt = nullptr, *boffs = nullptr, *d_cuts = nullptr;
  CU_TRY(cudaMalloc(&cnt, (size_t)nb * (n + 1) * 8));
  CU_TRY(cudaMalloc(&boffs, (size_t)nb * (n + 1) * 8));
  CU_TRY(cudaMalloc(&d_cuts, (nb + 1ull) * 8));
  CU_TRY(cudaMemcpyAsync(d_cuts, cuts.data(), (nb + 1ull) * 8, cudaMemcpyHostToDevice, s));
  k_colblock_count<<<nblk(n + 1), T256, 0, s>>>(n, A->offs, A->cols, d_cuts, nb, n_own, n_low, cnt);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));  // (cuts: pageable source, landed before the vector goes)
  cudaFree(d_cuts);
  std::vector<PlainCsr> sub(nb);
  std::vector<uint32_t *> h_cols(nb, nullptr);
  std::vector<double *> h_vals(nb, nullptr);
  int rc = B200_OK;
  for (uint32_t b = 0; b < nb && rc == B200_OK; b++) {
    rc = exclusive_scan<uint64_t>(s, cnt + (size_t)b * (n + 1), boffs + (size_t)b * (n + 1), n + 1);
    if (rc != B200_OK)
      break;
    sub[b].n = n;
    if (cudaMemcpy(&sub[b].nnz, boffs + (size_t)b * (n + 1) + n, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMalloc(&sub[b].offs, (n + 1) * 8) != cudaSuccess ||
        cudaMalloc(&sub[b].cols, (sub[b].nnz ? sub[b].nnz : 1) * 4) != cudaSuccess ||
        cudaMalloc(&sub[b].vals, (sub[b].nnz ? sub[b].nnz : 1) * 8) != cudaSuccess ||
        cudaMemcpyAsync(sub[b].offs, boffs + (size_t)b * (n + 1), (n + 1) * 8,
                        cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
      b200_set_error("b200: column block %u: %s", b, cudaGetErrorString(cudaGetLastError()));
      rc = B200_ENOMEM;
    }
    h_cols[b] = sub[b].cols, h_vals[b] = sub[b].vals;
  }
  uint32_t **d_cols = nullptr;
  double **d_vals = nullptr;
  if (rc == B200_OK) {
    CU_TRY(cudaMalloc(&d_cols, nb * sizeof(uint32_t *)));
    CU_TRY(cudaMalloc(&d_vals, nb * sizeof(double *)));
    // (in the stream that consumes them: a blocking copy on the legacy stream is not
    // ordered with a non-blocking stream)
    CU_TRY(cudaMemcpyAsync(d_cols, h_cols.data(), nb * sizeof(uint32_t *), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d_vals, h_vals.data(), nb * sizeof(double *), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));
    const uint64_t warps = n < (uint64_t)c->sm_count * 64 ? n : (uint64_t)c->sm_count * 64;
    k_colblock_fill<<<(unsigned)((warps * 32 + T256 - 1) / T256), T256, 0, s>>>(
        n, A->offs, A->cols, A->vals, nb, boffs, d_cols, d_vals);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
  }
  cudaFree(cnt), cudaFree(boffs);
  if (d_cols) cudaFree(d_cols);
  if (d_vals) cudaFree(d_vals);
  // ---- every range: a layout of its own ----------------------------------------------------------
  const uint32_t child_flags = flags & ~(uint32_t)(B200_MAT_COL_BLOCK | B200_MAT_SYM_UPPER);
  // B200_COL_BLOCK_SIGMA: widest length-sort window of a range (rows; 0 = the whole list, as a
  // matrix of its own would choose); B200_COL_BLOCK_KERNEL=plain: one slice per warp trip
  uint64_t sigma_cap = 128 * B2_SELL_SIGMA;
  if (const char *v = getenv("B200_COL_BLOCK_SIGMA"))
    sigma_cap = (uint64_t)atoll(v);
  const char *kv = getenv("B200_COL_BLOCK_KERNEL");
  const bool grouped = !(kv && !strcmp(kv, "plain"));
  std::vector<b200_mat *> kids(nb, nullptr);
  for (uint32_t b = 0; b < nb; b++) {
    if (rc == B200_OK) {
      b200_mat *child = new b200_mat();
      child->sell_sigma_cap = sigma_cap;
      child->grouped_slices = grouped && !(flags & B200_MAT_VALUES_F32);
      // (a column id inside the range, for the padding: first column of the range)
      const uint64_t o = cuts[b];
      const uint64_t first_col = o < n_low ? n_own + o : (o < n_low + n_own ? o - n_low : o);
      child->pad_col = grouped ? (uint32_t)first_col : 0xffffffffu;
      rc = build_layout(c, &sub[b], n_global, row_begin, child_flags, &child);
      kids[b] = child;
      if (child)
        M->device_bytes += child->device_bytes;
    }
    plain_free(&sub[b]);
  }
  // pass order: the ranges of owned columns first -- on several ranks they are multiplied while
  // the halo is on its way -- then the remote ones; block_in_col_order keeps the global order
  // (a row of the operator = its pieces in that order, b200_mat_export)
  M->block_in_col_order.assign(nb, 0);
  for (int want = 1; want >= 0; want--)
    for (uint32_t b = 0; b < nb; b++)
      if (local_piece[b] == want && kids[b]) {
        M->block_in_col_order[b] = (uint32_t)M->blocks.size();
        M->blocks.push_back(kids[b]);
        M->n_local_blocks += want;
      }
  return rc;
}

// ---------------------------------------------------------------------------
// C ABI
// plain device CSR (0-based columns, global rows) -> matrix; consumes A
int mat_from_plain(b200_ctx *c, PlainCsr *A, uint32_t flags, b200_mat **out) {
  const uint64_t nrows = A->n;
  uint32_t patsym = 1;
  int rc = B200_OK;
  if (flags & B200_MAT_SYM_UPPER)
    rc = sym_upper(c, A, &patsym);
  if (rc != B200_OK) {
    plain_free(A);
    return rc;
  }
  b200_mat *M = new b200_mat();
  M->pattern_symmetric = patsym;
  M->ctx = c;
  rc = partition_and_renumber(c, A, nrows, 0, M);
  if (rc == B200_OK)
    rc = build_layout_or_blocks(c, A, nrows, M->row_begin, flags, &M);
  plain_free(A);
  if (rc == B200_OK && c->nranks > 1)
    rc = halo_setup(M);
  if (rc != B200_OK) {
    b200_mat_destroy(M);
    return rc;
  }
  *out = M;
  return B200_OK;
}

extern "C" int b200_mat_from_csr(b200_ctx *c, uint32_t nrows, uint32_t base,
                                 const uint32_t *offs, const uint32_t *cols,
                                 const double *vals, uint32_t flags,
                                 b200_mat **out) {
  if (!c || !offs || !cols || !vals || !out)
    B_FAIL(B200_EINVAL, "b200_mat_from_csr: null argument");
  if (base > 1 || nrows == 0)
    B_FAIL(B200_EINVAL, "b200_mat_from_csr: nrows=%u base=%u", nrows, base);
  if ((flags & B200_MAT_FORCE_SELL) && (flags & B200_MAT_FORCE_VECTOR))
    B_FAIL(B200_EINVAL, "b200_mat_from_csr: contradictory FORCE flags");
  CU_TRY(cudaSetDevice(c->device));
  *out = nullptr;
  PlainCsr A;
  int rc;
  if (c->nranks > 1) {
    // Several ranks, one host CSR: a rank uploads what ITS row block [r0, r1) is made of,
    // not the whole matrix -- its own rows, and (upper triangle mirrored) the entries of
    // the rows above that fall into its columns, which become its mirrored entries.  The
    // rows it does not hold stay empty; the block is cut out after the mirror
    // (dist.cu partition_and_renumber).  Filtered on the host in one pass.
    uint64_t r0, r1;
    b200_row_block(nrows, c->rank, c->nranks, &r0, &r1);
    const bool sym = flags & B200_MAT_SYM_UPPER;
    std::vector<uint32_t> fo((size_t)nrows + 1, 0u), fc;
    std::vector<double> fv;
    for (uint64_t i = sym ? 0 : r0; i < r1; i++) {
      const bool own = i >= r0;
      for (uint32_t e = offs[i]; e < offs[i + 1]; e++) {
        const uint64_t cc = (uint64_t)cols[e] - base;
        if (own || (cc >= r0 && cc < r1))
          fc.push_back(cols[e]), fv.push_back(vals[e]);
      }
      fo[i + 1] = (uint32_t)fc.size();
    }
    for (uint64_t i = 0; i < nrows; i++)
      if (fo[i + 1] < fo[i])
        fo[i + 1] = fo[i];  // rows not held: empty
    if (fc.empty())
      fc.push_back(base), fv.push_back(0.0);
    rc = upload_csr(c, nrows, base, fo.data(), fc.data(), fv.data(), &A);
  } else {
    rc = upload_csr(c, nrows, base, offs, cols, vals, &A);
  }
  if (rc != B200_OK) {
    plain_free(&A);
    return rc;
  }
  return mat_from_plain(c, &A, flags, out);
}

extern "C" int b200_mat_destroy(b200_mat *M) {
  if (!M)
    return B200_OK;
  if (M->ctx)
    cudaSetDevice(M->ctx->device);
  cudaDeviceSynchronize();
  for (b200_mat *child : M->blocks)
    b200_mat_destroy(child);
  M->blocks.clear();
  small_free(M);
  halo_free(M);
  if (M->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)M->graph_exec);
  void *ptrs[] = {M->sell_off, M->sell_cols, M->sell_vals, M->sell_vals32, M->sell_perm,
                  M->w_d, M->w_rhs,
                  M->sell_meta, M->sell_dcols,
                  M->vec_row_ids, M->long_row_ids, M->vec_off, M->long_off,
                  M->vl_cols, M->vl_vals, M->dinv, M->row_len, M->w_r, M->w_p,
                  M->w_q, M->w_x, M->x_ext, M->partials, M->state, M->stage_b, M->stage_x, M->grp_work};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete M;
  return B200_OK;
}

extern "C" int b200_mat_get_info(const b200_mat *M, b200_mat_info *o) {
  if (!M || !o)
    B_FAIL(B200_EINVAL, "b200_mat_get_info: null argument");
  memset(o, 0, sizeof *o);
  o->n_global = M->n_global, o->row_begin = M->row_begin, o->n_local = M->n_local;
  o->n_halo = M->halo.n_halo, o->nnz = M->nnz;
  o->nnz_padded = M->sell_entries + M->vl_entries;
  o->sell_rows = M->sell_rows, o->sell_slices = M->sell_slices;
  o->sell_sigma = M->sell_sigma, o->sell_max_width = M->sell_max_width;
  o->vec_rows = M->vec_rows, o->long_rows = M->long_rows;
  o->vec_nnz = M->vec_nnz, o->long_nnz = M->long_nnz;
  o->interior_begin = M->interior_begin, o->interior_end = M->interior_end;
  memcpy(o->hist, M->hist, sizeof o->hist);
  o->max_row_len = M->max_row_len;
  o->pattern_symmetric = M->pattern_symmetric;
  o->sell_perm = M->sell_perm != nullptr;
  o->device_bytes = M->device_bytes;
  o->sell_uniform_slices = M->sell_uniform_slices;
  o->values_f32 = M->sell_vals32 ? (M->vals32_exact ? 1u : 2u) : 0u;
  o->col_blocks = (uint32_t)M->blocks.size();
  if (!M->blocks.empty()) {  // the sums over the column ranges
    b200_mat_info ci;
    for (const b200_mat *child : M->blocks) {
      B_TRY(b200_mat_get_info(child, &ci));
      o->nnz_padded += ci.nnz_padded, o->matrix_stream_bytes += ci.matrix_stream_bytes;
      o->sell_rows += ci.sell_rows, o->sell_slices += ci.sell_slices;
      o->vec_rows += ci.vec_rows, o->vec_nnz += ci.vec_nnz;
      o->long_rows += ci.long_rows, o->long_nnz += ci.long_nnz;
      o->values_f32 = ci.values_f32;
    }
    return B200_OK;
  }
  // what one SpMV (of the PCG iteration) streams from the matrix: values,
  // columns / deltas, slice and row offsets, the row permutation
  o->matrix_stream_bytes =
      (M->sell_vals32 ? 4 : 8) * M->sell_entries + 8 * M->vl_entries +
      4 * (M->sell_col_entries + M->sell_delta_entries + M->vl_entries) +
      (M->sell_meta ? 16 : 4) * (uint64_t)(M->sell_slices + 1) +
      (M->sell_perm ? 4 * (uint64_t)M->sell_slices * B2_SLICE : 0) +
      12 * ((uint64_t)M->vec_rows + M->long_rows);
  return B200_OK;
}

extern "C" int b200_mat_algorithmic_bytes(const b200_mat *M, uint64_t *spmv,
                                          uint64_t *pcg) {
  if (!M)
    B_FAIL(B200_EINVAL, "b200_mat_algorithmic_bytes: null matrix");
  uint64_t mat = 12 * M->nnz + 4 * (M->n_local + 1);
  if (spmv) *spmv = mat + 16 * M->n_local;
  if (pcg) *pcg = mat + 104 * M->n_local;
  return B200_OK;
}

// ---- export: what the layout represents, in original row order -----------------
__global__ void k_export_sell(uint32_t nslices, const uint32_t *list, uint64_t n,
                              const uint32_t *len, const uint32_t *sell_off,
                              const uint4 *meta, const uint32_t *scols,
                              const int32_t *dcols, const double *svals,
                              const float *svals32,
                              const uint64_t *ooffs, uint32_t *ocols, double *ovals) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices)
    return;
  uint64_t pos = (uint64_t)s * B2_SLICE + lane;
  uint32_t row = list ? list[pos] : (pos < n ? (uint32_t)pos : 0xffffffffu);
  if (row == 0xffffffffu)
    return;
  uint64_t src = (uint64_t)sell_off[s] * B2_SLICE + lane, dst = ooffs[row];
  const uint4 m = meta ? meta[s] : make_uint4(0u, 0u, 0u, 0u);
  const uint32_t c = m.z;
  if (meta && (m.y >> 31)) {  // uniform slice: column = row + delta
    const int32_t *dp = dcols + c;
    for (uint32_t k = 0; k < len[row]; k++, src += B2_SLICE)
      ocols[dst + k] = row + (uint32_t)dp[k],
                 ovals[dst + k] = svals ? svals[src] : (double)svals32[src];
    return;
  }
  uint64_t csrc = meta ? (uint64_t)c * B2_SLICE + lane : src;
  for (uint32_t k = 0; k < len[row]; k++, src += B2_SLICE, csrc += B2_SLICE)
    ocols[dst + k] = scols[csrc], ovals[dst + k] = svals ? svals[src] : (double)svals32[src];
}

__global__ void k_export_vl(uint32_t nrows, const uint32_t *ids,
                            const uint32_t *len, const uint64_t *voff,
                            const uint32_t *vcols, const double *vvals,
                            const uint64_t *ooffs, uint32_t *ocols, double *ovals) {
  uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= nrows)
    return;
  uint32_t row = ids[r];
  uint64_t src = voff[r], dst = ooffs[row];
  for (uint32_t k = lane; k < len[row]; k += 32)
    ocols[dst + k] = vcols[src + k], ovals[dst + k] = vvals[src + k];
}

__global__ void k_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[i];
  if (i == n)
    out[i] = 0;
}

extern "C" int b200_mat_export(const b200_mat *M, uint64_t *offs,
                               uint32_t *cols, double *vals) {
  if (!M)
    B_FAIL(B200_EINVAL, "b200_mat_export: null matrix");
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  uint64_t n = M->n_local;
  uint64_t *l64, *ooffs;
  CU_TRY(cudaMalloc(&l64, (n + 1) * 8));
  CU_TRY(cudaMalloc(&ooffs, (n + 1) * 8));
  k_u32_to_u64<<<nblk(n + 1), T256, 0, s>>>(M->row_len, l64, n);
  B_TRY(exclusive_scan<uint64_t>(s, l64, ooffs, n + 1));
  if (offs)
    CU_TRY(cudaMemcpy(offs, ooffs, (n + 1) * 8, cudaMemcpyDeviceToHost));
  if (cols && vals && !M->blocks.empty()) {
    // a row of the operator = its pieces in the column ranges, in range order
    std::vector<uint64_t> ho(n + 1), at(n + 1);
    CU_TRY(cudaMemcpy(ho.data(), ooffs, (n + 1) * 8, cudaMemcpyDeviceToHost));
    at = ho;
    for (uint32_t b : M->block_in_col_order) {
      const b200_mat *child = M->blocks[b];
      std::vector<uint64_t> co(n + 1);
      std::vector<uint32_t> cc(child->nnz ? child->nnz : 1);
      std::vector<double> cv(child->nnz ? child->nnz : 1);
      B_TRY(b200_mat_export(child, co.data(), cc.data(), cv.data()));
      for (uint64_t i = 0; i < n; i++)
        for (uint64_t e = co[i]; e < co[i + 1]; e++)
          cols[at[i]] = cc[e], vals[at[i]++] = cv[e];
    }
    for (uint64_t i = 0; i < n; i++)
      if (at[i] != ho[i + 1])
        B_FAIL(B200_EINVAL, "b200_mat_export: column blocks do not add up in row %llu",
               (unsigned long long)i);
  } else if (cols && vals) {
    uint32_t *oc;
    double *ov;
    CU_TRY(cudaMalloc(&oc, (M->nnz ? M->nnz : 1) * 4));
    CU_TRY(cudaMalloc(&ov, (M->nnz ? M->nnz : 1) * 8));
    if (M->sell_slices)
      k_export_sell<<<nblk((uint64_t)M->sell_slices * 32), T256, 0, s>>>(
          M->sell_slices, M->sell_perm, n, M->row_len, M->sell_off,
          (const uint4 *)M->sell_meta, M->sell_cols, M->sell_dcols, M->sell_vals,
          M->sell_vals32, ooffs, oc, ov);
    if (M->vec_rows)
      k_export_vl<<<nblk((uint64_t)M->vec_rows * 32), T256, 0, s>>>(
          M->vec_rows, M->vec_row_ids, M->row_len, M->vec_off, M->vl_cols,
          M->vl_vals, ooffs, oc, ov);
    if (M->long_rows)
      k_export_vl<<<nblk((uint64_t)M->long_rows * 32), T256, 0, s>>>(
          M->long_rows, M->long_row_ids, M->row_len, M->long_off, M->vl_cols,
          M->vl_vals, ooffs, oc, ov);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(cols, oc, M->nnz * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(vals, ov, M->nnz * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(oc), cudaFree(ov);
  }
  cudaFree(l64), cudaFree(ooffs);
  return B200_OK;
}

extern "C" int b200_mat_inv_diag(const b200_mat *M, double *h) {
  if (!M || !h)
    B_FAIL(B200_EINVAL, "b200_mat_inv_diag: null argument");
  CU_TRY(cudaSetDevice(M->ctx->device));
  CU_TRY(cudaMemcpy(h, M->dinv, M->n_local * 8, cudaMemcpyDeviceToHost));
  return B200_OK;
}
