// generate.cu -- synthetic operators built on the device, one row block per
// rank (BASELINE.json configs 3-5).  The reference has no generator: its
// inputs go through the COO text reader (src/lsbench-csr.c:29-92), whose
// `unsigned nnz` (:35) and qsort (:54) cannot carry 256^3 / 512^3 grids, so
// `struct csr` stays a descriptor and the rows are produced here as plain CSR
// and then pass through the same layout conversion as a file matrix.
//
// Definitions (checked row for row against the CPU specification in tests):
//   poisson7 / poisson27  N^3 grid, row = x + N (y + N z), Dirichlet
//                         truncation, diag 6 / 26, neighbours -1, ascending.
//   powerlaw              L_i = min(Lmax, floor(3 u^(-1/1.2))) from integer
//                         thresholds (host-computed table, so host and device
//                         agree exactly), columns by stratified sampling:
//                         sorted and distinct by construction.
#include "common.cuh"
#include <cmath>
#include <cub/cub.cuh>

#define T256 256
static inline unsigned nblk(uint64_t n, unsigned t = T256) {
  return (unsigned)((n + t - 1) / t);
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash2(uint64_t s, uint64_t a) {
  return mix64(s ^ mix64(a));
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t s, uint64_t a,
                                                   uint64_t b) {
  return mix64(hash2(s, a) ^ mix64(b ^ 0xD1B54A32D192ED03ull));
}

#define PL_LMIN 3u
#define PL_LMAX 65536u

// ---- stencils ----------------------------------------------------------------
__global__ void k_stencil_len(uint32_t N, uint64_t row0, uint64_t nloc, int full27,
                              uint64_t *len) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > nloc)
    return;
  if (i == nloc) {
    len[i] = 0;
    return;
  }
  uint64_t r = row0 + i;
  uint32_t x = r % N, y = (r / N) % N, z = r / ((uint64_t)N * N);
  uint32_t nx = 1 + (x > 0) + (x + 1 < N), ny = 1 + (y > 0) + (y + 1 < N),
           nz = 1 + (z > 0) + (z + 1 < N);
  len[i] = full27 ? (uint64_t)nx * ny * nz : (uint64_t)nx + ny + nz - 2;
}

__global__ void k_stencil_fill(uint32_t N, uint64_t row0, uint64_t nloc, int full27,
                               const uint64_t *offs, uint32_t *cols, double *vals) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nloc)
    return;
  uint64_t r = row0 + i, w = offs[i];
  int x = r % N, y = (r / N) % N, z = r / ((uint64_t)N * N);
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        int nzc = (dz != 0) + (dy != 0) + (dx != 0);
        if (!full27 && nzc > 1)
          continue;
        int xx = x + dx, yy = y + dy, zz = z + dz;
        if (xx < 0 || yy < 0 || zz < 0 || xx >= (int)N || yy >= (int)N || zz >= (int)N)
          continue;
        cols[w] = (uint32_t)(xx + (uint64_t)N * (yy + (uint64_t)N * zz));
        vals[w] = nzc == 0 ? (full27 ? 26.0 : 6.0) : -1.0;
        w++;
      }
}

// ---- power law -----------------------------------------------------------------
__device__ __forceinline__ uint32_t pl_lmax(uint64_t n) {
  uint64_t c = n / 4;
  if (c < PL_LMIN) c = PL_LMIN;
  return (uint32_t)(c < PL_LMAX ? c : PL_LMAX);
}
__device__ __forceinline__ uint64_t pl_half(uint64_t n) {
  uint64_t h = n / 8;
  return h < 4096 ? h : 4096;
}

__global__ void k_pl_len(uint64_t n, uint64_t seed, uint64_t row0, uint64_t nloc,
                         const uint64_t *thr, uint64_t *len) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > nloc)
    return;
  if (i == nloc) {
    len[i] = 0;
    return;
  }
  uint64_t U = (hash2(seed, row0 + i) >> 11) + 1;
  uint32_t lo = PL_LMIN, hi = pl_lmax(n);
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo + 1) / 2;
    if (U <= thr[mid])
      lo = mid;
    else
      hi = mid - 1;
  }
  len[i] = lo;
}

// one warp per row
__global__ void k_pl_fill(uint64_t n, uint64_t seed, uint64_t row0, uint64_t nloc,
                          const uint64_t *offs, uint32_t *cols, double *vals) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (r >= nloc)
    return;
  const uint64_t i = row0 + r, o = offs[r], L = offs[r + 1] - o;
  const uint64_t H = pl_half(n), W = 2 * H + 1, F = n - W;
  uint64_t wlo = i > H ? i - H : 0;
  if (wlo > n - W)
    wlo = n - W;
  uint64_t Ln = (L + 1) / 2;
  if (Ln > W)
    Ln = W;
  const uint64_t Lf = L - Ln;
  uint64_t kf = 0;
  if (Lf) {
    kf = (wlo * Lf) / F;
    while (kf < Lf && ((kf + 1) * F) / Lf <= wlo)
      kf++;
    while (kf > 0 && (kf * F) / Lf > wlo)
      kf--;
    if (kf < Lf) {
      uint64_t lo = (kf * F) / Lf, hi = ((kf + 1) * F) / Lf;
      if (lo < wlo && lo + hash3(seed, i, 2 * kf + 1) % (hi - lo) < wlo)
        kf++;
    }
  }
  for (uint64_t pos = lane; pos < L; pos += 32) {
    uint64_t c;
    if (pos >= kf && pos < kf + Ln) {
      uint64_t k = pos - kf, lo = (k * W) / Ln, hi = ((k + 1) * W) / Ln;
      c = wlo + lo + hash3(seed, i, 2 * k) % (hi - lo);
    } else {
      uint64_t k = pos < kf ? pos : pos - Ln;
      uint64_t lo = (k * F) / Lf, hi = ((k + 1) * F) / Lf;
      uint64_t cp = lo + hash3(seed, i, 2 * k + 1) % (hi - lo);
      c = cp < wlo ? cp : cp + W;
    }
    cols[o + pos] = (uint32_t)c;
    vals[o + pos] = (double)(hash3(seed ^ 0xA5A5A5A5A5A5A5A5ull, i, c) >> 11) *
                        (1.0 / 4503599627370496.0) -
                    1.0;
  }
}

static int scan_u64(cudaStream_t s, const uint64_t *in, uint64_t *out, uint64_t n) {
  void *tmp = nullptr;
  size_t bytes = 0;
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
  CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, s));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaFree(tmp));
  return B200_OK;
}

// Row block owned by `rank`: contiguous, boundaries rounded to 32 rows (and so
// to whole z-planes for the stencils whenever P divides N).
extern "C" int b200_row_block(uint64_t n, int rank, int nranks, uint64_t *r0,
                              uint64_t *r1) {
  if (!r0 || !r1 || nranks < 1 || rank < 0 || rank >= nranks)
    B_FAIL(B200_EINVAL, "b200_row_block: rank %d of %d", rank, nranks);
  auto cut = [&](int k) -> uint64_t {
    if (k <= 0) return 0;
    if (k >= nranks) return n;
    uint64_t c = (uint64_t)((__uint128_t)n * k / nranks);
    c = (c + 16) / 32 * 32;
    return c > n ? n : c;
  };
  *r0 = cut(rank), *r1 = cut(rank + 1);
  return B200_OK;
}

extern "C" int b200_mat_generate(b200_ctx *c, int kind, uint64_t size,
                                 uint64_t seed, uint32_t flags, b200_mat **out) {
  if (!c || !out)
    B_FAIL(B200_EINVAL, "b200_mat_generate: null argument");
  if (flags & B200_MAT_SYM_UPPER)
    B_FAIL(B200_EINVAL, "b200_mat_generate: synthetic operators are symmetric already");
  uint64_t n;
  if (kind == B200_GEN_POISSON7 || kind == B200_GEN_POISSON27) {
    if (size < 2 || size > 1625)  // N^3 must stay below 2^32
      B_FAIL(B200_EINVAL, "b200_mat_generate: grid edge %llu", (unsigned long long)size);
    n = size * size * size;
  } else if (kind == B200_GEN_POWERLAW) {
    if (size < 64 || size >= 0xffffffffull)
      B_FAIL(B200_EINVAL, "b200_mat_generate: powerlaw rows %llu", (unsigned long long)size);
    n = size;
  } else {
    B_FAIL(B200_EINVAL, "b200_mat_generate: unknown kind %d", kind);
  }
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  uint64_t r0, r1;
  b200_row_block(n, c->rank, c->nranks, &r0, &r1);
  const uint64_t nloc = r1 - r0;
  *out = nullptr;

  PlainCsr A;
  A.n = nloc;
  uint64_t *len = nullptr;
  CU_TRY(cudaMalloc(&len, (nloc + 1) * 8));
  CU_TRY(cudaMalloc(&A.offs, (nloc + 1) * 8));
  uint64_t *d_thr = nullptr;
  if (kind == B200_GEN_POWERLAW) {
    // thr[L] = floor((Lmin/L)^a 2^53); same expression as the specification
    uint64_t *thr = (uint64_t *)malloc((PL_LMAX + 1) * 8);
    for (uint32_t L = 0; L <= PL_LMAX; L++)
      thr[L] = L <= PL_LMIN
                   ? (1ull << 53)
                   : (uint64_t)floor(pow((double)PL_LMIN / (double)L, 1.2) *
                                     9007199254740992.0);
    CU_TRY(cudaMalloc(&d_thr, (PL_LMAX + 1) * 8));
    CU_TRY(cudaMemcpyAsync(d_thr, thr, (PL_LMAX + 1) * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));  // pageable source: landed before it is freed
    free(thr);
    k_pl_len<<<nblk(nloc + 1), T256, 0, s>>>(n, seed, r0, nloc, d_thr, len);
  } else {
    k_stencil_len<<<nblk(nloc + 1), T256, 0, s>>>((uint32_t)size, r0, nloc,
                                                  kind == B200_GEN_POISSON27, len);
  }
  CU_TRY(cudaGetLastError());
  B_TRY(scan_u64(s, len, A.offs, nloc + 1));
  CU_TRY(cudaMemcpy(&A.nnz, A.offs + nloc, 8, cudaMemcpyDeviceToHost));
  cudaFree(len);
  CU_TRY(cudaMalloc(&A.cols, (A.nnz ? A.nnz : 1) * 4));
  CU_TRY(cudaMalloc(&A.vals, (A.nnz ? A.nnz : 1) * 8));
  if (kind == B200_GEN_POWERLAW)
    k_pl_fill<<<nblk(nloc * 32), T256, 0, s>>>(n, seed, r0, nloc, A.offs, A.cols, A.vals);
  else
    k_stencil_fill<<<nblk(nloc), T256, 0, s>>>((uint32_t)size, r0, nloc,
                                               kind == B200_GEN_POISSON27, A.offs,
                                               A.cols, A.vals);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  if (d_thr)
    cudaFree(d_thr);

  b200_mat *M = new b200_mat();
  M->ctx = c;
  int rc = partition_and_renumber(c, &A, n, r0, M);
  if (rc == B200_OK)
    rc = build_layout_or_blocks(c, &A, n, r0, flags, &M);
  plain_free(&A);
  if (rc == B200_OK && c->nranks > 1)
    rc = halo_setup(M);
  if (rc != B200_OK) {
    b200_mat_destroy(M);
    return rc;
  }
  *out = M;
  return B200_OK;
}
