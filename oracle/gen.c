/*
 * gen.c -- CPU definitions of the synthetic operators of BASELINE.json
 * configs 3-5.  TEST INFRASTRUCTURE.  The reference has no generators (its
 * inputs are the COO text files under tests/); these definitions are the
 * specification the device generators are checked against, row for row.
 *
 * poisson7  : N^3 grid, row = x + N*(y + N*z), Dirichlet truncation,
 *             diag 6, the six face neighbours -1, columns ascending.
 * poisson27 : same grid, all 26 neighbours -1, diag 26 (HPCG-style).
 * powerlaw  : row length L_i = min(Lmax, floor(Lmin * u^(-1/a))), a = 1.2,
 *             Lmin = 3, Lmax = min(65536, n/4); half the entries fall in a
 *             window of +-H around the diagonal (H = min(4096, n/8)), the
 *             other half anywhere else.  Columns are produced already sorted
 *             and distinct by stratified sampling, so host and device agree
 *             bit for bit with no sort: see powerlaw_entry().
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

/* ---- counter-based hash (Steele et al. splitmix64 finaliser) ----------- */
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t hash2(uint64_t s, uint64_t a) {
  return mix64(s ^ mix64(a));
}
static inline uint64_t hash3(uint64_t s, uint64_t a, uint64_t b) {
  return mix64(hash2(s, a) ^ mix64(b ^ 0xD1B54A32D192ED03ull));
}

#define PL_LMIN 3u
#define PL_LMAX 65536u
#define PL_ALPHA 1.2

/* thr[L] = floor((Lmin/L)^a * 2^53) for L in [Lmin, Lmax]; L_i is the
 * largest L with U <= thr[L], U = (hash >> 11) + 1 in [1, 2^53]. */
void orc_powerlaw_table(uint64_t *thr) {
  for (uint32_t L = 0; L <= PL_LMAX; L++) {
    if (L <= PL_LMIN)
      thr[L] = 1ull << 53;
    else
      thr[L] = (uint64_t)floor(pow((double)PL_LMIN / (double)L, PL_ALPHA) *
                               9007199254740992.0);
  }
}

static uint64_t g_thr[PL_LMAX + 1];
static int g_thr_ready;

static uint32_t pl_lmax(uint64_t n) {
  uint64_t c = n / 4;
  if (c < PL_LMIN)
    c = PL_LMIN;
  return (uint32_t)(c < PL_LMAX ? c : PL_LMAX);
}
static uint64_t pl_half(uint64_t n) {
  uint64_t h = n / 8;
  return h < 4096 ? h : 4096;
}

uint32_t orc_powerlaw_rowlen(uint64_t n, uint64_t seed, uint64_t row) {
  if (!g_thr_ready)
    orc_powerlaw_table(g_thr), g_thr_ready = 1;
  uint64_t U = (hash2(seed, row) >> 11) + 1;
  uint32_t lo = PL_LMIN, hi = pl_lmax(n); /* thr[lo] >= U always */
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo + 1) / 2;
    if (U <= g_thr[mid])
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

/* Entry `pos` (0-based, ascending column) of row i, which has L entries. */
static void powerlaw_entry(uint64_t n, uint64_t seed, uint64_t i, uint32_t L,
                           uint32_t pos, uint32_t *col, double *val) {
  uint64_t H = pl_half(n), W = 2 * H + 1, F = n - W;
  uint64_t wlo = i > H ? i - H : 0;
  if (wlo > n - W)
    wlo = n - W;
  uint64_t Ln = (L + 1) / 2;
  if (Ln > W)
    Ln = W; /* at most one pick per window column */
  uint64_t Lf = L - Ln;
  /* far picks below the window: strata entirely below wlo, plus the one that
   * straddles it if its pick happens to fall below. */
  uint64_t kf = 0;
  if (Lf) {
    kf = (wlo * Lf) / F; /* strata 0..kf-1 have hi <= wlo, up to rounding */
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
  *col = (uint32_t)c;
  /* U(-1,1), exactly representable: 53-bit integer * 2^-52 - 1 */
  *val = (double)(hash3(seed ^ 0xA5A5A5A5A5A5A5A5ull, i, c) >> 11) *
             (1.0 / 4503599627370496.0) -
         1.0;
}

static orc_op *alloc_rows(uint64_t nloc, uint64_t nnz) {
  orc_op *M = (orc_op *)calloc(1, sizeof(orc_op));
  M->n = nloc;
  M->offs = (uint64_t *)calloc(nloc + 1, sizeof(uint64_t));
  M->cols = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
  M->vals = (double *)malloc((nnz ? nnz : 1) * sizeof(double));
  return M;
}

orc_op *orc_gen_powerlaw(uint64_t n, uint64_t seed, uint64_t row0,
                         uint64_t row1) {
  uint64_t nloc = row1 - row0, nnz = 0;
  uint32_t *len = (uint32_t *)malloc((nloc ? nloc : 1) * sizeof(uint32_t));
  for (uint64_t r = 0; r < nloc; r++)
    nnz += (len[r] = orc_powerlaw_rowlen(n, seed, row0 + r));
  orc_op *M = alloc_rows(nloc, nnz);
  for (uint64_t r = 0; r < nloc; r++)
    M->offs[r + 1] = M->offs[r] + len[r];
  for (uint64_t r = 0; r < nloc; r++)
    for (uint32_t k = 0; k < len[r]; k++)
      powerlaw_entry(n, seed, row0 + r, len[r], k, &M->cols[M->offs[r] + k],
                     &M->vals[M->offs[r] + k]);
  free(len);
  return M;
}

static orc_op *gen_stencil(uint32_t N, uint64_t row0, uint64_t row1,
                           int full27) {
  uint64_t nloc = row1 - row0;
  uint64_t cap = nloc * (full27 ? 27 : 7);
  orc_op *M = alloc_rows(nloc, cap);
  uint64_t w = 0;
  for (uint64_t r = row0; r < row1; r++) {
    int64_t x = (int64_t)(r % N), y = (int64_t)((r / N) % N),
            z = (int64_t)(r / ((uint64_t)N * N));
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          int nz = (dz != 0) + (dy != 0) + (dx != 0);
          if (!full27 && nz > 1)
            continue;
          int64_t xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 0 || yy < 0 || zz < 0 || xx >= N || yy >= N || zz >= N)
            continue;
          M->cols[w] = (uint32_t)(xx + (int64_t)N * (yy + (int64_t)N * zz));
          M->vals[w] = nz == 0 ? (full27 ? 26.0 : 6.0) : -1.0;
          w++;
        }
    M->offs[r - row0 + 1] = w;
  }
  return M;
}

orc_op *orc_gen_poisson7(uint32_t N, uint64_t row0, uint64_t row1) {
  return gen_stencil(N, row0, row1, 0);
}
orc_op *orc_gen_poisson27(uint32_t N, uint64_t row0, uint64_t row1) {
  return gen_stencil(N, row0, row1, 1);
}
