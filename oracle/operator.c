/*
 * operator.c -- which matrix is actually solved.  TEST INFRASTRUCTURE.
 *
 * orc_op_upper_mirror restates src/cholmod-impl.h:5-21: the triplet is
 * allocated with stype = -1 (:6), each row contributes only the entries from
 * the first one with (col - base) >= row to the end of the row (:13-16), and
 * cholmod_l_triplet_to_sparse (:21) turns that into a symmetric matrix, i.e.
 * a_ij = a_ji = file value at (min(i,j), max(i,j)).  Lower-triangle values
 * in the file are never looked at.
 *
 * Deviation, on purpose: the reference scan at :13 has no end-of-row bound;
 * a row with no entry at or right of the diagonal is read as contributing
 * nothing here instead of running into the next row.
 *
 * orc_op_full restates what the other backends are handed (cols - base, all
 * stored entries): src/cusparse.c:55-63, src/amgx.c:41.
 *
 * orc_op_perm_lower_mirror restates what the reference's cuSOLVER backend
 * actually solves, src/cusparse.c:66-99 + :181-197: the matrix is renumbered
 * with the ordering Q (B = Q A Q^T, :66-99) and handed to
 * cusolverSpDcsrlsvchol, a Cholesky solver that reads ONE triangle of what it
 * is given -- the lower one of B.  On a file that is not exactly symmetric
 * that is a_ij = a_ji = the file value at (i, j) where i is the vertex
 * numbered LATER by Q: an operator that depends on the ordering.  Pinned by
 * the reference's own output: tests/golden/cusolver_x.npz holds the x its
 * cusparse_bench returned on a B200 and the Q cuSOLVER produced, and the
 * direct solve of this operator reproduces that x to 1e-13 (the upper-mirror
 * and as-stored operators are 1e-6 ... 3e-5 away).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static orc_op *op_alloc(uint64_t n, uint64_t nnz) {
  orc_op *M = (orc_op *)calloc(1, sizeof(orc_op));
  M->n = n;
  M->offs = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  M->cols = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
  M->vals = (double *)malloc((nnz ? nnz : 1) * sizeof(double));
  return M;
}

void orc_op_free(orc_op *M) {
  if (!M)
    return;
  free(M->offs), free(M->cols), free(M->vals), free(M);
}

uint64_t orc_op_nnz(const orc_op *M) { return M->offs[M->n]; }

orc_op *orc_op_full(const orc_csr *A) {
  uint64_t n = A->nrows, nnz = A->offs[n];
  orc_op *M = op_alloc(n, nnz);
  for (uint64_t i = 0; i <= n; i++)
    M->offs[i] = A->offs[i];
  for (uint64_t k = 0; k < nnz; k++)
    M->cols[k] = A->cols[k] - A->base, M->vals[k] = A->vals[k];
  return M;
}

orc_op *orc_op_upper_mirror(const orc_csr *A) {
  uint64_t n = A->nrows;
  /* first stored entry of each row with col - base >= row */
  uint32_t *ustart = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
  uint64_t *cnt = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++) {
    uint32_t j = A->offs[i], je = A->offs[i + 1];
    while (j < je && (uint64_t)(A->cols[j] - A->base) < i)
      j++;
    ustart[i] = j;
    for (; j < je; j++) {
      uint64_t c = A->cols[j] - A->base;
      cnt[i]++;
      if (c != i && c < n)
        cnt[c]++; /* its mirror image lands in row c */
    }
  }
  uint64_t nnz = 0;
  for (uint64_t i = 0; i < n; i++)
    nnz += cnt[i];
  orc_op *M = op_alloc(n, nnz);
  for (uint64_t i = 0; i < n; i++)
    M->offs[i + 1] = M->offs[i] + cnt[i];

  /* Row r of the result = [mirrors (c < r), ascending source row] followed by
   * [its own upper entries, ascending col].  Sweeping source rows in order
   * fills the mirror part already sorted. */
  uint64_t *fill = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
  memcpy(fill, M->offs, (n + 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++) {
    for (uint32_t j = ustart[i]; j < A->offs[i + 1]; j++) {
      uint64_t c = A->cols[j] - A->base;
      if (c != i && c < n) {
        M->cols[fill[c]] = (uint32_t)i, M->vals[fill[c]] = A->vals[j];
        fill[c]++;
      }
    }
  }
  for (uint64_t i = 0; i < n; i++) {
    for (uint32_t j = ustart[i]; j < A->offs[i + 1]; j++) {
      M->cols[fill[i]] = A->cols[j] - A->base, M->vals[fill[i]] = A->vals[j];
      fill[i]++;
    }
  }
  free(fill), free(cnt), free(ustart);
  return M;
}

/* q[new] = old, as cusolverSpXcsrsym*Host return it (src/cusparse.c:66-85).
 * The result is in the caller's numbering. */
orc_op *orc_op_perm_lower_mirror(const orc_csr *A, const int32_t *q) {
  uint64_t n = A->nrows;
  uint32_t *pos = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint64_t *cnt = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++)
    pos[q[i]] = (uint32_t)i;
  /* entry (i, c) is read iff it lies in the lower triangle of B: pos[i] >= pos[c] */
  for (uint64_t i = 0; i < n; i++)
    for (uint32_t j = A->offs[i]; j < A->offs[i + 1]; j++) {
      uint64_t c = A->cols[j] - A->base;
      if (c < n && pos[i] >= pos[c]) {
        cnt[i]++;
        if (c != i)
          cnt[c]++;
      }
    }
  uint64_t nnz = 0;
  for (uint64_t i = 0; i < n; i++)
    nnz += cnt[i];
  orc_op *M = op_alloc(n, nnz);
  for (uint64_t i = 0; i < n; i++)
    M->offs[i + 1] = M->offs[i] + cnt[i];
  uint64_t *fill = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
  memcpy(fill, M->offs, (n + 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; i++)
    for (uint32_t j = A->offs[i]; j < A->offs[i + 1]; j++) {
      uint64_t c = A->cols[j] - A->base;
      if (c < n && pos[i] >= pos[c]) {
        M->cols[fill[i]] = (uint32_t)c, M->vals[fill[i]++] = A->vals[j];
        if (c != i)
          M->cols[fill[c]] = (uint32_t)i, M->vals[fill[c]++] = A->vals[j];
      }
    }
  /* rows in ascending column order (insertion sort: rows are short) */
  for (uint64_t i = 0; i < n; i++)
    for (uint64_t a = M->offs[i] + 1; a < M->offs[i + 1]; a++) {
      uint32_t ck = M->cols[a];
      double vk = M->vals[a];
      uint64_t b = a;
      for (; b > M->offs[i] && M->cols[b - 1] > ck; b--)
        M->cols[b] = M->cols[b - 1], M->vals[b] = M->vals[b - 1];
      M->cols[b] = ck, M->vals[b] = vk;
    }
  free(fill), free(cnt), free(pos);
  return M;
}

void orc_rhs(uint64_t n, double *b) {
  for (uint64_t i = 0; i < n; i++)
    b[i] = (double)i; /* src/lsbench.c:159-160 */
}
