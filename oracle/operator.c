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

void orc_rhs(uint64_t n, double *b) {
  for (uint64_t i = 0; i < n; i++)
    b[i] = (double)i; /* src/lsbench.c:159-160 */
}
