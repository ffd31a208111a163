/*
 * ldlt.c -- sparse direct solve, the stand-in for the CHOLMOD calls the
 * reference makes.  TEST INFRASTRUCTURE.
 *
 *   src/cholmod-impl.h:25   cholmod_l_analyze    -> ordering + symbolic
 *   src/cholmod-impl.h:26   cholmod_l_factorize  -> numeric
 *   src/cholmod-impl.h:46,60 cholmod_l_solve(CHOLMOD_A, L, r)
 *
 * SuiteSparse v7.0.1 (libs/suitesparse.cmake:9) is not available offline, so
 * this restates the published up-looking sparse LDL^T algorithm (T. A. Davis,
 * "Algorithm 849", ACM TOMS 31(4) 2005: elimination tree from the upper
 * triangle, one sparse triangular solve per row of L) with a reverse
 * Cuthill-McKee ordering in place of CHOLMOD's AMD.  The answer x = A^-1 b is
 * ordering-independent up to rounding; only the timing differs from the real
 * CHOLMOD ("parity unpinned": checked against scipy SuperLU and the analytic
 * I1 answer instead).  Unlike the reference (:46-61) the solution is returned.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

struct orc_ldlt {
  int64_t n;
  int64_t *perm;  /* new -> old */
  int64_t *colp;  /* L column pointers (n+1), strictly-lower part */
  int64_t *rowi;  /* L row indices */
  double *lx;     /* L values */
  double *d;      /* D */
  int bad_pivot;
};

/* ---- reverse Cuthill-McKee --------------------------------------------- */
static int64_t bfs_levels(const orc_op *M, int64_t root, int64_t *queue,
                          int64_t *level, int64_t stamp_base, int64_t *mark,
                          int64_t *last, int64_t *nvisited) {
  /* returns eccentricity; mark[v] == stamp_base means visited this sweep */
  int64_t head = 0, tail = 0, depth = 0;
  queue[tail++] = root, mark[root] = stamp_base, level[root] = 0;
  while (head < tail) {
    int64_t v = queue[head++];
    depth = level[v];
    for (uint64_t k = M->offs[v]; k < M->offs[v + 1]; k++) {
      int64_t w = M->cols[k];
      if (w < (int64_t)M->n && mark[w] != stamp_base)
        mark[w] = stamp_base, level[w] = level[v] + 1, queue[tail++] = w;
    }
  }
  *last = queue[tail - 1], *nvisited = tail;
  return depth;
}

static int64_t *g_deg_for_sort;
static int cmp_by_degree(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  int64_t dx = g_deg_for_sort[x], dy = g_deg_for_sort[y];
  if (dx != dy)
    return dx < dy ? -1 : 1;
  return x < y ? -1 : (x > y);
}

static void rcm_order(const orc_op *M, int64_t *perm) {
  int64_t n = (int64_t)M->n;
  int64_t *deg = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *mark = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *level = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *queue = (int64_t *)malloc(n * sizeof(int64_t));
  char *done = (char *)calloc(n, 1);
  for (int64_t i = 0; i < n; i++)
    deg[i] = (int64_t)(M->offs[i + 1] - M->offs[i]), mark[i] = -1;
  g_deg_for_sort = deg;

  int64_t placed = 0, stamp = 0;
  for (int64_t seed = 0; seed < n; seed++) {
    if (done[seed])
      continue;
    /* pseudo-peripheral start: repeat BFS from the far end while depth grows */
    int64_t root = seed, last, nv, ecc;
    ecc = bfs_levels(M, root, queue, level, stamp++, mark, &last, &nv);
    for (int tries = 0; tries < 8; tries++) {
      int64_t l2, nv2;
      int64_t e2 = bfs_levels(M, last, queue, level, stamp++, mark, &l2, &nv2);
      if (e2 <= ecc)
        break;
      root = last, last = l2, ecc = e2;
    }
    /* Cuthill-McKee from root: BFS, neighbours by increasing degree */
    int64_t head = placed, tail = placed;
    perm[tail++] = root, done[root] = 1;
    while (head < tail) {
      int64_t v = perm[head++], first = tail;
      for (uint64_t k = M->offs[v]; k < M->offs[v + 1]; k++) {
        int64_t w = M->cols[k];
        if (w < n && !done[w])
          done[w] = 1, perm[tail++] = w;
      }
      qsort(perm + first, (size_t)(tail - first), sizeof(int64_t),
            cmp_by_degree);
    }
    placed = tail;
  }
  for (int64_t i = 0, j = n - 1; i < j; i++, j--) {
    int64_t t = perm[i];
    perm[i] = perm[j], perm[j] = t;
  }
  free(deg), free(mark), free(level), free(queue), free(done);
}

/* ---- factorisation ----------------------------------------------------- */
orc_ldlt *orc_ldlt_factor(const orc_op *M, int ordering) {
  int64_t n = (int64_t)M->n;
  orc_ldlt *F = (orc_ldlt *)calloc(1, sizeof(orc_ldlt));
  F->n = n;
  F->perm = (int64_t *)malloc(n * sizeof(int64_t));
  if (ordering == ORC_ORDER_RCM)
    rcm_order(M, F->perm);
  else
    for (int64_t i = 0; i < n; i++)
      F->perm[i] = i;
  int64_t *pinv = (int64_t *)malloc(n * sizeof(int64_t));
  for (int64_t i = 0; i < n; i++)
    pinv[F->perm[i]] = i;

  /* Upper triangle of B = P M P^T by columns.  M is symmetric, so old row
   * perm[k] supplies column k; keep entries with new row index <= k. */
  int64_t *parent = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *stamp = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *cnt = (int64_t *)calloc(n + 1, sizeof(int64_t));

  /* symbolic: elimination tree + column counts of L */
  for (int64_t k = 0; k < n; k++) {
    parent[k] = -1, stamp[k] = k;
    int64_t old = F->perm[k];
    for (uint64_t t = M->offs[old]; t < M->offs[old + 1]; t++) {
      int64_t i = pinv[M->cols[t]];
      if (i >= k)
        continue;
      while (stamp[i] != k) {
        if (parent[i] < 0)
          parent[i] = k;
        cnt[i]++, stamp[i] = k;
        i = parent[i];
      }
    }
  }
  F->colp = (int64_t *)malloc((n + 1) * sizeof(int64_t));
  F->colp[0] = 0;
  for (int64_t k = 0; k < n; k++)
    F->colp[k + 1] = F->colp[k] + cnt[k];
  int64_t lnz = F->colp[n];
  F->rowi = (int64_t *)malloc((lnz ? lnz : 1) * sizeof(int64_t));
  F->lx = (double *)malloc((lnz ? lnz : 1) * sizeof(double));
  F->d = (double *)malloc(n * sizeof(double));

  /* numeric: row k of L from a sparse triangular solve with rows < k */
  double *y = (double *)calloc(n, sizeof(double));
  int64_t *path = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *order = (int64_t *)malloc(n * sizeof(int64_t));
  int64_t *fillc = (int64_t *)calloc(n, sizeof(int64_t));
  for (int64_t k = 0; k < n; k++) {
    int64_t top = n, old = F->perm[k];
    stamp[k] = k;
    double dk = 0.0;
    for (uint64_t t = M->offs[old]; t < M->offs[old + 1]; t++) {
      int64_t i = pinv[M->cols[t]];
      if (i > k)
        continue;
      if (i == k) {
        dk += M->vals[t];
        continue;
      }
      y[i] += M->vals[t];
      int64_t len = 0;
      while (stamp[i] != k)
        path[len++] = i, stamp[i] = k, i = parent[i];
      while (len > 0)
        order[--top] = path[--len];
    }
    for (; top < n; top++) {
      int64_t i = order[top];
      double yi = y[i];
      y[i] = 0.0;
      int64_t pe = F->colp[i] + fillc[i];
      for (int64_t p = F->colp[i]; p < pe; p++)
        y[F->rowi[p]] -= F->lx[p] * yi;
      double lki = yi / F->d[i];
      dk -= lki * yi;
      F->rowi[pe] = k, F->lx[pe] = lki, fillc[i]++;
    }
    F->d[k] = dk;
    if (!(dk > 0.0))
      F->bad_pivot = 1;
  }
  free(y), free(path), free(order), free(fillc);
  free(parent), free(stamp), free(cnt), free(pinv);
  return F;
}

uint64_t orc_ldlt_nnz(const orc_ldlt *F) { return (uint64_t)F->colp[F->n]; }
int orc_ldlt_status(const orc_ldlt *F) { return F->bad_pivot; }

void orc_ldlt_solve(const orc_ldlt *F, const double *b, double *x) {
  int64_t n = F->n;
  double *w = (double *)malloc(n * sizeof(double));
  for (int64_t i = 0; i < n; i++)
    w[i] = b[F->perm[i]];
  for (int64_t j = 0; j < n; j++) { /* L w = Pb */
    double wj = w[j];
    for (int64_t p = F->colp[j]; p < F->colp[j + 1]; p++)
      w[F->rowi[p]] -= F->lx[p] * wj;
  }
  for (int64_t j = 0; j < n; j++)
    w[j] /= F->d[j];
  for (int64_t j = n - 1; j >= 0; j--) { /* L^T w = w */
    double s = w[j];
    for (int64_t p = F->colp[j]; p < F->colp[j + 1]; p++)
      s -= F->lx[p] * w[F->rowi[p]];
    w[j] = s;
  }
  for (int64_t i = 0; i < n; i++)
    x[F->perm[i]] = w[i];
  free(w);
}

void orc_ldlt_free(orc_ldlt *F) {
  if (!F)
    return;
  free(F->perm), free(F->colp), free(F->rowi), free(F->lx), free(F->d);
  free(F);
}
