/*
 * krylov.c -- host CSR SpMV and Jacobi-preconditioned CG.  TEST
 * INFRASTRUCTURE.  The reference has no SpMV / CG of its own (its only
 * Krylov precedent is the Ginkgo wiring, src/ginkgo.cpp:55-64: Jacobi
 * preconditioner, relative residual stop, x reset per trial :67,92); the
 * protocol around it is src/lsbench.c:157-160 (b[i] = i, x0 = 0).
 *
 * The recurrences are the textbook ones the CUDA path implements:
 *   r = b - A x0; z = D^-1 r; p = z; rz = r.z
 *   loop: q = A p; alpha = rz / p.q; x += alpha p; r -= alpha q;
 *         z = D^-1 r; rz' = r.z; stop if ||r|| <= tol ||b||;
 *         beta = rz'/rz; p = z + beta p
 * Summation order differs from the GPU (plain left-to-right here), so
 * agreement is to rounding, not bit for bit; see tests for the tolerances.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_spmv(const orc_op *M, const double *x, double *y, double *yabs) {
  for (uint64_t i = 0; i < M->n; i++) {
    double s = 0.0, sa = 0.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++) {
      double t = M->vals[k] * x[M->cols[k]];
      s += t, sa += fabs(t);
    }
    y[i] = s;
    if (yabs)
      yabs[i] = sa;
  }
}

/* Same product with every multiply-add fused (one rounding per term), summed
 * left to right: the arithmetic a GPU lane does in the SELL kernel, so that
 * path can be compared bit for bit. */
void orc_spmv_fma(const orc_op *M, const double *x, double *y) {
  for (uint64_t i = 0; i < M->n; i++) {
    double s = 0.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      s = fma(M->vals[k], x[M->cols[k]], s);
    y[i] = s;
  }
}

void orc_spmv_omp(const orc_op *M, const double *x, double *y) {
  int64_t n = (int64_t)M->n;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double s = 0.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      s += M->vals[k] * x[M->cols[k]];
    y[i] = s;
  }
}

static double *inv_diag(const orc_op *M) {
  double *d = (double *)malloc(M->n * sizeof(double));
  for (uint64_t i = 0; i < M->n; i++) {
    double a = 0.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      if (M->cols[k] == i)
        a = M->vals[k];
    d[i] = a != 0.0 ? 1.0 / a : 1.0;
  }
  return d;
}

double orc_true_relres(const orc_op *M, const double *b, const double *x) {
  long double rr = 0, bb = 0;
  for (uint64_t i = 0; i < M->n; i++) {
    long double s = b[i];
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      s -= (long double)M->vals[k] * x[M->cols[k]];
    rr += s * s, bb += (long double)b[i] * b[i];
  }
  return bb > 0 ? (double)sqrtl(rr / bb) : (double)sqrtl(rr);
}

#define PCG_BODY(PRAGMA_FOR, PRAGMA_RED1, PRAGMA_RED2, SPMV)                                \
  int64_t n = (int64_t)M->n;                                                   \
  double *dinv = inv_diag(M);                                                  \
  double *r = (double *)malloc(4 * (size_t)n * sizeof(double));                \
  double *p = r + n, *q = p + n, *z = q + n;                                   \
  (void)z;                                                                     \
  double bb = 0, rz = 0, rr = 0;                                               \
  SPMV(M, x, q);                                                               \
  PRAGMA_RED2(bb, rz)                                                          \
  for (int64_t i = 0; i < n; i++) {                                            \
    r[i] = b[i] - q[i];                                                        \
    p[i] = dinv[i] * r[i];                                                     \
    bb += b[i] * b[i], rz += r[i] * p[i];                                      \
  }                                                                            \
  PRAGMA_RED1(rr)                                                              \
  for (int64_t i = 0; i < n; i++)                                              \
    rr += r[i] * r[i];                                                         \
  double bnorm = sqrt(bb), thr = tol * bnorm;                                  \
  int it = 0, rc = 1;                                                          \
  if (sqrt(rr) <= thr)                                                         \
    rc = 0;                                                                    \
  while (rc == 1 && it < maxit) {                                              \
    double pq = 0;                                                             \
    SPMV(M, p, q);                                                             \
    PRAGMA_RED1(pq)                                                            \
    for (int64_t i = 0; i < n; i++)                                            \
      pq += p[i] * q[i];                                                       \
    if (!(pq > 0.0)) {                                                         \
      rc = 2;                                                                  \
      break;                                                                   \
    }                                                                          \
    double alpha = rz / pq, rzn = 0;                                           \
    rr = 0;                                                                    \
    PRAGMA_RED2(rzn, rr)                                                       \
    for (int64_t i = 0; i < n; i++) {                                          \
      x[i] += alpha * p[i];                                                    \
      double ri = r[i] - alpha * q[i];                                         \
      r[i] = ri;                                                               \
      rzn += ri * (dinv[i] * ri), rr += ri * ri;                               \
    }                                                                          \
    it++;                                                                      \
    if (sqrt(rr) <= thr) {                                                     \
      rc = 0;                                                                  \
      break;                                                                   \
    }                                                                          \
    double beta = rzn / rz;                                                    \
    rz = rzn;                                                                  \
    PRAGMA_FOR                                                                 \
    for (int64_t i = 0; i < n; i++)                                            \
      p[i] = dinv[i] * r[i] + beta * p[i];                                     \
  }                                                                            \
  if (iters)                                                                   \
    *iters = it;                                                               \
  if (relres)                                                                  \
    *relres = bnorm > 0 ? sqrt(rr) / bnorm : sqrt(rr);                         \
  free(r), free(dinv);                                                         \
  return rc;

static void spmv_serial(const orc_op *M, const double *x, double *y) {
  orc_spmv(M, x, y, NULL);
}

#define NOP_FOR
#define NOP_RED1(a)
#define NOP_RED2(a, b)
int orc_pcg(const orc_op *M, const double *b, double *x, double tol,
            int maxit, int *iters, double *relres) {
  PCG_BODY(NOP_FOR, NOP_RED1, NOP_RED2, spmv_serial)
}

#define DO_PRAGMA(s) _Pragma(#s)
#define OMP_FOR DO_PRAGMA(omp parallel for schedule(static))
#define OMP_RED1(a) DO_PRAGMA(omp parallel for schedule(static) reduction(+ : a))
#define OMP_RED2(a, b)                                                         \
  DO_PRAGMA(omp parallel for schedule(static) reduction(+ : a, b))
int orc_pcg_omp(const orc_op *M, const double *b, double *x, double tol,
                int maxit, int *iters, double *relres) {
  PCG_BODY(OMP_FOR, OMP_RED1, OMP_RED2, orc_spmv_omp)
}
