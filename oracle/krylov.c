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
 *   exit: if the recurrence met the bar and ||b - A x|| does not (drift after
 *         many iterations), r = b - A x, rz' = r.z, p = z + (rz'/rz) p, go on
 * Summation order differs from the GPU (plain left-to-right here), so
 * agreement is to rounding, not bit for bit; see tests for the tolerances.
 */
#define _POSIX_C_SOURCE 200809L
#include "oracle.h"
#include <math.h>
#include <time.h>
#include <unistd.h>
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

int orc_set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0)
    nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads < 1)
    nthreads = 1;
  omp_set_dynamic(0);
  omp_set_num_threads(nthreads);
  return nthreads;
#else
  (void)nthreads;
  return 1;
#endif
}

/* One Jacobi-PCG iteration's memory traffic and arithmetic on a row slab (see
 * oracle.h): K1 q = A p + p.q, K2 x += a p, r -= a q + r.z, r.r, K3 p = z + b p
 * -- the passes orc_pcg_omp makes, on rows [row0, row0 + M->n) of an operator
 * with n_global columns.  alpha and beta are kept harmless (the slab alone is
 * not a linear system); what is measured is time. */
double orc_pcg_slab_seconds(const orc_op *M, uint64_t n_global, uint64_t row0, int its) {
  const int64_t n = (int64_t)M->n;
  if (row0 + M->n > n_global || its < 1)
    return -1.0;
  double *p = (double *)malloc(n_global * sizeof(double));
  double *v = (double *)malloc(4 * (size_t)n * sizeof(double));
  if (!p || !v) {
    free(p), free(v);
    return -1.0;
  }
  double *q = v, *r = q + n, *x = r + n, *dinv = x + n;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_global; i++)
    p[i] = 1.0 / (double)(1 + (i & 1023));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double a = 1.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      if (M->cols[k] == row0 + (uint64_t)i)
        a = M->vals[k];
    q[i] = 0.0, r[i] = (double)(row0 + i), x[i] = 0.0, dinv[i] = 1.0 / a;
  }
  struct timespec t0, t1;
  double sink = 0.0;
  for (int it = -1; it < its; it++) {
    if (it == 0)
      clock_gettime(CLOCK_MONOTONIC, &t0);
    double pq = 0.0, rz = 0.0, rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < n; i++) {
      double s = 0.0;
      for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
        s += M->vals[k] * p[M->cols[k]];
      q[i] = s;
      pq += s * p[row0 + i];
    }
    const double alpha = 1e-3 / (1.0 + fabs(pq));
#pragma omp parallel for schedule(static) reduction(+ : rz, rr)
    for (int64_t i = 0; i < n; i++) {
      x[i] += alpha * p[row0 + i];
      const double ri = r[i] - alpha * q[i];
      r[i] = ri;
      rz += ri * (dinv[i] * ri), rr += ri * ri;
    }
    const double beta = 0.5;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
      p[row0 + i] = dinv[i] * r[i] + beta * p[row0 + i];
    sink += rz + rr;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(p), free(v);
  return sink == 12345.678 ? -dt : dt;  /* keeps the sums alive */
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
  int it_stop = maxit;                                                         \
  for (int replaced = 0;;) {                                                   \
  while (rc == 1 && it < it_stop) {                                            \
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
  /* residual replacement: the recurrence met the bar; if b - A x does not,    \
   * go on from the true residual with the last direction */                   \
  if (rc == 1 && it < maxit)                                                   \
    rc = 4; /* the bounded tail after a replacement: stagnated */              \
  if (rc != 0 || it == 0 || it >= maxit || replaced >= 4)                      \
    break;                                                                     \
  SPMV(M, x, q);                                                               \
  double trr = 0;                                                              \
  PRAGMA_RED1(trr)                                                             \
  for (int64_t i = 0; i < n; i++)                                              \
    trr += (b[i] - q[i]) * (b[i] - q[i]);                                      \
  if (!(trr > thr * thr))                                                      \
    break;                                                                     \
  replaced++;                                                                  \
  double rzn = 0;                                                              \
  PRAGMA_RED1(rzn)                                                             \
  for (int64_t i = 0; i < n; i++) {                                            \
    r[i] = b[i] - q[i];                                                        \
    rzn += r[i] * (dinv[i] * r[i]);                                            \
  }                                                                            \
  rr = trr;                                                                    \
  double beta = rzn / rz;                                                      \
  rz = rzn;                                                                    \
  PRAGMA_FOR                                                                   \
  for (int64_t i = 0; i < n; i++)                                              \
    p[i] = dinv[i] * r[i] + beta * p[i];                                       \
  rc = 1;                                                                      \
  it_stop = it + (it / 8 > 8 ? it / 8 : 8);                                    \
  it_stop = it_stop < maxit ? it_stop : maxit;                                 \
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

/* ---- fp32-stored operator + fp64 refinement --------------
 * What b200_pcg_solve does on a B200_MAT_VALUES_F32 matrix whose values do not
 * all survive the rounding to fp32: A32 = fl32(A); repeat { r = b - A x in
 * fp64 with the fp64 values; stop if ||r|| <= tol ||b||; solve A32 d = r by
 * Jacobi-PCG (fp64 vectors) to max(eta, tol ||b|| / (2 ||r||)); x += d }.
 * When every value survives, A32 == A and this is orc_pcg. */
int orc_pcg_refine32(const orc_op *M, const double *b, double *x, double tol,
                     int maxit, double eta, int *iters, int *outer, double *relres) {
  int64_t n = (int64_t)M->n;
  uint64_t nnz = M->offs[n];
  orc_op M32 = *M;
  M32.vals = (double *)malloc((nnz ? nnz : 1) * sizeof(double));
  int exact = 1;
  for (uint64_t k = 0; k < nnz; k++) {
    M32.vals[k] = (double)(float)M->vals[k];
    exact &= M32.vals[k] == M->vals[k];
  }
  if (exact) { /* lossless storage: the product runs the plain solve */
    free(M32.vals);
    if (outer)
      *outer = 0;
    return orc_pcg(M, b, x, tol, maxit, iters, relres);
  }
  double *r = (double *)malloc(2 * (size_t)n * sizeof(double) + 8), *d = r + n;
  double bb = 0, rr = 0;
  for (int64_t i = 0; i < n; i++)
    bb += b[i] * b[i];
  int total = 0, passes = 0, rc = 1;
  for (;;) {
    orc_spmv(M, x, r, NULL);
    rr = 0;
    for (int64_t i = 0; i < n; i++) {
      r[i] = b[i] - r[i];
      rr += r[i] * r[i];
    }
    if (rr <= tol * tol * bb) {
      rc = 0;
      break;
    }
    if (total >= maxit || passes >= 40)
      break;
    double t = 0.5 * tol * sqrt(bb / rr);
    if (t < eta)
      t = eta;
    memset(d, 0, (size_t)n * sizeof(double));
    int it = 0;
    int irc = orc_pcg(&M32, r, d, t, maxit - total, &it, NULL);
    total += it, passes++;
    if (irc == 2) {
      rc = 2;
      break;
    }
    for (int64_t i = 0; i < n; i++)
      x[i] += d[i];
  }
  if (iters)
    *iters = total;
  if (outer)
    *outer = passes;
  if (relres)
    *relres = bb > 0 ? sqrt(rr / bb) : sqrt(rr);
  free(M32.vals), free(r);
  return rc;
}

/* ---- Chebyshev-Jacobi preconditioned CG ---------------------------
 * The reference reaches for algebraic multigrid on these systems (src/hypre.c:126-188
 * BoomerAMG, src/amgx.c:78-85); a polynomial in D^-1 A is the preconditioner of that
 * family that needs nothing but the product the solver already has.  Chebyshev
 * iteration (Saad, Iterative Methods, Alg. 12.1) for B z = c with B = D^-1 A, c = D^-1 r,
 * z0 = 0, eigenvalues of B assumed in [a, b]:
 *   theta = (b + a) / 2, delta = (b - a) / 2, sigma = theta / delta, rho = 1 / sigma
 *   rh = c; d = rh / theta; z = d
 *   repeat degree - 1 times:  rh -= B d;  rho' = 1 / (2 sigma - rho);
 *                             d = rho' rho d + (2 rho' / delta) rh;  z += d;  rho = rho'
 * z = P(B) c with P a fixed polynomial, so the preconditioner is a fixed SPD operator
 * as long as b bounds the spectrum, and CG theory applies unchanged. */
double orc_cheb_lmax(const orc_op *M) {
  const int64_t n = (int64_t)M->n;
  double *dinv = inv_diag(M);
  double gersh = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double s = 0.0;
    for (uint64_t k = M->offs[i]; k < M->offs[i + 1]; k++)
      s += fabs(M->vals[k]);
    s *= fabs(dinv[i]);
    gersh = s > gersh ? s : gersh;
  }
  double *v = (double *)malloc(2 * (size_t)n * sizeof(double)), *w = v + n;
  for (int64_t i = 0; i < n; i++)
    v[i] = 1.0;
  double lam = 0.0;
  for (int it = 0; it < 40; it++) {
    orc_spmv(M, v, w, NULL);
    double nw = 0.0, nv = 0.0;
    for (int64_t i = 0; i < n; i++) {
      w[i] *= dinv[i];
      nw += w[i] * w[i], nv += v[i] * v[i];
    }
    lam = sqrt(nw / nv);
    const double s = 1.0 / sqrt(nw);
    for (int64_t i = 0; i < n; i++)
      v[i] = w[i] * s;
  }
  free(v), free(dinv);
  const double est = 1.15 * lam;
  return est < gersh ? est : gersh;
}

int orc_pcg_cheb(const orc_op *M, const double *b, double *x, double tol, int maxit,
                 int degree, double lmax, double ratio, int *iters, double *relres) {
  const int64_t n = (int64_t)M->n;
  double *dinv = inv_diag(M);
  double *r = (double *)malloc(7 * (size_t)n * sizeof(double));
  double *p = r + n, *q = p + n, *z = q + n, *rh = z + n, *d = rh + n, *t = d + n;
  const double lb = lmax, la = lmax / ratio;
  const double theta = 0.5 * (lb + la), delta = 0.5 * (lb - la), sigma = theta / delta;
#define ORC_CHEB_APPLY()                                                       \
  do {                                                                         \
    double rho = 1.0 / sigma;                                                  \
    for (int64_t i = 0; i < n; i++) {                                          \
      rh[i] = dinv[i] * r[i];                                                  \
      d[i] = rh[i] / theta;                                                    \
      z[i] = d[i];                                                             \
    }                                                                          \
    for (int j = 1; j < degree; j++) {                                         \
      orc_spmv(M, d, t, NULL);                                                 \
      const double rhon = 1.0 / (2.0 * sigma - rho);                           \
      for (int64_t i = 0; i < n; i++) {                                        \
        rh[i] -= dinv[i] * t[i];                                               \
        d[i] = rhon * rho * d[i] + (2.0 * rhon / delta) * rh[i];               \
        z[i] += d[i];                                                          \
      }                                                                        \
      rho = rhon;                                                              \
    }                                                                          \
  } while (0)
  double bb = 0, rz = 0, rr = 0;
  orc_spmv(M, x, q, NULL);
  for (int64_t i = 0; i < n; i++) {
    r[i] = b[i] - q[i];
    bb += b[i] * b[i], rr += r[i] * r[i];
  }
  ORC_CHEB_APPLY();
  for (int64_t i = 0; i < n; i++)
    p[i] = z[i], rz += r[i] * z[i];
  const double bnorm = sqrt(bb), thr = tol * bnorm;
  int it = 0, rc = 1;
  if (sqrt(rr) <= thr)
    rc = 0;
  while (rc == 1 && it < maxit) {
    double pq = 0;
    orc_spmv(M, p, q, NULL);
    for (int64_t i = 0; i < n; i++)
      pq += p[i] * q[i];
    if (!(pq > 0.0)) {
      rc = 2;
      break;
    }
    const double alpha = rz / pq;
    rr = 0;
    for (int64_t i = 0; i < n; i++) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rr += r[i] * r[i];
    }
    it++;
    if (sqrt(rr) <= thr) {
      rc = 0;
      break;
    }
    ORC_CHEB_APPLY();
    double rzn = 0;
    for (int64_t i = 0; i < n; i++)
      rzn += r[i] * z[i];
    const double beta = rzn / rz;
    rz = rzn;
    for (int64_t i = 0; i < n; i++)
      p[i] = z[i] + beta * p[i];
  }
#undef ORC_CHEB_APPLY
  if (iters)
    *iters = it;
  if (relres)
    *relres = bnorm > 0 ? sqrt(rr) / bnorm : sqrt(rr);
  free(r), free(dinv);
  return rc;
}

/* ---- block-Jacobi PCG -----------------------------------------------
 * z = B^-1 r with B the diagonal blocks of A over a given partition of the rows
 * (block_of_row[i] = block id of row i; ids are arbitrary, a block may be any subset).
 * A block is symmetrised, inverted through its Cholesky factor and -- as the product
 * keeps it in shared memory -- cut to the high 32 bits of every fp64 entry (rounded to
 * nearest); everything else is fp64.  The method
 * csrc/small.cu runs with B200_PCG_BLOCK_JACOBI, on the partition
 * b200_mat_block_jacobi_partition reports.  What the reference reaches for on these
 * systems is algebraic multigrid (src/hypre.c:126-188, src/amgx.c:78-85).  Same
 * contract as orc_pcg; rc 3 when a block is not positive definite. */
typedef struct {
  uint32_t m;     /* rows */
  uint32_t *rows; /* ascending */
  double *inv;    /* m x m, symmetric: fp64 values cut to their high 32 bits */
} bj_block;

/* what the product keeps of an entry of an inverted block: the high word of the fp64
 * value, rounded to nearest (20 bits of mantissa; widening it back costs nothing) */
static double high_word(double v) {
  uint64_t bits;
  memcpy(&bits, &v, 8);
  bits = (bits + 0x80000000ull) & 0xffffffff00000000ull;
  memcpy(&v, &bits, 8);
  return v;
}

static int cmp_u64(const void *a, const void *b) {
  const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

int orc_pcg_bj(const orc_op *M, const double *b, double *x, double tol, int maxit,
               const uint32_t *block_of_row, int *iters, double *relres) {
  const int64_t n = (int64_t)M->n;
  /* rows grouped by block id: sort (id, row) pairs */
  uint64_t *key = (uint64_t *)malloc((size_t)(n ? n : 1) * sizeof(uint64_t));
  for (int64_t i = 0; i < n; i++)
    key[i] = ((uint64_t)block_of_row[i] << 32) | (uint64_t)i;
  qsort(key, (size_t)n, sizeof(uint64_t), cmp_u64);
  int64_t nblocks = 0;
  for (int64_t i = 0; i < n; i++)
    nblocks += i == 0 || (key[i] >> 32) != (key[i - 1] >> 32);
  bj_block *blk = (bj_block *)calloc((size_t)(nblocks ? nblocks : 1), sizeof(bj_block));
  uint32_t *pos = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t)); /* row -> index in its block */
  int rc = 1;
  int64_t k = -1;
  for (int64_t i = 0; i < n;) {
    int64_t j = i;
    while (j < n && (key[j] >> 32) == (key[i] >> 32))
      j++;
    bj_block *B = &blk[++k];
    B->m = (uint32_t)(j - i);
    B->rows = (uint32_t *)malloc(B->m * sizeof(uint32_t));
    for (int64_t t = i; t < j; t++)
      B->rows[t - i] = (uint32_t)(key[t] & 0xffffffffu), pos[B->rows[t - i]] = (uint32_t)(t - i);
    i = j;
  }
  for (k = 0; k < nblocks && rc == 1; k++) {
    bj_block *B = &blk[k];
    const uint32_t m = B->m;
    double *D = (double *)calloc((size_t)m * m * 3, sizeof(double)), *L = D + (size_t)m * m, *Li = L + (size_t)m * m;
    for (uint32_t a = 0; a < m; a++) {
      const uint32_t r = B->rows[a];
      for (uint64_t e = M->offs[r]; e < M->offs[r + 1]; e++) {
        const uint32_t c = M->cols[e];
        if (c < (uint64_t)n && block_of_row[c] == block_of_row[r])
          D[(size_t)a * m + pos[c]] += M->vals[e];
      }
    }
    for (uint32_t a = 0; a < m; a++)
      for (uint32_t q = 0; q < a; q++) {
        const double v = 0.5 * (D[(size_t)a * m + q] + D[(size_t)q * m + a]);
        D[(size_t)a * m + q] = D[(size_t)q * m + a] = v;
      }
    for (uint32_t a = 0; a < m && rc == 1; a++) /* D = L L^T */
      for (uint32_t q = 0; q <= a; q++) {
        double v = D[(size_t)a * m + q];
        for (uint32_t t = 0; t < q; t++)
          v -= L[(size_t)a * m + t] * L[(size_t)q * m + t];
        if (a == q) {
          if (!(v > 0.0)) {
            rc = 3;
            break;
          }
          L[(size_t)a * m + a] = sqrt(v);
        } else {
          L[(size_t)a * m + q] = v / L[(size_t)q * m + q];
        }
      }
    B->inv = (double *)calloc((size_t)m * m, sizeof(double));
    if (rc == 1) {
      for (uint32_t q = 0; q < m; q++) { /* Li = L^-1 */
        Li[(size_t)q * m + q] = 1.0 / L[(size_t)q * m + q];
        for (uint32_t a = q + 1; a < m; a++) {
          double v = 0.0;
          for (uint32_t t = q; t < a; t++)
            v -= L[(size_t)a * m + t] * Li[(size_t)t * m + q];
          Li[(size_t)a * m + q] = v / L[(size_t)a * m + a];
        }
      }
      for (uint32_t a = 0; a < m; a++) /* D^-1 = Li^T Li, rounded to fp32 */
        for (uint32_t q = 0; q <= a; q++) {
          double v = 0.0;
          for (uint32_t t = a; t < m; t++)
            v += Li[(size_t)t * m + a] * Li[(size_t)t * m + q];
          B->inv[(size_t)a * m + q] = B->inv[(size_t)q * m + a] = high_word(v);
        }
    }
    free(D);
  }
  double *r = (double *)malloc(4 * (size_t)(n ? n : 1) * sizeof(double));
  double *p = r + n, *q = p + n, *z = q + n;
#define ORC_BJ_APPLY()                                                         \
  for (int64_t kb = 0; kb < nblocks; kb++) {                                   \
    const bj_block *B = &blk[kb];                                              \
    for (uint32_t a = 0; a < B->m; a++) {                                      \
      double s = 0.0;                                                          \
      for (uint32_t c = 0; c < B->m; c++)                                      \
        s += B->inv[(size_t)a * B->m + c] * r[B->rows[c]];                     \
      z[B->rows[a]] = s;                                                       \
    }                                                                          \
  }
  double bb = 0, rz = 0, rr = 0;
  int it = 0;
  if (rc == 1) {
    orc_spmv(M, x, q, NULL);
    for (int64_t i = 0; i < n; i++) {
      r[i] = b[i] - q[i];
      bb += b[i] * b[i], rr += r[i] * r[i];
    }
    ORC_BJ_APPLY();
    for (int64_t i = 0; i < n; i++)
      p[i] = z[i], rz += r[i] * z[i];
  }
  const double bnorm = sqrt(bb), thr = tol * bnorm;
  if (rc == 1 && sqrt(rr) <= thr)
    rc = 0;
  while (rc == 1 && it < maxit) {
    double pq = 0;
    orc_spmv(M, p, q, NULL);
    for (int64_t i = 0; i < n; i++)
      pq += p[i] * q[i];
    if (!(pq > 0.0)) {
      rc = 2;
      break;
    }
    const double alpha = rz / pq;
    rr = 0;
    for (int64_t i = 0; i < n; i++) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rr += r[i] * r[i];
    }
    it++;
    if (sqrt(rr) <= thr) {
      rc = 0;
      break;
    }
    ORC_BJ_APPLY();
    double rzn = 0;
    for (int64_t i = 0; i < n; i++)
      rzn += r[i] * z[i];
    const double beta = rzn / rz;
    rz = rzn;
    for (int64_t i = 0; i < n; i++)
      p[i] = z[i] + beta * p[i];
  }
#undef ORC_BJ_APPLY
  if (iters)
    *iters = it;
  if (relres)
    *relres = bnorm > 0 ? sqrt(rr) / bnorm : sqrt(rr);
  for (k = 0; k < nblocks; k++)
    free(blk[k].rows), free(blk[k].inv);
  free(blk), free(pos), free(key), free(r);
  return rc;
}
