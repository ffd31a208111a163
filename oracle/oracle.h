/*
 * oracle.h -- CPU restatement of the lsbench hot path (TEST INFRASTRUCTURE).
 *
 * This directory is the checker, never the product: only tests/, the smoke()
 * entry and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under lsbench_b200/ links, imports or executes anything from here.
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths are relative to the lsbench reference tree).  Pure C11, no
 * third-party code.
 *
 * Parity status: the reference ships no golden vectors for this path
 * (tests/ holds inputs only) and its CHOLMOD solve cannot be built here
 * (SuiteSparse v7.0.1 is fetched at configure time).  What IS pinned:
 *   - orc_matrix_read against the reference's own lsbench_matrix_read,
 *     compiled from the reference sources into oracle/_ref (bit-exact CSR);
 *   - the solve against THE REFERENCE'S OWN OUTPUT: x as its cuSOLVER backend
 *     (src/cusparse.c, compiled from the reference sources by `make
 *     ref-cusolver`, run on a B200) returned it for the seven Nek matrices
 *     (tests/golden/cusolver_x.npz; orc_op_perm_lower_mirror is the operator
 *     that backend solves), reproduced by the direct solve here to 1.5e-13;
 *   - the direct solve against scipy SuperLU on CHOLMOD's operator and the
 *     analytic answer for I1_05x05 (tests/golden/direct.npz).
 * Only the CHOLMOD backend's own arithmetic stays "parity unpinned".
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Host CSR as the reference builds it: src/lsbench-impl.h:22-26.
 * offs are 0-based, cols keep the file's base. */
typedef struct {
  uint32_t nrows, base;
  uint32_t *offs, *cols;
  double *vals;
} orc_csr;

/* The operator a solver sees: square, 0-based, columns ascending per row,
 * full (both triangles) storage.  64-bit offsets so 27-pt 512^3 fits. */
typedef struct {
  uint64_t n;
  uint64_t *offs;
  uint32_t *cols;
  double *vals;
} orc_op;

/* ---- ingest: src/lsbench-csr.c:29-92 ---------------------------------- */
orc_csr *orc_matrix_read(const char *fname);
void orc_matrix_free(orc_csr *A);

/* ---- operator definitions --------------------------------------------- */
/* src/cholmod-impl.h:5-21: entries with col-base >= row, mirrored. */
orc_op *orc_op_upper_mirror(const orc_csr *A);
/* As stored (what cusolver / ginkgo / amgx are handed): cols - base. */
orc_op *orc_op_full(const orc_csr *A);
/* what src/cusparse.c solves: lower triangle of Q A Q^T mirrored; q[new] = old */
orc_op *orc_op_perm_lower_mirror(const orc_csr *A, const int32_t *q);
void orc_op_free(orc_op *M);
uint64_t orc_op_nnz(const orc_op *M);

/* ---- right-hand side: src/lsbench.c:157-160 (r[i] = i, x = 0) --------- */
void orc_rhs(uint64_t n, double *b);

/* ---- SpMV (host CSR product in row order) ------------------ */
/* y = M x; if yabs != NULL also yabs[i] = sum_j |a_ij x_j| (error scale). */
void orc_spmv(const orc_op *M, const double *x, double *y, double *yabs);
void orc_spmv_fma(const orc_op *M, const double *x, double *y);
void orc_spmv_omp(const orc_op *M, const double *x, double *y);

/* ---- Jacobi-preconditioned CG ----------------------------------------- */
/* x holds x0 on entry.  Stops when ||r||_2 <= tol*||b||_2 (recurrence r).
 * Returns 0 converged, 1 maxit reached, 2 breakdown (p.Ap <= 0 or NaN). */
int orc_pcg(const orc_op *M, const double *b, double *x, double tol,
            int maxit, int *iters, double *relres);
int orc_pcg_omp(const orc_op *M, const double *b, double *x, double tol,
                int maxit, int *iters, double *relres);
/* SURVEY 8(f) row 2, the preconditioner half: Chebyshev-Jacobi.  z = P(D^-1 A) D^-1 r
 * with P the degree-(k-1) Chebyshev polynomial for the interval [lmax / ratio, lmax]
 * of D^-1 A (k = 1 is plain Jacobi): k - 1 more products per iteration, no more
 * reductions, fewer iterations.  Same contract as orc_pcg.  What the on-chip
 * coarse-grid kernel (csrc/small.cu) runs with B200_PCG_CHEBYSHEV. */
int orc_pcg_cheb(const orc_op *M, const double *b, double *x, double tol, int maxit,
                 int degree, double lmax, double ratio, int *iters, double *relres);
/* block-Jacobi.  z = B^-1 r, B = the diagonal blocks of A over the
 * partition block_of_row[] (a block symmetrised, inverted by Cholesky, its entries cut to
 * the high word of the fp64 value as the product stores them).  Same contract as orc_pcg; 3 = a block is not positive definite.
 * What csrc/small.cu runs with B200_PCG_BLOCK_JACOBI. */
int orc_pcg_bj(const orc_op *M, const double *b, double *x, double tol, int maxit,
               const uint32_t *block_of_row, int *iters, double *relres);
/* Upper bound of the spectrum of D^-1 A the product uses: min(Gershgorin bound,
 * 1.15 x power-iteration estimate after 40 steps from the vector of ones). */
double orc_cheb_lmax(const orc_op *M);
/* fp32-stored operator + fp64 iterative refinement; iters =
 * inner iterations in total, outer = refinement passes, relres = true one. */
int orc_pcg_refine32(const orc_op *M, const double *b, double *x, double tol,
                     int maxit, double eta, int *iters, int *outer, double *relres);
/* ||b - M x||_2 / ||b||_2 with a long-double accumulator. */
double orc_true_relres(const orc_op *M, const double *b, const double *x);

/* ---- direct solve: stand-in for cholmod_l_analyze/factorize/solve ----- */
/* src/cholmod-impl.h:25-26 (setup) and :46,60 (solve). */
typedef struct orc_ldlt orc_ldlt;
enum { ORC_ORDER_NATURAL = 0, ORC_ORDER_RCM = 1 };
orc_ldlt *orc_ldlt_factor(const orc_op *M, int ordering);
uint64_t orc_ldlt_nnz(const orc_ldlt *F);
/* Returns 0 if every pivot was > 0 (SPD), else 1. */
int orc_ldlt_status(const orc_ldlt *F);
void orc_ldlt_solve(const orc_ldlt *F, const double *b, double *x);
void orc_ldlt_free(orc_ldlt *F);

/* ---- synthetic operators (BASELINE.json configs 3-5) ------------------ */
/* Rows [row0,row1) of the N^3 grid operator, natural order (x fastest),
 * global column ids; offs are relative to row0.  diag 6 / 26, off -1. */
orc_op *orc_gen_poisson7(uint32_t N, uint64_t row0, uint64_t row1);
orc_op *orc_gen_poisson27(uint32_t N, uint64_t row0, uint64_t row1);
/* Power-law row lengths, see DESIGN.md "powerlaw generator". */
orc_op *orc_gen_powerlaw(uint64_t n, uint64_t seed, uint64_t row0,
                         uint64_t row1);
uint32_t orc_powerlaw_rowlen(uint64_t n, uint64_t seed, uint64_t row);
/* 65537 thresholds shared with the device generator (see gen.c). */
void orc_powerlaw_table(uint64_t *thr);

int orc_num_threads(void);
/* OpenMP threads for the *_omp functions: nthreads <= 0 means every online
 * core, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1). */
int orc_set_threads(int nthreads);

/* bench.py's CPU legs: `its` iterations of the Jacobi-PCG kernel sequence on a
 * ROW SLAB of a larger operator -- M holds rows [row0, row0 + M->n) with
 * global column ids (orc_gen_poisson27 with a row range), the vectors p, q,
 * r, x live on the slab, p on all n_global columns.  Not a solve: the
 * per-iteration cost of real rows of the real operator, OpenMP over all
 * threads.  One untimed iteration first (pages, caches).  Returns the
 * seconds the `its` timed iterations took (CLOCK_MONOTONIC), < 0 on error. */
double orc_pcg_slab_seconds(const orc_op *M, uint64_t n_global, uint64_t row0, int its);

#ifdef __cplusplus
}
#endif
#endif
