/*
 * lsbench.h -- public interface of the linear-solver benchmark harness.
 *
 * Source-compatible with the reference's src/lsbench.h:8-40 (same eight entry
 * points, same enumerator values) so that the reference's bin/driver.c builds
 * against this library unchanged; extended with the B200 solver id and a few
 * accessors.  Only the `--solver b200` path is implemented natively here; the
 * third-party wrappers of the reference (cuSOLVER, hypre, AmgX, CHOLMOD,
 * parAlmond, Ginkgo) are out of scope and report "not built".
 */
#ifndef _LSBENCH_
#define _LSBENCH_

#ifdef __cplusplus
extern "C" {
#endif

/* Solver, precision and ordering ids: the reference's names and values
 * (src/lsbench.h:8-29), so code written against either header compiles
 * against the other; LSBENCH_SOLVER_B200 is the one addition.
 *
 *   id  --solver     backend
 *   -1  (none)       unknown string (the reference falls back to CHOLMOD, src/lsbench.c:31-33)
 *    0  cusolver     cuSOLVER-Sp sparse Cholesky      not built in this tree
 *    1  hypre        BoomerAMG                        not built
 *    2  amgx         AmgX classical AMG (fp32)        not built
 *    3  cholmod      CHOLMOD direct solve             not built (stand-in: tests/cholmod_standin.c)
 *    4  paralmond    parAlmond AMG                    not built
 *    5  ginkgo       BiCGSTAB + Jacobi                not built
 *    6  b200         fp64 SELL SpMV + fused Jacobi-PCG, sm_100a   this tree
 *
 * Precision: FP64 and FP32 are accepted (FP32 = matrix values stored as fp32, fp64
 * arithmetic and the same fp64 bars: lossless on stencils, iterative refinement
 * otherwise); FP16 is rejected as in src/lsbench.c:140-141.  Ordering: parsed and
 * printed; the b200 backend applies RCM when LSBENCH_B200_ORDERING says so (AMD and
 * METIS only reduce the fill of a factorisation: reported, not applied) and
 * renumbers on its own inside its coarse-grid kernel. */
typedef enum {
  LSBENCH_SOLVER_NONE = -1, LSBENCH_SOLVER_CUSOLVER, LSBENCH_SOLVER_HYPRE, LSBENCH_SOLVER_AMGX,
  LSBENCH_SOLVER_CHOLMOD, LSBENCH_SOLVER_PARALMOND, LSBENCH_SOLVER_GINKGO, LSBENCH_SOLVER_B200
} lsbench_solver_t;
typedef enum { LSBENCH_PRECISION_FP64, LSBENCH_PRECISION_FP32, LSBENCH_PRECISION_FP16 } lsbench_precision_t;
typedef enum {
  LSBENCH_ORDERING_NONE = -1, LSBENCH_ORDERING_RCM, LSBENCH_ORDERING_AMD, LSBENCH_ORDERING_METIS
} lsbench_ordering_t;

/* ---- matrix ------------------------------------------------------------ */
struct csr;
/* COO text file ("nnz base" then nnz lines "row col val") -> CSR, with the
 * semantics of src/lsbench-csr.c:29-92.  Besides file names, the pseudo-names
 * "poisson7:N", "poisson27:N" and "powerlaw:n[:seed]" give a descriptor whose
 * rows are generated on the device by the b200 backend. */
struct csr *lsbench_matrix_read(const char *fname);
void lsbench_matrix_print(const struct csr *A);
void lsbench_matrix_free(struct csr *A);
/* extensions */
unsigned lsbench_matrix_rows(const struct csr *A);
unsigned long long lsbench_matrix_nnz(const struct csr *A);

/* ---- harness ----------------------------------------------------------- */
struct lsbench;
/* Parses --matrix/--test, --solver, --ordering, --precision, --verbose,
 * --trials, --help (src/lsbench.c:82-135); both "--opt value" and
 * "--opt=value" are accepted. */
struct lsbench *lsbench_init(int argc, char *argv[]);
const char *lsbench_get_matrix_name(struct lsbench *cb);
/* b[i] = i, x0 = 0 (src/lsbench.c:157-160), then the selected backend's
 * X_bench(x, A, b, cb). */
void lsbench_bench(struct csr *A, const struct lsbench *cb);
void lsbench_finalize(struct lsbench *cb);
/* extension: like lsbench_bench, but the solution is copied to x_out
 * (nrows doubles) when it is not NULL; returns the backend's status. */
int lsbench_solve(struct csr *A, const struct lsbench *cb, double *x_out);

#ifdef __cplusplus
}
#endif

#endif
