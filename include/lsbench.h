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

/* Solver ids.  0..5 are the reference's (src/lsbench.h:8-16); 6 is new. */
typedef enum {
  LSBENCH_SOLVER_NONE = -1,
  LSBENCH_SOLVER_CUSOLVER = 0,
  LSBENCH_SOLVER_HYPRE = 1,
  LSBENCH_SOLVER_AMGX = 2,
  LSBENCH_SOLVER_CHOLMOD = 3,
  LSBENCH_SOLVER_PARALMOND = 4,
  LSBENCH_SOLVER_GINKGO = 5,
  LSBENCH_SOLVER_B200 = 6 /* fp64 SELL SpMV + fused Jacobi-PCG, sm_100a */
} lsbench_solver_t;

/* src/lsbench.h:18-22.  Only FP64 is accepted (src/lsbench.c:140-141). */
typedef enum {
  LSBENCH_PRECISION_FP64 = 0,
  LSBENCH_PRECISION_FP32 = 1,
  LSBENCH_PRECISION_FP16 = 2
} lsbench_precision_t;

/* src/lsbench.h:24-29.  The b200 backend does not reorder. */
typedef enum {
  LSBENCH_ORDERING_NONE = -1,
  LSBENCH_ORDERING_RCM = 0,
  LSBENCH_ORDERING_AMD = 1,
  LSBENCH_ORDERING_METIS = 2
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
