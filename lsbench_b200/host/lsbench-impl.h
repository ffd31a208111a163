/*
 * lsbench-impl.h -- private ABI between the harness and its backends.
 *
 * The leading members of `struct lsbench` and `struct csr` are laid out as in
 * the reference (src/lsbench-impl.h:14-26) so a backend file written against
 * either tree compiles against the other; members after the marker are
 * extensions of this tree.
 */
#ifndef _LSBENCH_IMPL_
#define _LSBENCH_IMPL_

#include "lsbench.h"
#include <err.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

struct lsbench {
  char *matrix;
  lsbench_solver_t solver;
  lsbench_ordering_t ordering;
  lsbench_precision_t precision;
  unsigned verbose, trials;
};

struct csr {
  unsigned nrows, base;
  unsigned *offs, *cols; /* offs 0-based; cols keep `base` */
  double *vals;
  /* ---- extensions: a generated matrix has offs == NULL and these set ---- */
#define LSBENCH_HAS_SYNTHETIC 1
  int gen_kind; /* 0 = from file; else B200_GEN_* */
  unsigned long long gen_size, gen_seed, gen_nnz;
};

#define tcalloc(T, n) ((T *)calloc((n), sizeof(T)))
#define tfree(p) free((void *)(p))

/* Backend convention (src/lsbench-impl.h:42-68): X_init / X_finalize return 1
 * if already (un)initialised, X_bench returns 1 if the backend is not usable,
 * 0 on success; x has nrows zeros on entry, r is read-only. */
#define LSBENCH_BENCH_FN(name)                                                 \
  int name##_bench(double *x, struct csr *A, const double *r,                  \
                   const struct lsbench *cb)
int b200_init(void);
int b200_finalize(void);
LSBENCH_BENCH_FN(b200);

/* The reference's third-party wrappers are not built in this tree; their
 * bench entry points exist (lsbench.c) and report "not built". */
LSBENCH_BENCH_FN(cusparse);
LSBENCH_BENCH_FN(hypre);
LSBENCH_BENCH_FN(amgx);
LSBENCH_BENCH_FN(cholmod);
LSBENCH_BENCH_FN(paralmond);
LSBENCH_BENCH_FN(ginkgo);

#ifdef __cplusplus
}
#endif
#endif
