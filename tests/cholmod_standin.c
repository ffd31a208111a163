/*
 * cholmod_standin.c -- TEST INFRASTRUCTURE: a `cholmod_bench` for the harness
 * built on the CPU oracle's sparse LDL^T (oracle/ldlt.c), so BASELINE.json
 * config 1 ("driver --solver cholmod --test tests/I1_05x05.txt on CPU") runs
 * end to end without SuiteSparse.  Follows the reference wrapper's protocol
 * (src/cholmod-impl.h:34-74): operator from the upper triangle (:5-21),
 * analyze + factorize untimed (:25-26), trials warm-up solves (:45-55), trials
 * timed solves (:58-63), CSV row (:68-70) -- and, unlike the reference
 * (:46-61), copies the solution into x.  Linked only into the test build
 * liblsbench_test.so (tests/test_host_shell.py); never into the product.
 */
#define _GNU_SOURCE
#include "lsbench-impl.h"
#include "oracle.h"
#include <string.h>

int cholmod_bench(double *x, struct csr *A, const double *r,
                  const struct lsbench *cb) {
  if (A->offs == NULL)
    return 1;
  orc_csr H = {A->nrows, A->base, A->offs, A->cols, A->vals};
  orc_op *M = orc_op_upper_mirror(&H);
  orc_ldlt *F = orc_ldlt_factor(M, ORC_ORDER_RCM);
  for (unsigned t = 0; t < cb->trials; t++)
    orc_ldlt_solve(F, r, x);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (unsigned t = 0; t < cb->trials; t++)
    orc_ldlt_solve(F, r, x);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double el = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  unsigned m = A->nrows, nnz = A->offs[m];
  printf("===matrix,n,nnz,trials,solver,ordering,elapsed===\n");
  printf("%s,%u,%u,%u,%u,%d,%.15lf\n", cb->matrix, m, nnz, cb->trials,
         cb->solver, cb->ordering, el);
  fflush(stdout);
  orc_ldlt_free(F), orc_op_free(M);
  return 0;
}
