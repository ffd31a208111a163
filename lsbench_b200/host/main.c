/*
 * main.c -- stand-alone command line of this tree:
 *   driver --solver b200 --matrix FILE|poisson27:512 [--trials=N] [--verbose=1]
 * The reference's own bin/driver.c (five API calls, bin/driver.c:5-15) builds
 * against liblsbench.so unchanged and does the same; this one adds --dump-x.
 */
#include "lsbench.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
  /* --dump-x FILE is ours: strip it before the harness parses the rest */
  const char *dump = NULL;
  int n = 0;
  char **args = (char **)calloc((size_t)argc + 1, sizeof(char *));
  for (int i = 0; i < argc; i++) {
    if (strcmp(argv[i], "--dump-x") == 0 && i + 1 < argc)
      dump = argv[++i];
    else
      args[n++] = argv[i];
  }

  struct lsbench *cb = lsbench_init(n, args);
  struct csr *A = lsbench_matrix_read(lsbench_get_matrix_name(cb));
  int rc;
  if (dump) {
    unsigned m = lsbench_matrix_rows(A);
    double *x = (double *)calloc(m ? m : 1, sizeof(double));
    rc = lsbench_solve(A, cb, x);
    FILE *fp = fopen(dump, "wb");
    if (!fp || fwrite(x, sizeof(double), m, fp) != m)
      rc = 2;
    if (fp)
      fclose(fp);
    free(x);
  } else {
    lsbench_bench(A, cb);
    rc = 0;
  }
  lsbench_matrix_free(A);
  lsbench_finalize(cb);
  free(args);
  return rc;
}
