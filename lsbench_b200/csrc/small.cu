// small.cu -- on-chip PCG for matrices that fit in distributed shared memory.
// STUB for the first milestone: never selected.
#include "common.cuh"
int small_try_build(b200_mat *M) { (void)M; return B200_OK; }
void small_free(b200_mat *M) { (void)M; }
int small_solve(b200_mat *M, const double *b, double *x, const b200_pcg_opts *o,
                b200_pcg_result *r) {
  (void)M, (void)b, (void)x, (void)o, (void)r;
  B_FAIL(B200_EINVAL, "small path not built");
}
