// dist.cu -- row-block partition, halo exchange and scalar all-reduce
// (north_star piece 5).  STUB for the first single-GPU milestone.
#include "common.cuh"

int dist_comm_init(b200_ctx *c, const void *nccl_id) {
  (void)c, (void)nccl_id;
  B_FAIL(B200_ENCCL, "b200: multi-rank contexts are not built yet");
}
void dist_comm_destroy(b200_ctx *c) { (void)c; }

extern "C" int b200_nccl_unique_id(void *id_out) {
  (void)id_out;
  B_FAIL(B200_ENCCL, "b200: multi-rank contexts are not built yet");
}

int partition_and_renumber(b200_ctx *c, PlainCsr *A, uint64_t n_global,
                           uint64_t row_begin, b200_mat *M) {
  (void)A, (void)n_global;
  if (c->nranks != 1)
    B_FAIL(B200_ENCCL, "b200: multi-rank contexts are not built yet");
  M->row_begin = row_begin;
  return B200_OK;
}
int halo_setup(b200_mat *M) { (void)M; return B200_OK; }
int halo_exchange_begin(b200_mat *M, double *x) { (void)M, (void)x; return B200_OK; }
int halo_exchange_wait(b200_mat *M) { (void)M; return B200_OK; }
void halo_free(b200_mat *M) { (void)M; }
int allreduce_sum(b200_ctx *c, double *v, int n) { (void)c, (void)v, (void)n; return B200_OK; }

extern "C" int b200_mat_halo_cols(const b200_mat *M, uint64_t *g) {
  (void)M, (void)g;
  return B200_OK;
}
