/*
 * b200.h -- C ABI of the B200-native sparse solver layer (libb200.so).
 *
 * This is the thin layer `src/b200.c` (the lsbench `--solver b200` backend,
 * see INTEGRATION.md) calls; it is also what tests/ and bench.py bind with
 * ctypes.  Plain pointers and sizes only; every function returns 0 on success
 * and a B200_E* code otherwise, with a message in b200_last_error().  There is
 * no CPU fallback: on a machine without a usable CUDA device the calls fail.
 *
 * Which reference interface each entry point replaces (paths relative to the
 * lsbench tree):
 *
 *   b200_ctx_create / _destroy     X_init / X_finalize file-scope handles,
 *                                  src/cusparse.c:33-36,138-162
 *   b200_mat_from_csr              backend csr_init: host `struct csr`
 *                                  (src/lsbench-impl.h:22-26) -> backend
 *                                  layout + H2D, src/cusparse.c:47-125; with
 *                                  B200_MAT_SYM_UPPER the operator is the one
 *                                  CHOLMOD factorises, src/cholmod-impl.h:5-21
 *   b200_mat_destroy               csr_finalize, src/cusparse.c:127-136
 *   b200_text_to_csr               the record loop of lsbench_matrix_read,
 *                                  src/lsbench-csr.c:49-53, plus the body below
 *   b200_coo_to_csr / _mat_from_coo the sort / fold / row-compress / fill body
 *                                  of lsbench_matrix_read once the text is
 *                                  tokenised, src/lsbench-csr.c:54-86
 *                                  (device-side ingest)
 *   b200_pcg_solve                 the solve inside the X_bench timed loop,
 *                                  src/cusparse.c:189-197 /
 *                                  src/cholmod-impl.h:58-63 /
 *                                  src/ginkgo.cpp:91-99 (x reset per trial)
 *   b200_spmv*                     no reference counterpart
 *   b200_mat_generate              no reference counterpart: BASELINE.json
 *                                  configs 3-5 cannot go through the text
 *                                  reader (src/lsbench-csr.c:35,54)
 */
#ifndef B200_H_
#define B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: b200_mat_info grew (index compression), ingest and row-block entry points
 * 3: b200_mat_info.values_f32, b200_pcg_result.outer_iters, new flags
 * 4: b200_pcg_result.replacements, status 4, Chebyshev flags; single-reduction flag gone
 * 5: B200_PCG_BLOCK_JACOBI, b200_pcg_result.block_jacobi, b200_mat_block_jacobi_partition */
#define B200_ABI_VERSION 5

enum {
  B200_OK = 0,
  B200_EINVAL = 1,   /* bad argument */
  B200_ECUDA = 2,    /* CUDA runtime / driver error */
  B200_ENOMEM = 3,
  B200_ENCCL = 4,    /* NCCL missing or failed */
  B200_ENOTSPD = 5,  /* PCG breakdown: p.Ap <= 0 or NaN */
  B200_ERANGE = 6    /* index width exceeded */
};

typedef struct b200_ctx b200_ctx;
typedef struct b200_mat b200_mat;

const char *b200_last_error(void);
int b200_abi_version(void);
/* Number of CUDA devices visible; fails (B200_ECUDA) without a driver. */
int b200_device_count(int *count);

/* ---- context: one per rank, one GPU per rank ---------------------------- */
int b200_ctx_create(int device, b200_ctx **ctx);
/* Row-block distributed context.  nccl_id is B200_NCCL_ID_BYTES bytes from
 * b200_nccl_unique_id() on rank 0, handed to every rank by the caller
 * (torch.distributed broadcast, or shared memory in a threaded host). */
#define B200_NCCL_ID_BYTES 128
int b200_nccl_unique_id(void *id_out);
int b200_ctx_create_dist(int device, int rank, int nranks, const void *nccl_id,
                         b200_ctx **ctx);
int b200_ctx_destroy(b200_ctx *ctx);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) so the
 * caller's CUDA events bracket the kernels.  NULL = the context's own. */
int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream);
int b200_ctx_sync(b200_ctx *ctx);
int b200_ctx_rank(const b200_ctx *ctx, int *rank, int *nranks);
/* The row block [r0, r1) rank `rank` of `nranks` owns of an n-row operator:
 * contiguous, cuts at n*k/nranks rounded to 32 rows.  Pure host arithmetic
 * (no device needed): callers use it to slice b and x per rank. */
int b200_row_block(uint64_t n, int rank, int nranks, uint64_t *r0, uint64_t *r1);

/* Device memory owned by the library (so a C host needs no CUDA headers). */
int b200_malloc(b200_ctx *ctx, size_t bytes, void **dptr);
int b200_free(b200_ctx *ctx, void *dptr);
int b200_memcpy_h2d(b200_ctx *ctx, void *dst, const void *src, size_t bytes);
int b200_memcpy_d2h(b200_ctx *ctx, void *dst, const void *src, size_t bytes);
int b200_memset(b200_ctx *ctx, void *dst, int byte, size_t bytes);
/* Pinned host memory for the end-to-end path. */
int b200_host_alloc(size_t bytes, void **hptr);
int b200_host_free(void *hptr);

/* ---- matrix -------------------------------------------------------------- */
enum {
  B200_MAT_SYM_UPPER = 1u << 0, /* solve the upper triangle mirrored */
  B200_MAT_FORCE_VECTOR = 1u << 1, /* kernel sweep: no SELL bin */
  B200_MAT_FORCE_SELL = 1u << 2,   /* kernel sweep: everything in SELL */
  B200_MAT_NO_SORT = 1u << 3,      /* SELL without the length-sort window */
  B200_MAT_NO_COMPRESS = 1u << 4,  /* keep one explicit u32 column per entry */
  /* SURVEY 8f row 4 (--precision FP32): store the SELL values as fp32.  If every
   * value is exactly representable (any stencil) that is lossless: the fp64
   * copy is dropped, products and PCG iterates keep their bits, the value
   * stream halves.  Otherwise both copies stay resident and b200_pcg_solve
   * runs iterative refinement: inner PCG on the fp32-valued operator (fp64
   * vectors), residual b - A x with the fp64 values, to the same fp64 bar. */
  B200_MAT_VALUES_F32 = 1u << 5,
  /* Column blocking for operators whose SpMV is bound by random gathers (the
   * power-law matrix: 101 B of DRAM traffic per 8-byte gather): the columns are
   * cut into ranges of B200_COL_BLOCK_MB (environment, default 64) megabytes of
   * x, every range becomes a matrix of its own over all rows, and y = A x runs
   * as y = A_0 x; y += A_1 x; ... so that the gathers of one pass stay inside
   * one L2-sized piece of x.  A range sorts its rows by length inside windows of
   * B200_COL_BLOCK_SIGMA rows (default 131072; 0 = the whole list) and multiplies
   * four slices per warp trip (k_spmv_sell_grp; B200_COL_BLOCK_KERNEL=plain keeps
   * one slice per trip).  On several ranks the ranges follow the global column
   * order -- remote columns below the own rows, owned columns, remote columns
   * above -- and the owned ranges are multiplied while the halo travels.  SpMV
   * only (b200_pcg_solve refuses); row sums are formed block by block, i.e.
   * equal to the CSR product to rounding, not bit for bit.  Ignored when one
   * block would hold everything. */
  B200_MAT_COL_BLOCK = 1u << 6
};

/* Host CSR exactly as lsbench_matrix_read leaves it: offs 0-based, cols
 * carrying `base` (src/lsbench-csr.c:79-86).  The matrix is global; in a
 * distributed context each rank keeps its row block [n*rank/P, n*(rank+1)/P)
 * rounded to 32 rows. */
int b200_mat_from_csr(b200_ctx *ctx, uint32_t nrows, uint32_t base,
                      const uint32_t *offs, const uint32_t *cols,
                      const double *vals, uint32_t flags, b200_mat **M);

/* ---- device-side ingest ---------------------------------- */
/* The records of a COO text file in file order (rows / cols as written, i.e.
 * still carrying the file's base) -> CSR with the reference reader's
 * semantics (src/lsbench-csr.c:54-86): ordered by (row, col) with a stable
 * sort, equal (row, col) records summed left to right in file order, absent
 * row ids compressed away, offs 0-based, cols keep the base.  Sort, fold and
 * fill run on the device.  0 < nnz < 2^32 (the reference's `unsigned nnz`).
 * Output arrays are caller-owned with room for nnz+1 / nnz / nnz elements;
 * pass NULL for all three to get the sizes only. */
int b200_coo_to_csr(b200_ctx *ctx, uint64_t nnz, const uint32_t *rows,
                    const uint32_t *cols, const double *vals,
                    uint32_t *nrows_out, uint64_t *nnz_out, uint32_t *offs,
                    uint32_t *cols_out, double *vals_out);
/* The same from text: `body` is the file content after the header line
 * ("nnz base\n"), len bytes; its first nnz lines are the records.  Lines are
 * found and parsed on the device by a strict, exact parser (one record per
 * line, "row col val\n"; doubles on Clinger's exact path); the lines it does
 * not accept are parsed by strtoul / strtod on the host (*n_host_parsed of
 * them).  B200_EINVAL when the text is not one strict record per line -- the
 * caller then tokenises with fscanf semantics and uses b200_coo_to_csr. */
int b200_text_to_csr(b200_ctx *ctx, const char *body, uint64_t len, uint64_t nnz,
                     uint32_t *nrows_out, uint64_t *nnz_out, uint32_t *offs,
                     uint32_t *cols_out, double *vals_out,
                     uint64_t *n_host_parsed);
/* Same ingest, then straight into the device layout (no host CSR). */
int b200_mat_from_coo(b200_ctx *ctx, uint64_t nnz, uint32_t base,
                      const uint32_t *rows, const uint32_t *cols,
                      const double *vals, uint32_t flags, b200_mat **M);

enum { B200_GEN_POISSON7 = 1, B200_GEN_POISSON27 = 2, B200_GEN_POWERLAW = 3 };
/* Synthetic operator generated on the device, only this rank's rows.
 * size = N (grid edge) for the Poisson kinds, n (rows) for powerlaw. */
int b200_mat_generate(b200_ctx *ctx, int kind, uint64_t size, uint64_t seed,
                      uint32_t flags, b200_mat **M);
int b200_mat_destroy(b200_mat *M);

#define B200_HIST_BINS 24
typedef struct {
  uint64_t n_global;      /* rows of the whole operator */
  uint64_t row_begin;     /* first global row owned by this rank */
  uint64_t n_local;       /* rows owned */
  uint64_t n_halo;        /* remote x entries this rank reads */
  uint64_t nnz;           /* stored entries of the local rows (after mirror) */
  uint64_t nnz_padded;    /* entries in the device streams incl. padding */
  uint64_t sell_rows, sell_slices, sell_sigma, sell_max_width;
  uint64_t vec_rows, vec_nnz;      /* warp-per-row bin */
  uint64_t long_rows, long_nnz;    /* block-per-row bin */
  uint64_t interior_begin, interior_end; /* local rows with no halo column */
  uint64_t hist[B200_HIST_BINS];  /* rows with 2^(b-1) < len <= 2^b; hist[0]: len 0..1 */
  uint64_t max_row_len;
  uint32_t pattern_symmetric;     /* SYM_UPPER: mirror slots all existed */
  uint32_t sell_perm;             /* 1 if SELL rows are permuted */
  uint64_t device_bytes;
  /* index compression: SELL slices stored as w column deltas instead of
   * 32 w columns, and the bytes of matrix streams (offsets, columns, values,
   * permutation) one SpMV actually reads -- compare with 12 nnz + 4 (n+1) */
  uint64_t sell_uniform_slices;
  uint64_t matrix_stream_bytes;
  /* B200_MAT_VALUES_F32: 0 off, 1 lossless (fp64 copy dropped), 2 rounded
   * (both copies resident, PCG = iterative refinement) */
  uint32_t values_f32;
  uint32_t col_blocks;   /* B200_MAT_COL_BLOCK: number of column ranges, else 0 */
} b200_mat_info;
int b200_mat_get_info(const b200_mat *M, b200_mat_info *info);

/* Reads the local rows back as plain CSR in ORIGINAL row order with local
 * column ids (halo columns >= n_local), padding removed: what the device
 * layout represents.  Pass NULL arrays to query sizes. */
int b200_mat_export(const b200_mat *M, uint64_t *offs, uint32_t *cols,
                    double *vals);
/* Global column id of halo slot j (j < n_halo). */
int b200_mat_halo_cols(const b200_mat *M, uint64_t *gcols);
int b200_mat_inv_diag(const b200_mat *M, double *h_dinv);
/* The partition B200_PCG_BLOCK_JACOBI works on: block_of_row[i] = id of the diagonal
 * block row i (caller's numbering) belongs to, *block_size = rows per block (0 and
 * B200_OK when the matrix does not take the on-chip path or no block size fits). */
int b200_mat_block_jacobi_partition(b200_mat *M, uint32_t *block_of_row, uint32_t *block_size);

/* ---- SpMV ------------------------------------------------------------------ */
/* y = A x on the local rows.  d_x holds the n_local owned entries; the halo is
 * exchanged inside.  d_y: n_local.  Device pointers. */
int b200_spmv(b200_mat *M, const double *d_x, double *d_y);
/* Host buffers: H2D x, SpMV, D2H y. */
int b200_spmv_host(b200_mat *M, const double *h_x, double *h_y);
/* Times `reps` back-to-back launches with CUDA events on the launch stream
 * and returns the mean milliseconds per SpMV (all bins, halo included). */
int b200_spmv_time(b200_mat *M, const double *d_x, double *d_y, int reps,
                   float *ms_per_spmv);

/* ---- Jacobi-preconditioned CG --------------------------------------------- */
typedef struct {
  double tol;        /* stop when ||r||_2 <= tol * ||b||_2 (recurrence r) */
  int32_t maxit;
  int32_t check_every; /* iterations queued between host looks; 0 = default */
  uint32_t flags;
} b200_pcg_opts;
enum {
  B200_PCG_TIME_KERNELS = 1u << 0, /* per-class CUDA-event timing */
  B200_PCG_NO_GRAPH = 1u << 1,
  B200_PCG_NO_SMALL = 1u << 2,     /* never take the on-chip small-matrix path */
  /* (1u << 3 was B200_PCG_SINGLE_REDUCTION, Chronopoulos-Gear CG: measured on 1, 2 and
   * 8 B200 in round 2, slower everywhere -- 27-point 512^3 on 8 GPUs 1.199 s against
   * 1.115 s -- and removed; profiles/r02_single_reduction_lost.txt) */
  /* SURVEY 8f row 2, the preconditioner half: Chebyshev-Jacobi of degree 2 / 3 on the
   * on-chip coarse-grid path (z = P(D^-1 A) D^-1 r, interval [lmax / 30, lmax]): one /
   * two more products per iteration, about 1/2 / 1/3 of the iterations -- and of the
   * cluster-wide reductions each of them waits for.  Ignored on the streaming path,
   * where the product is the cost and CG is already optimal per product. */
  B200_PCG_CHEBYSHEV2 = 1u << 4,
  B200_PCG_CHEBYSHEV3 = 1u << 5,
  /* SURVEY 8f row 2, block-Jacobi on the on-chip coarse-grid path: z = B^-1 r, B the
   * diagonal blocks of 32 (16 when shared memory is short) consecutive rows of a CTA's
   * row chunk, inverted once on the host and kept in shared memory as the high 32 bits of
   * every fp64 entry (a preconditioner may be rounded; widening costs nothing).  No exchange
   * and no reduction more than Jacobi, 0.4 - 0.8 x the iterations on the Nek matrices.
   * What the reference reaches for on these systems is algebraic multigrid
   * (src/hypre.c:126-188, src/amgx.c:78-85); this is the step in that direction that
   * costs a CTA nothing it has to wait for.  Ignored on the streaming path and when a
   * diagonal block is not positive definite. */
  B200_PCG_BLOCK_JACOBI = 1u << 6
};

typedef struct {
  int32_t iters;
  int32_t status;      /* 0 converged, 1 maxit, 2 breakdown (not SPD / NaN),
                          4 stagnated: the recurrence residual met the bar, b - A x
                          did not, and iterating on from the true residual did not
                          get there either (the bar is below what fp64 reaches) */
  double relres;       /* recurrence ||r|| / ||b|| at exit */
  double true_relres;  /* ||b - A x|| / ||b|| recomputed with one SpMV */
  double bnorm;
  float solve_ms;      /* CUDA events around the whole solve */
  float spmv_ms, update_ms, pupdate_ms; /* B200_PCG_TIME_KERNELS only */
  int32_t kernel_launches;
  int32_t path;        /* 0 = streaming kernels, 1 = on-chip small-matrix */
  int32_t outer_iters; /* refinement passes (values_f32 == 2); on the on-chip path the
                          degree of the preconditioner that ran (1 = Jacobi); else 0 */
  int32_t replacements; /* residual replacements: exit checks that found ||b - A x|| above
                           the bar with the recurrence below it, after which the solve went on */
  uint32_t block_jacobi; /* block size of the block-Jacobi preconditioner that ran, else 0 */
} b200_pcg_result;

/* x: in x0, out solution (n_local).  b: n_local.  Device pointers. */
int b200_pcg_solve(b200_mat *M, const double *d_b, double *d_x,
                   const b200_pcg_opts *opts, b200_pcg_result *res);
/* Host buffers: H2D b and x0, solve, D2H x -- the X_bench call shape. */
int b200_pcg_solve_host(b200_mat *M, const double *h_b, double *h_x,
                        const b200_pcg_opts *opts, b200_pcg_result *res);

/* ---- bytes the roofline is quoted on --------------------------- */
/* 12*nnz + 4*(n+1) + 16*n  and  12*nnz + 4*(n+1) + 104*n, local rows. */
int b200_mat_algorithmic_bytes(const b200_mat *M, uint64_t *spmv_bytes,
                               uint64_t *pcg_iter_bytes);

#ifdef __cplusplus
}
#endif
#endif
