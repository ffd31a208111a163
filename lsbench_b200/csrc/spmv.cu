// spmv.cu -- north_star piece (2): fp64 SpMV y = A x on the device layout of
// convert.cu, optionally fused with the CG dot product p.Ap.
//
// No reference counterpart (the reference has no SpMV of its own;
// its backends hand the CSR to cuSOLVER / CHOLMOD / Ginkgo).  HBM-bound:
// algorithmic bytes 12 nnz + 4 (n+1) + 16 n, no tensor cores.
//
//   k_spmv_sell  one warp per 32-row slice, one row per lane, column-major
//                slice: cols / vals warp loads are single aligned 128 B /
//                256 B segments (read once, streaming hint), x is gathered
//                through L1/L2 (__ldg), each lane sums its row left to right
//                (same order as a host CSR loop => bit-identical to it when
//                the host uses fma).
//                With index compression (the default) uniform slices read w
//                column deltas instead of 32 w columns.
//   k_spmv_vec   one warp per row, 128-bit col / val loads, fixed butterfly
//                warp-shuffle reduction.
//   k_spmv_long  one CTA per row for the power-law tail.
//
// All grids are persistent (a multiple of the SM count) so the per-CTA
// partials of the fused dot have a fixed count and a fixed summation order.
#include "common.cuh"

#include "sell_kernels.cuh"

// ---------------------------------------------------------------------------
static int persistent_grid(b200_ctx *c, const void *kernel, uint64_t work_ctas,
                           int threads = SPMV_THREADS) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel,
                                                    threads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 4;
  uint64_t g = (uint64_t)c->sm_count * per_sm;
  if (g > work_ctas)
    g = work_ctas;
  return g < 1 ? 1 : (int)g;
}

// chunks of 9 where the rows are a multiple of 9 and not of 8 wide (27-point: 9 + 9 + 9
// instead of 8 + 8 + 8 + 3: three dependent rounds of loads per slice instead of four, same
// 48 registers; inside the headline solve 5.794 -> 5.468 ms per SpMV); index-compressed
// layout only.  B200_SPMV_CH=8 keeps 8.
static bool use_chunks_of_9(const b200_mat *M) {
  static const int ch_env = [] {
    const char *v = getenv("B200_SPMV_CH");
    return v ? atoi(v) : 0;
  }();
  return M->sell_meta && ch_env != 8 && M->sell_max_width % 9 == 0 && M->sell_max_width % 8 != 0;
}

static SpmvPlan compute_plan(b200_mat *M, int phase) {
  SpmvPlan P = {0, 0, 0, 0, 0, 0, 0};
  b200_ctx *c = M->ctx;
  uint32_t ns = M->sell_slices;
  // Slices that hold only interior rows.  With a permuted SELL list the
  // windows are B2_SELL_SIGMA rows wide, so round inward to window multiples.
  uint32_t ib = 0, ie = ns;
  if (M->halo.n_halo) {
    if (M->sell_perm && (M->vec_rows || M->long_rows)) {
      ib = ie = 0;  // compacted list: slice <-> row range is not monotone
    } else {
      uint64_t g = M->sell_perm ? M->sell_sigma : B2_SLICE;
      uint64_t lo = (M->interior_begin + g - 1) / g * g;
      uint64_t hi = M->interior_end / g * g;
      if (hi > lo)
        ib = (uint32_t)(lo / B2_SLICE), ie = (uint32_t)(hi / B2_SLICE);
      else
        ib = ie = 0;
    }
  }
  bool others = true;
  if (phase == 0)
    P.b0 = 0, P.e0 = ns;
  else if (phase == 1)
    P.b0 = ib, P.e0 = ie, others = false;
  else
    P.b0 = 0, P.e0 = ib, P.b1 = ie, P.e1 = ns;
  uint32_t nv = (P.e0 - P.b0) + (P.e1 - P.b1);
  if (nv) {
    // one grid per matrix (fixed count of dot partials = fixed summation
    // order): sized for the instantiation that will run
    const bool ch9 = use_chunks_of_9(M);
    const void *fn =
        M->sell_vals32
            ? (M->sell_meta ? (ch9 ? (const void *)k_spmv_sellc<true, float, false, 9> : (const void *)k_spmv_sellc<true, float>)
                            : (const void *)k_spmv_sell<true, float>)
            : (M->sell_meta ? (ch9 ? (const void *)k_spmv_sellc<true, double, false, 9> : (const void *)k_spmv_sellc<true, double>)
                            : (const void *)k_spmv_sell<true, double>);
    P.g_sell = persistent_grid(c, fn, (nv + SPMV_WARPS - 1) / SPMV_WARPS);
  }
  if (others && M->vec_rows)
    P.g_vec = persistent_grid(c, (const void *)k_spmv_vec<true>,
                              (M->vec_rows + SPMV_WARPS - 1) / SPMV_WARPS);
  if (others && M->long_rows)
    P.g_long = persistent_grid(c, (const void *)k_spmv_long<true>, M->long_rows);
  return P;
}

static const SpmvPlan &plan_phase(b200_mat *M, int phase) {
  if (!M->plan_ready) {
    for (int p = 0; p < 3; p++)
      M->plan[p] = compute_plan(M, p);
    if (M->grouped_slices && M->sell_slices)
      M->plan_grp = persistent_grid(M->ctx, (const void *)k_spmv_sell_grp<true>,
                                    ((M->sell_slices + SELL_GRP - 1) / SELL_GRP + SPMV_WARPS - 1) / SPMV_WARPS +
                                        M->long_rows + M->vec_rows);
    M->plan_ready = true;
  }
  return M->plan[phase];
}

// Partials layout for the fused dot: phase-1 CTAs first, then phase-2 (or the
// phase-0 CTAs alone).  Totals are fixed per matrix => fixed summation order.
int launch_spmv(b200_mat *M, const double *x, double *y, bool dot, int phase,
                const XrArgs *xrp) {
  const XrArgs xr = (dot && xrp) ? *xrp : XrArgs{nullptr, nullptr, 1, 0, 0, 0ull};
  b200_ctx *c = M->ctx;
  cudaStream_t s = c->stream;
  const SpmvPlan P = plan_phase(M, phase);
  unsigned slot_base = 0, total = 0;
  if (dot) {
    if (phase == 0) {
      total = P.g_sell + P.g_vec + P.g_long;
    } else {
      const SpmvPlan P1 = plan_phase(M, 1), P2 = plan_phase(M, 2);
      unsigned t1 = P1.g_sell, t2 = P2.g_sell + P2.g_vec + P2.g_long;
      total = t1 + t2;
      slot_base = phase == 1 ? 0 : t1;
    }
    if (total == 0)
      B_FAIL(B200_EINVAL, "launch_spmv: empty matrix");
  }
  uint32_t n = (uint32_t)M->n_local;
  double *dot_out = dot ? (c->nranks > 1 ? &M->state->pq_loc : &M->state->pq) : nullptr;
  if (P.g_sell && M->grouped_slices && !dot && !M->sell_meta && M->sell_vals && phase == 0) {
    // a column range of a column-blocked operator: four slices per warp trip (sell_kernels.cuh)
    k_spmv_sell_grp<false><<<M->plan_grp, SPMV_THREADS, 0, s>>>(
        M->sell_off, M->sell_cols, M->sell_vals, M->sell_perm, x, y, M->sell_slices, n, M->grp_work,
        M->long_rows, M->long_row_ids, M->long_off, M->vec_rows, M->vec_row_ids, M->vec_off, M->vl_cols, M->vl_vals);
    c->launches += 1;
    CU_TRY(cudaGetLastError());
    return B200_OK;  // (the row-major bins are units of the same launch)
  } else if (P.g_sell) {
#define B2_SELL_ARGS(DOTV)                                                        \
  M->sell_perm, x, y, P.b0, P.e0, P.b1, P.e1, n, DOTV ? M->partials : nullptr,    \
      DOTV ? slot_base : 0u, DOTV ? total : 0u, DOTV ? M->state : nullptr,        \
      DOTV ? dot_out : nullptr, xr
    const uint4 *meta = (const uint4 *)M->sell_meta;
    const bool f32 = M->sell_vals32 && (M->spmv_use32 || !M->sell_vals);
    const bool ch9 = use_chunks_of_9(M);
#define B2_SELL_LAUNCH(VT, VALS)                                                  \
  do {                                                                            \
    if (dot && meta && ch9)                                                       \
      k_spmv_sellc<true, VT, false, 9><<<P.g_sell, SPMV_THREADS, 0, s>>>(         \
          meta, M->sell_cols, M->sell_dcols, VALS, B2_SELL_ARGS(true));           \
    else if (meta && ch9)                                                         \
      k_spmv_sellc<false, VT, false, 9><<<P.g_sell, SPMV_THREADS, 0, s>>>(        \
          meta, M->sell_cols, M->sell_dcols, VALS, B2_SELL_ARGS(false));          \
    else if (dot && meta)                                                         \
      k_spmv_sellc<true, VT><<<P.g_sell, SPMV_THREADS, 0, s>>>(                   \
          meta, M->sell_cols, M->sell_dcols, VALS, B2_SELL_ARGS(true));           \
    else if (meta)                                                                \
      k_spmv_sellc<false, VT><<<P.g_sell, SPMV_THREADS, 0, s>>>(                  \
          meta, M->sell_cols, M->sell_dcols, VALS, B2_SELL_ARGS(false));          \
    else if (dot)                                                                 \
      k_spmv_sell<true, VT><<<P.g_sell, SPMV_THREADS, 0, s>>>(                    \
          M->sell_off, M->sell_cols, VALS, B2_SELL_ARGS(true));                   \
    else                                                                          \
      k_spmv_sell<false, VT><<<P.g_sell, SPMV_THREADS, 0, s>>>(                   \
          M->sell_off, M->sell_cols, VALS, B2_SELL_ARGS(false));                  \
  } while (0)
    if (f32)
      B2_SELL_LAUNCH(float, M->sell_vals32);
    else
      B2_SELL_LAUNCH(double, M->sell_vals);
#undef B2_SELL_LAUNCH
#undef B2_SELL_ARGS
    slot_base += P.g_sell;
  }
  if (P.g_vec) {
    if (dot)
      k_spmv_vec<true><<<P.g_vec, SPMV_THREADS, 0, s>>>(
          M->vec_rows, M->vec_row_ids, M->vec_off, M->vl_cols, M->vl_vals, x, y,
          M->partials, slot_base, total, M->state, dot_out, xr);
    else
      k_spmv_vec<false><<<P.g_vec, SPMV_THREADS, 0, s>>>(
          M->vec_rows, M->vec_row_ids, M->vec_off, M->vl_cols, M->vl_vals, x, y,
          nullptr, 0, 0, nullptr, nullptr, xr);
    slot_base += P.g_vec;
  }
  if (P.g_long) {
    if (dot)
      k_spmv_long<true><<<P.g_long, SPMV_THREADS, 0, s>>>(
          M->long_rows, M->long_row_ids, M->long_off, M->vl_cols, M->vl_vals, x,
          y, M->partials, slot_base, total, M->state, dot_out, xr);
    else
      k_spmv_long<false><<<P.g_long, SPMV_THREADS, 0, s>>>(
          M->long_rows, M->long_row_ids, M->long_off, M->vl_cols, M->vl_vals, x,
          y, nullptr, 0, 0, nullptr, nullptr, xr);
  }
  c->launches += (P.g_sell > 0) + (P.g_vec > 0) + (P.g_long > 0);
  CU_TRY(cudaGetLastError());
  return B200_OK;
}

// y += A x on all bins (no fused dot, no phases): what the second and later
// column blocks of a column-blocked matrix run (convert.cu build_col_blocked).
static int launch_spmv_acc(b200_mat *M, const double *x, double *y) {
  b200_ctx *c = M->ctx;
  cudaStream_t s = c->stream;
  const SpmvPlan P = plan_phase(M, 0);
  const XrArgs xr = XrArgs{nullptr, nullptr, 1, 0, 0, 0ull};
  const uint32_t n = (uint32_t)M->n_local;
  if (P.g_sell && M->grouped_slices && !M->sell_meta && M->sell_vals) {
    k_spmv_sell_grp<true><<<M->plan_grp, SPMV_THREADS, 0, s>>>(
        M->sell_off, M->sell_cols, M->sell_vals, M->sell_perm, x, y, M->sell_slices, n, M->grp_work,
        M->long_rows, M->long_row_ids, M->long_off, M->vec_rows, M->vec_row_ids, M->vec_off, M->vl_cols, M->vl_vals);
    c->launches += 1;
    CU_TRY(cudaGetLastError());
    return B200_OK;  // (the row-major bins are units of the same launch)
  } else if (P.g_sell) {
    const uint4 *meta = (const uint4 *)M->sell_meta;
    const bool f32 = M->sell_vals32 && (M->spmv_use32 || !M->sell_vals);
#define B2_ACC_ARGS M->sell_perm, x, y, P.b0, P.e0, P.b1, P.e1, n, nullptr, 0u, 0u, nullptr, nullptr, xr
    if (meta && f32)
      k_spmv_sellc<false, float, true><<<P.g_sell, SPMV_THREADS, 0, s>>>(
          meta, M->sell_cols, M->sell_dcols, M->sell_vals32, B2_ACC_ARGS);
    else if (meta)
      k_spmv_sellc<false, double, true><<<P.g_sell, SPMV_THREADS, 0, s>>>(
          meta, M->sell_cols, M->sell_dcols, M->sell_vals, B2_ACC_ARGS);
    else if (f32)
      k_spmv_sell<false, float, true><<<P.g_sell, SPMV_THREADS, 0, s>>>(
          M->sell_off, M->sell_cols, M->sell_vals32, B2_ACC_ARGS);
    else
      k_spmv_sell<false, double, true><<<P.g_sell, SPMV_THREADS, 0, s>>>(
          M->sell_off, M->sell_cols, M->sell_vals, B2_ACC_ARGS);
#undef B2_ACC_ARGS
  }
  if (P.g_vec)
    k_spmv_vec<false, true><<<P.g_vec, SPMV_THREADS, 0, s>>>(
        M->vec_rows, M->vec_row_ids, M->vec_off, M->vl_cols, M->vl_vals, x, y, nullptr, 0, 0,
        nullptr, nullptr, xr);
  if (P.g_long)
    k_spmv_long<false, true><<<P.g_long, SPMV_THREADS, 0, s>>>(
        M->long_rows, M->long_row_ids, M->long_off, M->vl_cols, M->vl_vals, x, y, nullptr, 0, 0,
        nullptr, nullptr, xr);
  c->launches += (P.g_sell > 0) + (P.g_vec > 0) + (P.g_long > 0);
  CU_TRY(cudaGetLastError());
  return B200_OK;
}

// A column-blocked operator: y = A_0 x, then y += A_b x block after block, so that
// the random gathers of one pass stay inside one L2-sized range of x.  On several ranks the
// ranges of owned columns come first and are multiplied while the halo is on its way.
static int spmv_col_blocked(b200_mat *M, double *x, double *y, bool exchange) {
  if (exchange)
    B_TRY(halo_exchange_begin(M, x));
  for (size_t b = 0; b < M->blocks.size(); b++) {
    if (exchange && b == M->n_local_blocks)
      B_TRY(halo_exchange_wait(M));
    if (b == 0)
      B_TRY(launch_spmv(M->blocks[b], x, y, false, 0, nullptr));
    else
      B_TRY(launch_spmv_acc(M->blocks[b], x, y));
  }
  if (exchange && M->n_local_blocks >= M->blocks.size())
    B_TRY(halo_exchange_wait(M));
  return B200_OK;
}

// Full SpMV with halo exchange overlapped with the interior rows (piece 5).
// peer_seq != 0 (PCG iterations only, x_ext = the p vector): the halo goes over peer
// memory with that sequence offset (dist.cu), not through NCCL.
static int spmv_full(b200_mat *M, double *x_ext, double *y, bool dot,
                     const XrArgs *xr = nullptr, unsigned peer_seq = 0) {
  if (!M->blocks.empty()) {
    if (dot)
      B_FAIL(B200_EINVAL, "a column-blocked matrix is SpMV-only");
    return spmv_col_blocked(M, x_ext, y, M->ctx->nranks > 1);
  }
  if (!M->halo.n_halo && M->ctx->nranks == 1)
    return launch_spmv(M, x_ext, y, dot, 0, nullptr);
  if (peer_seq && M->halo.peer_ready && x_ext == M->w_p) {
    B_TRY(halo_peer_push(M, x_ext, peer_seq));
    B_TRY(launch_spmv(M, x_ext, y, dot, 1, xr));
    B_TRY(halo_peer_wait(M, peer_seq));
    return launch_spmv(M, x_ext, y, dot, 2, xr);
  }
  B_TRY(halo_exchange_begin(M, x_ext));
  B_TRY(launch_spmv(M, x_ext, y, dot, 1, xr));
  B_TRY(halo_exchange_wait(M));
  return launch_spmv(M, x_ext, y, dot, 2, xr);
}

int spmv_full_internal(b200_mat *M, double *x_ext, double *y, bool dot, const XrArgs *xr,
                       unsigned peer_seq) {
  return spmv_full(M, x_ext, y, dot, xr, peer_seq);
}

extern "C" int b200_spmv(b200_mat *M, const double *d_x, double *d_y) {
  if (!M || !d_x || !d_y)
    B_FAIL(B200_EINVAL, "b200_spmv: null argument");
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  if (!M->blocks.empty() && c->nranks == 1)
    return spmv_col_blocked(M, const_cast<double *>(d_x), d_y, false);
  if (c->nranks == 1)
    return launch_spmv(M, d_x, d_y, false, 0, nullptr);
  B_TRY(ensure_workspace(M));
  CU_TRY(cudaMemcpyAsync(M->x_ext, d_x, M->n_local * 8,
                         cudaMemcpyDeviceToDevice, c->stream));
  return spmv_full(M, M->x_ext, d_y, false);
}

extern "C" int b200_spmv_host(b200_mat *M, const double *h_x, double *h_y) {
  if (!M || !h_x || !h_y)
    B_FAIL(B200_EINVAL, "b200_spmv_host: null argument");
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  B_TRY(ensure_workspace(M));
  cudaStream_t s = c->stream;
  CU_TRY(cudaMemcpyAsync(M->x_ext, h_x, M->n_local * 8, cudaMemcpyHostToDevice, s));
  B_TRY(spmv_full(M, M->x_ext, M->w_q, false));
  CU_TRY(cudaMemcpyAsync(h_y, M->w_q, M->n_local * 8, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

extern "C" int b200_spmv_time(b200_mat *M, const double *d_x, double *d_y,
                              int reps, float *ms) {
  if (!M || !d_x || !d_y || !ms || reps < 1)
    B_FAIL(B200_EINVAL, "b200_spmv_time: bad argument");
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  B_TRY(b200_spmv(M, d_x, d_y));  // warm
  CU_TRY(cudaEventRecord(c->ev_a, c->stream));
  for (int i = 0; i < reps; i++)
    B_TRY(b200_spmv(M, d_x, d_y));
  CU_TRY(cudaEventRecord(c->ev_b, c->stream));
  CU_TRY(cudaEventSynchronize(c->ev_b));
  float t = 0;
  CU_TRY(cudaEventElapsedTime(&t, c->ev_a, c->ev_b));
  *ms = t / reps;
  return B200_OK;
}
