// pcg.cu -- north_star piece (3): Jacobi-preconditioned CG with the vector
// work fused into three HBM passes per iteration.
//
// Stands where the timed solve of a reference backend stands
// (src/cusparse.c:189-197 cusolverSpDcsrlsvchol, src/cholmod-impl.h:58-63
// cholmod_l_solve, src/ginkgo.cpp:91-99 solver->apply); protocol per the
// Ginkgo wiring, the one Krylov precedent: Jacobi preconditioner, relative
// residual stop (src/ginkgo.cpp:55-64), x reset by the caller per trial (:92).
//
//   K1  q = A p            + p.q          (spmv.cu, fused dot)
//   K2  x += a p; r -= a q + r.D^-1 r, r.r
//   K3  p = D^-1 r + b p
//
// Algorithmic bytes / iteration: 12 nnz + 4 (n+1) + 104 n.
// alpha, beta, the convergence test and the iteration count live on the
// device (PcgState); the host queues `check_every` iterations at a time and
// only then looks at one pinned word, so there is no per-iteration sync, and
// iterations queued past convergence return at their first instruction.
// Every reduction has a fixed order (common.cuh grid_sum_finish), hence
// bit-identical iterates and iteration counts run to run.
//
// SURVEY 8(f) row 4, matrices converted with B200_MAT_VALUES_F32 whose values do
// not all survive the rounding: iterative refinement (pcg_refine below) -- inner
// iterations stream the fp32 values, the residual b - A x the fp64 ones.
#include "common.cuh"

#include "pcg_kernels.cuh"

int spmv_full_internal(b200_mat *M, double *x_ext, double *y, bool dot,
                       const XrArgs *xr = nullptr, unsigned peer_seq = 0);

// residual replacements per solve at most (each one is an exit check that found
// ||b - A x|| above the bar although the recurrence was below it), and the
// iterations queued between two looks at the device afterwards
#define B2_MAX_REPLACEMENTS 4
#define B2_TAIL_CHUNK 4

// ---------------------------------------------------------------------------
int ensure_workspace(b200_mat *M) {
  if (M->state)
    return B200_OK;
  b200_ctx *c = M->ctx;
  uint64_t n = M->n_local, ne = n + M->halo.n_halo;
  B_TRY(dev_alloc(M, (void **)&M->w_r, (n + 2) * 8));
  B_TRY(dev_alloc(M, (void **)&M->w_p, (ne + 2) * 8));
  B_TRY(dev_alloc(M, (void **)&M->w_q, (n + 2) * 8));
  B_TRY(dev_alloc(M, (void **)&M->w_x, (n + 2) * 8));
  B_TRY(dev_alloc(M, (void **)&M->x_ext, (ne + 2) * 8));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_update,
                                                    EW_THREADS, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 4;
  if (per_sm > 8)
    per_sm = 8;
  uint64_t g = (uint64_t)c->sm_count * per_sm;
  uint64_t need = (n + EW_THREADS - 1) / EW_THREADS;
  if (g > need)
    g = need;
  M->grid_ew = g < 1 ? 1 : (int)g;
  // partial slots: the larger of the SpMV CTA count and the element-wise grid
  unsigned slots = (unsigned)c->sm_count * 32 * 3 + 64;
  M->partial_stride = slots;
  B_TRY(dev_alloc(M, (void **)&M->partials, (size_t)slots * 3 * 8));
  B_TRY(dev_alloc(M, (void **)&M->state, sizeof(PcgState)));
  CU_TRY(cudaMemsetAsync(M->state, 0, sizeof(PcgState), c->stream));
  // several ranks: map the neighbours' p vectors (collective: every rank gets here in
  // its first solve or product)
  B_TRY(halo_peer_setup(M));
  return B200_OK;
}

// where a reduction kernel stores its sum: straight into `red` on one rank,
// into the matching `loc` slot (all-reduced into `red` afterwards) on several
static double *sum_target(b200_mat *M, double *red, double *loc) {
  return M->ctx->nranks > 1 ? loc : red;
}
static int reduce_ranks(b200_mat *M, double *red, double *loc, int count) {
  if (M->ctx->nranks == 1)
    return B200_OK;
  return allreduce_sum(M->ctx, loc, red, count);
}

// seq: position of the iteration inside its chunk, 1-based (xr_chunk_begin)
static XrArgs xr_args(b200_ctx *c, int kind, unsigned seq) {
  XrArgs a = {nullptr, nullptr, c->nranks, c->rank, kind, 0ull, nullptr};
  if (c->xr_on)
    a.peers = c->xr_peers, a.mine = c->xr_mail, a.seq = seq, a.base = c->d_seq;
  return a;
}

// One iteration.  ev (4 events) brackets the three kernel classes when the
// caller times them.  With the peer-memory all-reduce the two NCCL calls
// disappear: K1's last CTA ships p.q to every rank and K2 collects it, K2's
// last CTA ships {r.z, r.r} and K3 collects them; the halo of p travels the
// same way (dist.cu), so nothing in here needs the host or NCCL.
static int queue_iteration(b200_mat *M, int par, unsigned seq, cudaEvent_t *ev = nullptr) {
  b200_ctx *c = M->ctx;
  cudaStream_t s = c->stream;
  uint64_t n = M->n_local;
  PcgState *st = M->state;
  const int nx = (par ^ 1) * 2;
  const XrArgs x_pq = xr_args(c, 0, seq), x_rz = xr_args(c, 1, seq);
  const XrArgs none = {nullptr, nullptr, 1, 0, 0, 0ull, nullptr};
  if (ev) CU_TRY(cudaEventRecord(ev[0], s));
  B_TRY(spmv_full_internal(M, M->w_p, M->w_q, true, c->xr_on ? &x_pq : nullptr,
                           c->xr_on ? seq : 0u));  // K1
  if (!c->xr_on)
    B_TRY(reduce_ranks(M, &st->pq, &st->pq_loc, 1));
  if (ev) CU_TRY(cudaEventRecord(ev[1], s));
  k_pcg_update<<<M->grid_ew, EW_THREADS, 0, s>>>(
      n, M->w_x, M->w_r, M->w_p, M->w_q, M->dinv, M->partials, M->partial_stride,
      st, par, sum_target(M, &st->red[nx], &st->loc[nx]), c->xr_on ? x_pq : none,
      c->xr_on ? x_rz : none);
  if (!c->xr_on)
    B_TRY(reduce_ranks(M, &st->red[nx], &st->loc[nx], 2));
  if (ev) CU_TRY(cudaEventRecord(ev[2], s));
  k_pcg_pupdate<<<M->grid_ew, EW_THREADS, 0, s>>>(n, M->w_r, M->dinv, M->w_p,
                                                  M->state, par, c->xr_on ? x_rz : none, 0);
  if (ev) CU_TRY(cudaEventRecord(ev[3], s));
  c->launches += 2;
  CU_TRY(cudaGetLastError());
  return B200_OK;
}

static int queue_chunk(b200_mat *M, int chunk) {
  B_TRY(xr_chunk_begin(M->ctx, (unsigned)chunk));
  for (int i = 0; i < chunk; i++)
    B_TRY(queue_iteration(M, i & 1, (unsigned)i + 1u));
  return B200_OK;
}

// One graph = `chunk` (even) iterations; replayed until the device says done.
static int ensure_graph(b200_mat *M, int chunk) {
  if (M->graph_exec && M->graph_chunk == chunk && M->graph_stream == (void *)M->ctx->stream)
    return B200_OK;
  if (M->graph_exec) {
    cudaGraphExecDestroy((cudaGraphExec_t)M->graph_exec);
    M->graph_exec = nullptr;
  }
  cudaStream_t s = M->ctx->stream;
  cudaGraph_t g = nullptr;
  CU_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = B200_OK;
  const uint64_t before = M->ctx->launches;
  rc = queue_chunk(M, chunk);
  cudaError_t e = cudaStreamEndCapture(s, &g);
  M->graph_kernels = (int)(M->ctx->launches - before);
  M->ctx->launches = before;  // captured, not launched
  if (rc != B200_OK)
    return rc;
  CU_TRY(e);
  cudaGraphExec_t ge = nullptr;
  CU_TRY(cudaGraphInstantiate(&ge, g, 0));
  cudaGraphDestroy(g);
  M->graph_exec = ge, M->graph_chunk = chunk, M->graph_stream = (void *)s;
  return B200_OK;
}

// The streaming solve: start-up, chunks of queued iterations, exit residual.
static int pcg_stream(b200_mat *M, const double *d_b, double *d_x,
                      const b200_pcg_opts *o, b200_pcg_result *res) {
  b200_ctx *c = M->ctx;
  B_TRY(ensure_workspace(M));
  cudaStream_t s = c->stream;
  const uint64_t n = M->n_local;
  const bool timing = o->flags & B200_PCG_TIME_KERNELS;
  int chunk = o->check_every > 0 ? o->check_every : 32;
  chunk = (chunk + 1) & ~1;  // even: the parity pattern repeats per chunk
  // (several ranks: a graph only when nothing in the iteration is an NCCL call, i.e. the
  // sums and the halo both go over peer memory)
  const bool use_graph = !(o->flags & B200_PCG_NO_GRAPH) &&
                         (c->nranks == 1 || (c->xr_on && M->halo.peer_ready));
  const uint64_t launches0 = c->launches;

  CU_TRY(cudaEventRecord(c->ev_a, s));
  // ---- start-up: r = b - A x0, p = D^-1 r ---------------------------------------
  CU_TRY(cudaMemcpyAsync(M->w_x, d_x, n * 8, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaMemcpyAsync(M->w_p, d_x, n * 8, cudaMemcpyDeviceToDevice, s));
  B_TRY(spmv_full_internal(M, M->w_p, M->w_q, false));
  k_pcg_init<<<M->grid_ew, EW_THREADS, 0, s>>>(n, d_b, M->w_q, M->dinv, M->w_r,
                                               M->w_p, M->partials,
                                               M->partial_stride, M->state,
                                               sum_target(M, &M->state->red[4], &M->state->loc[4]));
  B_TRY(reduce_ranks(M, &M->state->red[4], &M->state->loc[4], 3));
  k_pcg_start<<<1, 1, 0, s>>>(M->state, o->tol, o->maxit);
  c->launches += 2;
  CU_TRY(cudaGetLastError());

  // ---- iterations ---------------------------------------------------------------
  volatile int *flag = c->h_flag;  // {iter, done, status, maxit}
  int queued = 0;
  float t_cls[3] = {0, 0, 0};
  int timed_iters = 0;
  for (;;) {
    CU_TRY(cudaMemcpyAsync((void *)flag, &M->state->iter, 16, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaEventRecord(c->ev_poll, s));
    CU_TRY(cudaEventSynchronize(c->ev_poll));
    if (flag[1] || queued >= o->maxit)
      break;
    if (timing && timed_iters == 0) {
      // per-class device time over one chunk, events on the launch stream
      cudaEvent_t ev[4];
      for (auto &e : ev)
        CU_TRY(cudaEventCreate(&e));
      B_TRY(xr_chunk_begin(c, (unsigned)chunk));
      for (int i = 0; i < chunk; i++) {
        B_TRY(queue_iteration(M, i & 1, (unsigned)i + 1u, ev));
        CU_TRY(cudaEventSynchronize(ev[3]));
        for (int k = 0; k < 3; k++) {
          float t;
          CU_TRY(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
          t_cls[k] += t;
        }
      }
      timed_iters = chunk;
      for (auto &e : ev)
        cudaEventDestroy(e);
    } else if (use_graph) {
      B_TRY(ensure_graph(M, chunk));
      CU_TRY(cudaGraphLaunch((cudaGraphExec_t)M->graph_exec, s));
      c->launches += M->graph_kernels;
    } else {
      B_TRY(queue_chunk(M, chunk));
    }
    queued += chunk;
  }

  // ---- exit: true residual with one more SpMV; residual replacement -------------------
  // x goes through x_ext so that p survives the check.  If the recurrence said
  // "converged" and b - A x says "not quite" (drift after ~10^3 iterations), r is
  // replaced by the true residual, p keeps its direction, and the iteration goes
  // on (pcg_kernels.cuh k_pcg_replace) -- plain launches, a few iterations at a
  // time: this is the tail of a solve, not its body.
  PcgState h;
  int replacements = 0;
  bool stagnated = false;
  for (;;) {
    CU_TRY(cudaMemcpyAsync(M->x_ext, M->w_x, n * 8, cudaMemcpyDeviceToDevice, s));
    B_TRY(spmv_full_internal(M, M->x_ext, M->w_q, false));
    k_true_resid<<<M->grid_ew, EW_THREADS, 0, s>>>(
        n, d_b, M->w_q, M->partials, M->partial_stride, M->state,
        sum_target(M, &M->state->true_rr, &M->state->true_rr_loc));
    B_TRY(reduce_ranks(M, &M->state->true_rr, &M->state->true_rr_loc, 1));
    c->launches += 1;
    CU_TRY(cudaMemcpyAsync(&h, M->state, sizeof h, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (stagnated) {
      h.status = 4;
      break;
    }
    if (h.status != 0 || h.iter == 0 || !(h.true_rr > h.thr2) || h.iter >= o->maxit ||
        replacements >= B2_MAX_REPLACEMENTS)
      break;
    replacements++;
    const int par_last = (h.iter - 1) & 1, nx = (par_last ^ 1) * 2;
    k_pcg_replace<<<M->grid_ew, EW_THREADS, 0, s>>>(
        n, d_b, M->w_q, M->dinv, M->w_r, M->partials, M->partial_stride, M->state,
        sum_target(M, &M->state->red[nx], &M->state->loc[nx]));
    B_TRY(reduce_ranks(M, &M->state->red[nx], &M->state->loc[nx], 2));
    k_pcg_resume<<<1, 1, 0, s>>>(M->state, nx);
    const XrArgs none = {nullptr, nullptr, 1, 0, 0, 0ull, nullptr};
    k_pcg_pupdate<<<M->grid_ew, EW_THREADS, 0, s>>>(n, M->w_r, M->dinv, M->w_p, M->state,
                                                    par_last, none, 1);
    c->launches += 3;
    CU_TRY(cudaGetLastError());
    // (a replaced residual that is at the accuracy fp64 can reach for this system
    // never meets the bar: the tail is bounded, then status 4)
    const int tail = h.iter + (h.iter / 8 > 8 ? h.iter / 8 : 8);
    for (int q = h.iter;;) {
      B_TRY(xr_chunk_begin(c, B2_TAIL_CHUNK));
      for (int i = 0; i < B2_TAIL_CHUNK; i++, q++)
        B_TRY(queue_iteration(M, q & 1, (unsigned)i + 1u));
      CU_TRY(cudaMemcpyAsync((void *)flag, &M->state->iter, 16, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      if (flag[1])
        break;
      if (q >= tail) {
        stagnated = true;
        break;
      }
    }
  }
  CU_TRY(cudaMemcpyAsync(d_x, M->w_x, n * 8, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaEventRecord(c->ev_b, s));
  CU_TRY(cudaEventSynchronize(c->ev_b));
  CU_TRY(cudaEventElapsedTime(&res->solve_ms, c->ev_a, c->ev_b));
  res->replacements = replacements;

  const int parity = h.iter & 1;  // {rz, rr} of the last finished iteration
  double rr = h.iter == 0 ? h.red[1] : h.red[parity * 2 + 1];
  res->iters = h.iter, res->status = h.status;
  res->bnorm = sqrt(h.bb);
  res->relres = h.bb > 0 ? sqrt(rr / h.bb) : sqrt(rr);
  res->true_relres = h.bb > 0 ? sqrt(h.true_rr / h.bb) : sqrt(h.true_rr);
  res->kernel_launches = (int32_t)(c->launches - launches0);
  res->path = 0;
  if (timed_iters) {
    res->spmv_ms = t_cls[0] / timed_iters;
    res->update_ms = t_cls[1] / timed_iters;
    res->pupdate_ms = t_cls[2] / timed_iters;
  }
  if (h.status == 2)
    B_FAIL(B200_ENOTSPD, "b200_pcg_solve: breakdown at iteration %d (p.Ap = %g)",
           h.iter, h.pq);
  if (h.status == 3)
    B_FAIL(B200_ENCCL, "b200_pcg_solve: a rank's partial sums did not arrive over peer "
                       "memory (iteration %d)", h.iter);
  return B200_OK;
}

// SURVEY 8(f) row 4.  A B200_MAT_VALUES_F32 matrix whose values do not all
// survive the rounding keeps both value streams; the solve is iterative
// refinement to the same fp64 bar:
//   repeat  r = b - A x            fp64 values (spmv_use32 = false)
//           stop if ||r|| <= tol ||b||
//           solve A32 d = r        streaming PCG on the fp32 values, fp64 vectors,
//                                  to max(eta, tol ||b|| / (2 ||r||)), eta = 1e-4
//           x += d
// (tests hold it against a CPU statement of the same loop).
static int pcg_refine(b200_mat *M, const double *d_b, double *d_x,
                      const b200_pcg_opts *o, b200_pcg_result *res) {
  b200_ctx *c = M->ctx;
  B_TRY(ensure_workspace(M));
  cudaStream_t s = c->stream;
  const uint64_t n = M->n_local;
  if (!M->w_d) {
    B_TRY(dev_alloc(M, (void **)&M->w_d, (n + 2) * 8));
    B_TRY(dev_alloc(M, (void **)&M->w_rhs, (n + 2) * 8));
  }
  const double eta = 1e-4;
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  CU_TRY(cudaEventRecord(e0, s));
  const uint64_t launches0 = c->launches;
  int total = 0, passes = 0, status = 1, rc = B200_OK;
  double rr = 0, bb = 0;
  for (;;) {
    M->spmv_use32 = false;
    CU_TRY(cudaMemcpyAsync(M->w_p, d_x, n * 8, cudaMemcpyDeviceToDevice, s));
    B_TRY(spmv_full_internal(M, M->w_p, M->w_q, false));
    k_refine_resid<<<M->grid_ew, EW_THREADS, 0, s>>>(
        n, d_b, M->w_q, M->w_rhs, M->partials, M->partial_stride, M->state,
        sum_target(M, &M->state->red[4], &M->state->loc[4]));
    B_TRY(reduce_ranks(M, &M->state->red[4], &M->state->loc[4], 2));
    c->launches += 1;
    double h[2];
    CU_TRY(cudaMemcpyAsync(h, &M->state->red[4], 16, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    rr = h[0], bb = h[1];
    if (rr <= o->tol * o->tol * bb) {
      status = 0;
      break;
    }
    if (!(rr == rr)) {
      status = 2;
      break;
    }
    if (total >= o->maxit || passes >= 40)
      break;
    double t = 0.5 * o->tol * sqrt(bb / rr);
    b200_pcg_opts in = *o;
    in.tol = t < eta ? eta : t;
    in.maxit = o->maxit - total;
    in.flags &= ~(uint32_t)B200_PCG_TIME_KERNELS;
    b200_pcg_result ir;
    memset(&ir, 0, sizeof ir);
    CU_TRY(cudaMemsetAsync(M->w_d, 0, n * 8, s));
    M->spmv_use32 = true;
    rc = pcg_stream(M, M->w_rhs, M->w_d, &in, &ir);
    M->spmv_use32 = false;
    total += ir.iters, passes++;
    if (rc != B200_OK)
      break;
    k_add_into<<<M->grid_ew, EW_THREADS, 0, s>>>(n, d_x, M->w_d);
    c->launches += 1;
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaEventRecord(e1, s));
  CU_TRY(cudaEventSynchronize(e1));
  CU_TRY(cudaEventElapsedTime(&res->solve_ms, e0, e1));
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  res->iters = total, res->outer_iters = passes;
  res->status = rc == B200_ENOTSPD ? 2 : status;
  res->bnorm = sqrt(bb);
  res->relres = res->true_relres = bb > 0 ? sqrt(rr / bb) : sqrt(rr);
  res->kernel_launches = (int32_t)(c->launches - launches0);
  res->path = 0;
  if (rc != B200_OK)
    return rc;
  if (status == 2)
    B_FAIL(B200_ENOTSPD, "b200_pcg_solve: NaN residual in refinement pass %d", passes);
  return B200_OK;
}

extern "C" int b200_pcg_solve(b200_mat *M, const double *d_b, double *d_x,
                              const b200_pcg_opts *o, b200_pcg_result *res) {
  if (!M || !d_b || !d_x || !o || !res)
    B_FAIL(B200_EINVAL, "b200_pcg_solve: null argument");
  if (!(o->tol > 0.0) || o->maxit < 0)
    B_FAIL(B200_EINVAL, "b200_pcg_solve: tol=%g maxit=%d", o->tol, o->maxit);
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  memset(res, 0, sizeof *res);
  if (!M->blocks.empty())
    B_FAIL(B200_EINVAL, "b200_pcg_solve: a column-blocked matrix (B200_MAT_COL_BLOCK) is SpMV-only");
  if (!(o->flags & B200_PCG_NO_SMALL) && c->nranks == 1) {
    B_TRY(small_try_build(M));
    if (M->small)
      return small_solve(M, d_b, d_x, o, res);
  }
  if (M->sell_vals32 && M->sell_vals)  // rounded fp32 values: refinement
    return pcg_refine(M, d_b, d_x, o, res);
  return pcg_stream(M, d_b, d_x, o, res);
}

extern "C" int b200_pcg_solve_host(b200_mat *M, const double *h_b, double *h_x,
                                   const b200_pcg_opts *o, b200_pcg_result *res) {
  if (!M || !h_b || !h_x || !o || !res)
    B_FAIL(B200_EINVAL, "b200_pcg_solve_host: null argument");
  b200_ctx *c = M->ctx;
  CU_TRY(cudaSetDevice(c->device));
  const uint64_t n = M->n_local;
  // device staging for the host-buffer call shape, kept with the matrix so the
  // timed X_bench loop does not pay a cudaMalloc / cudaFree pair per solve
  if (!M->stage_b) {
    B_TRY(dev_alloc(M, (void **)&M->stage_b, (n + 1) * 8));
    B_TRY(dev_alloc(M, (void **)&M->stage_x, (n + 1) * 8));
  }
  double *d_b = M->stage_b, *d_x = M->stage_x;
  cudaStream_t s = c->stream;
  CU_TRY(cudaMemcpyAsync(d_b, h_b, n * 8, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(d_x, h_x, n * 8, cudaMemcpyHostToDevice, s));
  int rc = b200_pcg_solve(M, d_b, d_x, o, res);
  if (rc == B200_OK || rc == B200_ENOTSPD) {
    CU_TRY(cudaMemcpyAsync(h_x, d_x, n * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
  }
  return rc;
}
