// pcg_emul.cpp -- TEST INFRASTRUCTURE.  The product's Jacobi-PCG kernels
// (lsbench_b200/csrc/pcg_kernels.cuh) and its SELL SpMV with the fused dot
// product (sell_kernels.cuh), compiled for the host and run on the SIMT
// emulator of simt_emul.hpp.  The loop around them restates what pcg.cu queues
// (start-up, K1 K2 K3 per iteration, exit check and residual replacement);
// the arithmetic -- every fma, every reduction tree -- is the product's.
//   g++ -std=c++20 -O1 -DB2_SIMT_EMUL -I include -I lsbench_b200/csrc -I $CUDA/include
#include <cuda_runtime.h>  // first: its own declarations of threadIdx etc. stay untouched
#include "simt_emul.hpp"   // then the emulator's spellings ...
#include "common.cuh"      // ... for the device helpers (B2_SIMT_EMUL) and the kernels
#include "sell_kernels.cuh"
#include "pcg_kernels.cuh"

#include <algorithm>
#include <cstring>

struct Layout {  // the index-compressed SELL layout, built by the caller (numpy)
  uint32_t n, ns;
  const uint4 *meta;
  const uint32_t *cols;
  const int32_t *dcols;
  const double *vals;
  const uint32_t *perm;
  const double *dinv;
  const float *vals32;  // the same values as fp32 (exact), for the fp32-value kernels
  int kernel;           // 0 k_spmv_sellc fp64 | 1 k_spmv_sellc fp32
  int wmax;             // (unused)
};

static const XrArgs NOXR = {nullptr, nullptr, 1, 0, 0, 0ull};

template <bool DOT, typename VT>
static void plain_launch(const Layout &L, const VT *vals, unsigned grid, const double *x, double *y,
                         double *partials, PcgState *st) {
  simt::launch(grid, SPMV_THREADS, [&] {
    k_spmv_sellc<DOT, VT>(L.meta, L.cols, L.dcols, vals, L.perm, x, y, 0, L.ns, 0, 0, L.n,
                          DOT ? partials : nullptr, 0, DOT ? grid : 0u, DOT ? st : nullptr,
                          DOT ? &st->pq : nullptr, NOXR);
  });
}

static void spmv(const Layout &L, unsigned grid, const double *x, double *y, bool dot,
                 double *partials, PcgState *st) {
  switch (L.kernel * 2 + (dot ? 1 : 0)) {
  case 0: plain_launch<false, double>(L, L.vals, grid, x, y, partials, st); break;
  case 1: plain_launch<true, double>(L, L.vals, grid, x, y, partials, st); break;
  case 2: plain_launch<false, float>(L, L.vals32, grid, x, y, partials, st); break;
  case 3: plain_launch<true, float>(L, L.vals32, grid, x, y, partials, st); break;
  }
}

// returns 0; out: iters, status, relres (recurrence), x
extern "C" int emul_pcg(uint32_t n, uint32_t ns, const uint4 *meta, const uint32_t *cols,
                        const int32_t *dcols, const double *vals, const uint32_t *perm,
                        const double *dinv, const double *b, double *x, double tol, int maxit,
                        int /*unused*/, unsigned grid_spmv, unsigned grid_ew, int *iters,
                        int *status, double *relres, const float *vals32, int kernel, int wmax,
                        int *replacements, double *true_relres) {
  Layout L{n, ns, meta, cols, dcols, vals, perm, dinv, vals32, kernel, wmax};
  std::vector<double> r(n + 2), p(n + 2), q(n + 2);
  const unsigned stride = 148 * 32 * 3 + 64;
  std::vector<double> partials((size_t)stride * 3, 0.0);
  PcgState st;
  std::memset(&st, 0, sizeof st);
  // ---- start-up (pcg.cu pcg_stream) ---------------------------------------------------
  std::copy(x, x + n, p.begin());
  spmv(L, grid_spmv, p.data(), q.data(), false, nullptr, nullptr);
  simt::launch(grid_ew, EW_THREADS, [&] {
    k_pcg_init(n, b, q.data(), dinv, r.data(), p.data(), partials.data(), stride, &st, &st.red[4]);
  });
  simt::launch(1, 1, [&] { k_pcg_start(&st, tol, maxit); });
  // ---- iterations, in chunks as the product queues them -----------------------------------
  const int chunk = 8;
  int queued = 0;
  while (!st.done && queued < maxit) {
    for (int i = 0; i < chunk; i++) {
      const int par = i & 1, nx = (par ^ 1) * 2;
      {
        spmv(L, grid_spmv, p.data(), q.data(), true, partials.data(), &st);
        simt::launch(grid_ew, EW_THREADS, [&] {
          k_pcg_update(n, x, r.data(), p.data(), q.data(), dinv, partials.data(), stride, &st, par,
                       &st.red[nx], NOXR, NOXR);
        });
        simt::launch(grid_ew, EW_THREADS, [&] {
          k_pcg_pupdate(n, r.data(), dinv, p.data(), &st, par, NOXR, 0);
        });
      }
    }
    queued += chunk;
  }
  // ---- exit check and residual replacement (pcg.cu pcg_stream) --------------------------
  std::vector<double> xe(n + 2);
  int replaced = 0;
  bool stagnated = false;
  for (;;) {
    std::copy(x, x + n, xe.begin());
    spmv(L, grid_spmv, xe.data(), q.data(), false, nullptr, nullptr);
    simt::launch(grid_ew, EW_THREADS, [&] {
      k_true_resid(n, b, q.data(), partials.data(), stride, &st, &st.true_rr);
    });
    if (stagnated) {
      st.status = 4;
      break;
    }
    if (st.status != 0 || st.iter == 0 || !(st.true_rr > st.thr2) ||
        st.iter >= maxit || replaced >= 4)
      break;
    replaced++;
    const int par_last = (st.iter - 1) & 1, nx = (par_last ^ 1) * 2;
    simt::launch(grid_ew, EW_THREADS, [&] {
      k_pcg_replace(n, b, q.data(), dinv, r.data(), partials.data(), stride, &st, &st.red[nx]);
    });
    simt::launch(1, 1, [&] { k_pcg_resume(&st, nx); });
    simt::launch(grid_ew, EW_THREADS, [&] {
      k_pcg_pupdate(n, r.data(), dinv, p.data(), &st, par_last, NOXR, 1);
    });
    const int tail = st.iter + (st.iter / 8 > 8 ? st.iter / 8 : 8);
    for (int it = st.iter; !st.done; it++) {
      if (it >= tail && (it - st.iter) % 4 == 0) {  // the product looks every 4 iterations
        stagnated = true;
        break;
      }
      const int par = it & 1, nx2 = (par ^ 1) * 2;
      spmv(L, grid_spmv, p.data(), q.data(), true, partials.data(), &st);
      simt::launch(grid_ew, EW_THREADS, [&] {
        k_pcg_update(n, x, r.data(), p.data(), q.data(), dinv, partials.data(), stride, &st, par,
                     &st.red[nx2], NOXR, NOXR);
      });
      simt::launch(grid_ew, EW_THREADS, [&] {
        k_pcg_pupdate(n, r.data(), dinv, p.data(), &st, par, NOXR, 0);
      });
    }
  }
  if (replacements)
    *replacements = replaced;
  if (true_relres)
    *true_relres = st.bb > 0 ? std::sqrt(st.true_rr / st.bb) : std::sqrt(st.true_rr);
  const int parity = st.iter & 1;
  double rr = st.iter == 0 ? st.red[1] : st.red[parity * 2 + 1];
  *iters = st.iter, *status = st.status;
  *relres = st.bb > 0 ? std::sqrt(rr / st.bb) : std::sqrt(rr);
  return 0;
}

// ---- the row-major bins (power-law tail): k_spmv_vec, one warp per row, and
// k_spmv_long, one CTA per row; rows padded to multiples of 4 entries -------------
extern "C" int emul_rowmajor(int long_kernel, unsigned grid, uint32_t nrows, const uint32_t *ids,
                             const uint64_t *off, const uint32_t *cols, const double *vals,
                             const double *x, double *y, int dot, double *dot_out) {
  const unsigned stride = 148 * 32 * 3 + 64;
  std::vector<double> partials((size_t)stride * 3, 0.0);
  PcgState st;
  std::memset(&st, 0, sizeof st);
  simt::launch(grid, SPMV_THREADS, [&] {
    if (long_kernel && dot)
      k_spmv_long<true>(nrows, ids, off, cols, vals, x, y, partials.data(), 0, grid, &st, &st.pq, NOXR);
    else if (long_kernel)
      k_spmv_long<false>(nrows, ids, off, cols, vals, x, y, nullptr, 0, 0, nullptr, nullptr, NOXR);
    else if (dot)
      k_spmv_vec<true>(nrows, ids, off, cols, vals, x, y, partials.data(), 0, grid, &st, &st.pq, NOXR);
    else
      k_spmv_vec<false>(nrows, ids, off, cols, vals, x, y, nullptr, 0, 0, nullptr, nullptr, NOXR);
  });
  if (dot)
    *dot_out = st.pq;
  return 0;
}

// ---- the two-launch form of the overlapped multi-GPU SpMV: interior slices, then
// boundary slices, ONE fused dot product (the partial slots and the ticket span
// both launches; the last CTA of the second launch adds all of them) -----------------
extern "C" int emul_spmv_two_phase(uint32_t n, uint32_t ns, const uint4 *meta, const uint32_t *cols,
                                   const int32_t *dcols, const double *vals, const float *vals32,
                                   int kernel, int wmax, uint32_t ib, uint32_t ie, const double *x,
                                   double *y, double *dot_out) {
  Layout L{n, ns, meta, cols, dcols, vals, nullptr, nullptr, vals32, kernel, wmax};
  const unsigned stride = 148 * 32 * 3 + 64;
  std::vector<double> partials((size_t)stride * 3, 0.0);
  PcgState st;
  std::memset(&st, 0, sizeof st);
  {
    const unsigned g1 = (ie - ib + SPMV_WARPS - 1) / SPMV_WARPS,
                   g2 = (ib + (ns - ie) + SPMV_WARPS - 1) / SPMV_WARPS;
    simt::launch(g1, SPMV_THREADS, [&] {
      k_spmv_sellc<true, double>(meta, cols, dcols, vals, nullptr, x, y, ib, ie, 0, 0, n, partials.data(),
                                 0, g1 + g2, &st, &st.pq, NOXR);
    });
    simt::launch(g2, SPMV_THREADS, [&] {
      k_spmv_sellc<true, double>(meta, cols, dcols, vals, nullptr, x, y, 0, ib, ie, ns, n, partials.data(),
                                 g1, g1 + g2, &st, &st.pq, NOXR);
    });
  }
  *dot_out = st.pq;
  return st.ticket[0] == 0 ? 0 : 1;  // the last CTA must have reset the ticket
}

// ---- column blocking (csrc/colblock_kernels.cuh): cut a CSR into column ranges, then
// y = A_0 x; y += A_1 x; ... with the ACC instantiations of the SELL kernel ---------------
#include "colblock_kernels.cuh"

extern "C" int emul_colblock_split(uint64_t n, const uint64_t *offs, const uint32_t *cols,
                                   const double *vals, const uint64_t *cuts /* nb + 1, global order */,
                                   uint32_t nb, uint64_t n_own, uint64_t n_low,
                                   uint64_t *cnt /* nb x (n+1) */, uint64_t *boffs /* nb x (n+1) */,
                                   uint32_t **ocols, double **ovals, int stage, unsigned *unsorted) {
  if (stage == 0) {
    simt::launch((unsigned)((n + 1 + 255) / 256), 256,
                 [&] { k_colblock_count(n, offs, cols, cuts, nb, n_own, n_low, cnt); });
    simt::launch((unsigned)((n + 255) / 256), 256,
                 [&] { k_rows_sorted(n, offs, cols, n_own, n_low, unsorted); });
  } else {
    simt::launch(3, 256, [&] { k_colblock_fill(n, offs, cols, vals, nb, boffs, ocols, ovals); });
  }
  return 0;
}

// k_spmv_sell_grp as a column range runs it: a permuted slice list, and the rows of the
// row-major bins as units of the same work list (CTA-per-row and warp-per-row)
extern "C" int emul_sell_grp_bins(int acc, unsigned grid, uint32_t ns, const uint32_t *sell_off,
                                  const uint32_t *cols, const double *vals, const uint32_t *perm,
                                  const double *x, double *y, uint32_t n_rows, uint32_t long_rows,
                                  const uint32_t *long_ids, const uint64_t *long_off, uint32_t vec_rows,
                                  const uint32_t *vec_ids, const uint64_t *vec_off,
                                  const uint32_t *vl_cols, const double *vl_vals) {
  unsigned work[2] = {0, 0};
  if (acc)
    simt::launch(grid, SPMV_THREADS, [&] {
      k_spmv_sell_grp<true>(sell_off, cols, vals, perm, x, y, ns, n_rows, work, long_rows, long_ids, long_off,
                            vec_rows, vec_ids, vec_off, vl_cols, vl_vals);
    });
  else
    simt::launch(grid, SPMV_THREADS, [&] {
      k_spmv_sell_grp<false>(sell_off, cols, vals, perm, x, y, ns, n_rows, work, long_rows, long_ids, long_off,
                             vec_rows, vec_ids, vec_off, vl_cols, vl_vals);
    });
  return work[0] == 0 && work[1] == 0 ? 0 : 1;
}

// y (+)= A x with k_spmv_sell<false, double, ACC> on an explicit-column SELL layout
extern "C" int emul_sell_acc(int acc, unsigned grid, uint32_t ns, const uint32_t *sell_off,
                             const uint32_t *cols, const double *vals, const double *x, double *y,
                             uint32_t n_rows) {
  // bit 1: the grouped kernel (four slices per warp trip), what a column range runs by default
  static unsigned work[2] = {0, 0};
  if (acc & 2) {
    if (acc == 3)
      simt::launch(grid, SPMV_THREADS, [&] {
        k_spmv_sell_grp<true>(sell_off, cols, vals, nullptr, x, y, ns, n_rows, work, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
      });
    else
      simt::launch(grid, SPMV_THREADS, [&] {
        k_spmv_sell_grp<false>(sell_off, cols, vals, nullptr, x, y, ns, n_rows, work, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
      });
    return work[0] == 0 && work[1] == 0 ? 0 : 1;  // the last CTA out rearms the counters
  }
  else if (acc)
    simt::launch(grid, SPMV_THREADS, [&] {
      k_spmv_sell<false, double, true>(sell_off, cols, vals, nullptr, x, y, 0, ns, 0, 0, n_rows, nullptr,
                                       0, 0, nullptr, nullptr, NOXR);
    });
  else
    simt::launch(grid, SPMV_THREADS, [&] {
      k_spmv_sell<false, double, false>(sell_off, cols, vals, nullptr, x, y, 0, ns, 0, 0, n_rows, nullptr,
                                        0, 0, nullptr, nullptr, NOXR);
    });
  return 0;
}
