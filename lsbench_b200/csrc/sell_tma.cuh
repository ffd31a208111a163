// sell_tma.cuh -- the index-compressed SELL SpMV with the value stream fed by the
// bulk-copy engine (included by spmv.cu through sell_kernels.cuh).
//
// Why.  k_spmv_sellc is bound by latency, not by bytes: a warp asks for one chunk
// of 8 values per lane, waits a DRAM round trip, gathers, waits an L2 round trip,
// and only then asks again -- on average about half of its 2 KB are in flight.
// Round 1 measured it: DRAM-side 5.76 TB/s = 89 % of the copy peak on 27-point
// 512^3, 62 % of the warp slots active, long-scoreboard stalls 31 per issue, and
// storing the values as fp32 (42 % fewer bytes) made it 5 % SLOWER.  Holding more
// of the stream in registers costs occupancy and lost every time it was tried
// (next-chunk prefetch 8.0-8.4 ms against 5.99; whole-slice register pipelining,
// round 2: 0.64 ms against 0.29 at 192^3 -- profiles/r02_pipe_lost.txt).
//
// What.  The values of one slice are ONE contiguous run of w x 32 entries
// (column-major inside the slice), so one cp.async.bulk moves them into shared
// memory and counts its bytes on an mbarrier: no registers, no LSU slots, no
// scoreboard entries.  Every warp owns a ring of NSTAGES stages; lane 0 keeps
// NSTAGES slices requested ahead, all lanes wait on the stage's mbarrier, read
// their row's values from shared memory (one conflict-free 256-byte wavefront
// pair per k), gather x through L1/L2 and run the SAME fma chain in the SAME
// order as k_spmv_sellc -- the result has the same bits.  With 8 warps x 3 stages
// x 6.9 KB (27-point, fp64) an SM keeps 110-166 KB of the value stream in flight
// whatever the warps are doing; what 6.4 TB/s x 1.5 us asks for is 65 KB.
// Explicit (non-uniform) slices take their columns with streaming loads as
// before; uniform slices need no column stream at all.
//
// Ordering between the generic proxy (the lanes' reads of a stage) and the async
// proxy (the next bulk copy into it): __syncwarp after the last read, then
// fence.proxy.async by the issuing lane, then the copy.
#pragma once

#define TMA_MAX_STAGES 8

#ifndef B2_SIMT_EMUL
__device__ __forceinline__ uint32_t tma_s32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tma_bar_init(uint64_t *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tma_s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// arm the stage's barrier with the byte count and start the copy (one lane).  No proxy
// fence here: the stage being overwritten was READ by the lanes (values already in
// registers and multiplied) before the __syncwarp that precedes this call; a
// fence.proxy.async compiles to MEMBAR.ALL.CTA and would wait for every load the warp
// has in flight -- the gathers requested ahead for the next slice -- once per slice.
__device__ __forceinline__ void tma_fetch(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                          uint64_t policy) {
  const uint32_t b = tma_s32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  if (bytes)
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(tma_s32(dst)),
        "l"(src), "r"(bytes), "r"(b), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = tma_s32(bar);
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// the value stream is read once: first to go when L2 needs room (x stays)
__device__ __forceinline__ uint64_t tma_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
#else
// host emulation (tests/simt_emul.hpp): the copy happens at issue, the wait is a
// warp barrier -- lane 0 has issued before it arrives there
static inline void tma_bar_init(uint64_t *bar) { *bar = 0; }
static inline void tma_fence_init() {}
static inline void tma_fetch(void *dst, const void *src, uint32_t bytes, uint64_t *, uint64_t) {
  memcpy(dst, src, bytes);
}
static inline void tma_wait(uint64_t *, uint32_t) { __syncwarp(); }
static inline uint64_t tma_policy_stream() { return 0; }
#endif

// What the first version of this kernel taught (profiles/r02_tma_v1.txt): feeding the
// values alone is not enough.  Per slice a warp still paid, one after the other, the
// load of the slice's descriptor, the load of its deltas, two rounds of gathers and a
// fence behind the y store -- five to six dependent memory round trips of ~1.5 us each,
// ~10 us per slice per warp whatever the ring depth (the plain kernel pays ~8 us per
// slice per warp and hides it with 40 warps; this one has 8-16).  So everything a
// slice needs is now requested AHEAD of the slice, each level one step before the next:
//
//   step it:   descriptors  a window of 32 per warp (one lane each), refreshed every 32
//              G(it+1)      the w gathers of slice it+1, all in flight at once (+ x[row] for p.q)
//              D(it+2)      columns of slice it+2 (deltas: lane l holds delta l) and its rows
//              C(it)        wait for the values of slice it (bulk copy, requested S
//                           slices ago), w fma in row order, y
//              F(it+S)      request the values of slice it+S into the stage just freed
//
// so a warp's step costs one memory round trip, not six, and 8 warps keep up with the
// value stream.  Slices with explicit columns (12 - 25 % on a stencil) go the same way:
// their w columns are requested in D (registers), their gathers in G.
template <bool DOT, typename VT, int WARPS, int WCAP>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmv_sellc_tma(const uint4 *__restrict__ meta, const uint32_t *__restrict__ cols,
                 const int32_t *__restrict__ dcols, const VT *__restrict__ vals,
                 const uint32_t *__restrict__ perm, const double *__restrict__ x,
                 double *__restrict__ y, uint32_t b0, uint32_t e0, uint32_t b1, uint32_t e1,
                 uint32_t n_rows, double *partials, unsigned slot_base, unsigned total_slots,
                 PcgState *st, double *dot_out, const XrArgs xr, uint32_t stage_bytes,
                 int nstages) {
  static_assert(WCAP <= 32, "one delta per lane");
  // explicit slices: columns requested a step ahead into registers (8 warps), or loaded on
  // demand inside C (12 warps: 384 threads x two gather buffers leave no room for them)
  constexpr bool CBUF = WARPS <= 8;
  constexpr int NCB = CBUF ? WCAP : 1;
  if (DOT && st->done)
    return;
#ifdef B2_SIMT_EMUL
  static unsigned char *smem_raw = (unsigned char *)aligned_alloc(128, 232448);
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  __shared__ double red[WARPS];
  __shared__ uint64_t bars[WARPS * TMA_MAX_STAGES];
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n0 = e0 - b0, nv = n0 + (e1 - b1);
  const uint32_t stride = gridDim.x * WARPS;
  unsigned char *ring = smem_raw + (size_t)warp * nstages * stage_bytes;
  uint64_t *bar = bars + warp * TMA_MAX_STAGES;
  const uint64_t policy = tma_policy_stream();
  if (lane == 0)
    for (int i = 0; i < nstages; i++)
      tma_bar_init(bar + i);
  tma_fence_init();
  __syncwarp();
  const uint32_t v0 = blockIdx.x * WARPS + warp;
  // the warp's slices are it = 0 .. cnt-1: slice_of(v0 + it * stride)
  const uint32_t cnt = v0 < nv ? (nv - v0 + stride - 1) / stride : 0u;
  auto slice_of = [&](uint32_t it) {
    const uint32_t v = v0 + it * stride;
    return v < n0 ? b0 + v : b1 + (v - n0);
  };
  // ---- descriptors: two windows of 32, lane l holds the one of step wbase (+32) + l ------
  auto load_meta = [&](uint32_t it) {
    return it < cnt ? __ldg(meta + slice_of(it)) : make_uint4(0u, 0u, 0u, 0u);
  };
  uint4 mA = load_meta(lane), mB = load_meta(32 + lane);
  uint32_t wbase = 0;
  auto meta_at = [&](uint32_t it) {  // every lane calls it; it in [wbase, wbase + 64)
    const uint32_t t = it - wbase, l = t & 31u;
    const bool a = t < 32u;
    uint4 r;
    r.x = __shfl_sync(FULL, a ? mA.x : mB.x, l);
    r.y = __shfl_sync(FULL, a ? mA.y : mB.y, l);
    r.z = __shfl_sync(FULL, a ? mA.z : mB.z, l);
    r.w = 0u;
    return r;
  };
  // F: request the values of step `it` into `stage` (meta by all lanes, the copy by lane 0)
  auto fetch = [&](uint32_t it, int stage) {
    if (it >= cnt)
      return;
    const uint4 m = meta_at(it);
    if (lane == 0)
      tma_fetch(ring + (size_t)stage * stage_bytes, vals + (size_t)m.x * B2_SLICE,
                (m.y & 0x7fffffffu) * B2_SLICE * (uint32_t)sizeof(VT), bar + stage, policy);
  };
  // D: what the gathers of step `it` need -- its row, and the columns: a uniform slice's
  // deltas (lane l holds delta l), an explicit slice's own w columns in cb[]
  auto D = [&](uint32_t it, int32_t &dl, uint32_t &row, uint32_t (&cb)[NCB]) {
    dl = 0, row = 0xffffffffu;
    if (it >= cnt)
      return;
    const uint4 m = meta_at(it);
    const uint32_t w = m.y & 0x7fffffffu;
    const uint32_t pos = slice_of(it) * B2_SLICE + lane;
    row = perm ? __ldg(perm + pos) : pos;
    if (m.y >> 31) {
      if (lane < w)
        dl = __ldg(dcols + m.z + lane);
    } else if (CBUF && w) {
      const uint32_t *cp = cols + (size_t)m.z * B2_SLICE + lane;
#pragma unroll
      for (int k = 0; k < NCB; k++)  // (entry index clamped: no select on a loaded value)
        cb[k] = ld_stream(cp + (size_t)((uint32_t)k < w ? (uint32_t)k : w - 1u) * B2_SLICE);
    }
  };
  // G: all gathers of step `it`, in flight at once, and x[row] for the fused dot
  auto G = [&](uint32_t it, int32_t dl, uint32_t row, const uint32_t (&cb)[NCB], double (&xv)[WCAP],
               double &xrow) {
    if (it >= cnt)
      return;
    const uint4 m = meta_at(it);
    const uint32_t w = m.y & 0x7fffffffu;
    if (DOT)
      xrow = __ldg(x + (row < n_rows ? row : 0u));
    // NEVER select on the loaded value (k < w ? load : 0): the select consumes the load
    // where it stands and the 27 gathers go out one after the other, each a full round
    // trip -- that was the ~8 us per slice per warp of the first two versions (ncu source
    // page: every sample on the conditional moves behind the loads).  Select the ADDRESS:
    // entries past the slice's width read an address already read and are never used.
    if (m.y >> 31) {
#pragma unroll
      for (int k = 0; k < WCAP; k++) {
        const int32_t d = __shfl_sync(FULL, dl, k);
        xv[k] = __ldg(x + ((uint32_t)k < w ? row + (uint32_t)d : row));
      }
    } else if (CBUF && w) {
#pragma unroll
      for (int k = 0; k < NCB; k++)
        xv[k] = __ldg(x + cb[k]);
    }
  };
  for (int i = 0; i < nstages; i++)
    fetch((uint32_t)i, i);
  int32_t dl_a, dl_b;
  uint32_t row_c, row_a, row_b;
  uint32_t cb[NCB];
  double xvA[WCAP], xvB[WCAP], xrow_c = 0.0, xrow_n = 0.0;
#pragma unroll
  for (int k = 0; k < WCAP; k++)
    xvA[k] = 0.0, xvB[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NCB; k++)
    cb[k] = 0u;
  {
    int32_t dl_c;
    D(0u, dl_c, row_c, cb);
    G(0u, dl_c, row_c, cb, xvA, xrow_c);
    D(1u, dl_a, row_a, cb);
  }
  double dot = 0.0;
  int stage = 0;
  uint32_t phase = 0;
  auto step = [&](uint32_t it, double (&cur)[WCAP], double (&nxt)[WCAP]) {
    if (it - wbase >= 32u) {  // the next window of descriptors
      mA = mB, wbase += 32u;
      mB = load_meta(wbase + 32u + lane);
    }
    // gathers of the next slice (its columns were requested a step ago), then the columns
    // of the one after it (cb is free again once the gathers have their addresses)
    G(it + 1u, dl_a, row_a, cb, nxt, xrow_n);
    D(it + 2u, dl_b, row_b, cb);
    // ---- C(it) -----------------------------------------------------------------------------
    const uint4 m = meta_at(it);
    const uint32_t w = m.y & 0x7fffffffu;
    const VT *sv = reinterpret_cast<const VT *>(ring + (size_t)stage * stage_bytes) + lane;
    double sum = 0.0;
    if (CBUF || (m.y >> 31)) {
      tma_wait(bar + stage, phase);
#pragma unroll
      for (int k = 0; k < WCAP; k++)
        if ((uint32_t)k < w)
          sum = fma((double)sv[(size_t)k * B2_SLICE], cur[k], sum);
    } else {  // explicit columns on demand, 8 at a time
      const uint32_t *cp = cols + (size_t)m.z * B2_SLICE + lane;
      bool waited = false;
      for (uint32_t k = 0; k < w; k += 8) {
        const uint32_t nk = w - k < 8u ? w - k : 8u;
        uint32_t c[8];
        double xe[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
          c[j] = ld_stream(cp + (size_t)(k + ((uint32_t)j < nk ? j : nk - 1u)) * B2_SLICE);
#pragma unroll
        for (int j = 0; j < 8; j++)
          xe[j] = __ldg(x + c[j]);
        if (!waited) {
          tma_wait(bar + stage, phase);
          waited = true;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          if ((uint32_t)j < nk)
            sum = fma((double)sv[(size_t)(k + j) * B2_SLICE], xe[j], sum);
      }
      if (!waited)
        tma_wait(bar + stage, phase);  // an empty slice still completes its phase
    }
    // the stage is free once every lane has read its values: refill it, then write y
    __syncwarp();
    fetch(it + (uint32_t)nstages, stage);
    if (row_c < n_rows) {
      y[row_c] = sum;
      if (DOT)
        dot = fma(sum, xrow_c, dot);
    }
    if (++stage == nstages)
      stage = 0, phase ^= 1;
    row_c = row_a, dl_a = dl_b, row_a = row_b, xrow_c = xrow_n;
  };
  for (uint32_t it = 0; it < cnt; it += 2) {
    step(it, xvA, xvB);
    if (it + 1 < cnt)
      step(it + 1, xvB, xvA);
  }
  if (DOT) {
    double b[1] = {block_sum<WARPS>(dot, red)};
    grid_sum_finish<1, WARPS>(b, partials, 0, slot_base + blockIdx.x, total_slots, &st->ticket[0],
                              dot_out, red, xr);
  }
}
