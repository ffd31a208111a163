// small.cu -- the coarse-grid regime (BASELINE.json config 2: Nek matrices,
// n ~ 3.5-6.4 k, nnz ~ 77-146 k): the whole Jacobi-PCG solve in ONE kernel on
// ONE thread-block cluster, matrix and vectors resident in (distributed)
// shared memory, no global-memory traffic and no cluster-wide barrier inside
// the iteration.
//
// Why: with three streaming kernels per iteration a 300-iteration solve is
// ~900 launches -- launch-bound (3.4 ms measured) although the data is 1.7 MB.
// Here the matrix is split over the shared memory of the C CTAs of a cluster
// (C = 16, else 8); each CTA keeps x, r, D^-1 of its rows and p on the window
// of columns its rows touch.  CTAs talk through DSMEM only:
//
//   SpMV    q = A p on the local window, 4 lanes per row; p.q partial
//   -- all-reduce 1: every CTA writes its partial into every CTA's slot with
//      st.async (data + mbarrier transaction bytes in one DSMEM hop); a CTA
//      goes on when its own mbarrier has counted 8 C bytes --
//   update  x += a p, r -= a q, z = D^-1 r; partials of r.z and r.r; z is
//           pushed (st.async) into the window of every CTA that reads the row
//   -- all-reduce 2 + halo in one wait: 16 C + 8 col_n bytes on mbarrier B --
//   p       p = z + b p on the whole local window (redundantly, instead of a
//           third synchronisation)
//
// Measured per-iteration cost on tj7a_A_12 (clock64, B200_SMALL_PROFILE=1): the
// first version of this kernel -- p exchanged through L2, three cluster
// barriers -- spent 1400-1850 cycles in each barrier.cluster and 12 400 cycles
// per iteration.  Two mbarriers used in strict alternation are enough to
// order every reuse of the slots and windows (see the kernel).
//
// All sums have a fixed order (4-lane butterfly, warp butterfly, 32 warp sums
// by butterfly, <= 16 CTA sums by butterfly), so iterates and iteration counts
// are bit-reproducible.  Same recurrences and stopping rule as pcg.cu; stands
// where the timed solve of src/cholmod-impl.h:58-63 / src/cusparse.c:189-197
// stands for these matrices.
#include "common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

#define SM_THREADS 512
#define SM_WARPS (SM_THREADS / 32)
#define SM_REG_STEPS 24  // matrix entries per lane kept in registers
static_assert(SM_WARPS <= 32, "send_partials reduces the warp sums with one butterfly");
#define SM_MAX_CLUSTER 16
#define SM_MAX_ROWS 16384
#define SM_MAX_NNZ 600000

struct SmallCta {         // one per CTA, in global memory
  uint32_t n_groups;      // 8-row groups owned
  uint32_t n_rows;        // rows owned (<= 8 n_groups)
  uint32_t ent;           // padded entries = 32 * sum of group widths
  uint32_t col_lo, col_n; // staged column window [col_lo, col_lo + col_n)
  uint32_t goff_at;       // into goff[]   (n_groups + 1 values, units of 32 entries)
  uint32_t row_at;        // into rowid[] / dinv[] / dmask[] (8 n_groups values, 0xffffffff = none)
  uint32_t n_recv;        // window entries some row of this CTA reads (incl. its own rows)
  uint32_t row_lo;        // first row of the chunk (block-Jacobi: blocks are counted from here)
  uint32_t rest_lo;       // first entry kept in shared memory: what comes before it lives in
                          // registers only (regmat_load), in every warp
  uint64_t ent_at;        // into vals[] / cols[]
};

struct SmallPlan {
  int C = 0;
  uint32_t n = 0;
  size_t smem = 0;
  SmallCta *d_cta = nullptr;
  uint32_t *d_goff = nullptr, *d_rowid = nullptr, *d_dmask = nullptr;
  uint32_t *d_orig = nullptr, *d_perm = nullptr;  // slot -> caller's row; on-chip row -> caller's row
  bool reordered = false;
  double *d_vals = nullptr, *d_dinv = nullptr;
  uint16_t *d_cols = nullptr;
  long long *d_prof = nullptr;
  PcgState *d_state = nullptr;
  // Chebyshev-Jacobi: bound of the spectrum of D^-1 A, largest degree whose shared
  // memory fits, shared-memory bytes per degree
  double lmax = 0.0;
  int max_kdeg = 1;
  size_t smem_k[4] = {0, 0, 0, 0};
  // block-Jacobi: block size in use (0 = none fits), the inverted diagonal
  // blocks as fp32 rows of bj + 1 values per chunk row, chunk row -> slot, shared memory, and
  // the block every row of the caller's numbering belongs to (b200_mat_block_jacobi_partition)
  int bj = 0;
  uint32_t *d_binv = nullptr;  // high words of the fp64 entries
  uint16_t *d_slotof = nullptr;
  size_t smem_bj = 0;
  std::vector<uint32_t> block_of_row;
  // largest per-CTA extents (shared memory carve-up is uniform)
  uint32_t max_ent = 0, max_groups = 0, max_stage = 0;
};

struct SmemMap {
  size_t stage, zwin, dwin0, dwin1, vals, xs, rs, ds, qs, rhs, dds, zzs, rrow, slots, wred, bars, win, goff, rowid,
      orig, dmask, binv, slotof, cols, total;
};

// kdeg > 1 (Chebyshev-Jacobi): two more windows (the direction d of the polynomial
// recurrence lands in them alternately) and three more vectors on the owned rows
// bj > 0 (block-Jacobi): r once more in chunk-row order (blocks bj + 2 entries apart: 16-byte
// aligned, and the two blocks a warp may read start in different banks), the inverted blocks
// (rows bj + 1 words apart: a warp's 32 rows in 32 banks), chunk row -> slot
__host__ __device__ inline SmemMap smem_map(uint32_t max_ent, uint32_t max_groups,
                                            uint32_t max_stage, int kdeg, int bj = 0) {
  SmemMap m;
  size_t rows = (size_t)max_groups * 8, o = 0;
  const size_t win_bytes = ((size_t)max_stage + 1) / 2 * 2 * 8;
  m.stage = o, o += win_bytes;
  m.zwin = o, o += win_bytes;
  m.dwin0 = o, o += kdeg > 1 ? win_bytes : 0;
  m.dwin1 = o, o += kdeg > 2 ? win_bytes : 0;
  m.vals = o, o += (size_t)max_ent * 8;
  m.xs = o, o += rows * 8;
  m.rs = o, o += rows * 8;
  m.ds = o, o += rows * 8;
  m.qs = o, o += rows * 8;
  m.rhs = o, o += kdeg > 1 ? rows * 8 : 0;
  m.dds = o, o += kdeg > 1 ? rows * 8 : 0;
  m.zzs = o, o += kdeg > 1 ? rows * 8 : 0;
  o = (o + 15) / 16 * 16;
  m.rrow = o, o += bj ? (rows / bj + 2) * (size_t)(bj + 2) * 8 : 0;
  m.slots = o, o += 4 * SM_MAX_CLUSTER * 8;  // pq | rz | rr | bb
  m.wred = o, o += 3 * SM_WARPS * 8;
  m.bars = o, o += 4 * 8;
  m.win = o, o += 2 * SM_MAX_CLUSTER * 4;
  m.goff = o, o += ((size_t)max_groups + 2) * 4;
  m.rowid = o, o += rows * 4;
  m.orig = o, o += rows * 4;
  m.dmask = o, o += rows * 4;
  m.binv = o, o += bj ? rows * (size_t)(bj + 1) * 4 : 0;
  m.slotof = o, o += bj ? (rows * 2 + 7) / 8 * 8 : 0;
  m.cols = o, o += (size_t)max_ent * 2;
  m.total = (o + 15) / 16 * 16;
  return m;
}

// The SpMV of a CTA is shared-memory-bandwidth bound when values (8 B), columns
// (2 B) and the gathered vector entry (8 B) all come from shared memory
// (3 200 of 8 800 cycles per iteration on tj7a_A_12).  So each lane keeps the
// first SM_REG_STEPS entries it multiplies in registers: a warp's groups j =
// warp, warp + SM_WARPS, ... are covered in that order while they fit; the
// rest stay in shared memory.  Slots are walked with compile-time indices
// (full unroll) and warp-uniform group boundaries.
struct RegMat {
  double a[SM_REG_STEPS];
  uint32_t c[SM_REG_STEPS / 2];  // two 16-bit window offsets per register
  uint32_t nsteps;               // register slots in use (warp-uniform)
  uint32_t j_rest;               // first group of this warp left in shared memory
};

// (vals / cols: the CTA's entries in GLOBAL memory -- the entries a lane keeps in registers
// are not staged in shared memory at all: up to 24 x 512 x 10 bytes = 120 KB per CTA that
// the inverted blocks of block-Jacobi can use)
__device__ __forceinline__ void regmat_load(RegMat &R, const SmallCta &me, const uint32_t *goff,
                                            const double *__restrict__ vals,
                                            const uint16_t *__restrict__ cols) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t total = 0, j = warp;
  for (; j < me.n_groups; j += SM_WARPS) {
    const uint32_t w = goff[j + 1] - goff[j];
    if (total + w > SM_REG_STEPS)
      break;
    total += w;
  }
  R.nsteps = total, R.j_rest = j;
  uint32_t g = warp, o = g < me.n_groups ? goff[g] : 0u, w = g < me.n_groups ? goff[g + 1] - o : 1u, t = 0;
#pragma unroll
  for (int i = 0; i < SM_REG_STEPS; i++) {
    double a = 0.0;
    uint32_t c = 0;
    if ((uint32_t)i < total) {
      const uint32_t idx = (o + t) * 32 + lane;
      a = vals[idx], c = cols[idx];
      if (++t == w) {
        g += SM_WARPS, t = 0;
        if (g < me.n_groups)
          o = goff[g], w = goff[g + 1] - o;
      }
    }
    R.a[i] = a;
    if (i & 1)
      R.c[i / 2] |= c << 16;
    else
      R.c[i / 2] = c;
  }
}

// y[lr] = sum_k a(lr,k) * v[col(lr,k)] for the rows of this CTA; v staged in
// shared memory.  4 lanes per row; the row value lands in every lane of its
// group of four.  With DOT the lane that owns the row also returns its share
// of y . v (the CG p.Ap): rows it handled, in order.  Every group has at
// least one step (the host pads).
template <bool DOT>
__device__ __forceinline__ double small_spmv(const SmallCta &me, const RegMat &R,
                                             const uint32_t *goff, const double *vals,
                                             const uint16_t *cols, const uint32_t *rowid,
                                             const double *v_s, double *q_s) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double dot = 0.0;
  auto finish = [&](uint32_t j, double s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) {
      const uint32_t slot = j * 8 + (lane >> 2);
      q_s[slot] = s;
      if (DOT) {
        const uint32_t row = rowid[slot];
        if (row != 0xffffffffu)
          dot = fma(s, v_s[row - me.col_lo], dot);
      }
    }
  };
  // ---- groups held in registers ------------------------------------------------------
  // (Measured with clock64: the 64-bit gathers from the window are what this
  // loop waits for -- ~3.5 LDS-pipe cycles per warp gather with the bank
  // conflicts of eight unrelated rows -- so batching the loads or keeping the
  // values in shared memory changes nothing; registers only free the
  // shared-memory port for the gathers.)
  // (Round 2, ncu source page of this kernel: ~670 issued instructions per warp and
  // iteration for ~40 fma -- the walk counted steps against group widths re-read from
  // shared memory and branched twice per step.  Now a slot that is not in use multiplies
  // by a stored 0.0, and one precomputed mask bit per slot says where a group ends.)
  // (Round 2 tried a leaner walk -- unused slots multiply a stored 0.0, one mask bit per
  // slot says where a group ends, no width re-read: 670 -> ~450 issued instructions per
  // warp and iteration -- and measured it 1.5 - 7 % SLOWER on all seven Nek matrices
  // (A/B of two builds, gpurun_out/r02m_small_*): removed.)
  {
    uint32_t g = warp, w = g < me.n_groups ? goff[g + 1] - goff[g] : 1u, t = 0;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < SM_REG_STEPS; i++) {
      if ((uint32_t)i < R.nsteps) {
        const uint32_t c = (i & 1) ? R.c[i / 2] >> 16 : R.c[i / 2] & 0xffffu;
        s = fma(R.a[i], v_s[c], s);
        if (++t == w) {
          finish(g, s);
          s = 0.0, t = 0, g += SM_WARPS;
          if (g < me.n_groups)
            w = goff[g + 1] - goff[g];
        }
      }
    }
  }
  // ---- the rest, from shared memory ----------------------------------------------------
  for (uint32_t j = R.j_rest; j < me.n_groups; j += SM_WARPS) {
    const uint32_t o = goff[j], w = goff[j + 1] - o;
    const uint32_t base = o * 32 + lane - me.rest_lo;  // (shared memory starts at entry rest_lo)
    double s = 0.0;
    uint32_t t = 0;
    for (; t + 2 <= w; t += 2) {
      const double a0 = vals[base + t * 32], a1 = vals[base + t * 32 + 32];
      const double x0 = v_s[cols[base + t * 32]], x1 = v_s[cols[base + t * 32 + 32]];
      s = fma(a0, x0, s);
      s = fma(a1, x1, s);
    }
    if (t < w)
      s = fma(vals[base + t * 32], v_s[cols[base + t * 32]], s);
    finish(j, s);
  }
  return dot;
}

// ---- DSMEM / mbarrier primitives ----------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t s_remote(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
// the one arrival of a phase, announcing how many bytes st.async will deliver
__device__ __forceinline__ void bar_arm(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = s_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// 8 bytes into another CTA's shared memory; the same hop adds 8 to the
// transaction count of that CTA's mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                   remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}

// CTA sum of NV per-thread values, delivered to slot[v][my rank] of EVERY CTA of
// the cluster together with 8 NV transaction bytes on that CTA's `bar`.  Fixed
// butterflies at every level.  (The __syncthreads inside also publishes what the
// callers wrote to shared memory before the call.)
template <int NV>
__device__ __forceinline__ void send_partials(unsigned C, unsigned me, double (&val)[NV],
                                              double *wred, double *slots, int slot0,
                                              uint64_t *bar) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s = warp_sum(val[v]);
    if (lane == 0)
      wred[v * SM_WARPS + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t rbar = s_remote(s_u32(bar), lane < C ? lane : 0);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const double s = warp_sum(lane < SM_WARPS ? wred[v * SM_WARPS + lane] : 0.0);
      if (lane < C)
        st_async_f64(s_remote(s_u32(slots + (slot0 + v) * SM_MAX_CLUSTER + me), lane), s, rbar);
    }
  }
}
// after bar_wait: the <= 16 CTA sums by butterfly, identical in every thread
template <int NV>
__device__ __forceinline__ void read_totals(unsigned C, const double *slots, int slot0,
                                            double (&tot)[NV]) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s = (lane & 15u) < C ? slots[(slot0 + v) * SM_MAX_CLUSTER + (lane & 15u)] : 0.0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    tot[v] = s;
  }
}

// PROF: thread 0 of CTA 0 accumulates clock64() per phase into prof[0..5]
// (spmv, wait 1, update + push, wait 2, p window, -) -- B200_SMALL_PROFILE=1.
//
// KDEG: degree of the Chebyshev-Jacobi preconditioner (Saad, Iterative
// Methods, Alg. 12.1).  KDEG = 1 is plain Jacobi, z = D^-1 r.  KDEG = 2, 3:
//   rh = D^-1 r;  d = rh / theta;  z = d
//   KDEG - 1 times:  rh -= D^-1 (A d);  rho' = 1 / (2 sigma - rho);
//                    d = rho' rho d + (2 rho' / delta) rh;  z += d;  rho = rho'
// for the interval [lmax / 30, lmax] of D^-1 A: KDEG - 1 more products per iteration,
// each behind one more exchange of d between the CTAs that share columns (its own
// mbarrier, no all-reduce), and about 1 / KDEG of the iterations -- i.e. of the two
// all-reduces that every iteration of this kernel waits for.
//
// BJ: block-Jacobi.  z = B^-1 r with B the diagonal blocks of BJ consecutive
// rows of a CTA's chunk, inverted on the host, stored as fp32 (a preconditioner may be
// rounded: symmetric, still positive definite; the recurrences stay fp64) -- more exactly as
// the HIGH WORD of the fp64 value, rounded to nearest: 20 bits of mantissa, and widening it
// back is a register move, where an fp32 -> fp64 conversion issues at a quarter of the fma
// rate and was half of what this phase cost.  No exchange
// and no reduction is added: the rows of a block live in one CTA.  The update runs in two
// passes: x, r in slot order (r also into chunk-row order); then, one thread per chunk row --
// the lanes of a warp are then rows of the same one or two blocks and read the block's r as
// broadcasts -- z = B^-1 r, the push of z and the partial sums.
template <bool PROF, int KDEG, int BJ>
__global__ void __launch_bounds__(SM_THREADS, 1)
k_pcg_small(const SmallCta *__restrict__ ctas, const uint32_t *__restrict__ g_goff,
            const uint32_t *__restrict__ g_rowid, const uint32_t *__restrict__ g_dmask,
            const uint32_t *__restrict__ g_orig, const uint32_t *__restrict__ g_perm,
            const double *__restrict__ g_vals,
            const uint16_t *__restrict__ g_cols, const double *__restrict__ g_dinv,
            const uint32_t *__restrict__ g_binv, const uint16_t *__restrict__ g_slotof,
            const double *__restrict__ b, double *__restrict__ x, PcgState *st,
            uint32_t max_ent, uint32_t max_groups, uint32_t max_stage, double tol,
            int maxit, long long *prof, double cheb_theta, double cheb_delta) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned C = cl.num_blocks(), rank = cl.block_rank();
  const SmallCta me = ctas[rank];
  static_assert(BJ == 0 || KDEG == 1, "block-Jacobi and Chebyshev-Jacobi do not combine");
  const SmemMap mp = smem_map(max_ent, max_groups, max_stage, KDEG, BJ);
  double *p_w = (double *)(smem + mp.stage), *z_w = (double *)(smem + mp.zwin);
  double *d_w[2] = {(double *)(smem + mp.dwin0), (double *)(smem + (KDEG > 2 ? mp.dwin1 : mp.dwin0))};
  double *rh_s = (double *)(smem + mp.rhs), *dd_s = (double *)(smem + mp.dds), *zz_s = (double *)(smem + mp.zzs);
  double *vals = (double *)(smem + mp.vals);
  double *x_s = (double *)(smem + mp.xs), *r_s = (double *)(smem + mp.rs);
  double *d_s = (double *)(smem + mp.ds), *q_s = (double *)(smem + mp.qs);
  double *slots = (double *)(smem + mp.slots), *wred = (double *)(smem + mp.wred);
  uint64_t *bar_a = (uint64_t *)(smem + mp.bars), *bar_b = bar_a + 1, *bar_x = bar_a + 2;
  uint32_t *win_lo = (uint32_t *)(smem + mp.win), *win_n = win_lo + SM_MAX_CLUSTER;
  uint32_t *goff = (uint32_t *)(smem + mp.goff), *rowid = (uint32_t *)(smem + mp.rowid);
  uint32_t *dmask = (uint32_t *)(smem + mp.dmask), *orig = (uint32_t *)(smem + mp.orig);
  uint16_t *cols = (uint16_t *)(smem + mp.cols);
  double *rrow_s = (double *)(smem + mp.rrow);
  uint32_t *binv_s = (uint32_t *)(smem + mp.binv);
  uint16_t *slotof = (uint16_t *)(smem + mp.slotof);
  const uint32_t tid = threadIdx.x, nslot = me.n_groups * 8;
  long long pt[6] = {0, 0, 0, 0, 0, 0}, t0 = 0;
#define B2_TICK(i)                                                                \
  if (PROF) {                                                                     \
    long long t1 = clock64();                                                     \
    pt[i] += t1 - t0, t0 = t1;                                                    \
  }

  // ---- make the matrix resident ----------------------------------------------------
  for (uint32_t i = me.rest_lo + tid; i < me.ent; i += SM_THREADS)
    vals[i - me.rest_lo] = g_vals[me.ent_at + i], cols[i - me.rest_lo] = g_cols[me.ent_at + i];
  for (uint32_t i = tid; i <= me.n_groups; i += SM_THREADS)
    goff[i] = g_goff[me.goff_at + i];
  for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
    uint32_t row = g_rowid[me.row_at + i];
    const uint32_t og = g_orig[me.row_at + i];  // the caller's numbering, for b and x
    rowid[i] = row, orig[i] = og;
    dmask[i] = g_dmask[me.row_at + i];
    d_s[i] = g_dinv[me.row_at + i];
    x_s[i] = row != 0xffffffffu ? x[og] : 0.0;
  }
  if (BJ) {
    for (uint32_t i = tid; i < me.n_rows * (BJ + 1); i += SM_THREADS)
      binv_s[i] = g_binv[(size_t)me.row_at * (BJ + 1) + i];
    for (uint32_t i = tid; i < me.n_rows; i += SM_THREADS)
      slotof[i] = g_slotof[me.row_at + i];
    for (uint32_t i = tid; i < (nslot / (BJ ? BJ : 1) + 2) * (BJ + 2); i += SM_THREADS)
      rrow_s[i] = 0.0;  // (rows past the end of the last block stay zero)
  }
  if (tid < SM_MAX_CLUSTER)  // window base of every CTA, as a byte offset into its z window
    win_lo[tid] = tid < C ? ctas[tid].col_lo : 0u, win_n[tid] = tid < C ? ctas[tid].col_n : 0u;
  for (uint32_t i = tid; i < me.col_n; i += SM_THREADS)
    p_w[i] = x[g_perm[me.col_lo + i]], z_w[i] = 0.0;  // the window of x0, for r = b - A x0
  if (tid == 0) {
    bar_init(bar_a, 1), bar_init(bar_b, 1), bar_init(bar_x, 1), bar_init(bar_x + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  RegMat R;
  regmat_load(R, me, goff, g_vals + me.ent_at, g_cols + me.ent_at);
  cl.sync();  // every CTA's barriers exist before anyone stores into a neighbour

  // Two mbarriers, used in strict alternation B A B A ... B | B A.  A CTA sends
  // the data of phase k+1 of one barrier only after it has passed phase k of
  // the other one, which needed a contribution every CTA sends only after it
  // has itself passed phase k of the first: so no store can land in a phase,
  // a slot or a window that its receiver has not finished with.
  uint32_t par_a = 0, par_b = 0, par_x[2] = {0, 0};
  const uint32_t my_bar_b = s_u32(bar_b), my_zw = s_u32(z_w);
  // the value of owned row `row`, into the window `win` of every CTA that reads it,
  // counted on that CTA's barrier `bar` (mask: bit k set when some row of CTA k has
  // an entry in column `row`, or owns it)
  auto push_to = [&](uint32_t win, uint32_t bar, uint32_t row, uint32_t mask, double v) {
    while (mask) {
      const unsigned k = __ffs(mask) - 1;
      mask &= mask - 1;
      st_async_f64(s_remote(win + (row - win_lo[k]) * 8u, k), v, s_remote(bar, k));
    }
  };
  auto push = [&](uint32_t row, uint32_t mask, double v) { push_to(my_zw, my_bar_b, row, mask, v); };
  // Chebyshev-Jacobi: zz_s = P(D^-1 A) D^-1 r_s on the owned rows (KDEG > 1 only).  Every
  // exchange of d has its own barrier, used once per application; two applications are
  // always separated by an all-to-all phase (barrier A or B), so no store of the next
  // one can land in a phase or a window its receiver has not finished with.
  const double cheb_sigma = cheb_theta / cheb_delta;
  auto chebyshev = [&]() {
    double rho = 1.0 / cheb_sigma;
    if (tid == 0)
      bar_arm(bar_x, me.n_recv * 8);
    for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
      const uint32_t row = rowid[i];
      if (row == 0xffffffffu)
        continue;
      const double c = d_s[i] * r_s[i], dd = c / cheb_theta;
      rh_s[i] = c, dd_s[i] = dd, zz_s[i] = dd;
      push_to(s_u32(d_w[0]), s_u32(bar_x), row, dmask[i], dd);
    }
#pragma unroll
    for (int j = 1; j < KDEG; j++) {
      bar_wait(bar_x + (j - 1), par_x[j - 1]), par_x[j - 1] ^= 1;
      small_spmv<false>(me, R, goff, vals, cols, rowid, d_w[(j - 1) & 1], q_s);  // t = A d
      __syncthreads();
      const double rhon = 1.0 / (2.0 * cheb_sigma - rho);
      if (j + 1 < KDEG && tid == 0)
        bar_arm(bar_x + j, me.n_recv * 8);
      for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
        const uint32_t row = rowid[i];
        if (row == 0xffffffffu)
          continue;
        const double rh = fma(-d_s[i], q_s[i], rh_s[i]);
        const double dd = fma(rhon * rho, dd_s[i], (2.0 * rhon / cheb_delta) * rh);
        rh_s[i] = rh, dd_s[i] = dd, zz_s[i] += dd;
        if (j + 1 < KDEG)
          push_to(s_u32(d_w[j & 1]), s_u32(bar_x + j), row, dmask[i], dd);
      }
      rho = rhon;
    }
    __syncthreads();
  };

  // block-Jacobi, second pass (after a __syncthreads behind the pass that wrote r_s and
  // rrow_s): thread t takes chunk row t; adds its r.z and r.r to acc[0], acc[1]
  auto rrow_at = [&](uint32_t lr) { return lr + 2 * (lr / (BJ ? BJ : 1)); };  // blocks BJ + 2 apart
  auto block_jacobi = [&](double *acc) {
    for (uint32_t lr = tid; lr < me.n_rows; lr += SM_THREADS) {
      const uint32_t i = slotof[lr];
      const uint32_t *bi = binv_s + (size_t)lr * (BJ + 1);
      const double2 *rb = reinterpret_cast<const double2 *>(rrow_s + (size_t)(lr / (BJ ? BJ : 1)) * (BJ + 2));
      double zi = 0.0;
#pragma unroll
      for (int j = 0; j < BJ; j += 2) {
        const double2 rj = rb[j / 2];  // one 16-byte broadcast per pair
        zi = fma(__hiloint2double((int)bi[j], 0), rj.x, zi);
        zi = fma(__hiloint2double((int)bi[j + 1], 0), rj.y, zi);
      }
      const double ri = r_s[i];
      push(me.row_lo + lr, dmask[i], zi);
      acc[0] = fma(ri, zi, acc[0]), acc[1] = fma(ri, ri, acc[1]);
    }
  };

  // ---- r = b - A x0, z = D^-1 r, p = z ----------------------------------------------
  small_spmv<false>(me, R, goff, vals, cols, rowid, p_w, q_s);
  __syncthreads();
  if (tid == 0)
    bar_arm(bar_b, (3 * C + me.n_recv) * 8);
  double acc3[3] = {0.0, 0.0, 0.0}, tot3[3];
  if (KDEG > 1) {
    for (uint32_t i = tid; i < nslot; i += SM_THREADS)
      if (rowid[i] != 0xffffffffu)
        r_s[i] = b[orig[i]] - q_s[i];
    __syncthreads();
    chebyshev();
  }
  if (BJ) {
    for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
      uint32_t row = rowid[i];
      if (row == 0xffffffffu)
        continue;
      const double bi = b[orig[i]], ri = bi - q_s[i];
      r_s[i] = ri, rrow_s[rrow_at(row - me.row_lo)] = ri;
      acc3[2] = fma(bi, bi, acc3[2]);
    }
    __syncthreads();
    block_jacobi(acc3);
  } else
  for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
    uint32_t row = rowid[i];
    if (row == 0xffffffffu)
      continue;
    double bi = b[orig[i]], ri = bi - q_s[i], zi = KDEG > 1 ? zz_s[i] : d_s[i] * ri;
    if (KDEG > 1)
      ri = r_s[i];
    r_s[i] = ri;
    push(row, dmask[i], zi);
    acc3[0] = fma(ri, zi, acc3[0]), acc3[1] = fma(ri, ri, acc3[1]), acc3[2] = fma(bi, bi, acc3[2]);
  }
  send_partials<3>(C, rank, acc3, wred, slots, 1, bar_b);  // slots 1,2,3 = rz, rr, bb
  bar_wait(bar_b, par_b), par_b ^= 1;
  read_totals<3>(C, slots, 1, tot3);
  for (uint32_t i = tid; i < me.col_n; i += SM_THREADS)
    p_w[i] = z_w[i];
  __syncthreads();
  double rz = tot3[0], rr = tot3[1];
  const double bb = tot3[2], thr2 = tol * tol * bb;
  int it = 0, status = 1;
  double pq = 0.0;
  if (rr <= thr2)
    status = 0;

  // ---- iterations ----------------------------------------------------------------------
  if (PROF)
    t0 = clock64();
  int replacements = 0, it_stop = maxit;
  double t4[1];
  for (;;) {  // iterate; check b - A x; go on from the true residual if it misses the bar
  while (status == 1 && it < it_stop) {
    if (tid == 0)
      bar_arm(bar_a, C * 8);
    double a1[1], t1[1];
    a1[0] = small_spmv<true>(me, R, goff, vals, cols, rowid, p_w, q_s);
    B2_TICK(0)
    send_partials<1>(C, rank, a1, wred, slots, 0, bar_a);
    bar_wait(bar_a, par_a), par_a ^= 1;
    read_totals<1>(C, slots, 0, t1);
    B2_TICK(1)
    pq = t1[0];
    if (!(pq > 0.0)) {
      status = 2;
      break;
    }
    const double alpha = rz / pq;
    if (tid == 0)
      bar_arm(bar_b, (2 * C + me.n_recv) * 8);
    double a2[2] = {0.0, 0.0}, t2[2];
    if (BJ) {
      for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
        uint32_t row = rowid[i];
        if (row == 0xffffffffu)
          continue;
        x_s[i] = fma(alpha, p_w[row - me.col_lo], x_s[i]);
        const double ri = fma(-alpha, q_s[i], r_s[i]);
        r_s[i] = ri, rrow_s[rrow_at(row - me.row_lo)] = ri;
      }
      __syncthreads();
      block_jacobi(a2);
    } else if (KDEG == 1) {
      for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
        uint32_t row = rowid[i];
        if (row == 0xffffffffu)
          continue;
        double pi = p_w[row - me.col_lo];
        x_s[i] = fma(alpha, pi, x_s[i]);
        double ri = fma(-alpha, q_s[i], r_s[i]);
        r_s[i] = ri;
        const double zi = d_s[i] * ri;
        push(row, dmask[i], zi);
        a2[0] = fma(ri, zi, a2[0]), a2[1] = fma(ri, ri, a2[1]);
      }
    } else {
      for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
        uint32_t row = rowid[i];
        if (row == 0xffffffffu)
          continue;
        x_s[i] = fma(alpha, p_w[row - me.col_lo], x_s[i]);
        r_s[i] = fma(-alpha, q_s[i], r_s[i]);
      }
      __syncthreads();
      chebyshev();
      for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
        uint32_t row = rowid[i];
        if (row == 0xffffffffu)
          continue;
        const double ri = r_s[i], zi = zz_s[i];
        push(row, dmask[i], zi);
        a2[0] = fma(ri, zi, a2[0]), a2[1] = fma(ri, ri, a2[1]);
      }
    }
    B2_TICK(2)
    send_partials<2>(C, rank, a2, wred, slots, 1, bar_b);
    bar_wait(bar_b, par_b), par_b ^= 1;
    read_totals<2>(C, slots, 1, t2);
    B2_TICK(3)
    const double rzn = t2[0];
    rr = t2[1];
    it++;
    if (rr <= thr2) {
      status = 0;
      break;
    }
    if (!(rr == rr)) {
      status = 2;
      break;
    }
    const double beta = rzn / rz;
    rz = rzn;
    for (uint32_t i = tid; i < me.col_n; i += SM_THREADS)
      p_w[i] = fma(beta, p_w[i], z_w[i]);  // p = z + b p on the whole window
    __syncthreads();
    B2_TICK(4)
  }
#undef B2_TICK
  if (status == 1 && it < maxit)
    status = 4;  // the bounded tail after a replacement did not get there: stagnated

  // ---- x out, true residual ---------------------------------------------------------------
  cl.sync();  // the loop may end on barrier B; the next use below is B again
  if (tid == 0)
    bar_arm(bar_b, me.n_recv * 8);
  for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
    uint32_t row = rowid[i];
    if (row != 0xffffffffu) {
      x[orig[i]] = x_s[i];
      push(row, dmask[i], x_s[i]);
    }
  }
  bar_wait(bar_b, par_b), par_b ^= 1;
  if (tid == 0)
    bar_arm(bar_a, C * 8);
  small_spmv<false>(me, R, goff, vals, cols, rowid, z_w, q_s);
  __syncthreads();
  double a4[1] = {0.0};
  for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
    uint32_t row = rowid[i];
    if (row != 0xffffffffu) {
      double d = b[orig[i]] - q_s[i];
      a4[0] = fma(d, d, a4[0]);
    }
  }
  send_partials<1>(C, rank, a4, wred, slots, 0, bar_a);
  bar_wait(bar_a, par_a), par_a ^= 1;
  read_totals<1>(C, slots, 0, t4);
  // Residual replacement (pcg_kernels.cuh k_pcg_replace, same rule as the streaming
  // path): the recurrence says converged, b - A x does not.  q_s holds A x: take
  // r = b - q, push z = D^-1 r, form r.z and r.r, p = z + (r.z / r.z_old) p with
  // the direction of the last iteration (rz still is r.z_old: the loop left
  // before overwriting it), and iterate on.  Barrier order stays B | B A | B, A B ...
  if (!(status == 0 && it > 0 && it < maxit && t4[0] > thr2 && replacements < 4))
    break;
  replacements++;
  if (tid == 0)
    bar_arm(bar_b, (2 * C + me.n_recv) * 8);
  double a5[2] = {0.0, 0.0}, t5[2];
  if (KDEG > 1) {
    for (uint32_t i = tid; i < nslot; i += SM_THREADS)
      if (rowid[i] != 0xffffffffu)
        r_s[i] = b[orig[i]] - q_s[i];
    __syncthreads();
    chebyshev();
  }
  if (BJ) {
    for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
      uint32_t row = rowid[i];
      if (row == 0xffffffffu)
        continue;
      const double ri = b[orig[i]] - q_s[i];
      r_s[i] = ri, rrow_s[rrow_at(row - me.row_lo)] = ri;
    }
    __syncthreads();
    block_jacobi(a5);
  } else
  for (uint32_t i = tid; i < nslot; i += SM_THREADS) {
    uint32_t row = rowid[i];
    if (row == 0xffffffffu)
      continue;
    const double ri = KDEG > 1 ? r_s[i] : b[orig[i]] - q_s[i], zi = KDEG > 1 ? zz_s[i] : d_s[i] * ri;
    r_s[i] = ri;
    push(row, dmask[i], zi);
    a5[0] = fma(ri, zi, a5[0]), a5[1] = fma(ri, ri, a5[1]);
  }
  send_partials<2>(C, rank, a5, wred, slots, 1, bar_b);
  bar_wait(bar_b, par_b), par_b ^= 1;
  read_totals<2>(C, slots, 1, t5);
  rr = t5[1];
  if (rr <= thr2)
    break;  // (cannot happen: the same sums as t4; keeps the loop honest)
  {
    const double beta = t5[0] / rz;
    rz = t5[0];
    for (uint32_t i = tid; i < me.col_n; i += SM_THREADS)
      p_w[i] = fma(beta, p_w[i], z_w[i]);
    __syncthreads();
  }
  status = 1;
  it_stop = it + (it / 8 > 8 ? it / 8 : 8);
  it_stop = it_stop < maxit ? it_stop : maxit;
  if (PROF)
    t0 = clock64();
  }  // for (;;)
  if (PROF && rank == 0 && tid == 0)
    for (int i = 0; i < 6; i++)
      prof[i] = pt[i];
  if (rank == 0 && tid == 0) {
    st->iter = it, st->status = status, st->done = 1;
    st->bb = bb, st->red[1] = rr, st->pq = pq, st->true_rr = t4[0];
    st->replacements = replacements;
  }
  cl.sync();  // nobody leaves while a neighbour may still be storing into it
}

// the instantiations: profiling on / off x preconditioner degree 1..3
typedef void (*small_kernel_t)(const SmallCta *, const uint32_t *, const uint32_t *, const uint32_t *,
                               const uint32_t *, const uint32_t *, const double *, const uint16_t *,
                               const double *, const uint32_t *, const uint16_t *, const double *, double *,
                               PcgState *, uint32_t, uint32_t, uint32_t, double, int, long long *, double,
                               double);
static small_kernel_t small_kernel(bool prof, int kdeg, int bj = 0) {
  if (bj == 32)
    return prof ? k_pcg_small<true, 1, 32> : k_pcg_small<false, 1, 32>;
  if (bj == 16)
    return prof ? k_pcg_small<true, 1, 16> : k_pcg_small<false, 1, 16>;
  if (prof)
    return kdeg == 3 ? k_pcg_small<true, 3, 0> : kdeg == 2 ? k_pcg_small<true, 2, 0> : k_pcg_small<true, 1, 0>;
  return kdeg == 3 ? k_pcg_small<false, 3, 0> : kdeg == 2 ? k_pcg_small<false, 2, 0> : k_pcg_small<false, 1, 0>;
}

// Upper bound of the spectrum of D^-1 A for the Chebyshev interval: min(Gershgorin
// bound, 1.15 x the power-iteration estimate after 40 steps from the vector of ones)
// (the checker restates the same rule).
static double cheb_lmax(uint64_t n, const std::vector<uint64_t> &offs, const std::vector<uint32_t> &cols,
                        const std::vector<double> &vals, const std::vector<double> &dinv) {
  double gersh = 0.0;
  for (uint64_t i = 0; i < n; i++) {
    double t = 0.0;
    for (uint64_t e = offs[i]; e < offs[i + 1]; e++)
      t += fabs(vals[e]);
    t *= fabs(dinv[i]);
    gersh = t > gersh ? t : gersh;
  }
  std::vector<double> v(n, 1.0), w(n);
  double lam = 0.0;
  for (int it = 0; it < 40; it++) {
    double nw = 0.0, nv = 0.0;
    for (uint64_t i = 0; i < n; i++) {
      double t = 0.0;
      for (uint64_t e = offs[i]; e < offs[i + 1]; e++)
        t += vals[e] * v[cols[e]];
      w[i] = t * dinv[i];
      nw += w[i] * w[i], nv += v[i] * v[i];
    }
    lam = sqrt(nw / nv);
    const double sc = 1.0 / sqrt(nw);
    for (uint64_t i = 0; i < n; i++)
      v[i] = w[i] * sc;
  }
  const double est = 1.15 * lam;
  return est < gersh ? est : gersh;
}

// ---------------------------------------------------------------------------
static void small_free_plan(b200_mat *M, void *&slot);
void small_free(b200_mat *M) {
  small_free_plan(M, M->small);
  small_free_plan(M, M->small_bj);
}
static void small_free_plan(b200_mat *M, void *&slot) {
  SmallPlan *P = (SmallPlan *)slot;
  if (!P)
    return;
  void *ptrs[] = {P->d_cta, P->d_goff, P->d_rowid, P->d_vals, P->d_dinv, P->d_cols,
                  P->d_state, P->d_prof, P->d_dmask, P->d_orig, P->d_perm, P->d_binv, P->d_slotof};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete P;
  slot = nullptr;
}

static int small_launch_config(SmallPlan *P, cudaLaunchConfig_t *cfg,
                               cudaLaunchAttribute *attr, cudaStream_t s) {
  memset(cfg, 0, sizeof *cfg);
  cfg->gridDim = dim3(P->C), cfg->blockDim = dim3(SM_THREADS);
  cfg->dynamicSmemBytes = P->smem, cfg->stream = s;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P->C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr, cfg->numAttrs = 1;
  return B200_OK;
}

// Reverse Cuthill-McKee on the pattern of A + A^T (used here for
// one purpose: a CTA's window is [min column, max column] of its row chunk, and
// both the shared memory it takes and the z entries pushed to it every
// iteration grow with the bandwidth -- xn3b has bandwidth 2 343 at n = 6 408).
// Returns perm[new] = old.  Deterministic: ties by original index.
static std::vector<uint32_t> rcm_order(uint64_t n, const std::vector<uint64_t> &offs,
                                       const std::vector<uint32_t> &cols) {
  std::vector<std::vector<uint32_t>> adj(n);
  for (uint64_t r = 0; r < n; r++)
    for (uint64_t e = offs[r]; e < offs[r + 1]; e++)
      if (cols[e] != r && cols[e] < n)
        adj[r].push_back(cols[e]), adj[cols[e]].push_back((uint32_t)r);
  for (auto &a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  auto by_degree = [&](uint32_t a, uint32_t b) {
    return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
  };
  std::vector<uint32_t> order, byd(n);
  std::vector<char> seen(n, 0);
  order.reserve(n);
  for (uint64_t i = 0; i < n; i++)
    byd[i] = (uint32_t)i;
  std::sort(byd.begin(), byd.end(), by_degree);
  for (uint32_t start : byd) {  // one BFS per connected component, lowest degree first
    if (seen[start])
      continue;
    size_t head = order.size();
    order.push_back(start), seen[start] = 1;
    while (head < order.size()) {
      const uint32_t u = order[head++];
      std::vector<uint32_t> nb;
      for (uint32_t v : adj[u])
        if (!seen[v])
          nb.push_back(v), seen[v] = 1;
      std::sort(nb.begin(), nb.end(), by_degree);
      order.insert(order.end(), nb.begin(), nb.end());
    }
  }
  std::reverse(order.begin(), order.end());
  return order;
}

// sum over the C row chunks of the column window each one needs
static uint64_t window_cost(uint64_t n, int C, const std::vector<uint64_t> &offs,
                            const std::vector<uint32_t> &cols) {
  uint64_t total = 0;
  for (int k = 0; k < C; k++) {
    uint64_t r0 = n * k / C, r1 = n * (k + 1) / C;
    if (r1 == r0)
      continue;
    uint64_t lo = r0, hi = r1 - 1;
    for (uint64_t e = offs[r0]; e < offs[r1]; e++)
      lo = std::min<uint64_t>(lo, cols[e]), hi = std::max<uint64_t>(hi, cols[e]);
    total += hi - lo + 1;
  }
  return total;
}

// Builds the on-chip plan when the matrix qualifies; leaves M->small null
// (streaming path) when it does not.  Never an error just for not fitting.
// Rows of every CTA chunk (n k / C .. n (k + 1) / C, C = 16) regrouped so that each run of
// `bs` consecutive rows is a set of strongly coupled ones: a block starts at the first row
// not yet taken and grows by the row of the chunk with the largest sum of
// |a_ij| / sqrt(a_ii a_jj) over the block's members (ties: the lower index).  The chunk keeps
// its set of rows, hence its column window.  Returns order[new] = old.
static std::vector<uint32_t> block_growth_order(uint64_t n, int C, uint32_t bs,
                                                const std::vector<uint64_t> &offs,
                                                const std::vector<uint32_t> &cols,
                                                const std::vector<double> &vals) {
  std::vector<double> diag(n, 1.0), score(n, 0.0);
  for (uint64_t r = 0; r < n; r++)
    for (uint64_t e = offs[r]; e < offs[r + 1]; e++)
      if (cols[e] == r && vals[e] > 0.0)
        diag[r] = vals[e];
  std::vector<char> taken(n, 0);
  std::vector<uint32_t> order, cand;
  order.reserve(n);
  for (int k = 0; k < C; k++) {
    const uint64_t r0 = n * k / C, r1 = n * (k + 1) / C;
    uint64_t next = r0, left = r1 - r0;
    while (left) {
      while (taken[next])
        next++;
      uint32_t size = 0;
      uint64_t row = next;
      cand.clear();
      for (;;) {
        taken[row] = 1, order.push_back((uint32_t)row), size++, left--;
        if (size == bs || !left)
          break;
        for (uint64_t e = offs[row]; e < offs[row + 1]; e++) {
          const uint32_t c = cols[e];
          if (c < r0 || c >= r1 || taken[c])
            continue;
          if (score[c] == 0.0)
            cand.push_back(c);
          score[c] += fabs(vals[e]) / sqrt(diag[row] * diag[c]) + 1e-300;
        }
        double best = 0.0;
        uint64_t pick = n;
        for (uint32_t c : cand)
          if (!taken[c] && (score[c] > best || (score[c] == best && c < pick)))
            best = score[c], pick = c;
        if (pick == n) {  // nothing coupled to the block is left in the chunk
          while (taken[next])
            next++;
          pick = next;
        }
        row = pick;
      }
      for (uint32_t c : cand)
        score[c] = 0.0;
    }
  }
  return order;
}

// `blocks`: the plan block-Jacobi runs on (M->small_bj) -- no RCM (compact chunks are worth
// more to the blocks than narrow windows: xn3b_A_10 takes 132 iterations on blocks grown
// inside the file's chunks, 180 inside RCM's), rows regrouped by block_growth_order, the
// inverted diagonal blocks; otherwise the plan Jacobi and Chebyshev-Jacobi run on (M->small).
int small_try_build(b200_mat *M, bool blocks) {
  void *&slot = blocks ? M->small_bj : M->small;
  bool &tried = blocks ? M->small_bj_tried : M->small_tried;
  if (slot || tried)
    return B200_OK;
  tried = true;
  b200_ctx *c = M->ctx;
  const uint64_t n = M->n_local;
  if (c->nranks != 1 || n > SM_MAX_ROWS || M->nnz > SM_MAX_NNZ || M->vec_rows ||
      M->long_rows || n == 0)
    return B200_OK;

  std::vector<uint64_t> offs(n + 1);
  std::vector<uint32_t> cols(M->nnz ? M->nnz : 1);
  std::vector<double> vals(M->nnz ? M->nnz : 1), dinv(n);
  B_TRY(b200_mat_export(M, offs.data(), cols.data(), vals.data()));
  B_TRY(b200_mat_inv_diag(M, dinv.data()));

  // Renumber with RCM when that shrinks the windows by a fifth or more; the
  // kernel then works in the new numbering and only b and x are addressed in
  // the caller's (orig / perm).
  std::vector<uint32_t> perm(n);
  for (uint64_t i = 0; i < n; i++)
    perm[i] = (uint32_t)i;
  bool reordered = false;
  // the operator renumbered by cand[new] = old (columns of a row ascending again); taken
  // when `accept` says so
  auto renumber = [&](const std::vector<uint32_t> &cand, bool always) {
    std::vector<uint32_t> inv(n);
    for (uint64_t i = 0; i < n; i++)
      inv[cand[i]] = (uint32_t)i;
    std::vector<uint64_t> o2(n + 1, 0);
    std::vector<uint32_t> c2(cols.size()), p2(n);
    std::vector<double> v2(vals.size()), d2(n);
    for (uint64_t i = 0; i < n; i++)
      o2[i + 1] = o2[i] + (offs[cand[i] + 1] - offs[cand[i]]);
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t r = cand[i], len = offs[r + 1] - offs[r];
      std::vector<std::pair<uint32_t, double>> row(len);
      for (uint64_t e = 0; e < len; e++)
        row[e] = {inv[cols[offs[r] + e]], vals[offs[r] + e]};
      std::sort(row.begin(), row.end(),
                [](const std::pair<uint32_t, double> &a, const std::pair<uint32_t, double> &b) {
                  return a.first < b.first;
                });
      for (uint64_t e = 0; e < len; e++)
        c2[o2[i] + e] = row[e].first, v2[o2[i] + e] = row[e].second;
      d2[i] = dinv[r], p2[i] = perm[r];
    }
    if (always || 5 * window_cost(n, 16, o2, c2) <= 4 * window_cost(n, 16, offs, cols)) {
      offs.swap(o2), cols.swap(c2), vals.swap(v2), dinv.swap(d2), perm.swap(p2);
      reordered = true;
    }
  };
  if (blocks)
    renumber(block_growth_order(n, 16, 32, offs, cols, vals), true);
  else if (!getenv("B200_SMALL_NO_RCM"))
    renumber(rcm_order(n, offs, cols), false);

  int dev_smem = 0;
  CU_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  for (int kd = 1; kd <= 3; kd++)
    for (int pf = 0; pf < 2; pf++)
      CU_TRY(cudaFuncSetAttribute((const void *)small_kernel(pf != 0, kd),
                                  cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const double lmax = cheb_lmax(n, offs, cols, vals, dinv);

  // B200_SMALL_CLUSTER=8|16 pins the cluster size (experiment switch)
  int only = 0;
  if (const char *e = getenv("B200_SMALL_CLUSTER"))
    only = atoi(e);
  for (int C : {16, 8}) {
    if (only && C != only)
      continue;
    // contiguous row chunks; inside a chunk rows sorted by decreasing length
    std::vector<SmallCta> ctas(C);
    std::vector<uint32_t> goff, rowid;
    std::vector<double> pv, pd;
    std::vector<uint16_t> pc;
    uint32_t max_ent = 0, max_groups = 0, max_stage = 0;
    bool ok = true;
    for (int k = 0; k < C && ok; k++) {
      uint64_t r0 = n * k / C, r1 = n * (k + 1) / C;
      std::vector<uint32_t> rows;
      for (uint64_t r = r0; r < r1; r++)
        rows.push_back((uint32_t)r);
      std::stable_sort(rows.begin(), rows.end(), [&](uint32_t a, uint32_t b) {
        return offs[a + 1] - offs[a] > offs[b + 1] - offs[b];
      });
      uint32_t lo = 0xffffffffu, hi = 0;
      for (uint32_t r : rows) {
        lo = std::min(lo, r), hi = std::max(hi, r);  // own p entries are read too
        for (uint64_t e = offs[r]; e < offs[r + 1]; e++)
          lo = std::min(lo, cols[e]), hi = std::max(hi, cols[e]);
      }
      if (rows.empty())
        lo = 0, hi = 0;
      SmallCta &T = ctas[k];
      T.n_rows = (uint32_t)rows.size();
      T.n_groups = (T.n_rows + 7) / 8;
      T.col_lo = lo, T.col_n = rows.empty() ? 0 : hi - lo + 1;
      T.goff_at = (uint32_t)goff.size(), T.row_at = (uint32_t)rowid.size();
      T.ent_at = pv.size(), T.n_recv = 0, T.row_lo = (uint32_t)r0;
      if (T.col_n > 65536)
        ok = false;
      uint32_t units = 0;
      for (uint32_t g = 0; g < T.n_groups; g++) {
        uint32_t w = 0;
        for (uint32_t i = 0; i < 8; i++) {
          uint32_t s = g * 8 + i;
          if (s < T.n_rows) {
            uint32_t len = (uint32_t)(offs[rows[s] + 1] - offs[rows[s]]);
            w = std::max(w, (len + 3) / 4);
          }
        }
        w = std::max(w, 1u);  // the register walk of small_spmv needs w >= 1
        goff.push_back(units);
        size_t base = pv.size();
        pv.resize(base + (size_t)w * 32, 0.0);
        pc.resize(base + (size_t)w * 32, 0);
        for (uint32_t i = 0; i < 8; i++) {
          uint32_t s = g * 8 + i;
          if (s >= T.n_rows)
            continue;
          uint32_t r = rows[s];
          uint64_t e0 = offs[r], len = offs[r + 1] - e0;
          for (uint64_t e = 0; e < len; e++) {
            // entry e of the row: lane 4i + (e % 4), step e / 4 -- so the four
            // lanes walk the row interleaved and the butterfly adds them
            size_t at = base + (e / 4) * 32 + i * 4 + (e % 4);
            pv[at] = vals[e0 + e], pc[at] = (uint16_t)(cols[e0 + e] - lo);
          }
        }
        units += w;
      }
      goff.push_back(units);
      T.ent = units * 32;
      {  // the first group some warp does not keep in registers (the rule of regmat_load)
        uint32_t first_rest = T.n_groups;
        for (uint32_t warp = 0; warp < SM_WARPS && warp < T.n_groups; warp++) {
          uint32_t total = 0, j = warp;
          for (; j < T.n_groups; j += SM_WARPS) {
            const uint32_t w = goff[T.goff_at + j + 1] - goff[T.goff_at + j];
            if (total + w > SM_REG_STEPS)
              break;
            total += w;
          }
          first_rest = std::min(first_rest, j);
        }
        T.rest_lo = goff[T.goff_at + first_rest] * 32;
      }
      for (uint32_t s = 0; s < T.n_groups * 8; s++) {
        rowid.push_back(s < T.n_rows ? rows[s] : 0xffffffffu);
        pd.push_back(s < T.n_rows ? dinv[rows[s]] : 1.0);
      }
      // (max_ent: the entries a CTA keeps in SHARED memory)
      max_ent = std::max(max_ent, T.ent - T.rest_lo), max_groups = std::max(max_groups, T.n_groups);
      max_stage = std::max(max_stage, T.col_n);
    }
    if (!ok)
      continue;
    // who reads which column: mask of CTAs per row, and per CTA the number of
    // window entries it receives each iteration
    std::vector<uint32_t> readers(n, 0u), dmask(rowid.size(), 0u), orig(rowid.size(), 0xffffffffu);
    for (size_t i = 0; i < rowid.size(); i++)
      if (rowid[i] != 0xffffffffu)
        orig[i] = perm[rowid[i]];
    for (int k = 0; k < C; k++) {
      uint64_t r0 = n * k / C, r1 = n * (k + 1) / C;
      for (uint64_t r = r0; r < r1; r++) {
        readers[r] |= 1u << k;
        for (uint64_t e = offs[r]; e < offs[r + 1]; e++)
          readers[cols[e]] |= 1u << k;
      }
    }
    for (int k = 0; k < C; k++)
      ctas[k].n_recv = 0;
    for (uint64_t r = 0; r < n; r++)
      for (int k = 0; k < C; k++)
        ctas[k].n_recv += (readers[r] >> k) & 1u;
    for (size_t i = 0; i < rowid.size(); i++)
      dmask[i] = rowid[i] == 0xffffffffu ? 0u : readers[rowid[i]];
    SmemMap mp = smem_map(max_ent, max_groups, max_stage, 1);
    if (mp.total > (size_t)dev_smem)
      continue;
    // ---- block-Jacobi: the largest block size whose shared memory fits (B200_SMALL_BJ pins
    // it, 0 = none); the diagonal blocks of bj consecutive chunk rows, inverted (Cholesky),
    // rounded to fp32 symmetrically.  A block that is not positive definite: no block-Jacobi.
    int bj = 0;
    std::vector<uint32_t> binv;
    std::vector<uint16_t> slotof(rowid.size(), 0);
    std::vector<uint32_t> block_of_row(n, 0);
    {
      int want = blocks ? -1 : 0;
      if (const char *e = getenv("B200_SMALL_BJ"))
        want = blocks ? atoi(e) : 0;
      for (int cand : {32, 16}) {
        if (want >= 0 && cand != want)
          continue;
        if (smem_map(max_ent, max_groups, max_stage, 1, cand).total <= (size_t)dev_smem) {
          bj = cand;
          break;
        }
      }
      for (int k = 0; k < C; k++)
        for (uint32_t sl = 0; sl < ctas[k].n_rows; sl++)
          slotof[ctas[k].row_at + (rowid[ctas[k].row_at + sl] - ctas[k].row_lo)] = (uint16_t)sl;
      if (bj) {
        binv.assign(rowid.size() * (size_t)(bj + 1), 0u);
        std::vector<double> B((size_t)bj * bj), L((size_t)bj * bj), Li((size_t)bj * bj), Bi((size_t)bj * bj);
        uint32_t block_id = 0;
        for (int k = 0; k < C && bj; k++) {
          const uint64_t r0 = ctas[k].row_lo, r1 = r0 + ctas[k].n_rows;
          for (uint64_t s0 = r0; s0 < r1 && bj; s0 += bj, block_id++) {
            const int m = (int)std::min<uint64_t>(bj, r1 - s0);
            std::fill(B.begin(), B.end(), 0.0);
            for (int a = 0; a < m; a++)
              for (uint64_t e = offs[s0 + a]; e < offs[s0 + a + 1]; e++)
                if (cols[e] >= s0 && cols[e] < s0 + m)
                  B[(size_t)a * bj + (cols[e] - s0)] += vals[e];
            for (int a = 0; a < m; a++)  // the operator may be unsymmetric at 1e-8 (as stored)
              for (int q = 0; q < a; q++) {
                const double v = 0.5 * (B[(size_t)a * bj + q] + B[(size_t)q * bj + a]);
                B[(size_t)a * bj + q] = B[(size_t)q * bj + a] = v;
              }
            // B = L L^T
            bool spd = true;
            std::fill(L.begin(), L.end(), 0.0);
            for (int a = 0; a < m && spd; a++)
              for (int q = 0; q <= a; q++) {
                double v = B[(size_t)a * bj + q];
                for (int t = 0; t < q; t++)
                  v -= L[(size_t)a * bj + t] * L[(size_t)q * bj + t];
                if (a == q) {
                  if (!(v > 0.0)) {
                    spd = false;
                    break;
                  }
                  L[(size_t)a * bj + a] = sqrt(v);
                } else {
                  L[(size_t)a * bj + q] = v / L[(size_t)q * bj + q];
                }
              }
            if (!spd) {
              bj = 0;
              break;
            }
            // Li = L^-1 (lower), B^-1 = Li^T Li
            std::fill(Li.begin(), Li.end(), 0.0);
            for (int q = 0; q < m; q++) {
              Li[(size_t)q * bj + q] = 1.0 / L[(size_t)q * bj + q];
              for (int a = q + 1; a < m; a++) {
                double v = 0.0;
                for (int t = q; t < a; t++)
                  v -= L[(size_t)a * bj + t] * Li[(size_t)t * bj + q];
                Li[(size_t)a * bj + q] = v / L[(size_t)a * bj + a];
              }
            }
            for (int a = 0; a < m; a++)
              for (int q = 0; q <= a; q++) {
                double v = 0.0;
                for (int t = a; t < m; t++)
                  v += Li[(size_t)t * bj + a] * Li[(size_t)t * bj + q];
                Bi[(size_t)a * bj + q] = Bi[(size_t)q * bj + a] = v;
              }
            for (int a = 0; a < m; a++) {
              uint32_t *dst = binv.data() + ((size_t)ctas[k].row_at + (s0 - r0) + a) * (bj + 1);
              for (int q = 0; q < m; q++) {  // high word of the fp64 value, rounded to nearest
                uint64_t bits;
                const double v = Bi[(size_t)a * bj + q];
                memcpy(&bits, &v, 8);
                dst[q] = (uint32_t)((bits + 0x80000000ull) >> 32);
              }
              block_of_row[perm[s0 + a]] = block_id;
            }
          }
        }
      }
    }
    SmallPlan *P = new SmallPlan();
    P->C = C, P->n = (uint32_t)n, P->smem = mp.total;
    P->max_ent = max_ent, P->max_groups = max_groups, P->max_stage = max_stage;
    P->lmax = lmax;
    bool fits = true;
    for (int kd = 1; kd <= 3 && fits; kd++) {
      const size_t bytes = smem_map(max_ent, max_groups, max_stage, kd).total;
      if (bytes > (size_t)dev_smem)
        break;
      for (int pf = 0; pf < 2; pf++)
        if (cudaFuncSetAttribute((const void *)small_kernel(pf != 0, kd),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
          cudaGetLastError();
          fits = false;
        }
      if (fits)
        P->max_kdeg = kd, P->smem_k[kd] = bytes;
    }
    if (!P->smem_k[1]) {
      delete P;
      continue;
    }
    if (bj) {
      const size_t bytes = smem_map(max_ent, max_groups, max_stage, 1, bj).total;
      for (int pf = 0; pf < 2; pf++)
        if (cudaFuncSetAttribute((const void *)small_kernel(pf != 0, 1, bj),
                                 cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
            cudaFuncSetAttribute((const void *)small_kernel(pf != 0, 1, bj),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
          cudaGetLastError();
          bj = 0;
        }
      if (bj)
        P->bj = bj, P->smem_bj = bytes, P->block_of_row = block_of_row;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    small_launch_config(P, &cfg, attr, c->stream);
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (const void *)small_kernel(false, 1), &cfg) != cudaSuccess ||
        nclusters < 1) {
      cudaGetLastError();
      delete P;
      continue;
    }
    auto up = [&](void **d, const void *h, size_t bytes) -> int {
      CU_TRY(cudaMalloc(d, bytes ? bytes : 8));
      CU_TRY(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, c->stream));
      CU_TRY(cudaStreamSynchronize(c->stream));
      M->device_bytes += bytes;
      return B200_OK;
    };
    B_TRY(up((void **)&P->d_cta, ctas.data(), ctas.size() * sizeof(SmallCta)));
    B_TRY(up((void **)&P->d_goff, goff.data(), goff.size() * 4));
    B_TRY(up((void **)&P->d_rowid, rowid.data(), rowid.size() * 4));
    B_TRY(up((void **)&P->d_dmask, dmask.data(), dmask.size() * 4));
    B_TRY(up((void **)&P->d_orig, orig.data(), orig.size() * 4));
    B_TRY(up((void **)&P->d_perm, perm.data(), perm.size() * 4));
    P->reordered = reordered;
    B_TRY(up((void **)&P->d_vals, pv.data(), pv.size() * 8));
    B_TRY(up((void **)&P->d_cols, pc.data(), pc.size() * 2));
    B_TRY(up((void **)&P->d_dinv, pd.data(), pd.size() * 8));
    if (P->bj)
      B_TRY(up((void **)&P->d_binv, binv.data(), binv.size() * 4));
    B_TRY(up((void **)&P->d_slotof, slotof.data(), slotof.size() * 2));
    CU_TRY(cudaMalloc(&P->d_state, sizeof(PcgState)));
    CU_TRY(cudaMalloc(&P->d_prof, 16 * sizeof(long long)));
    slot = P;
    return B200_OK;
  }
  return B200_OK;
}

extern "C" int b200_mat_block_jacobi_partition(b200_mat *M, uint32_t *block_of_row, uint32_t *block_size) {
  if (!M || !block_of_row || !block_size)
    B_FAIL(B200_EINVAL, "b200_mat_block_jacobi_partition: null argument");
  CU_TRY(cudaSetDevice(M->ctx->device));
  B_TRY(small_try_build(M, true));
  const SmallPlan *P = (const SmallPlan *)M->small_bj;
  *block_size = P ? (uint32_t)P->bj : 0u;
  if (P && P->bj)
    memcpy(block_of_row, P->block_of_row.data(), P->block_of_row.size() * sizeof(uint32_t));
  return B200_OK;
}

int small_solve(b200_mat *M, const double *d_b, double *d_x,
                const b200_pcg_opts *o, b200_pcg_result *res) {
  // block-Jacobi runs on a plan of its own (rows regrouped into strongly coupled blocks)
  SmallPlan *P = (SmallPlan *)M->small;
  const bool cheb = (o->flags & (B200_PCG_CHEBYSHEV2 | B200_PCG_CHEBYSHEV3)) != 0;
  if ((o->flags & B200_PCG_BLOCK_JACOBI) && !cheb) {
    B_TRY(small_try_build(M, true));
    if (M->small_bj && ((SmallPlan *)M->small_bj)->bj)
      P = (SmallPlan *)M->small_bj;
  }
  b200_ctx *c = M->ctx;
  cudaStream_t s = c->stream;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  small_launch_config(P, &cfg, attr, s);
  // preconditioner degree: B200_PCG_CHEBYSHEV2 / 3 (or B200_SMALL_CHEB=2|3), at most what fits
  static const int env_deg = [] {
    const char *v = getenv("B200_SMALL_CHEB");
    return v ? atoi(v) : 0;
  }();
  int kdeg = (o->flags & B200_PCG_CHEBYSHEV3) ? 3 : (o->flags & B200_PCG_CHEBYSHEV2) ? 2 : env_deg;
  kdeg = kdeg < 1 ? 1 : kdeg > P->max_kdeg ? P->max_kdeg : kdeg;
  // block-Jacobi: B200_PCG_BLOCK_JACOBI, when a block size fits (Chebyshev-Jacobi wins a tie)
  const int bj = ((o->flags & B200_PCG_BLOCK_JACOBI) && kdeg == 1) ? P->bj : 0;
  cfg.dynamicSmemBytes = bj ? P->smem_bj : P->smem_k[kdeg];
  const double la = P->lmax / 30.0, theta = 0.5 * (P->lmax + la), delta = 0.5 * (P->lmax - la);
  CU_TRY(cudaEventRecord(c->ev_a, s));
  static const bool prof = getenv("B200_SMALL_PROFILE") != nullptr;
  CU_TRY(cudaLaunchKernelEx(&cfg, small_kernel(prof, kdeg, bj),
                            (const SmallCta *)P->d_cta,
                            (const uint32_t *)P->d_goff, (const uint32_t *)P->d_rowid,
                            (const uint32_t *)P->d_dmask, (const uint32_t *)P->d_orig,
                            (const uint32_t *)P->d_perm, (const double *)P->d_vals, (const uint16_t *)P->d_cols,
                            (const double *)P->d_dinv, (const uint32_t *)P->d_binv,
                            (const uint16_t *)P->d_slotof, d_b, d_x,
                            P->d_state, P->max_ent, P->max_groups, P->max_stage,
                            o->tol, (int)o->maxit, P->d_prof, theta, delta));
  c->launches += 1;
  CU_TRY(cudaEventRecord(c->ev_b, s));
  PcgState h;
  CU_TRY(cudaMemcpyAsync(&h, P->d_state, sizeof h, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaEventElapsedTime(&res->solve_ms, c->ev_a, c->ev_b));
  res->iters = h.iter, res->status = h.status;
  res->bnorm = sqrt(h.bb);
  res->relres = h.bb > 0 ? sqrt(h.red[1] / h.bb) : sqrt(h.red[1]);
  res->true_relres = h.bb > 0 ? sqrt(h.true_rr / h.bb) : sqrt(h.true_rr);
  res->kernel_launches = 1;
  res->path = 1;
  res->replacements = h.replacements;
  res->outer_iters = kdeg;  // (on this path: the degree of the preconditioner that ran)
  res->block_jacobi = (uint32_t)bj;
  if (prof && h.iter > 0) {
    long long hp[6];
    CU_TRY(cudaMemcpy(hp, P->d_prof, sizeof hp, cudaMemcpyDeviceToHost));
    fprintf(stderr, "b200 small: n=%u C=%d reordered=%d bj=%d smem=%zu (entries in smem <= %u) iters=%d cycles/iter "
                    "spmv=%lld reduce1=%lld update+push=%lld reduce2+halo=%lld pwindow=%lld\n", P->n, P->C,
            (int)P->reordered, bj, (size_t)cfg.dynamicSmemBytes, P->max_ent, h.iter, hp[0] / h.iter,
            hp[1] / h.iter, hp[2] / h.iter, hp[3] / h.iter, hp[4] / h.iter);
  }
  if (h.status == 2)
    B_FAIL(B200_ENOTSPD, "b200_pcg_solve: breakdown at iteration %d (p.Ap = %g)",
           h.iter, h.pq);
  return B200_OK;
}
