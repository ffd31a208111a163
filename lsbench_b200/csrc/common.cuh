// common.cuh -- internal types and device helpers of libb200.so (sm_100a).
#pragma once
#include "b200.h"
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define B200_SM_COUNT_FALLBACK 148
#define B2_SLICE 32  // SELL slice height = one warp, one row per lane

void b200_set_error(const char *fmt, ...);

#define CU_TRY(expr)                                                           \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      b200_set_error("%s:%d cuda error: %s (%s)", __FILE__, __LINE__,          \
                     cudaGetErrorString(e_), #expr);                           \
      return e_ == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;       \
    }                                                                          \
  } while (0)

#define B_TRY(expr)                                                            \
  do {                                                                         \
    int r_ = (expr);                                                           \
    if (r_ != B200_OK)                                                         \
      return r_;                                                               \
  } while (0)

#define B_FAIL(code, ...)                                                      \
  do {                                                                         \
    b200_set_error(__VA_ARGS__);                                               \
    return (code);                                                             \
  } while (0)

// ---- NCCL, bound at run time (dist.cu) ------------------------------------
struct NcclApi;

// ---- peer-memory all-reduce of the CG scalars (dist.cu xr_setup) ---------------
// Every rank owns a mailbox [kind][source rank][4 doubles: v0, v1, v2, seq] that
// all ranks can write over NVLink (same process: peer access; other process:
// CUDA IPC).  The CTA that finishes a rank-local sum stores it, then a sequence
// number, into EVERY rank's mailbox (common.cuh xr_push, inside
// grid_sum_finish); the kernel that needs the global sum waits for that
// sequence number from all ranks in its own mailbox and adds the values in
// rank order (xr_wait_sum) -- the same additions on every rank.  No collective
// kernel, no extra launch: the reduction is fused into producer and consumer.
//
// Sequence numbers live on the device so that a chunk of iterations can be a CUDA
// graph on several ranks too: `base` points at a counter that a one-thread kernel
// at the head of every chunk advances by the chunk's length (k_xr_chunk_begin),
// `seq` is the position inside the chunk -- baked into the captured launch -- and
// the number that travels is *base + seq.  Every rank queues the same chunks, so
// the counters agree without ever being exchanged.
//
// Kind 2 is the flag of the halo exchange over peer memory (dist.cu halo_peer_setup,
// k_halo_push / k_halo_wait): the rows a neighbour reads are stored straight into
// its vector's halo section, then the sequence number into its mailbox.
#define B2_XR_MAX_RANKS 16
#define B2_XR_KINDS 3  // 0: p.q    1: r.z, r.r    2: halo of p has landed
struct XrArgs {
  double *const *peers;  // device array, peers[r] = rank r's mailbox; nullptr = off
  double *mine;          // this rank's mailbox
  int nranks, me, kind;
  unsigned long long seq;              // position inside the chunk (or the number itself)
  const unsigned long long *base;      // device counter added to seq; nullptr = 0
};
#if defined(__CUDACC__) || defined(B2_SIMT_EMUL)
__device__ __forceinline__ unsigned long long xr_seq_of(const XrArgs &xr) {
  return xr.seq + (xr.base ? *xr.base : 0ull);
}
#endif

struct b200_ctx {
  int device = 0;
  int sm_count = B200_SM_COUNT_FALLBACK;
  int rank = 0, nranks = 1;
  cudaStream_t own_stream = nullptr;   // created by the context
  cudaStream_t stream = nullptr;       // where kernels go (own or caller's)
  cudaStream_t comm_stream = nullptr;  // halo exchange, overlapped with interior
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_halo = nullptr,
              ev_ready = nullptr, ev_poll = nullptr;
  void *nccl_comm = nullptr;
  const NcclApi *nccl = nullptr;
  // peer-memory all-reduce (nullptr / false: NCCL all-reduce)
  double *xr_mail = nullptr;      // my mailbox
  double **xr_peers = nullptr;    // device array of the ranks' mailboxes
  void *xr_opened[B2_XR_MAX_RANKS] = {nullptr};  // IPC mappings to close
  double *xr_peers_h[B2_XR_MAX_RANKS] = {nullptr};  // the same pointers on the host
  bool xr_on = false;
  unsigned long long *d_seq = nullptr;  // device: {base of the running chunk, next base}
  int *h_flag = nullptr;  // pinned: PCG progress word read by the host
  uint64_t launches = 0;  // kernels of this library queued so far
};

// Plain CSR on the device: the intermediate every input goes through.
struct PlainCsr {
  uint64_t n = 0;      // rows held
  uint64_t nnz = 0;
  uint64_t *offs = nullptr;  // n+1
  uint32_t *cols = nullptr;
  double *vals = nullptr;
};

// the halo exchange over peer memory: where the entries I own go, and whose flags I wait for
struct HaloPushArgs {
  int n_dst;                                   // neighbours that read rows of mine
  uint32_t seg_begin[B2_XR_MAX_RANKS + 1];     // their runs in the send list
  double *dst[B2_XR_MAX_RANKS];                // first halo slot of mine in their vector
  unsigned long long *flag[B2_XR_MAX_RANKS];   // my flag word in their mailbox
};
struct HaloWaitArgs {
  int n_src;                                   // neighbours whose rows I read
  int src[B2_XR_MAX_RANKS];
};

struct HaloPlan {
  // receive side: halo slots are sorted by global column, grouped by owner
  uint64_t n_halo = 0;
  uint64_t n_low = 0;                    // halo slots whose global column lies below the own rows
  uint64_t *d_gcols = nullptr;           // n_halo global ids
  int n_peers = 0;                       // ranks we exchange with
  int *peer = nullptr;                   // host arrays, n_peers long
  uint64_t *recv_off = nullptr;          // n_peers+1, into the halo section
  uint64_t *send_off = nullptr;          // n_peers+1, into send buffer
  uint32_t *d_send_idx = nullptr;        // local row ids to pack
  double *d_send_buf = nullptr;
  uint64_t n_send = 0;
  // peer-memory path (dist.cu halo_peer_setup): set up once the solver workspace exists
  bool peer_tried = false, peer_ready = false;
  HaloPushArgs push;
  HaloWaitArgs wait;
  unsigned *d_push_ticket = nullptr;
  void *peer_opened[B2_XR_MAX_RANKS] = {nullptr};  // IPC mappings of the neighbours' p vectors
};

struct SpmvPlan {
  uint32_t b0, e0, b1, e1;  // SELL slice ranges of the phase
  int g_sell, g_vec, g_long;
};

// Device-resident scalars of one PCG solve.  red[] holds the values that are
// all-reduced across ranks: {rz, rr} for iteration parity 0 at [0..1], parity
// 1 at [2..3], and the start-up triple {rz, rr, bb} at [4..6].
// loc[] / pq_loc / true_rr_loc are this rank's partial sums; on a multi-rank
// context they are all-reduced out of place into red[] / pq / true_rr, so a
// repeated all-reduce (iterations queued past convergence are no-ops that
// leave loc[] untouched) reproduces the same values instead of compounding.
struct PcgState {
  double pq, pq_loc;
  double red[7];
  double loc[7];
  double true_rr_loc;
  double bb, thr2, tol;
  double true_rr;
  int iter, done, status, maxit;
  unsigned ticket[4];
  int replacements, pad0;  // residual replacements of the on-chip solve (small.cu)
};

struct b200_mat {
  b200_ctx *ctx = nullptr;
  uint64_t n_global = 0, row_begin = 0, n_local = 0, nnz = 0;
  uint32_t flags = 0;
  // --- SELL-32 bin: slice s holds rows perm[32s .. 32s+31], column-major ----
  uint32_t sell_rows = 0, sell_slices = 0, sell_sigma = 1, sell_max_width = 0;
  uint32_t *sell_off = nullptr;   // sell_slices+1, in units of 32 entries
  uint32_t *sell_cols = nullptr;
  double *sell_vals = nullptr;
  // B200_MAT_VALUES_F32: the SELL values as fp32 in the same layout.  When
  // every value survived the rounding (vals32_exact) the fp64 stream is freed
  // and all products read this one -- same fp64 arithmetic, same bits, half the
  // value bytes; otherwise both stay and the PCG runs as iterative refinement
  // (inner iterations on the fp32-valued operator, residuals on the fp64 one).
  float *sell_vals32 = nullptr;
  bool vals32_exact = false;
  bool spmv_use32 = false;        // which value stream launch_spmv multiplies with
  uint32_t *sell_perm = nullptr;  // nullptr == identity (row = 32 s + lane)
  // index compression (nullptr == off): sell_meta[s] = {o, w | uniform << 31,
  // column offset, 0} (uint4 per slice), the deltas of the uniform slices;
  // sell_cols then holds only the explicit slices
  uint32_t *sell_meta = nullptr;
  int32_t *sell_dcols = nullptr;
  uint64_t sell_uniform_slices = 0, sell_col_entries = 0, sell_delta_entries = 0;
  uint64_t sell_entries = 0;      // padded
  // --- warp-per-row and block-per-row bins: row-major, rows padded to 4 ----
  uint32_t vec_rows = 0, long_rows = 0;
  uint32_t *vec_row_ids = nullptr, *long_row_ids = nullptr;
  uint64_t *vec_off = nullptr, *long_off = nullptr;  // rows+1 each
  uint32_t *vl_cols = nullptr;
  double *vl_vals = nullptr;
  uint64_t vl_entries = 0, vec_nnz = 0, long_nnz = 0;
  // --- common -----------------------------------------------------------------
  double *dinv = nullptr;        // n_local
  uint32_t *row_len = nullptr;   // n_local, true lengths (export / checks)
  uint64_t hist[B200_HIST_BINS] = {0};
  uint64_t max_row_len = 0;
  uint32_t pattern_symmetric = 1;
  uint64_t interior_begin = 0, interior_end = 0;
  HaloPlan halo;
  uint64_t device_bytes = 0;
  // --- solver workspace (lazily allocated) ----------------------------------
  double *w_r = nullptr, *w_p = nullptr, *w_q = nullptr;  // p has halo room
  double *w_x = nullptr;          // iterate (graph-stable pointer)
  double *w_d = nullptr, *w_rhs = nullptr;  // refinement: correction and residual
  double *stage_b = nullptr, *stage_x = nullptr;  // b200_pcg_solve_host staging
  int grid_ew = 0;                // element-wise kernels
  unsigned partial_stride = 0;
  double *x_ext = nullptr;        // spmv staging: n_local + n_halo
  double *partials = nullptr;     // per-CTA partial sums, 4 lanes of them
  PcgState *state = nullptr;
  SpmvPlan plan[3];              // phase 0 all, 1 interior, 2 boundary
  bool plan_ready = false;
  // column-blocked operator (B200_MAT_COL_BLOCK): the column ranges as matrices of
  // their own (all n rows, global column ids); this one then holds no entries
  std::vector<b200_mat *> blocks;            // pass order: ranges of owned columns first
  std::vector<uint32_t> block_in_col_order;  // blocks[] indices in global column order
  uint32_t n_local_blocks = 0;               // how many of blocks[] hold owned columns only
  bool grouped_slices = false;   // a column range: k_spmv_sell_grp takes its slices four at a time
  int plan_grp = 1;
  unsigned *grp_work = nullptr;  // k_spmv_sell_grp: {next unit of work, CTAs done}
  uint64_t sell_sigma_cap = 0;   // a column range: widest sort window (0 = the whole list)
  uint32_t pad_col = 0xffffffffu; // a column range: padding gathers this column (inside the range), not x[own row]
  uint64_t col_block_width = 0;
  void *small = nullptr;          // on-chip small-matrix plan (small.cu)
  bool small_tried = false;
  void *small_bj = nullptr;       // the plan block-Jacobi runs on (rows regrouped into blocks)
  bool small_bj_tried = false;
  void *graph_exec = nullptr;     // cudaGraphExec_t of one iteration chunk
  int graph_chunk = 0;
  int graph_kernels = 0;         // kernel nodes in the captured chunk
  void *graph_stream = nullptr;
};

// ---- helpers implemented across the .cu files --------------------------------
int dev_alloc(b200_mat *M, void **p, size_t bytes);  // tracks device_bytes
int plain_free(PlainCsr *A);
int build_layout(b200_ctx *ctx, PlainCsr *A, uint64_t n_global,
                 uint64_t row_begin, uint32_t flags, b200_mat **out);
int build_layout_or_blocks(b200_ctx *ctx, PlainCsr *A, uint64_t n_global,
                           uint64_t row_begin, uint32_t flags, b200_mat **out);
int partition_and_renumber(b200_ctx *ctx, PlainCsr *A, uint64_t n_global,
                           uint64_t row_begin, b200_mat *M);
int halo_setup(b200_mat *M);
int halo_exchange_begin(b200_mat *M, double *d_x_ext);  // on comm stream
int halo_exchange_wait(b200_mat *M);
int halo_peer_setup(b200_mat *M);                 // collective; after w_p exists
int halo_peer_push(b200_mat *M, const double *x_ext, unsigned seq_off);   // on the comm stream
int halo_peer_wait(b200_mat *M, unsigned seq_off);                       // on the compute stream
int xr_chunk_begin(b200_ctx *c, unsigned count);  // advance the device sequence counter
void halo_free(b200_mat *M);
int ensure_workspace(b200_mat *M);
int launch_spmv(b200_mat *M, const double *x_ext, double *y, bool fuse_dot,
                int phase /*0 all, 1 interior, 2 boundary*/, const XrArgs *xr);
int allreduce_sum(b200_ctx *ctx, const double *d_src, double *d_dst, int count);
int small_try_build(b200_mat *M, bool blocks = false);
void small_free(b200_mat *M);
int small_solve(b200_mat *M, const double *d_b, double *d_x,
                const b200_pcg_opts *o, b200_pcg_result *res);

// ---- device-side helpers -----------------------------------------------------
// (B2_SIMT_EMUL: tests/simt_emul.hpp compiles them for the host, one host thread
// per CUDA thread, to run the kernels of pcg_kernels.cuh / sell_kernels.cuh
// without a GPU)
#if defined(__CUDACC__) || defined(B2_SIMT_EMUL)
__device__ __forceinline__ double warp_sum(double v) {
  // fixed butterfly: the same tree for every run
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of one value per thread, fixed order: warp trees, then the
// warp results in ascending warp order by warp 0.  Result valid in thread 0.
template <int NWARPS>
__device__ __forceinline__ double block_sum(double v, double *smem /*NWARPS*/) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0)
    smem[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = lane < NWARPS ? smem[lane] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// ---- peer-memory all-reduce, device side --------------------------------------
#ifndef B2_SIMT_EMUL
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

#else  // host emulation: one process, ordinary memory
static inline void st_relaxed_sys(double *p, double v) { __atomic_store(p, &v, __ATOMIC_RELAXED); }
static inline void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
static inline unsigned long long ld_acquire_sys(const unsigned long long *p) {
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
static inline double ld_relaxed_sys(const double *p) {
  double v;
  __atomic_load(p, &v, __ATOMIC_RELAXED);
  return v;
}
#endif

// One CTA, after its __syncthreads: thread r < nranks stores the NV values and
// then the sequence number into rank r's mailbox slot [kind][me].
template <int NV>
__device__ __forceinline__ void xr_push(const XrArgs &xr, const double *vals /*smem, NV*/) {
  const int r = threadIdx.x;
  if (r < xr.nranks) {
    double *slot = xr.peers[r] + ((size_t)xr.kind * B2_XR_MAX_RANKS + xr.me) * 4;
#pragma unroll
    for (int v = 0; v < NV; v++)
      st_relaxed_sys(slot + v, vals[v]);
    __threadfence_system();
    st_relaxed_sys(reinterpret_cast<unsigned long long *>(slot) + 3, xr_seq_of(xr));
  }
}

// Every thread of the CTA calls it.  Waits until all ranks' values of sequence
// xr.seq are in my mailbox, then tot[v] = sum over the ranks in rank order
// (identical on every rank).  smem: NV * B2_XR_MAX_RANKS doubles + 1 int.
// Returns false when a peer did not deliver within ~4 s (never hang the GPU).
template <int NV>
__device__ __forceinline__ bool xr_wait_sum(const XrArgs &xr, double (&tot)[NV], double *smem) {
  int *ok = reinterpret_cast<int *>(smem + NV * B2_XR_MAX_RANKS);
  const int r = threadIdx.x;
  if (r == 0)
    *ok = 1;
  __syncthreads();
  if (r < xr.nranks) {
    const double *slot = xr.mine + ((size_t)xr.kind * B2_XR_MAX_RANKS + r) * 4;
    const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(slot) + 3;
    const long long t0 = clock64();
    const unsigned long long want = xr_seq_of(xr);
    bool got = true;
    while (ld_acquire_sys(flag) != want)
      if (clock64() - t0 > 8000000000ll) {
        got = false;
        break;
      }
    if (!got)
      *ok = 0;
#pragma unroll
    for (int v = 0; v < NV; v++)
      smem[v * B2_XR_MAX_RANKS + r] = ld_relaxed_sys(slot + v);
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s = 0.0;
    for (int k = 0; k < xr.nranks; k++)
      s += smem[v * B2_XR_MAX_RANKS + k];
    tot[v] = s;
  }
  const bool good = *ok != 0;
  __syncthreads();
  return good;
}

// Grid-wide deterministic sum of NV values per CTA.  Each CTA stores its
// block sums to partials[v * stride + slot]; the CTA drawing the last ticket
// adds the `total` partials of every lane in a fixed order (strided per
// thread, then the block tree) and writes out[v].  The order depends only on
// `total`, i.e. on the launch geometry -- never on scheduling.  With xr.peers
// set, that CTA also ships the sums to every rank's mailbox (xr_push).
template <int NV, int NWARPS>
__device__ __forceinline__ void grid_sum_finish(const double (&block_val)[NV],
                                                double *partials,
                                                unsigned stride, unsigned slot,
                                                unsigned total, unsigned *ticket,
                                                double *out, double *smem,
                                                const XrArgs xr = XrArgs{nullptr, nullptr, 1, 0, 0, 0ull}) {
  static_assert(NV <= NWARPS, "the sums are staged in the warp-sum scratch");
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; v++)
      partials[v * stride + slot] = block_val[v];
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == total - 1);
  }
  __syncthreads();
  if (!is_last)
    return;
  __threadfence();
  double res[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < total; i += blockDim.x)
      acc += __ldcg(partials + v * stride + i);
    acc = block_sum<NWARPS>(acc, smem);
    res[v] = acc;
    if (threadIdx.x == 0)
      out[v] = acc;
  }
  if (threadIdx.x == 0)
    *ticket = 0;
  if (xr.peers) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int v = 0; v < NV; v++)
        smem[v] = res[v];
    }
    __syncthreads();
    xr_push<NV>(xr, smem);
  }
}

// streaming (read-once) loads: keep them out of L1, evict-first in L2
__device__ __forceinline__ double ld_stream(const double *p) {
  return __ldcs(p);
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
  return __ldcs(p);
}
__device__ __forceinline__ float ld_stream(const float *p) {
  return __ldcs(p);
}
#endif
