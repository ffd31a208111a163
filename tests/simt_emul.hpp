// simt_emul.hpp -- TEST INFRASTRUCTURE.  A small SIMT emulator: runs a CUDA
// kernel body compiled for the host with one host thread per CUDA thread of a
// block, blocks one after another.  __syncthreads is a barrier over the block,
// warp shuffles exchange through a per-warp buffer between two warp barriers,
// __shared__ variables are function-static (one block lives at a time), atomics
// are the host's.  That is enough for the kernels of pcg_kernels.cuh and
// sell_kernels.cuh including their fixed-order reductions (block_sum,
// grid_sum_finish: the last block to take its ticket adds the partials), so the
// product's kernels can be held against the oracle without a GPU.
//
// Include AFTER <cuda_runtime.h> (common.cuh pulls it in) and BEFORE the kernel
// headers; define B2_SIMT_EMUL on the command line.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace simt {
struct Idx {
  unsigned x, y, z;
};
struct BlockCtx {
  std::barrier<> cta;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<std::vector<unsigned long long>> xch;  // per warp: 32 lanes of 8 bytes
  explicit BlockCtx(unsigned threads) : cta(threads) {
    for (unsigned w = 0; w < (threads + 31) / 32; w++) {
      unsigned lanes = threads - w * 32 < 32 ? threads - w * 32 : 32;
      warp.emplace_back(new std::barrier<>(lanes));
      xch.emplace_back(32, 0ull);
    }
  }
};
inline thread_local Idx t_idx{0, 0, 0}, b_idx{0, 0, 0};
inline Idx g_dim{1, 1, 1}, b_dim{1, 1, 1};
inline thread_local BlockCtx *ctx = nullptr;

template <typename T> inline T shfl_xor(T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned w = t_idx.x >> 5, lane = t_idx.x & 31;
  unsigned long long raw = 0;
  __builtin_memcpy(&raw, &v, sizeof(T));
  ctx->xch[w][lane] = raw;
  ctx->warp[w]->arrive_and_wait();
  raw = ctx->xch[w][lane ^ (unsigned)o];
  ctx->warp[w]->arrive_and_wait();
  T r;
  __builtin_memcpy(&r, &raw, sizeof(T));
  return r;
}

// launch(grid, threads, [&] { kernel(args...); })
inline void launch(unsigned grid, unsigned threads, const std::function<void()> &body) {
  g_dim = Idx{grid, 1, 1}, b_dim = Idx{threads, 1, 1};
  for (unsigned b = 0; b < grid; b++) {
    BlockCtx bc(threads);
    std::vector<std::thread> th;
    th.reserve(threads);
    for (unsigned t = 0; t < threads; t++)
      th.emplace_back([&, t, b] {
        t_idx = Idx{t, 0, 0}, b_idx = Idx{b, 0, 0}, ctx = &bc;
        body();
        // a thread that leaves early (e.g. `if (st->done) return`) must not
        // strand the others at a barrier: CUDA kernels under test only return
        // early block-uniformly before any barrier, so dropping is safe
        bc.cta.arrive_and_drop();
        bc.warp[t >> 5]->arrive_and_drop();
      });
    for (auto &x : th)
      x.join();
  }
}
}  // namespace simt

// ---- the CUDA spellings the kernels use ---------------------------------------------
#undef __shared__
#define __shared__ static
#define threadIdx simt::t_idx
#define blockIdx simt::b_idx
#define gridDim simt::g_dim
#define blockDim simt::b_dim
#define __launch_bounds__(...)
static inline void __syncthreads() { simt::ctx->cta.arrive_and_wait(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __shfl_xor_sync(unsigned, double v, int o) { return simt::shfl_xor(v, o); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline long long clock64() { return 0; }
static inline double __longlong_as_double(long long v) {
  double d;
  __builtin_memcpy(&d, &v, 8);
  return d;
}
