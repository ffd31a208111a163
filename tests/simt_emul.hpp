// simt_emul.hpp -- TEST INFRASTRUCTURE.  A small SIMT emulator: runs a CUDA
// kernel body compiled for the host with one FIBER per CUDA thread of
// a block (a 30-byte x86-64 stack switch: swapcontext makes a system call per
// switch), blocks one after another, everything on one OS thread.  A fiber runs
// until it reaches a barrier and then yields to the next one, round robin:
// __syncthreads is a barrier over the live fibers of the block, a warp shuffle
// exchanges through a per-warp buffer between two warp barriers, __shared__
// variables are function-static (one block lives at a time), atomics are plain.
// Deterministic, no data races, and enough for the kernels of pcg_kernels.cuh
// and sell_kernels.cuh including their fixed-order reductions (block_sum,
// grid_sum_finish: the last block to take its ticket adds the partials), so the
// product's kernels can be held against the oracle without a GPU.
//
// Include AFTER <cuda_runtime.h> and BEFORE common.cuh / the kernel headers;
// define B2_SIMT_EMUL on the command line.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <vector>

#if !defined(__x86_64__)
#error "simt_emul.hpp switches stacks with x86-64 assembly"
#endif
// simt_switch(&save_sp, load_sp): push the callee-saved registers, park the stack
// pointer, adopt the other one, pop, return into the other fiber
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch, @function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch, .-simt_switch
)");

namespace simt {
struct Idx {
  unsigned x, y, z;
};
struct Barrier {
  unsigned expected = 0, count = 0, gen = 0;
};
struct Fiber {
  void *sp = nullptr;
  Idx t_idx{0, 0, 0};
  bool done = false;
  const Barrier *wait_bar = nullptr;  // parked until wait_bar->gen != wait_gen
  unsigned wait_gen = 0;
};
struct Block {
  std::vector<Fiber> fibers;
  Barrier cta;
  std::vector<Barrier> warp;
  std::vector<unsigned long long> xch;  // 32 lanes of 8 bytes per warp
  void *sched_sp = nullptr;
  const std::function<void()> *body = nullptr;
};
inline Block *blk = nullptr;
inline Fiber *cur = nullptr;
inline Idx b_idx{0, 0, 0}, g_dim{1, 1, 1}, b_dim{1, 1, 1};
constexpr size_t STACK = 256 * 1024;

inline void yield() { simt_switch(&cur->sp, blk->sched_sp); }

inline void arrive_and_wait(Barrier &b) {
  const unsigned gen = b.gen;
  if (++b.count >= b.expected) {
    b.count = 0, b.gen++;
  } else {
    cur->wait_bar = &b, cur->wait_gen = gen;  // the scheduler resumes us when the phase ends
    yield();
    cur->wait_bar = nullptr;
  }
}
// a fiber that has finished no longer takes part (arrive_and_drop)
inline void drop(Barrier &b) {
  b.expected--;
  if (b.expected && b.count >= b.expected)
    b.count = 0, b.gen++;
}

inline void trampoline() {
  (*blk->body)();
  cur->done = true;
  drop(blk->cta);
  drop(blk->warp[cur->t_idx.x >> 5]);
  simt_switch(&cur->sp, blk->sched_sp);
  std::abort();  // a finished fiber is never resumed
}

template <typename T> inline T shfl_xor(T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned w = cur->t_idx.x >> 5, lane = cur->t_idx.x & 31;
  unsigned long long raw = 0;
  __builtin_memcpy(&raw, &v, sizeof(T));
  blk->xch[w * 32 + lane] = raw;
  arrive_and_wait(blk->warp[w]);
  raw = blk->xch[w * 32 + (lane ^ (unsigned)o)];
  arrive_and_wait(blk->warp[w]);
  T r;
  __builtin_memcpy(&r, &raw, sizeof(T));
  return r;
}

// value of lane `src` (every lane of the warp calls it)
template <typename T> inline T shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned w = cur->t_idx.x >> 5, lane = cur->t_idx.x & 31;
  unsigned long long raw = 0;
  __builtin_memcpy(&raw, &v, sizeof(T));
  blk->xch[w * 32 + lane] = raw;
  arrive_and_wait(blk->warp[w]);
  raw = blk->xch[w * 32 + ((unsigned)src & 31u)];
  arrive_and_wait(blk->warp[w]);
  T r;
  __builtin_memcpy(&r, &raw, sizeof(T));
  return r;
}

// launch(grid, threads, [&] { kernel(args...); })
inline void launch(unsigned grid, unsigned threads, const std::function<void()> &body) {
  g_dim = Idx{grid, 1, 1}, b_dim = Idx{threads, 1, 1};
  static std::vector<void *> stacks;  // reused across launches
  while (stacks.size() < threads)
    stacks.push_back(std::malloc(STACK));
  for (unsigned b = 0; b < grid; b++) {
    Block B;
    B.body = &body;
    B.fibers.resize(threads);
    B.cta.expected = threads;
    const unsigned nw = (threads + 31) / 32;
    B.warp.resize(nw);
    B.xch.assign((size_t)nw * 32, 0ull);
    for (unsigned w = 0; w < nw; w++)
      B.warp[w].expected = threads - w * 32 < 32 ? threads - w * 32 : 32;
    blk = &B, b_idx = Idx{b, 0, 0};
    for (unsigned t = 0; t < threads; t++) {
      Fiber &f = B.fibers[t];
      f.t_idx = Idx{t, 0, 0};
      // a fresh stack that simt_switch can "return" into: six zero registers, then
      // the entry point, then a null return address; rsp is 8 mod 16 at entry
      void **top = (void **)(((uintptr_t)stacks[t] + STACK) & ~(uintptr_t)15);
      *--top = nullptr;
      *--top = (void *)trampoline;
      for (int k = 0; k < 6; k++)
        *--top = nullptr;
      f.sp = top;
    }
    unsigned live = threads;
    while (live) {
      for (unsigned t = 0; t < threads; t++) {
        Fiber &f = B.fibers[t];
        if (f.done || (f.wait_bar && f.wait_bar->gen == f.wait_gen))
          continue;
        cur = &f;
        simt_switch(&B.sched_sp, f.sp);
        if (f.done)
          live--;
      }
    }
    blk = nullptr, cur = nullptr;
  }
}
}  // namespace simt

// ---- the CUDA spellings the kernels use ---------------------------------------------
#undef __shared__
#define __shared__ static
#define threadIdx (simt::cur->t_idx)
#define blockIdx simt::b_idx
#define gridDim simt::g_dim
#define blockDim simt::b_dim
#define __launch_bounds__(...)
static inline void __syncthreads() { simt::arrive_and_wait(simt::blk->cta); }
static inline void __syncwarp() { simt::arrive_and_wait(simt::blk->warp[simt::cur->t_idx.x >> 5]); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline double __shfl_xor_sync(unsigned, double v, int o) { return simt::shfl_xor(v, o); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return simt::shfl_idx(v, src); }
static inline int __shfl_sync(unsigned, int v, int src) { return simt::shfl_idx(v, src); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) {
  const unsigned old = *p;
  *p = old + v;
  return old;
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
