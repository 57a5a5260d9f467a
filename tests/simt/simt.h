// TEST INFRASTRUCTURE ONLY -- a small SIMT emulator so that the product's CUDA kernels
// (crafter_b200/csrc/cr_kernels.h, unmodified) run on a CPU: one fiber (ucontext) per CUDA thread,
// one block at a time, cooperative scheduling.
//
//   __syncthreads()        block barrier: released when every LIVE thread of the block waits at it
//   __syncwarp(), __ballot_sync, __shfl_sync, __shfl_up_sync, __reduce_or_sync
//                          warp barrier / collectives over the 32 consecutive threads of a warp
//   __shared__             `static`: one copy, shared by the fibers of the (single) running block
//   extern __shared__      CR_DYN_SMEM -> simt::dyn_smem(), filled with 0xCD before every block
//   atomics                plain read-modify-write (one OS thread)
//
// What it checks that tests/hostsim (one lane, sequential phases) cannot: multi-lane logic (ballots,
// prefix sums, order-preserving compaction), block-level choreography (barriers between phases,
// shared-memory carve-up and aliasing) and barrier divergence -- a block in which some live threads
// wait at a barrier the others never reach is reported as a deadlock instead of hanging.
// What it does not model: memory-ordering races between barriers, streams, graphs, TMA.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace simt {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Fiber {
  ucontext_t ctx;
  char *stack = nullptr;  // malloc'ed once, pages are touched only as deep as a thread really goes
  int tid = 0;
  bool done = false;
  int wait = 0;  // 0 runnable, 1 at the block barrier, 2 at the warp barrier
};

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 * 1024;

struct Runtime {
  ucontext_t main;
  Fiber fibers[MAX_THREADS];
  int n = 0;  // threads of the running block
  Fiber *cur = nullptr;
  Dim3 block_idx, block_dim, grid_dim;
  unsigned char *dyn = nullptr;  // a fresh exact-size allocation per block: overruns are ASan's to catch (tools/simt_asan.sh)
  uint32_t slot[MAX_THREADS / 32][32];  // [warp][lane] exchange buffer of the warp collectives
  const std::function<void()> *body = nullptr;
  const char *kernel = "";
  long launches = 0, blocks = 0;
};
inline Runtime &rt() { static Runtime *r = new Runtime(); return *r; }

inline unsigned char *dyn_smem() { return rt().dyn; }

inline void yield_to_scheduler() {
  Runtime &r = rt();
  swapcontext(&r.cur->ctx, &r.main);
}
inline void barrier_block() { rt().cur->wait = 1; yield_to_scheduler(); }
inline void barrier_warp() { rt().cur->wait = 2; yield_to_scheduler(); }

inline void trampoline() {
  Runtime &r = rt();
  (*r.body)();
  r.cur->done = true;
  swapcontext(&r.cur->ctx, &r.main);
}

// Run one block to completion.
inline void run_block() {
  Runtime &r = rt();
  const int n = r.n;
  // CR_SIMT_ORDER: forward (default), "reverse", or "shuffle": which runnable thread goes next is
  // not defined on a GPU either, and code that needs one particular order has a race.
  static const char *order_env = getenv("CR_SIMT_ORDER");
  const int mode = !order_env ? 0 : order_env[0] == 'r' ? 1 : order_env[0] == 's' ? 2 : 0;
  static uint32_t lcg = 12345u;
  for (;;) {
    bool progressed = false;
    const uint32_t mul = mode == 2 ? ((lcg = lcg * 1664525u + 1013904223u) >> 8 | 1u) : 1u;  // odd: a bijection mod 2^k
    const uint32_t add = mode == 2 ? lcg >> 16 : 0u;
    int pow2 = 1;
    while (pow2 < n) pow2 <<= 1;
    for (int j = 0; j < pow2; ++j) {
      int i = mode == 1 ? n - 1 - j : mode == 2 ? (int)((j * mul + add) & (uint32_t)(pow2 - 1)) : j;
      if (i < 0 || i >= n) continue;
      Fiber &f = r.fibers[i];
      if (f.done || f.wait) continue;
      r.cur = &f;
      swapcontext(&r.main, &f.ctx);
      progressed = true;
    }
    int live = 0, at_block = 0;
    for (int i = 0; i < n; ++i) { live += !r.fibers[i].done; at_block += !r.fibers[i].done && r.fibers[i].wait == 1; }
    if (live == 0) return;
    bool released = false;
    if (at_block == live) {
      for (int i = 0; i < n; ++i) r.fibers[i].wait = 0;
      released = true;
    } else {
      for (int w = 0; w * 32 < n; ++w) {
        int wl = 0, ww = 0;
        for (int i = w * 32; i < n && i < w * 32 + 32; ++i) { wl += !r.fibers[i].done; ww += !r.fibers[i].done && r.fibers[i].wait == 2; }
        if (wl && ww == wl) {
          for (int i = w * 32; i < n && i < w * 32 + 32; ++i) r.fibers[i].wait = 0;
          released = true;
        }
      }
    }
    if (!progressed && !released) {
      fprintf(stderr, "simt: DEADLOCK in %s, block %u: %d live threads, %d at __syncthreads; (thread:wait)",
              r.kernel, r.block_idx.x, live, at_block);
      for (int i = 0; i < n; ++i)
        if (!r.fibers[i].done) fprintf(stderr, " %d:%d", i, r.fibers[i].wait);
      fprintf(stderr, "\n");
      abort();
    }
  }
}

// kernel<<<grid, block, smem>>>(args...)  ->  simt::launch("name", grid, block, smem, [&] { kernel(args...); })
inline void launch(const char *name, unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
  Runtime &r = rt();
  if (block > (unsigned)MAX_THREADS || block == 0) { fprintf(stderr, "simt: bad block size %u\n", block); abort(); }
  r.kernel = name;
  r.body = &body;
  r.grid_dim.x = grid; r.block_dim.x = block;
  r.n = (int)block;
  r.launches += 1;
  for (unsigned b = 0; b < grid; ++b) {
    r.block_idx.x = b;
    r.blocks += 1;
    free(r.dyn);
    r.dyn = (unsigned char *)malloc(smem ? smem : 1);
    memset(r.dyn, 0xCD, smem);  // uninitialised shared memory must not look like zeros
    for (unsigned t = 0; t < block; ++t) {
      Fiber &f = r.fibers[t];
      f.tid = (int)t; f.done = false; f.wait = 0;
      if (!f.stack) f.stack = (char *)malloc(STACK_BYTES);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = STACK_BYTES;
      f.ctx.uc_link = &r.main;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    run_block();
  }
  r.body = nullptr;
}

struct Idx { unsigned x, y, z; };
inline Idx thread_idx() { return Idx{(unsigned)rt().cur->tid, 0, 0}; }
inline Idx block_idx() { return Idx{rt().block_idx.x, 0, 0}; }
inline Idx block_dim() { return Idx{rt().block_dim.x, 1, 1}; }
inline Idx grid_dim() { return Idx{rt().grid_dim.x, 1, 1}; }

// warp collectives: write my slot, barrier, combine, barrier (nobody overwrites before all have read)
inline uint32_t *my_slots(int &lane) {
  Runtime &r = rt();
  lane = r.cur->tid & 31;
  return r.slot[r.cur->tid >> 5];
}
inline bool lane_live(int lane) {
  Runtime &r = rt();
  const int i = (r.cur->tid & ~31) + lane;
  return i < r.n && !r.fibers[i].done;
}

}  // namespace simt

// ---- the CUDA surface the kernels use ------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define threadIdx (simt::thread_idx())
#define blockIdx (simt::block_idx())
#define blockDim (simt::block_dim())
#define gridDim (simt::grid_dim())

inline void __syncthreads() { simt::barrier_block(); }
inline void __threadfence() {}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier_warp(); }
inline unsigned __ballot_sync(unsigned, int pred) {
  int lane; uint32_t *s = simt::my_slots(lane);
  s[lane] = pred ? 1u : 0u;
  simt::barrier_warp();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) if (simt::lane_live(l) && s[l]) m |= 1u << l;
  simt::barrier_warp();
  return m;
}
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) {
  int lane; uint32_t *s = simt::my_slots(lane);
  s[lane] = v;
  simt::barrier_warp();
  const uint32_t out = s[src & 31];
  simt::barrier_warp();
  return out;
}
inline int __shfl_up_sync(unsigned, int v, unsigned delta) {
  int lane; uint32_t *s = simt::my_slots(lane);
  s[lane] = (uint32_t)v;
  simt::barrier_warp();
  const int out = lane >= (int)delta ? (int)s[lane - delta] : v;
  simt::barrier_warp();
  return out;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  int lane; uint32_t *s = simt::my_slots(lane);
  s[lane] = v;
  simt::barrier_warp();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) if (simt::lane_live(l)) m |= s[l];
  simt::barrier_warp();
  return m;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __vcmpeq4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int k = 0; k < 4; ++k) if (((a >> (8 * k)) & 0xFF) == ((b >> (8 * k)) & 0xFF)) r |= 0xFFu << (8 * k);
  return r;
}
inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
