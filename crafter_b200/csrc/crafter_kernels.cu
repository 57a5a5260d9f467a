// crafter_b200: sm_100a kernels + the C ABI of include/crafter_b200.h.
//
// Step graph (one CUDA graph per handle, captured on first use):
//
//   memset(work lists) -> k_update -+-> k_post (balance) ------------+-> k_render ---------+-> end
//                  (warp per env)   |                                |                    |
//                                   +-> k_install (swap in the ------+                    |
//                                       prefetched worlds) -> k_wg_mat -> k_wg_obj -------+
//                                                                     \-> k_seed (ahead) -+
//                                       (prefetch the NEXT world of the finished envs)
//
// With CRAFTER_B200_DEFER_WG=1 (two prefetched worlds per env) the regeneration of
// the buffers consumed in step t-1 runs from the ROOT of step t's graph, beside k_update / k_post:
//
//   k_seed2 -> k_wg_mat -> (k_wg_obj || k_seed2 ahead) ------------> [install done] k_pending_copy -+
//   memset -> k_update -+-> k_post ------------------+-> k_render ---------------------------------+-> end
//                       +-> k_install (buffer CUR) --+
//
// World generation is FP64-heavy and latency-bound; it runs on a forked branch next to the render
// kernel (integer / LSU bound) and fills the `next_*` buffers, so it never delays the observation.
// Compile with -fmad=false: the reference's numpy / PIL arithmetic has no fused multiply-adds, and
// terrain thresholds / truncating casts see the last bit.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/crafter_b200.h"
#include "cr_common.h"
#include "cr_geom.h"
#include "cr_noise.h"
#include "cr_render.h"
#include "cr_update.h"
#include "cr_worldgen.h"
#include "cr_kernels.h"

using namespace cr;
using namespace cr::kernels;

namespace {

// launch the default-geometry instantiation when the handle's geometry matches
#define CR_LAUNCH(KERNEL, DEFAULT, GRID, BLOCK, SMEM, STREAM, ...)                       \
  do {                                                                                   \
    if (DEFAULT) KERNEL<true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);               \
    else KERNEL<false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                      \
  } while (0)

// the kernels themselves: cr_kernels.h

thread_local char g_error[512] = "";

int fail(const char *what, cudaError_t err, int line) {
  snprintf(g_error, sizeof(g_error), "crafter_b200: %s failed at line %d: %s", what, line,
           cudaGetErrorString(err));
  return -1;
}
int fail_msg(const char *msg) {
  snprintf(g_error, sizeof(g_error), "crafter_b200: %s", msg);
  return -2;
}

#define CR_CUDA(expr)                                        \
  do {                                                       \
    cudaError_t err_ = (expr);                               \
    if (err_ != cudaSuccess) return fail(#expr, err_, __LINE__); \
  } while (0)

}  // namespace

struct GraphSlot {
  cudaGraphExec_t exec;
  const void *actions, *obs, *reward, *done, *reward_host, *done_host;
  int kernels;
};

struct cr_handle {
  int device;  // the device current at cr_create; every entry point runs on it (DeviceGuard)
  Geom g;
  State st;
  RenderTables rt;
  int auto_reset;
  int use_graph;
  int num_sms;
  size_t update_smem, render_smem, balance_smem;
  int balance_threads;
  int render_staged;
  int64_t launches;
  cudaStream_t side, side2;     // worldgen branch, seed-ahead branch
  cudaEvent_t ev_fork, ev_join, ev_mat, ev_ahead, ev_inst;
  int fused;                    // CRAFTER_B200_FUSED=1 (experiment, needs DEFER_WG): k_tick_render
  int split_render;             // CRAFTER_B200_SPLIT=1 (experiment): early / late render launches
  cudaStream_t side3;           // early-render branch
  cudaEvent_t ev_early;
  int defer;                    // CRAFTER_B200_DEFER_WG=1: two prefetch buffers, regeneration beside the next tick
  cudaStream_t side_w, side_a;  // deferred mode: regeneration branch, its seed-ahead branch
  cudaEvent_t ev_root, ev_join_w;
  // CRAFTER_B200_TIMING=1: eager launches bracketed by events, per-kernel warm durations
  int is_default;  // geometry == the reference's defaults: launch the constant-folded kernels
  int timing;
  int debug_skip;  // CRAFTER_B200_DEBUG_SKIP: timing experiments only (1 no balance, 2 no entities)
  cudaEvent_t t_ev[8][2];
  double t_ms[8];
  int64_t t_n;
  GraphSlot slots[2];  // cached step graphs: [0] device buffers only, [1] with the host copies
  // cr_step_host: D2H of reward/done forks right after k_update (inside the graph)
  float *d2h_reward;
  uint8_t *d2h_done;
  cudaEvent_t ev_upd, ev_d2h;
};

namespace {

enum { TK_UPDATE = 0, TK_INSTALL, TK_RENDER, TK_SEED, TK_MAT, TK_OBJ, TK_AHEAD, TK_BALANCE, TK_COUNT };
inline void tmark(cr_handle *h, int id, int end, cudaStream_t s) {
  if (h->timing) cudaEventRecord(h->t_ev[id][end], s);
}

// seed -> terrain -> creatures into the next_* buffers of the listed envs, on stream `s`.
// With `ahead`, the seed of the following world is prepared on the second side stream while
// k_wg_obj runs (forked after k_wg_mat, which is the last reader of the permutation table).
int launch_worldgen(cr_handle *h, cudaStream_t s, int only_invalid, int ahead, int seeded) {
  const Geom &g = h->g;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  int seed_grid = (g.B + SEED_WPB - 1) / SEED_WPB;
  if (seed_grid > h->num_sms * 4) seed_grid = h->num_sms * 4;
  if (!seeded) {  // the step graph skips this: every listed env was seeded ahead and promoted
    tmark(h, TK_SEED, 0, s);
    k_seed<<<seed_grid, SEED_WPB * 32, 0, s>>>(g, h->st, only_invalid, 0);
    tmark(h, TK_SEED, 1, s);
  }
  long long want = (long long)g.B * tiles;
  int mat_grid = (int)(want < (long long)h->num_sms * 16 ? want : (long long)h->num_sms * 16);
  tmark(h, TK_MAT, 0, s);
  CR_LAUNCH(k_wg_mat, h->is_default, mat_grid, WG_THREADS, 0, s, g, h->st, only_invalid);
  tmark(h, TK_MAT, 1, s);
  int n = seeded ? 2 : 3;
  if (ahead) {
    CR_CUDA(cudaEventRecord(h->ev_mat, s));
    CR_CUDA(cudaStreamWaitEvent(h->side2, h->ev_mat, 0));
    tmark(h, TK_AHEAD, 0, h->side2);
    k_seed<<<seed_grid, SEED_WPB * 32, 0, h->side2>>>(g, h->st, 0, 1);
    tmark(h, TK_AHEAD, 1, h->side2);
    CR_CUDA(cudaEventRecord(h->ev_ahead, h->side2));
    n += 1;
  }
  int obj_grid = g.B < h->num_sms * 2 ? g.B : h->num_sms * 2;
  tmark(h, TK_OBJ, 0, s);
  CR_LAUNCH(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, s, g, h->st, only_invalid);
  tmark(h, TK_OBJ, 1, s);
  if (ahead) CR_CUDA(cudaStreamWaitEvent(s, h->ev_ahead, 0));
  CR_CUDA(cudaGetLastError());
  return n;
}

// Deferred mode: one regeneration pass over the entries of `stl`'s list (the handle's state with
// reset_list / reset_count pointing at the list to serve) on stream `s`.  `fork_ahead`: the seed
// of the following world runs on side_a next to k_wg_obj and is joined back into `s`.
int launch_worldgen2(cr_handle *h, cudaStream_t s, const State &stl, int fork_ahead) {
  const Geom &g = h->g;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  int seed_grid = (g.B + SEED_WPB - 1) / SEED_WPB;
  if (seed_grid > h->num_sms * 4) seed_grid = h->num_sms * 4;
  k_seed2<<<seed_grid, SEED_WPB * 32, 0, s>>>(g, stl, 0);
  long long want = (long long)g.B * tiles;
  int mat_grid = (int)(want < (long long)h->num_sms * 16 ? want : (long long)h->num_sms * 16);
  CR_LAUNCH(k_wg_mat, h->is_default, mat_grid, WG_THREADS, 0, s, g, stl, 0);
  cudaStream_t sa = fork_ahead ? h->side_a : s;
  if (fork_ahead) {
    CR_CUDA(cudaEventRecord(h->ev_mat, s));
    CR_CUDA(cudaStreamWaitEvent(sa, h->ev_mat, 0));
  }
  k_seed2<<<seed_grid, SEED_WPB * 32, 0, sa>>>(g, stl, 1);  // after k_wg_mat, the last reader of perm
  if (fork_ahead) CR_CUDA(cudaEventRecord(h->ev_ahead, sa));
  int obj_grid = g.B < h->num_sms * 2 ? g.B : h->num_sms * 2;
  CR_LAUNCH(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, s, g, stl, 0);
  if (fork_ahead) CR_CUDA(cudaStreamWaitEvent(s, h->ev_ahead, 0));
  CR_CUDA(cudaGetLastError());
  return 4;
}

// The handle's state with the pending list in the place of the reset list.
State pending_view(const cr_handle *h) {
  State v = h->st;
  v.reset_list = h->st.pend_list;
  v.reset_count = h->st.pend_count;
  return v;
}

int launch_install(cr_handle *h, cudaStream_t s) {
  int grid = h->g.B < h->num_sms * 8 ? h->g.B : h->num_sms * 8;
  tmark(h, TK_INSTALL, 0, s);
  CR_LAUNCH(k_install, h->is_default, grid, INSTALL_THREADS, 0, s, h->g, h->st);
  tmark(h, TK_INSTALL, 1, s);
  CR_CUDA(cudaGetLastError());
  return 1;
}

int launch_render(cr_handle *h, uint8_t *obs, cudaStream_t s, const int32_t *env_list = nullptr,
                  int n_envs = -1, const uint8_t *done = nullptr, int part = RENDER_ALL) {
  if (part == RENDER_ALL) {
    tmark(h, TK_RENDER, 0, s);
    CR_LAUNCH(k_render, h->is_default, n_envs < 0 ? h->g.B : n_envs, RENDER_THREADS, h->render_smem, s, h->g,
              h->st, h->rt, obs, h->render_staged, env_list, nullptr, 0);
    tmark(h, TK_RENDER, 1, s);
    CR_CUDA(cudaGetLastError());
    return 1;
  }
#define CR_RENDER_PART(DEF, PART)                                                                    \
  k_render<DEF, PART><<<h->g.B, RENDER_THREADS, h->render_smem, s>>>(h->g, h->st, h->rt, obs,       \
                                                                      h->render_staged, nullptr, done, h->auto_reset)
  if (part == RENDER_EARLY) { if (h->is_default) CR_RENDER_PART(true, RENDER_EARLY); else CR_RENDER_PART(false, RENDER_EARLY); }
  else if (part == RENDER_LATE) { if (h->is_default) CR_RENDER_PART(true, RENDER_LATE); else CR_RENDER_PART(false, RENDER_LATE); }
  else { if (h->is_default) CR_RENDER_PART(true, RENDER_RESET); else CR_RENDER_PART(false, RENDER_RESET); }
#undef CR_RENDER_PART
  CR_CUDA(cudaGetLastError());
  return 1;
}
// render on `s`, worldgen prefetch for the listed envs on the side stream, joined back into `s`.
// Works eagerly and under stream capture (the side stream joins the capture through the event).
int launch_render_and_prefetch(cr_handle *h, uint8_t *obs, cudaStream_t s, int seeded) {
  CR_CUDA(cudaEventRecord(h->ev_fork, s));
  CR_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  int n = 0, k;
  if (obs) {
    if ((k = launch_render(h, obs, s)) < 0) return k;
    n += k;
  }
  if ((k = launch_worldgen(h, h->side, 0, 1, seeded)) < 0) return k;
  n += k;
  CR_CUDA(cudaEventRecord(h->ev_join, h->side));
  CR_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
  return n;
}

// CRAFTER_B200_FUSED: W(prev) from the root | memset -> k_tick_render<BALANCE> || k_tick_render<PLAIN>
// -> k_install -> k_render<RESET>; the refill branch ends with k_pending_copy as in the deferred mode.
int enqueue_step_fused(cr_handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done,
                       cudaStream_t s) {
  const Geom &g = h->g;
  int n = 0, k;
  CR_CUDA(cudaEventRecord(h->ev_root, s));
  CR_CUDA(cudaStreamWaitEvent(h->side_w, h->ev_root, 0));
  if ((k = launch_worldgen2(h, h->side_w, pending_view(h), 1)) < 0) return k;
  n += k;
  if (h->st.balance_count == h->st.reset_count + 1) {
    CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, 2 * sizeof(int32_t), s));
  } else {
    CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, sizeof(int32_t), s));
    CR_CUDA(cudaMemsetAsync(h->st.balance_count, 0, sizeof(int32_t), s));
  }
  CR_CUDA(cudaEventRecord(h->ev_fork, s));
  CR_CUDA(cudaStreamWaitEvent(h->side3, h->ev_fork, 0));
#define CR_TICK(DEF, CLS, STREAM)                                                                     \
  k_tick_render<DEF, CLS><<<g.B, RENDER_THREADS, h->render_smem, STREAM>>>(g, h->st, h->rt, actions, obs, \
                                                                           reward, done, h->auto_reset)
  if (h->fused == 2) {  // one launch, every env: the measured-slower shape, now with the tick inside
    if (h->is_default) CR_TICK(true, TICK_ANY, s); else CR_TICK(false, TICK_ANY, s);
    n += 1;
  } else if (h->is_default) { CR_TICK(true, TICK_BALANCE, s); CR_TICK(true, TICK_PLAIN, h->side3); n += 2; }
  else { CR_TICK(false, TICK_BALANCE, s); CR_TICK(false, TICK_PLAIN, h->side3); n += 2; }
#undef CR_TICK
  CR_CUDA(cudaGetLastError());
  CR_CUDA(cudaEventRecord(h->ev_early, h->side3));
  CR_CUDA(cudaStreamWaitEvent(s, h->ev_early, 0));
  const bool d2h = h->d2h_reward && h->d2h_done;
  if (d2h) {
    CR_CUDA(cudaEventRecord(h->ev_upd, s));
    CR_CUDA(cudaStreamWaitEvent(h->side2, h->ev_upd, 0));
    CR_CUDA(cudaMemcpyAsync(h->d2h_reward, reward, (size_t)g.B * sizeof(float), cudaMemcpyDeviceToHost, h->side2));
    CR_CUDA(cudaMemcpyAsync(h->d2h_done, done, (size_t)g.B, cudaMemcpyDeviceToHost, h->side2));
    CR_CUDA(cudaEventRecord(h->ev_d2h, h->side2));
  }
  if ((k = launch_install(h, s)) < 0) return k;
  n += k;
  CR_CUDA(cudaEventRecord(h->ev_inst, s));
  if ((k = launch_render(h, obs, s, nullptr, -1, done, RENDER_RESET)) < 0) return k;
  n += k;
  CR_CUDA(cudaStreamWaitEvent(h->side_w, h->ev_inst, 0));
  k_pending_copy<<<g.B < 16384 ? 1 : 8, 256, 0, h->side_w>>>(h->st);
  CR_CUDA(cudaGetLastError());
  n += 1;
  CR_CUDA(cudaEventRecord(h->ev_join_w, h->side_w));
  CR_CUDA(cudaStreamWaitEvent(s, h->ev_join_w, 0));
  if (d2h) CR_CUDA(cudaStreamWaitEvent(s, h->ev_d2h, 0));
  return n;
}

// Enqueue one tick; returns the number of kernels or a negative error.
int enqueue_step(cr_handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done,
                 cudaStream_t s) {
  const Geom &g = h->g;
  int n = 0, k;
  if (h->fused && h->auto_reset) return enqueue_step_fused(h, actions, obs, reward, done, s);
  const bool defer = h->defer && h->auto_reset;
  if (defer) {
    // the buffers consumed by the previous step are refilled from the root of this one, beside
    // the tick: nothing they touch is read or written by k_update / k_install
    CR_CUDA(cudaEventRecord(h->ev_root, s));
    CR_CUDA(cudaStreamWaitEvent(h->side_w, h->ev_root, 0));
    if ((k = launch_worldgen2(h, h->side_w, pending_view(h), 1)) < 0) return k;
    n += k;
  }
  // reset_count and balance_count are adjacent words (see Env._alloc_state): one memset node
  if (h->st.balance_count == h->st.reset_count + 1) {
    CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, 2 * sizeof(int32_t), s));
  } else {
    CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, sizeof(int32_t), s));
    CR_CUDA(cudaMemsetAsync(h->st.balance_count, 0, sizeof(int32_t), s));
  }
  tmark(h, TK_UPDATE, 0, s);
  CR_LAUNCH(k_update, h->is_default, (g.B + UPDATE_WPB - 1) / UPDATE_WPB, UPDATE_WPB * 32, h->update_smem,
            s, g, h->st, h->rt.daylight, actions, reward, done, h->auto_reset, h->debug_skip);
  tmark(h, TK_UPDATE, 1, s);
  CR_CUDA(cudaGetLastError());
  n += 1;
  const bool d2h = h->d2h_reward && h->d2h_done;
  if (d2h) {  // reward / done are final after the tick: copy them out while the rest of the step runs
    CR_CUDA(cudaEventRecord(h->ev_upd, s));
    CR_CUDA(cudaStreamWaitEvent(h->side2, h->ev_upd, 0));
    CR_CUDA(cudaMemcpyAsync(h->d2h_reward, reward, (size_t)g.B * sizeof(float), cudaMemcpyDeviceToHost, h->side2));
    CR_CUDA(cudaMemcpyAsync(h->d2h_done, done, (size_t)g.B, cudaMemcpyDeviceToHost, h->side2));
    CR_CUDA(cudaEventRecord(h->ev_d2h, h->side2));
  }
  const int bal_ctas = g.B < h->num_sms * 4 ? g.B : h->num_sms * 4;
  if (!h->auto_reset) {
    CR_LAUNCH(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, s, g, h->st,
              h->rt.daylight, bal_ctas);
    if ((k = launch_render(h, obs, s)) < 0) return k;
    if (d2h) CR_CUDA(cudaStreamWaitEvent(s, h->ev_d2h, 0));
    return n + 1 + k;
  }
  // Two branches after the tick:
  //   main   k_post (balance) ---------------------> [wait install] k_render -------> [join]
  //   side   k_install -> k_wg_mat -> (k_wg_obj || k_seed ahead) ----------------------^
  // The render needs both the balanced and the re-installed envs; world generation only the
  // install, so it starts ~30 us earlier than behind a combined post kernel.
  CR_CUDA(cudaEventRecord(h->ev_fork, s));
  CR_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  if ((k = launch_install(h, h->side)) < 0) return k;
  n += k;
  CR_CUDA(cudaEventRecord(h->ev_inst, h->side));
  if (h->split_render) {  // envs the tick left final are drawn beside k_post / k_install
    CR_CUDA(cudaStreamWaitEvent(h->side3, h->ev_fork, 0));
    if ((k = launch_render(h, obs, h->side3, nullptr, -1, done, RENDER_EARLY)) < 0) return k;
    n += k;
    CR_CUDA(cudaEventRecord(h->ev_early, h->side3));
  }
  tmark(h, TK_BALANCE, 0, s);
  CR_LAUNCH(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, s, g, h->st,
            h->rt.daylight, bal_ctas);
  tmark(h, TK_BALANCE, 1, s);
  n += 1;
  CR_CUDA(cudaStreamWaitEvent(s, h->ev_inst, 0));
  if ((k = launch_render(h, obs, s, nullptr, -1, done, h->split_render ? RENDER_LATE : RENDER_ALL)) < 0) return k;
  n += k;
  if (h->split_render) CR_CUDA(cudaStreamWaitEvent(s, h->ev_early, 0));
  if (defer) {
    // tail of the regeneration branch: once k_install has named the consumed buffers, its list
    // becomes the pending list of the next step
    CR_CUDA(cudaStreamWaitEvent(h->side_w, h->ev_inst, 0));
    k_pending_copy<<<g.B < 16384 ? 1 : 8, 256, 0, h->side_w>>>(h->st);
    CR_CUDA(cudaGetLastError());
    n += 1;
    CR_CUDA(cudaEventRecord(h->ev_join_w, h->side_w));
    CR_CUDA(cudaStreamWaitEvent(s, h->ev_join_w, 0));
  } else {
    if ((k = launch_worldgen(h, h->side, 0, 1, 1)) < 0) return k;
    n += k;
    CR_CUDA(cudaEventRecord(h->ev_join, h->side));
    CR_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
  }
  if (d2h) CR_CUDA(cudaStreamWaitEvent(s, h->ev_d2h, 0));
  return n;
}

// Env.reset in deferred mode, everything in stream order on `s` (the explicit path is not the hot
// one): serve the pending regenerations, make sure both buffers of the listed envs hold the worlds
// of the next two episodes, consume the first, and refill it beside the render.
int reset_deferred(cr_handle *h, const uint8_t *mask, uint8_t *obs, cudaStream_t s) {
  const Geom &g = h->g;
  const int list_grid = (g.B + 255) / 256;
  int k;
  if ((k = launch_worldgen2(h, s, pending_view(h), 0)) < 0) return k;
  h->launches += k;
  CR_CUDA(cudaMemsetAsync(h->st.pend_count, 0, sizeof(int32_t), s));
  CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, sizeof(int32_t), s));
  k_fill_list<<<list_grid, 256, 0, s>>>(g.B, mask, h->st.reset_list, h->st.reset_count);
  for (int which = 0; which < 2; ++which) {
    k_prep<<<list_grid, 256, 0, s>>>(g, h->st, which);
    CR_CUDA(cudaGetLastError());
    if ((k = launch_worldgen2(h, s, h->st, 0)) < 0) return k;
    h->launches += k + 1;
  }
  if ((k = launch_install(h, s)) < 0) return k;  // entries now name the consumed buffers
  h->launches += k + 1;
  CR_CUDA(cudaEventRecord(h->ev_fork, s));
  CR_CUDA(cudaStreamWaitEvent(h->side_w, h->ev_fork, 0));
  if (obs) {
    if ((k = launch_render(h, obs, s)) < 0) return k;
    h->launches += k;
  }
  if ((k = launch_worldgen2(h, h->side_w, h->st, 1)) < 0) return k;
  h->launches += k;
  CR_CUDA(cudaEventRecord(h->ev_join_w, h->side_w));
  CR_CUDA(cudaStreamWaitEvent(s, h->ev_join_w, 0));
  return 0;
}

}  // namespace

// The handle's streams, events and graphs belong to the device that was current in cr_create;
// entry points switch to it (and back) when the caller's current device differs, so host code
// needs no device context manager around the calls.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int want) {
    int cur = want;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != want) { prev = cur; cudaSetDevice(want); }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" {

int cr_abi_version(void) { return CR_ABI_VERSION; }
const char *cr_last_error(void) { return g_error; }

int cr_create(const cr_config *c, const cr_tables *t, const cr_state *s, cr_handle **out) {
  if (!c || !t || !s || !out) return fail_msg("null argument");
  cr_handle *h = (cr_handle *)calloc(1, sizeof(cr_handle));
  Geom &g = h->g;
  if (const char *msg = geom_from_config(*c, g)) {
    free(h);
    return fail_msg(msg);
  }
  state_from_abi(*s, h->st);
  h->is_default = geom_is_default(g) ? 1 : 0;
  if (const char *nd = getenv("CRAFTER_B200_NO_SPECIALIZE")) if (nd[0] == '1') h->is_default = 0;
  h->rt.mat_tex = t->mat_tex; h->rt.obj_tex = t->obj_tex; h->rt.item_tile = t->item_tile;
  h->rt.vignette = t->vignette; h->rt.daylight = t->daylight; h->rt.colx = t->colx;
  h->rt.rowy = t->rowy;
  h->auto_reset = c->auto_reset;
  const char *ng = getenv("CRAFTER_B200_NO_GRAPH");
  const char *tm = getenv("CRAFTER_B200_TIMING");
  h->timing = tm && tm[0] == '1';
  const char *ds = getenv("CRAFTER_B200_DEBUG_SKIP");
  h->debug_skip = ds ? atoi(ds) : 0;
  h->use_graph = !(ng && ng[0] == '1') && !h->timing;
  const char *sp = getenv("CRAFTER_B200_SPLIT");
  h->split_render = sp && sp[0] == '1' && !h->timing;
  const char *dw = getenv("CRAFTER_B200_DEFER_WG");
  h->defer = dw && dw[0] == '1';
  if (h->defer && !state_has_defer_buffers(h->st)) {
    free(h);
    return fail_msg("CRAFTER_B200_DEFER_WG=1 needs next_mat2 / next_ents2 / next_meta2 / pend_list / pend_count");
  }
  if (h->defer && h->timing) { free(h); return fail_msg("CRAFTER_B200_TIMING is not available with CRAFTER_B200_DEFER_WG"); }
  g.defer = h->defer;
  // Pure work reductions, on unless switched off for an A/B run (=0): the tick's first 32 keyed draws
  // by all lanes at once; grass / path cells per chunk kept current by the terrain writes (needs
  // the caller's chunk_cnt buffer; without it every balance tick re-counts the cells).
  const char *dp = getenv("CRAFTER_B200_DRAW_PREFETCH");
  g.draw_prefetch = !(dp && dp[0] == '0');
  const char *ic = getenv("CRAFTER_B200_INCR_CENSUS");
  g.incr_census = !(ic && ic[0] == '0') && h->st.chunk_cnt != nullptr;
  int dev = 0;
  CR_CUDA(cudaGetDevice(&dev));
  h->device = dev;
  CR_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
  int max_smem = 0;
  CR_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  h->update_smem = UPDATE_WPB * update_smem_per_warp(g);
  if (h->update_smem > (size_t)max_smem) { free(h); return fail_msg("view too large for the update window"); }
  h->balance_smem = balance_smem(g);
  h->balance_threads = g.NCH * 3 > 4 * BALANCE_THREADS ? BALANCE_THREADS_MAX : BALANCE_THREADS;
  if (h->balance_smem > (size_t)max_smem) { free(h); return fail_msg("area too large for k_balance"); }
  CR_CUDA(cudaFuncSetAttribute(k_post<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->balance_smem));
  CR_CUDA(cudaFuncSetAttribute(k_post<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->balance_smem));
  size_t tile = align16((size_t)g.sw * g.sh * 3);
  size_t fixed = align16(sizeof(RenderShared)) +
                 (g.tile_cache ? align16((size_t)(N_TILES + 1) * g.ux * g.uy * sizeof(uint32_t)) : 16);
  if (fixed > (size_t)max_smem) { free(h); return fail_msg("unit too large for the tile cache"); }
  // keep at least two CTAs per SM when staging the output tile
  h->render_staged = fixed + tile <= (size_t)max_smem / 2;
  h->render_smem = fixed + (h->render_staged ? tile : 0);
  if (!h->render_staged) h->is_default = 0;  // the constant-folded k_render has no unstaged path
  CR_CUDA(cudaFuncSetAttribute(k_render<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->render_smem));
  CR_CUDA(cudaFuncSetAttribute(k_render<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->render_smem));
  {
    const char *fu = getenv("CRAFTER_B200_FUSED");
    const size_t tick = align16(sizeof(Ent) * ENT_SMEM) + align16(sizeof(uint32_t) * g.TW);
    const size_t bal = h->balance_smem - align16(sizeof(PlayerS));
    const size_t scratch = align16(sizeof(PlayerS)) + (tick > bal ? tick : bal);
    h->fused = fu && (fu[0] == '1' || fu[0] == '2') && h->defer && h->auto_reset && !h->timing &&
                       h->render_staged && g.tile_cache && scratch <= tile
                   ? fu[0] - '0' : 0;  // else the knob falls back to the deferred schedule
  }
  if (h->fused) {
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<true, TICK_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<true, TICK_BALANCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<false, TICK_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<false, TICK_BALANCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<true, TICK_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_tick_render<false, TICK_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_render<true, RENDER_RESET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_render<false, RENDER_RESET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
  }
  if (h->split_render || h->fused) {
    CR_CUDA(cudaStreamCreateWithFlags(&h->side3, cudaStreamNonBlocking));
    CR_CUDA(cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming));
  }
  if (h->split_render) {
    CR_CUDA(cudaFuncSetAttribute(k_render<true, RENDER_EARLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_render<true, RENDER_LATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_render<false, RENDER_EARLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
    CR_CUDA(cudaFuncSetAttribute(k_render<false, RENDER_LATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->render_smem));
  }
  CR_CUDA(cudaFuncSetAttribute(k_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->update_smem));
  CR_CUDA(cudaFuncSetAttribute(k_update<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h->update_smem));
  if (h->timing)
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 2; ++j) CR_CUDA(cudaEventCreate(&h->t_ev[i][j]));
  CR_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CR_CUDA(cudaStreamCreateWithFlags(&h->side2, cudaStreamNonBlocking));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_mat, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_ahead, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_inst, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_upd, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_d2h, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CR_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (h->defer) {
    CR_CUDA(cudaStreamCreateWithFlags(&h->side_w, cudaStreamNonBlocking));
    CR_CUDA(cudaStreamCreateWithFlags(&h->side_a, cudaStreamNonBlocking));
    CR_CUDA(cudaEventCreateWithFlags(&h->ev_root, cudaEventDisableTiming));
    CR_CUDA(cudaEventCreateWithFlags(&h->ev_join_w, cudaEventDisableTiming));
  }
  *out = h;
  return 0;
}

int cr_destroy(cr_handle *h) {
  if (!h) return 0;
  DeviceGuard on_device(h->device);
  for (int i = 0; i < 2; ++i)
    if (h->slots[i].exec) cudaGraphExecDestroy(h->slots[i].exec);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side2) cudaStreamDestroy(h->side2);
  if (h->ev_mat) cudaEventDestroy(h->ev_mat);
  if (h->ev_ahead) cudaEventDestroy(h->ev_ahead);
  if (h->ev_inst) cudaEventDestroy(h->ev_inst);
  if (h->ev_upd) cudaEventDestroy(h->ev_upd);
  if (h->ev_d2h) cudaEventDestroy(h->ev_d2h);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side3) cudaStreamDestroy(h->side3);
  if (h->ev_early) cudaEventDestroy(h->ev_early);
  if (h->side_w) cudaStreamDestroy(h->side_w);
  if (h->side_a) cudaStreamDestroy(h->side_a);
  if (h->ev_root) cudaEventDestroy(h->ev_root);
  if (h->ev_join_w) cudaEventDestroy(h->ev_join_w);
  free(h);
  return 0;
}

int cr_reset(cr_handle *h, const uint8_t *mask, uint8_t *obs, void *stream) {
  if (!h) return fail_msg("null handle");
  DeviceGuard on_device(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int k;
  if (h->defer) return reset_deferred(h, mask, obs, s);
  CR_CUDA(cudaMemsetAsync(h->st.reset_count, 0, sizeof(int32_t), s));
  k_fill_list<<<(h->g.B + 255) / 256, 256, 0, s>>>(h->g.B, mask, h->st.reset_list, h->st.reset_count);
  CR_CUDA(cudaGetLastError());
  h->launches += 1;
  // worlds that were never prefetched (first reset) are generated now, then swapped in ...
  if ((k = launch_worldgen(h, s, 1, 0, 0)) < 0) return k;
  h->launches += k;
  if ((k = launch_install(h, s)) < 0) return k;
  h->launches += k;
  // ... and the following episode's worlds are prefetched next to the render
  if ((k = launch_render_and_prefetch(h, obs, s, 0)) < 0) return k;
  h->launches += k;
  return 0;
}

int cr_step(cr_handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done,
            void *stream) {
  if (!h || !actions || !obs || !reward || !done) return fail_msg("null argument");
  DeviceGuard on_device(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  bool legacy = s == nullptr || s == cudaStreamLegacy;
  if (!h->use_graph || legacy) {
    int n = enqueue_step(h, actions, obs, reward, done, s);
    if (n < 0) return n;
    h->launches += n;
    if (h->timing && h->auto_reset) {
      CR_CUDA(cudaStreamSynchronize(s));
      for (int i = 0; i < TK_COUNT; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->t_ev[i][0], h->t_ev[i][1]) == cudaSuccess) h->t_ms[i] += ms;
      }
      h->t_n += 1;
    }
    return 0;
  }
  GraphSlot &gs = h->slots[h->d2h_reward ? 1 : 0];  // device-only step and host-buffer step
  if (!gs.exec || gs.actions != actions || gs.obs != obs || gs.reward != reward || gs.done != done ||
      gs.reward_host != h->d2h_reward || gs.done_host != h->d2h_done) {
    if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    CR_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int n = enqueue_step(h, actions, obs, reward, done, s);
    cudaError_t end = cudaStreamEndCapture(s, &graph);
    if (n < 0) { if (graph) cudaGraphDestroy(graph); return n; }
    if (end != cudaSuccess) return fail("cudaStreamEndCapture", end, __LINE__);
    cudaError_t inst = cudaGraphInstantiate(&gs.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) { gs.exec = nullptr; return fail("cudaGraphInstantiate", inst, __LINE__); }
    gs.actions = actions; gs.obs = obs; gs.reward = reward; gs.done = done;
    gs.reward_host = h->d2h_reward; gs.done_host = h->d2h_done;
    gs.kernels = n;
  }
  CR_CUDA(cudaGraphLaunch(gs.exec, s));
  h->launches += gs.kernels;
  return 0;
}

int cr_step_host(cr_handle *h, const int32_t *actions_host, uint8_t *obs_host, float *reward_host,
                 uint8_t *done_host, int32_t *actions_dev, uint8_t *obs_dev, float *reward_dev,
                 uint8_t *done_dev, void *stream) {
  if (!h || !actions_host || !reward_host || !done_host) return fail_msg("null argument");
  DeviceGuard on_device(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t B = (size_t)h->g.B;
  CR_CUDA(cudaMemcpyAsync(actions_dev, actions_host, B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  // reward / done leave on a branch of the step graph right after the tick (same host buffers
  // every call keep the graph; new ones re-capture it)
  h->d2h_reward = reward_host;
  h->d2h_done = done_host;
  int rc = cr_step(h, actions_dev, obs_dev, reward_dev, done_dev, stream);
  h->d2h_reward = nullptr;
  h->d2h_done = nullptr;
  if (rc) return rc;
  if (obs_host)
    CR_CUDA(cudaMemcpyAsync(obs_host, obs_dev, B * h->g.sw * h->g.sh * 3, cudaMemcpyDeviceToHost, s));
  CR_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int cr_render(cr_handle *h, uint8_t *obs, void *stream) {
  if (!h || !obs) return fail_msg("null argument");
  DeviceGuard on_device(h->device);
  int k = launch_render(h, obs, (cudaStream_t)stream);
  if (k < 0) return k;
  h->launches += k;
  return 0;
}

int cr_render_envs(cr_handle *h, const int32_t *env_ids, int n, uint8_t *obs, void *stream) {
  if (!h || !obs || !env_ids) return fail_msg("null argument");
  DeviceGuard on_device(h->device);
  if (n < 0 || n > h->g.B) return fail_msg("cr_render_envs: n out of range");
  if (n == 0) return 0;
  int k = launch_render(h, obs, (cudaStream_t)stream, env_ids, n);
  if (k < 0) return k;
  h->launches += k;
  return 0;
}

int cr_semantic(cr_handle *h, uint8_t *out, void *stream) {
  if (!h || !out) return fail_msg("null argument");
  DeviceGuard on_device(h->device);
  size_t n = (size_t)h->g.B * h->g.NC;
  k_semantic<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->g, h->st, out);
  CR_CUDA(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int cr_recount(cr_handle *h, void *stream) {
  if (!h) return fail_msg("null handle");
  if (!h->g.incr_census) return 0;
  DeviceGuard on_device(h->device);
  const int grid = h->g.B < h->num_sms * 8 ? h->g.B : h->num_sms * 8;
  k_recount<<<grid, INSTALL_THREADS, 0, (cudaStream_t)stream>>>(h->g, h->st);
  CR_CUDA(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int64_t cr_launch_count(const cr_handle *h) { return h ? h->launches : 0; }

/* Profiling aid (CRAFTER_B200_TIMING=1): mean device milliseconds per kernel of the step, in the
 * order update, install, render, seed, wg_mat, wg_obj, seed_ahead; returns the number of steps. */
int64_t cr_timing(cr_handle *h, double *out_ms) {
  if (!h || !h->timing || h->t_n == 0) return 0;
  for (int i = 0; i < TK_COUNT; ++i) out_ms[i] = h->t_ms[i] / (double)h->t_n;
  int64_t n = h->t_n;
  for (int i = 0; i < 8; ++i) h->t_ms[i] = 0;
  h->t_n = 0;
  return n;
}

}  // extern "C"
