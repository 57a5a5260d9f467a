// crafter_b200: the sm_100a kernels (device code only; the C ABI and the step graph that launches them
// live in crafter_kernels.cu).  A separate header so that tests/simt can compile the very same kernels
// for the host under a SIMT emulator (fibers, block / warp barriers, warp collectives): CPU-only CI
// then covers the kernels' block-level choreography too, not just the per-lane logic of tests/hostsim.
#pragma once
#include "cr_common.h"
#include "cr_geom.h"
#include "cr_noise.h"
#include "cr_render.h"
#include "cr_update.h"
#include "cr_worldgen.h"

#ifdef CR_SIMT
#define CR_DYN_SMEM(name) unsigned char *name = simt::dyn_smem()
#else
#define CR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace cr {
namespace kernels {

#ifndef CR_UPDATE_WPB
#define CR_UPDATE_WPB 4
#endif
constexpr int UPDATE_WPB = CR_UPDATE_WPB;  // warps (= envs) per CTA of k_update
constexpr int SEED_WPB = 4;
constexpr int RENDER_THREADS = RENDER_NT;
#ifndef CR_RENDER_MIN_CTAS
#define CR_RENDER_MIN_CTAS (RENDER_NT <= 128 ? 8 : DEF ? 6 : 5)  // default geometry: 40 registers, no spills
// (121.4 vs 123.1 us/step; 7 CTAs spill and are slower); the generic instantiation keeps 48
#endif
#ifndef CR_WG_THREADS
#define CR_WG_THREADS 256
#endif
#ifndef CR_WG_MIN_CTAS
#define CR_WG_MIN_CTAS 3
#endif
constexpr int WG_THREADS = CR_WG_THREADS;
constexpr int OBJ_THREADS = 1024;
constexpr int INSTALL_THREADS = 256;

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// per-warp shared memory of k_update: player copy, mirrored slot records, touched set
__host__ __device__ inline size_t update_smem_per_warp(const Geom &g) {
  return align16(sizeof(PlayerS)) + align16(sizeof(Ent) * ENT_SMEM) + align16(sizeof(uint32_t) * g.TW);
}

// ---- k_update: Env.step minus render (env.py:83-118) ------------------------------------------
template <bool DEF>
__global__ void __launch_bounds__(UPDATE_WPB * 32)
k_update(Geom g, State st, const double *__restrict__ daylight, const int32_t *__restrict__ actions,
         float *reward, uint8_t *done, int auto_reset, int debug_skip) {
  geom_specialize<DEF>(g);
  CR_DYN_SMEM(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * UPDATE_WPB + warp;
  if (env >= g.B) return;
  unsigned char *base = smem + warp * update_smem_per_warp(g);
  PlayerS *P = reinterpret_cast<PlayerS *>(base);
  Ent *sents = reinterpret_cast<Ent *>(base + align16(sizeof(PlayerS)));
  uint32_t *stouched = reinterpret_cast<uint32_t *>(
      base + align16(sizeof(PlayerS)) + align16(sizeof(Ent) * ENT_SMEM));
  int action = actions[env];
  if (action < 0 || action >= N_ACTIONS) action = ACT_NOOP;
  env_step(g, st, daylight, env, lane, action, P, sents, stouched, reward, done, auto_reset,
           debug_skip);
}

// ---- k_balance: spawn / despawn balancing, one CTA per env on a multiple-of-10 step -------------
#ifndef CR_BALANCE_THREADS
#define CR_BALANCE_THREADS 128
#endif
constexpr int BALANCE_THREADS = CR_BALANCE_THREADS;  // default area: 36 chunks, 108 (chunk, class) pairs
constexpr int BALANCE_THREADS_MAX = 512;  // large areas: one thread per few pairs, more loads in flight
__host__ __device__ inline size_t balance_smem(const Geom &g) {
  return align16(sizeof(PlayerS)) + align16((size_t)g.NCH * 5 * sizeof(uint16_t)) +
         align16((size_t)g.NCH * 3 * BAL_MEMBERS * sizeof(uint16_t)) + align16(sizeof(Ent) * ENT_SMEM) + align16(sizeof(uint32_t) * g.TW) +
         align16(sizeof(uint32_t) * g.NCH * 3);
}
// ---- k_post: after the tick, balance the envs on a multiple-of-10 step (env_balance), one CTA
// each; `bal_ctas` CTAs stride over the balance list (a finished env with auto-reset is not on it).
template <bool DEF>
__global__ void __launch_bounds__(BALANCE_THREADS_MAX)
k_post(Geom g, State st, const double *__restrict__ daylight, int bal_ctas) {
  geom_specialize<DEF>(g);
  CR_DYN_SMEM(smem);
  unsigned char *q = smem;
  PlayerS *P = reinterpret_cast<PlayerS *>(q); q += align16(sizeof(PlayerS));
  uint16_t *cnt = reinterpret_cast<uint16_t *>(q); q += align16((size_t)g.NCH * 5 * sizeof(uint16_t));
  uint16_t *members = reinterpret_cast<uint16_t *>(q);
  q += align16((size_t)g.NCH * 3 * BAL_MEMBERS * sizeof(uint16_t));
  Ent *sents = reinterpret_cast<Ent *>(q); q += align16(sizeof(Ent) * ENT_SMEM);
  uint32_t *stouched = reinterpret_cast<uint32_t *>(q); q += align16(sizeof(uint32_t) * g.TW);
  uint32_t *dec = reinterpret_cast<uint32_t *>(q);
  const int count = *st.balance_count;
  for (int r = blockIdx.x; r < count; r += bal_ctas)
    env_balance(g, st, daylight, st.balance_list[r], threadIdx.x, DEF ? BALANCE_THREADS : (int)blockDim.x, P, cnt, members, sents,
                stouched, dec);
}

// ---- reset list ---------------------------------------------------------------------------------
__global__ void k_fill_list(int B, const uint8_t *__restrict__ mask, int32_t *list, int32_t *count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  if (mask == nullptr) {
    list[i] = i;
    if (i == 0) *count = B;
  } else if (mask[i]) {
    list[atomicAdd(count, 1)] = i;
  }
}

// `only_invalid`: skip listed envs whose prefetched world is still valid (explicit reset path).
// List entries are env indices; the deferred mode adds the buffer to fill and a skip flag.
__device__ __forceinline__ bool wg_skip(const State &st, int32_t entry, int only_invalid) {
  if (entry & ENTRY_SKIP) return true;
  return only_invalid && st.next_meta[(size_t)(entry & ENTRY_ENV) * NM_COUNT + NM_VALID] != 0;
}
__device__ __forceinline__ int entry_buf(int32_t entry) { return (entry & ENTRY_BUF) ? 1 : 0; }

// ---- k_seed: one warp per listed world (see wg_seed for `ahead`) -------------------------------
__global__ void __launch_bounds__(SEED_WPB * 32) k_seed(Geom g, State st, int only_invalid, int ahead) {
  __shared__ SeedScratch scratch[SEED_WPB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = *st.reset_count;
  for (int r = blockIdx.x * SEED_WPB + warp; r < count; r += gridDim.x * SEED_WPB) {
    const int env = st.reset_list[r];
    if (!wg_skip(st, env, only_invalid)) wg_seed(g, st, env, lane, scratch[warp], ahead);
    __syncwarp();
  }
}

// Deferred mode: head (ahead = 0) and tail (ahead = 1) seeds of a regeneration pass (wg2_seed_*).
__global__ void __launch_bounds__(SEED_WPB * 32) k_seed2(Geom g, State st, int ahead) {
  __shared__ SeedScratch scratch[SEED_WPB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = *st.reset_count;
  for (int r = blockIdx.x * SEED_WPB + warp; r < count; r += gridDim.x * SEED_WPB) {
    const int32_t e = st.reset_list[r];
    if (!(e & ENTRY_SKIP)) {
      if (ahead) wg2_seed_ahead(g, st, e & ENTRY_ENV, entry_buf(e), lane, scratch[warp]);
      else wg2_seed_head(g, st, e & ENTRY_ENV, entry_buf(e), lane, scratch[warp]);
    }
    __syncwarp();
  }
}

// Deferred mode, explicit reset path: which buffers of the listed envs still need a world.
__global__ void k_prep(Geom g, State st, int which) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= *st.reset_count) return;
  st.reset_list[r] = wg2_prepare(st, st.reset_list[r] & ENTRY_ENV, which);
}

// Deferred mode, tail of the step: the entries k_install rewrote become the pending list.
__global__ void k_pending_copy(State st) {
  const int n = *st.reset_count;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    st.pend_list[r] = st.reset_list[r];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st.pend_count[0] = n;
    st.pend_count[1] += 1;  // step epoch of the two-launch fused schedule (k_tick_render)
  }
}

// ---- k_wg_mat: terrain, a tile of WG_TILE cells per CTA iteration, persistent over (world, tile) --
constexpr int WG_CELLS = WG_TILE;
template <bool DEF>
__global__ void __launch_bounds__(WG_THREADS, CR_WG_MIN_CTAS) k_wg_mat(Geom g, State st, int only_invalid) {
  geom_specialize<DEF>(g);
  __shared__ uint8_t s_perm[256], s_pgi[256];
  __shared__ int8_t s_grad[72];
  __shared__ uint64_t s_ext[N_EXT_CASES];
  __shared__ WgTile T;
  const int tid = threadIdx.x;
  const int count = *st.reset_count;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  const int total = count * tiles;
  for (int i = tid; i < 72; i += WG_THREADS) s_grad[i] = noise_gradient_component(i);
  for (int i = tid; i < N_EXT_CASES; i += WG_THREADS) s_ext[i] = noise_ext_case(i);
  NoiseTables t;
  t.perm = s_perm; t.pgi = s_pgi; t.grad = s_grad; t.ext = s_ext;
  int cur = -1;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int r = w / tiles, tile = w - r * tiles;
    const int32_t entry = st.reset_list[r];
    const int env = entry & ENTRY_ENV, buf = entry_buf(entry);
    if (wg_skip(st, entry, only_invalid)) continue;  // uniform per CTA
    if (env != cur) {
      __syncthreads();
      for (int i = tid; i < 256; i += WG_THREADS) {
        const uint8_t p = st.perm[(size_t)env * 256 + i];
        s_perm[i] = p;
        s_pgi[i] = (uint8_t)((p % 24) * 3);
      }
      cur = env;
      __syncthreads();
    }
    const uint32_t ws = (uint32_t)next_meta_of(st, env, buf)[NM_WORLD_SEED];
    const int cell0 = tile * WG_CELLS;
    wg_material_tile(g, t, ws, next_mat_of(st, g, env, buf), cell0, imin(WG_CELLS, g.NC - cell0),
                     tid, WG_THREADS, T);
  }
}

// ---- k_wg_obj: initial creatures -> slots in x-major cell order (worldgen.py:16-18) -----------
template <bool DEF>
__global__ void __launch_bounds__(OBJ_THREADS) k_wg_obj(Geom g, State st, int only_invalid) {
  geom_specialize<DEF>(g);
  __shared__ int s_warp[OBJ_THREADS / 32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int count = *st.reset_count;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    const int32_t entry = st.reset_list[r];
    const int env = entry & ENTRY_ENV, buf = entry_buf(entry);
    if (wg_skip(st, entry, only_invalid)) continue;  // uniform per CTA
    uint8_t *mat = next_mat_of(st, g, env, buf);
    Ent *ents = next_ents_of(st, g, env, buf);
    int32_t *nm = next_meta_of(st, env, buf);
    const int cpt = (g.NC + OBJ_THREADS - 1) / OBJ_THREADS;
    const int c0 = imin(g.NC, tid * cpt), c1 = imin(g.NC, c0 + cpt);
    // the per-cell creature decisions were made by k_wg_mat (bits 4-5); count, scan, emit in order
    const bool words = (cpt & 3) == 0 && (g.NC & 3) == 0;  // whole aligned words per thread
    int mine = 0;
    if (words) {
      const uint32_t *mw = reinterpret_cast<const uint32_t *>(mat);
      for (int c = c0; c < c1; c += 4) {
        const uint32_t w = mw[c >> 2];
        mine += __popc(((w >> OBJ_SHIFT) | (w >> (OBJ_SHIFT + 1))) & 0x01010101u);
      }
    } else {
      for (int c = c0; c < c1; ++c) mine += ((mat[c] >> OBJ_SHIFT) & 3) != 0;
    }
    // block-wide exclusive prefix sum of `mine`
    int incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int i = 0; i < OBJ_THREADS / 32; ++i) { int v = s_warp[i]; s_warp[i] = run; run += v; }
      s_total = run;
    }
    __syncthreads();
    int slot = 2 + s_warp[warp] + incl - mine;  // slot 1 is the player (env.py:76-78)
    if (words) {
      uint32_t *mw = reinterpret_cast<uint32_t *>(mat);
      for (int c = c0; c < c1; c += 4) {
        const uint32_t w = mw[c >> 2];
        if ((w & 0xF0F0F0F0u) == 0) continue;
        mw[c >> 2] = w & 0x0F0F0F0Fu;
        for (int q = 0; q < 4; ++q) {
          const int type = (w >> (8 * q + OBJ_SHIFT)) & 3;  // 0 none, 1 cow, 2 zombie, 3 skeleton
          if (!type) continue;
          const int cc = c + q, x = cc / g.H;
          if (slot < g.CAP) ents[slot] = wg_make_entity(type + 1, x, cc - x * g.H);
          ++slot;
        }
      }
    } else {
      for (int c = c0; c < c1; ++c) {
        const uint8_t m = mat[c];
        const int type = (m >> OBJ_SHIFT) & 3;
        if (m & ~MAT_MASK) mat[c] = m & MAT_MASK;
        if (type) {
          const int x = c / g.H;
          if (slot < g.CAP) ents[slot] = wg_make_entity(type + 1, x, c - x * g.H);
          ++slot;
        }
      }
    }
    if (tid == 0) {
      int n = 2 + s_total, valid = 1;
      if (n > g.CAP) {
        n = g.CAP;
        // deferred mode: the tick may be rewriting this env's scalars right now; the flag rides in
        // the buffer's row (bit 1 of NM_VALID) and lands in PS_ERROR when the world is installed
        if (g.defer) valid |= 2;
        else st.pstate[(size_t)env * PS_COUNT + PS_ERROR] |= ERR_SLOT_OVERFLOW;
      }
      nm[NM_NSLOTS] = n;
      nm[NM_VALID] = valid;
    }
    __syncthreads();
  }
}

// ---- k_install: prefetched world -> live state for the listed envs (one CTA each) --------------
template <bool DEF>
__global__ void __launch_bounds__(INSTALL_THREADS) k_install(Geom g, State st) {
  geom_specialize<DEF>(g);
  const int count = *st.reset_count;
  for (int r = blockIdx.x; r < count; r += gridDim.x) {
    const int env = st.reset_list[r] & ENTRY_ENV;
    // deferred mode: consume the buffer whose turn it is and name it in the entry, which becomes
    // next step's order to refill it (every thread reads CUR before thread 0 flips it)
    const int c = g.defer ? (st.next_meta2[(size_t)env * NM_COUNT + NM2_CUR] & 1) : 0;
    wg_install_clear(g, st, env, threadIdx.x, INSTALL_THREADS, c);
    __syncthreads();
    if (g.incr_census)  // the fresh terrain's grass / path cells per chunk (block syncs inside)
      census_recount(g, st.mat + (size_t)env * g.NC, st.chunk_cnt + (size_t)env * g.NCH * 2, threadIdx.x,
                     INSTALL_THREADS);
    wg_install_scatter(g, st, env, threadIdx.x, INSTALL_THREADS, c);
    if (threadIdx.x == 0) {
      if (g.defer) {
        wg2_install_player(g, st, env, c);
        st.reset_list[r] = env | (c ? ENTRY_BUF : 0);
      } else {
        wg_install_player(g, st, env);
      }
    }
    __syncthreads();
  }
}

// The finished tile (shared memory) -> the observation row of the env (global memory).  16-byte
// multiples leave the SM as ONE bulk copy (TMA, `UBLKCP` in SASS): generic-proxy writes are fenced
// to the async proxy, then one thread issues, commits and waits for the copy.
__device__ __forceinline__ void store_tile(uint8_t *out, uint8_t *tile, size_t bytes, int tid) {
#ifndef CR_SIMT
  if ((bytes & 15) == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(out), "r"(saddr), "r"((uint32_t)bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
#endif
  __syncthreads();
  for (size_t i = tid; i < bytes; i += RENDER_THREADS) out[i] = tile[i];
}

// CRAFTER_B200_SPLIT=1 (experiment): the step draws in two launches -- RENDER_EARLY right after
// k_update for the envs whose tick is already final, RENDER_LATE for the ones k_post balances or
// k_install regenerates.  RENDER_ALL is the product instantiation and carries no predicate.
enum RenderPart : int { RENDER_ALL = 0, RENDER_EARLY = 1, RENDER_LATE = 2, RENDER_RESET = 3 };

// cr_recount: the incremental census of every env from its terrain (the caller wrote `mat` itself)
__global__ void __launch_bounds__(INSTALL_THREADS) k_recount(Geom g, State st) {
  for (int env = blockIdx.x; env < g.B; env += gridDim.x) {
    __syncthreads();
    census_recount(g, st.mat + (size_t)env * g.NC, st.chunk_cnt + (size_t)env * g.NCH * 2, threadIdx.x,
                   INSTALL_THREADS);
  }
}

// ---- k_render: one CTA per env; tile staged in shared memory, one bulk (TMA) store out --------
template <bool DEF, int PART = RENDER_ALL>
__global__ void __launch_bounds__(RENDER_THREADS, CR_RENDER_MIN_CTAS)
k_render(Geom g, State st, RenderTables rt, uint8_t *__restrict__ obs, int staged,
         const int32_t *__restrict__ env_list, const uint8_t *__restrict__ done, int auto_reset) {
  geom_specialize<DEF>(g);
  if (PART == RENDER_RESET) {  // CRAFTER_B200_FUSED: only the envs k_install has just regenerated
    if (!(auto_reset && done[blockIdx.x])) return;
  } else if (PART != RENDER_ALL) {
    // `done` is written by k_update only; PS_STEP of an env that is not re-installed is stable
    const int e = (int)blockIdx.x;
    const bool late = (auto_reset && done[e]) || st.pstate[(size_t)e * PS_COUNT + PS_STEP] % 10 == 0;
    if (late != (PART == RENDER_LATE)) return;  // uniform per CTA
  }
  CR_DYN_SMEM(smem);
  RenderShared &S = *reinterpret_cast<RenderShared *>(smem);
  uint32_t *tiles = reinterpret_cast<uint32_t *>(smem + align16(sizeof(RenderShared)));
  uint8_t *tile = smem + align16(sizeof(RenderShared)) +
                  (g.tile_cache ? align16((size_t)(N_TILES + 1) * g.ux * g.uy * sizeof(uint32_t)) : 16);
  const int tid = threadIdx.x;
  const int env = env_list ? env_list[blockIdx.x] : (int)blockIdx.x;  // cr_render_envs: a subset
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  const double daylight = rt.daylight[imin(ps[PS_STEP], g.n_daylight - 1)];
  const size_t bytes = (size_t)g.sw * g.sh * 3;
  uint8_t *out = obs + (size_t)blockIdx.x * bytes;
  render_stage(g, st, rt, env, tid, RENDER_THREADS, S, daylight);  // warp 0 also plans the tiles
  __syncthreads();
  render_tiles(g, rt, S, tiles, tid, RENDER_THREADS, daylight < 0.5, ps[PS_SLEEPING]);
  __syncthreads();
  if (!DEF && !staged) {  // the default geometry always stages (12 KB tile; cr_create checks)
    render_assemble(g, st, rt, S, tiles, env, tid, RENDER_THREADS, out, daylight, (bytes & 3) == 0);
    return;
  }
  render_assemble(g, st, rt, S, tiles, env, tid, RENDER_THREADS, tile, daylight, true);
  store_tile(out, tile, bytes, tid);
}

// ---- k_tick_render (CRAFTER_B200_FUSED=1, experiment): tick, balance and observation of ONE env
// in one CTA.  Warp 0 ticks (env_step) while warps 1.. build the frame's FP64 tables; a balancing
// env (step % 10 == 0, known before the tick) then runs env_balance on the whole CTA; the frame
// follows at once.  Envs wait for nobody else's tick, so the latency-bound phases of k_update and
// k_post overlap with other envs' rendering on the same SM.  Two launches side by side -- the
// balancing class (10 % of the envs, long CTAs) and the plain class -- keep the long CTAs from
// forming the tail of one big launch (that fusion was measured slower: profiles/README.md).
// Finished envs (auto-reset) only tick here; k_install and k_render<RENDER_RESET> draw them.
// The tick's and the balance's shared-memory scratch aliases the output tile, written last.
enum TickClass : int { TICK_PLAIN = 0, TICK_BALANCE = 1, TICK_ANY = 2 };
#ifndef CR_FUSED_MIN_CTAS
#define CR_FUSED_MIN_CTAS 5
#endif
template <bool DEF, int CLS>
__global__ void __launch_bounds__(RENDER_THREADS, CR_FUSED_MIN_CTAS)
k_tick_render(Geom g, State st, RenderTables rt, const int32_t *__restrict__ actions,
              uint8_t *__restrict__ obs, float *reward, uint8_t *done, int auto_reset) {
  geom_specialize<DEF>(g);
  CR_DYN_SMEM(smem);
  RenderShared &S = *reinterpret_cast<RenderShared *>(smem);
  uint32_t *tiles = reinterpret_cast<uint32_t *>(smem + align16(sizeof(RenderShared)));
  uint8_t *tile = smem + align16(sizeof(RenderShared)) +
                  align16((size_t)(N_TILES + 1) * g.ux * g.uy * sizeof(uint32_t));
  const int tid = threadIdx.x, env = (int)blockIdx.x;
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  int step_next;  // env.py:84; every thread reads it before the tick rewrites the row
  if (CLS == TICK_ANY) {
    step_next = ps[PS_STEP] + 1;
  } else {
    // Two launches run side by side and each env must be ticked by exactly one of them.  The class
    // of an env is a function of its step counter, which the tick itself advances, so the CTA
    // that takes an env first stamps it with the step's epoch (pend_count[1], advanced once per
    // step by k_pending_copy) BEFORE it ticks; the other launch's CTA reads the counter first and
    // the stamp second: if it saw the advanced counter it also sees the stamp (fences below).
    const int32_t stamp = ((volatile const int32_t *)st.pend_count)[1] + 1;
    volatile int32_t *mark = (volatile int32_t *)(st.next_meta2 + (size_t)env * NM_COUNT + NM2_TICK);
    step_next = ((volatile const int32_t *)ps)[PS_STEP] + 1;
    __threadfence();
    if (*mark == stamp) return;                                   // uniform per CTA
    if ((step_next % 10 == 0) != (CLS == TICK_BALANCE)) return;   // uniform per CTA
    __syncthreads();  // every thread has taken its decision before the stamp appears
    if (tid == 0) { *mark = stamp; __threadfence(); }
  }
  const bool balancing = CLS == TICK_ANY ? step_next % 10 == 0 : CLS == TICK_BALANCE;
  const double daylight = rt.daylight[imin(step_next, g.n_daylight - 1)];
  __syncthreads();
  unsigned char *q = tile;  // scratch until render_assemble
  PlayerS *P = reinterpret_cast<PlayerS *>(q); q += align16(sizeof(PlayerS));
  if (tid < 32) {
    Ent *sents = reinterpret_cast<Ent *>(q);
    uint32_t *stouched = reinterpret_cast<uint32_t *>(q + align16(sizeof(Ent) * ENT_SMEM));
    int action = actions[env];
    if (action < 0 || action >= N_ACTIONS) action = ACT_NOOP;
    env_step(g, st, rt.daylight, env, tid, action, P, sents, stouched, reward, done, auto_reset, 0);
  } else {
    render_tables(tid, RENDER_THREADS, S, daylight);
  }
  __syncthreads();
  const bool regen = auto_reset && done[env];  // written by lane 0 before the barrier
  if (balancing && !regen) {
    uint16_t *cnt = reinterpret_cast<uint16_t *>(q); q += align16((size_t)g.NCH * 5 * sizeof(uint16_t));
    uint16_t *members = reinterpret_cast<uint16_t *>(q);
    q += align16((size_t)g.NCH * 3 * BAL_MEMBERS * sizeof(uint16_t));
    Ent *sents = reinterpret_cast<Ent *>(q); q += align16(sizeof(Ent) * ENT_SMEM);
    uint32_t *stouched = reinterpret_cast<uint32_t *>(q); q += align16(sizeof(uint32_t) * g.TW);
    uint32_t *dec = reinterpret_cast<uint32_t *>(q);
    env_balance(g, st, rt.daylight, env, tid, RENDER_THREADS, P, cnt, members, sents, stouched, dec);
  }
  if (regen) return;
  if (tid < 32) render_gather(g, st, rt, env, tid, S);
  __syncthreads();
  render_tiles(g, rt, S, tiles, tid, RENDER_THREADS, daylight < 0.5, ps[PS_SLEEPING]);
  __syncthreads();
  const size_t bytes = (size_t)g.sw * g.sh * 3;
  uint8_t *out = obs + (size_t)env * bytes;
  render_assemble(g, st, rt, S, tiles, env, tid, RENDER_THREADS, tile, daylight, true);
  store_tile(out, tile, bytes, tid);
}

__global__ void k_semantic(Geom g, State st, uint8_t *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)g.B * g.NC) return;
  int env = (int)(i / g.NC), cell = (int)(i - (size_t)env * g.NC);
  out[i] = semantic_cell(g, st, env, cell);
}

}  // namespace kernels
}  // namespace cr
