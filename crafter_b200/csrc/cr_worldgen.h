// World generation: Env.reset (env.py:70-81) + worldgen.generate_world (worldgen.py:10-91).
//
// A world depends only on (env seed, episode), so the world of an env's NEXT episode is generated
// ahead of time into `next_*` buffers, off the critical path of the step, and `wg_install_*` swaps
// it in when the episode ends.
//   wg_seed      one warp per world: episode, world seed, simplex seed, permutation table
//   wg_material  one thread per cell: FP64 simplex terrain (pass 1, worldgen.py:21-61)
//   wg_object    one thread per cell: initial creature decision (pass 2, worldgen.py:64-76);
//                the calling kernel turns the per-cell decisions into slots with an ordered
//                prefix sum so that slot order == x-major cell order (worldgen.py:16-18)
//   wg_install_* copy the prefetched world into the live state + Player / Env reset
#pragma once
#include "cr_common.h"
#include "cr_noise.h"

namespace cr {

constexpr uint8_t TUNNEL_BIT = 0x80;  // `tunnels[x, y]` (worldgen.py:12) carried in mat bit 7

#ifdef CR_HOSTSIM
CR_DEV void cr_atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
#else
CR_DEV void cr_atomic_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
#endif

struct SeedScratch {  // per-warp shared memory of the seeding kernel
  uint64_t lcg[256];
  uint16_t r[256];
  uint8_t source[256];
};

// env.py:72-74 + worldgen.py:11 for the env's next episode: world seed, simplex seed, permutation.
// One warp per world: lane 0 walks the 64-bit LCG, all lanes reduce the states to swap indices
// (64-bit modulo is the expensive part), lane 0 applies the serial shuffle in shared memory.
CR_DEV void wg_seed(const Geom &g, const State &st, int env, int lane, SeedScratch &S) {
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  int32_t *nm = st.next_meta + (size_t)env * NM_COUNT;
  uint8_t *perm = st.perm + (size_t)env * 256;
  if (lane == 0) {
    int episode = ps[PS_EPISODE] + 1;
    uint32_t ws = world_seed_of(g.seed + g.env_offset + env, episode);
    nm[NM_EPISODE] = episode;
    nm[NM_WORLD_SEED] = (int32_t)ws;
    Rng r = rng_ctx(ws, D_SEED, 0);
    uint64_t s = (uint64_t)rng_randint(r, 2147483647u);  // worldgen.py:11
    for (int k = 0; k < 3; ++k) s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    for (int i = 255; i >= 0; --i) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      S.lcg[i] = s;
    }
  }
  for (int i = lane; i < 256; i += CR_LANES) S.source[i] = (uint8_t)i;
#ifndef CR_HOSTSIM
  __syncwarp();
#endif
  for (int i = lane; i < 256; i += CR_LANES) {
    // Python: int((seed + 31) % (i + 1)) on an unbounded signed int, floored modulo.
    int64_t n = i + 1;
    int64_t a = (int64_t)S.lcg[i] % n;
    int64_t r = (a + 31 % n) % n;
    if (r < 0) r += n;
    S.r[i] = (uint16_t)r;
  }
#ifndef CR_HOSTSIM
  __syncwarp();
#endif
  if (lane == 0) {
    for (int i = 255; i >= 0; --i) {
      int r = S.r[i];
      perm[i] = S.source[r];
      S.source[r] = S.source[i];
    }
  }
}

CR_DEV double wg_n(const NoiseTables &t, double x, double y, double z, double size) {
  return noise3(t, x / size, y / size, z);  // _simplex with a single octave, worldgen.py:79-91
}

// worldgen.py:21-61.  Returns the material id, with TUNNEL_BIT set for tunnel cells.
CR_DEV uint8_t wg_material(const Geom &g, const NoiseTables &t, uint32_t world_seed, int x, int y) {
  const int px = g.W / 2, py = g.H / 2;  // env.py:71
  Rng rng = rng_ctx(world_seed, D_WG_MAT, (uint32_t)(x * g.H + y));
  const double fx = (double)x, fy = (double)y;
  int ddx = x - px, ddy = y - py;
  double start = 4 - sqrt((double)(ddx * ddx + ddy * ddy));
  start += 2 * wg_n(t, fx, fy, 8, 3);
  start = 1 / (1 + exp(-start));
  if (start > 0.5) return M_GRASS;  // water / mountain are unused on this branch (no draws skipped)
  double water = (0 + 1 * wg_n(t, fx, fy, 3, 15)) + 0.15 * wg_n(t, fx, fy, 3, 5);  // {15:1, 5:0.15}
  water = water + 0.1;
  water -= 2 * start;
  double mountain = (0 + 1 * wg_n(t, fx, fy, 0, 15)) + 0.3 * wg_n(t, fx, fy, 0, 5);  // {15:1, 5:0.3}
  mountain /= (1 + 0.3);
  mountain -= 4 * start + 0.3 * water;
  if (mountain > 0.15) {
    if (wg_n(t, fx, fy, 6, 7) > 0.15 && mountain > 0.3) return M_PATH;  // cave
    if (wg_n(t, (double)(2 * x), fy / 5, 7, 3) > 0.4) return M_PATH | TUNNEL_BIT;  // horizontal
    if (wg_n(t, fx / 5, (double)(2 * y), 7, 3) > 0.4) return M_PATH | TUNNEL_BIT;  // vertical
    if (wg_n(t, fx, fy, 1, 8) > 0 && rng_uniform(rng) > 0.85) return M_COAL;
    if (wg_n(t, fx, fy, 2, 6) > 0.4 && rng_uniform(rng) > 0.75) return M_IRON;
    if (mountain > 0.18 && rng_uniform(rng) > 0.994) return M_DIAMOND;
    if (mountain > 0.3 && wg_n(t, fx, fy, 6, 5) > 0.35) return M_LAVA;
    return M_STONE;
  }
  if (0.25 < water && water <= 0.35 && wg_n(t, fx, fy, 4, 9) > -0.2) return M_SAND;
  if (0.3 < water) return M_WATER;
  if (wg_n(t, fx, fy, 5, 7) > 0 && rng_uniform(rng) > 0.8) return M_TREE;
  return M_GRASS;
}

// worldgen.py:64-76.  `matbyte` still carries TUNNEL_BIT.  Returns EntType or T_NONE.
CR_DEV int wg_object(const Geom &g, uint32_t world_seed, int x, int y, uint8_t matbyte) {
  const int px = g.W / 2, py = g.H / 2;
  const int m = matbyte & 0x7F;
  if (!((WALKABLE >> m) & 1u)) return T_NONE;
  Rng rng = rng_ctx(world_seed, D_WG_OBJ, (uint32_t)(x * g.H + y));
  int ddx = x - px, ddy = y - py;
  double dist = sqrt((double)(ddx * ddx + ddy * ddy));
  if (dist > 3 && m == M_GRASS && rng_uniform(rng) > 0.985) return T_COW;
  if (dist > 10 && rng_uniform(rng) > 0.993) return T_ZOMBIE;
  if (m == M_PATH && (matbyte & TUNNEL_BIT) && rng_uniform(rng) > 0.95) return T_SKELETON;
  return T_NONE;
}

CR_DEV Ent wg_make_entity(int type, int x, int y) {  // objects.py:266-268,284-288,317-321
  Ent e;
  e.type = (uint8_t)type;
  e.health = (int8_t)(type == T_ZOMBIE ? 5 : 3);
  e.x = (int16_t)x; e.y = (int16_t)y; e.aux = 0;
  return e;
}

// ---- install: prefetched world -> live state (World.reset engine.py:33-39 + env.py:70-81) -----
// Phase A (all threads): terrain copy, empty object map, empty touched set.
CR_DEV void wg_install_clear(const Geom &g, const State &st, int env, int tid, int nthreads) {
  uint8_t *mat = st.mat + (size_t)env * g.NC;
  const uint8_t *src = st.next_mat + (size_t)env * g.NC;
  uint16_t *objmap = st.objmap + (size_t)env * g.NC;
  uint32_t *touched = st.touched + (size_t)env * g.TW;
  if ((g.NC & 15) == 0) {  // rows of every env stay 16-byte aligned
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(src);
    uint64_t *d8 = reinterpret_cast<uint64_t *>(mat), *o8 = reinterpret_cast<uint64_t *>(objmap);
    for (int i = tid; i < g.NC / 8; i += nthreads) d8[i] = s8[i];
    for (int i = tid; i < g.NC / 4; i += nthreads) o8[i] = 0;
  } else {
    for (int c = tid; c < g.NC; c += nthreads) { mat[c] = src[c]; objmap[c] = 0; }
  }
  for (int c = tid; c < g.TW; c += nthreads) touched[c] = 0;
}

// Phase B (all threads, after a barrier): creatures into slots 2.., object map, touched chunks.
CR_DEV void wg_install_scatter(const Geom &g, const State &st, int env, int tid, int nthreads) {
  const int n = st.next_meta[(size_t)env * NM_COUNT + NM_NSLOTS];
  const Ent *src = st.next_ents + (size_t)env * g.CAP;
  Ent *ents = st.ents + (size_t)env * g.CAP;
  uint16_t *objmap = st.objmap + (size_t)env * g.NC;
  uint32_t *touched = st.touched + (size_t)env * g.TW;
  for (int s = 2 + tid; s < n; s += nthreads) {
    Ent e = src[s];
    ents[s] = e;
    objmap[e.x * g.H + e.y] = (uint16_t)s;
    int ch = (e.x / CHUNK) * g.ncy + (e.y / CHUNK);
    cr_atomic_or(&touched[ch >> 5], 1u << (ch & 31));
  }
}

// Phase C (one thread): Player + per-episode scalars, env.py:75-79, objects.py:70-82,
// data.yaml:39-55; consumes the prefetched world.
CR_DEV void wg_install_player(const Geom &g, const State &st, int env) {
  int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  int32_t *inv = st.inventory + (size_t)env * N_ITEMS;
  int32_t *ach = st.achievements + (size_t)env * N_ACH;
  int32_t *nm = st.next_meta + (size_t)env * NM_COUNT;
  for (int i = 0; i < N_ITEMS; ++i) inv[i] = i < 4 ? 9 : 0;
  for (int i = 0; i < N_ACH; ++i) ach[i] = 0;
  ps[PS_HUNGER2] = 0; ps[PS_THIRST2] = 0; ps[PS_FATIGUE] = 0; ps[PS_RECOVER2] = 0;
  ps[PS_SLEEPING] = 0; ps[PS_P_LAST_HEALTH] = 9; ps[PS_LAST_HEALTH] = 9; ps[PS_UNLOCKED] = 0;
  ps[PS_NSLOTS] = nm[NM_NSLOTS]; ps[PS_STEP] = 0;
  ps[PS_EPISODE] = nm[NM_EPISODE]; ps[PS_WORLD_SEED] = nm[NM_WORLD_SEED];
  ps[PS_PX] = g.W / 2; ps[PS_PY] = g.H / 2;
  nm[NM_VALID] = 0;
  Ent p;
  p.type = T_PLAYER; p.health = 9; p.x = (int16_t)(g.W / 2); p.y = (int16_t)(g.H / 2);
  p.aux = 3;  // facing (0, 1) = down, objects.py:72
  Ent *ents = st.ents + (size_t)env * g.CAP;
  ents[1] = p;
  ents[0].type = T_NONE;
  st.objmap[(size_t)env * g.NC + (g.W / 2) * g.H + g.H / 2] = 1;  // env.py:76-78
  int ch = ((g.W / 2) / CHUNK) * g.ncy + ((g.H / 2) / CHUNK);
  cr_atomic_or(&st.touched[(size_t)env * g.TW + (ch >> 5)], 1u << (ch & 31));
}

}  // namespace cr
