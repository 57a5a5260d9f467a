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

// env.py:72-74 + worldgen.py:11: world seed, simplex seed and permutation table of a world.
//   ahead = 0  the world about to be generated (episode = live episode + 1); a no-op when an
//              earlier ahead pass prepared it and wg_install_player promoted it (NM_SEEDED).
//   ahead = 1  the world after the one just generated; runs next to k_wg_obj, off the chain.
// One warp per world: lane 0 walks the 64-bit LCG, all lanes reduce the states to swap indices
// (64-bit modulo is the expensive part), lane 0 applies the serial shuffle in shared memory.
CR_DEV void wg_seed(const Geom &g, const State &st, int env, int lane, SeedScratch &S, int ahead) {
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  int32_t *nm = st.next_meta + (size_t)env * NM_COUNT;
  uint8_t *perm = st.perm + (size_t)env * 256;
  if (!ahead && nm[NM_SEEDED]) return;  // promoted by wg_install_player (uniform across the warp)
  if (lane == 0) {
    int episode = ahead ? nm[NM_EPISODE] + 1 : ps[PS_EPISODE] + 1;
    uint32_t ws = world_seed_of(g.seed + g.env_offset + env, episode);
    nm[ahead ? NM_AHEAD_EPISODE : NM_EPISODE] = episode;
    nm[ahead ? NM_AHEAD_WORLD_SEED : NM_WORLD_SEED] = (int32_t)ws;
    if (ahead) nm[NM_AHEAD_VALID] = 1; else nm[NM_SEEDED] = 1;
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

// ---- terrain: worldgen.py:21-61, one cell per QUAD of lanes --------------------------------------
// The reference evaluates up to 11 simplex octaves per cell one after another.  They are pure
// functions, so a quad evaluates four of them at once (round 1: start, water x2, mountain 15;
// round 2: mountain 5 and the first candidates of both branches; round 3: tunnels, coal, iron;
// round 4: lava) and exchanges the values by shuffle.  Values that the reference would not have
// computed are simply unused; the uniform draws keep the reference's order and short-circuiting.
CR_DEV bool wg_eval_args(int round, int q, int x, int y, double &ax, double &ay, double &az) {
  const double fx = (double)x, fy = (double)y;  // _simplex: noise3(x / size, y / size, z)
  switch (round * 4 + q) {
    case 0: ax = fx / 3; ay = fy / 3; az = 8; return true;             // start      (x, y, 8, 3)
    case 1: ax = fx / 15; ay = fy / 15; az = 3; return true;           // water      (x, y, 3, 15)
    case 2: ax = fx / 5; ay = fy / 5; az = 3; return true;             // water      (x, y, 3, 5)
    case 3: ax = fx / 15; ay = fy / 15; az = 0; return true;           // mountain   (x, y, 0, 15)
    case 4: ax = fx / 5; ay = fy / 5; az = 0; return true;             // mountain   (x, y, 0, 5)
    case 5: ax = fx / 7; ay = fy / 7; az = 6; return true;             // cave       (x, y, 6, 7)
    case 6: ax = fx / 9; ay = fy / 9; az = 4; return true;             // sand       (x, y, 4, 9)
    case 7: ax = fx / 7; ay = fy / 7; az = 5; return true;             // tree       (x, y, 5, 7)
    case 8: ax = (double)(2 * x) / 3; ay = (fy / 5) / 3; az = 7; return true;   // (2x, y/5, 7, 3)
    case 9: ax = (fx / 5) / 3; ay = (double)(2 * y) / 3; az = 7; return true;   // (x/5, 2y, 7, 3)
    case 10: ax = fx / 8; ay = fy / 8; az = 1; return true;            // coal       (x, y, 1, 8)
    case 11: ax = fx / 6; ay = fy / 6; az = 2; return true;            // iron       (x, y, 2, 6)
    case 12: ax = fx / 5; ay = fy / 5; az = 6; return true;            // lava       (x, y, 6, 5)
    default: ax = ay = az = 0; return false;
  }
}

// v[k] = value computed by lane k of the quad for `round` (0 for lanes without work).
CR_DEV void wg_quad_noise(const NoiseTables &t, int round, int q, int x, int y, double v[4]) {
#ifdef CR_HOSTSIM
  (void)q;
  for (int k = 0; k < 4; ++k) {
    double ax, ay, az;
    v[k] = (round >= 0 && wg_eval_args(round, k, x, y, ax, ay, az)) ? noise3(t, ax, ay, az) : 0.0;
  }
#else
  double ax, ay, az, mine = 0.0;
  if (round >= 0 && wg_eval_args(round, q, x, y, ax, ay, az)) mine = noise3(t, ax, ay, az);
  const int lane0 = (threadIdx.x & 31) & ~3;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = __shfl_sync(0xffffffffu, mine, lane0 + k);
#endif
}

#ifdef CR_HOSTSIM
CR_DEV bool cr_any(bool p) { return p; }
#else
CR_DEV bool cr_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
#endif

// All four lanes of the quad return the material id (TUNNEL_BIT set for tunnel cells).  Every
// lane of the warp must call this (inactive quads pass active = false).
CR_DEV uint8_t wg_material_quad(const Geom &g, const NoiseTables &t, uint32_t world_seed, int x, int y,
                                int q, bool active) {
  const int px = g.W / 2, py = g.H / 2;  // env.py:71
  Rng rng = rng_ctx(world_seed, D_WG_MAT, (uint32_t)(x * g.H + y));
  int round = active ? 0 : -1;  // -1: finished
  int result = M_GRASS;
  double start = 0, water = 0, mountain = 0, m15 = 0;
  while (cr_any(round >= 0)) {
    double v[4];
    wg_quad_noise(t, round, q, x, y, v);
    if (round == 0) {
      int ddx = x - px, ddy = y - py;
      start = 4 - sqrt((double)(ddx * ddx + ddy * ddy));
      start += 2 * v[0];
      start = 1 / (1 + exp(-start));
      if (start > 0.5) { result = M_GRASS; round = -1; continue; }
      water = (0 + 1 * v[1]) + 0.15 * v[2];  // {15: 1, 5: 0.15}, unnormalised
      water = water + 0.1;
      water -= 2 * start;
      m15 = v[3];
      round = 1;
    } else if (round == 1) {
      mountain = (0 + 1 * m15) + 0.3 * v[0];  // {15: 1, 5: 0.3}
      mountain /= (1 + 0.3);
      mountain -= 4 * start + 0.3 * water;
      if (mountain > 0.15) {
        if (v[1] > 0.15 && mountain > 0.3) { result = M_PATH; round = -1; }  // cave
        else round = 2;
      } else {
        if (0.25 < water && water <= 0.35 && v[2] > -0.2) result = M_SAND;
        else if (0.3 < water) result = M_WATER;
        else if (v[3] > 0 && rng_uniform(rng) > 0.8) result = M_TREE;
        else result = M_GRASS;
        round = -1;
      }
    } else if (round == 2) {
      round = -1;
      if (v[0] > 0.4) result = M_PATH | TUNNEL_BIT;        // horizontal tunnel
      else if (v[1] > 0.4) result = M_PATH | TUNNEL_BIT;   // vertical tunnel
      else if (v[2] > 0 && rng_uniform(rng) > 0.85) result = M_COAL;
      else if (v[3] > 0.4 && rng_uniform(rng) > 0.75) result = M_IRON;
      else if (mountain > 0.18 && rng_uniform(rng) > 0.994) result = M_DIAMOND;
      else if (mountain > 0.3) round = 3;                   // lava needs one more octave
      else result = M_STONE;
    } else if (round == 3) {
      result = v[0] > 0.35 ? M_LAVA : M_STONE;
      round = -1;
    }  // round == -1: this quad is finished and only keeps the warp's shuffles converged
  }
  return (uint8_t)result;
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
  // the seed prepared ahead (next to k_wg_obj) becomes the seed of the world to generate next
  nm[NM_SEEDED] = nm[NM_AHEAD_VALID];
  if (nm[NM_AHEAD_VALID]) {
    nm[NM_EPISODE] = nm[NM_AHEAD_EPISODE];
    nm[NM_WORLD_SEED] = nm[NM_AHEAD_WORLD_SEED];
    nm[NM_AHEAD_VALID] = 0;
  }
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
