// Observation render: Env.render (env.py:120-130) = LocalView (engine.py:165-218) + ItemView
// (engine.py:227-248), written straight into the transposed (H, W, 3) uint8 observation.
//
// One CTA renders one environment.  Per env it stages in shared memory
//   * the 9x7 (gx x gy) window: material id + sprite id per cell (the LocalView gather),
//   * four 256-entry FP64 tables that fold the reference's float64 post-processing
//       out = daylight*c + (1-daylight)*(0.5*enh + 0.5*tint)                  (engine.py:189-206)
//     into one add per channel:  out = A[c] + B[k][enh]  -- same operations, same roundings,
//   * the finished observation tile, which leaves the SM as one bulk (TMA) store.
// Arithmetic follows the reference's dtypes: sprite alpha blend in float32 (engine.py:276-284),
// ImageEnhance.Color == PIL blend in float32 on the luma image, everything else float64, all
// casts truncating; -fmad=false keeps products and sums unfused.
#pragma once
#include "cr_common.h"

namespace cr {

struct RenderTables {
  const uint32_t *mat_tex;    // [13][ux*uy]   RGBX texels; id 0 = (127,127,127) (engine.py:168)
  const uint32_t *obj_tex;    // [14][ux*uy]   RGBA texels
  const uint32_t *item_tile;  // [16][10][ux*uy] RGBX: icon + digit composited over black
  const double *vignette;     // [lw][lh]      engine.py:213-218 (numpy on the host)
  const double *daylight;     // [n_daylight]  env.py:135-139   (numpy on the host)
  const uint16_t *colx;       // [sw] obs column -> (cell i << 8 | texel tx), 0xFFFF = border
  const uint16_t *rowy;       // [sh] obs row    -> (cell j << 8 | texel ty), 0xFFFF = border
};

struct RenderShared {      // per-CTA staging
  double A[256];           // daylight * c
  double B[3][256];        // (1 - daylight) * (0.5 * e + 0.5 * tint[k])
  int32_t inv[N_ITEMS];
  uint8_t tmat[256];       // window cells: material id (0 outside the map)
  uint8_t tobj[256];       // window cells: sprite id, 255 = no object
};

CR_DEV int luma(int r, int g, int b) {  // PIL convert('L'), ITU-R 601-2 in 16.16 fixed point
  return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16;
}

// engine.py:276-284 for one channel, float32 end to end.
CR_DEV int blend_f32(int alpha, int tex, int cur) {
  float a = (float)alpha / 255.0f, t = (float)tex / 255.0f, c = (float)cur / 255.0f;
  float blended = a * t + (1.0f - a) * c;
  return (int)(255.0f * blended);
}

// Sprite id of the object in a slot (the `texture` properties of objects.py).
CR_DEV int sprite_of(const Ent &e, int sleeping) {
  switch (e.type) {
    case T_PLAYER: return sleeping ? TEX_PLAYER_SLEEP : TEX_PLAYER_LEFT + e.aux;
    case T_COW: return TEX_COW;
    case T_ZOMBIE: return TEX_ZOMBIE;
    case T_SKELETON: return TEX_SKELETON;
    case T_ARROW: return TEX_ARROW_LEFT + e.aux;
    default: return e.aux > 300 ? TEX_PLANT_RIPE : TEX_PLANT;
  }
}

// Stage the per-env tables.  tid in [0, nthreads).  Caller synchronises afterwards.
CR_DEV void render_stage(const Geom &g, const State &st, const RenderTables &rt, int env, int tid,
                         int nthreads, RenderShared &S, double daylight) {
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  const int px = ps[PS_PX], py = ps[PS_PY], sleeping = ps[PS_SLEEPING];
  const uint8_t *mat = st.mat + (size_t)env * g.NC;
  const uint16_t *objmap = st.objmap + (size_t)env * g.NC;
  const Ent *ents = st.ents + (size_t)env * g.CAP;
  const int offx = g.gx / 2, offy = g.gy / 2;  // engine.py:161
  for (int c = tid; c < g.gx * g.gy; c += nthreads) {  // engine.py:169-181
    int i = c / g.gy, j = c - i * g.gy;
    int wx = px + i - offx, wy = py + j - offy;
    int m = 0, o = 255;
    if (wx >= 0 && wx < g.W && wy >= 0 && wy < g.H) {
      int cell = wx * g.H + wy;
      m = mat[cell] & 0x7F;
      int slot = objmap[cell];
      if (slot) o = sprite_of(ents[slot], sleeping);
    }
    S.tmat[c] = (uint8_t)m;
    S.tobj[c] = (uint8_t)o;
  }
  const double inv_d = 1 - daylight;
  for (int v = tid; v < 256; v += nthreads) {
    S.A[v] = daylight * (double)v;
    double half = (1 - 0.5) * (double)v;  // _tint, engine.py:204-206
    S.B[0][v] = inv_d * (half + 0.5 * 0.0);
    S.B[1][v] = inv_d * (half + 0.5 * 16.0);
    S.B[2][v] = inv_d * (half + 0.5 * 64.0);
  }
  const int32_t *inv = st.inventory + (size_t)env * N_ITEMS;
  for (int i = tid; i < N_ITEMS; i += nthreads) S.inv[i] = inv[i];
}

// One observation pixel -> packed 0x00BBGGRR.  (x, y) index the (H, W, 3) output.
CR_DEV uint32_t render_pixel(const Geom &g, const RenderTables &rt, const RenderShared &S, int x,
                             int y, double daylight, double amount, int sleeping,
                             uint32_t world_seed, uint32_t step) {
  const uint32_t cx = rt.colx[x], ry = rt.rowy[y];
  if (cx == 0xFFFFu || ry == 0xFFFFu) return 0;  // border stays zero, env.py:124
  const int i = cx >> 8, tx = cx & 0xFF, j = ry >> 8, ty = ry & 0xFF;
  const int texel = tx * g.uy + ty, tsize = g.ux * g.uy;
  if (j >= g.gy) {  // item strip, engine.py:227-248: inventory order, 9 per row
    int index = (j - g.gy) * g.vw + i;
    if (index >= N_ITEMS) return 0;
    int amount_i = S.inv[index];
    if (amount_i < 1) return 0;
    if (amount_i > 9) amount_i = 0;  // tile 0 = icon + 'unknown' glyph (engine.py:246)
    return rt.item_tile[(index * 10 + amount_i) * tsize + texel] & 0x00FFFFFFu;
  }
  const int cell = i * g.gy + j;
  uint32_t rgb = rt.mat_tex[S.tmat[cell] * tsize + texel];
  int c0 = rgb & 0xFF, c1 = (rgb >> 8) & 0xFF, c2 = (rgb >> 16) & 0xFF;
  const int o = S.tobj[cell];
  if (o != 255) {
    uint32_t t = rt.obj_tex[o * tsize + texel];
    int a = t >> 24;
    c0 = blend_f32(a, t & 0xFF, c0);
    c1 = blend_f32(a, (t >> 8) & 0xFF, c1);
    c2 = blend_f32(a, (t >> 16) & 0xFF, c2);
  }
  int n0 = c0, n1 = c1, n2 = c2;
  if (daylight < 0.5) {  // _noise, engine.py:208-211; one U(32,127) per pixel, keyed by (step, pixel)
    const uint32_t pix = (uint32_t)((i * g.ux + tx) * g.lh + (j * g.uy + ty));
    U4 w = philox4x32(world_seed, D_NOISE, pix >> 2, step, 0, 0);
    double u = 32.0 + (127.0 - 32.0) * ((double)w.w[pix & 3u] * (1.0 / 4294967296.0));
    double mask = amount * rt.vignette[pix];
    double om = 1 - mask, mu = mask * u;
    n0 = (int)(om * (double)c0 + mu);
    n1 = (int)(om * (double)c1 + mu);
    n2 = (int)(om * (double)c2 + mu);
  }
  // ImageEnhance.Color(night).enhance(0.4) == Image.blend(grey, night, 0.4), float32 per channel
  const int L = luma(n0, n1, n2);
  const float fL = (float)L;
  int e0 = (int)(fL + 0.4f * (float)(n0 - L));
  int e1 = (int)(fL + 0.4f * (float)(n1 - L));
  int e2 = (int)(fL + 0.4f * (float)(n2 - L));
  int r0 = (int)(S.A[c0] + S.B[0][e0]);  // engine.py:196
  int r1 = (int)(S.A[c1] + S.B[1][e1]);
  int r2 = (int)(S.A[c2] + S.B[2][e2]);
  if (sleeping) {  // _sleep, engine.py:198-202: grey of the truncated frame, tint (0,0,16) at 0.5
    int G = luma(r0, r1, r2) >> 1;  // (1-0.5)*G + 0.5*{0,0,16}, truncated
    r0 = G; r1 = G; r2 = G + 8;
  }
  return (uint32_t)r0 | ((uint32_t)r1 << 8) | ((uint32_t)r2 << 16);
}

// Render env into `tile` (sh*sw*3 bytes, shared memory on the device; 4-byte aligned).
// Threads take groups of four consecutive pixels = three aligned 32-bit words.
CR_DEV void render_env(const Geom &g, const State &st, const RenderTables &rt, const RenderShared &S,
                       int env, int tid, int nthreads, uint8_t *tile, double daylight,
                       bool words_ok = true) {
  const int32_t *ps = st.pstate + (size_t)env * PS_COUNT;
  const int sleeping = ps[PS_SLEEPING];
  const uint32_t ws = (uint32_t)ps[PS_WORLD_SEED], step = (uint32_t)ps[PS_STEP];
  const double amount = 2 * (0.5 - daylight);  // engine.py:192
  const int P = g.sw * g.sh, groups = (P + 3) >> 2;
  uint32_t *words = (uint32_t *)tile;
  for (int q = tid; q < groups; q += nthreads) {
    int p = q << 2;
    int y = p / g.sw, x = p - y * g.sw;
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      px[k] = (p + k < P) ? render_pixel(g, rt, S, x, y, daylight, amount, sleeping, ws, step) : 0u;
      if (++x == g.sw) { x = 0; ++y; }
    }
    if (words_ok && p + 3 < P) {
      words[q * 3 + 0] = px[0] | (px[1] << 24);
      words[q * 3 + 1] = (px[1] >> 8) | (px[2] << 16);
      words[q * 3 + 2] = (px[2] >> 16) | (px[3] << 8);
    } else {
      for (int k = 0; p + k < P; ++k) {
        tile[(p + k) * 3 + 0] = (uint8_t)px[k];
        tile[(p + k) * 3 + 1] = (uint8_t)(px[k] >> 8);
        tile[(p + k) * 3 + 2] = (uint8_t)(px[k] >> 16);
      }
    }
  }
}

// info['semantic'] (engine.py:251-264): material ids with 12 + type on object cells.
CR_DEV uint8_t semantic_cell(const Geom &g, const State &st, int env, int cell) {
  int slot = st.objmap[(size_t)env * g.NC + cell];
  if (slot) return (uint8_t)(12 + st.ents[(size_t)env * g.CAP + slot].type);
  return st.mat[(size_t)env * g.NC + cell] & 0x7F;
}

}  // namespace cr
