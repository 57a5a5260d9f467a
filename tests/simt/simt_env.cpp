// TEST INFRASTRUCTURE ONLY -- not a product path, never loaded by crafter_b200/.
//
// The product's CUDA kernels (crafter_b200/csrc/cr_kernels.h, the file nvcc compiles) built for the
// host on top of tests/simt/simt.h and launched in the order of crafter_kernels.cu's step graph
// (any topological order of the graph is a valid execution; the knobs pick the same schedules as
// the library: CRAFTER_B200_DEFER_WG, _SPLIT, _FUSED, _DRAW_PREFETCH, _NO_SPECIALIZE).  Same C
// interface as tests/hostsim so that the Python replay helpers drive either.
//
// Grids are sized for a 3-SM device: every grid-stride loop of the kernels really strides.
#define CR_SIMT 1
#include "simt.h"

#include "../../crafter_b200/csrc/cr_kernels.h"

using namespace cr;
using namespace cr::kernels;

namespace {

constexpr int NUM_SMS = 3;
constexpr size_t MAX_SMEM = 232448;  // 227 KB opt-in shared memory per block (sm_100)

struct Handle {
  Geom g;
  State st;
  RenderTables rt;
  int auto_reset, is_default, defer, split, fused;
  size_t update_smem, balance_smem, render_smem;
  int balance_threads, render_staged;
};

#define LAUNCH2(NAME, DEF, GRID, BLOCK, SMEM, ...)                                              \
  do {                                                                                          \
    if (DEF) simt::launch(#NAME, GRID, BLOCK, SMEM, [&] { NAME<true>(__VA_ARGS__); });          \
    else simt::launch(#NAME, GRID, BLOCK, SMEM, [&] { NAME<false>(__VA_ARGS__); });             \
  } while (0)

int imin_(long long a, long long b) { return (int)(a < b ? a : b); }

State pending_view(const Handle *h) {
  State v = h->st;
  v.reset_list = h->st.pend_list;
  v.reset_count = h->st.pend_count;
  return v;
}

void launch_render(Handle *h, uint8_t *obs, const uint8_t *done, int part) {
  const Geom &g = h->g;
  const int staged = h->render_staged, ar = h->auto_reset;
  State &st = h->st; RenderTables &rt = h->rt;
  const int32_t *none = nullptr;
#define R(DEF, PART) simt::launch("k_render", g.B, RENDER_THREADS, h->render_smem, [&] { k_render<DEF, PART>(g, st, rt, obs, staged, none, done, ar); })
  if (h->is_default) {
    if (part == RENDER_ALL) R(true, RENDER_ALL); else if (part == RENDER_EARLY) R(true, RENDER_EARLY);
    else if (part == RENDER_LATE) R(true, RENDER_LATE); else R(true, RENDER_RESET);
  } else {
    if (part == RENDER_ALL) R(false, RENDER_ALL); else if (part == RENDER_EARLY) R(false, RENDER_EARLY);
    else if (part == RENDER_LATE) R(false, RENDER_LATE); else R(false, RENDER_RESET);
  }
#undef R
}

// launch_worldgen of crafter_kernels.cu (default schedule)
void worldgen(Handle *h, int only_invalid, int ahead, int seeded) {
  const Geom &g = h->g;
  State &st = h->st;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  const int seed_grid = imin_((g.B + SEED_WPB - 1) / SEED_WPB, NUM_SMS * 4);
  if (!seeded) simt::launch("k_seed", seed_grid, SEED_WPB * 32, 0, [&] { k_seed(g, st, only_invalid, 0); });
  const int mat_grid = imin_((long long)g.B * tiles, NUM_SMS * 16);
  LAUNCH2(k_wg_mat, h->is_default, mat_grid, WG_THREADS, 0, g, st, only_invalid);
  if (ahead) simt::launch("k_seed", seed_grid, SEED_WPB * 32, 0, [&] { k_seed(g, st, 0, 1); });
  const int obj_grid = imin_(g.B, NUM_SMS * 2);
  LAUNCH2(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, g, st, only_invalid);
}

// launch_worldgen2 (deferred schedule): one regeneration pass over `stl`'s list
void worldgen2(Handle *h, const State &stl) {
  const Geom &g = h->g;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  const int seed_grid = imin_((g.B + SEED_WPB - 1) / SEED_WPB, NUM_SMS * 4);
  simt::launch("k_seed2", seed_grid, SEED_WPB * 32, 0, [&] { k_seed2(g, stl, 0); });
  const int mat_grid = imin_((long long)g.B * tiles, NUM_SMS * 16);
  LAUNCH2(k_wg_mat, h->is_default, mat_grid, WG_THREADS, 0, g, stl, 0);
  simt::launch("k_seed2", seed_grid, SEED_WPB * 32, 0, [&] { k_seed2(g, stl, 1); });
  const int obj_grid = imin_(g.B, NUM_SMS * 2);
  LAUNCH2(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, g, stl, 0);
}

void install(Handle *h) {
  const int grid = imin_(h->g.B, NUM_SMS * 8);
  LAUNCH2(k_install, h->is_default, grid, INSTALL_THREADS, 0, h->g, h->st);
}

void pending_copy(Handle *h) {
  State &st = h->st;
  simt::launch("k_pending_copy", h->g.B < 16384 ? 1 : 8, 256, 0, [&] { k_pending_copy(st); });
}

void clear_counts(Handle *h) { *h->st.reset_count = 0; *h->st.balance_count = 0; }

}  // namespace

extern "C" {

int hs_create(const cr_config *c, const cr_tables *t, const cr_state *s, Handle **out) {
  Handle *h = new Handle();
  Geom &g = h->g;
  if (geom_from_config(*c, g)) { delete h; return -2; }
  state_from_abi(*s, h->st);
  h->rt.mat_tex = t->mat_tex; h->rt.obj_tex = t->obj_tex; h->rt.item_tile = t->item_tile;
  h->rt.vignette = t->vignette; h->rt.daylight = t->daylight; h->rt.colx = t->colx;
  h->rt.rowy = t->rowy;
  h->auto_reset = c->auto_reset;
  auto on = [](const char *name) { const char *v = getenv(name); return v && v[0] == '1'; };
  h->is_default = geom_is_default(g) && !on("CRAFTER_B200_NO_SPECIALIZE");
  h->defer = on("CRAFTER_B200_DEFER_WG");
  if (h->defer && !state_has_defer_buffers(h->st)) { delete h; return -3; }
  g.defer = h->defer;
  auto off = [](const char *name) { const char *v = getenv(name); return v && v[0] == '0'; };
  g.draw_prefetch = !off("CRAFTER_B200_DRAW_PREFETCH");
  h->split = on("CRAFTER_B200_SPLIT");
  g.incr_census = !off("CRAFTER_B200_INCR_CENSUS") && h->st.chunk_cnt != nullptr;
  h->update_smem = UPDATE_WPB * update_smem_per_warp(g);
  h->balance_smem = balance_smem(g);
  h->balance_threads = g.NCH * 3 > 4 * BALANCE_THREADS ? BALANCE_THREADS_MAX : BALANCE_THREADS;
  const size_t tile = align16((size_t)g.sw * g.sh * 3);
  const size_t fixed = align16(sizeof(RenderShared)) +
                       (g.tile_cache ? align16((size_t)(N_TILES + 1) * g.ux * g.uy * sizeof(uint32_t)) : 16);
  h->render_staged = fixed + tile <= MAX_SMEM / 2;
  h->render_smem = fixed + (h->render_staged ? tile : 0);
  if (!h->render_staged) h->is_default = 0;
  {
    const size_t tick = align16(sizeof(Ent) * ENT_SMEM) + align16(sizeof(uint32_t) * g.TW);
    const size_t bal = h->balance_smem - align16(sizeof(PlayerS));
    const size_t scratch = align16(sizeof(PlayerS)) + (tick > bal ? tick : bal);
    const char *fu = getenv("CRAFTER_B200_FUSED");
    h->fused = fu && (fu[0] == '1' || fu[0] == '2') && h->defer && h->auto_reset && h->render_staged &&
                       g.tile_cache && scratch <= tile ? fu[0] - '0' : 0;
  }
  *out = h;
  return 0;
}
int hs_destroy(Handle *h) { delete h; return 0; }

int hs_render(Handle *h, uint8_t *obs) { launch_render(h, obs, nullptr, RENDER_ALL); return 0; }

// cr_reset / reset_deferred
int hs_reset(Handle *h, const uint8_t *mask, uint8_t *obs) {
  const Geom &g = h->g;
  State &st = h->st;
  const int list_grid = (g.B + 255) / 256;
  if (h->defer) {
    worldgen2(h, pending_view(h));
    *st.pend_count = 0;
    *st.reset_count = 0;
    simt::launch("k_fill_list", list_grid, 256, 0, [&] { k_fill_list(g.B, mask, st.reset_list, st.reset_count); });
    for (int which = 0; which < 2; ++which) {
      simt::launch("k_prep", list_grid, 256, 0, [&] { k_prep(g, st, which); });
      worldgen2(h, st);
    }
    install(h);
    if (obs) launch_render(h, obs, nullptr, RENDER_ALL);
    worldgen2(h, st);
    return 0;
  }
  *st.reset_count = 0;
  simt::launch("k_fill_list", list_grid, 256, 0, [&] { k_fill_list(g.B, mask, st.reset_list, st.reset_count); });
  worldgen(h, 1, 0, 0);
  install(h);
  if (obs) launch_render(h, obs, nullptr, RENDER_ALL);
  worldgen(h, 0, 1, 0);
  return 0;
}

// enqueue_step / enqueue_step_fused
int hs_step(Handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done) {
  const Geom &g = h->g;
  State &st = h->st;
  RenderTables &rt = h->rt;
  const int ar = h->auto_reset;
  const bool defer = h->defer && ar;
  const char *order = getenv("CR_HOSTSIM_DEFER_ORDER");  // the refill is concurrent with the tick
  const bool late = order && order[0] == 'l';
  if (h->fused && ar) {
    if (!late) worldgen2(h, pending_view(h));
    clear_counts(h);
#define T(DEF, CLS) simt::launch("k_tick_render", g.B, RENDER_THREADS, h->render_smem, [&] { k_tick_render<DEF, CLS>(g, st, rt, actions, obs, reward, done, ar); })
    // the two classes run side by side on the device: either order must do
    const bool plain_first = late;
    if (h->fused == 2) { if (h->is_default) T(true, TICK_ANY); else T(false, TICK_ANY); }
    else if (h->is_default) { if (plain_first) { T(true, TICK_PLAIN); T(true, TICK_BALANCE); } else { T(true, TICK_BALANCE); T(true, TICK_PLAIN); } }
    else { if (plain_first) { T(false, TICK_PLAIN); T(false, TICK_BALANCE); } else { T(false, TICK_BALANCE); T(false, TICK_PLAIN); } }
#undef T
    install(h);
    launch_render(h, obs, done, RENDER_RESET);
    if (late) worldgen2(h, pending_view(h));
    pending_copy(h);
    return 0;
  }
  if (defer && !late) worldgen2(h, pending_view(h));
  clear_counts(h);
  const double *daylight = rt.daylight;
  LAUNCH2(k_update, h->is_default, (g.B + UPDATE_WPB - 1) / UPDATE_WPB, UPDATE_WPB * 32, h->update_smem, g, st,
          daylight, actions, reward, done, ar, 0);
  const int bal_ctas = imin_(g.B, NUM_SMS * 4);
  if (!ar) {
    LAUNCH2(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, g, st, daylight, bal_ctas);
    launch_render(h, obs, done, RENDER_ALL);
    return 0;
  }
  if (h->split) launch_render(h, obs, done, RENDER_EARLY);  // concurrent with k_install / k_post on the device
  install(h);
  LAUNCH2(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, g, st, daylight, bal_ctas);
  launch_render(h, obs, done, h->split ? RENDER_LATE : RENDER_ALL);
  if (defer) {
    if (late) worldgen2(h, pending_view(h));
    pending_copy(h);
  } else {
    worldgen(h, 0, 1, 1);
  }
  return 0;
}

int hs_semantic(Handle *h, uint8_t *out) {
  const Geom &g = h->g;
  State &st = h->st;
  const size_t n = (size_t)g.B * g.NC;
  simt::launch("k_semantic", (unsigned)((n + 255) / 256), 256, 0, [&] { k_semantic(g, st, out); });
  return 0;
}

int hs_recount(Handle *h) {
  if (!h->g.incr_census) return 0;
  const Geom &g = h->g;
  State &st = h->st;
  simt::launch("k_recount", imin_(g.B, NUM_SMS * 8), INSTALL_THREADS, 0, [&] { k_recount(g, st); });
  return 0;
}

long hs_simt_blocks() { return simt::rt().blocks; }

}  // extern "C"
