// TEST INFRASTRUCTURE ONLY -- not a product path, never loaded by crafter_b200/.
//
// The product's CUDA kernels (crafter_b200/csrc/cr_kernels.h, the file nvcc compiles) built for the
// host on top of tests/simt/simt.h and launched in the order of crafter_kernels.cu's step graph
// (any topological order of the graph is a valid execution; the knobs pick the same schedules as
// the library: CRAFTER_B200_QUEUE, _DRAW_PREFETCH, _INCR_CENSUS, _NO_SPECIALIZE).  Same C
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
  int auto_reset, is_default, queue, parity;
  size_t update_smem, balance_smem, render_smem, consume_smem;
  int balance_threads, render_staged;
};

#define LAUNCH2(NAME, DEF, GRID, BLOCK, SMEM, ...)                                              \
  do {                                                                                          \
    if (DEF) simt::launch(#NAME, GRID, BLOCK, SMEM, [&] { NAME<true>(__VA_ARGS__); });          \
    else simt::launch(#NAME, GRID, BLOCK, SMEM, [&] { NAME<false>(__VA_ARGS__); });             \
  } while (0)

int imin_(long long a, long long b) { return (int)(a < b ? a : b); }

void launch_render(Handle *h, uint8_t *obs) {
  const int32_t *none = nullptr;
  // persistent CTAs on the pretend 3-SM device: the row loop really strides
  const int grid = getenv("CR_SIMT_ONE_SHOT") ? h->g.B : imin_(h->g.B, NUM_SMS * 2);
  LAUNCH2(k_render, h->is_default, grid, RENDER_THREADS, h->render_smem, h->g, h->st, h->rt, obs, h->render_staged, none, h->g.B);
}

// launch_worldgen of crafter_kernels.cu
void worldgen(Handle *h, const int32_t *list, const int32_t *count, int only_invalid, int ahead, int seeded) {
  const Geom &g = h->g;
  State &st = h->st;
  const int tiles = (g.NC + WG_CELLS - 1) / WG_CELLS;
  const int seed_grid = imin_((g.B + SEED_WPB - 1) / SEED_WPB, NUM_SMS * 4);
  if (!seeded) simt::launch("k_seed", seed_grid, SEED_WPB * 32, 0, [&] { k_seed(g, st, list, count, only_invalid, 0); });
  const int mat_grid = imin_((long long)g.B * tiles, NUM_SMS * 16);
  LAUNCH2(k_wg_mat, h->is_default, mat_grid, WG_THREADS, 0, g, st, list, count, only_invalid);
  // k_seed ahead and k_wg_obj run side by side on the device: either order must do
  const char *order = getenv("CR_SIMT_WG_ORDER");
  const bool obj_first = order && order[0] == 'o';
  const int obj_grid = imin_(g.B, NUM_SMS * 2);
  if (obj_first) LAUNCH2(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, g, st, list, count, only_invalid);
  if (ahead) simt::launch("k_seed", seed_grid, SEED_WPB * 32, 0, [&] { k_seed(g, st, list, count, 0, 1); });
  if (!obj_first) LAUNCH2(k_wg_obj, h->is_default, obj_grid, OBJ_THREADS, 0, g, st, list, count, only_invalid);
}

void install(Handle *h) {
  const int grid = imin_(h->g.B, NUM_SMS * 8);
  LAUNCH2(k_install, h->is_default, grid, INSTALL_THREADS, 0, h->g, h->st);
}

// drain_pending of crafter_kernels.cu
void drain_pending(Handle *h) {
  if (!h->queue || !h->auto_reset) return;
  const int q = h->parity ^ 1;
  worldgen(h, h->st.wg_list + (size_t)q * h->g.B, h->st.wg_count + q, 0, 1, 1);
  h->st.wg_count[q] = 0;
}

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
  auto off = [](const char *name) { const char *v = getenv(name); return v && v[0] == '0'; };
  h->is_default = geom_is_default(g) && !on("CRAFTER_B200_NO_SPECIALIZE");
  g.draw_prefetch = !off("CRAFTER_B200_DRAW_PREFETCH");
  g.incr_census = !off("CRAFTER_B200_INCR_CENSUS") && h->st.chunk_cnt != nullptr;
  h->update_smem = UPDATE_WPB * update_smem_per_warp(g);
  h->balance_smem = balance_smem(g);
  h->balance_threads = g.NCH * 3 > 4 * BALANCE_THREADS ? BALANCE_THREADS_MAX : BALANCE_THREADS;
  const size_t tile = align16((size_t)g.sw * g.sh * 3);
  const size_t fixed = render_tile_offset(g);
  h->render_staged = fixed + tile <= MAX_SMEM / 2;
  h->render_smem = fixed + (h->render_staged ? tile : 0);
  if (!h->render_staged) h->is_default = 0;
  h->consume_smem = consume_smem(g, h->render_smem);
  const bool have = h->st.work_queue && h->st.sched && h->st.wg_list && h->st.wg_count;
  h->queue = !off("CRAFTER_B200_QUEUE") && have && h->render_staged && g.tile_cache && h->consume_smem <= MAX_SMEM / 2;
  if (h->st.final_obs && !h->queue) { delete h; return -3; }
  h->parity = 0;
  *out = h;
  return 0;
}
int hs_destroy(Handle *h) { delete h; return 0; }

int hs_render(Handle *h, uint8_t *obs) { launch_render(h, obs); return 0; }

// cr_reset
int hs_reset(Handle *h, const uint8_t *mask, uint8_t *obs) {
  const Geom &g = h->g;
  State &st = h->st;
  const int list_grid = (g.B + 255) / 256;
  drain_pending(h);
  *st.reset_count = 0;
  simt::launch("k_fill_list", list_grid, 256, 0, [&] { k_fill_list(g.B, mask, st.reset_list, st.reset_count); });
  worldgen(h, st.reset_list, st.reset_count, 1, 0, 0);
  install(h);
  if (obs) launch_render(h, obs);
  worldgen(h, st.reset_list, st.reset_count, 0, 1, 0);
  return 0;
}

int hs_flush(Handle *h) { drain_pending(h); return 0; }
int hs_schedule(Handle *h) { return h->queue; }

// enqueue_step_queue / enqueue_step_chain
int hs_step(Handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done) {
  const Geom &g = h->g;
  State &st = h->st;
  RenderTables &rt = h->rt;
  const int ar = h->auto_reset;
  if (h->queue) {
    const int p = h->parity;
    h->parity ^= 1;
    // the side branch runs beside the tick on the device; on this one OS thread it has to come first
    // (an install that needs one of its worlds would wait for ever: cr_wait_flags aborts)
    if (ar) {
      worldgen(h, st.wg_list + (size_t)(p ^ 1) * g.B, st.wg_count + (p ^ 1), 0, 1, 1);
      st.wg_count[p ^ 1] = 0;
    }
    LAUNCH2(k_update, h->is_default, (g.B + UPDATE_WPB - 1) / UPDATE_WPB, UPDATE_WPB * 32, h->update_smem, g, st,
            rt.daylight, actions, reward, done, ar, 0, p);
    LAUNCH2(k_consume, h->is_default, getenv("CR_SIMT_ONE_SHOT") ? g.B : imin_(g.B, NUM_SMS * 2), RENDER_THREADS,
            h->consume_smem, g, st, rt, obs);
    for (int i = 0; i < SC_WORDS; ++i) if (st.sched[i] != 0) { fprintf(stderr, "k_consume left sched[%d] = %d\n", i, st.sched[i]); abort(); }
    for (int i = 0; i < g.B; ++i) if (st.work_queue[i] != 0) { fprintf(stderr, "k_consume left work_queue[%d]\n", i); abort(); }
    return 0;
  }
  *st.reset_count = 0; *st.balance_count = 0;
  const double *daylight = rt.daylight;
  LAUNCH2(k_update, h->is_default, (g.B + UPDATE_WPB - 1) / UPDATE_WPB, UPDATE_WPB * 32, h->update_smem, g, st,
          daylight, actions, reward, done, ar, 0, -1);
  const int bal_ctas = imin_(g.B, NUM_SMS * 4);
  if (!ar) {
    LAUNCH2(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, g, st, daylight, bal_ctas);
    launch_render(h, obs);
    return 0;
  }
  install(h);
  LAUNCH2(k_post, h->is_default, bal_ctas, h->balance_threads, h->balance_smem, g, st, daylight, bal_ctas);
  launch_render(h, obs);
  worldgen(h, st.reset_list, st.reset_count, 0, 1, 1);
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
