/*
 * crafter_b200 C ABI -- the drop-in boundary of the batched, B200-native Crafter environment.
 *
 * The reference (danijar/crafter) has no FFI: its boundary is the Python class `crafter.Env`
 * (crafter/env.py:25).  Each entry point below names the reference interface it replaces.  All
 * buffers are owned by the caller (torch tensors on the Python side) and passed as raw device
 * pointers; the library allocates no device memory (it owns a few auxiliary CUDA streams and events
 * per handle for the branches of the step graph, and one 4-byte word for cr_error_flags), starts no threads and is stream-ordered.
 * Every function returns 0 on success and a negative code on error; cr_last_error() describes the
 * last failure of the calling thread.  A handle is bound to the device that was current in cr_create and is not re-entrant;
 * every entry point switches to that device for the duration of the call when another is current.
 */
#ifndef CRAFTER_B200_H_
#define CRAFTER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CR_ABI_VERSION 3

typedef struct cr_handle cr_handle;

/* Constructor arguments of `Env.__init__` (env.py:27-56) plus the batch dimension. */
typedef struct cr_config {
  int32_t num_envs;      /* new: batch size on this device */
  int32_t area_w, area_h; /* area=(64, 64) */
  int32_t view_w, view_h; /* view=(9, 9) */
  int32_t size_w, size_h; /* size=(64, 64): observation is [size_h][size_w][3] uint8 */
  int32_t length;        /* length=10000, 0 = unbounded (env.py:106) */
  int32_t reward;        /* reward=True (env.py:116-117) */
  int32_t auto_reset;    /* new: regenerate finished episodes inside cr_step */
  int32_t slot_capacity; /* entity slots per env (slot 0 unused, slot 1 = player) */
  int32_t n_daylight;    /* entries of tables.daylight */
  int32_t item_w, item_h, digit_w, digit_h; /* int(0.8*unit), int(0.6*unit) (engine.py:240,247) */
  int64_t seed;          /* env i uses seed + env_offset + i where the reference uses `seed` */
  int64_t env_offset;    /* global index of env 0 of this handle (batch sharding across GPUs) */
} cr_config;

/* Device pointers to read-only tables built once by the host (Textures, engine.py:120-142;
 * _vignette, engine.py:213-218; _update_time, env.py:135-139). */
typedef struct cr_tables {
  const uint32_t *mat_tex;   /* [13][ux*uy] RGBX, texel index tx*uy+ty, id 0 = grey 127 */
  const uint32_t *obj_tex;   /* [14][ux*uy] RGBA */
  const uint32_t *item_tile; /* [16][10][ux*uy] RGBX */
  const double *vignette;    /* [gy*uy][gx*ux] (canvas row major) */
  const double *daylight;    /* [n_daylight] */
  const uint16_t *colx;      /* [size_w] */
  const uint16_t *rowy;      /* [size_h] */
} cr_tables;

/* Device pointers to the mutable SoA state (engine.World + objects.Player). */
typedef struct cr_state {
  uint8_t *mat;           /* [B][W*H] */
  uint16_t *objmap;       /* [B][W*H] */
  void *ents;             /* [B][slot_capacity] 8-byte records */
  int32_t *inventory;     /* [B][16]  info['inventory'] */
  int32_t *achievements;  /* [B][22]  info['achievements'] */
  int32_t *pstate;        /* [B][16]  see cr_common.h PState */
  uint32_t *touched;      /* [B][ceil(chunks/32)] */
  uint8_t *perm;          /* [B][2][256] */
  uint8_t *next_mat;      /* [B][W*H]  prefetched world of the next episode */
  void *next_ents;        /* [B][slot_capacity] */
  int32_t *next_meta;     /* [B][8] */
  int32_t *reset_list;    /* [B] */
  int32_t *reset_count;   /* [1] */
  double *ep_return;      /* [B][2]  StatsRecorder: running / last finished episode return */
  int32_t *final_stats;   /* [B][42] the terminal transition of the last finished episode as the reference's info
                           * shows it (env.py:108-115): achievements[22], length, dead flag, inventory[16], player_pos[2] */
  int32_t *balance_list;  /* [B] */
  int32_t *balance_count; /* [1] */
  /* Grass / path cells per 12x12 chunk, kept current by the library (NULL, or CRAFTER_B200_INCR_CENSUS=0:
   * every balance tick re-counts the cells instead).  Call cr_recount after writing `mat` yourself. */
  int32_t *chunk_cnt;     /* [B][chunks][2] */
  /* Optional (NULL: off), auto_reset only: the frame of the step that ended an episode, which the
   * reference returns with done=True (env.py:96,118), for the envs regenerated inside cr_step;
   * rows of other envs are left alone.  [B][size_h][size_w][3] */
  uint8_t *final_obs;
  uint8_t *final_semantic; /* optional with final_obs: the terminal info['semantic'] of those envs, [B][W][H] */
} cr_state;

int cr_abi_version(void);
/* Digest of the sources the library was compiled from (crafter_b200/build.py source_hash). */
const char *cr_source_hash(void);
const char *cr_last_error(void);

/* Env.__init__ (env.py:27-56). */
int cr_create(const cr_config *cfg, const cr_tables *tables, const cr_state *state, cr_handle **out);
int cr_destroy(cr_handle *h);

/* Env.reset (env.py:70-81) for the envs whose mask byte is non-zero (mask == NULL: all).
 * Writes the first observation of the reset envs into obs[B][size_h][size_w][3]. */
int cr_reset(cr_handle *h, const uint8_t *mask, uint8_t *obs, void *stream);

/* Env.step (env.py:83-118): actions int32[B] in, obs / reward float32[B] / done uint8[B] out.
 * The per-env info tensors are the cr_state buffers themselves (zero copy). */
int cr_step(cr_handle *h, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done,
            void *stream);

/* Same tick with HOST buffers: copies actions in and reward/done (and obs when non-NULL) out and
 * synchronises the stream -- what a non-torch caller of the reference's step() would bind. */
int cr_step_host(cr_handle *h, const int32_t *actions_host, uint8_t *obs_host, float *reward_host,
                 uint8_t *done_host, int32_t *actions_dev, uint8_t *obs_dev, float *reward_dev,
                 uint8_t *done_dev, void *stream);

/* Env.render (env.py:120-130) at the configured size into obs[B][size_h][size_w][3]. */
int cr_render(cr_handle *h, uint8_t *obs, void *stream);

/* The same render for a subset: obs[n][size_h][size_w][3], row r shows env env_ids[r] (device
 * array of n indices in [0, B)).  What a VideoRecorder (recorder.py:68-96) needs of a large batch:
 * 512x512 frames of a few envs, not of all of them. */
int cr_render_envs(cr_handle *h, const int32_t *env_ids, int n, uint8_t *obs, void *stream);

/* SemanticView (engine.py:251-264): out[B][W][H] uint8, info['semantic']. */
int cr_semantic(cr_handle *h, uint8_t *out, void *stream);

/* After the caller has written `mat` itself (state restore, tests): recount what the library keeps
 * incrementally about the terrain (the per-chunk counts of chunk_cnt; a no-op without that buffer
 * or with CRAFTER_B200_INCR_CENSUS=0). */
int cr_recount(cr_handle *h, void *stream);

/* OR of the envs' sticky error bits (pstate column 14) into *flags_host, synchronising the stream:
 * bit 0 an object did not fit the slot arena and was dropped (raise slot_capacity), bit 1 an env's step
 * counter ran past the daylight table (n_daylight entries; the last one is used from there on). */
int cr_error_flags(cr_handle *h, int32_t *flags_host, void *stream);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t cr_launch_count(const cr_handle *h);

/* Profiling aid: with CRAFTER_B200_TIMING=1 in the environment the step runs eagerly with events
 * around every kernel, with =2 it stays one graph and the events are nodes of it; writes the mean
 * device ms of [update, install, render, seed, wg_mat, wg_obj,
 * seed_ahead, balance] since the last call and returns the number of steps averaged (0 = off). */
int64_t cr_timing(cr_handle *h, double *out_ms);

#ifdef __cplusplus
}
#endif
#endif /* CRAFTER_B200_H_ */
