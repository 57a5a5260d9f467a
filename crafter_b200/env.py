"""Batched Crafter environment on one B200: the `crafter.Env` surface (crafter/env.py:25-130) with a
`num_envs` batch dimension and torch.cuda tensors in and out.

Host code stays Python; all simulation and rendering runs in hand-written sm_100a kernels behind
the C ABI of include/crafter_b200.h.  torch is used for device memory and streams only.
"""
import collections
import ctypes
import os

import numpy as np
import torch

from . import _cabi
from . import rules
from . import state as state_lib
from . import tables as tables_lib

DiscreteSpace = collections.namedtuple('DiscreteSpace', 'n')  # env.py:18-21 (gym is optional)
BoxSpace = collections.namedtuple('BoxSpace', 'low, high, shape, dtype')


class Info(dict):
  """`info` of Env.step (env.py:108-115) as batched tensors; expensive entries are computed on
  first access: 'semantic' (engine.py:251-264), 'discount' (env.py:111) and -- for auto_reset, where
  'inventory' / 'achievements' of an env that just finished already belong to its next episode --
  'final_inventory' / 'final_achievements' / 'final_player_pos' / 'final_observation' / 'final_semantic': the terminal transition as the
  reference's info shows it (rows of envs with done=False hold their last terminal values or zeros)."""

  def __init__(self, env, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._env = env

  def __missing__(self, key):
    env = self._env
    if key == 'semantic':
      value = env.semantic()
    elif key == 'discount':
      # env.py:111: 1 - float(dead).  Taken from the terminal record of the tick (a regenerated env's
      # inventory already shows the health of its next episode): dead is only ever set with done
      dead = env._done & (env._state['final_stats'][:, 23] != 0)
      value = 1.0 - dead.float()
    elif key == 'final_achievements':
      value = env._state['final_stats'][:, :22]
    elif key == 'final_inventory':
      value = env._state['final_stats'][:, 24:40]
    elif key == 'final_player_pos':
      value = env._state['final_stats'][:, 40:42]
    elif key in ('final_observation', 'final_semantic'):
      if env._final_obs is None:
        raise KeyError(f"{key} needs Env(..., auto_reset=True, final_obs=True)")
      value = env._final_obs if key == 'final_observation' else env._final_semantic.view(env.num_envs, *env._area)
    else:
      raise KeyError(key)
    self[key] = value
    return value


class Env:
  """Drop-in for `crafter.Env(area, view, size, reward, length, seed)` (env.py:27-29) plus:

  num_envs      batch size; every tensor gains a leading dimension of this size
  device        a CUDA device (there is no CPU path)
  auto_reset    False (reference semantics: the caller resets, `reset(done)` takes a mask) or True:
                finished episodes are regenerated inside `step()` and the returned observation of
                those envs is the first one of the new episode
  env_offset    global index of env 0, so that a batch sharded over GPUs matches one big batch:
                env i plays the reference's `Env(seed=seed + env_offset + i)`
  final_obs     with auto_reset: also draw the frame of the step that ended an episode (the one the
                reference returns with done=True, env.py:96,118) into info['final_observation']

  Randomness is counter-based (Philox keyed by the per-episode world seed), so a
  batch is reproducible and independent of how it is sharded.  Returned tensors are views of the
  env's output buffers: they are overwritten by the next `step()` / `reset()`; clone to keep them.
  """

  def __init__(self, num_envs=1, area=(64, 64), view=(9, 9), size=(64, 64), reward=True,
               length=10000, seed=None, device=None, auto_reset=False, env_offset=0,
               slot_capacity=None, final_obs=False):
    if not torch.cuda.is_available():
      raise RuntimeError('crafter_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    self._lib = _cabi.load()
    self._device = torch.device('cuda', torch.cuda.current_device()) if device is None else (
        torch.device(device))
    if self._device.type != 'cuda':
      raise ValueError('device must be a CUDA device')
    if self._device.index is None:
      self._device = torch.device('cuda', torch.cuda.current_device())
    geo = tables_lib.geometry(view, size)
    self._num_envs = int(num_envs)
    self._area = (int(area[0]), int(area[1]))
    self._view = geo['view']
    self._size = geo['size']
    self._reward = reward
    self._length = length
    self._seed = int(np.random.randint(0, 2 ** 31 - 1) if seed is None else seed)  # env.py:32
    if not 0 <= self._seed + env_offset + num_envs < 2 ** 61 - 1:
      raise ValueError('seed out of range')
    self._auto_reset = bool(auto_reset)
    if final_obs and not auto_reset:
      raise ValueError('final_obs only makes sense with auto_reset=True (otherwise obs IS the terminal frame)')
    self._want_final_obs = bool(final_obs)
    self._env_offset = int(env_offset)
    self._capacity = int(slot_capacity or state_lib.default_slot_capacity(self._area))
    # env.py:106: length None / 0 = no time limit.  Daylight (env.py:135-139) is a host-built table, so an
    # unbounded env gets a million steps of it (8 MB; a random agent lives 170); running past the table
    # keeps its last entry and raises the ERR_DAYLIGHT_CLAMP bit, which check_errors() reports
    # (a done env may be stepped on without a reset, as the reference allows: 1024 steps of slack)
    self._n_daylight = int(length) + 1026 if length else 1_000_002
    self.reward_range = None  # env.py:55-56
    self.metadata = None
    with torch.cuda.device(self._device):
      self._stream = torch.cuda.Stream(self._device)
      self._alloc_state()
      self._daylight = self._upload(tables_lib.daylight_table(self._n_daylight))
      self._handle, self._tables = self._create(tuple(int(v) for v in self._size))
    self._aux_handles = {}
    self._needs_reset = True
    # raw addresses of the fixed buffers: the per-step calls below hand them to the C ABI without
    # touching torch again (the library switches to its own device itself, see DeviceGuard)
    self._ptrs = (self._actions.data_ptr(), self._obs.data_ptr(), self._reward_buf.data_ptr(),
                  self._done.data_ptr())
    self._stream_ptr = self._stream.cuda_stream
    self._host_key, self._host_ptrs = None, None

  # ---- construction ---------------------------------------------------------------------------
  def _upload(self, array):
    return torch.from_numpy(np.ascontiguousarray(array)).to(self._device)

  def _alloc_state(self):
    B, nc = self._num_envs, self._area[0] * self._area[1]
    nch = -(-self._area[0] // 12) * -(-self._area[1] // 12)
    z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=self._device)
    self._state = dict(
        mat=z(B, nc, dtype=torch.uint8),
        objmap=z(B, nc, dtype=torch.int16),
        ents=z(B, self._capacity, dtype=torch.int64),
        inventory=z(B, len(rules.ITEMS), dtype=torch.int32),
        achievements=z(B, len(rules.ACHIEVEMENTS), dtype=torch.int32),
        pstate=z(B, len(rules.PSTATE), dtype=torch.int32),
        touched=z(B, (nch + 31) // 32, dtype=torch.int32),
        perm=z(B, 512, dtype=torch.uint8),
        next_mat=z(B, nc, dtype=torch.uint8),
        next_ents=z(B, self._capacity, dtype=torch.int64),
        next_meta=z(B, 8, dtype=torch.int32),
        reset_list=z(B, dtype=torch.int32),
        ep_return=z(B, 2, dtype=torch.float64),
        final_stats=z(B, 42, dtype=torch.int32),
        balance_list=z(B, dtype=torch.int32))
    counters = z(4, dtype=torch.int32)  # adjacent, so the step graph clears both with one memset
    self._state['reset_count'] = counters[0:1]
    self._state['balance_count'] = counters[1:2]
    if os.environ.get('CRAFTER_B200_INCR_CENSUS') != '0':
      # grass / path cells per chunk, maintained by the terrain writes instead of being re-counted by
      # every balance tick (=0 goes back to the census for A/B runs)
      self._state['chunk_cnt'] = z(B, nch * 2, dtype=torch.int32)
    self._obs = z(B, int(self._size[1]), int(self._size[0]), 3, dtype=torch.uint8)
    self._final_obs = z(B, int(self._size[1]), int(self._size[0]), 3, dtype=torch.uint8) if self._want_final_obs else None
    self._final_semantic = z(B, nc, dtype=torch.uint8) if self._want_final_obs else None
    self._reward_buf = z(B, dtype=torch.float32)
    self._zero_reward = z(B, dtype=torch.float32)  # reward=False (env.py:116-117); info['reward'] keeps the real one
    self._done = z(B, dtype=torch.bool)
    self._actions = z(B, dtype=torch.int32)

  def _create(self, size):
    t = tables_lib.render_tables(tuple(int(v) for v in self._view), size)
    dev = {k: self._upload(t[k]) for k in ('mat_tex', 'obj_tex', 'item_tile', 'vignette', 'colx',
                                           'rowy')}
    dev['daylight'] = self._daylight
    cfg = _cabi.CrConfig(
        num_envs=self._num_envs, area_w=self._area[0], area_h=self._area[1],
        view_w=int(self._view[0]), view_h=int(self._view[1]), size_w=size[0], size_h=size[1],
        length=int(self._length or 0), reward=int(bool(self._reward)),
        auto_reset=int(self._auto_reset), slot_capacity=self._capacity,
        n_daylight=self._n_daylight, item_w=t['item_size'][0], item_h=t['item_size'][1],
        digit_w=t['digit_size'][0], digit_h=t['digit_size'][1], seed=self._seed,
        env_offset=self._env_offset)
    tabs = _cabi.CrTables(**{k: v.data_ptr() for k, v in dev.items()})
    st = _cabi.CrState(**{k: v.data_ptr() for k, v in self._state.items()})
    if self._final_obs is not None and size == tuple(int(v) for v in self._size):
      st.final_obs = self._final_obs.data_ptr()
      st.final_semantic = self._final_semantic.data_ptr()
    handle = ctypes.c_void_p()
    _cabi.check(self._lib.cr_create(
        ctypes.byref(cfg), ctypes.byref(tabs), ctypes.byref(st), ctypes.byref(handle)))
    return handle, dev

  def close(self):
    for h, _ in list(getattr(self, '_aux_handles', {}).values()) + [
        (getattr(self, '_handle', None), None)]:
      if h:
        self._lib.cr_destroy(h)
    self._aux_handles = {}
    self._handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  # ---- spaces (env.py:58-68) ------------------------------------------------------------------
  @property
  def num_envs(self):
    return self._num_envs

  @property
  def device(self):
    return self._device

  @property
  def observation_space(self):
    return BoxSpace(0, 255, (int(self._size[1]), int(self._size[0]), 3), np.uint8)

  @property
  def action_space(self):
    return DiscreteSpace(len(rules.ACTIONS))

  @property
  def action_names(self):
    return rules.ACTIONS

  # ---- stream plumbing ------------------------------------------------------------------------
  def _enter(self):
    self._stream.wait_stream(torch.cuda.current_stream(self._device))
    return self._stream.cuda_stream

  def _exit(self):
    torch.cuda.current_stream(self._device).wait_stream(self._stream)

  # ---- Env.reset (env.py:70-81) ---------------------------------------------------------------
  def reset(self, mask=None):
    """Start a new episode in every env (or in those where `mask` is True); returns obs."""
    with torch.cuda.device(self._device):
      ptr = None
      if mask is not None:
        mask = torch.as_tensor(mask, device=self._device).to(torch.bool).contiguous()
        assert mask.shape == (self._num_envs,)
        if self._needs_reset and not bool(mask.all()):
          raise RuntimeError('the first reset() must cover every env (envs outside the mask have no world yet)')
        ptr = mask.data_ptr()
      s = self._enter()
      _cabi.check(self._lib.cr_reset(self._handle, ptr, self._obs.data_ptr(), s))
      self._exit()
    self._needs_reset = False
    return self._obs

  # ---- Env.step (env.py:83-118) ---------------------------------------------------------------
  def step(self, actions):
    """actions: int tensor / array of shape (num_envs,) -> (obs, reward, done, info)."""
    if self._needs_reset:
      raise RuntimeError('call reset() before step()')  # the reference fails on None state too
    if not (torch.is_tensor(actions) and actions.data_ptr() == self._ptrs[0]):
      a = torch.as_tensor(actions)
      self._actions.copy_(a.reshape(self._num_envs), non_blocking=True)
    s = self._enter()
    _cabi.check(self._lib.cr_step(self._handle, *self._ptrs, s))
    self._exit()
    info = Info(
        self, inventory=self._state['inventory'], achievements=self._state['achievements'],
        player_pos=self._state['pstate'][:, 12:14], reward=self._reward_buf)
    return self._obs, self._reward_buf if self._reward else self._zero_reward, self._done, info

  @property
  def actions_buffer(self):
    """Write actions here and pass this very tensor to step() to skip the copy."""
    return self._actions

  def step_host(self, actions_pinned, reward_pinned, done_pinned, obs_pinned=None):
    """One tick through `cr_step_host`: pinned host buffers in and out, copies and the stream
    synchronisation included -- the path a non-torch caller of the reference's step() binds."""
    if self._needs_reset:
      raise RuntimeError('call reset() before step()')
    key = (id(actions_pinned), id(reward_pinned), id(done_pinned), id(obs_pinned))
    if key != self._host_key:  # same buffers every step in a rollout loop: look the addresses up once
      self._host_ptrs = (actions_pinned.data_ptr(), obs_pinned.data_ptr() if obs_pinned is not None else None,
                         reward_pinned.data_ptr(), done_pinned.data_ptr())
      self._host_key, self._host_keep = key, (actions_pinned, reward_pinned, done_pinned, obs_pinned)
    _cabi.check(self._lib.cr_step_host(self._handle, *self._host_ptrs, *self._ptrs, self._stream_ptr))

  # ---- Env.render (env.py:120-130) ------------------------------------------------------------
  def render(self, size=None, env_ids=None):
    """Fresh (num_envs, H, W, 3) uint8 render, at `size` if given (e.g. 512 for videos); with
    `env_ids` only those envs are drawn, in that order: (len(env_ids), H, W, 3)."""
    with torch.cuda.device(self._device):
      if size is None:
        handle, sz = self._handle, tuple(int(v) for v in self._size)
      else:
        sz = tuple(size) if hasattr(size, '__len__') else (int(size), int(size))
        if sz not in self._aux_handles:
          self._aux_handles[sz] = self._create(sz)
        handle = self._aux_handles[sz][0]
      if env_ids is None:
        out = torch.empty(self._num_envs, sz[1], sz[0], 3, dtype=torch.uint8, device=self._device)
        s = self._enter()
        _cabi.check(self._lib.cr_render(handle, out.data_ptr(), s))
      else:
        ids = torch.as_tensor(env_ids, dtype=torch.int64).reshape(-1)
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self._num_envs):
          raise IndexError('env_ids out of range')
        ids = ids.to(device=self._device, dtype=torch.int32)
        out = torch.empty(ids.numel(), sz[1], sz[0], 3, dtype=torch.uint8, device=self._device)
        s = self._enter()
        _cabi.check(self._lib.cr_render_envs(handle, ids.data_ptr(), ids.numel(), out.data_ptr(), s))
      self._exit()
    return out

  def semantic(self):
    """info['semantic'] (engine.py:251-264): (num_envs, W, H) uint8."""
    with torch.cuda.device(self._device):
      out = torch.empty(self._num_envs, *self._area, dtype=torch.uint8, device=self._device)
      s = self._enter()
      _cabi.check(self._lib.cr_semantic(self._handle, out.data_ptr(), s))
      self._exit()
    return out

  # ---- inspection -----------------------------------------------------------------------------
  @property
  def state(self):
    """The raw SoA state tensors (zero copy); layout in csrc/cr_common.h."""
    return self._state

  @property
  def launch_count(self):
    return int(self._lib.cr_launch_count(self._handle))

  def error_flags(self):
    """OR of the envs' sticky error bits (synchronises): 1 = an object was dropped because the slot
    arena was full (raise slot_capacity), 2 = an env was stepped past its daylight table."""
    flags = ctypes.c_int32(0)
    s = self._enter()
    _cabi.check(self._lib.cr_error_flags(self._handle, ctypes.byref(flags), s))
    self._exit()
    return int(flags.value)

  def check_errors(self):
    """Raise if any env has a sticky error bit set (see error_flags)."""
    flags = self.error_flags()
    if flags & 1:
      bad = self._state['pstate'][:, 14].bitwise_and(1).nonzero().flatten().tolist()
      raise RuntimeError(f'crafter_b200: slot arena overflow in envs {bad[:8]}{"..." if len(bad) > 8 else ""}: an object '
                         f'was dropped (slot_capacity={self._capacity}); results of those envs differ from the reference')
    if flags & 2:
      raise RuntimeError('crafter_b200: an env was stepped past its daylight table '
                         f'({self._n_daylight} entries); daylight is frozen at the last entry there')

  def set_inventory(self, values, env_ids=None):
    """Overwrite inventory entries ({item: amount}), like poking `env._player.inventory` on the
    reference.  Health also resets the two `_last_health` trackers (env.py:77, objects.py:78)."""
    inv, ps = self._state['inventory'], self._state['pstate']
    idx = slice(None) if env_ids is None else torch.as_tensor(env_ids, device=self._device)
    for name, amount in values.items():
      inv[idx, rules.ITEMS.index(name)] = int(amount)
      if name == 'health':
        ps[idx, state_lib.PS['player_last_health']] = int(amount)
        ps[idx, state_lib.PS['env_last_health']] = int(amount)

  def snapshot(self, i):
    """Canonical host copy of env i's state (see state.canonical)."""
    torch.cuda.synchronize(self._device)
    g = lambda k: self._state[k][i].cpu().numpy()
    return state_lib.canonical(g('mat'), g('ents'), g('inventory'), g('achievements'), g('pstate'),
                               g('touched'), self._area)

  def state_dict(self):
    torch.cuda.synchronize(self._device)
    return {k: v.clone() for k, v in self._state.items()}

  def load_state_dict(self, sd):
    if set(sd) != set(self._state):
      raise ValueError(f'state_dict of another layout (keys differ: {sorted(set(sd) ^ set(self._state))})')
    torch.cuda.synchronize(self._device)
    for k, v in self._state.items():
      v.copy_(sd[k])
    self._needs_reset = False

  def recount(self):
    """After writing `state['mat']` directly: refresh what the library keeps incrementally about the
    terrain (the per-chunk grass / path counts; CRAFTER_B200_INCR_CENSUS=0 keeps nothing)."""
    s = self._enter()
    _cabi.check(self._lib.cr_recount(self._handle, s))
    self._exit()
