"""ORACLE harness: runs the UNMODIFIED reference (`/root/reference/crafter`) in this container.

Only usable where the reference is mounted (never on the GPU box); `tools/make_golden.py` uses it
to produce the committed fixtures under `tests/golden/`.  Nothing is edited in the reference: it is
imported under three dependency shims (`oracle/shims`) and canonicalised by
subclassing / module-attribute patching only:

  (i)  RNG: `World.reset` (engine.py:33-39) gets a `KeyedRandom` instead of `RandomState`; the
       draw context is set from hooks around `Env._update_time` (env.py:135), `Env._balance_object`
       (env.py:157), `worldgen.generate_world/_set_material/_set_object` (worldgen.py:10,21,64) and
       `LocalView._noise` (engine.py:208).  See oracle/keyed_rng.py for the contract.
  (ii) chunks: `World.chunks` (engine.py:46-48) returns the ever-touched chunks in sorted key order
       with members in slot order (the reference iterates `set`s of id()-hashed objects, F5).
"""
import pathlib
import sys

import numpy as np

from oracle import keyed_rng as kr

REFERENCE = pathlib.Path('/root/reference')
_SHIMS = pathlib.Path(__file__).resolve().parent / 'shims'

DIRS = ((-1, 0), (1, 0), (0, -1), (0, 1))  # objects.py:33-34
TYPE_IDS = {'Player': 1, 'Cow': 2, 'Zombie': 3, 'Skeleton': 4, 'Arrow': 5, 'Plant': 6}

_loaded = None


def available():
  return (REFERENCE / 'crafter' / 'env.py').exists()


def load():
  """Import the reference package under the shims and install the canonicalisations."""
  global _loaded
  if _loaded is not None:
    return _loaded
  if not available():
    raise RuntimeError('reference not mounted at /root/reference')
  for p in (str(_SHIMS), str(REFERENCE)):
    if p not in sys.path:
      sys.path.insert(0, p)
  import crafter  # noqa: the reference itself
  from crafter import engine, env as env_mod, worldgen, constants, objects

  class KeyedWorld(engine.World):

    def reset(self, seed=None):
      super().reset(seed)
      self.random = kr.KeyedRandom(seed)
      self.rng_step = 0

    @property
    def chunks(self):
      out = {}
      for key in sorted(self._chunks.keys()):
        members = sorted(
            self._chunks[key], key=lambda o: int(self._obj_map[tuple(o.pos)]))
        out[key] = members
      return out

  engine.World = KeyedWorld

  orig_generate = worldgen.generate_world
  orig_mat = worldgen._set_material
  orig_obj = worldgen._set_object

  def generate_world(world, player):
    world.random.set_ctx(kr.D_SEED)
    return orig_generate(world, player)

  def _set_material(world, pos, player, tunnels, simplex):
    world.random.set_ctx(kr.D_WG_MAT, pos[0] * world.area[1] + pos[1])
    return orig_mat(world, pos, player, tunnels, simplex)

  def _set_object(world, pos, player, tunnels):
    world.random.set_ctx(kr.D_WG_OBJ, pos[0] * world.area[1] + pos[1])
    return orig_obj(world, pos, player, tunnels)

  worldgen.generate_world = generate_world
  worldgen._set_material = _set_material
  worldgen._set_object = _set_object

  orig_noise = engine.LocalView._noise

  def _noise(self, canvas, amount, stddev):
    self._world.random.set_ctx(kr.D_NOISE, self._world.rng_step)
    return orig_noise(self, canvas, amount, stddev)

  engine.LocalView._noise = _noise

  cls_index = {objects.Zombie: 0, objects.Skeleton: 1, objects.Cow: 2}

  class RefEnv(env_mod.Env):
    """The reference Env with keyed randomness; logic is inherited, not restated."""

    def _update_time(self):
      super()._update_time()
      self._world.rng_step = self._step
      self._world.random.set_ctx(kr.D_UPDATE, self._step)

    def _balance_object(self, chunk, objs, cls, *args):
      ncy = -(-self._world.area[1] // 12)
      cidx = (chunk[0] // 12) * ncy + chunk[2] // 12
      self._world.random.set_ctx(kr.D_BALANCE, self._step, cidx, cls_index[cls])
      return super()._balance_object(chunk, objs, cls, *args)

  _loaded = dict(
      crafter=crafter, engine=engine, env=env_mod, worldgen=worldgen, constants=constants,
      objects=objects, RefEnv=RefEnv)
  return _loaded


def export_state(env):
  """Canonical integer state of a reference env (see oracle/canon.py for the layout)."""
  mods = load()
  constants = mods['constants']
  w = env._world
  rows = []
  for obj in w._objects:
    if obj is None:
      continue
    name = type(obj).__name__
    a = b = 0
    if name == 'Player':
      a, b = DIRS.index(tuple(obj.facing)), int(obj.sleeping)
    elif name == 'Zombie':
      a = obj.cooldown
    elif name == 'Skeleton':
      a = obj.reload
    elif name == 'Arrow':
      a = DIRS.index(tuple(int(v) for v in obj.facing))
    elif name == 'Plant':
      a = obj.grown
    rows.append([TYPE_IDS[name], int(obj.pos[0]), int(obj.pos[1]), int(obj.health), int(a), b])
  p = env._player
  unlocked = 0
  for i, name in enumerate(constants.achievements):
    if name in env._unlocked:
      unlocked |= 1 << i
  player = (
      [int(p.inventory[k]) for k in constants.items] +
      [int(p.achievements[k]) for k in constants.achievements] +
      [int(round(p._hunger * 2)), int(round(p._thirst * 2)), int(p._fatigue),
       int(round(p._recover * 2)), int(p.sleeping), DIRS.index(tuple(p.facing)),
       int(p._last_health), int(p.pos[0]), int(p.pos[1]), int(env._last_health), unlocked])
  ncy = -(-w.area[1] // 12)
  touched = sorted((k[0] // 12) * ncy + k[2] // 12 for k in w._chunks.keys())
  return dict(
      mat=w._mat_map.copy(), objs=np.array(rows, np.int32).reshape(-1, 6),
      player=np.array(player, np.int64), touched=np.array(touched, np.int32),
      daylight=float(w.daylight))


def make_env(seed, area=(64, 64), view=(9, 9), size=(64, 64), length=10000, reward=True):
  return load()['RefEnv'](area=area, view=view, size=size, length=length, seed=seed, reward=reward)


def boost_inventory(env, boost):
  """Test hook: overwrite inventory entries right after reset (dict name -> amount)."""
  for k, v in boost.items():
    env._player.inventory[k] = int(v)
  env._last_health = env._player.health
  env._player._last_health = env._player.health
