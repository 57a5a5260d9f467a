"""ORACLE: canonical integer state + digests shared by the reference harness, the C oracle and the
CUDA parity tests.  A state is a dict:

  mat      uint8 [W, H]    material ids, 0 = None, 1..12 = data.yaml:20-32 order (engine.py:29-30)
  objs     int32 [n, 6]    live objects in SLOT order (only relative order is semantic):
                           (type, x, y, health, a, b); type 1 Player 2 Cow 3 Zombie 4 Skeleton
                           5 Arrow 6 Plant; a = facing idx (Player, Arrow: 0 left 1 right 2 up
                           3 down) | cooldown (Zombie) | reload (Skeleton) | grown (Plant);
                           b = sleeping (Player)
  player   int64 [49]      inventory[16], achievements[22] (counts), 2*hunger, 2*thirst, fatigue,
                           2*recover, sleeping, facing idx, player._last_health, x, y,
                           env._last_health, unlocked bitmask
  touched  int32 [k]       sorted indices of chunks that ever held an object (engine.py:36)
"""
import zlib

import numpy as np

KEYS = ('mat', 'objs', 'player', 'touched')
DTYPES = dict(mat=np.uint8, objs=np.int32, player=np.int64, touched=np.int32)


def crc(array, dtype):
  return zlib.crc32(np.ascontiguousarray(array, dtype).tobytes())


def digest(state):
  return {k: crc(state[k], DTYPES[k]) for k in KEYS}


def diff(a, b):
  """Human-readable first difference between two states, or None."""
  for k in KEYS:
    x, y = np.asarray(a[k]), np.asarray(b[k])
    if x.shape != y.shape:
      return f'{k}: shape {x.shape} vs {y.shape}\n{x}\n{y}'
    if (x != y).any():
      idx = np.argwhere(x != y)[:8]
      return f'{k}: differs at {idx.tolist()}: {x[tuple(idx.T)]} vs {y[tuple(idx.T)]}'
  return None
