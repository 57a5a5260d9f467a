"""Views of the SoA state tensors (layout: csrc/cr_common.h) for inspection, tests and recorders."""
import numpy as np

from . import rules

ENT_DTYPE = np.dtype([('type', 'u1'), ('health', 'i1'), ('x', '<i2'), ('y', '<i2'), ('aux', '<i2')])
assert ENT_DTYPE.itemsize == 8
PS = {name: i for i, name in enumerate(rules.PSTATE)}
T_PLAYER, T_COW, T_ZOMBIE, T_SKELETON, T_ARROW, T_PLANT = 1, 2, 3, 4, 5, 6


def canonical(mat, ents, inventory, achievements, pstate, touched, area):
  """One env's arrays (numpy, host) -> the canonical state dict compared by the parity tests:
  mat uint8 [W, H]; objs int32 [n, 6] live objects in slot order (type, x, y, health, a, b);
  player int64 [49]; touched int32 [k].  See oracle/canon.py for the field meanings."""
  ents = np.asarray(ents).view(ENT_DTYPE).reshape(-1)
  n = int(pstate[PS['n_slots']])
  rows = []
  for slot in range(1, n):
    e = ents[slot]
    t = int(e['type'])
    if t == 0:
      continue
    health, a, b = int(e['health']), int(e['aux']), 0
    if t == T_PLAYER:
      health, b = int(inventory[0]), int(pstate[PS['sleeping']])
    elif t == T_COW:
      a = 0
    rows.append([t, int(e['x']), int(e['y']), health, a, b])
  player = list(int(v) for v in inventory) + list(int(v) for v in achievements) + [
      int(pstate[PS['hunger2']]), int(pstate[PS['thirst2']]), int(pstate[PS['fatigue']]),
      int(pstate[PS['recover2']]), int(pstate[PS['sleeping']]), int(ents[1]['aux']),
      int(pstate[PS['player_last_health']]), int(pstate[PS['player_x']]),
      int(pstate[PS['player_y']]), int(pstate[PS['env_last_health']]),
      int(np.uint32(pstate[PS['unlocked']]))]
  bits = np.unpackbits(np.asarray(touched, np.uint32).view(np.uint8), bitorder='little')
  return dict(
      mat=np.asarray(mat, np.uint8).reshape(area) & 0x7F,
      objs=np.array(rows, np.int32).reshape(-1, 6),
      player=np.array(player, np.int64),
      touched=np.flatnonzero(bits).astype(np.int32))


def default_slot_capacity(area):
  """Entity slots per env.  The reference's slot list is unbounded (engine.py:54-55); here slots are
  compacted in order when half full, so capacity bounds LIVE objects (about 50 at 64x64, 720 at
  256x256) with generous headroom."""
  return int(min(65535, max(256, (area[0] * area[1]) // 16)))
