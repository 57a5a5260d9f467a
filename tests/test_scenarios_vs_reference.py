"""Perturbation fuzz against the LIVE reference (build container only; skipped where
/root/reference is absent): random-policy trajectories rarely reach the corner cases of the rules
(crafting at the map edge, lava, dying mobs that still act, arrows hitting things, ripe plants,
full inventories ...), so this test teleports the player, rewrites terrain, spawns creatures with
odd attributes and sets inventories *inside the reference env through its own API*, loads the very
same canonical state into the device logic (host-sim build of csrc/cr_*.h), and then steps both
with random actions, comparing the full integer state, reward, done and the observation each step."""
import numpy as np
import pytest

from oracle import canon
from oracle import ref_harness as rh
from tests import hostsim_env
from crafter_b200 import rules
from crafter_b200 import state as state_lib

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference not mounted')

MATERIALS = [None] + rules.MATERIALS


def perturb(env, rs, mods):
  """Random edits of a reference env through World / object APIs (engine.py, objects.py)."""
  objects, world, player = mods['objects'], env._world, env._player
  W, H = world.area
  # terrain patches, including the materials that random walks seldom meet
  for _ in range(rs.randint(5, 40)):
    x, y = rs.randint(0, W), rs.randint(0, H)
    world[x, y] = rs.choice(['lava', 'water', 'table', 'furnace', 'tree', 'stone', 'coal', 'iron',
                             'diamond', 'grass', 'sand', 'path'])
  # teleport the player, often to an edge or corner
  for _ in range(20):
    x = rs.choice([0, 1, W - 2, W - 1, rs.randint(0, W)])
    y = rs.choice([0, 1, H - 2, H - 1, rs.randint(0, H)])
    if world[(x, y)][1] is None and world[(x, y)][0] in ('grass', 'sand', 'path'):
      world.move(player, (x, y))
      break
  px, py = player.pos
  for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1)):  # useful neighbours
    x, y = px + dx, py + dy
    if 0 <= x < W and 0 <= y < H and world[(x, y)][1] is None and rs.rand() < 0.6:
      world[x, y] = rs.choice(['table', 'furnace', 'lava', 'water', 'tree', 'stone', 'iron', 'diamond', 'grass'])
  player.facing = [(-1, 0), (1, 0), (0, -1), (0, 1)][rs.randint(4)]
  # creatures around the player with odd attributes (dying mobs, loaded skeletons, ripe plants ...)
  for _ in range(rs.randint(3, 14)):
    x, y = px + rs.randint(-6, 7), py + rs.randint(-6, 7)
    if not (0 <= x < W and 0 <= y < H) or world[(x, y)][1] is not None:
      continue
    kind = rs.randint(5)
    mat = world[(x, y)][0]
    if kind == 0 and mat in ('grass', 'sand', 'path'):
      o = objects.Zombie(world, (x, y), player); o.health = rs.randint(0, 6); o.cooldown = rs.randint(0, 6)
    elif kind == 1 and mat in ('grass', 'sand', 'path'):
      o = objects.Skeleton(world, (x, y), player); o.health = rs.randint(0, 4); o.reload = rs.randint(0, 5)
    elif kind == 2 and mat in ('grass', 'sand', 'path'):
      o = objects.Cow(world, (x, y)); o.health = rs.randint(0, 4)
    elif kind == 3 and mat in ('grass', 'sand', 'path', 'water', 'lava'):
      o = objects.Arrow(world, (x, y), [(-1, 0), (1, 0), (0, -1), (0, 1)][rs.randint(4)])
    elif kind == 4 and mat == 'grass':
      o = objects.Plant(world, (x, y)); o.grown = rs.choice([0, 299, 300, 301, 500]); o.health = rs.randint(0, 2)
    else:
      continue
    world.add(o)
  for name in player.inventory:
    player.inventory[name] = int(rs.choice([0, 0, 1, 2, 5, 9]))
  player.inventory['health'] = int(rs.randint(1, 10))
  player.sleeping = bool(rs.rand() < 0.1)
  if rs.rand() < 0.5:
    player.inventory['energy'] = 9
  player._hunger, player._thirst = float(rs.randint(0, 26)), float(rs.randint(0, 21))
  player._fatigue, player._recover = int(rs.randint(-10, 31)), float(rs.randint(-15, 26))
  player._last_health = player.health
  env._last_health = player.health
  env._step = int(rs.choice([env._step, 140, 147, 200, 271, 299, 9]))
  env._update_time()


def load_into_hostsim(hs, i, env, st):
  """Write a canonical reference state into env i of a HostSimEnv (layout: csrc/cr_common.h)."""
  s = hs.state
  W, H = hs.area
  s['mat'][i] = st['mat'].reshape(-1)
  s['objmap'][i] = 0
  ents = s['ents'][i].view(state_lib.ENT_DTYPE)
  ents[:] = 0
  for k, (t, x, y, health, a, b) in enumerate(st['objs']):
    slot = k + 1
    ents[slot] = (t, min(health, 127), x, y, a)
    s['objmap'][i][x * H + y] = slot
  p = st['player']
  s['inventory'][i] = p[:16]
  s['achievements'][i] = p[16:38]
  ps = s['pstate'][i]
  ps[:] = 0
  ps[state_lib.PS['hunger2']], ps[state_lib.PS['thirst2']] = p[38], p[39]
  ps[state_lib.PS['fatigue']], ps[state_lib.PS['recover2']] = p[40], p[41]
  ps[state_lib.PS['sleeping']] = p[42]
  ps[state_lib.PS['player_last_health']] = p[44]
  ps[state_lib.PS['player_x']], ps[state_lib.PS['player_y']] = p[45], p[46]
  ps[state_lib.PS['env_last_health']] = p[47]
  ps[state_lib.PS['unlocked']] = np.int64(p[48]).astype(np.int32)
  ps[state_lib.PS['n_slots']] = len(st['objs']) + 1
  ps[state_lib.PS['step']] = env._step
  ps[state_lib.PS['episode']] = env._episode
  ps[state_lib.PS['world_seed']] = env._world.random.seed
  s['touched'][i] = 0
  for c in st['touched']:
    s['touched'][i][c >> 5] |= np.uint32(1 << (c & 31))


@pytest.mark.parametrize('geometry', [dict(), dict(area=(24, 20))])
def test_perturbed_states_step_like_the_reference(geometry):
  mods = rh.load()
  rs = np.random.RandomState(2024)
  rounds, steps = (40, 45) if not geometry else (30, 40)
  for r in range(rounds):
    seed = 900 + r
    ref = rh.make_env(seed, **geometry)
    ref.reset()
    hs = hostsim_env.HostSimEnv(num_envs=1, seed=seed, **geometry)
    hs.reset()
    perturb(ref, rs, mods)
    st = rh.export_state(ref)
    load_into_hostsim(hs, 0, ref, st)
    assert canon.diff(st, hs.snapshot(0)) is None
    assert (ref.render() == hs.render()[0]).all(), ('render after load', r)
    for t in range(steps):
      a = int(rs.randint(0, 17)) if rs.rand() < 0.5 else int(rs.choice([5, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]))
      obs, reward, done, info = ref.step(a)
      hobs, hreward, hdone = hs.step(np.array([a]))
      problem = canon.diff(rh.export_state(ref), hs.snapshot(0))
      assert problem is None, (geometry, r, t, a, problem)
      assert np.float32(reward) == hreward[0] and done == bool(hdone[0]), (r, t, reward, hreward[0])
      assert (obs == hobs[0]).all(), (geometry, r, t, 'obs')
      if done:
        break
