"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

The reference holds no tests or golden vectors, so the fixtures are produced by
running /root/reference/crafter itself under oracle/ref_harness.py (dependency shims + keyed RNG +
canonical chunk order).  Each fixture is a batch of K trajectories with env seeds seed0..seed0+K-1,
reset-on-done like crafter/run_random.py:36-43, and stores per step the digests of the canonical
state (oracle/canon.py), reward, done and the observation digest, plus sparse full snapshots.

    python tools/make_golden.py            # writes every fixture listed in SPECS
"""
import pathlib
import sys
import zlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import canon  # noqa: E402
from oracle import oracle_env  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

RICH = dict(sapling=9, wood=9, stone=9, coal=9, iron=9, diamond=2, wood_pickaxe=1,
            stone_pickaxe=1, iron_pickaxe=1, wood_sword=1)

SPECS = {
    # name: (env kwargs, seed0, K, T, policy, boost)
    'default_random': (dict(), 0, 8, 600, 'random', None),
    'default_rich': (dict(), 100, 4, 600, 'random', RICH),
    'default_sleepy': (dict(), 200, 4, 900, 'sleepy', None),
    'default_fighter': (dict(), 700, 4, 800, 'fighter', dict(RICH, iron_sword=1)),
    'default_short': (dict(length=50), 300, 2, 160, 'random', None),
    'big_view': (dict(view=(15, 15), size=(128, 128)), 400, 2, 300, 'random', RICH),
    'odd_geometry': (dict(area=(40, 52), view=(7, 9), size=(58, 75)), 500, 2, 300, 'random', RICH),
    'big_area': (dict(area=(256, 256)), 600, 1, 120, 'random', None),
    # map smaller than the view: out-of-map cells in every frame, crafting / placing / arrows at the
    # map edges, clipped chunks
    'tiny_area': (dict(area=(18, 14)), 800, 4, 500, 'random', RICH),
    # even view: the player is not in the middle of the local grid (offset = grid // 2, engine.py:161);
    # a wide flat one: 12 x 3 local cells over 2 item rows
    'even_view': (dict(view=(8, 8)), 900, 2, 150, 'random', RICH),
    'wide_view': (dict(view=(12, 5), size=(96, 45)), 950, 2, 150, 'random', RICH),
}
SNAP_EVERY = 100


def actions_for(policy, seed, T):
  rs = np.random.RandomState(10_000 + seed)
  if policy == 'random':
    return rs.randint(0, 17, T).astype(np.int32)
  if policy == 'sleepy':  # mostly sleep/noop with some wandering and hitting: reaches night + sleep
    return rs.choice(np.array([0, 6, 6, 6, 6, 5, 1, 2, 3, 4], np.int32), T)
  if policy == 'fighter':  # face-and-hit heavy: cows, zombies, plants, mining
    return rs.choice(np.array([5, 5, 5, 5, 1, 2, 3, 4, 10], np.int32), T)
  raise ValueError(policy)


def run_one(kwargs, seed, T, policy, boost):
  env = rh.make_env(seed, **kwargs)
  actions = actions_for(policy, seed, T)
  rec = dict(actions=actions, reward=np.zeros(T), done=np.zeros(T, bool),
             obs_crc=np.zeros(T, np.int64), player=np.zeros((T, 49), np.int64),
             daylight=np.zeros(T))
  for k in canon.KEYS:
    rec[k + '_crc'] = np.zeros(T, np.int64)
  resets = dict(step=[], obs_crc=[], player=[])
  for k in canon.KEYS:
    resets[k + '_crc'] = []
  snaps = {}
  stats = dict(night=0, sleeping=0, episodes=0)

  def do_reset(step):
    obs = env.reset()
    if boost:
      rh.boost_inventory(env, boost)
      obs = env.render()
    st = rh.export_state(env)
    resets['step'].append(step)
    resets['obs_crc'].append(zlib.crc32(obs.tobytes()))
    resets['player'].append(st['player'])
    for k, v in canon.digest(st).items():
      resets[k + '_crc'].append(v)
    return obs, st

  obs, st = do_reset(-1)
  snaps['reset_obs'] = obs
  snaps['reset_mat'] = st['mat']
  snaps['reset_objs'] = st['objs']
  for t in range(T):
    obs, reward, done, info = env.step(int(actions[t]))
    st = rh.export_state(env)
    rec['reward'][t], rec['done'][t] = reward, done
    rec['obs_crc'][t] = zlib.crc32(obs.tobytes())
    rec['player'][t] = st['player']
    rec['daylight'][t] = st['daylight']
    for k, v in canon.digest(st).items():
      rec[k + '_crc'][t] = v
    stats['night'] += st['daylight'] < 0.5
    stats['sleeping'] += int(st['player'][42])
    sem = info['semantic']
    if t % SNAP_EVERY == SNAP_EVERY - 1 or t == T - 1:
      snaps[f'obs_{t}'] = obs
      snaps[f'mat_{t}'] = st['mat']
      snaps[f'objs_{t}'] = st['objs']
      snaps[f'semantic_{t}'] = sem
    if done:
      stats['episodes'] += 1
      do_reset(t)
  out = dict(rec)
  for k, v in resets.items():
    out['reset_' + k] = np.array(v, np.int64)
  out.update(snaps)
  out['unlocked_total'] = np.array(
      [int(st['player'][48])], np.int64)
  return out, stats


def main(names):
  outdir = ROOT / 'tests' / 'golden'
  outdir.mkdir(parents=True, exist_ok=True)
  for name in names:
    kwargs, seed0, K, T, policy, boost = SPECS[name]
    blob = dict(
        meta_area=np.array(kwargs.get('area', (64, 64))), meta_view=np.array(kwargs.get('view', (9, 9))),
        meta_size=np.array(kwargs.get('size', (64, 64))), meta_length=np.array(kwargs.get('length', 10000)),
        meta_seed0=np.array(seed0), meta_K=np.array(K), meta_T=np.array(T),
        meta_boost_items=np.array(list((boost or {}).keys())),
        meta_boost_values=np.array(list((boost or {}).values()), np.int64))
    agg = dict(night=0, sleeping=0, episodes=0)
    ach = 0
    for i in range(K):
      out, stats = run_one(kwargs, seed0 + i, T, policy, boost)
      for k, v in out.items():
        blob[f'e{i}_{k}'] = v
      for k in agg:
        agg[k] += int(stats[k])
      ach |= int(out['player'][:, 16:38].max(0).astype(bool) @ (1 << np.arange(22)))
    # host tables used during generation (cross-machine libm / SVML check)
    view = kwargs.get('view', (9, 9))
    size = kwargs.get('size', (64, 64))
    t = oracle_env.render_tables(view, size)
    blob['table_vignette'] = t['vignette']
    blob['table_daylight'] = oracle_env.daylight_table(int(kwargs.get('length', 10000)) + 2)
    path = outdir / f'{name}.npz'
    np.savez_compressed(path, **blob)
    names_unlocked = [a for j, a in enumerate(oracle_env.ACHIEVEMENTS) if ach >> j & 1]
    print(f'{name}: K={K} T={T} {path.stat().st_size/1024:.0f} KiB  coverage {agg}  '
          f'achievements {len(names_unlocked)}/22: {names_unlocked}')


if __name__ == '__main__':
  main(sys.argv[1:] or list(SPECS))
