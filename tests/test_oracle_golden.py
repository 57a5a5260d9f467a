"""The C oracle replays every golden trajectory of the unmodified reference bit for bit:
grid, slots, inventory, achievements, vitals, touched chunks, reward, done and observation."""
import zlib

import numpy as np
import pytest

from oracle import canon
from oracle import oracle_env as oe
from tests.golden_util import Fixture, NAMES


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference_trajectories(name):
  fx = Fixture(name)
  for i in range(fx.K):
    env = oe.OracleEnv(seed=fx.seed0 + i, **fx.kwargs)
    actions = fx.env(i, 'actions')
    n_reset = 0

    def do_reset():
      nonlocal n_reset
      obs = env.reset()
      if fx.boost:
        env.set_inventory(fx.boost)
        obs = env.render()
      st = env.export_state()
      for k, v in canon.digest(st).items():
        assert v == fx.env(i, f'reset_{k}_crc')[n_reset], (name, i, 'reset', n_reset, k)
      assert zlib.crc32(obs.tobytes()) == fx.env(i, 'reset_obs_crc')[n_reset]
      n_reset += 1

    do_reset()
    assert (env.export_state()['mat'] == fx.env(i, 'reset_mat')).all()
    for t in range(fx.T):
      obs, reward, done = env.step(int(actions[t]))
      st = env.export_state()
      assert (st['player'] == fx.env(i, 'player')[t]).all(), (name, i, t)
      for k, v in canon.digest(st).items():
        assert v == fx.env(i, f'{k}_crc')[t], (name, i, t, k)
      assert reward == fx.env(i, 'reward')[t] and done == fx.env(i, 'done')[t], (name, i, t)
      assert st['daylight'] == fx.env(i, 'daylight')[t]
      assert zlib.crc32(obs.tobytes()) == fx.env(i, 'obs_crc')[t], (name, i, t, 'obs')
      if fx.has(i, f'semantic_{t}'):
        assert (env.semantic() == fx.env(i, f'semantic_{t}')).all()
        assert (obs == fx.env(i, f'obs_{t}')).all()
      if done:
        do_reset()


def test_host_tables_match_generation_machine():
  """daylight/vignette are numpy expressions of the reference (env.py:135-139, engine.py:213-218);
  numpy's exp/cos may differ by an ulp between CPUs (SVML vs libm), which would perturb obs."""
  fx = Fixture('default_random')
  t = oe.render_tables(fx.view, fx.size)
  assert (oe.daylight_table(fx.length + 2) == fx.z['table_daylight']).all()
  assert (t['vignette'] == fx.z['table_vignette']).all()


def test_world_seed_is_cpython_tuple_hash():
  for seed, ep in [(0, 1), (0, 2), (1, 1), (42, 1), (123456, 789), (2 ** 40 + 17, 3)]:
    assert oe.world_seed(seed, ep) == hash((seed, ep)) % (2 ** 31 - 1)
