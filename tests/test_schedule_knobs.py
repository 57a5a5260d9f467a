"""Reset patterns and the A/B knobs of the tick on the host-sim (the same device functions as the
kernels, one lane; tests/test_simt_kernels.py repeats them on the kernels themselves, both step
schedules): episodes of length 1 / 2 / 3 (an env finishes again while its next world is the one
generated last), explicit reset(mask) between auto-resets, the terminal frame of auto-reset
(final_obs), DRAW_PREFETCH and INCR_CENSUS on / off."""
import numpy as np
import pytest

from oracle import canon
from tests import hostsim_env
from tests import parity
from tests.golden_util import Fixture


def check_against_oracle(make_env, to_numpy, length, steps, K=3, seed=40, **kwargs):
  """auto_reset=True against the C oracle with reset-on-done: after a finished episode the batch
  shows the first frame / state of the next one (DESIGN.md 'Semantics added by batching')."""
  from oracle import oracle_env
  env = make_env(num_envs=K, seed=seed, length=length, auto_reset=True, **kwargs)
  refs = [oracle_env.OracleEnv(seed=seed + i, length=length, **kwargs) for i in range(K)]
  obs = to_numpy(env.reset())
  for i, ref in enumerate(refs):
    assert (ref.reset() == obs[i]).all(), ('reset', i)
  rs = np.random.RandomState(3)
  episodes = 0
  for t in range(steps):
    actions = rs.randint(0, 17, K).astype(np.int32)
    obs, reward, done = env.step(actions)[:3]
    obs, reward, done = to_numpy(obs), to_numpy(reward), to_numpy(done).astype(bool)
    for i, ref in enumerate(refs):
      o, r, d = ref.step(int(actions[i]))
      assert np.float32(r) == reward[i] and d == bool(done[i]), (length, t, i)
      if d:
        o = ref.reset()
        episodes += 1
      assert canon.diff(ref.export_state(), env.snapshot(i)) is None, (length, t, i)
      assert (o == obs[i]).all(), (length, t, i, 'obs')
  return episodes


@pytest.mark.parametrize('length', [1, 2, 3, 7])
def test_back_to_back_resets(length):
  episodes = check_against_oracle(hostsim_env.HostSimEnv, np.asarray, length, steps=14)
  assert episodes >= 3 * (14 // length)


def check_mixed_resets_and_masks(make_env):
  """reset(mask) while other envs have a world pending, then more auto-resets."""
  from oracle import oracle_env
  K, seed, length = 4, 90, 3
  env = make_env(num_envs=K, seed=seed, length=length, auto_reset=True)
  refs = [oracle_env.OracleEnv(seed=seed + i, length=length) for i in range(K)]
  obs = np.asarray(env.reset())
  for i, ref in enumerate(refs):
    assert (ref.reset() == obs[i]).all()
  rs = np.random.RandomState(5)
  for t in range(13):
    if t in (2, 3, 7):  # explicit reset of a subset right after / before auto-resets
      mask = np.array([t % 2 == 0, True, False, t == 7])
      obs = np.asarray(env.reset(mask)).copy()
      for i in np.flatnonzero(mask):
        assert (refs[i].reset() == obs[i]).all(), (t, i)
        assert canon.diff(refs[i].export_state(), env.snapshot(i)) is None
    actions = rs.randint(0, 17, K).astype(np.int32)
    obs, reward, done = env.step(actions)[:3]
    for i, ref in enumerate(refs):
      o, r, d = ref.step(int(actions[i]))
      assert d == bool(done[i])
      if d:
        o = ref.reset()
      assert canon.diff(ref.export_state(), env.snapshot(i)) is None, (t, i)
      assert (o == np.asarray(obs)[i]).all(), (t, i)


def test_mixed_resets_and_masks():
  check_mixed_resets_and_masks(hostsim_env.HostSimEnv)


def check_terminal_frames(make_env, to_numpy, length, steps, K=3, seed=11, **kwargs):
  """auto_reset=True, final_obs=True: when an episode ends inside step(), obs shows the first frame of
  the next episode and info['final_observation'] / the env's final_obs buffer the frame the reference
  returns with done=True (env.py:96,118) -- drawn after the step's balance (env.py:90-95), which only
  this frame can tell.  Also the terminal inventory / achievements (env.py:108-115)."""
  from oracle import oracle_env
  env = make_env(num_envs=K, seed=seed, length=length, auto_reset=True, final_obs=True, **kwargs)
  refs = [oracle_env.OracleEnv(seed=seed + i, length=length, **kwargs) for i in range(K)]
  obs = to_numpy(env.reset())
  for i, ref in enumerate(refs):
    assert (ref.reset() == obs[i]).all()
  rs = np.random.RandomState(8)
  ended = 0
  for t in range(steps):
    actions = rs.randint(0, 17, K).astype(np.int32)
    out = env.step(actions)
    obs, done = to_numpy(out[0]), to_numpy(out[2]).astype(bool)
    final = to_numpy(env.final_obs if hasattr(env, 'final_obs') else out[3]['final_observation'])
    final_sem = to_numpy(env.final_semantic if hasattr(env, 'final_semantic') else out[3]['final_semantic'])
    for i, ref in enumerate(refs):
      o, r, d = ref.step(int(actions[i]))
      assert d == bool(done[i]), (t, i)
      if d:
        assert (o == final[i]).all(), (length, t, i, 'terminal frame')
        assert (ref.semantic() == final_sem[i]).all(), (length, t, i, 'terminal semantic')
        ended += 1
        o = ref.reset()
      assert (o == obs[i]).all(), (length, t, i, 'obs')
  return ended


@pytest.mark.parametrize('length', [1, 10, 20])
def test_terminal_frames(length):
  """length 10 / 20: every terminal step is also a balance step."""
  assert check_terminal_frames(hostsim_env.HostSimEnv, np.asarray, length, steps=41) >= 3 * (41 // length)


# ---- draw prefetch (default on; CRAFTER_B200_DRAW_PREFETCH=0 is the plain per-draw Philox) -------------
@pytest.mark.parametrize('value', ['0', '1'])
@pytest.mark.parametrize('name', ['default_random', 'default_fighter', 'tiny_area'])
def test_draw_prefetch_replays_golden(monkeypatch, name, value):
  monkeypatch.setenv('CRAFTER_B200_DRAW_PREFETCH', value)
  parity.replay(Fixture(name), hostsim_env.HostSimEnv, auto_reset=False, steps=400, check_obs=False)


@pytest.mark.parametrize('value', ['0', '1'])
@pytest.mark.parametrize('group', ['directed_default', 'fuzz_small'])
def test_draw_prefetch_replays_scenarios(monkeypatch, group, value):
  """`many_objects` draws far more than 32 times per tick: table and computed draws in one step."""
  from tests import scenario_util as su
  from tests.test_scenarios_golden import replay_group
  monkeypatch.setenv('CRAFTER_B200_DRAW_PREFETCH', value)
  replay_group(group, hostsim_env.HostSimEnv, su.load_numpy)


# ---- incremental census (default on; CRAFTER_B200_INCR_CENSUS=0 re-counts on every balance tick) ------
def counts_are_current(env):
  before = env.state['chunk_cnt'].copy()
  env.recount()
  return (before == env.state['chunk_cnt']).all() and int(before.sum()) > 0


@pytest.mark.parametrize('name,steps', [('default_random', 400), ('default_rich', 400), ('tiny_area', 300),
                                        ('odd_geometry', 200), ('big_area', 60)])
def test_incremental_census_replays_golden(monkeypatch, name, steps):
  monkeypatch.setenv('CRAFTER_B200_INCR_CENSUS', '1')
  env = parity.replay(Fixture(name), hostsim_env.HostSimEnv, auto_reset=False, steps=steps, check_obs=False)
  assert counts_are_current(env)


@pytest.mark.parametrize('group', ['directed_default', 'fuzz_default', 'fuzz_big_area'])
def test_incremental_census_replays_scenarios(monkeypatch, group):
  """Collecting, placing (stone on water / lava, tables, furnaces), arrows turning tables into path:
  every terrain write of the tick, then balance ticks that read the counts."""
  from tests import scenario_util as su
  from tests.test_scenarios_golden import replay_group
  monkeypatch.setenv('CRAFTER_B200_INCR_CENSUS', '1')
  env = replay_group(group, hostsim_env.HostSimEnv, su.load_numpy)
  assert counts_are_current(env)


def test_incremental_census_with_auto_reset(monkeypatch):
  monkeypatch.setenv('CRAFTER_B200_INCR_CENSUS', '1')
  env = parity.replay(Fixture('default_short'), hostsim_env.HostSimEnv, auto_reset=True)
  assert counts_are_current(env)


@pytest.mark.parametrize('name,steps', [('default_random', 300), ('big_area', 40)])
def test_census_by_recount_replays_golden(monkeypatch, name, steps):
  """The A/B fallback: no chunk_cnt buffer, every balance tick counts the cells."""
  monkeypatch.setenv('CRAFTER_B200_INCR_CENSUS', '0')
  env = parity.replay(Fixture(name), hostsim_env.HostSimEnv, auto_reset=False, steps=steps, check_obs=False)
  assert 'chunk_cnt' not in env.state
