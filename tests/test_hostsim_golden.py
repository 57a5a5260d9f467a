"""CPU-only CI for the kernel logic: the device headers compiled for the host (tests/hostsim) replay
the reference's golden trajectories bit for bit.  The same `parity.replay` runs against the CUDA
library in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests import hostsim_env
from tests import parity
from tests.golden_util import Fixture, NAMES


@pytest.mark.parametrize('name', NAMES)
def test_device_logic_matches_reference(name):
  parity.replay(Fixture(name), hostsim_env.HostSimEnv, auto_reset=False)


@pytest.mark.parametrize('name', ['default_random', 'default_short'])
def test_device_logic_auto_reset(name):
  parity.replay(Fixture(name), hostsim_env.HostSimEnv, auto_reset=True)


def test_slot_compaction_preserves_semantics():
  """A small slot arena forces order-preserving compaction almost every step (csrc/cr_update.h
  compact_slots); trajectories must not change (only relative slot order is semantic)."""
  import functools
  fx = Fixture('default_fighter')
  env = parity.replay(fx, functools.partial(hostsim_env.HostSimEnv, slot_capacity=128), steps=400)
  assert (env.state['pstate'][:, 14] == 0).all()  # no slot overflow


def test_render_uncached_object_cells():
  """With a one-entry object-tile cache nearly every visible object cell takes the per-pixel path
  of csrc/cr_render.h (render_uncached); frames must not change."""
  import functools
  for name in ('default_fighter', 'big_view'):
    parity.replay(Fixture(name), functools.partial(hostsim_env.HostSimEnv, max_obj_tiles=1), steps=320)


@pytest.mark.parametrize('size', [(128, 128), (96, 80), (512, 512)])
def test_render_at_other_sizes(size):
  """Env.render(size) (env.py:120-123): the same state rendered at another size equals the
  reference pipeline at that size (reset frame, day; (512, 512) takes the uncached per-pixel path)."""
  import numpy as np
  from oracle import oracle_env
  hs = hostsim_env.HostSimEnv(num_envs=2, seed=77, size=size)
  obs = hs.reset()
  for i in range(2):
    ref = oracle_env.OracleEnv(seed=77 + i, size=size)
    assert (ref.reset() == obs[i]).all(), (size, i, np.argwhere(ref.reset() != obs[i])[:4])


class _TorchView:
  """Minimal torch facade over HostSimEnv so that crafter_b200.recorder.StatsRecorder can wrap it."""

  def __init__(self, **kwargs):
    import torch
    self._e = hostsim_env.HostSimEnv(**kwargs)
    self._torch = torch
    self._env_offset = 0

  @property
  def state(self):
    return {k: self._torch.from_numpy(v) for k, v in self._e.state.items() if v.dtype.kind != 'u' or v.dtype.itemsize == 1}

  def reset(self, mask=None):
    m = None if mask is None else np.asarray(mask)
    return self._torch.from_numpy(self._e.reset(m))

  def step(self, actions):
    t = self._torch
    obs, reward, done = self._e.step(np.asarray(actions))
    st = self._e.state
    info = dict(
        inventory=t.from_numpy(st['inventory']), achievements=t.from_numpy(st['achievements']),
        player_pos=t.from_numpy(st['pstate'][:, 12:14]), reward=t.from_numpy(reward),
        semantic=t.from_numpy(self._e.semantic()),
        discount=t.from_numpy(1.0 - (st['inventory'][:, 0] <= 0).astype(np.float32)))
    return t.from_numpy(obs), t.from_numpy(reward), t.from_numpy(done.copy()), info


@pytest.mark.parametrize('name', ['default_short', 'default_fighter'])
def test_stats_recorder_lines_match_reference(name, tmp_path):
  """crafter_b200.recorder.StatsRecorder writes the lines the reference's StatsRecorder would."""
  import json
  import numpy as np
  from crafter_b200 import recorder
  from tests import stats_util
  fx = Fixture(name)
  env = recorder.StatsRecorder(_TorchView(num_envs=fx.K, seed=fx.seed0, auto_reset=True, **fx.kwargs), tmp_path)
  env.reset()
  if fx.boost:
    env._env._e.set_inventory(fx.boost)
  actions = np.stack([fx.env(i, 'actions') for i in range(fx.K)], 1)
  for t in range(fx.T):
    done = env.step(actions[t])[2].numpy()
    if fx.boost and done.any():
      env._env._e.set_inventory(fx.boost, env_ids=np.flatnonzero(done))
  env.close()
  got = [json.loads(l) for l in (tmp_path / 'stats.jsonl').read_text().splitlines()]
  want = stats_util.expected_lines(fx)
  assert len(got) == len(want) > 0
  assert got == want


def test_episode_recorder_npz_matches_reference(tmp_path):
  """crafter_b200.recorder.EpisodeRecorder against the fixture of the reference."""
  from crafter_b200 import recorder
  from tests import stats_util
  fx = Fixture('default_short')
  env = recorder.EpisodeRecorder(
      _TorchView(num_envs=fx.K, seed=fx.seed0, auto_reset=False, **fx.kwargs), tmp_path, env_ids=range(fx.K))
  stats_util.record_and_check_episodes(fx, env, lambda a: a)
