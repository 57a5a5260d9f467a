"""`-m gpu`: the CUDA library (through the C ABI and crafter_b200.Env) replays the golden
trajectories of the unmodified reference bit for bit -- integer state, reward, done, observation."""
import functools
import os

import numpy as np
import pytest

from tests import parity
from tests.golden_util import Fixture, NAMES

pytestmark = pytest.mark.gpu


def make_env(**kwargs):
  import crafter_b200
  return crafter_b200.Env(**kwargs)


@pytest.mark.parametrize('name', NAMES)
def test_cuda_matches_reference(name):
  parity.replay(Fixture(name), make_env, auto_reset=False)


@pytest.mark.parametrize('name', ['default_random', 'default_short', 'default_rich'])
def test_cuda_auto_reset(name):
  parity.replay(Fixture(name), make_env, auto_reset=True)


def test_cuda_slot_compaction():
  fx = Fixture('default_fighter')
  env = parity.replay(fx, functools.partial(make_env, slot_capacity=128), steps=400)
  assert int(env.state['pstate'][:, 14].abs().sum()) == 0


def test_cuda_native_library_loaded():
  """The product path is the in-tree CUDA library; there is no fallback to fail over to."""
  from crafter_b200 import _cabi
  lib = _cabi.load()
  assert lib.cr_abi_version() == _cabi.ABI_VERSION
  with open('/proc/self/maps') as f:
    assert 'libcrafter_b200.so' in f.read()


class HostStepEnv:
  """crafter_b200.Env driven through cr_step_host (pinned host buffers in and out)."""

  def __init__(self, **kwargs):
    import torch
    import crafter_b200
    self._env = crafter_b200.Env(**kwargs)
    B = self._env.num_envs
    self._a = torch.zeros(B, dtype=torch.int32).pin_memory()
    self._r = torch.zeros(B, dtype=torch.float32).pin_memory()
    self._d = torch.zeros(B, dtype=torch.bool).pin_memory()

  def __getattr__(self, name):
    return getattr(self._env, name)

  def step(self, actions):
    import torch
    self._a.copy_(torch.as_tensor(np.asarray(actions), dtype=torch.int32))
    self._env.step_host(self._a, self._r, self._d)
    return self._env._obs, self._r.clone(), self._d.clone(), {}


def test_cuda_step_host_matches_reference():
  """The host-buffer entry point (H2D actions, D2H reward/done on a graph branch) replays the
  golden trajectories too; alternating it with the device entry point keeps both graphs valid."""
  parity.replay(Fixture('default_short'), HostStepEnv, auto_reset=True)
  parity.replay(Fixture('default_random'), HostStepEnv, auto_reset=False, steps=200)


def test_cuda_step_host_copies_obs_when_asked():
  """cr_step_host with an obs_host buffer: the pinned copy equals the device observation."""
  import torch
  import crafter_b200
  env = crafter_b200.Env(num_envs=5, seed=9, auto_reset=True)
  env.reset()
  a = torch.zeros(5, dtype=torch.int32).pin_memory()
  r = torch.zeros(5, dtype=torch.float32).pin_memory()
  d = torch.zeros(5, dtype=torch.bool).pin_memory()
  o = torch.zeros(5, 64, 64, 3, dtype=torch.uint8).pin_memory()
  for t in range(12):
    a.copy_(torch.randint(0, 17, (5,), dtype=torch.int32))
    env.step_host(a, r, d, o)
    assert torch.equal(o, env._obs.cpu()) and int(o.sum()) > 0
    assert torch.equal(r, env._reward_buf.cpu()) and torch.equal(d, env._done.cpu())


@pytest.mark.parametrize('size', [(128, 128), (96, 80), (512, 512)])
def test_cuda_render_at_other_sizes(size):
  from oracle import oracle_env
  env = make_env(num_envs=3, seed=77)
  env.reset()
  frames = env.render(size).cpu().numpy()
  for i in range(3):
    ref = oracle_env.OracleEnv(seed=77 + i, size=size)
    assert (ref.reset() == frames[i]).all(), (size, i)


def test_cuda_info_tensors():
  import torch
  env = make_env(num_envs=4, seed=3)
  env.reset()
  obs, reward, done, info = env.step(torch.zeros(4, dtype=torch.int32, device='cuda'))
  assert info['inventory'].shape == (4, 16) and info['achievements'].shape == (4, 22)
  assert info['player_pos'].tolist() == [[32, 32]] * 4
  assert info['semantic'].shape == (4, 64, 64) and int(info['semantic'][0, 32, 32]) == 13
  assert info['discount'].tolist() == [1.0] * 4
  assert obs.dtype == torch.uint8 and obs.shape == (4, 64, 64, 3) and obs.is_cuda


def test_cuda_generic_kernels_on_default_geometry(monkeypatch):
  """The default geometry normally runs the constant-folded kernel instantiations; the generic ones
  must give the same answer on it."""
  monkeypatch.setenv('CRAFTER_B200_NO_SPECIALIZE', '1')
  parity.replay(Fixture('default_random'), make_env, auto_reset=True, steps=300)


def test_cuda_stats_recorder_and_vector_api(tmp_path):
  """N2/N3 of SURVEY.md 8(f): stats.jsonl lines equal the reference StatsRecorder's; the vector
  adaptor splits done into terminated / truncated."""
  import json
  import torch
  from crafter_b200 import recorder, vector
  from tests import stats_util
  fx = Fixture('default_short')
  env = recorder.StatsRecorder(make_env(num_envs=fx.K, seed=fx.seed0, auto_reset=True, **fx.kwargs), tmp_path)
  env.reset()
  actions = np.stack([fx.env(i, 'actions') for i in range(fx.K)], 1)
  for t in range(fx.T):
    env.step(torch.as_tensor(actions[t], device='cuda'))
  env.close()
  got = [json.loads(l) for l in (tmp_path / 'stats.jsonl').read_text().splitlines()]
  assert got == stats_util.expected_lines(fx) and len(got) > 0
  venv = vector.make('CrafterNoReward-v1', num_envs=3, seed=fx.seed0, length=20)
  obs, info = venv.reset()
  for t in range(20):
    obs, reward, terminated, truncated, info = venv.step(torch.zeros(3, dtype=torch.int32, device='cuda'))
  assert truncated.all() and not terminated.any() and float(reward.abs().sum()) == 0.0


def test_cuda_episode_recorder_npz(tmp_path):
  """N2 of SURVEY.md 8(f): the .npz episodes carry the reference EpisodeRecorder's keys
  (recorder.py:125-152) and, row by row, the values the reference produced (golden fixture)."""
  import torch
  from crafter_b200 import recorder
  from tests import stats_util
  fx = Fixture('default_short')
  env = recorder.EpisodeRecorder(
      make_env(num_envs=fx.K, seed=fx.seed0, auto_reset=False, **fx.kwargs), tmp_path, env_ids=range(fx.K))
  stats_util.record_and_check_episodes(fx, env, lambda a: torch.as_tensor(a, device='cuda'))


def test_cuda_episode_recorder_over_auto_reset(tmp_path):
  """The same .npz episodes from an AUTO-RESETTING batch: the terminal row of every episode comes from
  info['final_observation' / 'final_semantic' / 'final_inventory' / 'final_achievements'] (k_terminal and
  the tick's terminal record), and the composite Recorder (recorder.py:9-25) chains the three recorders."""
  import torch
  from crafter_b200 import recorder
  from tests import stats_util
  fx = Fixture('default_short')
  with pytest.raises(ValueError):
    recorder.EpisodeRecorder(make_env(num_envs=2, seed=1, auto_reset=True), tmp_path)
  env = recorder.EpisodeRecorder(
      make_env(num_envs=fx.K, seed=fx.seed0, auto_reset=True, final_obs=True, **fx.kwargs), tmp_path / 'a', env_ids=range(fx.K))
  stats_util.record_and_check_episodes(fx, env, lambda a: torch.as_tensor(a, device='cuda'))
  both = recorder.Recorder(make_env(num_envs=fx.K, seed=fx.seed0, auto_reset=True, final_obs=True, **fx.kwargs),
                           tmp_path / 'b', video_size=(96, 96), env_ids=(0, 1))
  both.reset()
  actions = np.stack([fx.env(i, 'actions') for i in range(fx.K)], 1)
  for t in range(fx.T):
    both.step(torch.as_tensor(actions[t], device='cuda'))
  files = sorted(p.name for p in (tmp_path / 'b').iterdir())
  assert 'stats.jsonl' in files and any(f.endswith('.npz') for f in files)
  assert any(f.endswith(('.mp4', '.gif')) or (f.endswith('.npz') and 'frames' in np.load(tmp_path / 'b' / f).files) for f in files)


def test_cuda_vector_env_follows_gymnasium(tmp_path):
  """N3 of SURVEY.md 8(f): batched spaces, five-tuple step, same-step autoreset with the terminal
  transition under info['final_obs'] / info['final_info'] and their masks (gymnasium 1.x), the legacy
  `final_observation` alias, terminated vs truncated, the reference's two ids."""
  import torch
  from crafter_b200 import vector
  fx = Fixture('default_short')  # length 50: every env truncates at step 50, some die before
  venv = vector.make('CrafterReward-v1', num_envs=fx.K, seed=fx.seed0, **fx.kwargs)
  assert venv.single_observation_space.shape == (64, 64, 3) and venv.observation_space.shape == (fx.K, 64, 64, 3)
  assert venv.single_action_space.n == 17 and str(venv.metadata['autoreset_mode']).lower().endswith('same_step')
  assert venv.observation_space.contains(venv.observation_space.sample())
  ref = make_env(num_envs=fx.K, seed=fx.seed0, auto_reset=False, **fx.kwargs)  # the same batch, caller resets
  obs, info = venv.reset()
  assert torch.equal(obs, ref.reset()) and info == {}
  actions = np.stack([fx.env(i, 'actions') for i in range(fx.K)], 1)
  ended = 0
  for t in range(fx.T):
    a = torch.as_tensor(actions[t], device='cuda')
    obs, reward, terminated, truncated, info = venv.step(a)
    robs, rreward, rdone, rinfo = ref.step(a)
    done = terminated | truncated
    assert torch.equal(done, rdone) and torch.equal(reward, rreward) and not bool((terminated & truncated).any())
    assert torch.equal(info['_final_obs'], done) and torch.equal(info['_final_info'], done)
    assert info['final_observation'] is info['final_obs']
    if bool(done.any()):
      idx = done.nonzero().flatten()
      assert torch.equal(info['final_obs'][idx], robs[idx])  # the frame the reference returns with done=True
      assert torch.equal(info['final_info']['inventory'][idx], rinfo['inventory'][idx])
      assert torch.equal(info['final_info']['achievements'][idx], rinfo['achievements'][idx])
      assert torch.equal(info['discount'][idx], rinfo['discount'][idx])  # 0 for a death even though the env was regenerated
      assert torch.equal(terminated[idx], rinfo['inventory'][idx, 0] <= 0)
      robs = ref.reset(done).clone()
      ended += int(done.sum())
    assert torch.equal(obs, robs)  # first frame of the next episode where one ended
  assert ended >= 2 * fx.K
  with pytest.raises(ValueError):
    venv.reset(seed=fx.seed0 + 1)
  venv.close()
  nr = vector.make('CrafterNoReward-v1', num_envs=2, seed=1, length=5, to_numpy=True)
  nr.reset()
  for t in range(5):
    obs, reward, terminated, truncated, info = nr.step(np.zeros(2, np.int64))
  assert isinstance(obs, np.ndarray) and truncated.all() and float(np.abs(reward).sum()) == 0.0
  assert isinstance(info['final_info']['inventory'], np.ndarray)


def test_cuda_error_flags_and_unbounded_length():
  """A slot arena that is too small raises through check_errors (the dropped object is no longer
  silent), and length=None steps past the old 100k clamp of the daylight table."""
  import torch
  env = make_env(num_envs=8, seed=3, slot_capacity=24)  # worlds start with ~50 creatures: overflow at the first reset
  env.reset()
  assert env.error_flags() & 1
  with pytest.raises(RuntimeError, match='slot arena overflow'):
    env.check_errors()
  ok = make_env(num_envs=2, seed=3, length=None)
  ok.reset()
  ok.state['pstate'][:, 9] = 150_000  # far beyond the old table
  ok.set_inventory({'health': 9, 'food': 9, 'drink': 9, 'energy': 9})
  ok.step(torch.zeros(2, dtype=torch.int32, device='cuda'))
  assert ok.error_flags() == 0 and int(ok.state['pstate'][0, 9]) == 150_001
  from crafter_b200 import tables
  assert float(tables.daylight_table(150_002)[150_001]) == 1 - abs(np.cos(np.pi * ((150_001 / 300) % 1 + 0.3))) ** 3
  ok.state['pstate'][:, 9] = 1_000_005
  ok.step(torch.zeros(2, dtype=torch.int32, device='cuda'))
  assert ok.error_flags() & 2


def test_cuda_render_subset_and_video_recorder(tmp_path):
  """cr_render_envs draws the rows cr_render draws; the VideoRecorder (recorder.py:68-99) writes
  one file per finished episode with reset frame + one frame per step."""
  import torch
  from crafter_b200 import recorder
  env = make_env(num_envs=6, seed=3, length=12)
  env.reset()
  for t in range(5):
    env.step(torch.full((6,), 1 + t % 4, dtype=torch.int32, device='cuda'))
  for size in (None, (96, 80)):
    full = env.render(size)
    part = env.render(size, env_ids=[4, 1, 1])
    assert torch.equal(part, full[[4, 1, 1]])
  with pytest.raises(IndexError):
    env.render(None, env_ids=[6])
  video = recorder.VideoRecorder(make_env(num_envs=3, seed=3, length=12), tmp_path, size=(128, 128), env_ids=(0, 2))
  video.reset()
  for t in range(12):
    obs, reward, done, info = video.step(torch.zeros(3, dtype=torch.int32, device='cuda'))
  assert bool(done.all()) and len(video.saved) == 2
  assert all(p.exists() and p.stat().st_size > 0 and '-len12' in p.name for p in video.saved)
  if video.saved[0].suffix == '.gif':
    from PIL import Image
    im = Image.open(video.saved[0])
    assert im.size == (128, 128) and 2 <= im.n_frames <= 13  # Pillow merges identical frames


def test_cuda_state_dict_roundtrip_replays_exactly():
  """N4 (first half) of SURVEY.md 8(f): a snapshot of the device state restores an exact replay --
  the keyed random contract has no hidden generator state, so obs / reward / done repeat bit for bit,
  including the auto-resets that fall into the replayed window."""
  import torch
  env = make_env(num_envs=64, seed=11, length=60, auto_reset=True)
  env.reset()
  gen = torch.Generator(device='cuda').manual_seed(5)
  actions = torch.randint(0, 17, (140, 64), generator=gen, device='cuda', dtype=torch.int32)
  for t in range(50):
    env.step(actions[t])
  saved = env.state_dict()
  first = []
  for t in range(50, 140):
    obs, reward, done, info = env.step(actions[t])
    first.append((obs.clone(), reward.clone(), done.clone(), info['inventory'].clone()))
  assert any(bool(d.any()) for _, _, d, _ in first)  # episodes ended (length=60) inside the window
  env.load_state_dict(saved)
  for t in range(50, 140):
    obs, reward, done, info = env.step(actions[t])
    o, r, d, inv = first[t - 50]
    assert torch.equal(obs, o) and torch.equal(reward, r) and torch.equal(done, d)
    assert torch.equal(info['inventory'], inv)


def test_cuda_full_size_batch_equals_its_shards():
  """BASELINE.json's batch (4096 envs) against the same envs run as two shards with `env_offset`:
  a size-independent property at the full size -- per-env seeds depend on the global index only, so
  observations, rewards, dones and the whole integer state agree bit for bit, auto-resets included."""
  import torch
  import crafter_b200
  B, T = 4096, 90
  whole = crafter_b200.Env(num_envs=B, seed=21, length=40, auto_reset=True)
  parts = [crafter_b200.Env(num_envs=B // 2, seed=21, length=40, auto_reset=True, env_offset=o) for o in (0, B // 2)]
  gen = torch.Generator(device='cuda').manual_seed(3)
  actions = torch.randint(0, 17, (T, B), generator=gen, device='cuda', dtype=torch.int32)
  obs = whole.reset()
  assert torch.equal(obs, torch.cat([p.reset() for p in parts]))
  resets = 0
  for t in range(T):
    obs, reward, done, info = whole.step(actions[t])
    outs = [p.step(actions[t, i * (B // 2):(i + 1) * (B // 2)].contiguous()) for i, p in enumerate(parts)]
    assert torch.equal(obs, torch.cat([o[0] for o in outs])), t
    assert torch.equal(reward, torch.cat([o[1] for o in outs])) and torch.equal(done, torch.cat([o[2] for o in outs])), t
    resets += int(done.sum())
  assert resets >= 2 * B  # length 40: every env went through two regenerated worlds
  for key in ('mat', 'ents', 'inventory', 'achievements', 'touched'):
    assert torch.equal(whole.state[key], torch.cat([p.state[key] for p in parts])), key
  ps = torch.cat([p.state['pstate'] for p in parts])
  assert torch.equal(whole.state['pstate'], ps)
