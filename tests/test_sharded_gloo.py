"""Multi-process host logic (CPU, gloo, world_size 2): a batch sharded over ranks reproduces the
single-process batch exactly, and ShardedEnv.gather reassembles it.  The device logic behind the
env is the host-sim build of the kernels' headers; on the B200 box the same class wraps the CUDA Env."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

K, T = 4, 60


class TorchHostSim:
  """HostSimEnv with torch tensors out (what ShardedEnv.gather expects)."""

  def __init__(self, **kwargs):
    from tests import hostsim_env
    self._e = hostsim_env.HostSimEnv(**kwargs)

  def reset(self, mask=None):
    return torch.from_numpy(self._e.reset(mask).copy())

  def step(self, actions):
    obs, reward, done = self._e.step(np.asarray(actions))
    return torch.from_numpy(obs.copy()), torch.from_numpy(reward.copy()), torch.from_numpy(done.copy()), {}

  def snapshot(self, i):
    return self._e.snapshot(i)


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from crafter_b200.sharded import ShardedEnv
  env = ShardedEnv(num_envs=K, seed=11, env_factory=TorchHostSim, auto_reset=True, length=40)
  assert env.local_num_envs == K // world and env.offset == rank * (K // world)
  actions = np.random.RandomState(5).randint(0, 17, (T, K))
  obs = env.reset()
  frames = [env.gather(obs)]
  for t in range(T):
    obs, reward, done, _ = env.step(actions[t, env.local_slice()])
    full_obs, full_reward, full_done = env.gather(obs, reward, done)
    frames.append(full_obs)
    if rank == 0:
      out.setdefault('reward', []).append(full_reward.numpy().copy())
      out.setdefault('done', []).append(full_done.numpy().copy())
  if rank == 0:
    out['obs'] = np.stack([f.numpy() for f in frames])
  dist.destroy_process_group()


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def test_two_rank_shards_equal_one_batch():
  ctx = mp.get_context('spawn')  # never fork a multi-threaded pytest process
  manager = ctx.Manager()
  out = manager.dict()
  port = _free_port()
  procs = [ctx.Process(target=_worker_entry, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(300)
    assert p.exitcode == 0
  from tests import hostsim_env
  ref = hostsim_env.HostSimEnv(num_envs=K, seed=11, auto_reset=True, length=40)
  actions = np.random.RandomState(5).randint(0, 17, (T, K))
  frames = [ref.reset().copy()]
  rewards, dones = [], []
  for t in range(T):
    obs, reward, done = ref.step(actions[t])
    frames.append(obs.copy()); rewards.append(reward.copy()); dones.append(done.copy())
  assert (np.stack(frames) == out['obs']).all()
  assert (np.stack(rewards) == np.stack(out['reward'])).all()
  assert (np.stack(dones) == np.stack(out['done'])).all()
  assert np.stack(dones).any(), 'the run should cross at least one auto-reset'


def _worker_entry(rank, world, port, shared):
  local = {}
  _worker(rank, world, port, local)
  if rank == 0:
    for k, v in local.items():
      shared[k] = v


def test_shard_arithmetic():
  from crafter_b200.sharded import shard_of
  assert shard_of(32768, 3, 8) == (12288, 4096)
  with pytest.raises(ValueError):
    shard_of(10, 0, 4)
