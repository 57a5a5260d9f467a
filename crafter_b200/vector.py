"""Gymnasium-style vector-env adaptor over `crafter_b200.Env` and the reference's gym ids
(crafter/__init__.py:4-17: `CrafterReward-v1`, `CrafterNoReward-v1`, max 10000 steps).

    venv = crafter_b200.vector.make('CrafterReward-v1', num_envs=4096, seed=0)
    obs, info = venv.reset()
    obs, reward, terminated, truncated, info = venv.step(actions)     # torch.cuda tensors, auto-reset

`terminated` = the player died (`discount` 0 in the reference, env.py:105,111), `truncated` = the
episode hit `length`.  gym / gymnasium are optional (absent in this image): `register()` adds the
two ids when either is importable.
"""
import numpy as np

IDS = {'CrafterReward-v1': dict(reward=True), 'CrafterNoReward-v1': dict(reward=False)}


class VectorEnv:

  def __init__(self, num_envs, **kwargs):
    from .env import Env
    kwargs.setdefault('auto_reset', True)
    self.env = Env(num_envs=num_envs, **kwargs)
    self.num_envs = self.env.num_envs
    self.single_observation_space = self.env.observation_space
    self.single_action_space = self.env.action_space

  def reset(self, seed=None, options=None):
    if seed is not None:
      raise ValueError('the seed is fixed at construction (per-episode seeds derive from it)')
    return self.env.reset(), {}

  def step(self, actions):
    obs, reward, done, info = self.env.step(actions)
    dead = self.env.state['final_stats'][:, 23].bool()
    terminated = done & dead
    truncated = done & ~dead
    return obs, reward, terminated, truncated, info

  def close(self):
    self.env.close()


def make(env_id, num_envs, **kwargs):
  spec = dict(IDS[env_id])
  spec.setdefault('length', 10000)  # max_episode_steps of the reference registration
  spec.update(kwargs)
  return VectorEnv(num_envs, **spec)


def register():
  """Register the reference's ids as vector entry points with gym or gymnasium, when installed."""
  done = []
  for modname in ('gymnasium', 'gym'):
    try:
      mod = __import__(modname)
    except ImportError:
      continue
    for env_id in IDS:
      try:
        mod.register(id=env_id, vector_entry_point=lambda num_envs=1, _id=env_id, **kw: make(_id, num_envs, **kw))
        done.append((modname, env_id))
      except Exception:  # already registered, or an older API without vector entry points
        pass
  return done
