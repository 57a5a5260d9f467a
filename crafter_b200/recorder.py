"""Batched counterpart of the reference's `StatsRecorder` (crafter/recorder.py:28-66): one
`stats.jsonl` line per finished episode with the reference's keys -- `length`, `reward` (sum of
info['reward'] rounded to one decimal) and `achievement_<name>` counts -- so that
`analysis/read_metrics.py` of the reference can score GPU rollouts unchanged.

The per-episode sums are kept on the device by the step kernel (`ep_return`, `final_stats` in
csrc/cr_common.h), so this works with `auto_reset=True` too, where the terminal state is replaced
inside `step()`.  Reading `done` costs one small device-to-host copy per step.
"""
import json
import pathlib

from . import rules


class StatsRecorder:

  def __init__(self, env, directory, env_ids=True):
    self._env = env
    self._directory = pathlib.Path(directory).expanduser()
    self._directory.mkdir(exist_ok=True, parents=True)
    self._file = (self._directory / 'stats.jsonl').open('a')
    self._env_ids = env_ids
    self.episodes = 0

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return getattr(self._env, name)

  def reset(self, mask=None):
    return self._env.reset(mask)

  def step(self, actions):
    obs, reward, done, info = self._env.step(actions)
    finished = done.nonzero().flatten()
    if finished.numel():
      state = self._env.state
      stats = state['final_stats'][finished].cpu().numpy()
      returns = state['ep_return'][finished, 1].cpu().numpy()
      for k, env_id in enumerate(finished.tolist()):
        line = {'length': int(stats[k, 22]), 'reward': round(float(returns[k]), 1)}
        for i, name in enumerate(rules.ACHIEVEMENTS):
          line[f'achievement_{name}'] = int(stats[k, i])
        if self._env_ids:
          line['env'] = env_id + getattr(self._env, '_env_offset', 0)
        self._file.write(json.dumps(line) + '\n')
        self.episodes += 1
      self._file.flush()
    return obs, reward, done, info

  def close(self):
    self._file.close()
