"""CPU-side checks of crafter_b200.vector (no GPU: the env itself is exercised under -m gpu): the space
stand-ins used when gymnasium is not installed, and `register()` against a minimal stand-in module with
gymnasium's register() signature (crafter/__init__.py:4-17 registers the same two ids with gym)."""
import sys
import types

import numpy as np

from crafter_b200 import vector


def test_space_stand_ins_without_gymnasium():
  single_obs, single_act, obs, act = vector._spaces(5, (64, 64, 3), 17)
  assert single_obs.shape == (64, 64, 3) and obs.shape == (5, 64, 64, 3) and single_obs.dtype == np.uint8
  assert single_act.n == 17 and act.shape == (5,)
  assert obs.contains(obs.sample()) and act.contains(act.sample()) and single_act.contains(single_act.sample())
  assert not act.contains(np.full(5, 17))


def test_register_adds_the_reference_ids(monkeypatch):
  calls = []
  fake = types.ModuleType('gymnasium')
  fake.register = lambda id, entry_point=None, vector_entry_point=None, max_episode_steps=None: calls.append(
      (id, entry_point, vector_entry_point, max_episode_steps))
  monkeypatch.setitem(sys.modules, 'gymnasium', fake)
  monkeypatch.delitem(sys.modules, 'gym', raising=False)
  done = vector.register()
  assert [d for d in done if d[0] == 'gymnasium'] == [('gymnasium', 'CrafterReward-v1'), ('gymnasium', 'CrafterNoReward-v1')]
  assert {c[0] for c in calls} == set(vector.IDS) and all(c[3] == 10000 and callable(c[1]) and callable(c[2]) for c in calls)
  assert vector.IDS['CrafterReward-v1'] == dict(reward=True) and vector.IDS['CrafterNoReward-v1'] == dict(reward=False)
