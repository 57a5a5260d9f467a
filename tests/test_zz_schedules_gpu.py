"""`-m gpu`: every knob combination of the CUDA library's step graph, each in its own process (the knobs
are read at cr_create / import): the default, the generic instantiations, the A/B fallbacks of
the tick (CRAFTER_B200_DRAW_PREFETCH=0 / CRAFTER_B200_INCR_CENSUS=0), no L2 hint on the obs rows, eager
launches instead of the graph.  Each process replays reference-recorded fixtures with and without
auto-reset, back-to-back resets (length 1 / 3), the terminal frames (final_obs), explicit reset(mask)
between auto-resets, and a 512-env rollout that must agree bit for bit with the default."""
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
KNOBS = ('CRAFTER_B200_OBS_EVICT_FIRST', 'CRAFTER_B200_DRAW_PREFETCH', 'CRAFTER_B200_INCR_CENSUS',
         'CRAFTER_B200_NO_SPECIALIZE', 'CRAFTER_B200_NO_GRAPH', 'CRAFTER_B200_FRAME_ORDER',
         'CRAFTER_B200_VIEW_AHEAD')

CODE = r'''
import functools, os, sys
import numpy as np
import torch
sys.path.insert(0, os.getcwd())
import crafter_b200
from tests import parity
from tests.golden_util import Fixture
from tests.test_schedule_knobs import check_against_oracle, check_terminal_frames, check_mixed_resets_and_masks
from tests.test_gpu_parity import HostStepEnv

to_numpy = lambda x: x.detach().cpu().numpy()
parity.replay(Fixture('default_short'), crafter_b200.Env, auto_reset=True)
parity.replay(Fixture('default_random'), crafter_b200.Env, auto_reset=True, steps=150)
parity.replay(Fixture('default_short'), crafter_b200.Env, auto_reset=False)
parity.replay(Fixture('default_short'), HostStepEnv, auto_reset=True)
for length in (1, 3):
  check_against_oracle(crafter_b200.Env, to_numpy, length, steps=10)
for length in (1, 10):
  assert check_terminal_frames(crafter_b200.Env, to_numpy, length, steps=31) >= 9

class NumpyEnv(crafter_b200.Env):  # the CPU helpers hand numpy masks / actions in and read arrays out
  def reset(self, mask=None):
    return to_numpy(super().reset(mask))
  def step(self, actions):
    o, r, d, i = super().step(torch.as_tensor(actions, device='cuda'))
    return to_numpy(o), to_numpy(r), to_numpy(d)
check_mixed_resets_and_masks(NumpyEnv)

# a larger batch for a while; the default must agree
def rollout():
  e = crafter_b200.Env(num_envs=512, seed=5, length=40, auto_reset=True)
  e.reset()
  g = torch.Generator(device='cuda').manual_seed(1)
  a = torch.randint(0, 17, (100, 512), generator=g, device='cuda', dtype=torch.int32)
  acc = torch.zeros((), dtype=torch.int64, device='cuda')
  for t in range(100):
    obs, reward, done, info = e.step(a[t])
    acc += obs.to(torch.int64).sum() + (reward * 10).round().to(torch.int64).sum() + done.sum()
  return int(acc), e.state_dict()['pstate']
a1, p1 = rollout()
for k in KNOBS:
  os.environ.pop(k, None)
a0, p0 = rollout()
assert a0 == a1 and torch.equal(p0, p1), (a0, a1)
print('schedule ok')
'''.replace('KNOBS', repr(KNOBS))


@pytest.mark.gpu
@pytest.mark.parametrize('knobs', [
    dict(), dict(CRAFTER_B200_NO_SPECIALIZE='1'),
    dict(CRAFTER_B200_INCR_CENSUS='0', CRAFTER_B200_DRAW_PREFETCH='0', CRAFTER_B200_OBS_EVICT_FIRST='0'),
    dict(CRAFTER_B200_NO_GRAPH='1'), dict(CRAFTER_B200_FRAME_ORDER='0'), dict(CRAFTER_B200_VIEW_AHEAD='0')],
    ids=['default', 'generic', 'plain_tick', 'eager', 'env_order', 'no_view_ahead'])
def test_cuda_step_schedules_in_subprocess(knobs):
  out = subprocess.run([sys.executable, '-c', CODE], env=dict(os.environ, **knobs),
                       capture_output=True, text=True, timeout=420, cwd=str(ROOT))
  assert out.returncode == 0 and 'schedule ok' in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
