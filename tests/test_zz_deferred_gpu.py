"""`-m gpu`: the experimental step schedules (CRAFTER_B200_DEFER_WG=1 deferred world generation,
CRAFTER_B200_SPLIT=1 early / late render, CRAFTER_B200_FUSED=1 tick + balance + frame of an env in one CTA; all
default OFF -- plus the A/B fallbacks CRAFTER_B200_DRAW_PREFETCH=0 / CRAFTER_B200_INCR_CENSUS=0) on the real CUDA library.  The schedule was written in a container without a
GPU (the host-sim replays of tests.test_schedule_knobs.py cover its logic, not its streams and
graph), so until its first hardware run is recorded under profiles/ this test is allowed to fail
(xfail, non-strict) and runs in a subprocess, last in the suite: a fault in the opt-in schedule
cannot take the product's own GPU tests with it."""
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]

CODE = r'''
import functools, os, sys
import numpy as np
import torch
sys.path.insert(0, os.getcwd())
import crafter_b200
from tests import parity
from tests.golden_util import Fixture
from tests.test_schedule_knobs import check_against_oracle
from tests.test_gpu_parity import HostStepEnv

knobs = {k: os.environ.get(k) for k in ('CRAFTER_B200_DEFER_WG', 'CRAFTER_B200_SPLIT')}
to_numpy = lambda x: x.detach().cpu().numpy()
env = parity.replay(Fixture('default_short'), crafter_b200.Env, auto_reset=True)
assert ('next_mat2' in env.state) == (knobs['CRAFTER_B200_DEFER_WG'] == '1')
parity.replay(Fixture('default_random'), crafter_b200.Env, auto_reset=True, steps=150)
parity.replay(Fixture('default_short'), crafter_b200.Env, auto_reset=False)
parity.replay(Fixture('default_short'), HostStepEnv, auto_reset=True)
for length in (1, 3):
  check_against_oracle(crafter_b200.Env, to_numpy, length, steps=10)
# a larger batch for a while: many refills in flight beside the tick; the default schedule must agree
def rollout():
  e = crafter_b200.Env(num_envs=512, seed=5, length=40, auto_reset=True)
  e.reset()
  g = torch.Generator(device='cuda').manual_seed(1)
  a = torch.randint(0, 17, (100, 512), generator=g, device='cuda', dtype=torch.int32)
  acc = torch.zeros((), dtype=torch.int64, device='cuda')
  for t in range(100):
    obs, reward, done, info = e.step(a[t])
    acc += obs.to(torch.int64).sum() + (reward * 10).round().to(torch.int64).sum() + done.sum()
  return int(acc), e.state['pstate'].clone()
a1, p1 = rollout()
os.environ['CRAFTER_B200_DEFER_WG'] = '0'
os.environ['CRAFTER_B200_SPLIT'] = '0'
os.environ.pop('CRAFTER_B200_DRAW_PREFETCH', None)
os.environ['CRAFTER_B200_FUSED'] = '0'
os.environ.pop('CRAFTER_B200_INCR_CENSUS', None)
a0, p0 = rollout()
assert a0 == a1 and torch.equal(p0, p1), (a0, a1)
print('deferred ok')
'''


@pytest.mark.gpu
@pytest.mark.xfail(reason='experimental opt-in schedules (default off), first hardware run '
                          'pending; the default schedule is covered by tests/test_gpu_parity.py',
                   strict=False)
@pytest.mark.parametrize('knobs', [
    dict(CRAFTER_B200_DEFER_WG='1'), dict(CRAFTER_B200_SPLIT='1'),
    dict(CRAFTER_B200_DEFER_WG='1', CRAFTER_B200_SPLIT='1'),
    dict(CRAFTER_B200_DEFER_WG='1', CRAFTER_B200_FUSED='1'),
    dict(CRAFTER_B200_DEFER_WG='1', CRAFTER_B200_FUSED='2'),
    dict(CRAFTER_B200_INCR_CENSUS='0', CRAFTER_B200_DRAW_PREFETCH='0')],
    ids=['defer', 'split', 'defer+split', 'defer+fused', 'defer+fused_one_launch', 'plain_tick'])
def test_cuda_experimental_schedules_in_subprocess(knobs):
  out = subprocess.run([sys.executable, '-c', CODE], env=dict(os.environ, **knobs),
                       capture_output=True, text=True, timeout=420, cwd=str(ROOT))
  assert out.returncode == 0 and 'deferred ok' in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
