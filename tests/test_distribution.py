"""Distribution-level parity with the reference's PUBLISHED random-agent runs.

The reference holds no value-level golden data for this path, but it does publish what its own
environment -- real `opensimplex` included -- does to a uniform random agent: 29,784 episodes over
5 runs of 1M steps (scores/crafter_noreward-random.json, summarised in
tests/golden/published_random_agent.json by tools/make_published_scores.py).  Episode length and the
unlock rate of every achievement depend on what the generated worlds look like around the player
(trees, water, grass, cows, zombies at night ...), so they test the one boundary no bit-exact fixture
made in this container can reach: that terrain produced through the RESTATED 3-D noise is the
terrain of the real package, statistically.  Unbiased sampling: a fixed number of whole episodes per
env, never "the episodes that happened to finish".

Two-sample z tests at |z| < 4.5 (22 rates + the mean length; a false alarm is a < 1e-4 event), and
achievements the published runs never unlocked may show up at most a handful of times."""
import json
import pathlib

import numpy as np
import pytest

from crafter_b200 import rules

PUBLISHED = json.loads((pathlib.Path(__file__).resolve().parent / 'golden' / 'published_random_agent.json').read_text())


def compare(lengths, unlocked):
  """lengths [n], unlocked bool [n, 22] of n whole episodes against the published summary."""
  n, N = len(lengths), PUBLISHED['episodes']
  assert n >= 1500
  z_len = (lengths.mean() - PUBLISHED['length_mean']) / np.sqrt(lengths.var() / n + PUBLISHED['length_var'] / N)
  report = {'length': (PUBLISHED['length_mean'], float(lengths.mean()), float(z_len))}
  assert abs(z_len) < 4.5, report
  for k, name in enumerate(rules.ACHIEVEMENTS):
    p, q = PUBLISHED['unlocked'][name] / N, float(unlocked[:, k].mean())
    if PUBLISHED['unlocked'][name] < 50:  # (almost) never unlocked by a random agent
      assert unlocked[:, k].sum() <= 8 + 10 * n * max(p, 1e-4), (name, int(unlocked[:, k].sum()))
      continue
    z = (q - p) / np.sqrt(p * (1 - p) * (1 / n + 1 / N))
    report[name] = (100 * p, 100 * q, float(z))
    assert abs(z) < 4.5, (name, report[name])
  for q_, want in PUBLISHED['length_quantiles'].items():  # the shape, not just the mean
    got = float(np.quantile(lengths, float(q_)))
    assert abs(got - want) <= max(4.0, 0.06 * want), ('length quantile', q_, want, got)
  return report


def test_oracle_random_agent_matches_published_scores():
  """The C oracle (the checker the CUDA path is compared with, bit for bit): 2,400 whole episodes."""
  from oracle import oracle_env
  rs = np.random.RandomState(0)
  lengths, unlocked = [], []
  for i in range(60):
    env = oracle_env.OracleEnv(seed=31000 + i, reward=False)
    for episode in range(40):
      env.reset()
      n = 0
      while True:
        n += 1
        if env.step_norender(int(rs.randint(0, 17)))[1]:
          break
      lengths.append(n)
      unlocked.append(env.export_state()['player'][16:38] >= 1)
  report = compare(np.array(lengths), np.array(unlocked))
  assert abs(report['length'][1] - 167.9) < 5


@pytest.mark.gpu
def test_cuda_random_agent_matches_published_scores():
  """The CUDA library: 4096 envs, the first 6 whole episodes of each (24,576 episodes)."""
  import torch
  import crafter_b200
  B, M = 4096, 6
  env = crafter_b200.Env(num_envs=B, seed=52000, auto_reset=True, reward=False)
  env.reset()
  gen = torch.Generator(device='cuda').manual_seed(7)
  count = torch.zeros(B, dtype=torch.int64, device='cuda')
  lengths, unlocked = [], []
  for t in range(12000):
    actions = torch.randint(0, 17, (B,), generator=gen, device='cuda', dtype=torch.int32)
    obs, reward, done, info = env.step(actions)
    idx = torch.nonzero(done & (count < M)).flatten()
    if idx.numel():
      fs = env.state['final_stats'][idx]  # achievements[22], length, terminated flag of the episode that ended
      lengths.append(fs[:, 22].cpu().numpy())
      unlocked.append((fs[:, :22] >= 1).cpu().numpy())
    count += done.to(torch.int64)
    if t % 50 == 49 and bool((count >= M).all()):
      break
  assert bool((count >= M).all()), 'some env did not finish its episodes'
  lengths, unlocked = np.concatenate(lengths), np.concatenate(unlocked)
  assert len(lengths) == B * M
  compare(lengths, unlocked)
