"""Timeline of ONE k_consume launch at steady state (CRAFTER_B200_TRACE=1: every CTA stamps %globaltimer at
start, when it holds its work item, and at its end): how long consumers wait, how long a frame / a
balance + frame / an install + frame takes, and when the last of each ends."""
import os
import pathlib
import sys

os.environ['CRAFTER_B200_TRACE'] = '1'
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import crafter_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = crafter_b200.Env(num_envs=B, seed=0, auto_reset=True)
gen = torch.Generator(device='cuda').manual_seed(1234)
actions = torch.randint(0, 17, (256, B), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
for t in range(1000):
  env.step(actions[t % 256])
for rep in range(3):
  env.step(actions[rep])
  torch.cuda.synchronize()
  tr = env.state['trace'].cpu().numpy()
  t0 = tr[:, 0].min()
  start, item, end, word = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3, tr[:, 3]
  tick = word == -1  # (none: k_update's CTAs are not traced)
  kind = np.where(tick, -1, (word >> 24) - 1)
  q = lambda a: ' '.join(f'{np.percentile(a, p):7.1f}' for p in (0, 10, 50, 90, 100)) if len(a) else '-'
  print(f'--- launch {rep}: k_consume spans {end.max():.1f} us from its first CTA, {(~tick).sum()} CTAs   (percentiles 0 10 50 90 100, us)')
  for k, name in ((0, 'frame'), (1, 'balance+frame'), (2, 'install+frame'), (3, 'install(+balance)')):
    m = kind == k
    if m.any():
      print(f'{name:18s} n={m.sum():5d} start', q(start[m]), ' | wait', q((item - start)[m]), ' | work', q((end - item)[m]), ' | end', q(end[m]))
  order = np.argsort(start)
  # how many CTAs are alive over time
  ts = np.linspace(0, end.max(), 12)
  alive = [(int(((start <= x) & (end > x)).sum()), int(((start <= x) & (end > x) & tick).sum())) for x in ts]
  print('alive CTAs (all, ticking) at', ' '.join(f'{x:.0f}us:{a}/{b}' for x, (a, b) in zip(ts, alive)))
