"""Timeline of k_post at steady state (profiling aid cr_debug_trace: %globaltimer stamps at the phase
boundaries of env_balance): how long the phases of one env's balance take and when the CTAs start."""
import ctypes
import os
import pathlib
import sys

os.environ.setdefault('CRAFTER_B200_LIB', str(pathlib.Path(__file__).resolve().parents[1] / 'crafter_b200/_lib/variants/libcrafter_b200_trace.so'))

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import crafter_b200  # noqa: E402

B = 4096
env = crafter_b200.Env(num_envs=B, seed=0, auto_reset=True)
lib = env._lib
lib.cr_debug_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
gen = torch.Generator(device='cuda').manual_seed(1234)
actions = torch.randint(0, 17, (256, B), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
for t in range(1000):
  env.step(actions[t % 256])
for rep in range(3):
  lib.cr_debug_trace(1, None, 0)
  env.step(actions[rep])
  torch.cuda.synchronize()
  lib.cr_debug_trace(0, None, 0)
  buf = np.zeros((3 * 4096, 8), np.int64)
  lib.cr_debug_trace(0, buf.ctypes.data, buf.size)
  envs = buf[:4096]
  rows = envs[envs[:, 5] > 0]
  ctas = buf[4096:8192, 0]
  ctas = ctas[ctas > 0]
  t0 = ctas.min()
  q = lambda a: ' '.join(f'{np.percentile(a, p) / 1e3:6.1f}' for p in (0, 10, 50, 90, 100))
  print(f'--- step {rep}: {len(rows)} envs balanced by {len(ctas)} CTAs (us; percentiles 0 10 50 90 100)')
  print('CTA start              ', q(ctas - t0))
  print('env start (list entry) ', q(rows[:, 0] - t0))
  names = ['state loaded', 'census', 'decided', 'resolved+scan', 'emitted+written']
  for k, name in enumerate(names):
    print(f'  {name:20s}', q(rows[:, k + 1] - rows[:, k]))
  print('env total              ', q(rows[:, 5] - rows[:, 0]), '   last end', f'{(rows[:, 5].max() - t0) / 1e3:.1f}')
  ticks = buf[8192:]
  ticks = ticks[ticks[:, 4] > 0]
  u0 = ticks[:, 0].min()
  print(f'k_update: {len(ticks)} ticks; start', q(ticks[:, 0] - u0), ' end', q(ticks[:, 4] - u0))
  for k, name in enumerate(['state + slots loaded', 'player', 'entities', 'reward / write-back']):
    print(f'  {name:22s}', q(ticks[:, k + 1] - ticks[:, k]))
  n_upd, n_slots = ticks[:, 5], ticks[:, 6]
  print('  in-radius entities    ', ' '.join(f'{np.percentile(n_upd, p):6.0f}' for p in (0, 10, 50, 90, 100)), '  slots', ' '.join(f'{np.percentile(n_slots, p):4.0f}' for p in (50, 100)))
  dur = (ticks[:, 3] - ticks[:, 2]) / 1e3
  print('  us per entity update   %.2f (least squares), tick total' % (np.polyfit(n_upd, dur, 1)[0]), q(ticks[:, 4] - ticks[:, 0]))
