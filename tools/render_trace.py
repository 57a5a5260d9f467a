"""Timeline of k_render (profiling aid cr_debug_trace, build variant `trace`: %globaltimer stamps of warps 0
and 1 of every frame CTA at the phase boundaries): alone (`env.render()`) and inside the step graph."""
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
q = lambda a: ' '.join(f'{np.percentile(a, p) / 1e3:6.2f}' for p in (0, 10, 50, 90, 100))
NAMES = ['launch -> stage done (own part)', 'wait at barrier 1', 'tile cache (own part)', 'wait at barrier 2',
         'assemble', 'fence + barrier + bulk store']
for what in ('alone', 'alone', 'in the step graph', 'in the step graph'):
  torch.cuda.synchronize()
  lib.cr_debug_trace(1, None, 0)
  if what == 'alone':
    env.render()
  else:
    env.step(actions[7])
  torch.cuda.synchronize()
  lib.cr_debug_trace(0, None, 0)
  buf = np.zeros((5 * 4096, 8), np.int64)
  lib.cr_debug_trace(0, buf.ctypes.data, buf.size)
  w0, w1 = buf[3 * 4096:4 * 4096], buf[4 * 4096:]
  ok = w0[:, 6] > 0
  t0 = w0[ok, 0].min()
  night = w0[:, 7] >= 1000
  print(f'=== k_render {what}: {ok.sum()} frames, {int((night & ok).sum())} at night; us, percentiles 0 10 50 90 100')
  print('CTA start                         ', q(w0[ok, 0] - t0), '   last CTA ends', f'{(w0[ok, 6].max() - t0) / 1e3:.1f}')
  for label, sel in (('day', ok & ~night), ('night', ok & night)):
    if not sel.any():
      continue
    print(f'-- {label} frames: CTA lifetime     ', q(w0[sel, 6] - w0[sel, 0]), '  tile jobs', ' '.join(f'{np.percentile(w0[sel, 7] % 1000, p):3.0f}' for p in (10, 50, 90)))
    for name, w in (('warp 0 (gathers, plans)', w0), ('warp 1', w1)):
      print(f'   {name}')
      for k, n in enumerate(NAMES):
        print(f'     {n:32s}', q(w[sel, k + 1] - w[sel, k]))
