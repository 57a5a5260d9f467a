"""A/B of environment knobs: device-timed us/step of the bench workload, one subprocess per combination.

    python tools/ab_knobs.py - CRAFTER_B200_DRAW_PREFETCH=0 CRAFTER_B200_INCR_CENSUS=0 CRAFTER_B200_DEFER_WG=1 CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_SPLIT=1

Knobs: CRAFTER_B200_LIB=<other .so> (a build variant), CRAFTER_B200_NO_GRAPH=1, CRAFTER_B200_NO_SPECIALIZE=1,
CRAFTER_B200_OBS_EVICT_FIRST=0 (no L2 evict-first hint on the observation rows),
CRAFTER_B200_DRAW_PREFETCH=0 (switch OFF the tick's up-front keyed draws), CRAFTER_B200_INCR_CENSUS=0 (switch OFF the
maintained per-chunk grass / path counts: every balance tick re-counts the cells)."""
import os
import subprocess
import sys

CODE = r'''
import torch, crafter_b200
import os
kw = dict(default=dict(num_envs=4096), area256=dict(num_envs=1024, area=(256, 256)), view15=dict(num_envs=4096, view=(15, 15), size=(128, 128)))[os.environ.get('AB_CONFIG', 'default')]
env = crafter_b200.Env(seed=0, auto_reset=True, **kw)
a = torch.randint(0, 17, (64, env.num_envs), device='cuda', dtype=torch.int32)
env.reset()
for t in range(700): env.step(a[t % 64])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(env._stream)
  for t in range(1000): env.step(a[t % 64])
  e1.record(env._stream); torch.cuda.synchronize()
  best = min(best, e0.elapsed_time(e1) / 1000)
print(f'{best * 1000:.1f} us/step')
'''
# usage: tools/ab_knobs.py [K=V[,K=V...]] ...   (one run per argument; '-' = no knob)
COMBOS = [dict(kv.split('=', 1) for kv in arg.split(',') if '=' in kv) for arg in (sys.argv[1:] or ['-'])]
for combo in COMBOS:
  out = subprocess.run([sys.executable, '-c', CODE], env=dict(os.environ, **combo), capture_output=True, text=True)
  print(combo, out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
