"""Per-kernel device times of the step INSIDE its graph (CRAFTER_B200_TIMING=2: event-record nodes
around every kernel; =1 as first argument: eager launches instead)."""
import ctypes
import os
import pathlib
import sys

os.environ.setdefault('CRAFTER_B200_TIMING', '2')
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import crafter_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
area = int(sys.argv[2]) if len(sys.argv) > 2 else 64
view = int(sys.argv[3]) if len(sys.argv) > 3 else 9
size = int(sys.argv[4]) if len(sys.argv) > 4 else 64
env = crafter_b200.Env(num_envs=B, seed=0, auto_reset=True, area=(area, area), view=(view, view), size=(size, size))
gen = torch.Generator(device='cuda').manual_seed(1234)
actions = torch.randint(0, 17, (256, B), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
out = (ctypes.c_double * 8)()
names = ['update', 'install', 'render', 'seed', 'wg_mat', 'wg_obj', 'seed_ahead', 'balance']
print('timing mode', os.environ['CRAFTER_B200_TIMING'])
t = 0
for phase, steps in (('steps 0-100 (day)', 100), ('steps 100-148', 48), ('steps 148-272 (night, first death wave)', 124),
                     ('steps 272-600', 328), ('steps 600-1600 (desynchronised)', 1000)):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for _ in range(steps):
    env.step(actions[t % 256]); t += 1
  n = env._lib.cr_timing(env._handle, out)
  print(f'{phase:45s} n={n:5d}  ' + '  '.join(f'{k}={1e3*out[i]:6.1f}us' for i, k in enumerate(names) if out[i] > 0))
