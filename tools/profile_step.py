"""Workload for ncu captures: reset + N random-policy steps of the BASELINE configs[1] batch.
Graphs are disabled (CRAFTER_B200_NO_GRAPH=1) so that every kernel is a plain launch for ncu."""
import argparse
import os
import pathlib
import sys

os.environ.setdefault('CRAFTER_B200_NO_GRAPH', '1')
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))

import torch  # noqa: E402
import crafter_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=200)
ap.add_argument('--num-envs', type=int, default=4096)
ap.add_argument('--area', type=int, default=64)
ap.add_argument('--view', type=int, default=9)
ap.add_argument('--size', type=int, default=64)
args = ap.parse_args()
env = crafter_b200.Env(num_envs=args.num_envs, seed=0, auto_reset=True, area=(args.area, args.area),
                       view=(args.view, args.view), size=(args.size, args.size))
gen = torch.Generator(device='cuda').manual_seed(1234)
actions = torch.randint(0, 17, (256, args.num_envs), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
for t in range(args.steps):
  env.step(actions[t % 256])
torch.cuda.synchronize()
print('done', env.launch_count, 'launches; steps', int(env.state['pstate'][:, 9].max()))
