"""What the opt-in whole-batch gather costs next to the step: per rank 4096 envs, the obs
all-gather moves (N - 1) x 50 MB in and 50 MB out per GPU and step.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gather_cost.py

Prints device-timed us per step for: the step alone; step + blocking ShardedEnv.gather(obs, reward, done);
step + ShardedEnv.gather_async(obs) waited for one step later (the collective overlaps the next step)."""
import json
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from crafter_b200.sharded import ShardedEnv  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
world, rank = dist.get_world_size(), dist.get_rank()
B = 4096
env = ShardedEnv(num_envs=B * world, seed=0, auto_reset=True)
gen = torch.Generator(device='cuda').manual_seed(rank)
actions = torch.randint(0, 17, (64, B), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
for t in range(600):
  env.step(actions[t % 64])


def timed(body, steps=300):
  for t in range(20):
    body(t)
  dist.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for t in range(steps):
    body(t)
  e1.record(); torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / steps], device='cuda')
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return float(ms) * 1e3


def plain(t):
  env.step(actions[t % 64])


def blocking(t):
  obs, reward, done, info = env.step(actions[t % 64])
  env.gather(obs, reward, done)


pending = [None]


def overlapped(t):
  obs, reward, done, info = env.step(actions[t % 64])
  if pending[0] is not None:
    pending[0].wait()  # last step's whole-batch obs is ready for the learner here
  pending[0] = env.gather_async(obs)


res = {'n_gpus': world, 'envs_per_gpu': B, 'obs_mb_per_gpu': B * 64 * 64 * 3 / 1e6, 'step_us': timed(plain),
       'step_plus_blocking_gather_us': timed(blocking), 'step_plus_overlapped_gather_us': timed(overlapped)}
if rank == 0:
  print(json.dumps(res))
dist.destroy_process_group()
