"""bench.py -- env steps/sec of a batched random-policy Crafter rollout on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one tick of every env of the batch: `num_envs=4096, area=(64,64), view=(9,9)` per GPU
(BASELINE.json configs[1]; configs[2] is 8 x 4096 with no collective on the step path, i.e. weak
scaling).  Actions come from a pre-generated synthetic (T, B) int32 tensor; finished episodes are
regenerated inside the step (natural reset rate, about 1 env in 170 per step).

Timing: W untimed steps, then K steps; each timed step is bracketed by CUDA events on the env's
stream with a 256 MiB L2 flush between steps (outside the events); ms_per_step is the mean of those
K device durations, max over ranks.  `e2e` is the same metric through `cr_step_host` with pinned
HOST buffers (H2D actions, D2H reward/done, stream sync, every step).  One JSON line on rank 0.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NUM_ENVS = 4096
METRIC = 'env_steps_per_sec_random_policy'
UNIT = 'env-steps/s'
# algorithmic bytes per env-step of the render kernel =
# W*H*3 obs store + 3*vx*(vy-rows) window read (u8 mat + u16 slot) + 256 B of player/entity records.
RENDER_BYTES_PER_ENV = 64 * 64 * 3 + 3 * 9 * 7 + 256


def workload_config(n_gpus):
  return {
      'workload': f'crafter random-policy rollout, num_envs={NUM_ENVS} per GPU, area=(64,64), '
                  f'view=(9,9), size=(64,64), length=10000, auto-reset (BASELINE.json configs[1])',
      'global_num_envs': NUM_ENVS * n_gpus, 'parallelism': f'env-batch sharded x{n_gpus}, '
      'no collective on the step path',
      'l2': 'flushed between timed steps (256 MiB memset outside the per-step CUDA events)',
  }


class ClockSampler:
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.lines, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
           '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc:
      time.sleep(0.15)
      self.proc.terminate()
    sm, mx, reasons = [], [], set()
    for line in self.lines:
      p = [x.strip() for x in line.split(',')]
      if len(p) < 9:
        continue
      try:
        sm.append(float(p[1])); mx.append(float(p[2]))
      except ValueError:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                          'sw_power_cap'), p[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_baseline(cores, budget_s=12.0):
  """The oracle port (oracle/crafter_oracle.c: the reference's algorithm restated in C, pinned to
  the unmodified reference by tests/golden) on the host cores: same workload shape, bounded sample."""
  import numpy as np
  from oracle import oracle_env
  n_env = min(NUM_ENVS, 32 * cores)
  batch = oracle_env.OracleBatch(n_env, cores, seed=0)
  batch.reset()
  rs = np.random.RandomState(0)
  for _ in range(3):
    batch.step(rs.randint(0, 17, n_env))
  t0, n = time.perf_counter(), 0
  while time.perf_counter() - t0 < budget_s:
    batch.step(rs.randint(0, 17, n_env))
    n += 1
  dt = time.perf_counter() - t0
  return {'value': n * n_env / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
          'sample': f'{n} batched ticks of {n_env} of the {NUM_ENVS} envs (step + render + '
                    f'reset-on-done), C oracle port, {cores} threads, {dt:.1f} s'}


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU implementation of the path on the host cores.  The
  reference is pure Python and cannot travel to the GPU box, so this times the oracle port."""
  if rank != 0:
    return
  import numpy as np
  from oracle import oracle_env
  cores = os.cpu_count() or 1
  n_env = min(NUM_ENVS, 32 * cores)
  batch = oracle_env.OracleBatch(n_env, cores, seed=0)
  batch.reset()
  rs = np.random.RandomState(0)
  for _ in range(args.warmup):
    batch.step(rs.randint(0, 17, n_env))
  t0 = time.perf_counter()
  for _ in range(args.steps):
    batch.step(rs.randint(0, 17, n_env))
  dt = time.perf_counter() - t0
  value = args.steps * n_env / dt
  sample = (f'each step = one tick of {n_env} of the {NUM_ENVS} envs (bounded sample), oracle C port '
            f'of the reference env on {cores} host threads')
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32/f64',
      'data': 'synthetic', 'config': workload_config(args.gpus),
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                       'sample': sample},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=2000)
  ap.add_argument('--warmup', type=int, default=200)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  rank = int(os.environ.get('RANK', 0))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  if args.impl == 'reference':
    return run_reference(args, rank, world)

  import numpy as np
  import torch
  import torch.distributed as dist
  import crafter_b200

  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)
  assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)'

  B, K, W = NUM_ENVS, args.steps, args.warmup
  env = crafter_b200.Env(num_envs=B, seed=0, auto_reset=True, env_offset=rank * B, device=device)
  T = 512
  gen = torch.Generator(device=device).manual_seed(1234 + rank)
  actions = torch.randint(0, 17, (T, B), generator=gen, device=device, dtype=torch.int32)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
  stream = env._stream
  env.reset()
  for t in range(W):
    env.step(actions[t % T])
  torch.cuda.synchronize(device)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(device)

  # ---- timed region: K steps, per-step events on the env's stream, L2 flushed in between --------
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
  ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
  sampler = ClockSampler(local_rank)
  launches0 = env.launch_count
  barrier()
  sampler.start()
  wall0 = time.perf_counter()
  for k in range(K):
    env._actions.copy_(actions[(W + k) % T])
    flush.zero_()
    stream.wait_stream(torch.cuda.current_stream(device))
    starts[k].record(stream)
    env.step(env.actions_buffer)
    ends[k].record(stream)
  barrier()
  wall = time.perf_counter() - wall0
  clocks = sampler.stop()
  launches = env.launch_count - launches0
  step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
  total_ms = float(sum(step_ms))

  # back-to-back (warm L2, graph launches pipelined) for comparison
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(stream)
  for k in range(K):
    env.step(actions[k % T])
  e1.record(stream)
  barrier()
  warm_ms = e0.elapsed_time(e1)

  # ---- end to end through cr_step_host with pinned host buffers ---------------------------------
  h_actions = torch.randint(0, 17, (T, B), dtype=torch.int32).pin_memory()
  h_reward = torch.zeros(B, dtype=torch.float32).pin_memory()
  h_done = torch.zeros(B, dtype=torch.bool).pin_memory()
  for k in range(min(W, 20)):
    env.step_host(h_actions[k % T], h_reward, h_done)
  barrier()
  t0 = time.perf_counter()
  for k in range(K):
    env.step_host(h_actions[k % T], h_reward, h_done)
  barrier()
  e2e_s = time.perf_counter() - t0

  # ---- render kernel alone (the north-star roofline kernel), events on its launch stream --------
  R = 50
  r0 = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
  r1 = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
  for k in range(R):
    flush.zero_()
    stream.wait_stream(torch.cuda.current_stream(device))
    r0[k].record(stream)
    crafter_b200.env._cabi.check(env._lib.cr_render(env._handle, env._obs.data_ptr(), stream.cuda_stream))
    r1[k].record(stream)
  barrier()
  render_ms = sorted(a.elapsed_time(b) for a, b in zip(r0, r1))
  render_ms_avg = float(sum(render_ms) / len(render_ms))

  # ---- the same call with the observation batch copied to pinned host memory too (what a host-side
  # learner that consumes pixels pays: B*64*64*3 bytes over PCIe every step); reported beside e2e.
  K_obs, e2e_obs_s = min(K, 200), -1.0
  try:
    h_obs = torch.empty(B, 64, 64, 3, dtype=torch.uint8).pin_memory()
    for k in range(5):
      env.step_host(h_actions[k % T], h_reward, h_done, h_obs)
    barrier()
    t0 = time.perf_counter()
    for k in range(K_obs):
      env.step_host(h_actions[k % T], h_reward, h_done, h_obs)
    barrier()
    e2e_obs_s = time.perf_counter() - t0
  except Exception as exc:  # reported as null; every other number is already measured
    print(f'e2e_obs_to_host skipped: {exc!r}', file=sys.stderr)

  stats = torch.tensor([total_ms, warm_ms, e2e_s * 1e3, render_ms_avg, e2e_obs_s * 1e3],
                       dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
  total_ms, warm_ms, e2e_ms, render_ms_avg, e2e_obs_ms = stats.tolist()

  if rank == 0:
    peaks_path = ROOT / 'MEASURED_PEAKS.json'
    if peaks_path.exists():
      peak, peak_src = float(json.loads(peaks_path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    else:
      peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    algo_bytes = RENDER_BYTES_PER_ENV * B
    achieved = algo_bytes / (render_ms_avg * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = ROOT / 'profiles' / 'render_traffic.json'
    if tpath.exists():
      tj = json.loads(tpath.read_text())
      traffic = tj.get('dram_bytes_per_launch')
      traffic_src = tj.get('source', 'profiles/render_traffic.json (last ncu --set full capture of k_render)')
    out = {
        'metric': METRIC, 'value': world * B * K / (total_ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': K, 'warmup': W, 'ms_per_step': total_ms / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32/f64', 'data': 'synthetic',
        'config': workload_config(world),
        'value_warm_l2': world * B * K / (warm_ms * 1e-3), 'ms_per_step_warm_l2': warm_ms / K,
        'clocks': clocks,
        'e2e': {'value': world * B * K / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': 4 * B,
                'd2h_bytes_per_step': 5 * B, 'api': 'Env.step_host -> cr_step_host (pinned host '
                'actions in, reward+done out, stream sync per step); obs stays in HBM'},
        'e2e_obs_to_host': None if e2e_obs_ms <= 0 else {
            'value': world * B * K_obs / (e2e_obs_ms * 1e-3), 'unit': UNIT, 'steps': K_obs,
            'h2d_bytes_per_step': 4 * B, 'd2h_bytes_per_step': 5 * B + B * 64 * 64 * 3,
            'api': 'cr_step_host with obs_host: the observation batch is copied to pinned host memory too'},
        'gpu_launches': launches,
        'roofline': {'kernel': 'k_render', 'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                     'peak_source': peak_src, 'algorithmic_bytes_per_launch': algo_bytes,
                     'ms_per_launch': render_ms_avg,
                     'how': f'{R} stand-alone k_render launches over the {B} envs, CUDA events on the '
                            'launch stream, L2 flushed before each'},
        'wall_s_timed_region': wall,
    }
    if not args.no_cpu_baseline and world == 1:
      out['cpu_baseline'] = cpu_baseline(os.cpu_count() or 1)
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
