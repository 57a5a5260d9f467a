"""Device-timed env-steps/s of the other BASELINE.json configs (parity-test cases, not bench lines)."""
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import crafter_b200  # noqa: E402

CONFIGS = {
    'configs[1] B=4096 area=64 view=9 size=64': dict(num_envs=4096),
    'configs[3] B=1024 area=256 view=9 size=64': dict(num_envs=1024, area=(256, 256)),
    'configs[4] B=4096 area=64 view=15 size=128': dict(num_envs=4096, view=(15, 15), size=(128, 128)),
    'B=16384 area=64 view=9 size=64 (larger batch)': dict(num_envs=16384),
}
for name, kw in CONFIGS.items():
  env = crafter_b200.Env(seed=0, auto_reset=True, **kw)
  B = env.num_envs
  actions = torch.randint(0, 17, (64, B), device='cuda', dtype=torch.int32)
  t0 = time.time(); env.reset(); torch.cuda.synchronize(); reset_s = time.time() - t0
  for t in range(300):
    env.step(actions[t % 64])
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  K = 500
  e0.record(env._stream)
  for t in range(K):
    env.step(actions[t % 64])
  e1.record(env._stream)
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  err = int(env.state['pstate'][:, 14].abs().sum())
  print(json.dumps({'config': name, 'env_steps_per_s': B * K / (ms * 1e-3), 'ms_per_step': ms / K,
                    'reset_all_s': reset_s, 'slot_overflow_flags': err}))
  env.close()
  del env
  torch.cuda.empty_cache()
