"""crafter_b200: a B200-native, batched implementation of the Crafter environment.

    import crafter_b200
    env = crafter_b200.Env(num_envs=4096, seed=0, auto_reset=True)
    obs = env.reset()                                  # (4096, 64, 64, 3) uint8 on cuda
    obs, reward, done, info = env.step(actions)        # actions: int tensor (4096,)

Same constructor arguments, action / item / achievement orders, rules and rendering as
danijar/crafter's `crafter.Env`; simulation and rendering run in CUDA kernels for sm_100a.
"""
from .rules import ACTIONS, ACHIEVEMENTS, ITEMS, MATERIALS  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require CUDA (docs, CPU-side tests)
  if name == 'Env':
    from .env import Env
    return Env
  if name == 'ShardedEnv':
    from .sharded import ShardedEnv
    return ShardedEnv
  raise AttributeError(name)
