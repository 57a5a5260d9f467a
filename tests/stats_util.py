"""Expected `stats.jsonl` lines of the reference StatsRecorder (recorder.py:53-66), derived from a
golden fixture: per episode the step count, round(sum of info['reward'], 1) and achievement counts."""
from crafter_b200 import rules


def expected_lines(fx, steps=None):
  T = fx.T if steps is None else min(steps, fx.T)
  events = []  # (t, env, line)
  for i in range(fx.K):
    length, total = 0, 0
    reward, done, player = fx.env(i, 'reward'), fx.env(i, 'done'), fx.env(i, 'player')
    for t in range(T):
      length += 1
      total += float(reward[t])
      if done[t]:
        line = {'length': length, 'reward': round(total, 1)}
        for j, name in enumerate(rules.ACHIEVEMENTS):
          line[f'achievement_{name}'] = int(player[t][16 + j])
        line['env'] = fx.seed0 * 0 + i
        events.append((t, i, line))
        length, total = 0, 0
  return [e[2] for e in sorted(events, key=lambda e: (e[0], e[1]))]


def record_and_check_episodes(fx, env, to_actions):
  """Drives a crafter_b200.recorder.EpisodeRecorder over the fixture's first episodes and compares
  every saved .npz with what the reference's EpisodeRecorder (recorder.py:125-152) would hold."""
  import zlib
  import numpy as np
  env.reset()
  actions = np.stack([fx.env(i, 'actions') for i in range(fx.K)], 1)
  first_done = [int(np.argmax(fx.env(i, 'done'))) for i in range(fx.K)]
  for t in range(max(first_done) + 1):
    obs, reward, done, info = env.step(to_actions(actions[t]))
    if bool(done.any()) and not getattr(env, '_auto', False):
      env.reset(done)  # the fixture resets on done too (an auto-resetting batch did it inside step())
  assert len(env.saved) >= fx.K
  for i in range(fx.K):
    path = [p for p in env.saved if f'-env{i}-' in p.name][0]
    n = first_done[i] + 1
    assert path.name.endswith(f'-len{n}.npz')
    ep = np.load(path)
    want = {'image', 'action', 'reward', 'done', 'discount', 'semantic', 'player_pos'}
    want |= {f'achievement_{k}' for k in rules.ACHIEVEMENTS} | {f'ainventory_{k}' for k in rules.ITEMS}
    assert set(ep.files) == want
    assert ep['image'].shape == (n + 1, *fx.size[::-1], 3) and ep['image'].dtype == np.uint8
    assert zlib.crc32(ep['image'][0].tobytes()) == fx.env(i, 'reset_obs_crc')[0]
    player = fx.env(i, 'player')
    for t in range(n):
      assert zlib.crc32(ep['image'][t + 1].tobytes()) == fx.env(i, 'obs_crc')[t]
      assert ep['action'][t + 1] == actions[t, i] and ep['reward'][t + 1] == fx.env(i, 'reward')[t]
      assert bool(ep['done'][t + 1]) == bool(fx.env(i, 'done')[t])
      assert ep['discount'][t + 1] == (0.0 if player[t][0] <= 0 else 1.0)
      assert list(ep['player_pos'][t + 1]) == list(player[t][45:47])
      for j, k in enumerate(rules.ACHIEVEMENTS):
        assert ep[f'achievement_{k}'][t + 1] == player[t][16 + j]
      for j, k in enumerate(rules.ITEMS):
        assert ep[f'ainventory_{k}'][t + 1] == player[t][j]
    assert ep['action'][0] == 0 and ep['reward'][0] == 0 and not ep['done'][0]
    assert ep['semantic'].shape == (n + 1, *fx.area) and ep['semantic'][0].sum() == 0
    if fx.has(i, f'semantic_{n - 1}'):
      assert (ep['semantic'][n] == fx.env(i, f'semantic_{n - 1}')).all()
