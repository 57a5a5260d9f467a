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
