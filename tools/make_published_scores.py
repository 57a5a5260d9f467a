"""Summarise the reference's PUBLISHED random-agent runs (/root/reference/scores/
crafter_noreward-random.json: 5 runs of 1M steps recorded by the authors with the real environment,
real `opensimplex` included) into tests/golden/published_random_agent.json: episodes, episode-length
moments and, per achievement, the number of episodes that unlocked it.

The only golden data the reference holds for this path are distributions, not values; they pin what
no bit-exact fixture made in this container can: that worlds generated through the RESTATED noise
look to an agent like the worlds of the real package (tests/test_distribution.py).
"""
import json
import pathlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
SRC = pathlib.Path('/root/reference/scores/crafter_noreward-random.json')

runs = json.loads(SRC.read_text())
lengths = np.concatenate([np.array(r['length']) for r in runs])
names = sorted(k[len('achievement_'):] for k in runs[0] if k.startswith('achievement_'))
out = dict(
    source='danijar/crafter scores/crafter_noreward-random.json (5 seeds x 1M steps, uniform random actions)',
    episodes=int(len(lengths)), length_mean=float(lengths.mean()), length_var=float(lengths.var()),
    length_quantiles={str(q): float(np.quantile(lengths, q)) for q in (0.1, 0.25, 0.5, 0.75, 0.9, 0.99)},
    unlocked={a: int(sum((np.array(r['achievement_' + a]) >= 1).sum() for r in runs)) for a in names})
path = ROOT / 'tests' / 'golden' / 'published_random_agent.json'
path.write_text(json.dumps(out, indent=1))
print(path, out['episodes'], 'episodes, mean length %.2f' % out['length_mean'])
