"""Loader for tests/golden/*.npz (written by tools/make_golden.py from the unmodified reference)."""
import pathlib

import numpy as np

GOLDEN = pathlib.Path(__file__).resolve().parent / 'golden'
NAMES = sorted(p.stem for p in GOLDEN.glob('*.npz'))


class Fixture:

  def __init__(self, name):
    self.name = name
    self.z = np.load(GOLDEN / f'{name}.npz')
    g = lambda k: self.z['meta_' + k]
    self.area = tuple(int(v) for v in g('area'))
    self.view = tuple(int(v) for v in g('view'))
    self.size = tuple(int(v) for v in g('size'))
    self.length = int(g('length'))
    self.seed0, self.K, self.T = int(g('seed0')), int(g('K')), int(g('T'))
    self.boost = dict(zip([str(s) for s in g('boost_items')], [int(v) for v in g('boost_values')]))
    self.kwargs = dict(area=self.area, view=self.view, size=self.size, length=self.length)

  def env(self, i, key):
    return self.z[f'e{i}_{key}']

  def has(self, i, key):
    return f'e{i}_{key}' in self.z.files
