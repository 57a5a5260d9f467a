"""Pin the OpenSimplex restatement against the real PyPI package (run where `pip install
opensimplex` works; the build container has no network).  Writes tests/golden/noise_pypi.npz with
noise3 values on fixed (seed, x, y, z) lattices; tests/test_noise.py picks the file up when present.

    pip install opensimplex && python tools/make_noise_golden.py
"""
import pathlib

import numpy as np


def main():
  import opensimplex  # the real package, not oracle/shims
  assert 'shims' not in opensimplex.__file__, 'run without oracle/shims on sys.path'
  rs = np.random.RandomState(0)
  out = {}
  for seed in (0, 1, 42, 1256191933, 2 ** 31 - 2):
    gen = opensimplex.OpenSimplex(seed=seed)
    fn = gen.noise3 if hasattr(gen, 'noise3') else gen.noise3d
    pts = np.concatenate([rs.uniform(-40, 40, (2000, 3)), rs.randint(-8, 8, (200, 3)).astype(float),
                          np.stack(np.meshgrid(np.arange(16) / 3, np.arange(16) / 15, [0.0, 3.0, 8.0]), -1).reshape(-1, 3)])
    out[f'pts_{seed}'] = pts
    out[f'val_{seed}'] = np.array([fn(*p) for p in pts])
  path = pathlib.Path(__file__).resolve().parents[1] / 'tests' / 'golden' / 'noise_pypi.npz'
  np.savez_compressed(path, **out)
  print('wrote', path)


if __name__ == '__main__':
  main()
