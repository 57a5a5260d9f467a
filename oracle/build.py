"""ORACLE build recipe: gcc -> oracle/_build/liboracle.so (git-ignored, travels with gpurun)."""
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
OUT = ROOT / '_build' / 'liboracle.so'
SOURCES = ['opensimplex_ref.c', 'crafter_oracle.c']
FLAGS = ['-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-std=c11', '-Wall']


def build(force=False):
  srcs = [ROOT / s for s in SOURCES if (ROOT / s).exists()]
  if not force and OUT.exists() and all(OUT.stat().st_mtime >= s.stat().st_mtime for s in srcs):
    return OUT
  OUT.parent.mkdir(exist_ok=True)
  cmd = ['gcc'] + FLAGS + ['-o', str(OUT)] + [str(s) for s in srcs] + ['-lm', '-lpthread']
  subprocess.run(cmd, check=True)
  return OUT


def ensure():
  return build()


if __name__ == '__main__':
  print(build(force=True))
