"""Build recipe of the CUDA library: nvcc (sm_100a only) -> crafter_b200/_lib/libcrafter_b200.so.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  `-fmad=false`: the
reference arithmetic (numpy / PIL) has no fused multiply-adds and parity is bit-exact.
"""
import os
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
SRC = ROOT / 'csrc' / 'crafter_kernels.cu'
OUT = ROOT / '_lib' / 'libcrafter_b200.so'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-fmad=false', '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def needs_build():
  if not OUT.exists():
    return True
  deps = list((ROOT / 'csrc').glob('*')) + [ROOT.parent / 'include' / 'crafter_b200.h']
  return any(d.stat().st_mtime > OUT.stat().st_mtime for d in deps)


def build(force=False, verbose=False):
  if not force and not needs_build():
    return OUT
  OUT.parent.mkdir(exist_ok=True)
  cmd = [NVCC] + FLAGS + ['-o', str(OUT), str(SRC)]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or res.returncode:
    print(res.stdout)
    print(res.stderr)
  if res.returncode:
    raise RuntimeError('nvcc failed')
  (OUT.parent / 'ptxas.log').write_text(res.stderr)
  return OUT


if __name__ == '__main__':
  print(build(force=True, verbose=True))
