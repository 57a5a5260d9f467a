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


def can_build():
  return os.path.exists(NVCC)


def source_hash():
  """Digest of everything the library is compiled from; the build embeds it (cr_source_hash) and
  _cabi.load compares, so a library older than its sources is never loaded silently."""
  import hashlib
  h = hashlib.sha256()
  for d in sorted((ROOT / 'csrc').glob('*')) + [ROOT.parent / 'include' / 'crafter_b200.h']:
    h.update(d.name.encode())
    h.update(d.read_bytes())
  return h.hexdigest()[:16]


def needs_build():
  if not OUT.exists():
    return True
  import ctypes
  try:
    lib = ctypes.CDLL(str(OUT))
    lib.cr_source_hash.restype = ctypes.c_char_p
    return lib.cr_source_hash().decode() != source_hash()
  except (OSError, AttributeError):
    return True


# Build-time knobs for A/B runs (tools/ab_knobs.py with CRAFTER_B200_LIB=<variant>): name -> defines
VARIANTS = {
    'upd2': ['-DCR_UPDATE_WPB=2'], 'upd8': ['-DCR_UPDATE_WPB=8'],
    'bal256': ['-DCR_BALANCE_THREADS=256'],
    'memset_first': ['-DCR_MEMSET_FIRST'],
    'trace': ['-DCR_TRACE'],  # phase stamps of the balance (tools/balance_trace.py)
    # cells per k_wg_mat tile below the thread count: fewer octave items per thread and round
    'wgc192': ['-DCR_WG_TILE=192'], 'wgc128': ['-DCR_WG_TILE=128'], 'wgc64': ['-DCR_WG_TILE=64'],
    'wgc64t128': ['-DCR_WG_TILE=64', '-DCR_WG_THREADS=128', '-DCR_WG_MIN_CTAS=6'],
    'wgc32t128': ['-DCR_WG_TILE=32', '-DCR_WG_THREADS=128', '-DCR_WG_MIN_CTAS=6'],
    'obj256': ['-DCR_OBJ_THREADS=256'],
    'wg3': ['-DCR_WG_MIN_CTAS=3'], 'wg5': ['-DCR_WG_MIN_CTAS=5'], 'wgt128': ['-DCR_WG_TILE=128', '-DCR_WG_THREADS=128', '-DCR_WG_MIN_CTAS=6'],
}


def build_variant(name):
  """crafter_b200/_lib/variants/libcrafter_b200_<name>.so (in-tree, so it travels with gpurun)."""
  out = OUT.parent / 'variants' / f'libcrafter_b200_{name}.so'
  out.parent.mkdir(parents=True, exist_ok=True)
  res = subprocess.run([NVCC] + FLAGS + VARIANTS[name] + [f'-DCR_SOURCE_HASH="{source_hash()}"', '-o', str(out), str(SRC)],
                       capture_output=True, text=True)
  if res.returncode:
    print(res.stderr)
    raise RuntimeError(f'nvcc failed for variant {name}')
  return out


def build(force=False, verbose=False):
  if not force and not needs_build():
    return OUT
  OUT.parent.mkdir(exist_ok=True)
  cmd = [NVCC] + FLAGS + [f'-DCR_SOURCE_HASH="{source_hash()}"', '-o', str(OUT), str(SRC)]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or res.returncode:
    print(res.stdout)
    print(res.stderr)
  if res.returncode:
    raise RuntimeError('nvcc failed')
  (OUT.parent / 'ptxas.log').write_text(res.stderr)
  return OUT


if __name__ == '__main__':
  import sys
  if len(sys.argv) > 1 and sys.argv[1] == 'variants':
    for name in (sys.argv[2:] or VARIANTS):
      print(build_variant(name))
  else:
    print(build(force=True, verbose=True))
