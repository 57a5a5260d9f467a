"""The C-ABI library builds for sm_100a, loads without a GPU and exports every entry point that
include/crafter_b200.h declares; the Python package itself refuses to run without CUDA."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def declared_functions():
  text = (ROOT / 'include' / 'crafter_b200.h').read_text()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(cr_[a-z_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  from crafter_b200 import build, _cabi
  lib = ctypes.CDLL(str(build.build()))
  names = declared_functions()
  assert set(_cabi.EXPORTS) == set(names), (names, _cabi.EXPORTS)
  for name in names:
    assert hasattr(lib, name), name
  lib.cr_abi_version.restype = ctypes.c_int
  assert lib.cr_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_the_header():
  """ctypes mirrors of cr_config / cr_tables / cr_state have the C sizes (LP64)."""
  from crafter_b200 import _cabi
  assert ctypes.sizeof(_cabi.CrConfig) == 16 * 4 + 2 * 8
  assert ctypes.sizeof(_cabi.CrTables) == 7 * 8
  header = (ROOT / 'include' / 'crafter_b200.h').read_text()
  fields = re.search(r'typedef struct cr_state \{(.*?)\} cr_state;', header, re.S).group(1)
  names = re.findall(r'\*\s*(\w+);', fields)
  assert names == [f[0] for f in _cabi.CrState._fields_]


def test_no_cpu_fallback():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  import crafter_b200
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    crafter_b200.Env(num_envs=2)


def test_product_never_imports_the_oracle():
  for path in (ROOT / 'crafter_b200').rglob('*.py'):
    text = path.read_text()
    assert 'import oracle' not in text and 'from oracle' not in text, path
    assert 'hostsim' not in text, path
