"""ORACLE shim for the PyPI `opensimplex` module (reference crafter/worldgen.py:4,11,84-87),
backed by the C restatement oracle/opensimplex_ref.c (parity with the package: seeding pinned by two
published values, the 3-D arithmetic statistically by the published random-agent runs)."""
import ctypes

import numpy as np

from oracle import build as _build

_lib = None


def _get():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(str(_build.ensure()))
    _lib.osn_noise3.restype = ctypes.c_double
    _lib.osn_noise3.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    _lib.osn_init.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
  return _lib


class OpenSimplex:

  def __init__(self, seed=0):
    self._perm = np.zeros(256, np.int16)
    self._pgi = np.zeros(256, np.int16)
    _get().osn_init(int(seed), self._perm.ctypes.data, self._pgi.ctypes.data)
    self._p, self._g = self._perm.ctypes.data, self._pgi.ctypes.data

  def noise3(self, x, y, z):
    return _get().osn_noise3(self._p, self._g, float(x), float(y), float(z))
