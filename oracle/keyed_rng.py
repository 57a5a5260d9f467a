"""ORACLE (test infrastructure): the keyed, counter-based random contract.

The reference shares ONE sequential ``np.random.RandomState`` between worldgen, every entity,
spawn balancing and the renderer (crafter/engine.py:34, objects.py:12, env.py:161,
engine.py:209) and is not reproducible against itself.  BASELINE.json's
north star therefore asks for parity "under matched counter-based seeds": every draw site is
addressed by (world_seed, domain, counters) through Philox4x32-10, on the reference (injected
through `oracle/ref_harness.py` without editing reference files), in the C oracle and in the
CUDA kernels.  This file is the Python statement of that contract.

    key      = (world_seed, domain)
    counter  = (k, c1, c2, c3)           k = index of the draw inside its context

    domain 0 SEED     ctr (0,0,0,0)            simplex seed = randint(2**31-1)  worldgen.py:11
    domain 1 WG_MAT   ctr (k, cell, 0, 0)      cell = x*H + y                   worldgen.py:43-58
    domain 2 WG_OBJ   ctr (k, cell, 0, 0)                                       worldgen.py:71-75
    domain 3 UPDATE   ctr (k, step, 0, 0)      k runs over the whole slot-ordered update loop
                                               objects.py:65,226,277,298-299,333-340
    domain 4 BALANCE  ctr (k, step, chunk, cls) chunk = (xmin//12)*ncy + ymin//12; cls 0 zombie,
                                               1 skeleton, 2 cow              env.py:165,169,175,176
    domain 5 NOISE    ctr (x>>2, step, y, 0)   canvas pixel (x, y) uses word x&3 engine.py:209

    uniform()      = ((w1<<32 | w0) >> 11) * 2**-53
    randint(0, n)  = (w0 * n) >> 32
    noise pixel    = 32 + 95 * (w[x&3] * 2**-32)      (U(32,127) of engine.py:209)
"""
import numpy as np

D_SEED, D_WG_MAT, D_WG_OBJ, D_UPDATE, D_BALANCE, D_NOISE = range(6)

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(key, ctr):
  """Philox4x32-10 (Salmon et al. 2011). key: 2 words, ctr: 4 words -> 4 words."""
  k0, k1 = key[0] & _MASK, key[1] & _MASK
  c0, c1, c2, c3 = (c & _MASK for c in ctr)
  for _ in range(10):
    p0 = _M0 * c0
    p1 = _M1 * c2
    c0, c1, c2, c3 = (
        ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK,
        ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK)
    k0 = (k0 + _W0) & _MASK
    k1 = (k1 + _W1) & _MASK
  return c0, c1, c2, c3


def philox4x32_vec(key, c0, c1, c2, c3):
  """Vectorised Philox over numpy uint64 arrays of 32-bit counters."""
  c0, c1, c2, c3 = (np.asarray(c, np.uint64) & np.uint64(_MASK) for c in (c0, c1, c2, c3))
  k0, k1 = np.uint64(key[0] & _MASK), np.uint64(key[1] & _MASK)
  m = np.uint64(_MASK)
  s32 = np.uint64(32)
  for _ in range(10):
    p0 = np.uint64(_M0) * c0
    p1 = np.uint64(_M1) * c2
    c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & m, p1 & m, ((p0 >> s32) ^ c3 ^ k1) & m, p0 & m
    k0 = (k0 + np.uint64(_W0)) & m
    k1 = (k1 + np.uint64(_W1)) & m
  return c0, c1, c2, c3


def u53(w0, w1):
  return float(((w1 << 32) | w0) >> 11) * (2.0 ** -53)


class KeyedRandom:
  """Drop-in for the three RandomState methods the reference uses (`uniform`, `randint`),
  addressed by an explicit context instead of a sequential stream."""

  def __init__(self, seed=None):
    self.seed = 0 if seed is None else int(seed)
    self.set_ctx(D_SEED)
    self.draws = 0  # census, for tests

  def set_ctx(self, domain, c1=0, c2=0, c3=0):
    self._domain = domain
    self._c = (int(c1), int(c2), int(c3))
    self._k = 0

  def _next(self):
    w = philox4x32((self.seed, self._domain), (self._k,) + self._c)
    self._k += 1
    self.draws += 1
    return w

  def uniform(self, low=0.0, high=1.0, size=None):
    if size is not None:
      assert self._domain == D_NOISE, 'vector draws exist only at engine.py:209'
      xs, ys = np.meshgrid(
          np.arange(size[0], dtype=np.uint64), np.arange(size[1], dtype=np.uint64), indexing='ij')
      w = philox4x32_vec((self.seed, D_NOISE), xs >> np.uint64(2), self._c[0], ys, 0)
      sel = (xs & np.uint64(3)).astype(np.int64)
      word = np.choose(sel, w).astype(np.float64)
      u = word * (2.0 ** -32)
      return low + (high - low) * u
    w = self._next()
    return low + (high - low) * u53(w[0], w[1])

  def randint(self, low, high=None):
    if high is None:
      low, high = 0, low
    n = int(high) - int(low)
    assert 0 < n < 2 ** 32
    w = self._next()
    return int(low) + ((w[0] * n) >> 32)
