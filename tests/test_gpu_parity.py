"""`-m gpu`: the CUDA library (through the C ABI and crafter_b200.Env) replays the golden
trajectories of the unmodified reference bit for bit -- integer state, reward, done, observation."""
import functools
import os

import numpy as np
import pytest

from tests import parity
from tests.golden_util import Fixture, NAMES

pytestmark = pytest.mark.gpu


def make_env(**kwargs):
  import crafter_b200
  return crafter_b200.Env(**kwargs)


@pytest.mark.parametrize('name', NAMES)
def test_cuda_matches_reference(name):
  parity.replay(Fixture(name), make_env, auto_reset=False)


@pytest.mark.parametrize('name', ['default_random', 'default_short', 'default_rich'])
def test_cuda_auto_reset(name):
  parity.replay(Fixture(name), make_env, auto_reset=True)


def test_cuda_slot_compaction():
  fx = Fixture('default_fighter')
  env = parity.replay(fx, functools.partial(make_env, slot_capacity=128), steps=400)
  assert int(env.state['pstate'][:, 14].abs().sum()) == 0


def test_cuda_native_library_loaded():
  """The product path is the in-tree CUDA library; there is no fallback to fail over to."""
  from crafter_b200 import _cabi
  lib = _cabi.load()
  assert lib.cr_abi_version() == _cabi.ABI_VERSION
  with open('/proc/self/maps') as f:
    assert 'libcrafter_b200.so' in f.read()
