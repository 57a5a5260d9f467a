"""memcheck without a GPU: the product's kernels on the SIMT emulator (tests/simt), built with
AddressSanitizer (dynamic shared memory is a fresh exact-size allocation per block), replaying reference-recorded
trajectories under every schedule.  Any out-of-bounds access of shared / global (numpy-owned) memory
or a shared-memory carve-up that overruns aborts with an ASan report.

    bash tools/simt_asan.sh        # builds /tmp/libsimt_asan.so and runs this file under libasan
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np
from tests import hostsim_env, parity
from tests.golden_util import Fixture
from crafter_b200 import _cabi
# load the ASan build in the place of the regular emulator library
L = ctypes.CDLL('/tmp/libsimt_asan.so')
vp = ctypes.c_void_p
L.hs_create.argtypes = [ctypes.POINTER(_cabi.CrConfig), ctypes.POINTER(_cabi.CrTables), ctypes.POINTER(_cabi.CrState), ctypes.POINTER(vp)]
L.hs_destroy.argtypes = [vp]; L.hs_reset.argtypes = [vp, vp, vp]; L.hs_step.argtypes = [vp, vp, vp, vp, vp]
L.hs_render.argtypes = [vp, vp]; L.hs_semantic.argtypes = [vp, vp]; L.hs_simt_blocks.restype = ctypes.c_long
hostsim_env._libs['simt'] = L
for knobs in ({}, dict(CRAFTER_B200_DRAW_PREFETCH='0', CRAFTER_B200_INCR_CENSUS='0'), dict(CRAFTER_B200_DEFER_WG='1'), dict(CRAFTER_B200_DEFER_WG='1', CRAFTER_B200_FUSED='1'), dict(CRAFTER_B200_DEFER_WG='1', CRAFTER_B200_FUSED='2'), dict(CRAFTER_B200_SPLIT='1', CRAFTER_B200_NO_SPECIALIZE='1')):
  for k in ('CRAFTER_B200_DEFER_WG','CRAFTER_B200_FUSED','CRAFTER_B200_DRAW_PREFETCH','CRAFTER_B200_INCR_CENSUS','CRAFTER_B200_SPLIT','CRAFTER_B200_NO_SPECIALIZE'): os.environ.pop(k, None)
  os.environ.update(knobs)
  parity.replay(Fixture('default_short'), hostsim_env.SimtEnv, auto_reset=True, steps=80)
  parity.replay(Fixture('tiny_area'), hostsim_env.SimtEnv, auto_reset=False, steps=40)
  parity.replay(Fixture('big_view'), hostsim_env.SimtEnv, auto_reset=False, steps=30)
  print('ok', knobs, flush=True)
from tests import scenario_util as su
from tests.test_scenarios_golden import replay_group
for k in ('CRAFTER_B200_DEFER_WG','CRAFTER_B200_FUSED','CRAFTER_B200_DRAW_PREFETCH','CRAFTER_B200_INCR_CENSUS','CRAFTER_B200_SPLIT','CRAFTER_B200_NO_SPECIALIZE'): os.environ.pop(k, None)
replay_group('directed_default', hostsim_env.SimtEnv, su.load_numpy)
print('scenarios ok')
