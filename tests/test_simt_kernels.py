"""The product's CUDA KERNELS on a CPU: crafter_b200/csrc/cr_kernels.h -- the file nvcc compiles --
built for the host on a small SIMT emulator (tests/simt: one fiber per CUDA thread, block and warp
barriers, warp collectives, shared memory, grid-stride loops on a pretend 3-SM device) and launched
in the order of crafter_kernels.cu's step graph.  tests/hostsim checks the per-lane rule / render
logic with one lane and sequential phases; this checks what only showed on the GPU before: the
32-lane paths (radius ballots, order-preserving slot compaction, the draw table), the CTA
choreography (k_post's census / decide / apply, k_wg_mat's work lists, k_wg_obj's block prefix sum,
k_render's four phases, k_terminal's scratch laid over the output tile) and barrier
divergence (reported as a deadlock).  Streams, graphs and TMA are not modelled; `-m gpu` covers them.

Every replay compares with what the UNMODIFIED reference recorded (tests/golden), bit for bit."""
import functools

import numpy as np
import pytest

from tests import hostsim_env
from tests import parity
from tests import scenario_util as su
from tests.golden_util import Fixture
from tests.test_schedule_knobs import check_against_oracle
from tests.test_scenarios_golden import replay_group

SIMT = hostsim_env.SimtEnv

KNOBS = {
    'default': {},
    'generic': dict(CRAFTER_B200_NO_SPECIALIZE='1'),
    'no_draw_prefetch': dict(CRAFTER_B200_DRAW_PREFETCH='0'),
    'no_incr_census': dict(CRAFTER_B200_INCR_CENSUS='0'),
    'plain_tick': dict(CRAFTER_B200_INCR_CENSUS='0', CRAFTER_B200_DRAW_PREFETCH='0'),
    'obj_before_seed': dict(CR_SIMT_WG_ORDER='obj'),  # k_seed ahead beside k_wg_mat -> k_wg_obj: after both (default: before)
    'seed_between': dict(CR_SIMT_WG_ORDER='mat'),  # ... or between them
    'view_late': dict(CR_SIMT_VIEW_LATE='1'),  # k_view beside both branches: after them instead of before
    'late_first': dict(CR_SIMT_LATE_FIRST='1'),  # k_post before the side branch (k_terminal, k_install)
    'generic_plain': dict(CRAFTER_B200_NO_SPECIALIZE='1', CRAFTER_B200_INCR_CENSUS='0', CRAFTER_B200_DRAW_PREFETCH='0'),
}
ALL_KNOBS = ('CRAFTER_B200_NO_SPECIALIZE', 'CRAFTER_B200_DRAW_PREFETCH', 'CRAFTER_B200_INCR_CENSUS',
             'CR_SIMT_WG_ORDER', 'CR_SIMT_LATE_FIRST', 'CR_SIMT_VIEW_LATE')


def set_knobs(monkeypatch, name):
  for k in ALL_KNOBS:
    monkeypatch.delenv(k, raising=False)
  for k, v in KNOBS[name].items():
    monkeypatch.setenv(k, v)


@pytest.mark.parametrize('name,steps', [
    ('default_random', 330), ('default_sleepy', 300), ('default_rich', 200), ('big_view', 160),
    ('odd_geometry', 160), ('tiny_area', 200), ('big_area', 40), ('even_view', 150), ('wide_view', 100)])
def test_kernels_replay_golden(name, steps):
  parity.replay(Fixture(name), SIMT, auto_reset=False, steps=steps)


def test_kernels_slot_compaction():
  """A small arena: compact_slots (ballot / popc ranks over 32 lanes) runs almost every step."""
  env = parity.replay(Fixture('default_fighter'), functools.partial(SIMT, slot_capacity=128), steps=260)
  assert (env.state['pstate'][:, 14] == 0).all()


@pytest.mark.parametrize('knobs', list(KNOBS))
def test_kernels_auto_reset_schedules(monkeypatch, knobs):
  """Every schedule the library can run, auto-reset on: the golden episodes (length 50, so worlds
  are consumed and refilled all the time) and a few steps of a longer fixture."""
  set_knobs(monkeypatch, knobs)
  parity.replay(Fixture('default_short'), SIMT, auto_reset=True)
  parity.replay(Fixture('default_random'), SIMT, auto_reset=True, steps=60)


@pytest.mark.parametrize('knobs', ['default', 'generic', 'late_first'])
def test_kernels_explicit_resets_knobs(monkeypatch, knobs):
  set_knobs(monkeypatch, knobs)
  parity.replay(Fixture('default_short'), SIMT, auto_reset=False)
  parity.replay(Fixture('odd_geometry'), SIMT, auto_reset=False, steps=80)


@pytest.mark.parametrize('knobs', ['default', 'obj_before_seed', 'seed_between', 'view_late', 'plain_tick', 'late_first'])
@pytest.mark.parametrize('length', [1, 2, 3])
def test_kernels_back_to_back_resets(monkeypatch, knobs, length):
  set_knobs(monkeypatch, knobs)
  check_against_oracle(SIMT, np.asarray, length, steps=9)


@pytest.mark.parametrize('knobs', ['default', 'late_first'])
def test_kernels_mixed_resets_and_masks(monkeypatch, knobs):
  """reset(mask) between auto-resets."""
  from tests.test_schedule_knobs import check_mixed_resets_and_masks
  set_knobs(monkeypatch, knobs)
  check_mixed_resets_and_masks(SIMT)


@pytest.mark.parametrize('length,kwargs', [(1, {}), (10, {}), (20, dict(view=(7, 9), size=(70, 72), area=(48, 40)))])
def test_kernels_terminal_frames(length, kwargs):
  """final_obs: k_terminal balances (when due) and draws the terminal frame before k_install."""
  from tests.test_schedule_knobs import check_terminal_frames
  assert check_terminal_frames(SIMT, np.asarray, length, steps=41, **kwargs) >= 3 * (41 // length)


@pytest.mark.parametrize('knobs,group', [
    ('default', 'directed_default'), ('default', 'fuzz_default'), ('default', 'directed_short'),
    ('no_incr_census', 'directed_default'), ('no_incr_census', 'fuzz_small'), ('no_draw_prefetch', 'directed_default')])
def test_kernels_replay_scenarios(monkeypatch, knobs, group):
  """Corner-case scenarios (auto-reset off: the schedules only differ with auto-reset on):
  `many_objects` needs several ballot rounds per tick and compacts its arena while arrows append."""
  set_knobs(monkeypatch, knobs)
  replay_group(group, SIMT, su.load_numpy)


# Thread order: the emulator's default runs threads 0, 1, 2 ... until each yields; the order is read
# once per process (CR_SIMT_ORDER), so other orders run in a subprocess.
@pytest.mark.parametrize('order', ['reverse', 'shuffle'])
def test_kernels_do_not_depend_on_thread_order(order):
  import os
  import subprocess
  import sys
  code = (
      "import numpy as np\n"
      "from tests import hostsim_env, parity, scenario_util as su\n"
      "from tests.golden_util import Fixture\n"
      "from tests.test_scenarios_golden import replay_group\n"
      "import os\n"
      "parity.replay(Fixture('default_short'), hostsim_env.SimtEnv, auto_reset=True)\n"
      "parity.replay(Fixture('default_rich'), hostsim_env.SimtEnv, auto_reset=False, steps=120)\n"
      "replay_group('directed_default', hostsim_env.SimtEnv, su.load_numpy)\n"
      "print('order ok')\n")
  out = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, CR_SIMT_ORDER=order),
                       capture_output=True, text=True, cwd=str(hostsim_env.HERE.parent), timeout=1200)
  assert out.returncode == 0 and 'order ok' in out.stdout, (out.stdout[-800:], out.stderr[-2500:])


def test_frame_order_is_night_first_and_stable():
  """The frame kernel's CTA -> env map (frame_partition, one CTA of k_post): a permutation of the envs, the ones
  whose next frame is a night frame first, each group in env order; and the tick's flags say what they claim
  (night = daylight of the env's step below 0.5; final = neither balanced nor regenerated this step)."""
  B = 37
  env = SIMT(num_envs=B, seed=3, length=400, auto_reset=True)
  env.reset()
  env.state['pstate'][:, 9] = np.arange(B) * 9  # steps 0 .. 324: days and nights (nightfall near step 148)
  order, flags = np.zeros(B, np.int32), np.zeros(B, np.uint8)
  seen_night = seen_day = False
  for t in range(12):
    _, _, done = env.step(np.full(B, t % 17, np.int32))
    env._L.hs_frame_order(env.h, order.ctypes.data, flags.ctypes.data)
    step = env.state['pstate'][:, 9]
    night = env.tables['daylight'][np.minimum(step, len(env.tables['daylight']) - 1)] < 0.5
    night &= ~done  # a regenerated env starts by day
    assert ((flags & 1) != 0).tolist() == night.tolist()
    prev_step = np.where(done, -1, step)  # the tick balanced on its own step count
    final = ~done & (prev_step % 10 != 0)
    assert ((flags & 2) != 0).tolist() == final.tolist()
    want = [e for e in range(B) if night[e]] + [e for e in range(B) if not night[e]]
    assert order.tolist() == want
    seen_night |= bool(night.any()); seen_day |= bool((~night).any())
  assert seen_night and seen_day


def test_emulator_reports_barrier_divergence():
  """The emulator's own contract: it really ran blocks, and it is the kernels' file it compiled."""
  L = hostsim_env.simt_lib()
  before = L.hs_simt_blocks()
  env = SIMT(num_envs=2, seed=1)
  env.reset()
  env.step(np.zeros(2, np.int32))
  assert L.hs_simt_blocks() > before
  src = (hostsim_env.HERE / 'simt' / 'simt_env.cpp').read_text()
  assert 'crafter_b200/csrc/cr_kernels.h' in src


@pytest.mark.parametrize('size', [(128, 128), (96, 80), (512, 512)])
def test_kernels_render_at_other_sizes(size):
  """k_render's generic instantiation: other units, the unstaged path, and (512, 512) without a tile
  cache (every cell per pixel) against the C oracle at that size."""
  from oracle import oracle_env
  env = SIMT(num_envs=2, seed=77, size=size)
  obs = env.reset()
  for i in range(2):
    ref = oracle_env.OracleEnv(seed=77 + i, size=size)
    assert (ref.reset() == obs[i]).all(), (size, i)


def test_kernels_semantic_view():
  from oracle import oracle_env
  env = SIMT(num_envs=2, seed=3)
  env.reset()
  env.step(np.array([2, 4], np.int32))
  sem = env.semantic()
  for i, a in enumerate([2, 4]):
    ref = oracle_env.OracleEnv(seed=3 + i)
    ref.reset()
    ref.step(a)
    assert (ref.semantic() == sem[i]).all()
