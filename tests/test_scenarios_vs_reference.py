"""Perturbation fuzz against the LIVE reference (build container only; skipped where
/root/reference is absent): random-policy trajectories rarely reach the corner cases of the rules
(crafting at the map edge, lava, dying mobs that still act, arrows hitting things, ripe plants,
full inventories ...), so this test teleports the player, rewrites terrain, spawns creatures with
odd attributes and sets inventories *inside the reference env through its own API*, loads the very
same canonical state into the device logic (host-sim build of csrc/cr_*.h), and then steps both
with random actions, comparing the full integer state, reward, done and the observation each step."""
import numpy as np
import pytest

from oracle import canon
from oracle import ref_harness as rh
from tests import hostsim_env
from tests import scenario_util as su

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference not mounted')

def load_into_hostsim(hs, i, env, st):
  """Write a canonical reference state into env i of a HostSimEnv (layout: csrc/cr_common.h)."""
  su.load_numpy(hs.state, i, su.raw_arrays(st, su.extras_of(env), hs.area, hs.capacity))


@pytest.mark.parametrize('geometry', [dict(), dict(area=(24, 20))])
def test_perturbed_states_step_like_the_reference(geometry):
  mods = rh.load()
  rs = np.random.RandomState(4711)  # tests/golden/scenarios holds the 2024 / seed 900+ stream
  rounds, steps = (40, 45) if not geometry else (30, 40)
  for r in range(rounds):
    seed = 5000 + r
    ref = rh.make_env(seed, **geometry)
    ref.reset()
    hs = hostsim_env.HostSimEnv(num_envs=1, seed=seed, **geometry)
    hs.reset()
    su.perturb(ref, rs, mods)
    st = rh.export_state(ref)
    load_into_hostsim(hs, 0, ref, st)
    hs.recount()
    assert canon.diff(st, hs.snapshot(0)) is None
    assert (ref.render() == hs.render()[0]).all(), ('render after load', r)
    for t, a in enumerate(su.fuzz_actions(rs, steps)):
      obs, reward, done, info = ref.step(a)
      hobs, hreward, hdone = hs.step(np.array([a]))
      problem = canon.diff(rh.export_state(ref), hs.snapshot(0))
      assert problem is None, (geometry, r, t, a, problem)
      assert np.float32(reward) == hreward[0] and done == bool(hdone[0]), (r, t, reward, hreward[0])
      assert (obs == hobs[0]).all(), (geometry, r, t, 'obs')
      if done:
        break
