"""Build-time properties of the CUDA library that the measured performance depends on, checked
without a GPU from ptxas' report and the SASS (nvcc cross-compiles sm_100a anywhere):

* `k_render<default>` fits 40 registers, i.e. 6 CTAs of 256 threads per SM (profiles/README.md:
  121.4 vs 123.1 us/step against 5 CTAs; 7 CTAs spill and lose);
* the observation leaves the SM as ONE bulk (TMA) store per env -- `UBLKCP` in the SASS;
* `k_wg_mat` stays within the 80 registers of its 3-CTAs-per-SM launch bound without spilling;
* nothing in the hot kernels spills more than a few words."""
import re
import shutil
import subprocess

import pytest

from crafter_b200 import build


@pytest.fixture(scope='module')
def ptxas():
  if shutil.which(build.NVCC) is None and not build.OUT.exists():
    pytest.skip('no nvcc and no built library')
  out = build.build()
  log = (out.parent / 'ptxas.log')
  if not log.exists():
    build.build(force=True)
  info, cur, props_of_entry = {}, None, False
  for line in log.read_text().splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
      k = re.search(r'\d+(k_[a-z_0-9]+?)(I(?:L[a-z]\d+E)+E)?E', m.group(1))
      cur = (k.group(1), k.group(2) or '') if k else None
      continue
    m = re.search(r'Function properties for (\S+)', line)
    if m:  # out-of-line device functions (consume_heavy, entity_update ...) report their own frames
      own = cur is not None and re.search(r'\d+(k_[a-z_0-9]+?)(I(?:L[a-z]\d+E)+E)?E', m.group(1))
      props_of_entry = bool(own) and (own.group(1), own.group(2) or '') == cur
      continue
    if cur is None or not props_of_entry:
      continue
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
    if m:
      info.setdefault(cur, {})['spill'] = int(m.group(2)) + int(m.group(3))
    m = re.search(r'Used (\d+) registers', line)
    if m:
      info.setdefault(cur, {})['regs'] = int(m.group(1))
  return info


def test_register_budgets(ptxas):
  render_default = [v for (name, targs), v in ptxas.items() if name == 'k_render' and targs.startswith('ILb1E')]
  assert render_default and all(v['regs'] <= 40 for v in render_default), render_default  # 6 CTAs x 256 threads
  wg = [v for (name, targs), v in ptxas.items() if name == 'k_wg_mat']
  assert wg and all(v['regs'] <= 64 and v['spill'] <= 64 for v in wg), wg  # 4 CTAs x 256 threads, a few words spilled
  for (name, targs), v in ptxas.items():
    # (k_terminal -- final_obs only, a few dozen CTAs per step -- inlines the balance next to the frame)
    limit = 320 if name == 'k_terminal' else 64
    assert v.get('spill', 0) <= limit, (name, targs, v)  # a few words at most, never a spilled array


def test_observation_leaves_as_one_bulk_store():
  cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
  if shutil.which(cuobjdump) is None:
    pytest.skip('no cuobjdump')
  sass = subprocess.run([cuobjdump, '-sass', str(build.build())], capture_output=True, text=True).stdout
  kernels = re.split(r'\n\s*Function : ', sass)
  render = [k for k in kernels if k.startswith('_Z') and ('k_render' in k.split('\n', 1)[0] or 'k_terminal' in k.split('\n', 1)[0])]
  assert len(render) >= 4, 'k_render / k_terminal not found in the SASS'
  for body in render:
    assert 'UBLKCP' in body, body.split('\n', 1)[0]  # cp.async.bulk shared -> global
  assert not any('HMMA' in k or 'UTCMMA' in k for k in kernels)  # no tensor-core op anywhere: none is needed
