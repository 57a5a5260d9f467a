"""When do the kernels of the step's two branches really run?  %globaltimer stamps (build variant `trace`) of
k_update's ticks, k_post's CTAs, k_wg_mat's / k_wg_obj's CTAs and the frame CTAs of ONE step graph launch,
on one time axis (us after the first tick starts)."""
import ctypes
import os
import pathlib
import sys

os.environ.setdefault('CRAFTER_B200_LIB', str(pathlib.Path(__file__).resolve().parents[1] / 'crafter_b200/_lib/variants/libcrafter_b200_trace.so'))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import crafter_b200  # noqa: E402

B = 4096
env = crafter_b200.Env(num_envs=B, seed=0, auto_reset=True)
lib = env._lib
lib.cr_debug_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
gen = torch.Generator(device='cuda').manual_seed(1234)
actions = torch.randint(0, 17, (256, B), generator=gen, device='cuda', dtype=torch.int32)
env.reset()
for t in range(1000):
  env.step(actions[t % 256])
for rep in range(3):
  torch.cuda.synchronize()
  lib.cr_debug_trace(1, None, 0)
  env.step(actions[rep])
  torch.cuda.synchronize()
  lib.cr_debug_trace(0, None, 0)
  buf = np.zeros((5 * 4096, 8), np.int64)
  lib.cr_debug_trace(0, buf.ctypes.data, buf.size)
  ticks = buf[8192:12288]; ticks = ticks[ticks[:, 4] > 0]
  t0 = ticks[:, 0].min()
  us = lambda v: (v - t0) / 1e3
  post = buf[4096:4096 + 1024, 0]; post = post[post > 0]
  bal = buf[:4096]; bal = bal[bal[:, 5] > 0]
  obj = buf[4096 + 1024:4096 + 2048]; objs = obj[obj[:, 0] > 0]; objw = obj[obj[:, 1] > 0]
  mat = buf[4096 + 2048:4096 + 3072]; mat = mat[mat[:, 1] > 0]
  fr = buf[3 * 4096:4 * 4096]; fr = fr[fr[:, 6] > 0]
  inst = buf[4096:4096 + 1023]; inst = inst[inst[:, 5] > 0]
  part = buf[4096 + 1023]
  view = buf[4096 + 3072:4096 + 4096]; vs = view[view[:, 0] > 0, 0]; ve = view[:, 1:5]; ve = ve[ve > 0]
  print(f'--- step {rep} (us after the first tick starts)')
  print(f'k_update  ticks        start {us(ticks[:, 0].min()):6.1f} .. {us(ticks[:, 0].max()):6.1f}   last end {us(ticks[:, 4].max()):6.1f}')
  print(f'k_post    {len(post):4d} CTAs    start {us(post.min()):6.1f} .. {us(post.max()):6.1f}   last balance end {us(bal[:, 5].max()):6.1f}')
  print(f'k_post    frame order    start {us(part[6]):6.1f}   end {us(part[7]):6.1f}')
  print(f'k_install {len(inst):4d} CTAs    start {us(inst[:, 4].min()):6.1f} .. {us(inst[:, 4].max()):6.1f}   last end {us(inst[:, 5].max()):6.1f}')
  if len(vs):
    print(f'k_view    {len(vs):4d} CTAs    start {us(vs.min()):6.1f} .. {us(vs.max()):6.1f}   last view written {us(ve.max()):6.1f}')
  print(f'k_wg_mat  {len(mat):4d} CTAs    start {us(mat[:, 0].min()):6.1f} .. {us(mat[:, 0].max()):6.1f}   end {us(mat[:, 1].min()):6.1f} .. {us(mat[:, 1].max()):6.1f}')
  print(f'k_wg_obj  {len(objs):4d} CTAs    start {us(objs[:, 0].min()):6.1f} .. {us(objs[:, 0].max()):6.1f}   ({len(objw)} with a world, first at {us(objw[:, 1].min()) if len(objw) else -1:6.1f})')
  print(f'k_render  {len(fr):4d} CTAs    start {us(fr[:, 0].min()):6.1f} .. {us(fr[:, 0].max()):6.1f}   last end {us(fr[:, 6].max()):6.1f}')
