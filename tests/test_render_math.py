"""Identities the render kernel relies on, checked exhaustively against PIL / numpy on the host."""
import numpy as np
from PIL import Image, ImageEnhance


def test_enhance_color_is_integer_division():
  """csrc/cr_render.h enhance(): PIL's ImageEnhance.Color(img).enhance(0.4) is
  Image.blend(grey, img, 0.4) = trunc(L + 0.4f * (c - L)) in float32, which for 8-bit L and c
  equals (3L + 2c) // 5 for all 65536 pairs."""
  L, c = np.meshgrid(np.arange(256), np.arange(256), indexing='ij')
  ref = (3 * L + 2 * c) // 5
  f32 = (L.astype(np.float32) + np.float32(0.4) * (c - L).astype(np.float32)).astype(np.uint8)
  assert (f32 == ref).all()
  blend = np.array(Image.blend(Image.fromarray(L.astype(np.uint8)), Image.fromarray(c.astype(np.uint8)), 0.4))
  assert (blend == ref).all()


def test_enhance_uses_rounded_luma():
  """The grey image of ImageEnhance.Color is convert('L') = (19595 R + 38470 G + 7471 B + 32768) >> 16."""
  rs = np.random.RandomState(0)
  rgb = rs.randint(0, 256, (64, 64, 3)).astype(np.uint8)
  out = np.array(ImageEnhance.Color(Image.fromarray(rgb)).enhance(0.4)).astype(np.int64)
  r, g, b = (rgb[..., k].astype(np.int64) for k in range(3))
  L = (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16
  ref = (3 * L[..., None] + 2 * rgb.astype(np.int64)) // 5
  assert (out == ref).all()
  grey = np.array(ImageEnhance.Color(Image.fromarray(rgb)).enhance(0.0))
  assert (grey == L[..., None]).all()


def test_alpha_opaque_and_transparent_are_exact():
  """engine.py:276-284 in float32: alpha 255 returns the texture, alpha 0 the canvas, exactly."""
  v = np.arange(256, dtype=np.uint8)
  t, c = np.meshgrid(v, v, indexing='ij')
  for alpha, want in ((255, t), (0, c)):
    a = np.float32(alpha) / 255
    out = (255 * (a * (t.astype(np.float32) / 255) + (1 - a) * (c.astype(np.float32) / 255))).astype(np.uint8)
    assert (out == want).all()
