"""ORACLE shim for `imageio.v3.imread(bytes)` (reference crafter/engine.py:5,125): all assets are
16x16 RGB/RGBA PNGs without palettes, so Pillow decodes them to the same arrays."""
import io

import numpy as np
from PIL import Image


def imread(data):
  return np.array(Image.open(io.BytesIO(data)))
