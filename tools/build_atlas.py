"""Pack the reference's 16x16 sprite PNGs (crafter/assets/*.png, MIT-licensed data, not code)
into crafter_b200/assets/atlas16.npz so that the product and the oracle can run on machines where
/root/reference does not exist.  Arrays are stored exactly as `imageio.v3.imread` returns them
(H, W, C) with C in {3, 4}; consumers apply the reference's transpose (engine.py:126).

Run in the build container:  python tools/build_atlas.py
"""
import io
import pathlib

import numpy as np
from PIL import Image

SRC = pathlib.Path('/root/reference/crafter/assets')
DST = pathlib.Path(__file__).resolve().parents[1] / 'crafter_b200' / 'assets' / 'atlas16.npz'


def main():
  arrays = {}
  for path in sorted(SRC.glob('*.png')):
    img = np.array(Image.open(io.BytesIO(path.read_bytes())))
    assert img.dtype == np.uint8 and img.shape[:2] == (16, 16) and img.shape[2] in (3, 4), path
    arrays[path.stem] = img
  DST.parent.mkdir(parents=True, exist_ok=True)
  np.savez_compressed(DST, **arrays)
  print(f'{len(arrays)} sprites -> {DST} ({DST.stat().st_size} bytes)')


if __name__ == '__main__':
  main()
