"""Decode the reference's own golden PNGs into one small .npz fixture.

Source (read-only, only present in the build container):
  /root/reference/rend3-test/tests/results/{simple,object,shadow,msaa}/*.png   (rend3-test/tests/*.rs)
  /root/reference/examples/src/cube/screenshot.png                              (examples/src/cube/mod.rs:189-200)
These are the known-answer images the reference's tests compare against with nv-flip
(rend3-test/src/runner.rs:227-290).  They are test DATA, decoded to RGBA8 arrays; no reference
source code is copied.  Run:  python tests/golden/make_reference_goldens.py
"""
import glob
import os

import numpy as np
from PIL import Image

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.npz")


def main():
    arrays = {}
    for p in sorted(glob.glob(f"{REF}/rend3-test/tests/results/*/*.png")):
        key = "/".join(p.split("/")[-2:])[:-4]
        arrays[key] = np.asarray(Image.open(p).convert("RGBA"), dtype=np.uint8)
    arrays["examples/cube"] = np.asarray(Image.open(f"{REF}/examples/src/cube/screenshot.png").convert("RGBA"), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    for k, v in arrays.items():
        print(f"{k:45s} {v.shape}")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
