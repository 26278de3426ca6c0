"""Regenerates the BC7 partition / anchor tables (format constants of the BC7 block layout, Khronos Data Format 1.3 tables
"Partition table for 2 subset" / "3 subset" and the anchor index tables) by probing an independent decoder: Pillow's DDS reader.

Every probe is a hand-assembled block whose subsets carry distinct endpoint colours and whose indices are all zero (the decoded colour of a
texel then names its subset) or all one (a texel whose index lost its implied-zero top bit is an anchor).  Prints the tables in the form
include/r3_bc7_tables.h and rend3_b200/bc.py hold them; tests/test_host_cpu.py re-runs the probes against the committed tables.
Needs Pillow (present in the image); not part of the product."""
import io
import struct
import sys

import numpy as np


class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, width):
        assert 0 <= value < (1 << width)
        self.v |= value << self.n
        self.n += width

    def block(self):
        assert self.n == 128, self.n
        return self.v.to_bytes(16, "little")


def pillow_decode(fmt_code, blocks, w, h):
    from PIL import Image

    ddsd = 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000
    pf = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    hdr = struct.pack("<IIIIIII44x", 124, ddsd, h, w, len(blocks), 0, 1) + pf + struct.pack("<IIIII", 0x1000, 0, 0, 0, 0)
    im = Image.open(io.BytesIO(b"DDS " + hdr + struct.pack("<IIIII", fmt_code, 3, 0, 1, 0) + bytes(blocks)))
    im.load()
    return np.asarray(im)


def mode1_block(partition, index_bits):
    b = Bits()
    b.put(0b10, 2)                          # mode 1: one zero, then the one
    b.put(partition, 6)
    for ch in range(3):                     # r, g, b: s0e0 s0e1 s1e0 s1e1, 6 bits each
        for ep in range(4):
            b.put(63 if (ch == 0 and ep >= 2) else (63 if (ch == 1 and ep % 2 == 1) else 0), 6)
    b.put(0, 1); b.put(0, 1)                # shared p bits
    b.put(index_bits, 46)
    return b.block()


def mode2_block(partition, index_bits):
    b = Bits()
    b.put(0b100, 3)
    b.put(partition, 6)
    for ch in range(3):                     # 6 endpoints x 5 bits; subset 1 is red, subset 2 blue; every e1 also carries green
        for ep in range(6):
            b.put(31 if (ch == 0 and ep in (2, 3)) or (ch == 2 and ep in (4, 5)) or (ch == 1 and ep % 2 == 1) else 0, 5)
    b.put(index_bits, 29)
    return b.block()


def derive():
    p2, p3, a2, a3a, a3b = [], [], [], [], []
    for p in range(64):
        img = pillow_decode(98, mode1_block(p, 0), 4, 4).reshape(16, 4)
        subset = (img[:, 0] > 128).astype(int)
        p2.append(sum(int(s) << t for t, s in enumerate(subset)))
        img1 = pillow_decode(98, mode1_block(p, (1 << 46) - 1), 4, 4).reshape(16, 4)     # green = e1 weight: largest where the index kept all its bits
        anchors = [t for t in range(16) if img1[t, 1] != img1[:, 1].max()]
        assert anchors[0] == 0 and len(anchors) == 2 and subset[anchors[1]] == 1, (p, anchors)
        a2.append(anchors[1])
        img = pillow_decode(98, mode2_block(p, 0), 4, 4).reshape(16, 4)
        subset = np.where(img[:, 0] > 128, 1, np.where(img[:, 2] > 128, 2, 0))
        p3.append(sum(int(s) << (2 * t) for t, s in enumerate(subset)))
        img1 = pillow_decode(98, mode2_block(p, (1 << 29) - 1), 4, 4).reshape(16, 4)
        anchors = [t for t in range(16) if img1[t, 1] != img1[:, 1].max()]
        assert anchors[0] == 0 and len(anchors) == 3, (p, anchors)
        by_subset = {int(subset[t]): t for t in anchors}
        a3a.append(by_subset[1]); a3b.append(by_subset[2])
    return p2, p3, a2, a3a, a3b


def main():
    p2, p3, a2, a3a, a3b = derive()

    def rows(vals, fmt, per):
        return ",\n    ".join(", ".join(fmt % v for v in vals[i:i + per]) for i in range(0, len(vals), per))

    print("P2 = [\n    " + rows(p2, "0x%04X", 8) + "]")
    print("P3 = [\n    " + rows(p3, "0x%08X", 8) + "]")
    print("A2 = [" + ", ".join(map(str, a2)) + "]")
    print("A3A = [" + ", ".join(map(str, a3a)) + "]")
    print("A3B = [" + ", ".join(map(str, a3b)) + "]")


if __name__ == "__main__":
    sys.exit(main())
