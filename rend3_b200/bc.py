"""Block-compressed texture data (BC1 - BC5) for the scenes and tests.

rend3 itself never compresses anything: rend3-gltf's ktx2 / dds loaders (rend3-gltf/src/lib.rs:1300-1335, 1556-1602) pass the blocks they
read to Renderer::add_texture_2d with a Bc* TextureFormat and the hardware sampler decodes them.  This module stands in for the asset:
a plain bounding-box encoder (quality is irrelevant — parity is about the DECODE) and a float64 numpy decoder of the published palettes
(D3D11 functional spec 19.5 / Khronos Data Format 1.3 ch. 18-20) that the tests hold both the oracle's and the CUDA kernels' rule-R11
decode against.  Nothing here runs on the product path."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .layouts import (TEXFMT_BC1_RGBA_UNORM, TEXFMT_BC1_RGBA_UNORM_SRGB, TEXFMT_BC2_RGBA_UNORM, TEXFMT_BC2_RGBA_UNORM_SRGB, TEXFMT_BC3_RGBA_UNORM,
                      TEXFMT_BC3_RGBA_UNORM_SRGB, TEXFMT_BC4_R_SNORM, TEXFMT_BC4_R_UNORM, TEXFMT_BC5_RG_SNORM, TEXFMT_BC5_RG_UNORM, TEXFMT_BC7_RGBA_UNORM,
                      TEXFMT_BC7_RGBA_UNORM_SRGB)

# name -> (unorm / snorm format, sRGB format or None, bytes per block)
BLOCK_FORMATS = {
    "bc1": (TEXFMT_BC1_RGBA_UNORM, TEXFMT_BC1_RGBA_UNORM_SRGB, 8),
    "bc2": (TEXFMT_BC2_RGBA_UNORM, TEXFMT_BC2_RGBA_UNORM_SRGB, 16),
    "bc3": (TEXFMT_BC3_RGBA_UNORM, TEXFMT_BC3_RGBA_UNORM_SRGB, 16),
    "bc4": (TEXFMT_BC4_R_UNORM, None, 8),
    "bc4s": (TEXFMT_BC4_R_SNORM, None, 8),
    "bc5": (TEXFMT_BC5_RG_UNORM, None, 16),
    "bc5s": (TEXFMT_BC5_RG_SNORM, None, 16),
    "bc7": (TEXFMT_BC7_RGBA_UNORM, TEXFMT_BC7_RGBA_UNORM_SRGB, 16),
}


def _blocks(img: np.ndarray) -> Tuple[np.ndarray, int, int]:
    """(h, w, c) -> (block rows * block columns, 16, c), edge texels repeated into the padding; texel t of a block = 4 * py + px."""
    h, w = img.shape[:2]
    bh, bw = (h + 3) // 4, (w + 3) // 4
    pad = np.pad(img, ((0, bh * 4 - h), (0, bw * 4 - w), (0, 0)), mode="edge")
    return pad.reshape(bh, 4, bw, 4, -1).transpose(0, 2, 1, 3, 4).reshape(bh * bw, 16, -1), bh, bw


# ------------------------------------------------------------------ palettes (float64, the published ratios)
def _colour_palette(c0: np.ndarray, c1: np.ndarray, bc1: bool) -> np.ndarray:
    """c0, c1: (n,) uint16 RGB565 -> (n, 4 codes, 4 channels)."""
    def rgb(c):
        return np.stack([(c >> 11) / 31.0, ((c >> 5) & 63) / 63.0, (c & 31) / 31.0], axis=-1)

    e0, e1 = rgb(c0.astype(np.int64)), rgb(c1.astype(np.int64))
    four = np.ones(len(c0), dtype=bool) if not bc1 else c0 > c1
    pal = np.zeros((len(c0), 4, 4))
    pal[:, 0, :3], pal[:, 1, :3] = e0, e1
    pal[:, 2, :3] = np.where(four[:, None], (2 * e0 + e1) / 3.0, (e0 + e1) / 2.0)
    pal[:, 3, :3] = np.where(four[:, None], (e0 + 2 * e1) / 3.0, 0.0)
    pal[..., 3] = 1.0
    pal[:, 3, 3] = np.where(four, 1.0, 0.0)
    return pal


def _channel_palette(r0: np.ndarray, r1: np.ndarray, snorm: bool) -> np.ndarray:
    """r0, r1: raw endpoint bytes (uint8) -> (n, 8 codes) values."""
    a, b = (r0.astype(np.int8).astype(np.int64), r1.astype(np.int8).astype(np.int64)) if snorm else (r0.astype(np.int64), r1.astype(np.int64))
    wide = a > b
    full = 127.0 if snorm else 255.0
    if snorm:
        a, b = np.maximum(a, -127), np.maximum(b, -127)
    pal = np.zeros((len(r0), 8))
    pal[:, 0], pal[:, 1] = a / full, b / full
    for k in range(2, 8):
        eight = ((8 - k) * a + (k - 1) * b) / (7.0 * full)
        six = ((6 - k) * a + (k - 1) * b) / (5.0 * full) if k < 6 else np.full(len(r0), (-1.0 if snorm else 0.0) if k == 6 else 1.0)
        pal[:, k] = np.where(wide, eight, six)
    return pal


# ------------------------------------------------------------------ encoder
def _to565(rgb8: np.ndarray) -> np.ndarray:
    r = np.rint(rgb8[..., 0] / 255.0 * 31).astype(np.uint16)
    g = np.rint(rgb8[..., 1] / 255.0 * 63).astype(np.uint16)
    b = np.rint(rgb8[..., 2] / 255.0 * 31).astype(np.uint16)
    return (r << 11) | (g << 5) | b


def _encode_colour(blk: np.ndarray, bc1: bool) -> np.ndarray:
    """blk (n, 16, 4) uint8 -> (n, 8) bytes.  BC1 blocks holding a texel with alpha < 128 use the three-colour + transparent mode."""
    n = len(blk)
    hi, lo = _to565(blk[..., :3].max(axis=1).astype(np.float64)), _to565(blk[..., :3].min(axis=1).astype(np.float64))
    c0, c1 = np.maximum(hi, lo), np.minimum(hi, lo)
    punch = (blk[..., 3] < 128).any(axis=1) if bc1 else np.zeros(n, dtype=bool)
    c0, c1 = np.where(punch, c1, c0).astype(np.uint16), np.where(punch, np.maximum(hi, lo), c1).astype(np.uint16)   # punch-through wants c0 <= c1
    pal = _colour_palette(c0, c1, bc1)[..., :3]                                                                      # (n, 4, 3)
    usable = np.ones((n, 4), dtype=bool)
    usable[:, 3] = ~(bc1 & (c0 <= c1))                                                                              # code 3 = transparent there
    d = ((blk[:, :, None, :3] / 255.0 - pal[:, None, :, :]) ** 2).sum(axis=-1)                                      # (n, 16, 4)
    d = np.where(usable[:, None, :], d, np.inf)
    code = d.argmin(axis=-1).astype(np.uint32)
    if bc1:
        code = np.where((blk[..., 3] < 128) & (c0 <= c1)[:, None], 3, code).astype(np.uint32)
    bits = (code << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    out = np.zeros((n, 8), dtype=np.uint8)
    out[:, 0:2] = c0.astype("<u2").view(np.uint8).reshape(n, 2)
    out[:, 2:4] = c1.astype("<u2").view(np.uint8).reshape(n, 2)
    out[:, 4:8] = bits.astype("<u4").view(np.uint8).reshape(n, 4)
    return out


def _encode_channel(vals: np.ndarray, snorm: bool) -> np.ndarray:
    """vals (n, 16): uint8 (unorm) or int8-range integers (snorm) -> (n, 8) bytes.  Every fifth block (and every flat one) uses the
    six-value mode (e0 <= e1) so both palettes are exercised."""
    n = len(vals)
    v = vals.astype(np.int64)
    hi, lo = v.max(axis=1), v.min(axis=1)
    six = (np.arange(n) % 5 == 4) | (hi == lo)
    e0, e1 = np.where(six, lo, hi), np.where(six, hi, lo)
    raw0, raw1 = (e0.astype(np.int8).view(np.uint8), e1.astype(np.int8).view(np.uint8)) if snorm else (e0.astype(np.uint8), e1.astype(np.uint8))
    pal = _channel_palette(raw0, raw1, snorm)
    target = v / (127.0 if snorm else 255.0)
    code = np.abs(target[:, :, None] - pal[:, None, :]).argmin(axis=-1).astype(np.uint64)
    bits = (code << (3 * np.arange(16, dtype=np.uint64))[None, :]).sum(axis=1, dtype=np.uint64)
    out = np.zeros((n, 8), dtype=np.uint8)
    out[:, 0], out[:, 1] = raw0, raw1
    out[:, 2:8] = bits.astype("<u8").view(np.uint8).reshape(n, 8)[:, :6]
    return out


def snorm_source(u8: np.ndarray) -> np.ndarray:
    """The signed values an snorm format stores for a uint8 source image: v - 128 clamped to [-127, 127]."""
    return np.clip(u8.astype(np.int64) - 128, -127, 127)


def encode(name: str, rgba8: np.ndarray) -> np.ndarray:
    """(h, w, 4) uint8 -> the level's blocks as a flat uint8 array (row-major blocks)."""
    blk, _, _ = _blocks(rgba8)
    if name == "bc1":
        out = _encode_colour(blk, True)
    elif name == "bc2":
        a4 = np.rint(blk[..., 3] / 17.0).astype(np.uint64)
        alpha = (a4 << (4 * np.arange(16, dtype=np.uint64))[None, :]).sum(axis=1, dtype=np.uint64).astype("<u8").view(np.uint8).reshape(len(blk), 8)
        out = np.concatenate([alpha, _encode_colour(blk, False)], axis=1)
    elif name == "bc3":
        out = np.concatenate([_encode_channel(blk[..., 3], False), _encode_colour(blk, False)], axis=1)
    elif name in ("bc4", "bc4s"):
        s = name.endswith("s")
        out = _encode_channel(snorm_source(blk[..., 0]) if s else blk[..., 0], s)
    elif name in ("bc5", "bc5s"):
        s = name.endswith("s")
        out = np.concatenate([_encode_channel(snorm_source(blk[..., k]) if s else blk[..., k], s) for k in (0, 1)], axis=1)
    elif name == "bc7":
        out = _encode_bc7_mode6(blk)
    else:
        raise ValueError(f"unknown block format {name!r}")
    return np.ascontiguousarray(out).reshape(-1)


# ------------------------------------------------------------------ reference decoder (float64)
def _decode_colour(b: np.ndarray, bc1: bool) -> np.ndarray:
    c0, c1 = b[:, 0:2].copy().view("<u2")[:, 0], b[:, 2:4].copy().view("<u2")[:, 0]
    bits = b[:, 4:8].copy().view("<u4")[:, 0].astype(np.uint64)
    code = ((bits[:, None] >> (2 * np.arange(16, dtype=np.uint64))[None, :]) & 3).astype(np.int64)
    pal = _colour_palette(c0, c1, bc1)
    return np.take_along_axis(pal, code[:, :, None].repeat(4, axis=2), axis=1)                # (n, 16, 4)


def _decode_channel(b: np.ndarray, snorm: bool) -> np.ndarray:
    wide = np.zeros((len(b), 8), dtype=np.uint8)
    wide[:, :6] = b[:, 2:8]
    bits = wide.view("<u8")[:, 0]
    code = ((bits[:, None] >> (3 * np.arange(16, dtype=np.uint64))[None, :]) & 7).astype(np.int64)
    return np.take_along_axis(_channel_palette(b[:, 0], b[:, 1], snorm), code, axis=1)       # (n, 16)


def decode(name: str, data: np.ndarray, width: int, height: int, srgb: bool = False) -> np.ndarray:
    """Blocks of one level -> (height, width, 4) float64 as the sampler returns the texels (missing channels (0, 0, 1), sRGB decoded)."""
    size = BLOCK_FORMATS[name][2]
    bh, bw = (height + 3) // 4, (width + 3) // 4
    b = np.asarray(data, dtype=np.uint8).reshape(bh * bw, size)
    out = np.zeros((bh * bw, 16, 4))
    out[..., 3] = 1.0
    if name == "bc7":
        out = np.stack([decode_bc7_block(bytes(row)) for row in b]).astype(np.float64) / 255.0
    elif name == "bc1":
        out = _decode_colour(b, True)
    elif name == "bc2":
        out = _decode_colour(b[:, 8:], False)
        a = b[:, :8].copy().view("<u8")[:, 0]
        out[..., 3] = ((a[:, None] >> (4 * np.arange(16, dtype=np.uint64))[None, :]) & 15).astype(np.float64) / 15.0
    elif name == "bc3":
        out = _decode_colour(b[:, 8:], False)
        out[..., 3] = _decode_channel(b[:, :8], False)
    elif name in ("bc4", "bc4s"):
        out[..., 0] = _decode_channel(b, name.endswith("s"))
    else:
        out[..., 0] = _decode_channel(b[:, :8], name.endswith("s"))
        out[..., 1] = _decode_channel(b[:, 8:], name.endswith("s"))
    if srgb:
        c = out[..., :3]
        out[..., :3] = np.where(c > 0.04045, ((c + 0.055) / 1.055) ** 2.4, c / 12.92)
    img = out.reshape(bh, bw, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(bh * 4, bw * 4, 4)
    return img[:height, :width]


# ------------------------------------------------------------------ BC7 (integer-exact by the format: every decoder gives the same 8-bit texels)
# mode -> (subsets, partition bits, rotation bits, index-selection bits, colour bits, alpha bits, per-endpoint p bits, shared p bits, index bits, secondary index bits)
BC7_MODES = (
    (3, 4, 0, 0, 4, 0, 1, 0, 3, 0), (2, 6, 0, 0, 6, 0, 0, 1, 3, 0), (3, 6, 0, 0, 5, 0, 0, 0, 2, 0), (2, 6, 0, 0, 7, 0, 1, 0, 2, 0),
    (1, 0, 2, 1, 5, 6, 0, 0, 2, 3), (1, 0, 2, 0, 7, 8, 0, 0, 2, 2), (1, 0, 0, 0, 7, 7, 1, 0, 4, 0), (2, 6, 0, 0, 5, 5, 1, 0, 2, 0),
)
BC7_WEIGHTS = {2: (0, 21, 43, 64), 3: (0, 9, 18, 27, 37, 46, 55, 64), 4: (0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64)}
# partition of the 16 texels in 2 subsets (bit t) and 3 subsets (bits 2t, 2t + 1), anchor texels of subset 1 (2 subsets) and of subsets 1 / 2
# (3 subsets): constants of the format (Khronos Data Format 1.3, BC7 partition / anchor tables); tools/derive_bc7_tables.py regenerates them
BC7_P2 = (
    0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8, 0xFF00, 0xFFF0, 0xF000,
    0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110, 0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C,
    0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696, 0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660,
    0x0272, 0x04E4, 0x4E40, 0x2720, 0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22)
BC7_P3 = (
    0xAA685050, 0x6A5A5040, 0x5A5A4200, 0x5450A0A8, 0xA5A50000, 0xA0A05050, 0x5555A0A0, 0x5A5A5050, 0xAA550000, 0xAA555500, 0xAAAA5500, 0x90909090,
    0x94949494, 0xA4A4A4A4, 0xA9A59450, 0x2A0A4250, 0xA5945040, 0x0A425054, 0xA5A5A500, 0x55A0A0A0, 0xA8A85454, 0x6A6A4040, 0xA4A45000, 0x1A1A0500,
    0x0050A4A4, 0xAAA59090, 0x14696914, 0x69691400, 0xA08585A0, 0xAA821414, 0x50A4A450, 0x6A5A0200, 0xA9A58000, 0x5090A0A8, 0xA8A09050, 0x24242424,
    0x00AA5500, 0x24924924, 0x24499224, 0x50A50A50, 0x500AA550, 0xAAAA4444, 0x66660000, 0xA5A0A5A0, 0x50A050A0, 0x69286928, 0x44AAAA44, 0x66666600,
    0xAA444444, 0x54A854A8, 0x95809580, 0x96969600, 0xA85454A8, 0x80959580, 0xAA141414, 0x96960000, 0xAAAA1414, 0xA05050A0, 0xA0A5A5A0, 0x96000000,
    0x40804080, 0xA9A8A9A8, 0xAAAAAA44, 0x2A4A5254)
BC7_A2 = (15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2,
          15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6, 6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15)
BC7_A3A = (3, 3, 15, 15, 8, 3, 15, 15, 8, 8, 6, 6, 6, 5, 3, 3, 3, 3, 8, 15, 3, 3, 6, 10, 5, 8, 8, 6, 8, 5, 15, 15,
           8, 15, 3, 5, 6, 10, 8, 15, 15, 3, 15, 5, 15, 15, 15, 15, 3, 15, 5, 5, 5, 8, 5, 10, 5, 10, 8, 13, 15, 12, 3, 3)
BC7_A3B = (15, 8, 8, 3, 15, 15, 3, 8, 15, 15, 15, 15, 15, 15, 15, 8, 15, 8, 15, 3, 15, 8, 15, 8, 3, 15, 6, 10, 15, 15, 10, 8,
           15, 3, 15, 10, 10, 8, 9, 10, 6, 15, 8, 15, 3, 6, 6, 8, 15, 3, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 3, 15, 15, 8)


def decode_bc7_block(block: bytes) -> np.ndarray:
    """One 16-byte BC7 block -> (16, 4) uint8 RGBA, texel t = 4 py + px (plain Python: the test-side statement of the format)."""
    v = int.from_bytes(block, "little")
    out = np.zeros((16, 4), dtype=np.uint8)
    mode = 0
    while mode < 8 and not (v >> mode) & 1:
        mode += 1
    if mode == 8:
        return out                                                     # reserved: all zero
    ns, pb, rb, isb, cb, ab, epb, spb, ib, ib2 = BC7_MODES[mode]
    pos = mode + 1

    def take(width):
        nonlocal pos
        x = (v >> pos) & ((1 << width) - 1)
        pos += width
        return x

    partition, rotation, idxsel = take(pb), take(rb), take(isb)
    ep = np.zeros((2 * ns, 4), dtype=np.int64)
    for ch in range(3):
        for e in range(2 * ns):
            ep[e, ch] = take(cb)
    for e in range(2 * ns):
        ep[e, 3] = take(ab) if ab else 0
    cbits, abits = cb, ab
    if epb:
        for e in range(2 * ns):
            p = take(1)
            ep[e] = (ep[e] << 1) | p
        cbits, abits = cb + 1, (ab + 1 if ab else 0)
    if spb:
        for s in range(ns):
            p = take(1)
            ep[2 * s] = (ep[2 * s] << 1) | p
            ep[2 * s + 1] = (ep[2 * s + 1] << 1) | p
        cbits, abits = cb + 1, (ab + 1 if ab else 0)
    ep[:, :3] = (ep[:, :3] << (8 - cbits)) | (ep[:, :3] >> (2 * cbits - 8))
    ep[:, 3] = ((ep[:, 3] << (8 - abits)) | (ep[:, 3] >> (2 * abits - 8))) if ab else 255
    subset = [0] * 16 if ns == 1 else [(BC7_P2[partition] >> t) & 1 for t in range(16)] if ns == 2 else [(BC7_P3[partition] >> (2 * t)) & 3 for t in range(16)]
    anchors = {0} if ns == 1 else {0, BC7_A2[partition]} if ns == 2 else {0, BC7_A3A[partition], BC7_A3B[partition]}
    primary = [take(ib - 1 if t in anchors else ib) for t in range(16)]
    secondary = [take(ib2 - 1 if t == 0 else ib2) for t in range(16)] if ib2 else None
    for t in range(16):
        e0, e1 = ep[2 * subset[t]], ep[2 * subset[t] + 1]
        ci, cw = (primary[t], ib)
        ai, aw = (primary[t], ib)
        if ib2:
            if idxsel:
                ci, cw, ai, aw = secondary[t], ib2, primary[t], ib
            else:
                ai, aw = secondary[t], ib2
        wc, wa = BC7_WEIGHTS[cw][ci], BC7_WEIGHTS[aw][ai]
        rgba = [int(((64 - wc) * e0[k] + wc * e1[k] + 32) >> 6) for k in range(3)] + [int(((64 - wa) * e0[3] + wa * e1[3] + 32) >> 6)]
        if rotation:
            rgba[rotation - 1], rgba[3] = rgba[3], rgba[rotation - 1]
        out[t] = rgba
    assert pos == 128, (mode, pos)
    return out


def _encode_bc7_mode6(blk: np.ndarray) -> np.ndarray:
    """(n, 16, 4) uint8 -> (n, 16) bytes, mode 6 only (one subset, 7-bit RGBA endpoints + p bit, 4-bit indices): the bounding box of the block
    as end points, every texel on the nearest of the 16 palette entries; the anchor texel keeps its top index bit clear by swapping the ends."""
    out = np.zeros((len(blk), 16), dtype=np.uint8)
    w = np.array(BC7_WEIGHTS[4], dtype=np.int64)
    for n, b in enumerate(blk.astype(np.int64)):
        lo, hi = b.min(axis=0), b.max(axis=0)
        q0, q1 = lo >> 1, hi >> 1                                     # 7 bits; p bits 0 / 1 keep the ends inside [lo, hi]
        p0, p1 = 0, 1
        e0, e1 = (q0 << 1) | p0, (q1 << 1) | p1
        pal = ((64 - w)[:, None] * e0[None, :] + w[:, None] * e1[None, :] + 32) >> 6              # (16, 4)
        idx = ((b[:, None, :] - pal[None, :, :]) ** 2).sum(axis=-1).argmin(axis=-1)
        if idx[0] >= 8:
            q0, q1, p0, p1, idx = q1, q0, p1, p0, 15 - idx
        v, pos = 1 << 6, 7
        for ch in range(4):
            for q in (q0, q1):
                v |= int(q[ch]) << pos
                pos += 7
        v |= p0 << pos; pos += 1
        v |= p1 << pos; pos += 1
        for t in range(16):
            width = 3 if t == 0 else 4
            v |= int(idx[t]) << pos
            pos += width
        assert pos == 128
        out[n] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    return out
