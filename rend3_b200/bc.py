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
                      TEXFMT_BC3_RGBA_UNORM_SRGB, TEXFMT_BC4_R_SNORM, TEXFMT_BC4_R_UNORM, TEXFMT_BC5_RG_SNORM, TEXFMT_BC5_RG_UNORM)

# name -> (unorm / snorm format, sRGB format or None, bytes per block)
BLOCK_FORMATS = {
    "bc1": (TEXFMT_BC1_RGBA_UNORM, TEXFMT_BC1_RGBA_UNORM_SRGB, 8),
    "bc2": (TEXFMT_BC2_RGBA_UNORM, TEXFMT_BC2_RGBA_UNORM_SRGB, 16),
    "bc3": (TEXFMT_BC3_RGBA_UNORM, TEXFMT_BC3_RGBA_UNORM_SRGB, 16),
    "bc4": (TEXFMT_BC4_R_UNORM, None, 8),
    "bc4s": (TEXFMT_BC4_R_SNORM, None, 8),
    "bc5": (TEXFMT_BC5_RG_UNORM, None, 16),
    "bc5s": (TEXFMT_BC5_RG_SNORM, None, 16),
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
    if name == "bc1":
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
