"""The uncompressed texture formats beyond RGBA8 / RGBA32F / R8 / RG8 (include/r3_layouts.h, formats 17-30): what rend3-gltf's ktx2 loader can
hand to add_texture_2d besides the common ones (rend3-gltf/src/lib.rs:1195-1285).  `pack` turns an RGBA8 source image into the stored texels
(the scenes' and tests' stand-in for the asset), `unpack` is the float64 statement of how the sampler reads them back (missing channels
(0, 0, 1)) that the tests hold the oracle and the CUDA kernels against.  Nothing here runs on the product path."""
from __future__ import annotations

import numpy as np

from . import layouts as L

# name -> (format, bytes per texel)
STORAGE = {
    "r8s": (L.TEXFMT_R8_SNORM, 1), "rg8s": (L.TEXFMT_RG8_SNORM, 2), "rgba8s": (L.TEXFMT_RGBA8_SNORM, 4),
    "bgra8": (L.TEXFMT_BGRA8_UNORM, 4), "bgra8_srgb": (L.TEXFMT_BGRA8_UNORM_SRGB, 4), "rgb10a2": (L.TEXFMT_RGB10A2_UNORM, 4),
    "r16f": (L.TEXFMT_R16_FLOAT, 2), "rg16f": (L.TEXFMT_RG16_FLOAT, 4), "rgba16f": (L.TEXFMT_RGBA16_FLOAT, 8),
    "r32f": (L.TEXFMT_R32_FLOAT, 4), "rg32f": (L.TEXFMT_RG32_FLOAT, 8),
    "r16": (L.TEXFMT_R16_UNORM, 2), "rg16": (L.TEXFMT_RG16_UNORM, 4), "rgba16": (L.TEXFMT_RGBA16_UNORM, 8),
}
_CHANNELS = {"r8s": 1, "rg8s": 2, "rgba8s": 4, "r16f": 1, "rg16f": 2, "rgba16f": 4, "r32f": 1, "rg32f": 2, "r16": 1, "rg16": 2, "rgba16": 4}


def pack(name: str, rgba8: np.ndarray) -> np.ndarray:
    """(h, w, 4) uint8 -> the level's texels as a flat uint8 array."""
    src = np.ascontiguousarray(rgba8, dtype=np.uint8)
    if name in ("r8s", "rg8s", "rgba8s"):            # v - 128, so the image also holds -128 (which reads as -1, like -127)
        out = (src[..., : _CHANNELS[name]].astype(np.int16) - 128).astype(np.int8)
    elif name in ("bgra8", "bgra8_srgb"):
        out = src[..., [2, 1, 0, 3]]
    elif name == "rgb10a2":
        c = src.astype(np.uint32)
        ten = (c[..., :3] << 2) | (c[..., :3] >> 6)
        out = (ten[..., 0] | (ten[..., 1] << 10) | (ten[..., 2] << 20) | ((c[..., 3] >> 6) << 30)).astype("<u4")
    elif name in ("r16f", "rg16f", "rgba16f"):       # an HDR-ish range with negative values and a few subnormals
        v = src[..., : _CHANNELS[name]].astype(np.float64) / 255.0 * 6.0 - 1.0
        out = np.where(src[..., : _CHANNELS[name]] == 7, 3.0e-6, v).astype("<f2")
    elif name in ("r32f", "rg32f"):
        out = (src[..., : _CHANNELS[name]].astype(np.float64) / 255.0 * 6.0 - 1.0).astype("<f4")
    elif name in ("r16", "rg16", "rgba16"):
        out = (src[..., : _CHANNELS[name]].astype(np.uint16) * 257 ^ 0x0055).astype("<u2")   # not only multiples of 257
    else:
        raise ValueError(f"unknown storage {name!r}")
    return np.ascontiguousarray(out).view(np.uint8).reshape(-1)


def unpack(name: str, data: np.ndarray, width: int, height: int) -> np.ndarray:
    """Stored texels of one level -> (height, width, 4) float64 as the sampler returns them."""
    raw = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros((height, width, 4))
    out[..., 3] = 1.0
    if name in ("r8s", "rg8s", "rgba8s"):
        n = _CHANNELS[name]
        out[..., :n] = np.maximum(raw.view(np.int8).reshape(height, width, n).astype(np.float64), -127.0) / 127.0
    elif name in ("bgra8", "bgra8_srgb"):
        c = raw.reshape(height, width, 4).astype(np.float64) / 255.0
        out = c[..., [2, 1, 0, 3]]
        if name == "bgra8_srgb":
            rgb = out[..., :3]
            out[..., :3] = np.where(rgb > 0.04045, ((rgb + 0.055) / 1.055) ** 2.4, rgb / 12.92)
    elif name == "rgb10a2":
        v = raw.view("<u4").reshape(height, width).astype(np.uint64)
        for k in range(3):
            out[..., k] = ((v >> (10 * k)) & 1023) / 1023.0
        out[..., 3] = (v >> 30) / 3.0
    elif name in ("r16f", "rg16f", "rgba16f"):
        n = _CHANNELS[name]
        out[..., :n] = raw.view("<f2").reshape(height, width, n).astype(np.float64)
    elif name in ("r32f", "rg32f"):
        n = _CHANNELS[name]
        out[..., :n] = raw.view("<f4").reshape(height, width, n).astype(np.float64)
    else:
        n = _CHANNELS[name]
        out[..., :n] = raw.view("<u2").reshape(height, width, n).astype(np.float64) / 65535.0
    return out
