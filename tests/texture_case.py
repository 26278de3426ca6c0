"""Shared scenes for the texture tests: a full-screen unlit quad whose texture coordinates map texel centres onto pixel
centres (raw identity camera), so the expected image is the texture itself (or one of its mip levels)."""
import numpy as np

from rend3_b200 import glam
from rend3_b200.runner import TestRunner
from rend3_b200.world import LEFT, Camera, MeshBuilder, Object, PbrMaterial, Texture


def checker_texture(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (size, size, 4), dtype=np.uint8)


def srgb_decode(u8: np.ndarray) -> np.ndarray:
    c = u8.astype(np.float64) / 255.0
    out = c.copy()
    out[..., :3] = np.where(c[..., :3] > 0.04045, ((c[..., :3] + 0.055) / 1.055) ** 2.4, c[..., :3] / 12.92)
    return out


def build(backend, texture: Texture, sample_type="linear", uv_scale=1.0):
    r = TestRunner(backend, LEFT)
    pos = [(-1, -1, 0.5), (-1, 1, 0.5), (1, 1, 0.5), (1, -1, 0.5)]
    uv = [((x + 1) / 2, (1 - y) / 2) for x, y, _ in pos]
    mesh = MeshBuilder.new(pos, LEFT).with_indices([0, 1, 2, 0, 2, 3]).with_vertex_texture_coordinates_0(uv).build()
    tex = r.renderer.add_texture_2d(texture)
    ut = np.array([[uv_scale, 0, 0], [0, uv_scale, 0], [0, 0, 1]], dtype=np.float32)   # columns of uv_transform0
    mat = r.renderer.add_material(PbrMaterial(albedo_texture=tex, unlit=True, sample_type=sample_type, uv_transform0=ut))
    r.renderer.add_object(Object(r.renderer.add_mesh(mesh), mat, glam.identity()))
    r.renderer.set_camera_data(Camera(("raw", glam.identity()), glam.identity()))
    return r


def random_bc7_blocks(n: int, seed: int) -> np.ndarray:
    """n random 16-byte BC7 blocks, block i forced into mode i % 9 (8 = the reserved mode: no mode bit in byte 0): any bit pattern is a valid
    block, so random bits reach every partition, rotation, index selection and p bit of every mode."""
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    for i, b in enumerate(blocks):
        mode, v = i % 9, int.from_bytes(bytes(b), "little")
        v = (v & ~0xFF) if mode == 8 else ((v & ~((1 << (mode + 1)) - 1)) | (1 << mode))
        b[:] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    return blocks
