"""Shared scene for the blend-routine tests: an opaque unlit background and two unlit transparent triangles that cover the
whole 64x64 target, viewed through a raw identity camera (clip = world position, reverse-Z: larger z is nearer)."""
import numpy as np

from rend3_b200 import glam
from rend3_b200.runner import TestRunner
from rend3_b200.world import BLEND, LEFT, Camera, MeshBuilder, Object, PbrMaterial

BACK = (0.8, 0.6, 0.2, 1.0)
FAR = (0.1, 0.9, 0.3, 0.5)     # transparent, depth 0.3
NEAR = (0.7, 0.2, 0.9, 0.25)   # transparent, depth 0.6


def f16(x):
    return np.float32(np.float16(np.float32(x)))


def blend(src, dst):
    """Rule R8 (oracle/r3_oracle_forward.inc): ALPHA_BLENDING into an rgba16f target."""
    a = np.float32(src[3])
    inv = np.float32(1.0) - a
    rgb = [f16(np.float32(src[k]) * a + np.float32(dst[k]) * inv) for k in range(3)]
    return rgb + [f16(a + np.float32(dst[3]) * inv)]


def build(backend, far_first: bool):
    """far_first=True: the far triangle's object is farther from the camera location (the origin) than the near one, so
    batch_objects draws it first (back to front) and both layers blend.  far_first=False swaps the object distances: the near
    layer is drawn first, writes depth, and the far layer then FAILS the depth test (depth write is on for the blend routine)."""
    r = TestRunner(backend, LEFT)
    tri = MeshBuilder.new([(0.5, -0.5, 0), (-0.5, -0.5, 0), (0, 0.5, 0)], LEFT).build()
    mesh = r.renderer.add_mesh(tri)

    def place(color, z, shift, transparency):
        mat = r.renderer.add_material(PbrMaterial(albedo_value=color, unlit=True, transparency=transparency))
        m = glam.mul(glam.from_translation((shift, 0.0, z)), glam.from_scale((40.0, 40.0, 1.0)))
        r.renderer.add_object(Object(mesh, mat, m))

    place(BACK, 0.1, 0.0, 0)
    place(FAR, 0.3, 3.0 if far_first else 0.0, BLEND)
    place(NEAR, 0.6, 0.0 if far_first else 3.0, BLEND)
    r.renderer.set_camera_data(Camera(("raw", glam.identity()), glam.identity()))
    return r


def expected(far_first: bool):
    dst = [f16(v) for v in BACK]
    if far_first:
        dst = blend(FAR, dst)
    return np.array(blend(NEAR, dst), dtype=np.float32)
