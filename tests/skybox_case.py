"""Skybox scenes: six solid-colour faces seen through a 90-degree camera at the origin looking down an axis — the whole target shows
that face — and a gradient cube for the (s, t) orientation of rule R10."""
import numpy as np

from rend3_b200 import glam
from rend3_b200.runner import TestRunner
from rend3_b200.world import LEFT, Camera

FACE_COLOURS = np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [255, 255, 0, 255], [0, 255, 255, 255], [255, 0, 255, 255]], dtype=np.uint8)
# looking along +X, -X, +Y, -Y, +Z, -Z (up vectors chosen so that look_at is well defined)
LOOK = [((1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 1, 0)), ((0, 1, 0), (0, 0, -1)), ((0, -1, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((0, 0, -1), (0, 1, 0))]


def solid_faces(size=8):
    return [np.broadcast_to(FACE_COLOURS[f], (size, size, 4)).copy() for f in range(6)]


def build(backend, faces, face_index, srgb=False, mips="generated"):
    r = TestRunner(backend, LEFT)
    r.renderer.set_skybox(faces, srgb=srgb, mips=mips)
    direction, up = LOOK[face_index]
    view = glam.look_at_lh(np.zeros(3, dtype=np.float32), np.array(direction, dtype=np.float32), np.array(up, dtype=np.float32))
    r.renderer.set_aspect_ratio(1.0)
    r.renderer.set_camera_data(Camera(("perspective", 90.0, 0.1), view))
    return r
