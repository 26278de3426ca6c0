"""Mirror of rend3-test's TestRunner (rend3-test/src/runner.rs:100-187, helpers.rs:27-131) on top
of a `Backend`, so the parity tests can be written like rend3-test/tests/*.rs."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import glam
from .backend import Backend
from .routines import BaseRenderGraph, BaseRenderGraphSettings
from .world import LEFT, Camera, DirectionalLight, MeshBuilder, Object, PbrMaterial, Renderer


class TestRunner:
    __test__ = False  # not a pytest class

    def __init__(self, backend: Backend, handedness: str = LEFT):
        self.renderer = Renderer(handedness)
        self.backend = backend
        self.base_rendergraph = BaseRenderGraph(backend)
        self.last_eval = None

    # ---- helpers.rs:27-131
    def add_directional_light(self, direction) -> int:
        return self.renderer.add_directional_light(
            DirectionalLight(color=(1, 1, 1), resolution=256, distance=5.0, intensity=1.0, direction=tuple(direction))
        )

    def add_unlit_material(self, color) -> int:
        return self.renderer.add_material(PbrMaterial(albedo_value=tuple(color), unlit=True))

    def add_lit_material(self, color) -> int:
        return self.renderer.add_material(PbrMaterial(albedo_value=tuple(color), unlit=False))

    def plane(self, material: int, transform) -> int:
        mesh = (
            MeshBuilder.new([(-1, -1, 0), (-1, 1, 0), (1, 1, 0), (1, -1, 0)], LEFT).with_indices([0, 2, 1, 0, 3, 2]).build()
        )
        return self.renderer.add_object(Object(self.renderer.add_mesh(mesh), material, transform))

    def cube(self, material: int, transform) -> int:
        return self.renderer.add_object(Object(self.renderer.add_mesh(cube_mesh()), material, transform))

    # ---- runner.rs:121-187
    def render_frame(self, size: int = 64, samples: int = 1, resolution: Tuple[int, int] = None,
                     settings: BaseRenderGraphSettings = BaseRenderGraphSettings()) -> np.ndarray:
        res = resolution or (size, size)
        if resolution is not None:
            self.renderer.set_aspect_ratio(res[0] / res[1])
        ev = self.renderer.evaluate()
        self.last_eval = ev
        self.base_rendergraph.add_to_graph(ev, res, samples, settings, srgb_target=True)
        return self.backend.readback_ldr()


def cube_positions():
    """helpers.rs:78-109."""
    return [
        (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1),
        (-1, 1, -1), (1, 1, -1), (1, -1, -1), (-1, -1, -1),
        (1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1),
        (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), (-1, -1, -1),
        (1, 1, -1), (-1, 1, -1), (-1, 1, 1), (1, 1, 1),
        (1, -1, 1), (-1, -1, 1), (-1, -1, -1), (1, -1, -1),
    ]


def cube_indices():
    """helpers.rs:111-118."""
    out = []
    for f in range(6):
        b = 4 * f
        out += [b, b + 1, b + 2, b + 2, b + 3, b]
    return out


def cube_mesh():
    return MeshBuilder.new(cube_positions(), LEFT).with_indices(cube_indices()).build()
