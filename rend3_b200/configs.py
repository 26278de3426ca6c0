"""The five BASELINE.json configurations as seeded synthetic scenes.  Each returns (EvalOutput, resolution)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .runner import cube_mesh
from .scenes import (bulk_object_records, cube_example_camera, cube_field_scene, eval_with_bulk_objects, random_unit_quaternions,
                     subdivided_cube_mesh, trs_matrices)
from .world import LEFT, DirectionalLight, PbrMaterial, PointLight, Renderer

f32 = np.float32


def config1(resolution: Tuple[int, int] = (1920, 1080)):
    """C1: 10k untextured cubes, 1 directional light (2048^2 shadow map), 1080p, single camera (seed 1)."""
    return cube_field_scene(n_objects=10_000, seed=1, resolution=resolution), resolution


def config3(resolution: Tuple[int, int] = (3840, 2160), n_objects: int = 200_000):
    """C3: scene_viewer 'bistro'-shaped synthetic: 200k objects over 1000 mesh slots (cube + subdivided cubes, 12..768
    triangles, mean ~100), centres in a 120x20x80 box with 70% within 3 m of 6 'street' planes, 4 directional lights
    (4 shadow maps 2048^2, distance 100) + 4 point lights r=15, 3840x2160 (seed 3)."""
    rng = np.random.default_rng(3)
    r = Renderer(LEFT, aspect_ratio=resolution[0] / resolution[1])
    ks = [1, 1, 2, 2, 2, 3, 3, 4, 5, 8]                      # 12..768 triangles, mean ~ 124
    built = {k: (cube_mesh() if k == 1 else subdivided_cube_mesh(k)) for k in set(ks)}
    mesh_slots = [r.add_mesh(built[ks[i % len(ks)]]) for i in range(1000)]
    for m in range(8):
        g = 0.25 + 0.5 * m / 7.0
        r.add_material(PbrMaterial(albedo_value=(0.6, g, 1.0 - g, 1.0), roughness_factor=0.35 + 0.05 * m, metallic_factor=0.1 * (m % 3)))
    box = np.array([60.0, 10.0, 40.0])
    centers = rng.uniform(-box, box, (n_objects, 3))
    near = rng.random(n_objects) < 0.7
    plane = rng.integers(0, 6, n_objects)                     # 6 street planes: x = -40, 0, 40 and z = -25, 0, 25
    xs, zs = np.array([-40.0, 0.0, 40.0]), np.array([-25.0, 0.0, 25.0])
    off = rng.uniform(-3.0, 3.0, n_objects)
    on_x = near & (plane < 3)
    on_z = near & (plane >= 3)
    centers[on_x, 0] = xs[plane[on_x]] + off[on_x]
    centers[on_z, 2] = zs[plane[on_z] - 3] + off[on_z]
    scale = rng.uniform(0.15, 0.6, (n_objects, 1)).astype(f32)
    transforms = trs_matrices(centers.astype(f32), random_unit_quaternions(rng, n_objects), scale)
    mesh_ids = np.asarray(mesh_slots, dtype=np.int64)[rng.integers(0, 1000, n_objects)]
    material_ids = rng.integers(0, 8, n_objects).astype(np.uint32)
    r.set_camera_data(cube_example_camera(8.0))               # eye (24, 24, -40) looking into the box
    for d in [(-1.0, -4.0, 2.0), (2.0, -3.0, -1.0), (-2.0, -5.0, -3.0), (1.0, -2.0, 3.0)]:
        r.add_directional_light(DirectionalLight(color=(1, 1, 1), intensity=0.25, direction=d, distance=100.0, resolution=2048))
    for _ in range(4):
        p = rng.uniform(-box, box)
        r.add_point_light(PointLight(position=tuple(p), color=tuple(rng.uniform(0.3, 1.0, 3)), radius=15.0, intensity=4.0))
    rec, loc = bulk_object_records(r, transforms, mesh_ids, material_ids)
    return eval_with_bulk_objects(r, rec, loc, n_objects), resolution


def config5(resolution: Tuple[int, int] = (3840, 2160)):
    """C5: 4K PBR forward, ~500k triangles (4400 subdivided cubes + a 3-slab room), 64 point lights + 4 directional
    lights with 2048^2 shadow maps (seed 5)."""
    ev = cube_field_scene(n_objects=4400, seed=5, resolution=resolution, extent=30.0, pull_back=7.0, n_point_lights=64, n_dir_lights=4,
                          shadow_resolution=2048, shadow_distance=200.0, subdivisions=(2, 3, 3, 4), scale_range=(0.6, 2.4), slabs=True)
    return ev, resolution
