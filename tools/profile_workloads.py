"""Small drivers for ncu captures (see profiles/README.md):  python tools/profile_workloads.py cullbake|frame|c1 [iters]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rend3_b200 import load_cuda_backend  # noqa: E402
from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL  # noqa: E402
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings, per_camera_header  # noqa: E402
from rend3_b200.scenes import cloud_camera, cube_field_scene, object_cloud_records  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "cullbake"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    b = load_cuda_backend(0)
    if mode == "cullbake":
        n = int(os.environ.get("R3_OBJECTS", "10000000"))
        rec = object_cloud_records(n, seed=4)
        header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n)
        b.set_objects(rec)
        for _ in range(iters):
            b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
        b.sync()
        print("visible", b.visible_count(CAMERA_VIEWPORT))
    else:
        if mode == "frame":
            res = (3840, 2160)
            ev = cube_field_scene(n_objects=4400, seed=5, resolution=res, extent=30.0, pull_back=7.0, n_point_lights=64, n_dir_lights=4,
                                  shadow_resolution=2048, shadow_distance=200.0, subdivisions=(2, 3, 3, 4), scale_range=(0.6, 2.4), slabs=True)
        elif mode == "c3":
            from rend3_b200 import configs
            ev, res = configs.config3()
        else:
            res = (1920, 1080)
            ev = cube_field_scene(n_objects=10_000, seed=1, resolution=res)
        g = BaseRenderGraph(b)
        for i in range(iters):
            g.add_to_graph(ev, res, 1, BaseRenderGraphSettings(clear_color=(0.1, 0.05, 0.1, 1.0)), upload=(i == 0))
        b.sync()
        print("stats", b.forward_stats(), "visible", b.visible_count(CAMERA_VIEWPORT), "launches", b.launch_count())


if __name__ == "__main__":
    main()
