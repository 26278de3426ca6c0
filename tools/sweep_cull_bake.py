"""Time the three modes of the cull + bake stream on one GPU: python tools/sweep_cull_bake.py [objects]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rend3_b200 import load_cuda_backend  # noqa: E402
from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL  # noqa: E402
from rend3_b200.routines import per_camera_header  # noqa: E402
from rend3_b200.scenes import cloud_camera, object_cloud_records  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10_000_000
    b = load_cuda_backend(0)
    b.set_objects(object_cloud_records(n, seed=4))
    header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n)
    for mode, name in ((CB_BAKE | CB_CULL, "bake+cull"), (CB_CULL, "cull-only"), (CB_BAKE, "bake-only")):
        for _ in range(5):
            b.object_uniform_upload(CAMERA_VIEWPORT, header, mode)
        b.sync()
        t0 = time.perf_counter()
        iters = 30
        for _ in range(iters):
            b.object_uniform_upload(CAMERA_VIEWPORT, header, mode)
        b.sync()
        ms = (time.perf_counter() - t0) / iters * 1e3
        print(f"{name:10s} {ms:.4f} ms  {n / ms / 1e6:.2f} G objects/s", flush=True)


if __name__ == "__main__":
    main()
