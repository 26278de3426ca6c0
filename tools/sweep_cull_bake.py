"""Sweep the cull_bake tuning knobs on one GPU: python tools/sweep_cull_bake.py  (spawns one process per variant)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from rend3_b200 import load_cuda_backend
    from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL
    from rend3_b200.routines import per_camera_header
    from rend3_b200.scenes import cloud_camera, object_cloud_records

    n = int(os.environ.get("R3_OBJECTS", "10000000"))
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
        print(f"WT={os.environ.get('R3_CB_WT')} MINB={os.environ.get('R3_CB_MINB')} {name:10s} {ms:.4f} ms  {n / ms / 1e6:.2f} Gobj/s", flush=True)
else:
    for wt in (2, 4, 8):
        for minb in (3, 4):
            env = dict(os.environ, R3_CB_WT=str(wt), R3_CB_MINB=str(minb))
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
