"""Time whole frames of the BASELINE configurations on one GPU: python tools/run_configs.py [c1 c3 c5] [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rend3_b200 import configs, load_cuda_backend  # noqa: E402
from rend3_b200.backend import CAMERA_VIEWPORT  # noqa: E402
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if a.startswith("c")] or ["c1", "c3", "c5"]
    frames = next((int(a) for a in sys.argv[1:] if a.isdigit()), 10)
    for name in names:
        t0 = time.perf_counter()
        ev, res = {"c1": configs.config1, "c3": configs.config3, "c5": configs.config5}[name]()
        gen = time.perf_counter() - t0
        b = load_cuda_backend(0)
        g = BaseRenderGraph(b)
        settings = BaseRenderGraphSettings(clear_color=(0.1, 0.05, 0.1, 1.0))
        t0 = time.perf_counter()
        g.add_to_graph(ev, res, 1, settings, upload=True)
        b.sync()
        first = time.perf_counter() - t0
        for _ in range(3):
            g.add_to_graph(ev, res, 1, settings, upload=False)
        b.sync()
        l0 = b.launch_count()
        t0 = time.perf_counter()
        for _ in range(frames):
            g.add_to_graph(ev, res, 1, settings, upload=False)
        b.sync()
        ms = (time.perf_counter() - t0) / frames * 1e3
        st = b.forward_stats()
        tris = int(b.readback_draw_calls(CAMERA_VIEWPORT, 0)["vertex_count"].sum()) // 3
        n_obj = int((ev.object_live != 0).sum())
        print(f"{name}: {n_obj} objects, {res[0]}x{res[1]}, {len(ev.shadows)} shadow maps | scene gen {gen:.1f}s first frame {first * 1e3:.1f} ms | steady frame {ms:.3f} ms"
              f" ({(b.launch_count() - l0) // frames} launches) | visible objects {b.visible_count(CAMERA_VIEWPORT)} predicted tris {tris} set-up {st[0]}"
              f" rasterised {st[1]} shaded {st[2]} -> {st[2] / ms / 1e3:.1f} Mfrag/s shaded, {n_obj / ms / 1e3:.2f} Mobj/s through the whole frame", flush=True)
        b.set_stage_timing(True)
        for _ in range(3):
            g.add_to_graph(ev, res, 1, settings, upload=False)
        st = b.stage_times()
        b.set_stage_timing(False)
        print("    stage ms/frame: " + ", ".join(f"{k} {v['ms'] / 3:.3f} ({v['launches'] // 3})" for k, v in st.items()), flush=True)
        b.close()


if __name__ == "__main__":
    main()
