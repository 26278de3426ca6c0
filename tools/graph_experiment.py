"""Upper bound of what CUDA-graph submission could buy: steady-state frames of a config, launched (a) call by call through the C ABI and
(b) as ONE captured graph of the same stream work, replayed.  The replay repeats the SAME frame (kernel arguments are frozen at capture),
so (b) is an experiment that measures launch / inter-kernel gaps, not a product path.   python tools/graph_experiment.py c1|c3|c5 [frames]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rend3_b200 import configs, load_cuda_backend  # noqa: E402
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ev, res = {"c1": configs.config1, "c3": configs.config3, "c5": configs.config5}[name]()
    b = load_cuda_backend(0)
    s = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", 0))
    g = BaseRenderGraph(b)
    settings = BaseRenderGraphSettings(clear_color=(0.1, 0.05, 0.1, 1.0))
    g.add_to_graph(ev, res, 1, settings, upload=True)
    for _ in range(4):
        g.add_to_graph(ev, res, 1, settings, upload=False)
    b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = b.launch_count()
    t0 = time.perf_counter()
    e0.record(s)
    for _ in range(frames):
        g.add_to_graph(ev, res, 1, settings, upload=False)
    e1.record(s)
    host_ms = (time.perf_counter() - t0) * 1e3 / frames
    b.sync()
    stream_ms = e0.elapsed_time(e1) / frames
    launches = (b.launch_count() - l0) // frames
    # the culling buffers ping-pong with period 2: capture TWO consecutive frames, replay the pair
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph, stream=s, capture_error_mode="relaxed"):
            g.add_to_graph(ev, res, 1, settings, upload=False)
            g.add_to_graph(ev, res, 1, settings, upload=False)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0.record(s)
        with torch.cuda.stream(s):
            for _ in range(frames // 2):
                graph.replay()
        e1.record(s)
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / (2 * (frames // 2))
    except Exception as e:   # noqa: BLE001
        graph_ms = None
        print("capture failed:", e)
    print(f"{name}: {launches} launches/frame | call-by-call: {stream_ms:.3f} ms/frame on the stream (host issue time {host_ms:.3f} ms/frame) | "
          f"one graph per 2 frames, replayed: {graph_ms if graph_ms is None else round(graph_ms, 3)} ms/frame")


if __name__ == "__main__":
    main()
