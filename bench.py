#!/usr/bin/env python
"""bench.py — rend3 hot path on B200: culled objects/s (cull + uniform bake) and shaded Mfrag/s (PBR forward).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic input:
  * headline workload (BASELINE config 4, the one the target metric is quoted on): fused per-object frustum cull +
    object-uniform bake over 10 M object records PER GPU (weak scaling), 1% disabled, visible list ascending;
    at N > 1 the visible sets are all-gathered with NCCL (as 1-bit-per-object words) as north_star asks;
  * `value` = objects culled+baked per second, inputs resident in HBM, CUDA events on the library's stream;
  * `e2e`   = the same through the C ABI with HOST buffers: r3_set_objects (pinned H2D of every record) +
    r3_object_uniform_upload + r3_readback_visible (D2H) inside the timed region;
  * `forward` = BASELINE config 5 (4K, ~500k triangles, 64 point lights + 4 shadow-mapped directional lights):
    whole frames through BaseRenderGraph.add_to_graph; Mfrag/s = fs_main invocations / frame time; at N > 1 the
    screen is split in row tiles, one per rank, and the rgba16f rows are all-gathered.
Inputs (0.8 GB of transforms + spheres read, 1.28 GB of matrices written per step) exceed the 126 MB L2, so no explicit
flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL  # noqa: E402
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings, per_camera_header  # noqa: E402
from rend3_b200.scenes import cloud_camera, cube_field_scene, object_cloud_records  # noqa: E402

METRIC = "culled objects/s (fused frustum cull + object-uniform bake)"
BYTES_PER_OBJECT = 212   # 84 B read (transform 64 + sphere 16 + enabled 4) + 128 B MV/MVP written
BYTES_PER_VISIBLE = 4
NCU_TRAFFIC_PER_OBJECT = (801_563_648 + 1_213_981_000) / 10_000_000   # ncu --set full capture of cull_bake_kernel<bake,cull>, 10 M objects (profiles/)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of this
    workload is a few milliseconds, far shorter than one `nvidia-smi` invocation, so the sampler polls NVML directly
    (nvidia_ml_py: the library nvidia-smi itself reads) every ~0.5 ms and only falls back to nvidia-smi without it."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, uuid=None):
        self.index, self.rows, self.stop, self.nvml, self.handle, self.source = index, [], False, None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if str(uuid).startswith("GPU-") else f"GPU-{uuid}") if uuid else None
            except Exception:
                self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.rows.append((sm, self.max_mhz, mask, time.perf_counter()))
            return
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [x.strip() for x in out.split(",")]
            mask = sum(bit for k, bit in zip(range(2, 6), (0x8, 0x40, 0x20, 0x4)) if len(c) > k and c[k].lower().startswith("active"))
            self.rows.append((float(c[0]), float(c[1]), mask, time.perf_counter()))

    def run(self):
        while not self.stop:
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(0.0002 if self.nvml is not None else 0.1)

    def sample_now(self):
        try:
            self.sample()
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def summary(self):
        """Samples stamped inside [mark_begin, mark_end] (the timed region).  The region is only milliseconds long, so when
        fewer than 3 samples fell inside it the nearest samples of the surrounding warm-up / drain (same workload) are added."""
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0, "source": self.source}
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for r in self.rows if t0 <= r[3] <= t1]
        used = inside
        if len(inside) < 3:
            mid = 0.5 * (t0 + min(t1, self.rows[-1][3]))
            used = sorted(self.rows, key=lambda r: abs(r[3] - mid))[:max(3, len(inside))]
        sm = sorted(r[0] for r in used)
        mask = 0
        for r in used:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": used[0][1], "reasons": [k for k, bit in self.BITS.items() if mask & bit],
                "samples": len(used), "samples_in_timed_region": len(inside), "source": self.source}


class DeviceView:
    """__cuda_array_interface__ wrapper so torch.distributed can move library-owned device memory without a host bounce."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False), "version": 2}


def reference_arm(args):
    """The reference's own CPU arithmetic for the path.  The reference (Rust + wgpu) cannot be built or run here (no
    cargo, no Vulkan ICD: SURVEY 8c), so this arm times the oracle port — uniform_prep.wgsl + batch_objects' frustum filter
    restated in C — with every host thread, on bounded samples of the same workload."""
    import oracle

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.objects, 2_000_000)   # bounded sample: 2 M of the 10 M records per step
    cores = oracle.set_threads(os.cpu_count() or 1)
    rec = object_cloud_records(n, seed=4)
    header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n)
    b = oracle.load_oracle_backend()
    b.set_objects(rec)
    for _ in range(args.warmup):
        b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "objects/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 4: cull + uniform bake over 10 M object records per GPU", "objects_per_step_sampled": n,
                   "note": "reference itself (Rust/wgpu) cannot run here; CPU oracle port of uniform_prep.wgsl + Frustum::contains_sphere"},
        "cpu_baseline": {"value": value, "unit": "objects/s", "cores": cores, "kind": "port", "sample": f"{n} of {args.objects} records per step, OpenMP over all host cores"},
        "e2e": {"value": value, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(n_total):
    import oracle

    oracle.set_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
    n = min(n_total, 2_000_000)
    rec = object_cloud_records(n, seed=4)
    header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n)
    b = oracle.load_oracle_backend()
    b.set_objects(rec)
    b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 50):
        b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": n / dt, "unit": "objects/s", "cores": oracle.set_threads(os.cpu_count() or 1), "kind": "port",
            "sample": f"{n} of {n_total} records, {reps} repetitions, OpenMP over all host cores (oracle/r3_oracle.c)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--objects", type=int, default=10_000_000, help="object records per GPU")
    ap.add_argument("--no-forward", action="store_true")
    ap.add_argument("--forward-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    # keep stdout clean for the single JSON line: anything libraries print (e.g. NCCL's version banner) goes to stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from rend3_b200 import load_cuda_backend   # fails loudly when librend3_b200.so is missing: no CPU fallback

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.objects
    backend = load_cuda_backend(local)
    stream = torch.cuda.ExternalStream(backend.stream(), device=torch.device("cuda", local))

    # ---- inputs: this rank's shard of the object set, built on the host, staged in pinned memory
    rec = object_cloud_records(n, seed=4 + rank)
    pinned = torch.empty(n * 128, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = rec.view(np.uint8).reshape(-1)
    host_rec = pinned.numpy().view(rec.dtype)
    del rec
    header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n)
    backend.set_objects(host_rec)
    vis_host = torch.empty(n, dtype=torch.int32, pin_memory=True)
    gathered_words = torch.empty(world * ((n + 31) // 32), dtype=torch.int32, device=f"cuda:{local}")

    # ---- north-star exchange: every rank ends up with the visible set of all shards, in its 1-bit-per-object form
    # (n/8 bytes per shard instead of 4 B per visible object: fixed size, no count exchange, no host synchronisation).
    # Preferred: the library's compaction kernel stores the words straight into every peer's buffer over NVLink (r3_exchange_*),
    # so no collective kernel runs at all.  If the IPC set-up is not possible on this box the NCCL all-gather (on a high-priority
    # side stream, overlapped with the next cull) takes over; the JSON line says which one ran.
    exchange, exchange_kind = None, "single GPU"
    side = None
    if world > 1:
        try:
            from rend3_b200.parallel import VisibilityExchange
            exchange = VisibilityExchange(backend, CAMERA_VIEWPORT, n, rank, world)
            exchange_kind = "peer-memory stores fused into the compaction kernel (NVLink P2P, CUDA IPC)"
        except Exception as e:   # noqa: BLE001
            print(f"[rank {rank}] peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
            exchange = None
        ok = torch.tensor([1 if exchange is not None else 0], device=f"cuda:{local}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if exchange is not None:
                exchange.close()
            exchange, exchange_kind = None, "NCCL all-gather of the visibility words on a side stream"
            side = torch.cuda.Stream(device=f"cuda:{local}", priority=-1)

    def gather_visible():
        if world == 1 or exchange is not None:
            return   # nothing to launch: the exchange is part of r3_object_uniform_upload
        wptr, wbytes = backend.device_ptr(CAMERA_VIEWPORT, 4)
        done = torch.cuda.Event()
        done.record(stream)
        side.wait_event(done)
        with torch.cuda.stream(side):
            mine = torch.as_tensor(DeviceView(wptr, wbytes, "<i4", 4), device=f"cuda:{local}")
            dist.all_gather_into_tensor(gathered_words, mine)

    def step_resident():
        backend.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
        gather_visible()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None))
    clocks.__enter__()               # polls through warm-up, timed region and drain; only the stamped window is reported
    for _ in range(args.warmup):
        step_resident()
    barrier()
    exchange_verified = None
    if exchange is not None:
        # outside the timed region: the rows the peers stored must equal an NCCL all-gather of the same words
        wptr, wbytes = backend.device_ptr(CAMERA_VIEWPORT, 4)
        with torch.cuda.stream(stream):
            mine = torch.as_tensor(DeviceView(wptr, wbytes, "<i4", 4), device=f"cuda:{local}")
            ref = torch.empty(world * mine.numel(), dtype=torch.int32, device=f"cuda:{local}")
            dist.all_gather_into_tensor(ref, mine)
            got = exchange.gathered(f"cuda:{local}")[:, : mine.numel()].reshape(-1)
            same = torch.tensor([1 if torch.equal(got, ref) else 0], device=f"cuda:{local}")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        exchange_verified = bool(int(same.item()))
        if not exchange_verified:
            print(f"[rank {rank}] peer-memory exchange does not match the NCCL all-gather; using NCCL", file=sys.stderr)
            exchange.close()
            exchange, exchange_kind = None, "NCCL all-gather of the visibility words on a side stream"
            side = torch.cuda.Stream(device=f"cuda:{local}", priority=-1)
        barrier()
    launches0 = backend.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks.mark_begin()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        starts[i].record(stream)
        backend.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
        ends[i].record(stream)
        gather_visible()
    clocks.sample_now()              # GPU busy with the queued steps
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks.mark_end()
    clocks.__exit__()
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = starts[0].elapsed_time(ends[-1]) if world == 1 else wall * 1e3
    launches = backend.launch_count() - launches0
    n_vis = backend.visible_count(CAMERA_VIEWPORT)
    t = torch.tensor([total_ms], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    kern_s = float(np.mean(kernel_ms)) * 1e-3
    achieved = (BYTES_PER_OBJECT * n + BYTES_PER_VISIBLE * n_vis) / kern_s / 1e9

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    e2e_steps = max(3, min(args.steps, 5))
    backend.set_objects(host_rec)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        backend.set_objects(host_rec)                                         # pinned H2D of every record
        backend.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
        nv = backend.visible_count(CAMERA_VIEWPORT)
        backend._call("readback_visible", __import__("ctypes").c_uint32(CAMERA_VIEWPORT), __import__("ctypes").c_void_p(vis_host.data_ptr()),
                      __import__("ctypes").c_uint32(n), None)               # D2H of the visible list into pinned memory
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": world * n / float(te.item()), "unit": "objects/s", "h2d_bytes_per_step": n * 128 + 240, "d2h_bytes_per_step": 4 * nv + 4}

    # ---- forward: BASELINE config 5
    forward = None
    if not args.no_forward:
        res = (3840, 2160)
        ev = cube_field_scene(n_objects=4400, seed=5, resolution=res, extent=30.0, pull_back=7.0, n_point_lights=64, n_dir_lights=4,
                              shadow_resolution=2048, shadow_distance=200.0, subdivisions=(2, 3, 3, 4), scale_range=(0.6, 2.4), slabs=True)
        fb = load_cuda_backend(local)
        fstream = torch.cuda.ExternalStream(fb.stream(), device=torch.device("cuda", local))
        graph = BaseRenderGraph(fb)
        settings = BaseRenderGraphSettings(clear_color=(0.1, 0.05, 0.1, 1.0))
        rows = (res[1] * rank // world, res[1] * (rank + 1) // world)

        n_shadows = len(ev.shadows)

        def merge_shadow_maps():
            # each rank rendered the shadow maps i with i % world == rank into its (cleared) atlas: an integer MAX all-reduce of the
            # depth bits (reverse-Z, >= 0) gives every rank the complete atlas
            ptr, nbytes = fb.device_ptr(CAMERA_VIEWPORT, 5)
            with torch.cuda.stream(fstream):
                atlas = torch.as_tensor(DeviceView(ptr, nbytes, "<i4", 4), device=f"cuda:{local}")
                dist.all_reduce(atlas, op=dist.ReduceOp.MAX)

        def frame(upload):
            if world > 1:
                graph.add_to_graph(ev, res, 1, settings, upload=upload, scissor_rows=rows, shadow_filter=lambda i: i % world == rank,
                                   after_shadows=merge_shadow_maps if n_shadows else None)
                ptr, nbytes = fb.device_ptr(CAMERA_VIEWPORT, 1)
                with torch.cuda.stream(fstream):
                    img = torch.as_tensor(DeviceView(ptr, nbytes), device=f"cuda:{local}")
                    row_bytes = res[0] * 8
                    mine = img[rows[0] * row_bytes:rows[1] * row_bytes]
                    if res[1] % world == 0:
                        dist.all_gather_into_tensor(img, mine.clone())
            else:
                graph.add_to_graph(ev, res, 1, settings, upload=upload)

        split_verified = None
        if world > 1 and res[1] % world == 0:
            # outside the timed region: the frame assembled from the ranks' row tiles and merged shadow maps must equal, bit for bit,
            # the frame one GPU renders alone
            frame(True)
            barrier()
            ptr, nbytes = fb.device_ptr(CAMERA_VIEWPORT, 1)
            got = torch.as_tensor(DeviceView(ptr, nbytes), device=f"cuda:{local}").clone()
            vb = load_cuda_backend(local)
            BaseRenderGraph(vb).add_to_graph(ev, res, 1, settings)
            vb.sync()
            vptr, vbytes = vb.device_ptr(CAMERA_VIEWPORT, 1)
            want = torch.as_tensor(DeviceView(vptr, vbytes), device=f"cuda:{local}")
            same = torch.tensor([1 if torch.equal(got, want) else 0], device=f"cuda:{local}")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            split_verified = bool(int(same.item()))
            vb.close()
        frame(True)
        for _ in range(2):
            frame(False)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = fb.launch_count()
        tw = time.perf_counter()
        f0.record(fstream)
        for _ in range(args.forward_steps):
            frame(False)
        f1.record(fstream)
        barrier()
        frame_ms = (time.perf_counter() - tw) * 1e3 / args.forward_steps
        tf = torch.tensor([frame_ms], device=f"cuda:{local}")
        st = torch.tensor([float(x) for x in fb.forward_stats()[:3]], device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            dist.all_reduce(st, op=dist.ReduceOp.SUM)
        frame_ms = float(tf.item())
        tris = int(fb.readback_draw_calls(CAMERA_VIEWPORT, 0)["vertex_count"].sum()) // 3
        forward = {"workload": "BASELINE config 5: 3840x2160, 4400 meshes / ~500k triangles, 64 point lights + 4 directional lights with 2048^2 shadow maps",
                   "frame_ms": frame_ms, "shaded_mfrag_s": float(st[2].item()) / frame_ms / 1e3, "raster_mfrag_s": float(st[1].item()) / frame_ms / 1e3,
                   "shaded_fragments": int(st[2].item()), "depth_passing_fragments": int(st[1].item()), "triangles_after_cull": tris,
                   "gpu_launches_per_frame": (fb.launch_count() - l0) // max(args.forward_steps, 1),
                   "split": f"{world} row tiles (rgba16f rows all-gathered), shadow maps split by light and merged with a MAX all-reduce" if world > 1 else "single GPU",
                   "split_equals_single_gpu_frame": split_verified}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "objects/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: fused frustum cull + uniform bake, 10 M object records per GPU (128 B std430 records, 1% disabled)",
                       "objects_per_gpu": n, "visible_fraction": n_vis / n, "parallelism": f"object-range shards x{world}; visible set (1 bit/object) exchanged by {exchange_kind}" if world > 1 else "single GPU",
                       "exchange_verified_against_nccl": exchange_verified,
                       "l2": "inputs (0.8 GB) + outputs (1.28 GB) per step exceed the 126 MB L2; no explicit flush"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_PER_OBJECT * n,
                         "traffic_source": "profiles/r1_ncu_cull_bake_10M.txt: dram__bytes_read.sum + dram__bytes_write.sum of one 10 M-object launch, scaled per object",
                         "kernel": "cull_bake_kernel<bake,cull>", "kernel_ms": kern_s * 1e3, "algorithmic_bytes_per_launch": BYTES_PER_OBJECT * n + BYTES_PER_VISIBLE * n_vis,
                         "peak_source": peak_src},
            "cpu_baseline": cpu_baseline(n),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(), "forward": forward,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
