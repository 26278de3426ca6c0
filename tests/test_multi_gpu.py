"""Two-rank checks of the multi-GPU paths (need two GPUs, skipped otherwise):
  * object-range shards of the cull + bake with the NVLink peer-memory exchange (r3_exchange_*): the global visible list the consumer
    kernel builds on EVERY rank, chained on the epoch flags without a host barrier, against the CPU ORACLE culling the whole set;
  * the forward pass split in row tiles with the shadow maps split by light (r3_peer_*): the frame assembled on rank 0 against the
    single-GPU frame (bit for bit) and against the oracle (1e-4).
Run by hand with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

EXCHANGE_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["R3_ROOT"])
from rend3_b200 import load_cuda_backend
from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL
from rend3_b200.parallel import VisibilityExchange, shard_range
from rend3_b200.routines import per_camera_header
from rend3_b200.scenes import cloud_camera, object_cloud_records
from oracle import load_oracle_backend

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = 1_000_003                                  # ragged: the last shard is shorter and ends inside a 32-object word
rec = object_cloud_records(n_total, seed=8)
ranges = [shard_range(n_total, r, world) for r in range(world)]
lo, hi = ranges[rank]
b = load_cuda_backend(local)
b.set_objects(rec[lo:hi])
per_rank = max(h - l for l, h in ranges)
ex = VisibilityExchange(b, CAMERA_VIEWPORT, per_rank, rank, world, rank_objects=[h - l for l, h in ranges], rank_base=[l for l, h in ranges])
cams = [cloud_camera(), cloud_camera(pull_back=400.0), cloud_camera(pull_back=40.0)]
orc = load_oracle_backend()
orc.set_objects(rec)
for step, cam in enumerate(cams * 2):                # six epochs, no host barrier between them: both parities are reused
    b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cam, CAMERA_VIEWPORT, (1920, 1080), 1, hi - lo), CB_BAKE | CB_CULL)
    ex.merge()                                        # consumer kernels wait for the peers' epoch flags on the device
    if step in (2, 5):
        got = ex.merged(f"cuda:{local}").cpu().numpy().view(np.uint32).astype(np.int64)
        orc.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cam, CAMERA_VIEWPORT, (1920, 1080), 1, n_total), CB_CULL)
        want = orc.readback_visible(CAMERA_VIEWPORT).astype(np.int64)
        assert np.array_equal(got, want), (step, len(got), len(want))
        assert 0 < len(want) < n_total
        ex.count()                                   # the light consumer on the same epoch: per-shard counts, no list
        counts = ex.counts()
        assert int(counts[-1]) == len(want) and [int(x) for x in counts[:-1]] == [int(((want >= l) & (want < h)).sum()) for l, h in ranges]
# forty more epochs without ever joining the consumers (they lag behind the culls on the side stream): the acknowledgement / back-pressure
# protocol must neither deadlock nor let a row be overwritten under a reader — the last epoch's list is still exact
for step in range(40):
    cam = cams[step % 3]
    b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cam, CAMERA_VIEWPORT, (1920, 1080), 1, hi - lo), CB_BAKE | CB_CULL)
    (ex.merge if step % 2 else ex.count)()
b.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cams[1], CAMERA_VIEWPORT, (1920, 1080), 1, hi - lo), CB_BAKE | CB_CULL)
ex.merge()
got = ex.merged(f"cuda:{local}").cpu().numpy().view(np.uint32).astype(np.int64)
orc.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cams[1], CAMERA_VIEWPORT, (1920, 1080), 1, n_total), CB_CULL)
want = orc.readback_visible(CAMERA_VIEWPORT).astype(np.int64)
assert np.array_equal(got, want), ("after 41 unjoined epochs", len(got), len(want))
# the raw rows, for a host reader: stream sync + barrier, then the words of every rank equal the oracle's bits
b.sync()
dist.barrier()
words = ex.gathered(f"cuda:{local}").cpu().numpy().view(np.uint32)
mask = np.zeros(n_total, dtype=bool)
mask[want] = True
for r, (l, h) in enumerate(ranges):
    bits = np.unpackbits(words[r].view(np.uint8), bitorder="little")[: h - l].astype(bool)
    assert np.array_equal(bits, mask[l:h]), r
assert ex.verify_against_nccl(torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local)), f"cuda:{local}")
ex.close()
dist.barrier()
if rank == 0:
    print("EXCHANGE_OK", len(want))
'''

FORWARD_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["R3_ROOT"])
from rend3_b200 import load_cuda_backend
from rend3_b200.backend import CAMERA_VIEWPORT
from rend3_b200.parallel import ForwardSplit, tile_rows
from rend3_b200.routines import BaseRenderGraph, BaseRenderGraphSettings
from rend3_b200.scenes import cube_field_scene
from oracle import load_oracle_backend

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = (480, 270)
ev = cube_field_scene(n_objects=2500, seed=61, resolution=res, n_dir_lights=3, n_point_lights=6, shadow_resolution=512, shadow_distance=150.0, pull_back=10.0,
                      extent=24.0, subdivisions=(1, 2), material_count=4)
settings = BaseRenderGraphSettings(clear_color=(0.1, 0.05, 0.1, 1.0), ambient_color=(0.02, 0.02, 0.02, 1.0))
b = load_cuda_backend(local)
stream = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local))
graph = BaseRenderGraph(b)
split = ForwardSplit(b, stream, dev, rank, world, res, len(ev.shadows), root=0)
split.bind_scene(ev)
rows = tile_rows(res[1], rank, world)
for frame in range(3):                               # three frames back to back: the flags are the only flow control
    graph.add_to_graph(ev, res, 1, settings, upload=(frame == 0), scissor_rows=rows, shadow_filter=split.owns_shadow,
                       after_shadows=split.send_shadow_maps, before_resolve=split.wait_shadow_maps, after_target=split.begin_frame, tonemap=False)
    split.exchange_rows(rows)
b.sync()
dist.barrier()
# SURVEY 8e "triangle cull: shard by batch": every rank tested only its run of the viewport's workgroups, yet all of them must hold the
# same visibility bits, draw records and index lists afterwards
import hashlib
digest = hashlib.sha256()
for part in (0, 1):
    digest.update(b.readback_culling_results(CAMERA_VIEWPORT, part).tobytes())
    dc = b.readback_draw_calls(CAMERA_VIEWPORT, part)
    digest.update(dc.tobytes())
    idx = b.readback_indices(CAMERA_VIEWPORT, part)
    for r in range(len(dc)):
        b0, cnt = int(dc[r]["base_index"]), int(dc[r]["vertex_count"])
        digest.update(idx[b0:b0 + cnt].tobytes())
digests = [None] * world
dist.all_gather_object(digests, digest.hexdigest())
assert len(set(digests)) == 1, "the ranks disagree on the viewport's culled lists"
assert int(b.readback_draw_calls(CAMERA_VIEWPORT, 0)["vertex_count"].sum()) > 3000
if rank == 0:
    got16 = b.readback_hdr_f16()
    one = load_cuda_backend(local)
    g1 = BaseRenderGraph(one)
    for frame in range(3):
        g1.add_to_graph(ev, res, 1, settings, upload=(frame == 0))
    want16 = one.readback_hdr_f16()
    assert np.array_equal(got16.view(np.uint16), want16.view(np.uint16)), "assembled frame differs from the single-GPU frame"
    assert np.array_equal(b.readback_ldr(), one.readback_ldr())
    w, h = ev.shadow_target_size
    assert np.array_equal(b.readback_shadow_atlas(w, h).view(np.uint32), one.readback_shadow_atlas(w, h).view(np.uint32)), "exchanged shadow atlas differs"
    orc = load_oracle_backend()
    go = BaseRenderGraph(orc)
    for frame in range(3):
        go.add_to_graph(ev, res, 1, settings, upload=(frame == 0))
    o16 = orc.readback_hdr_f16().astype(np.float32)
    g = got16.astype(np.float32)
    ulp = np.maximum(np.abs(o16) * 2.0 ** -10, 2.0 ** -24)
    assert np.all(np.abs(g - o16) <= ulp + 1e-4 * np.maximum(1.0, np.abs(o16))), "assembled frame differs from the oracle"
    assert np.array_equal(b.readback_shadow_atlas(w, h).view(np.uint32), orc.readback_shadow_atlas(w, h).view(np.uint32))
split.close()
dist.barrier()
if rank == 0:
    print("FORWARD_SPLIT_OK")
'''


def _run(tmp_path, name, source, port, marker):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / name
    script.write_text(source)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, R3_ROOT=root, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                          str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and marker in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_two_rank_peer_memory_exchange_matches_the_oracle(tmp_path):
    _run(tmp_path, "exchange_worker.py", EXCHANGE_WORKER, 29533, "EXCHANGE_OK")


def test_two_rank_forward_split_matches_single_gpu_and_oracle(tmp_path):
    _run(tmp_path, "forward_worker.py", FORWARD_WORKER, 29534, "FORWARD_SPLIT_OK")
