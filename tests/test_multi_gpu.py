"""Two-rank check of the NVLink peer-memory exchange (r3_exchange_*): needs two GPUs, skipped otherwise.
Run by hand with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["R3_ROOT"])
from rend3_b200 import load_cuda_backend
from rend3_b200.backend import CAMERA_VIEWPORT, CB_BAKE, CB_CULL
from rend3_b200.parallel import VisibilityExchange, shard_range
from rend3_b200.routines import per_camera_header
from rend3_b200.scenes import cloud_camera, object_cloud_records

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = 1_000_003                                  # ragged: the last shard is shorter and ends inside a 32-object word
rec = object_cloud_records(n_total, seed=8)
lo, hi = shard_range(n_total, rank, world)
b = load_cuda_backend(local)
b.set_objects(rec[lo:hi])
per_rank = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
ex = VisibilityExchange(b, CAMERA_VIEWPORT, per_rank, rank, world)
header = per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, hi - lo)
for step in range(3):                                # repeated steps overwrite the rows in place
    b.object_uniform_upload(CAMERA_VIEWPORT, header, CB_BAKE | CB_CULL)
b.sync()
dist.barrier()
words = ex.gathered(f"cuda:{local}").cpu().numpy().view(np.uint32)          # (world, words_per_rank)
merged = []
for r in range(world):
    rlo, rhi = shard_range(n_total, r, world)
    bits = np.unpackbits(words[r].view(np.uint8), bitorder="little")[: rhi - rlo]
    merged.append(np.nonzero(bits)[0].astype(np.int64) + rlo)
merged = np.concatenate(merged)
# reference: the whole set culled by this rank alone
full = load_cuda_backend(local)
full.set_objects(rec)
full.object_uniform_upload(CAMERA_VIEWPORT, per_camera_header(cloud_camera(), CAMERA_VIEWPORT, (1920, 1080), 1, n_total), CB_CULL)
want = full.readback_visible(CAMERA_VIEWPORT).astype(np.int64)
assert np.array_equal(merged, want), (len(merged), len(want))
assert 0 < len(want) < n_total
ex.close()
dist.barrier()
if rank == 0:
    print("EXCHANGE_OK", len(want))
'''


def test_two_rank_peer_memory_exchange(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, R3_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                          str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "EXCHANGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
