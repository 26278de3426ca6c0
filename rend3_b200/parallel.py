"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed for the exchange.

The reference is single-device; these are the only collectives the B200 build introduces:
  * object cull + bake shards the object array in contiguous index ranges — no data-path collective; the visible
    lists are all-gathered only when a downstream stage needs the global list (`allgather_visible`): counts first,
    then the lists padded to the longest shard.  Rank order == ascending index order, so the merged list is
    bit-identical to the single-GPU list;
  * the forward pass is split sort-first into row tiles (`tile_rows`), every rank keeps the (small) culled
    lists and its rows of the rgba16f target are all-gathered (`allgather_rows`).
Works with the gloo backend on CPU tensors (tests) and NCCL on CUDA tensors (bench.py).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous object-index range [lo, hi) of `rank` (r*N/W .. (r+1)*N/W)."""
    return n * rank // world, n * (rank + 1) // world


def tile_rows(height: int, rank: int, world: int) -> Tuple[int, int]:
    """Pixel-row band rendered by `rank` in the screen-tile split."""
    return height * rank // world, height * (rank + 1) // world


def allgather_visible(local: torch.Tensor, offset: int, n_total: int, world: int) -> torch.Tensor:
    """All-gather per-shard visible lists (shard-local slot ids, ascending) into the global ascending list.
    `local` may live on the CPU (gloo) or on a GPU (NCCL)."""
    dev = local.device
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    padded = torch.zeros(m, dtype=local.dtype, device=dev)
    padded[: local.numel()] = local + offset
    out = [torch.zeros(m, dtype=local.dtype, device=dev) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def allgather_rows(image_rows: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather equally sized row bands of the rgba16f target (band r from rank r) into the full frame."""
    out = torch.empty((world,) + tuple(image_rows.shape), dtype=image_rows.dtype, device=image_rows.device)
    dist.all_gather_into_tensor(out, image_rows.contiguous())
    return out.reshape((-1,) + tuple(image_rows.shape[1:]))


class VisibilityExchange:
    """Every rank ends up with the visibility words (1 bit per object slot) of all shards, written by the cull's own compaction
    kernel straight into the peers' memory over NVLink (r3_exchange_*, include/rend3_b200.h).  torch.distributed only carries
    the 64-byte IPC handles once, at set-up."""

    def __init__(self, backend, camera: int, objects_per_rank: int, rank: int, world: int):
        self.backend, self.camera, self.rank, self.world = backend, camera, rank, world
        handle = backend.exchange_create(camera, world, rank, objects_per_rank)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        backend.exchange_connect(camera, b"".join(handles))
        self.ptr, self.nbytes, self.words_per_rank = backend.exchange_words(camera)
        dist.barrier()   # every rank has mapped every buffer before anybody's cull writes into them

    def gathered(self, device) -> torch.Tensor:
        """(world, words_per_rank) int32 view of the local gathered buffer.  Valid after the ranks synchronised their streams and
        passed a barrier."""
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.world, self.words_per_rank), "typestr": "<i4", "data": (self.ptr, False), "version": 2}
        return torch.as_tensor(v, device=device)

    def close(self):
        dist.barrier()
        self.backend.exchange_destroy(self.camera)
