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
