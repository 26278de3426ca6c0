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

import numpy as np
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
    kernel straight into the peers' memory over NVLink and published with per-row epoch flags (r3_exchange_*, include/rend3_b200.h);
    `merge()` chains the consumer kernels on those flags and leaves the GLOBAL ascending visible list on this rank.
    torch.distributed only carries the 64-byte IPC handles once, at set-up."""

    def __init__(self, backend, camera: int, objects_per_rank: int, rank: int, world: int, rank_objects=None, rank_base=None):
        self.backend, self.camera, self.rank, self.world = backend, camera, rank, world
        self.rank_objects = np.asarray(rank_objects if rank_objects is not None else [objects_per_rank] * world, dtype=np.uint32)
        self.rank_base = np.asarray(rank_base if rank_base is not None else np.arange(world, dtype=np.uint64) * objects_per_rank, dtype=np.uint32)
        handle = backend.exchange_create(camera, world, rank, objects_per_rank)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        backend.exchange_connect(camera, b"".join(handles))
        _, _, self.words_per_rank = backend.exchange_words(camera)
        dist.barrier()   # every rank has mapped every buffer before anybody's cull writes into them

    def merge(self, rank_objects=None):
        """Consumer of the last cull (epoch): two kernels on the library's stream that wait for every rank's epoch flag on the device and
        expand the rows into the global visible list.  `rank_objects` overrides the shard sizes of this step (strong-scaled runs)."""
        self.backend.exchange_merge(self.camera, self.rank_objects if rank_objects is None else np.asarray(rank_objects, dtype=np.uint32), self.rank_base)

    def count(self, rank_objects=None):
        """The light consumer: waits for every rank's epoch flag on the device and counts the visible objects per shard (no list)."""
        self.backend.exchange_count(self.camera, self.rank_objects if rank_objects is None else np.asarray(rank_objects, dtype=np.uint32))

    def counts(self) -> np.ndarray:
        return self.backend.exchange_counts(self.camera, self.world)

    def join(self):
        """Make the library's main stream wait for the consumer kernels in flight (they run on its side stream, overlapping the next cull)."""
        self.backend.exchange_merged(self.camera)

    def gathered(self, device) -> torch.Tensor:
        """(world, words_per_rank) int32 view of the rows of the LAST epoch in the local buffer.  Complete after merge() was enqueued and
        the stream synchronised, or after the ranks synchronised their streams and passed a barrier."""
        ptr, nbytes, wpr = self.backend.exchange_words(self.camera)
        return torch.as_tensor(_View(ptr, (self.world, wpr), "<i4"), device=device)

    def merged(self, device) -> torch.Tensor:
        """The global visible list (int32 view of uint32 ids) left by the last merge(); synchronises the library stream."""
        lst, cnt, cap = self.backend.exchange_merged(self.camera)
        self.backend.sync()
        n = int(torch.as_tensor(_View(cnt, (1,), "<i4"), device=device).item())
        return torch.as_tensor(_View(lst, (max(n, 1),), "<i4"), device=device)[:n]

    def verify_against_nccl(self, stream, device) -> bool:
        """Outside timed regions: the merged list must equal the list expanded from an NCCL all-gather of the same visibility words."""
        wptr, wbytes = self.backend.device_ptr(self.camera, 4)
        with torch.cuda.stream(stream):
            mine = torch.as_tensor(_View(wptr, (wbytes // 4,), "<i4"), device=device)
            padded = torch.zeros(self.words_per_rank, dtype=torch.int32, device=device)
            padded[: mine.numel()] = mine
            ref = torch.empty(self.world * self.words_per_rank, dtype=torch.int32, device=device)
            dist.all_gather_into_tensor(ref, padded)
        torch.cuda.synchronize()
        words = ref.view(self.world, self.words_per_rank).cpu().numpy().view(np.uint32)
        want = []
        for r in range(self.world):
            bits = np.unpackbits(words[r].view(np.uint8), bitorder="little")[: int(self.rank_objects[r])]
            want.append(np.nonzero(bits)[0].astype(np.int64) + int(self.rank_base[r]))
        want = np.concatenate(want)
        got = self.merged(device).cpu().numpy().view(np.uint32).astype(np.int64)
        same = torch.tensor([1 if np.array_equal(got, want) else 0], device=device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        return bool(int(same.item()))

    def close(self):
        self.backend.sync()      # our queued compaction kernels have stored into the peers' buffers before anybody frees one
        dist.barrier()
        self.backend.exchange_destroy(self.camera)


class _View:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


ATLAS, ROWS, FRAME_DONE = 0, 1, 2


class ForwardSplit:
    """Multi-GPU forward pass: shadow maps split by light, screen split in row tiles, everything moved by peer-memory
    stores + epoch flags (r3_peer_*, rend3_b200/csrc/r3_peer.cu) on the library's own stream — no collective, no host barrier in a frame.
    The assembled rgba16f frame lands on `root` (default rank 0; root=-1: on every rank)."""

    def __init__(self, backend, stream, device, rank: int, world: int, resolution, n_shadows: int, root: int = 0, shard_triangle_cull: bool = True):
        self.b, self.stream, self.device, self.rank, self.world, self.res, self.n_shadows, self.root = backend, stream, device, rank, world, resolution, n_shadows, root
        self.shard_triangle_cull = shard_triangle_cull
        self.connected, self.frame, self.shadows = False, 0, None

    def owns_shadow(self, i: int) -> bool:
        return i % self.world == self.rank

    def assembles_on(self, rank: int) -> bool:
        return self.root < 0 or rank == self.root

    def connect(self):
        """After the atlas and the render target exist (add_to_graph's after_target hook): map the peers' buffers."""
        if self.connected:
            return
        handle = self.b.peer_create(self.world, self.rank)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        self.b.peer_connect(b"".join(handles))
        dist.barrier()
        if self.shard_triangle_cull:
            self.b.set_cull_shard(self.rank, self.world)     # the viewport's triangle cull, sharded by runs of batches
        self.connected = True

    def bind_scene(self, ev):
        self.shadows = [(s.offset[0], s.offset[1], s.size) for s in ev.shadows]

    def begin_frame(self):
        """after_target hook: map the peers (first frame), then wait — on the device — until every rank has finished the previous frame:
        from here on this rank stores into the peers' atlases, frames and word staging arrays."""
        self.connect()
        if self.frame >= 1:
            self.b.peer_wait(FRAME_DONE, [self.frame] * self.world)

    def send_shadow_maps(self):
        """after_shadows hook: the rects this rank rendered go to every peer (stores over NVLink while the viewport is culled and rasterised)."""
        b, f = self.b, self.frame + 1
        for i, (ox, oy, size) in enumerate(self.shadows or []):
            if self.owns_shadow(i):
                b.peer_send_atlas_rect(ox, oy, size, size)
        b.peer_signal(ATLAS)

    def wait_shadow_maps(self):
        """before_resolve hook: the shading samples the atlas — wait (on the device) until every rank's rects have arrived here."""
        self.b.peer_wait(ATLAS, [self.frame + 1] * self.world)

    def exchange_shadow_maps(self):
        self.send_shadow_maps()
        self.wait_shadow_maps()

    def exchange_rows(self, rows):
        """After the frame: this rank's rows of the rgba16f target go to the assembling rank(s); those wait for everybody's rows."""
        b, f = self.b, self.frame + 1
        b.peer_send_rows(rows[0], rows[1], self.root)
        b.peer_signal(ROWS)
        if self.assembles_on(self.rank):
            b.peer_wait(ROWS, [f] * self.world)
            b.tonemap(True)
        b.peer_signal(FRAME_DONE)      # this rank no longer reads its atlas / (assembling ranks) its frame: the next frame's stores may land
        self.frame = f

    def describe(self) -> str:
        where = "every rank" if self.root < 0 else f"rank {self.root}"
        cull = "the viewport's triangle cull sharded by runs of batches (visibility words exchanged by peer stores); " if self.shard_triangle_cull else ""
        return (f"{self.world} row tiles; " + cull + "shadow maps split by light, each rect stored by its owner straight into every peer's atlas; rgba16f rows stored into the frame of "
                f"{where}; NVLink peer stores + st.release.sys / ld.acquire.sys epoch flags, no collective kernel, no host barrier inside a frame")

    def close(self):
        self.b.sync()
        dist.barrier()
        self.b.peer_destroy()
