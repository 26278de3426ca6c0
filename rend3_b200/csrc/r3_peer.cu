// r3_peer.cu — peer-memory plumbing of the multi-GPU forward pass (shadow maps split by light, screen split in row tiles).
//
// One process per GPU on one NVLink / NVSwitch node.  Every rank maps the shadow atlas, the rgba16f colour target and a small flag
// block of every other rank (CUDA IPC).  Producers write STRAIGHT into the consumers' memory with plain coalesced stores from a
// copy kernel — the rect of a shadow map its owner rendered into every peer's atlas, the rows a rank shaded into the assembling
// rank's frame — and publish them with an epoch flag per (kind, sender) written by st.release.sys; consumers wait for the flags
// with ld.acquire.sys in a one-CTA kernel on their own stream.  No collective kernel, no reduction (the r1 design merged the atlas
// with a 67 MB integer MAX all-reduce although the ranks own disjoint lights), no host barrier inside a frame.
//
// Protocol of one frame f on every rank (rend3_b200/parallel.py::ForwardSplit drives it):
//   clear the rects of the lights this rank owns; cull + shadow passes of those lights;
//   wait(FRAME_DONE, f - 1 from everybody)      nobody still samples last frame's atlas;
//   send_atlas_rect(own rects) ; signal(ATLAS)   ;  wait(ATLAS, f from everybody)
//   viewport cull + raster + resolve of the own row tile; send_rows ; signal(ROWS) ; signal(FRAME_DONE)
//   assembling rank(s): wait(ROWS, f from everybody), then tonemap.
// Every rank signals every kind every frame (also with nothing to send): the flags double as flow control, so a single buffer per
// target is enough — a producer can only be one frame ahead of the slowest consumer.
#include "r3_common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct PeerPtrs { void* p[R3_MAX_EXCHANGE_RANKS]; };

// rows x row_bytes block at (byte offset `first`, pitch) of `src` into the same place of every destination; 16-byte lanes
__global__ void __launch_bounds__(256) peer_copy_kernel(const uint8_t* __restrict__ src, const __grid_constant__ PeerPtrs dst, uint32_t n_dst, uint64_t first, uint64_t pitch,
                                                        uint32_t row_vecs, uint32_t rows) {
    const uint64_t total = (uint64_t)row_vecs * rows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_vecs, v = i - r * row_vecs, off = first + r * pitch + v * 16u;
        const uint4 x = *reinterpret_cast<const uint4*>(src + off);
#pragma unroll 1
        for (uint32_t d = 0; d < n_dst; ++d) *reinterpret_cast<uint4*>(static_cast<uint8_t*>(dst.p[d]) + off) = x;
    }
}
// the kernels in front of this one on the stream have completed, so their peer stores are performed; one release per destination
__global__ void peer_signal_kernel(const __grid_constant__ PeerPtrs flags, uint32_t n_dst, uint32_t slot, uint32_t epoch) {
    if (threadIdx.x < n_dst) {
        __threadfence_system();
        st_release_sys(static_cast<uint32_t*>(flags.p[threadIdx.x]) + slot, epoch);
    }
}
struct WaitParams { const uint32_t* flags; uint32_t slot0; uint32_t n; uint32_t expected[R3_MAX_EXCHANGE_RANKS]; };
__global__ void peer_wait_kernel(const __grid_constant__ WaitParams w) {
    if (threadIdx.x < w.n) {
        const uint32_t* f = w.flags + w.slot0 + threadIdx.x;
        while ((int32_t)(ld_acquire_sys(f) - w.expected[threadIdx.x]) < 0) __nanosleep(64);
    }
}

}  // namespace

// the context's least-priority side stream (shared with the exchange consumer), created on first use
static int peer_side_stream(r3_ctx* c) {
    if (c->side_stream) return R3_OK;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    R3_CUDA(c, cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, lo));
    return R3_OK;
}
static int peer_copy(r3_ctx* c, cudaStream_t stream, const void* src, void* const* dsts, uint32_t n_dst, uint64_t first, uint64_t pitch, uint64_t row_bytes, uint32_t rows) {
    if (!n_dst || !rows || !row_bytes) return R3_OK;
    if ((first | pitch | row_bytes) & 15u) return r3_fail(c, R3_E_INVALID, "peer copy: rect not 16-byte aligned");
    PeerPtrs d{};
    for (uint32_t k = 0; k < n_dst; ++k) d.p[k] = dsts[k];
    const uint64_t vecs = row_bytes / 16 * rows;
    const uint32_t grid = (uint32_t)((vecs + 255) / 256 < (uint64_t)R3_SM_COUNT * 8 ? (vecs + 255) / 256 : (uint64_t)R3_SM_COUNT * 8);
    peer_copy_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(src), d, n_dst, first, pitch, (uint32_t)(row_bytes / 16), rows);
    R3_CHECK_LAUNCH(c, "peer_copy_kernel");
    return R3_OK;
}

R3_EXPORT int r3_peer_create(r3_ctx* c, uint32_t n_ranks, uint32_t my_rank, uint8_t handles_out[4 * R3_IPC_HANDLE_BYTES]) {
    if (!c || !handles_out) return r3_fail(c, R3_E_INVALID, "peer_create: null");
    if (n_ranks == 0 || n_ranks > R3_MAX_EXCHANGE_RANKS || my_rank >= n_ranks) return r3_fail(c, R3_E_INVALID, "peer_create: bad rank layout");
    if (!c->d_hdr16) return r3_fail(c, R3_E_STATE, "peer_create before set_render_target");
    if (c->peer.created) return r3_fail(c, R3_E_STATE, "peer_create: already created");
    cudaSetDevice(c->device);
    R3_CUDA(c, cudaMalloc((void**)&c->peer.d_flags, 1024));
    R3_CUDA(c, cudaMemsetAsync(c->peer.d_flags, 0, 1024, c->stream));
    R3_CUDA(c, r3_stream_sync(c));
    memset(handles_out, 0, 4 * R3_IPC_HANDLE_BYTES);
    cudaIpcMemHandle_t h;
    R3_CUDA(c, cudaIpcGetMemHandle(&h, c->peer.d_flags));
    memcpy(handles_out, &h, sizeof h);
    if (c->d_atlas) { R3_CUDA(c, cudaIpcGetMemHandle(&h, c->d_atlas)); memcpy(handles_out + R3_IPC_HANDLE_BYTES, &h, sizeof h); }
    R3_CUDA(c, cudaIpcGetMemHandle(&h, c->d_hdr16));
    memcpy(handles_out + 2 * R3_IPC_HANDLE_BYTES, &h, sizeof h);
    // staging arrays of the sharded triangle test: [pred words | resid words], sized by the upper bound of the world's padded invocations
    R3_TRY(r3_compute_max_invocations(c));
    c->peer.tri_cap_words = (c->max_total_invocations + 31) / 32 + 64;
    R3_CUDA(c, cudaMalloc((void**)&c->peer.d_tri_words, c->peer.tri_cap_words * 2 * 4));
    R3_CUDA(c, cudaMemsetAsync(c->peer.d_tri_words, 0, c->peer.tri_cap_words * 2 * 4, c->stream));
    R3_CUDA(c, r3_stream_sync(c));
    R3_CUDA(c, cudaIpcGetMemHandle(&h, c->peer.d_tri_words));
    memcpy(handles_out + 3 * R3_IPC_HANDLE_BYTES, &h, sizeof h);
    c->peer.created = true; c->peer.connected = false; c->peer.n_ranks = n_ranks; c->peer.rank = my_rank;
    c->peer.has_atlas = c->d_atlas != nullptr;
    c->peer.atlas_at_create = c->d_atlas; c->peer.hdr_at_create = c->d_hdr16;
    for (auto& s : c->peer.sent) s = 0;
    return R3_OK;
}
R3_EXPORT int r3_peer_connect(r3_ctx* c, const uint8_t* handles) {
    if (!c || !handles) return r3_fail(c, R3_E_INVALID, "peer_connect: null");
    if (!c->peer.created) return r3_fail(c, R3_E_STATE, "peer_connect before peer_create");
    cudaSetDevice(c->device);
    for (uint32_t r = 0; r < c->peer.n_ranks; ++r) {
        if (r == c->peer.rank) { c->peer.flags[r] = c->peer.d_flags; c->peer.atlas[r] = c->d_atlas; c->peer.hdr16[r] = c->d_hdr16; c->peer.tri_words[r] = c->peer.d_tri_words; continue; }
        const uint8_t* base = handles + (size_t)r * 4 * R3_IPC_HANDLE_BYTES;
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, base, sizeof h);
        R3_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer.flags[r] = (uint32_t*)p;
        if (c->peer.has_atlas) {
            memcpy(&h, base + R3_IPC_HANDLE_BYTES, sizeof h);
            R3_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->peer.atlas[r] = (float*)p;
        }
        memcpy(&h, base + 2 * R3_IPC_HANDLE_BYTES, sizeof h);
        R3_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer.hdr16[r] = (uint16_t*)p;
        memcpy(&h, base + 3 * R3_IPC_HANDLE_BYTES, sizeof h);
        R3_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer.tri_words[r] = (uint32_t*)p;
    }
    c->peer.connected = true;
    return R3_OK;
}
static int peer_ready(r3_ctx* c, const char* who) {
    if (!c) return R3_E_INVALID;
    if (!c->peer.connected) return r3_fail(c, R3_E_STATE, who);
    if (c->d_hdr16 != c->peer.hdr_at_create || c->d_atlas != c->peer.atlas_at_create) return r3_fail(c, R3_E_STATE, "peer buffers were reallocated after r3_peer_create: destroy and create again");
    cudaSetDevice(c->device);
    return R3_OK;
}
R3_EXPORT int r3_peer_send_atlas_rect(r3_ctx* c, uint32_t ox, uint32_t oy, uint32_t w, uint32_t h) {
    R3_TRY(peer_ready(c, "peer_send_atlas_rect before peer_connect"));
    if (!c->peer.has_atlas) return r3_fail(c, R3_E_STATE, "peer_send_atlas_rect: no shadow atlas");
    if ((uint64_t)ox + w > c->atlas_w || (uint64_t)oy + h > c->atlas_h) return r3_fail(c, R3_E_INVALID, "peer_send_atlas_rect: rect outside the atlas");
    void* dst[R3_MAX_EXCHANGE_RANKS];
    uint32_t n = 0;
    for (uint32_t r = 0; r < c->peer.n_ranks; ++r) if (r != c->peer.rank) dst[n++] = c->peer.atlas[r];
    // The copy (16.7 MB per 2048^2 map and peer) runs on the side stream behind the shadow passes recorded so far, so that this rank's
    // viewport cull and raster do not queue behind it; r3_peer_signal(ATLAS) follows it there.  Nobody writes the rect again before every
    // peer has signalled FRAME_DONE, which it does after it has seen the ATLAS flag, i.e. after the copy has completed.
    R3_TRY(peer_side_stream(c));
    if (!c->peer.side_event) R3_CUDA(c, cudaEventCreateWithFlags(&c->peer.side_event, cudaEventDisableTiming));
    R3_CUDA(c, cudaEventRecord(c->peer.side_event, c->stream));
    R3_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->peer.side_event, 0));
    c->peer.atlas_on_side = true;
    return peer_copy(c, c->side_stream, c->d_atlas, dst, n, ((uint64_t)oy * c->atlas_w + ox) * 4, (uint64_t)c->atlas_w * 4, (uint64_t)w * 4, h);
}
R3_EXPORT int r3_peer_send_rows(r3_ctx* c, uint32_t row_begin, uint32_t row_end, int root) {
    R3_TRY(peer_ready(c, "peer_send_rows before peer_connect"));
    if (row_begin > row_end || row_end > c->height || root >= (int)c->peer.n_ranks) return r3_fail(c, R3_E_INVALID, "peer_send_rows: bad arguments");
    void* dst[R3_MAX_EXCHANGE_RANKS];
    uint32_t n = 0;
    for (uint32_t r = 0; r < c->peer.n_ranks; ++r) if (r != c->peer.rank && (root < 0 || (int)r == root)) dst[n++] = c->peer.hdr16[r];
    const uint64_t pitch = (uint64_t)c->width * 8;
    return peer_copy(c, c->stream, c->d_hdr16, dst, n, (uint64_t)row_begin * pitch, pitch, pitch, row_end - row_begin);
}
R3_EXPORT int r3_peer_signal(r3_ctx* c, uint32_t kind) {
    R3_TRY(peer_ready(c, "peer_signal before peer_connect"));
    if (kind >= R3_PEER_KINDS) return r3_fail(c, R3_E_INVALID, "peer_signal: kind");
    PeerPtrs f{};
    for (uint32_t r = 0; r < c->peer.n_ranks; ++r) f.p[r] = c->peer.flags[r];
    const uint32_t epoch = ++c->peer.sent[kind];
    cudaStream_t stream = c->stream;
    if (kind == 0u && c->peer.atlas_on_side) { stream = c->side_stream; c->peer.atlas_on_side = false; }   // behind the atlas copies of this frame
    peer_signal_kernel<<<1, 32, 0, stream>>>(f, c->peer.n_ranks, kind * R3_MAX_EXCHANGE_RANKS + c->peer.rank, epoch);
    R3_CHECK_LAUNCH(c, "peer_signal_kernel");
    return R3_OK;
}
R3_EXPORT int r3_peer_wait(r3_ctx* c, uint32_t kind, const uint32_t* expected) {
    R3_TRY(peer_ready(c, "peer_wait before peer_connect"));
    if (kind >= R3_PEER_KINDS || !expected) return r3_fail(c, R3_E_INVALID, "peer_wait: bad arguments");
    WaitParams w{};
    w.flags = c->peer.d_flags; w.slot0 = kind * R3_MAX_EXCHANGE_RANKS; w.n = c->peer.n_ranks;
    for (uint32_t r = 0; r < c->peer.n_ranks; ++r) w.expected[r] = expected[r];
    peer_wait_kernel<<<1, 32, 0, c->stream>>>(w);
    R3_CHECK_LAUNCH(c, "peer_wait_kernel");
    return R3_OK;
}
R3_EXPORT int r3_peer_destroy(r3_ctx* c) {
    if (!c) return R3_E_INVALID;
    if (!c->peer.created) return R3_OK;
    cudaSetDevice(c->device);
    r3_stream_sync(c);
    if (c->peer.connected)
        for (uint32_t r = 0; r < c->peer.n_ranks; ++r) {
            if (r == c->peer.rank) continue;
            if (c->peer.flags[r]) cudaIpcCloseMemHandle(c->peer.flags[r]);
            if (c->peer.atlas[r]) cudaIpcCloseMemHandle(c->peer.atlas[r]);
            if (c->peer.hdr16[r]) cudaIpcCloseMemHandle(c->peer.hdr16[r]);
            if (c->peer.tri_words[r]) cudaIpcCloseMemHandle(c->peer.tri_words[r]);
        }
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);
    if (c->peer.side_event) cudaEventDestroy(c->peer.side_event);
    cudaFree(c->peer.d_flags); cudaFree(c->peer.d_tri_words);
    c->peer = r3_peer_state{};
    c->tri_shard_index = 0; c->tri_shard_count = 1;
    return R3_OK;
}
// SURVEY 8e "triangle cull: shard by batch": from now on r3_cull of the VIEWPORT camera tests only the shard's run of workgroups (= batches)
// and exchanges the visibility words with the peers (r3_tri_cull.cu); count <= 1 switches it off.  Every rank must use the same count.
R3_EXPORT int r3_set_cull_shard(r3_ctx* c, uint32_t shard_index, uint32_t shard_count) {
    if (!c) return R3_E_INVALID;
    if (shard_count > 1u) {
        if (!c->peer.connected) return r3_fail(c, R3_E_STATE, "set_cull_shard before peer_connect");
        if (shard_count != c->peer.n_ranks || shard_index != c->peer.rank) return r3_fail(c, R3_E_INVALID, "set_cull_shard: the shards are the peer ranks");
    }
    c->tri_shard_index = shard_count > 1u ? shard_index : 0u; c->tri_shard_count = shard_count > 1u ? shard_count : 1u;
    return R3_OK;
}
// the rect of one shadow map (a rank that owns only some of the lights clears only their rects: the others arrive from their owners)
R3_EXPORT int r3_clear_shadow_rect(r3_ctx* c, uint32_t ox, uint32_t oy, uint32_t w, uint32_t h) {
    if (!c || !c->d_atlas) return r3_fail(c, R3_E_STATE, "clear_shadow_rect before set_directional_lights");
    if ((uint64_t)ox + w > c->atlas_w || (uint64_t)oy + h > c->atlas_h) return r3_fail(c, R3_E_INVALID, "clear_shadow_rect: rect outside the atlas");
    cudaSetDevice(c->device);
    if (w && h) R3_CUDA(c, cudaMemset2DAsync(c->d_atlas + (size_t)oy * c->atlas_w + ox, (size_t)c->atlas_w * 4, 0, (size_t)w * 4, h, c->stream));
    return R3_OK;
}
