// r3_cull_bake.cu — fused per-object frustum cull + object-uniform bake, then ordered visible-list compaction.
//
// Replaces (reference paths):
//   * uniform_prep.wgsl::cs_main          rend3-routine/shaders/src/uniform_prep.wgsl:9-27
//       MV = view * T, MVP = view_proj * T for every enabled slot < object_count
//   * the sphere/frustum filter of batch_objects   rend3-routine/src/culling/batching.rs:144-148
//       Frustum::contains_sphere, 5 planes        rend3/src/util/frustum.rs:148-161
// and emits the visible slots as one ASCENDING u32 list (the canonical, bit-exact artefact).
//
// Design (HBM-bound: 224 flop against >= 212 bytes per object).
//   hot/cold split — the reference's 128-byte Object record (object.rs:23-36) stays the canonical store for the kernels
//     that need its cold fields (first_index, index_count, material, attribute offsets); the 84 bytes this path reads
//     (transform, bounding sphere, enabled) are ALSO kept as dense arrays — transforms[4n] float4, spheres[n] float4,
//     enabled 1 bit per slot — filled by split_objects_kernel whenever records are uploaded (r3_set_objects,
//     r3_set_objects_device) or scattered (r3_update_objects).  With the AoS records every 32-byte sector of a record is
//     touched, i.e. 128 B read per object; the dense arrays bring that to 80 B + 1 bit (and to 16 B for cull-only).
//   stream kernel  — no inter-CTA dependency, no shared memory, no shuffles:
//     * a warp owns 32 consecutive slots.  Four fully coalesced 512-byte loads fetch their 128 transform columns; the
//       lane holding column j of slot o multiplies it by `view` and by `view_proj` (operands straight from the constant
//       bank) and stores column j of MV and of MVP — 64-byte runs, whole sectors.  Arithmetic is __fmul_rn/__fadd_rn in
//       WGSL's accumulation order, never contracted: MV/MVP are bit-identical to the CPU oracle;
//     * lane l loads the sphere of slot base+l (one coalesced 512-byte load) and tests it; the ballot IS the 32-bit
//       visibility word of the 32 slots (1 bit per object goes to HBM);
//     * each CTA (1024 slots) also leaves its survivor count.
//   compact kernel — one CTA per 32768 objects: sums the CTA counts in front of it (<= 40 KB, L2 resident),
//     scans its 1024 visibility words and writes the surviving slot ids in ascending order.  It moves
//     N/8 + 4*visible bytes, ~1% of the stream kernel's traffic.
// (History, profiles/README.md: a single-pass kernel with a decoupled look-back lost 20% to look-back stalls; the
//  two-kernel AoS version ran at the copy bandwidth but moved 256 B per object.)
#include <cstdlib>

#include "r3_common.cuh"

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_WARPS = CB_THREADS / 32;
constexpr int CB_WT = 4;                                   // 32-object warp tiles per warp
constexpr int CB_CTA_OBJECTS = CB_WARPS * CB_WT * 32;      // 1024
constexpr int CP_THREADS = 1024;                           // compact kernel: one visibility word per thread
constexpr int CP_OBJECTS = CP_THREADS * 32;                // 32768 objects per compact CTA
constexpr int CP_CTAS_PER_TILE = CP_OBJECTS / CB_CTA_OBJECTS;   // 32 stream-CTA counts per compact CTA

struct CullBakeParams {
    float view[16];
    float view_proj[16];
    float frustum[5][4];
    uint32_t object_count;
};

// AoS Object records -> the dense hot arrays.  8 lanes per 128-byte record (coalesced 512-byte loads); float4 #0-3 =
// transform, #4 = bounding sphere, #7.y = `enabled` (byte 116).
__global__ void __launch_bounds__(256) split_objects_kernel(const float4* __restrict__ objects, uint32_t n, float4* __restrict__ transforms,
                                                            float4* __restrict__ spheres, uint32_t* __restrict__ enabled_bits) {
    const int lane = threadIdx.x & 31, k = lane & 7, g = lane >> 3;
    const uint32_t wtile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, base = wtile * 32u;
    if (base >= n) return;
    uint32_t bits = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const uint32_t obj = base + it * 4 + g;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (obj < n) {
            r = __ldcs(&objects[(size_t)obj * 8 + k]);
            if (k < 4) transforms[(size_t)obj * 4 + k] = r;
            else if (k == 4) spheres[obj] = r;
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, k == 7 && obj < n && __float_as_uint(r.y) != 0u);   // lanes 7, 15, 23, 31
        bits |= (((b >> 7) & 1u) | (((b >> 15) & 1u) << 1) | (((b >> 23) & 1u) << 2) | (((b >> 31) & 1u) << 3)) << (it * 4);
    }
    if (lane == 0) enabled_bits[wtile] = bits;
}
// the same for the records r3_update_objects has just scattered (ScatterCopy, util/scatter_copy.rs:69-136)
__global__ void __launch_bounds__(256) split_slots_kernel(const float4* __restrict__ objects, const uint32_t* __restrict__ slots, uint32_t n_updates, uint32_t n_slots,
                                                          float4* __restrict__ transforms, float4* __restrict__ spheres, uint32_t* __restrict__ enabled_bits) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, k = threadIdx.x & 7;
    if (i >= n_updates) return;
    const uint32_t s = slots[i];
    if (s >= n_slots) return;
    const float4 r = objects[(size_t)s * 8 + k];
    if (k < 4) transforms[(size_t)s * 4 + k] = r;
    else if (k == 4) spheres[s] = r;
    else if (k == 7) {
        if (__float_as_uint(r.y) != 0u) atomicOr(&enabled_bits[s >> 5], 1u << (s & 31u));
        else atomicAnd(&enabled_bits[s >> 5], ~(1u << (s & 31u)));
    }
}

template <bool BAKE, bool CULL, bool LIVE>
__global__ void __launch_bounds__(CB_THREADS)
cull_bake_kernel(const float4* __restrict__ transforms, const float4* __restrict__ spheres, const uint32_t* __restrict__ enabled_bits,
                 const uint32_t* __restrict__ live_bits, float4* __restrict__ matrices, uint32_t* __restrict__ words, uint32_t* __restrict__ cta_counts,
                 const __grid_constant__ CullBakeParams p) {
    __shared__ uint32_t s_count[CB_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane & 3, sub = lane >> 2;   // transform column / slot within an 8-slot group
    uint32_t count = 0;
#pragma unroll 2
    for (int wt = 0; wt < CB_WT; ++wt) {
        const uint32_t wtile = (blockIdx.x * CB_WARPS + warp) * CB_WT + wt;   // visibility word index
        const uint32_t base = wtile * 32u;
        if (base >= p.object_count) break;
        float4 t[4];
        float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t enabled = 0u;
        if (BAKE) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t obj = base + it * 8 + sub;
                t[it] = obj < p.object_count ? __ldcs(&transforms[(size_t)base * 4 + it * 32 + lane]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (CULL && base + lane < p.object_count) sp = __ldcs(&spheres[base + lane]);
        if (BAKE || (CULL && !LIVE)) enabled = __ldg(&enabled_bits[wtile]);
        if (BAKE) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t slot = it * 8 + sub, obj = base + slot;
                if (obj < p.object_count && ((enabled >> slot) & 1u)) {   // uniform_prep.wgsl:18-20 skips disabled slots
                    float4* dst = &matrices[(size_t)obj * 8 + col];
                    __stcs(dst, mat_vec_rn(p.view, t[it].x, t[it].y, t[it].z, t[it].w));
                    __stcs(dst + 4, mat_vec_rn(p.view_proj, t[it].x, t[it].y, t[it].z, t[it].w));
                }
            }
        }
        if (CULL) {
            // one object per lane: Plane::distance = abc.dot(center) + d with glam's scalar dot order (util/frustum.rs:79-81,148-161)
            const uint32_t live = LIVE ? __ldg(&live_bits[wtile]) : enabled;
            const float neg_radius = -sp.w;
            bool inside = true;
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) {
                const float d = add_rn(add_rn(add_rn(mul_rn(p.frustum[pl][0], sp.x), mul_rn(p.frustum[pl][1], sp.y)), mul_rn(p.frustum[pl][2], sp.z)), p.frustum[pl][3]);
                inside = inside && (d >= neg_radius);
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, base + lane < p.object_count && ((live >> lane) & 1u) && inside);
            if (lane == 0) words[wtile] = word;
            count += __popc(word);
        }
    }
    if (CULL) {
        if (lane == 0) s_count[warp] = count;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < CB_WARPS; ++w) t += s_count[w];
            cta_counts[blockIdx.x] = t;
        }
    }
}

// visible-set exchange fused into the compaction: every word this rank produced is also stored, coalesced, into the gathered
// buffer of every rank (its own included) through NVLink peer mappings — no collective kernel, no extra pass over the words
struct ExchangeParams { uint32_t* peers[R3_MAX_EXCHANGE_RANKS]; uint32_t n_ranks, word_offset, words_per_rank; };

__global__ void __launch_bounds__(CP_THREADS)
compact_visible_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ cta_counts, uint32_t n_words, uint32_t n_cta_counts,
                       uint32_t* __restrict__ visible, uint32_t* __restrict__ visible_count, const __grid_constant__ ExchangeParams ex) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (1) survivors in front of this tile: sum of the stream CTAs' counts before it
    uint32_t before = 0;
    const uint32_t first_cta = blockIdx.x * CP_CTAS_PER_TILE;
    for (uint32_t i = threadIdx.x; i < first_cta && i < n_cta_counts; i += CP_THREADS) before += __ldg(&cta_counts[i]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, s);
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, s);
        if (lane == 0) s_base = v;
    }
    __syncthreads();
    const uint32_t tile_base = s_base;
    __syncthreads();
    // (2) block-wide exclusive scan of the word popcounts (one word per thread)
    const uint32_t wi = blockIdx.x * CP_THREADS + threadIdx.x;
    const uint32_t word = wi < n_words ? __ldg(&words[wi]) : 0u;
    if (ex.n_ranks && wi < ex.words_per_rank) {
#pragma unroll 1
        for (uint32_t r = 0; r < ex.n_ranks; ++r) ex.peers[r][ex.word_offset + wi] = word;
    }
    const uint32_t c = __popc(word);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi2 = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi2, d);
            if (lane >= d) wi2 += n;
        }
        s_warp[lane] = wi2 - w;
        if (lane == 31 && blockIdx.x == gridDim.x - 1) *visible_count = tile_base + wi2;
    }
    __syncthreads();
    const uint32_t excl = tile_base + s_warp[warp] + incl - c;
    // (3) ascending slot ids: the 32 words of a warp are expanded one after the other, one bit per lane
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
        const uint32_t wj = __shfl_sync(0xFFFFFFFFu, word, j), ej = __shfl_sync(0xFFFFFFFFu, excl, j);
        if ((wj >> lane) & 1u) visible[ej + __popc(wj & ((1u << lane) - 1u))] = (blockIdx.x * CP_THREADS + warp * 32 + j) * 32u + lane;
    }
}

}  // namespace

int r3_launch_cull_bake(r3_ctx* c, r3_camera* cam, uint32_t mode) {
    const uint32_t n = cam->header.object_count;
    const bool bake = mode & R3_CB_BAKE, cull = mode & R3_CB_CULL;
    if (cull) R3_CUDA(c, cudaMemsetAsync(cam->d_visible_count, 0, 4, c->stream));
    if (n == 0 || (!bake && !cull)) return R3_OK;
    CullBakeParams p;
    memcpy(p.view, cam->header.view, 64);
    memcpy(p.view_proj, cam->header.view_proj, 64);
    memcpy(p.frustum, cam->header.frustum, 80);
    p.object_count = n;
    const uint32_t n_ctas = (n + CB_CTA_OBJECTS - 1) / CB_CTA_OBJECTS, n_words = (n + 31) / 32;
    // scratch: two sets of { visibility words [n_words] | stream-CTA counts [n_ctas] }, used alternately: the words of
    // call k stay intact while call k+1 runs, so an exchange of the visible set (r3_device_ptr which = 4) can overlap the
    // next cull on another stream.  (d_tile_state is a u64 array; each set is padded to a 256-byte multiple.)
    const uint64_t set_words = (((uint64_t)n_words + n_ctas + 63) / 64) * 64;
    R3_TRY(r3_reserve_t(c, &cam->d_tile_state, &cam->tile_state_cap, set_words + 2));
    if (cull) cam->words_set ^= 1u;
    uint32_t* words = reinterpret_cast<uint32_t*>(cam->d_tile_state) + (size_t)cam->words_set * set_words;
    cam->d_words = words;
    uint32_t* cta_counts = words + n_words;
    if (!c->hot_valid) return r3_fail(c, R3_E_STATE, "object_uniform_upload before set_objects");
    float4* mats = reinterpret_cast<float4*>(cam->d_matrices);
    const bool live = c->have_live && cull;
#define R3_CB_LAUNCH(B, C, L) \
    cull_bake_kernel<B, C, L><<<n_ctas, CB_THREADS, 0, c->stream>>>(c->d_hot_transform, c->d_hot_sphere, c->d_enabled_bits, c->d_live_bits, mats, words, cta_counts, p)
    if (bake && cull) { if (live) R3_CB_LAUNCH(true, true, true); else R3_CB_LAUNCH(true, true, false); }
    else if (bake) R3_CB_LAUNCH(true, false, false);
    else { if (live) R3_CB_LAUNCH(false, true, true); else R3_CB_LAUNCH(false, true, false); }
#undef R3_CB_LAUNCH
    R3_CHECK_LAUNCH(c, "cull_bake_kernel");
    if (cull) {
        const uint32_t n_tiles = (n_words + CP_THREADS - 1) / CP_THREADS;
        ExchangeParams ex{};
        if (cam->ex_connected) {
            if (n_words > cam->ex_words_per_rank) return r3_fail(c, R3_E_INVALID, "object_uniform_upload: more objects than the exchange was created for");
            for (uint32_t r = 0; r < cam->ex_ranks; ++r) ex.peers[r] = cam->ex_peers[r];
            ex.n_ranks = cam->ex_ranks; ex.word_offset = cam->ex_rank * cam->ex_words_per_rank; ex.words_per_rank = cam->ex_words_per_rank;
        }
        // with an exchange every slot of this rank's row is written each step (the tail beyond n_words as zeros)
        const uint32_t n_tiles_ex = cam->ex_connected ? (cam->ex_words_per_rank + CP_THREADS - 1) / CP_THREADS : 0u;
        compact_visible_kernel<<<n_tiles > n_tiles_ex ? n_tiles : n_tiles_ex, CP_THREADS, 0, c->stream>>>(words, cta_counts, n_words, n_ctas, cam->d_visible, cam->d_visible_count, ex);
        R3_CHECK_LAUNCH(c, "compact_visible_kernel");
    }
    return R3_OK;
}

// (re)build the dense hot arrays from the AoS records (all slots)
int r3_split_objects(r3_ctx* c) {
    const uint32_t n = c->n_slots;
    const uint64_t want = n ? n : 1;
    if (want > c->hot_cap) {
        cudaFree(c->d_hot_transform); cudaFree(c->d_hot_sphere); cudaFree(c->d_enabled_bits);
        c->d_hot_transform = nullptr; c->d_hot_sphere = nullptr; c->d_enabled_bits = nullptr; c->hot_cap = 0;
        uint64_t cap = 1024;
        while (cap < want) cap *= 2;
        if (cap > want + want / 8 && want > (1u << 20)) cap = want + want / 8;   // large worlds: 12.5% head room instead of a power of two
        R3_CUDA(c, cudaMalloc((void**)&c->d_hot_transform, cap * 64));
        R3_CUDA(c, cudaMalloc((void**)&c->d_hot_sphere, cap * 16));
        R3_CUDA(c, cudaMalloc((void**)&c->d_enabled_bits, ((cap + 31) / 32 + 1) * 4));
        c->hot_cap = cap;
    }
    if (n) {
        const uint32_t warps = (n + 31) / 32;
        split_objects_kernel<<<(warps + 7) / 8, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->d_objects), n, c->d_hot_transform, c->d_hot_sphere, c->d_enabled_bits);
        R3_CHECK_LAUNCH(c, "split_objects_kernel");
    }
    c->hot_valid = true;
    return R3_OK;
}
// refresh the hot copies of the `n` slots listed in d_slots (device pointer) after a scatter
int r3_split_slots(r3_ctx* c, const uint32_t* d_slots, uint32_t n) {
    if (!c->hot_valid || n == 0) return R3_OK;
    split_slots_kernel<<<(n * 8 + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->d_objects), d_slots, n, c->n_slots, c->d_hot_transform, c->d_hot_sphere,
                                                                    c->d_enabled_bits);
    R3_CHECK_LAUNCH(c, "split_slots_kernel");
    return R3_OK;
}

// ------------------------------------------------------------------ multi-GPU exchange of the visible set
// One process per GPU.  Every rank owns gathered[n_ranks][words_per_rank]; rank r's compact kernel stores row r into the
// buffer of every rank through CUDA IPC peer mappings over NVLink / NVSwitch.  The rows are complete once the ranks have
// synchronised their streams and met at a barrier (the caller's: torch.distributed / MPI / the frame fence).
R3_EXPORT int r3_exchange_create(r3_ctx* c, uint32_t camera, uint32_t n_ranks, uint32_t my_rank, uint32_t max_objects_per_rank, uint8_t handle_out[R3_IPC_HANDLE_BYTES]) {
    if (!c || !handle_out) return r3_fail(c, R3_E_INVALID, "exchange_create: null");
    if (n_ranks == 0 || n_ranks > R3_MAX_EXCHANGE_RANKS || my_rank >= n_ranks || max_objects_per_rank == 0) return r3_fail(c, R3_E_INVALID, "exchange_create: bad rank layout");
    r3_camera* cam = r3_get_camera(c, camera);
    if (!cam) return r3_fail(c, R3_E_INVALID, "exchange_create: bad camera");
    static_assert(sizeof(cudaIpcMemHandle_t) == R3_IPC_HANDLE_BYTES, "IPC handle size");
    cudaSetDevice(c->device);
    if (cam->d_gathered) return r3_fail(c, R3_E_STATE, "exchange_create: already created for this camera");
    const uint32_t wpr = (((max_objects_per_rank + 31u) / 32u) + 63u) & ~63u;   // rows start 256-byte aligned
    R3_CUDA(c, cudaMalloc((void**)&cam->d_gathered, (size_t)n_ranks * wpr * 4));
    R3_CUDA(c, cudaMemsetAsync(cam->d_gathered, 0, (size_t)n_ranks * wpr * 4, c->stream));
    R3_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaIpcMemHandle_t h;
    R3_CUDA(c, cudaIpcGetMemHandle(&h, cam->d_gathered));
    memcpy(handle_out, &h, sizeof h);
    cam->ex_ranks = n_ranks; cam->ex_rank = my_rank; cam->ex_words_per_rank = wpr; cam->ex_connected = false;
    return R3_OK;
}
R3_EXPORT int r3_exchange_connect(r3_ctx* c, uint32_t camera, const uint8_t* handles) {
    if (!c || !handles) return r3_fail(c, R3_E_INVALID, "exchange_connect: null");
    r3_camera* cam = r3_get_camera(c, camera);
    if (!cam || !cam->d_gathered) return r3_fail(c, R3_E_STATE, "exchange_connect before exchange_create");
    cudaSetDevice(c->device);
    for (uint32_t r = 0; r < cam->ex_ranks; ++r) {
        if (r == cam->ex_rank) { cam->ex_peers[r] = cam->d_gathered; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * R3_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        R3_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // maps the peer allocation and enables NVLink peer access
        cam->ex_peers[r] = (uint32_t*)p;
    }
    cam->ex_connected = true;
    return R3_OK;
}
R3_EXPORT int r3_exchange_words(r3_ctx* c, uint32_t camera, void** device_ptr, uint64_t* nbytes, uint32_t* words_per_rank) {
    if (!c || !device_ptr || !nbytes) return r3_fail(c, R3_E_INVALID, "exchange_words: null");
    r3_camera* cam = r3_get_camera(c, camera);
    if (!cam || !cam->d_gathered) return r3_fail(c, R3_E_STATE, "exchange_words before exchange_create");
    *device_ptr = cam->d_gathered; *nbytes = (uint64_t)cam->ex_ranks * cam->ex_words_per_rank * 4;
    if (words_per_rank) *words_per_rank = cam->ex_words_per_rank;
    return R3_OK;
}
R3_EXPORT int r3_exchange_destroy(r3_ctx* c, uint32_t camera) {
    if (!c) return R3_E_INVALID;
    r3_camera* cam = r3_get_camera(c, camera);
    if (!cam) return r3_fail(c, R3_E_INVALID, "exchange_destroy: bad camera");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (uint32_t r = 0; r < cam->ex_ranks; ++r)
        if (cam->ex_connected && r != cam->ex_rank && cam->ex_peers[r]) cudaIpcCloseMemHandle(cam->ex_peers[r]);
    cudaFree(cam->d_gathered);
    cam->d_gathered = nullptr; cam->ex_connected = false; cam->ex_ranks = 0;
    for (auto& p : cam->ex_peers) p = nullptr;
    return R3_OK;
}
