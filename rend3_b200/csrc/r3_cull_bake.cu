// r3_cull_bake.cu — fused per-object frustum cull + object-uniform bake, then ordered visible-list compaction.
//
// Replaces (reference paths):
//   * uniform_prep.wgsl::cs_main          rend3-routine/shaders/src/uniform_prep.wgsl:9-27
//       MV = view * T, MVP = view_proj * T for every enabled slot < object_count
//   * the sphere/frustum filter of batch_objects   rend3-routine/src/culling/batching.rs:144-148
//       Frustum::contains_sphere, 5 planes        rend3/src/util/frustum.rs:148-161
// and emits the visible slots as one ASCENDING u32 list (the canonical, bit-exact artefact).
//
// Design (HBM-bound: 128 B read + 128 B written per object, 224 flop).
//   stream kernel  — no inter-CTA dependency, so it runs at copy speed:
//     * 8 lanes own one 128-byte object record: lane k loads float4 k, so a warp load instruction covers 512
//       contiguous bytes (4 records) and each lane keeps 8 independent 16-byte loads in flight;
//     * lanes 0-3 multiply `view` by transform column k, lanes 4-7 multiply `view_proj` by column k-4 (fetched
//       by shuffle); lane k then stores float4 k of the 128-byte MV|MVP record, so stores are as coalesced as
//       the loads.  Arithmetic is __fmul_rn/__fadd_rn in WGSL's accumulation order, never contracted: MV/MVP
//       are bit-identical to the CPU oracle;
//     * the sphere test runs one object per lane (spheres parked in shared memory by the lanes that loaded
//       them); its ballot IS the 32-bit visibility word of 32 objects (1 bit per object goes to HBM);
//     * each CTA (1024 objects) also leaves its survivor count.
//   compact kernel — one CTA per 32768 objects: sums the CTA counts in front of it (<= 40 KB, L2 resident),
//     scans its 1024 visibility words and writes the surviving slot ids in ascending order.  It moves
//     N/8 + 4*visible bytes, ~1% of the stream kernel's traffic.
// (A single-pass variant with a decoupled look-back was measured first: the look-back stalls cost 20% at 10 M
//  objects — profiles/README.md — while the streaming half alone already ran at the measured copy bandwidth.)
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

template <bool BAKE, bool CULL, bool LIVE>
__global__ void __launch_bounds__(CB_THREADS, 4)
cull_bake_kernel(const float4* __restrict__ objects, float4* __restrict__ matrices, const uint32_t* __restrict__ live_bits,
                 uint32_t* __restrict__ words, uint32_t* __restrict__ cta_counts, const __grid_constant__ CullBakeParams p) {
    __shared__ float s_mat[32];
    __shared__ float s_frustum[20];
    __shared__ float4 s_sphere[CB_WARPS][32];
    __shared__ uint32_t s_enabled[CB_WARPS][32];
    __shared__ uint32_t s_count[CB_WARPS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = lane & 7, g = lane >> 3;
    if (threadIdx.x < 32) s_mat[threadIdx.x] = threadIdx.x < 16 ? p.view[threadIdx.x] : p.view_proj[threadIdx.x - 16];
    if (threadIdx.x >= 32 && threadIdx.x < 52) s_frustum[threadIdx.x - 32] = (&p.frustum[0][0])[threadIdx.x - 32];
    __syncthreads();

    uint32_t count = 0;
#pragma unroll 1
    for (int wt = 0; wt < CB_WT; ++wt) {
        const uint32_t wtile = (blockIdx.x * CB_WARPS + warp) * CB_WT + wt;   // visibility word index
        const uint32_t base = wtile * 32u;
        if (base >= p.object_count) break;
        float4 r[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t obj = base + it * 4 + g;
            const bool want = BAKE ? true : (k == 4 || k == 7);
            r[it] = (obj < p.object_count && want) ? __ldcs(&objects[(size_t)obj * 8 + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t obj = base + it * 4 + g;
            // float4 #4 is the bounding sphere, float4 #7 carries `enabled` in .y: park them for the per-lane cull below
            if (CULL && k == 4) s_sphere[warp][it * 4 + g] = r[it];
            if (CULL && !LIVE && k == 7) s_enabled[warp][it * 4 + g] = __float_as_uint(r[it].y);
            if (BAKE) {
                const uint32_t enabled = __float_as_uint(__shfl_sync(0xFFFFFFFFu, r[it].y, (lane & 24) | 7));
                const int src = (k < 4) ? lane : lane - 4;
                const float cx = __shfl_sync(0xFFFFFFFFu, r[it].x, src), cy = __shfl_sync(0xFFFFFFFFu, r[it].y, src);
                const float cz = __shfl_sync(0xFFFFFFFFu, r[it].z, src), cw = __shfl_sync(0xFFFFFFFFu, r[it].w, src);
                const float4 o = mat_vec_rn(&s_mat[(k >> 2) * 16], cx, cy, cz, cw);
                if (obj < p.object_count && enabled != 0u) __stcs(&matrices[(size_t)obj * 8 + k], o);
            }
        }
        if (CULL) {
            __syncwarp();
            // one object per lane: Plane::distance = abc.dot(center) + d with glam's scalar dot order (util/frustum.rs:79-81,148-161)
            const uint32_t obj = base + lane;
            const float4 sp = s_sphere[warp][lane];
            bool live;
            if (LIVE) live = (__ldg(&live_bits[wtile]) >> lane) & 1u;
            else live = s_enabled[warp][lane] != 0u;
            const float neg_radius = -sp.w;
            bool inside = true;
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) {
                const float d = add_rn(add_rn(add_rn(mul_rn(s_frustum[pl * 4 + 0], sp.x), mul_rn(s_frustum[pl * 4 + 1], sp.y)),
                                              mul_rn(s_frustum[pl * 4 + 2], sp.z)), s_frustum[pl * 4 + 3]);
                inside = inside && (d >= neg_radius);
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, obj < p.object_count && live && inside);
            if (lane == 0) words[wtile] = word;
            count += __popc(word);
            __syncwarp();
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

__global__ void __launch_bounds__(CP_THREADS)
compact_visible_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ cta_counts, uint32_t n_words, uint32_t n_cta_counts,
                       uint32_t* __restrict__ visible, uint32_t* __restrict__ visible_count) {
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
    // scratch: visibility words [n_words] | stream-CTA counts [n_ctas]   (d_tile_state is a u64 array)
    R3_TRY(r3_reserve_t(c, &cam->d_tile_state, &cam->tile_state_cap, ((uint64_t)n_words + n_ctas + 3) / 2 + 1));
    uint32_t* words = reinterpret_cast<uint32_t*>(cam->d_tile_state);
    uint32_t* cta_counts = words + n_words;
    const float4* obj = reinterpret_cast<const float4*>(c->d_objects);
    float4* mats = reinterpret_cast<float4*>(cam->d_matrices);
    const bool live = c->have_live && cull;
#define R3_CB_LAUNCH(B, C, L) cull_bake_kernel<B, C, L><<<n_ctas, CB_THREADS, 0, c->stream>>>(obj, mats, c->d_live_bits, words, cta_counts, p)
    if (bake && cull) { if (live) R3_CB_LAUNCH(true, true, true); else R3_CB_LAUNCH(true, true, false); }
    else if (bake) R3_CB_LAUNCH(true, false, false);
    else { if (live) R3_CB_LAUNCH(false, true, true); else R3_CB_LAUNCH(false, true, false); }
#undef R3_CB_LAUNCH
    R3_CHECK_LAUNCH(c, "cull_bake_kernel");
    if (cull) {
        const uint32_t n_tiles = (n_words + CP_THREADS - 1) / CP_THREADS;
        compact_visible_kernel<<<n_tiles, CP_THREADS, 0, c->stream>>>(words, cta_counts, n_words, n_ctas, cam->d_visible, cam->d_visible_count);
        R3_CHECK_LAUNCH(c, "compact_visible_kernel");
    }
    return R3_OK;
}
