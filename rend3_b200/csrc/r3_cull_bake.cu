// r3_cull_bake.cu — fused per-object frustum cull + object-uniform bake + ordered visible-list compaction.
//
// Replaces (reference paths):
//   * uniform_prep.wgsl::cs_main          rend3-routine/shaders/src/uniform_prep.wgsl:9-27
//       MV = view * T, MVP = view_proj * T for every enabled slot < object_count
//   * the sphere/frustum filter of batch_objects   rend3-routine/src/culling/batching.rs:144-148
//       Frustum::contains_sphere, 5 planes        rend3/src/util/frustum.rs:148-161
// and emits the visible slots as one ASCENDING u32 list (the canonical, bit-exact artefact).
//
// Design (HBM-bound: 128 B read + 128 B written per object, 224 flop):
//   * one persistent CTA (256 threads) per SM slot pulls 2048-object tiles from an atomic ticket;
//   * 8 lanes own one 128-byte object record: lane k loads float4 k, so a warp load instruction covers
//     512 contiguous bytes (4 records) and each lane keeps 8 independent 16-byte loads in flight;
//   * lanes 0-3 multiply `view` by transform column k, lanes 4-7 multiply `view_proj` by column k-4
//     (fetched by shuffle); lane k then stores float4 k of the 128-byte MV|MVP record: stores are as
//     coalesced as the loads.  All arithmetic is __fmul_rn/__fadd_rn in WGSL's accumulation order,
//     never contracted, so MV/MVP are bit-identical to the CPU oracle;
//   * visibility is a warp ballot -> one 32-bit word per 32 objects, kept in shared memory; the CTA's
//     64 words are scanned by warp 0 and the tile's base offset comes from a decoupled look-back over
//     the preceding tiles' descriptors (single pass, no second kernel), after which every warp writes
//     its surviving slot ids in ascending order.
#include "r3_common.cuh"

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_WARPS = CB_THREADS / 32;
constexpr int CB_WTILES_PER_WARP = 8;                       // 32-object warp tiles per warp per CTA tile
constexpr int CB_WORDS = CB_WARPS * CB_WTILES_PER_WARP;     // 64 visibility words per CTA tile
constexpr int CB_TILE_OBJECTS = CB_WORDS * 32;              // 2048

struct CullBakeParams {
    float view[16];
    float view_proj[16];
    float frustum[5][4];
    uint32_t object_count;
    uint32_t n_tiles;
};

constexpr unsigned long long DESC_AGGREGATE = 1ull << 32, DESC_PREFIX = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

template <bool BAKE, bool CULL, bool LIVE>
__global__ void __launch_bounds__(CB_THREADS)
cull_bake_kernel(const float4* __restrict__ objects, float4* __restrict__ matrices, const uint32_t* __restrict__ live_bits,
                 uint32_t* __restrict__ visible, uint32_t* __restrict__ visible_count, unsigned long long* tile_state,
                 const __grid_constant__ CullBakeParams p) {
    __shared__ float s_mat[32];
    __shared__ float s_frustum[20];
    __shared__ uint32_t s_words[CB_WORDS];
    __shared__ uint32_t s_excl[CB_WORDS];
    __shared__ uint32_t s_tile, s_base;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = lane & 7, g = lane >> 3;
    if (threadIdx.x < 32) s_mat[threadIdx.x] = threadIdx.x < 16 ? p.view[threadIdx.x] : p.view_proj[threadIdx.x - 16];
    if (threadIdx.x >= 32 && threadIdx.x < 52) s_frustum[threadIdx.x - 32] = (&p.frustum[0][0])[threadIdx.x - 32];

    for (;;) {
        __syncthreads();   // protects s_tile / s_words reuse across tiles (and publishes s_mat the first time)
        if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(&tile_state[0], 1ull);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= p.n_tiles) return;

#pragma unroll 1
        for (int wt = 0; wt < CB_WTILES_PER_WARP; ++wt) {
            const uint32_t word_idx = warp * CB_WTILES_PER_WARP + wt;
            const uint32_t base = tile * CB_TILE_OBJECTS + word_idx * 32;
            float4 r[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t obj = base + it * 4 + g;
                const bool want = BAKE ? true : (k == 4 || k == 7);
                r[it] = (obj < p.object_count && want) ? __ldcs(&objects[(size_t)obj * 8 + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t live_word = 0xFFFFFFFFu;
            if (LIVE && CULL) live_word = (base < p.object_count) ? __ldg(&live_bits[base >> 5]) : 0u;
            uint32_t word = 0;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t obj = base + it * 4 + g;
                const uint32_t enabled = __float_as_uint(__shfl_sync(0xFFFFFFFFu, r[it].y, (lane & 24) | 7));
                if (BAKE) {
                    const int src = (k < 4) ? lane : lane - 4;
                    const float cx = __shfl_sync(0xFFFFFFFFu, r[it].x, src), cy = __shfl_sync(0xFFFFFFFFu, r[it].y, src);
                    const float cz = __shfl_sync(0xFFFFFFFFu, r[it].z, src), cw = __shfl_sync(0xFFFFFFFFu, r[it].w, src);
                    const float4 o = mat_vec_rn(&s_mat[(k >> 2) * 16], cx, cy, cz, cw);
                    if (obj < p.object_count && enabled != 0u) __stcs(&matrices[(size_t)obj * 8 + k], o);
                }
                if (CULL) {
                    bool vis = false;
                    if (k == 4 && obj < p.object_count) {
                        const bool live = LIVE ? ((live_word >> (it * 4 + g)) & 1u) : (enabled != 0u);
                        // Plane::distance = abc.dot(center) + d with glam's scalar dot order (util/frustum.rs:79-81)
                        const float neg_radius = -r[it].w;
                        bool inside = true;
#pragma unroll
                        for (int pl = 0; pl < 5; ++pl) {
                            const float d = add_rn(add_rn(add_rn(mul_rn(s_frustum[pl * 4 + 0], r[it].x), mul_rn(s_frustum[pl * 4 + 1], r[it].y)),
                                                          mul_rn(s_frustum[pl * 4 + 2], r[it].z)), s_frustum[pl * 4 + 3]);
                            inside = inside && (d >= neg_radius);
                        }
                        vis = live && inside;
                    }
                    const uint32_t b = __ballot_sync(0xFFFFFFFFu, vis);     // bits 4,12,20,28
                    word |= (((b >> 4) & 1u) | ((b >> 11) & 2u) | ((b >> 18) & 4u) | ((b >> 25) & 8u)) << (it * 4);
                }
            }
            if (CULL && lane == 0) s_words[word_idx] = word;
        }
        if (!CULL) continue;
        __syncthreads();

        if (warp == 0) {
            // exclusive scan of the 64 word popcounts (2 per lane)
            const uint32_t c0 = __popc(s_words[2 * lane]), c1 = __popc(s_words[2 * lane + 1]);
            uint32_t incl = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += n;
            }
            const uint32_t excl = incl - (c0 + c1);
            s_excl[2 * lane] = excl;
            s_excl[2 * lane + 1] = excl + c0;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            // decoupled look-back for the tile's base offset
            uint32_t running = 0;
            if (tile > 0) {
                if (lane == 0) st_desc(&tile_state[1 + tile], DESC_AGGREGATE | total);
                int pred = (int)tile - 1;
                for (;;) {
                    const int idx = pred - lane;
                    unsigned long long d = (idx >= 0) ? ld_desc(&tile_state[1 + idx]) : (DESC_PREFIX | 0ull);
                    while (__any_sync(0xFFFFFFFFu, (d >> 32) == 0ull)) {
                        if ((d >> 32) == 0ull) d = ld_desc(&tile_state[1 + idx]);
                    }
                    const uint32_t pmask = __ballot_sync(0xFFFFFFFFu, (d >> 32) == 2ull);
                    const int first = pmask ? (__ffs(pmask) - 1) : 31;
                    uint32_t v = (lane <= first) ? (uint32_t)d : 0u;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, s);
                    running += v;
                    if (pmask) break;
                    pred -= 32;
                }
            }
            if (lane == 0) {
                st_desc(&tile_state[1 + tile], DESC_PREFIX | (unsigned long long)(running + total));
                s_base = running;
                if (tile == p.n_tiles - 1) *visible_count = running + total;
            }
        }
        __syncthreads();
        const uint32_t tile_base = s_base;
#pragma unroll 1
        for (int wt = 0; wt < CB_WTILES_PER_WARP; ++wt) {
            const uint32_t word_idx = warp * CB_WTILES_PER_WARP + wt;
            const uint32_t word = s_words[word_idx];
            if ((word >> lane) & 1u) {
                const uint32_t dst = tile_base + s_excl[word_idx] + __popc(word & ((1u << lane) - 1u));
                visible[dst] = tile * CB_TILE_OBJECTS + word_idx * 32 + lane;
            }
        }
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
    p.n_tiles = (n + CB_TILE_OBJECTS - 1) / CB_TILE_OBJECTS;
    R3_TRY(r3_reserve_t(c, &cam->d_tile_state, &cam->tile_state_cap, (uint64_t)p.n_tiles + 1));
    R3_CUDA(c, cudaMemsetAsync(cam->d_tile_state, 0, ((size_t)p.n_tiles + 1) * 8, c->stream));
    const uint32_t grid = p.n_tiles < (uint32_t)(R3_SM_COUNT * 8) ? p.n_tiles : (uint32_t)(R3_SM_COUNT * 8);
    const float4* obj = reinterpret_cast<const float4*>(c->d_objects);
    float4* mats = reinterpret_cast<float4*>(cam->d_matrices);
    const bool live = c->have_live && cull;
#define R3_CB_LAUNCH(B, C, L) \
    cull_bake_kernel<B, C, L><<<grid, CB_THREADS, 0, c->stream>>>(obj, mats, c->d_live_bits, cam->d_visible, cam->d_visible_count, cam->d_tile_state, p)
    if (bake && cull) { if (live) R3_CB_LAUNCH(true, true, true); else R3_CB_LAUNCH(true, true, false); }
    else if (bake) R3_CB_LAUNCH(true, false, false);
    else { if (live) R3_CB_LAUNCH(false, true, true); else R3_CB_LAUNCH(false, true, false); }
#undef R3_CB_LAUNCH
    R3_CHECK_LAUNCH(c, "cull_bake_kernel");
    return R3_OK;
}
