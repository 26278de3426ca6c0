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
//   * persistent CTAs (256 threads, 4 per SM) pull 512-object tiles from an atomic ticket;
//   * 8 lanes own one 128-byte object record: lane k loads float4 k, so a warp load instruction covers
//     512 contiguous bytes (4 records) and each lane keeps 8 independent 16-byte loads in flight;
//   * lanes 0-3 multiply `view` by transform column k, lanes 4-7 multiply `view_proj` by column k-4
//     (fetched by shuffle); lane k then stores float4 k of the 128-byte MV|MVP record: stores are as
//     coalesced as the loads.  All arithmetic is __fmul_rn/__fadd_rn in WGSL's accumulation order,
//     never contracted, so MV/MVP are bit-identical to the CPU oracle;
//   * the sphere test runs one object per lane (spheres parked in shared memory by the lane that loaded
//     them), so its ballot is directly the 32-bit visibility word of 32 objects; the CTA's
//     16 words are scanned by warp 0 and the tile's base offset comes from a decoupled look-back over
//     the preceding tiles' descriptors (single pass, no second kernel), after which every warp writes
//     its surviving slot ids in ascending order.
#include <cstdlib>

#include "r3_common.cuh"

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_WARPS = CB_THREADS / 32;
// WT = 32-object warp tiles per warp per CTA tile (template parameter): a CTA tile is 8 * WT * 32 objects

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

template <bool BAKE, bool CULL, bool LIVE, int WT, int MINB>
__global__ void __launch_bounds__(CB_THREADS, MINB)
cull_bake_kernel(const float4* __restrict__ objects, float4* __restrict__ matrices, const uint32_t* __restrict__ live_bits,
                 uint32_t* __restrict__ visible, uint32_t* __restrict__ visible_count, unsigned long long* tile_state,
                 const __grid_constant__ CullBakeParams p) {
    constexpr int CB_WTILES_PER_WARP = WT, CB_WORDS = CB_WARPS * WT, CB_TILE_OBJECTS = CB_WORDS * 32;
    static_assert(CB_WORDS <= 64, "the word scan handles at most two words per lane");
    __shared__ float s_mat[32];
    __shared__ float s_frustum[20];
    __shared__ uint32_t s_words[CB_WORDS];
    __shared__ uint32_t s_excl[CB_WORDS];
    __shared__ uint32_t s_tile, s_base;
    __shared__ float4 s_sphere[CB_WARPS][32];
    __shared__ uint32_t s_enabled[CB_WARPS][32];

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
                if (LIVE) live = (((base < p.object_count) ? __ldg(&live_bits[base >> 5]) : 0u) >> lane) & 1u;
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
                if (lane == 0) s_words[word_idx] = word;
                __syncwarp();
            }
        }
        if (!CULL) continue;
        __syncthreads();

        if (warp == 0) {
            // exclusive scan of the CB_WORDS word popcounts (up to two per lane)
            const uint32_t c0 = (2 * lane < CB_WORDS) ? __popc(s_words[2 * lane]) : 0u, c1 = (2 * lane + 1 < CB_WORDS) ? __popc(s_words[2 * lane + 1]) : 0u;
            uint32_t incl = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += n;
            }
            if (2 * lane < CB_WORDS) s_excl[2 * lane] = incl - (c0 + c1);
            if (2 * lane + 1 < CB_WORDS) s_excl[2 * lane + 1] = incl - c1;
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

template <int WT, int MINB>
static void launch_variant(r3_ctx* c, r3_camera* cam, const CullBakeParams& p, uint32_t grid, bool bake, bool cull, bool live) {
    const float4* obj = reinterpret_cast<const float4*>(c->d_objects);
    float4* mats = reinterpret_cast<float4*>(cam->d_matrices);
#define R3_CB_LAUNCH(B, C, L) \
    cull_bake_kernel<B, C, L, WT, MINB><<<grid, CB_THREADS, 0, c->stream>>>(obj, mats, c->d_live_bits, cam->d_visible, cam->d_visible_count, cam->d_tile_state, p)
    if (bake && cull) { if (live) R3_CB_LAUNCH(true, true, true); else R3_CB_LAUNCH(true, true, false); }
    else if (bake) R3_CB_LAUNCH(true, false, false);
    else { if (live) R3_CB_LAUNCH(false, true, true); else R3_CB_LAUNCH(false, true, false); }
#undef R3_CB_LAUNCH
}

int r3_launch_cull_bake(r3_ctx* c, r3_camera* cam, uint32_t mode) {
    const uint32_t n = cam->header.object_count;
    const bool bake = mode & R3_CB_BAKE, cull = mode & R3_CB_CULL;
    if (cull) R3_CUDA(c, cudaMemsetAsync(cam->d_visible_count, 0, 4, c->stream));
    if (n == 0 || (!bake && !cull)) return R3_OK;
    // tuning knobs (profiles/README.md): R3_CB_WT = warp tiles per warp (tile = 256*WT objects), R3_CB_MINB = CTAs per SM
    static const int wt = getenv("R3_CB_WT") ? atoi(getenv("R3_CB_WT")) : 8;
    static const int minb = getenv("R3_CB_MINB") ? atoi(getenv("R3_CB_MINB")) : 3;
    const uint32_t tile_objects = CB_WARPS * 32u * (uint32_t)(wt == 2 ? 2 : wt == 4 ? 4 : 8);
    CullBakeParams p;
    memcpy(p.view, cam->header.view, 64);
    memcpy(p.view_proj, cam->header.view_proj, 64);
    memcpy(p.frustum, cam->header.frustum, 80);
    p.object_count = n;
    p.n_tiles = (n + tile_objects - 1) / tile_objects;
    R3_TRY(r3_reserve_t(c, &cam->d_tile_state, &cam->tile_state_cap, (uint64_t)p.n_tiles + 1));
    R3_CUDA(c, cudaMemsetAsync(cam->d_tile_state, 0, ((size_t)p.n_tiles + 1) * 8, c->stream));
    const uint32_t resident = (uint32_t)(R3_SM_COUNT * (minb == 4 ? 4 : 3));
    const uint32_t grid = p.n_tiles < resident ? p.n_tiles : resident;
    const bool live = c->have_live && cull;
    if (wt == 2) { if (minb == 4) launch_variant<2, 4>(c, cam, p, grid, bake, cull, live); else launch_variant<2, 3>(c, cam, p, grid, bake, cull, live); }
    else if (wt == 4) { if (minb == 4) launch_variant<4, 4>(c, cam, p, grid, bake, cull, live); else launch_variant<4, 3>(c, cam, p, grid, bake, cull, live); }
    else { if (minb == 4) launch_variant<8, 4>(c, cam, p, grid, bake, cull, live); else launch_variant<8, 3>(c, cam, p, grid, bake, cull, live); }
    R3_CHECK_LAUNCH(c, "cull_bake_kernel");
    return R3_OK;
}
