// r3_gpu_batching.cu — batch_objects on the device: sort the visible objects and build the ShaderBatchData records
// without leaving the GPU, so a frame needs no mid-frame host synchronisation.
//
// Replaces the host half of GpuCuller::add_culling_to_graph (rend3-routine/src/culling/culler.rs:682-713):
//   batch_objects            rend3-routine/src/culling/batching.rs:120-250
//   ShaderJobSortingKey::cmp batching.rs:53-79   (material key, sorting reason, distance; bind group is DUMMY)
// The reference does this single-threaded on the CPU, per camera, under the data_core mutex — the object-level
// bottleneck SURVEY §8 (row a4) calls out.  Here:
//   1. key generation: 64-bit key = [material_key:6 | reason:1 | sortable(distance²):32 | position in the visible list:24];
//   2. stable LSD radix sort on the 39 significant bits (5 passes of 8 bits; histogram, scan, stable scatter with
//      __match_any_sync ranking) — ties keep the ascending-handle order of the visible list, the same tie rule the
//      oracle and the host path use;
//   3. batch build: one CTA per 256 sorted objects (block scans give invocation_start, region boundaries, local ids),
//      one scan over the batches (batch_base_invocation, global region ids), one fix-up pass that also swaps the
//      per-camera previous-invocation map (batching.rs:226,230).
// All counts stay on the device in the job header; the cull / raster kernels read them there.
// Not covered on the device (the host path remains for them): material keys >= 64, more than 2^24 visible objects, a
// single batch exceeding max_compute_workgroups_per_dimension*256 invocations (reported through the overflow flag).
#include <cstring>

#include <cooperative_groups.h>

#include "r3_common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_KEYS_PER_THREAD = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_KEYS_PER_THREAD;   // 2048 keys per block
constexpr int KEY_SHIFT0 = 24, SORT_PASSES = 5;

// order-preserving map of OrderedFloat's total order (batching.rs:37): every NaN is one value above +inf, -0.0 == +0.0
__device__ __forceinline__ uint32_t sortable_f32(float f) {
    if (f != f) return 0xFFC00000u;                              // the image of the canonical quiet NaN 0x7FC00000
    const uint32_t b = __float_as_uint(f) == 0x80000000u ? 0u : __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ unsigned long long make_sort_key(const uint32_t* __restrict__ visible, const uint8_t* __restrict__ key8, const float* __restrict__ loc,
                                                            float vx, float vy, float vz, uint32_t j) {
    const uint32_t h = visible ? visible[j] : j;                 // visible == nullptr: the frame-wide sort over every slot (see r3_device_batch_objects)
    const uint8_t k = key8[h];                                   // (material_key << 1 | reason) << 1 | back_to_front
    const float dx = sub_rn(vx, loc[3 * (size_t)h]), dy = sub_rn(vy, loc[3 * (size_t)h + 1]), dz = sub_rn(vz, loc[3 * (size_t)h + 2]);
    float d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));   // Vec3A::distance_squared (batching.rs:156-157)
    if (k & 1u) d2 = -d2;                                                        // SortingOrder::BackToFront (batching.rs:158-160)
    return ((unsigned long long)(k >> 1) << 56) | ((unsigned long long)sortable_f32(d2) << 24) | (unsigned long long)j;
}

// (Measured: raising the limit to 12288 keys — 225 KB of shared memory — puts the 10 000-object config on this path and makes its sort
//  SLOWER, 85 us against 61 us for the cooperative kernel on 5 CTAs; the limit stays at 8192.)
// cap <= SMALL_SORT_MAX: key generation + the same stable LSD radix sort, but by ONE CTA entirely in shared memory — one launch
// instead of 16, which is what a frame with a few thousand objects per camera is made of.  Warp w owns the w-th contiguous
// chunk of the keys; per pass: per-warp digit counts (__match_any_sync, leader adds), a scan over (digit, warp), then the
// warps re-walk their chunks in order and scatter (rank inside the round from the peer mask).  A pass whose digit is the same
// for every key (typical for the material-key byte) is skipped.
constexpr uint32_t SMALL_SORT_MAX = 8192;
constexpr int SMALL_SORT_THREADS = 1024, SMALL_SORT_WARPS = SMALL_SORT_THREADS / 32;
__host__ __device__ inline uint32_t small_sort_pad(uint32_t n) { return ((n + SMALL_SORT_THREADS - 1) / SMALL_SORT_THREADS) * SMALL_SORT_THREADS; }
inline size_t small_sort_smem(uint32_t cap) { return (size_t)small_sort_pad(cap) * 16 + (size_t)SMALL_SORT_WARPS * 256 * 4 + 256 * 4; }
__global__ void __launch_bounds__(SMALL_SORT_THREADS) small_sort_kernel(const uint32_t* __restrict__ visible, const uint32_t* __restrict__ visible_count,
                                                                        const uint8_t* __restrict__ key8, const float* __restrict__ loc, float vx, float vy, float vz,
                                                                        unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ header, uint32_t cap_pad, uint32_t count_imm) {
    extern __shared__ unsigned long long s_keys[];                               // [2][cap_pad]
    uint32_t* s_count = reinterpret_cast<uint32_t*>(s_keys + 2 * (size_t)cap_pad);   // [warps][256]
    uint32_t* s_base = s_count + SMALL_SORT_WARPS * 256;                          // [256]
    __shared__ uint32_t s_wsum[8];
    __shared__ int s_skip;
    const uint32_t nv = min(visible_count ? *visible_count : count_imm, cap_pad);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { header[0] = nv; header[4] = 0u; }
    for (uint32_t j = threadIdx.x; j < nv; j += SMALL_SORT_THREADS) s_keys[j] = make_sort_key(visible, key8, loc, vx, vy, vz, j);
    const uint32_t rounds = (nv + SMALL_SORT_THREADS - 1) / SMALL_SORT_THREADS, chunk = rounds * 32u;
    int src = 0;
    for (int pass = 0; pass < SORT_PASSES; ++pass) {
        const int shift = KEY_SHIFT0 + 8 * pass;
        const unsigned long long* in = s_keys + (size_t)src * cap_pad;
        unsigned long long* out = s_keys + (size_t)(src ^ 1) * cap_pad;
        for (uint32_t i = threadIdx.x; i < SMALL_SORT_WARPS * 256; i += SMALL_SORT_THREADS) s_count[i] = 0u;
        if (threadIdx.x == 0) s_skip = 0;
        __syncthreads();
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t idx = warp * chunk + r * 32u + lane;
            const bool valid = idx < nv;
            const uint32_t digit = valid ? (uint32_t)(in[idx] >> shift) & 255u : 256u + lane;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
            if (valid && lane == __ffs(peers) - 1) s_count[warp * 256 + digit] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        uint32_t total = 0, incl = 0;
        if (threadIdx.x < 256) {
            for (int w = 0; w < SMALL_SORT_WARPS; ++w) { const uint32_t t = s_count[w * 256 + threadIdx.x]; s_count[w * 256 + threadIdx.x] = total; total += t; }
            if (total == nv) s_skip = 1;
            incl = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += n; }
            if (lane == 31) s_wsum[warp] = incl;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += s_wsum[w];
            s_base[threadIdx.x] = before + incl - total;
        }
        __syncthreads();
        const int skip = s_skip;
        __syncthreads();        // s_skip is reset at the top of the next pass
        if (skip) continue;     // block-uniform
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t idx = warp * chunk + r * 32u + lane;
            const bool valid = idx < nv;
            const unsigned long long key = valid ? in[idx] : 0ull;
            const uint32_t digit = valid ? (uint32_t)(key >> shift) & 255u : 256u + lane;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
            if (valid) {
                const uint32_t pos = s_base[digit] + s_count[warp * 256 + digit] + __popc(peers & ((1u << lane) - 1u));
                out[pos] = key;
            }
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) s_count[warp * 256 + digit] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        src ^= 1;
    }
    const unsigned long long* fin = s_keys + (size_t)src * cap_pad;
    for (uint32_t j = threadIdx.x; j < nv; j += SMALL_SORT_THREADS) keys_out[j] = fin[j];
}

__global__ void keygen_kernel(const uint32_t* __restrict__ visible, const uint32_t* __restrict__ visible_count, const uint8_t* __restrict__ key8,
                              const float* __restrict__ loc, float vx, float vy, float vz, unsigned long long* __restrict__ keys, uint32_t* __restrict__ header,
                              uint32_t count_imm) {
    const uint32_t nv = visible_count ? *visible_count : count_imm;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) { header[0] = nv; header[4] = (nv >= (1u << 24)) ? 1u : 0u; }
    if (j >= nv) return;
    keys[j] = make_sort_key(visible, key8, loc, vx, vy, vz, j);
}

__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ header,
                                                                  int shift, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[256];
    const uint32_t nv = header[0], nb = gridDim.x;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_KEYS_PER_THREAD; ++r) {
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        if (i < nv) atomicAdd(&s_hist[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nb + blockIdx.x] = s_hist[threadIdx.x];   // digit-major: a flat scan yields the scatter bases
}

// in-place exclusive scan by one block: every thread owns a contiguous run (serial sum, one block scan of the 1024 run totals,
// serial write-back) — three barriers in all instead of four per 1024 elements
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ data, uint32_t n) {
    __shared__ uint32_t s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (((n + 1023u) / 1024u) + 3u) & ~3u;          // multiple of 4: runs start 16-byte aligned
    const uint32_t lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    uint32_t sum = 0;
    uint32_t i = lo;
    for (; i + 4 <= hi; i += 4) { const uint4 v = *reinterpret_cast<const uint4*>(data + i); sum += v.x + v.y + v.z + v.w; }
    for (; i < hi; ++i) sum += data[i];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += t; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
    for (i = lo; i + 4 <= hi; i += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(data + i);
        uint4 o;
        o.x = run; o.y = o.x + v.x; o.z = o.y + v.y; o.w = o.z + v.z; run = o.w + v.w;
        *reinterpret_cast<uint4*>(data + i) = o;
    }
    for (; i < hi; ++i) { const uint32_t v = data[i]; data[i] = run; run += v; }
}

__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                                                                     const uint32_t* __restrict__ header, int shift, const uint32_t* __restrict__ hist_scanned) {
    __shared__ uint32_t s_cnt[SORT_THREADS / 32][256];   // per-warp digit counts of the current round
    __shared__ uint32_t s_run[256];                       // digits already emitted by this block in earlier rounds
    __shared__ uint32_t s_gbase[256];
    const uint32_t nv = header[0], nb = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_run[threadIdx.x] = 0;
    s_gbase[threadIdx.x] = hist_scanned[threadIdx.x * nb + blockIdx.x];
    const uint32_t base = blockIdx.x * SORT_TILE;
    if (base >= nv) return;
#pragma unroll 1
    for (int r = 0; r < SORT_KEYS_PER_THREAD; ++r) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) s_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        const bool valid = i < nv;
        const unsigned long long key = valid ? keys_in[i] : 0ull;
        const uint32_t d = valid ? ((uint32_t)(key >> shift) & 0xFFu) : 0x100u;     // invalid lanes form their own group
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) s_cnt[warp][d] = __popc(peers);              // one writer per (warp, digit)
        __syncthreads();
        // digit t: exclusive prefix over the warps, then advance the block's running count
        {
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) { const uint32_t c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = acc; acc += c; }
            __syncthreads();
            if (valid) keys_out[s_gbase[d] + s_run[d] + s_cnt[warp][d] + rank_in_warp] = key;
            __syncthreads();
            s_run[threadIdx.x] += acc;
        }
        __syncthreads();
    }
}

// ---- batch build (batching.rs:180-246), one CTA per 256 sorted objects
struct BuildParams {
    const unsigned long long* keys; const uint32_t* visible; const r3_object* objects;
    r3_batch_data* batches; uint32_t* header;
    uint32_t* batch_inv; uint32_t* batch_regions;         // per batch: total_invocations, number of regions
    uint32_t* region_key; uint32_t* region_start;         // per (batch, batch-local region): material key, first invocation in the batch
    r3_region* regions; uint32_t* region_first_inv;
    const uint32_t* prev_map; uint32_t* cur_map; uint32_t map_cap;
    uint32_t n_batches_cap; uint64_t dispatch_limit;
    uint32_t keys_hold_slots;                             // low 24 key bits = the object slot (frame-wide sort) instead of a position in `visible`
};

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; if (w < warp) wbase += c; tot += c; }
    __syncthreads();
    *total = tot;
    return wbase + incl - v;
}

// Mid-sized worlds (more than one CTA's worth, up to 256 tiles = 524288 visible objects): the whole sort — key generation and
// the five histogram / offset / scatter passes — in ONE cooperative launch with grid-wide barriers between the phases instead of
// 16 launches.  Every CTA derives its own scatter bases from the raw digit-major histogram table (one pass over <= 256 columns
// per thread), so no separate scan kernel runs.
constexpr uint32_t COOP_SORT_MAX_BLOCKS = 256;
struct SortCoopParams {
    const uint32_t* visible; const uint32_t* visible_count; const uint8_t* key8; const float* loc; float vx, vy, vz;
    unsigned long long* keys[2]; uint32_t* hist; uint32_t* header; uint32_t count_imm;
};
__global__ void __launch_bounds__(SORT_THREADS) radix_sort_coop_kernel(const __grid_constant__ SortCoopParams p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_cnt[SORT_THREADS / 32][256];
    __shared__ uint32_t s_run[256];
    __shared__ uint32_t s_gbase[256];
    __shared__ uint32_t s_warp[8];
    const uint32_t nv = p.visible_count ? *p.visible_count : p.count_imm, nb = gridDim.x, b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = b * SORT_TILE;
    if (b == 0 && threadIdx.x == 0) { p.header[0] = nv; p.header[4] = (nv >= (1u << 24)) ? 1u : 0u; }
#pragma unroll 1
    for (int r = 0; r < SORT_KEYS_PER_THREAD; ++r) {
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        if (i < nv) p.keys[0][i] = make_sort_key(p.visible, p.key8, p.loc, p.vx, p.vy, p.vz, i);
    }
    __syncthreads();
    int src = 0;
    for (int pass = 0; pass < SORT_PASSES; ++pass) {
        const int shift = KEY_SHIFT0 + 8 * pass;
        const unsigned long long* keys_in = p.keys[src];
        unsigned long long* keys_out = p.keys[src ^ 1];
        // histogram of this tile
        s_hist[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_KEYS_PER_THREAD; ++r) {
            const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
            if (i < nv) atomicAdd(&s_hist[(uint32_t)(keys_in[i] >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        p.hist[threadIdx.x * nb + b] = s_hist[threadIdx.x];
        grid.sync();
        // scatter bases of this tile: digits below mine in every tile + my digit in the tiles in front of me
        {
            uint32_t before = 0, total = 0;
            const uint32_t* col = p.hist + threadIdx.x * nb;
            for (uint32_t k = 0; k < nb; ++k) { const uint32_t v = col[k]; total += v; if (k < b) before += v; }
            uint32_t tot_all;
            const uint32_t digit_base = block_scan_excl(total, s_warp, &tot_all);
            s_gbase[threadIdx.x] = digit_base + before;
            s_run[threadIdx.x] = 0;
        }
        __syncthreads();
        // stable scatter, 256 keys per round
#pragma unroll 1
        for (int r = 0; r < SORT_KEYS_PER_THREAD; ++r) {
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) s_cnt[w][threadIdx.x] = 0;
            __syncthreads();
            const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
            const bool valid = i < nv;
            const unsigned long long key = valid ? keys_in[i] : 0ull;
            const uint32_t d = valid ? ((uint32_t)(key >> shift) & 0xFFu) : 0x100u;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
            const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank_in_warp == 0) s_cnt[warp][d] = __popc(peers);
            __syncthreads();
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) { const uint32_t c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = acc; acc += c; }
            __syncthreads();
            if (valid) keys_out[s_gbase[d] + s_run[d] + s_cnt[warp][d] + rank_in_warp] = key;
            __syncthreads();
            s_run[threadIdx.x] += acc;
            __syncthreads();
        }
        grid.sync();
        src ^= 1;
    }
}

__global__ void __launch_bounds__(256) batch_build_kernel(const __grid_constant__ BuildParams p) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_start[256];
    __shared__ uint32_t s_first[256];
    const uint32_t nv = p.header[0], b = blockIdx.x, i = threadIdx.x, j = b * 256u + i;
    if (b * 256u >= nv) { if (i == 0) { p.batch_inv[b] = 0; p.batch_regions[b] = 0; } return; }
    const bool valid = j < nv;
    unsigned long long key = 0ull, prev_key = 0ull;
    uint32_t h = 0, tri = 0;
    if (valid) {
        key = p.keys[j];
        h = p.keys_hold_slots ? (uint32_t)(key & 0xFFFFFFull) : p.visible[(uint32_t)(key & 0xFFFFFFull)];
        tri = p.objects[h].index_count / 3u;                                  // batching.rs:192
        if (j > 0) prev_key = p.keys[j - 1];
    }
    const uint32_t k7 = (uint32_t)(key >> 56), mat = k7 >> 1, prev_mat = (uint32_t)(prev_key >> 57);
    const uint32_t padded = valid ? ((tri + 255u) & ~255u) : 0u;              // round_up(invocation_count, WORKGROUP_SIZE) batching.rs:235
    uint32_t total_inv, n_regions;
    const uint32_t start = block_scan_excl(padded, s_warp, &total_inv);
    const uint32_t flag = (valid && (i == 0 || mat != prev_mat)) ? 1u : 0u;  // a region starts at a batch start or a key change (batching.rs:194-204)
    const uint32_t region_local = block_scan_excl(flag, s_warp, &n_regions) + flag - 1u;
    s_start[i] = start;
    // index of the first object of my region: running max of (flag ? i : 0)
    uint32_t first = flag ? i : 0u;
    {
        const int lane = i & 31, warp = i >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, first, d); if (lane >= d) first = max(first, t); }
        if (lane == 31) s_warp[warp] = first;
        __syncthreads();
        uint32_t wmax = 0;
        for (int w = 0; w < warp; ++w) wmax = max(wmax, s_warp[w]);
        first = max(first, wmax);
        __syncthreads();
    }
    s_first[i] = first;
    __syncthreads();
    r3_batch_data* bd = &p.batches[b];
    if (valid) {
        r3_object_culling_info info;
        info.invocation_start = start;
        info.invocation_end = start + tri;
        info.object_id = h;
        info.region_id = region_local;                     // made global by batch_finalize_kernel
        info.base_region_invocation = s_start[first];
        info.local_region_id = i - first;
        info.previous_global_invocation = (h < p.map_cap) ? p.prev_map[h] : R3_NO_PREVIOUS;   // batching.rs:226
        info.atomic_capable = (k7 & 1u) ? 0u : 1u;         // SortingReason::Optimization (batching.rs:227)
        bd->object_culling_information[i] = info;
        if (flag) { p.region_key[b * 256u + region_local] = mat; p.region_start[b * 256u + region_local] = start; }
    }
    if (i == 0) {
        const uint32_t n_obj = min(256u, nv - b * 256u);
        bd->total_objects = n_obj; bd->total_invocations = total_inv; bd->batch_base_invocation = 0;
        p.batch_inv[b] = total_inv; p.batch_regions[b] = n_regions;
        if ((uint64_t)total_inv >= p.dispatch_limit) atomicExch(&p.header[4], 1u);   // batching.rs:196 would have split the batch
    }
}

// one block: exclusive scans over the batches -> batch_base_invocation, global region ids; fills the header
__global__ void __launch_bounds__(1024) batch_scan_kernel(const __grid_constant__ BuildParams p) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry_inv, s_carry_reg;
    const uint32_t nv = p.header[0], nb = (nv + 255u) / 256u;
    if (threadIdx.x == 0) { s_carry_inv = 0; s_carry_reg = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t vi = b < nb ? p.batch_inv[b] : 0u, vr = b < nb ? p.batch_regions[b] : 0u;
        uint32_t ii = vi, ir = vr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ti = __shfl_up_sync(0xFFFFFFFFu, ii, d), tr = __shfl_up_sync(0xFFFFFFFFu, ir, d);
            if (lane >= d) { ii += ti; ir += tr; }
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = ii;
        __syncthreads();
        uint32_t wi = 0;
        for (int w = 0; w < warp; ++w) wi += s_warp[w];
        uint32_t tot_i = 0;
        for (int w = 0; w < 32; ++w) tot_i += s_warp[w];
        __syncthreads();
        if (lane == 31) s_warp[warp] = ir;
        __syncthreads();
        uint32_t wr = 0;
        for (int w = 0; w < warp; ++w) wr += s_warp[w];
        uint32_t tot_r = 0;
        for (int w = 0; w < 32; ++w) tot_r += s_warp[w];
        if (b < nb) { p.batch_inv[b] = s_carry_inv + wi + ii - vi; p.batch_regions[b] = s_carry_reg + wr + ir - vr; }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry_inv += tot_i; s_carry_reg += tot_r; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.header[1] = nb; p.header[2] = s_carry_reg; p.header[3] = s_carry_inv;
        p.region_first_inv[s_carry_reg] = s_carry_inv;
    }
}

__global__ void __launch_bounds__(256) batch_finalize_kernel(const __grid_constant__ BuildParams p) {
    const uint32_t nv = p.header[0], b = blockIdx.x, i = threadIdx.x, j = b * 256u + i;
    if (b * 256u >= nv) return;
    r3_batch_data* bd = &p.batches[b];
    const uint32_t base_inv = p.batch_inv[b], base_reg = p.batch_regions[b];
    if (i == 0) bd->batch_base_invocation = base_inv;
    if (j < nv) {
        r3_object_culling_info* info = &bd->object_culling_information[i];
        const uint32_t local_region = info->region_id;
        info->region_id = base_reg + local_region;
        if (info->object_id < p.map_cap) p.cur_map[info->object_id] = info->invocation_start + base_inv;   // batching.rs:230
        if (info->local_region_id == 0) {
            r3_region r;
            r.job_index = b; r.bind_group_index = 0u; r.material_key = p.region_key[b * 256u + local_region];
            p.regions[base_reg + local_region] = r;                                                            // batching.rs:199,238
            p.region_first_inv[base_reg + local_region] = base_inv + p.region_start[b * 256u + local_region];
        }
    }
}

// ---- frame-wide sort shared by the cameras of a frame.  batch_objects sorts by (material key, sorting reason, distance to the
// VIEWPORT camera) for every camera, shadow cameras included (batching.rs:156-157 uses viewport_camera_state) — the key of an object
// is the same in all of them, only the visible sets differ.  So the slots are sorted ONCE per frame and each camera takes its
// visible objects out of that order with a stream compaction (ties resolve by slot in both forms: the results are identical).
// On config 3 this replaces five 148-us cooperative sorts by one sort and five ~10-us compactions.
constexpr int RC_THREADS = 1024;
__device__ __forceinline__ bool rank_visible(const unsigned long long* __restrict__ gkeys, uint32_t r, uint32_t n, const uint32_t* __restrict__ words, uint32_t cap,
                                             unsigned long long* key) {
    if (r >= n) return false;
    const unsigned long long k = gkeys[r];
    const uint32_t slot = (uint32_t)(k & 0xFFFFFFull);
    *key = k;
    return slot < cap && ((__ldg(&words[slot >> 5]) >> (slot & 31u)) & 1u);
}
__global__ void __launch_bounds__(RC_THREADS) rank_count_kernel(const unsigned long long* __restrict__ gkeys, uint32_t n, const uint32_t* __restrict__ words, uint32_t cap,
                                                                uint32_t* __restrict__ tile_counts) {
    unsigned long long k;
    const int c = __syncthreads_count(rank_visible(gkeys, blockIdx.x * RC_THREADS + threadIdx.x, n, words, cap, &k) ? 1 : 0);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (uint32_t)c;
}
__global__ void __launch_bounds__(RC_THREADS) rank_scatter_kernel(const unsigned long long* __restrict__ gkeys, uint32_t n, const uint32_t* __restrict__ words, uint32_t cap,
                                                                  const uint32_t* __restrict__ tile_counts, unsigned long long* __restrict__ keys_out,
                                                                  uint32_t* __restrict__ header) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t before = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += RC_THREADS) before += __ldg(&tile_counts[t]);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, sft);
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane];
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, sft);
        if (lane == 0) s_base = v;
    }
    __syncthreads();
    const uint32_t base = s_base;
    __syncthreads();
    unsigned long long key = 0ull;
    const bool vis = rank_visible(gkeys, blockIdx.x * RC_THREADS + threadIdx.x, n, words, cap, &key);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, vis);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) { const uint32_t c = s_warp[w]; if (w < warp) wbase += c; total += c; }
    if (vis) keys_out[base + wbase + __popc(bal & ((1u << lane) - 1u))] = key;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { header[0] = base + total; header[4] = 0u; }
}

// out[0] = sum over the slots of round_up(triangles, 256), out[1] = the largest such term
__global__ void max_invocations_kernel(const r3_object* __restrict__ objects, uint32_t n, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0, big = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long v = ((objects[i].index_count / 3u) + 255u) & ~255u;
        acc += v; big = max(big, v);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s); big = max(big, __shfl_xor_sync(0xFFFFFFFFu, big, s)); }
    if ((threadIdx.x & 31) == 0 && acc) { atomicAdd(out, acc); atomicMax(out + 1, big); }
}

}  // namespace

// sum over every slot of round_up(index_count / 3, 256): the bound the culling buffers are sized with when the per-frame
// totals stay on the device.  One small reduction + 8-byte readback at upload time, never per frame.
int r3_compute_max_invocations(r3_ctx* c) {
    if (c->max_invocations_valid) return R3_OK;
    c->max_total_invocations = 0; c->max_object_invocations = 0;
    if (c->n_slots) {
        R3_CUDA(c, cudaMemsetAsync(c->d_stats + 4, 0, 16, c->stream));
        max_invocations_kernel<<<R3_SM_COUNT * 4, 256, 0, c->stream>>>(c->d_objects, c->n_slots, c->d_stats + 4);
        R3_CHECK_LAUNCH(c, "max_invocations_kernel");
        unsigned long long v[2] = {0, 0};
        R3_CUDA(c, cudaMemcpyAsync(v, c->d_stats + 4, 16, cudaMemcpyDeviceToHost, c->stream));
        R3_CUDA(c, r3_stream_sync(c));
        c->max_total_invocations = v[0]; c->max_object_invocations = v[1];
    }
    c->max_invocations_valid = true;
    return R3_OK;
}

// key generation + stable LSD radix sort of `cap` candidates: the entries of `visible` (count on the device) or, with visible == nullptr,
// the slots [0, cap) themselves.  keys[*src_out] holds the result.
static int r3_launch_sort(r3_ctx* c, const uint32_t* visible, const uint32_t* visible_count, uint32_t cap, const float vp_loc[3], unsigned long long* keys[2],
                          uint32_t** hist, uint64_t* hist_cap, uint32_t* header, int* src_out) {
    int src = 0;
    const uint32_t sort_blocks = (cap + SORT_TILE - 1) / SORT_TILE;
    R3_TRY(r3_reserve_t(c, hist, hist_cap, (uint64_t)sort_blocks * 256 + 1));
    r3_stage_begin(c, R3_STAGE_SORT);
    if (cap <= SMALL_SORT_MAX) {
        // small worlds: key generation + radix sort by one CTA in shared memory, one launch
        const size_t smem = small_sort_smem(cap);
        cudaFuncSetAttribute(small_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_sort_smem(SMALL_SORT_MAX));   // per device
        small_sort_kernel<<<1, SMALL_SORT_THREADS, smem, c->stream>>>(visible, visible_count, c->d_sort_key8, c->d_sort_loc, vp_loc[0], vp_loc[1], vp_loc[2], keys[0], header,
                                                                      small_sort_pad(cap), cap);
        R3_CHECK_LAUNCH(c, "small_sort_kernel");
    } else if (sort_blocks <= COOP_SORT_MAX_BLOCKS && c->coop_launch_ok) {
        SortCoopParams sp;
        sp.visible = visible; sp.visible_count = visible_count; sp.key8 = c->d_sort_key8; sp.loc = c->d_sort_loc;
        sp.vx = vp_loc[0]; sp.vy = vp_loc[1]; sp.vz = vp_loc[2];
        sp.keys[0] = keys[0]; sp.keys[1] = keys[1]; sp.hist = *hist; sp.header = header; sp.count_imm = cap;
        void* args[] = {&sp};
        R3_CUDA(c, cudaLaunchCooperativeKernel((const void*)radix_sort_coop_kernel, dim3(sort_blocks), dim3(SORT_THREADS), args, 0, c->stream));
        c->launches++;
        src = SORT_PASSES & 1;
    } else {
        keygen_kernel<<<(cap + 255) / 256, 256, 0, c->stream>>>(visible, visible_count, c->d_sort_key8, c->d_sort_loc, vp_loc[0], vp_loc[1], vp_loc[2], keys[0], header, cap);
        R3_CHECK_LAUNCH(c, "keygen_kernel");
        for (int pass = 0; pass < SORT_PASSES; ++pass) {
            const int shift = KEY_SHIFT0 + 8 * pass;
            radix_hist_kernel<<<sort_blocks, SORT_THREADS, 0, c->stream>>>(keys[src], header, shift, *hist);
            R3_CHECK_LAUNCH(c, "radix_hist_kernel");
            scan_u32_kernel<<<1, 1024, 0, c->stream>>>(*hist, sort_blocks * 256u);
            R3_CHECK_LAUNCH(c, "scan_u32_kernel");
            radix_scatter_kernel<<<sort_blocks, SORT_THREADS, 0, c->stream>>>(keys[src], keys[src ^ 1], header, shift, *hist);
            R3_CHECK_LAUNCH(c, "radix_scatter_kernel");
            src ^= 1;
        }
    }
    r3_stage_end(c);
    *src_out = src;
    return R3_OK;
}

// Should this camera take its order from the frame-wide sort?  Yes when several cameras batch per frame (known from the previous
// frame; on a first frame: when shadow-casting lights exist) and the world is small enough that sorting every slot once beats
// sorting each camera's visible set.  R3_FRAME_SORT=0 / 1 forces the choice (tests run both).
static bool r3_frame_sort_wanted(r3_ctx* c, r3_camera* cam, const float vp_loc[3]) {
    // a camera batching a second time or another viewport location starts a new frame epoch (as do r3_set_frame_uniforms and new sort data)
    if (cam->gsort_epoch_used == c->gsort_epoch || (c->gsort_cameras_this_epoch > 0 && memcmp(c->gsort_loc, vp_loc, 12) != 0)) r3_new_frame_epoch(c);
    if (c->gsort_cameras_this_epoch == 0) memcpy(c->gsort_loc, vp_loc, 12);
    cam->gsort_epoch_used = c->gsort_epoch;
    c->gsort_cameras_this_epoch++;
    const char* force = getenv("R3_FRAME_SORT");
    if (force) return force[0] != '0';
    if (c->n_slots > (1u << 22)) return false;
    return c->gsort_cameras_last_epoch >= 2 || (c->gsort_cameras_last_epoch == 0 && c->n_dir >= 1);
}
static int r3_frame_sort(r3_ctx* c, const float vp_loc[3]) {
    if (c->gsort_valid && c->gsort_sorted_epoch == c->gsort_epoch) return R3_OK;   // the location is fixed within an epoch
    const uint32_t n = (uint32_t)(c->sort_flags.size() < c->n_slots ? c->sort_flags.size() : c->n_slots);
    R3_TRY(r3_reserve_t(c, &c->d_gsort_keys[0], &c->gsort_cap[0], (uint64_t)n + 1));
    R3_TRY(r3_reserve_t(c, &c->d_gsort_keys[1], &c->gsort_cap[1], (uint64_t)n + 1));
    if (!c->d_gsort_header) R3_CUDA(c, cudaMalloc((void**)&c->d_gsort_header, 32));
    int src = 0;
    if (n) R3_TRY(r3_launch_sort(c, nullptr, nullptr, n, vp_loc, c->d_gsort_keys, &c->d_gsort_hist, &c->gsort_hist_cap, c->d_gsort_header, &src));
    c->gsort_src = src; c->gsort_n = n; c->gsort_valid = true; c->gsort_sorted_epoch = c->gsort_epoch;
    return R3_OK;
}

int r3_device_batch_objects(r3_ctx* c, r3_camera* cam, const float vp_loc[3], uint32_t max_dispatch_count) {
    if (!cam->header_set) return r3_fail(c, R3_E_STATE, "batch_objects before object_uniform_upload");
    const uint32_t cap = cam->header.object_count;
    R3_TRY(r3_compute_max_invocations(c));
    if (c->max_total_invocations >= (1ull << 31)) return r3_fail(c, R3_E_INVALID, "more than 2^31 padded invocations");
    const int w = (cam->cache_idx == 0) ? 1 : 0;   // never overwrite the DrawCallSet cached for the predicted pass
    cam->cur = w;
    r3_jobs& j = cam->jobs[w];
    const uint32_t nb_cap = (cap + 255u) / 256u + 1u, nr_cap = nb_cap + 64u;
    R3_TRY(r3_reserve_t(c, &j.d_batches, &j.batches_cap, nb_cap));
    uint32_t rcap = j.regions_cap;
    R3_TRY(r3_reserve_t(c, &j.d_regions, &j.regions_cap, nr_cap));
    if (!j.d_region_first_inv || rcap != j.regions_cap) {
        cudaFree(j.d_region_first_inv);
        j.d_region_first_inv = nullptr;
        R3_CUDA(c, cudaMalloc((void**)&j.d_region_first_inv, ((size_t)j.regions_cap + 2) * 4));
    }
    if (!j.d_header) R3_CUDA(c, cudaMalloc((void**)&j.d_header, 32));
    R3_TRY(r3_reserve_t(c, &cam->d_sort_keys[0], &cam->sort_keys_cap, (uint64_t)cap + 1));
    R3_TRY(r3_reserve_t(c, &cam->d_sort_keys[1], &cam->sort_keys_cap2, (uint64_t)cap + 1));
    const uint32_t sort_blocks = (cap + SORT_TILE - 1) / SORT_TILE;
    R3_TRY(r3_reserve_t(c, &cam->d_sort_hist, &cam->sort_hist_cap, (uint64_t)sort_blocks * 256 + 1));
    // batch scratch: batch_inv[nb] | batch_regions[nb] | region_key[nb*256] | region_start[nb*256]
    R3_TRY(r3_reserve_t(c, &cam->d_batch_tmp, &cam->batch_tmp_cap, (uint64_t)nb_cap * (2 + 512)));
    if (cam->prev_inv_cap < cap || !cam->d_prev_inv[0]) {
        for (int k = 0; k < 2; ++k) {
            uint32_t* n = nullptr;
            R3_CUDA(c, cudaMalloc((void**)&n, ((size_t)cap + 1) * 4));
            R3_CUDA(c, cudaMemsetAsync(n, 0xFF, ((size_t)cap + 1) * 4, c->stream));
            if (cam->d_prev_inv[k]) {   // keep last frame's entries across a capacity growth
                R3_CUDA(c, cudaMemcpyAsync(n, cam->d_prev_inv[k], (size_t)cam->prev_inv_cap * 4, cudaMemcpyDeviceToDevice, c->stream));
                R3_CUDA(c, r3_stream_sync(c));
                cudaFree(cam->d_prev_inv[k]);
            }
            cam->d_prev_inv[k] = n;
        }
        cam->prev_inv_cap = cap;
    }
    const int prev = cam->prev_inv_cur, cur = prev ^ 1;
    // get_and_reset_camera (batching.rs:111-113): the WHOLE map starts empty, also the entries beyond this frame's object_count
    // (a world that shrinks and grows again must not see invocations from two frames ago)
    R3_CUDA(c, cudaMemsetAsync(cam->d_prev_inv[cur], 0xFF, (size_t)cam->prev_inv_cap * 4, c->stream));
    R3_CUDA(c, cudaMemsetAsync(j.d_header, 0, 32, c->stream));

    if (cap) {
        int src = 0;
        const unsigned long long* sorted_keys = nullptr;
        bool keys_hold_slots = false;
        if (r3_frame_sort_wanted(c, cam, vp_loc)) {
            // one sort for the frame's cameras, then this camera's visible objects in that order
            R3_TRY(r3_frame_sort(c, vp_loc));
            const uint32_t n = c->gsort_n, tiles = (n + RC_THREADS - 1) / RC_THREADS;
            R3_TRY(r3_reserve_t(c, &cam->d_sort_hist, &cam->sort_hist_cap, (uint64_t)tiles + 1));
            if (tiles) {                                   // no sortable slot at all: the zeroed header already says "no visible objects"
                rank_count_kernel<<<tiles, RC_THREADS, 0, c->stream>>>(c->d_gsort_keys[c->gsort_src], n, cam->d_words, cap, cam->d_sort_hist);
                R3_CHECK_LAUNCH(c, "rank_count_kernel");
                rank_scatter_kernel<<<tiles, RC_THREADS, 0, c->stream>>>(c->d_gsort_keys[c->gsort_src], n, cam->d_words, cap, cam->d_sort_hist, cam->d_sort_keys[0], j.d_header);
                R3_CHECK_LAUNCH(c, "rank_scatter_kernel");
            }
            sorted_keys = cam->d_sort_keys[0];
            keys_hold_slots = true;
            cam->batching_path = 3;
        } else {
            R3_TRY(r3_launch_sort(c, cam->d_visible, cam->d_visible_count, cap, vp_loc, cam->d_sort_keys, &cam->d_sort_hist, &cam->sort_hist_cap, j.d_header, &src));
            sorted_keys = cam->d_sort_keys[src];
        }
        BuildParams p;
        p.keys = sorted_keys; p.visible = cam->d_visible; p.objects = c->d_objects; p.keys_hold_slots = keys_hold_slots ? 1u : 0u;
        p.batches = j.d_batches; p.header = j.d_header;
        p.batch_inv = cam->d_batch_tmp; p.batch_regions = p.batch_inv + nb_cap; p.region_key = p.batch_regions + nb_cap; p.region_start = p.region_key + (size_t)nb_cap * 256;
        p.regions = j.d_regions; p.region_first_inv = j.d_region_first_inv;
        p.prev_map = cam->d_prev_inv[prev]; p.cur_map = cam->d_prev_inv[cur]; p.map_cap = cam->prev_inv_cap;
        p.n_batches_cap = nb_cap; p.dispatch_limit = (uint64_t)max_dispatch_count * R3_WORKGROUP_SIZE;
        batch_build_kernel<<<nb_cap - 1, 256, 0, c->stream>>>(p);
        R3_CHECK_LAUNCH(c, "batch_build_kernel");
        batch_scan_kernel<<<1, 1024, 0, c->stream>>>(p);
        R3_CHECK_LAUNCH(c, "batch_scan_kernel");
        batch_finalize_kernel<<<nb_cap - 1, 256, 0, c->stream>>>(p);
        R3_CHECK_LAUNCH(c, "batch_finalize_kernel");
    }
    cam->prev_inv_cur = cur;
    j.device_built = true; j.valid = true;
    j.n_batches = nb_cap - 1; j.n_regions = nr_cap - 1;                 // upper bounds; exact counts live in d_header
    j.total_invocations = (uint32_t)c->max_total_invocations;
    j.batches.clear(); j.regions.clear();
    return R3_OK;
}

// device-built jobs -> host vectors (r3_batch_counts / r3_readback_batches / tests); blocks
int r3_download_jobs(r3_ctx* c, r3_camera* cam) {
    r3_jobs& j = cam->jobs[cam->cur];
    if (!j.device_built || !j.batches.empty() || !j.d_header) return R3_OK;
    uint32_t hdr[8] = {0};
    R3_CUDA(c, cudaMemcpyAsync(hdr, j.d_header, 32, cudaMemcpyDeviceToHost, c->stream));
    R3_CUDA(c, r3_stream_sync(c));
    if (hdr[4]) return r3_fail(c, R3_E_INVALID, "device batch_objects overflow (batch beyond the dispatch limit or > 2^24 visible objects): use host batching");
    j.batches.resize(hdr[1]); j.regions.resize(hdr[2]);
    if (hdr[1]) R3_CUDA(c, cudaMemcpyAsync(j.batches.data(), j.d_batches, (size_t)hdr[1] * sizeof(r3_batch_data), cudaMemcpyDeviceToHost, c->stream));
    if (hdr[2]) R3_CUDA(c, cudaMemcpyAsync(j.regions.data(), j.d_regions, (size_t)hdr[2] * sizeof(r3_region), cudaMemcpyDeviceToHost, c->stream));
    R3_CUDA(c, r3_stream_sync(c));
    // the 8448-byte records carry 244 bytes of padding the host path leaves zero
    for (auto& b : j.batches) {
        std::memset(b._pad, 0, sizeof b._pad);
        for (uint32_t o = b.total_objects; o < R3_BATCH_SIZE; ++o) std::memset(&b.object_culling_information[o], 0, sizeof(r3_object_culling_info));
    }
    return R3_OK;
}
