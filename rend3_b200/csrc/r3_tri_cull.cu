// r3_tri_cull.cu — per-triangle cull + ORDERED index compaction.
//
// Replaces GpuCuller::cull + cull.wgsl::cs_main (rend3-routine/src/culling/culler.rs:531-659,
// rend3-routine/shaders/src/cull.wgsl:264-390): back-face determinant, misses-pixel-centre, hi-Z occlusion,
// predicted/residual index lists, indirect draw records and the per-invocation visibility bits, written into
// the same ping-pong CullingBuffers layout (culler.rs:88-125, suballoc.rs) the forward stage reads.
//
// B200 design.  The reference launches one dispatch per 256-object batch and appends survivors with a global
// atomicAdd per triangle on a handful of contended counters, which also makes the list order nondeterministic.
// Here every batch is handled by the same launches, all of them plain streaming kernels without inter-CTA waits:
//   test     persistent grid sized by the occupancy API; the reference workgroups (256 invocations) are dealt to WARPS
//            (warp g takes workgroups g, g + warps, ...), each warp stages the 400-byte run of indices of its next word
//            with cp.async.bulk + an mbarrier while it tests the current one; the warp ballot IS the 32-bit visibility
//            word the reference assembles with workgroup atomics (1 bit per invocation to HBM), a second word marks
//            the residual triangles; per-superblock (1024 words) survivor counts by one 64-bit RED per word.
//            With r3_set_cull_shard the viewport's workgroups are split in runs across ranks and the words travel
//            over NVLink peer stores (r3_peer.cu);
//   scan     one block: exclusive prefix of the superblock counts;
//   regions  one CTA per region: survivors in front of the region -> its two IndirectCall records;
//   compact  one CTA per superblock: block scan of the word popcounts, then each surviving triangle re-reads its three
//            indices (12 B) and lands at  region base + (survivors before it in the region): ascending invocation order,
//            one legal outcome of the reference's atomics, reproducible run to run.
// (Round-1 history: a single-pass version with a per-workgroup decoupled look-back measured 1.66 ms per camera on the
//  200k-object config — ticket + look-back latency per 256 invocations; see profiles/README.md.)
#include <cstdlib>

#include "r3_common.cuh"

namespace {

constexpr int TC_THREADS = 256;                      // = the reference's workgroup.  (Quarter-workgroup CTAs, so that all-padding quarters retire early,
                                                     //  were measured 34% SLOWER on config 3: four times the CTAs to dispatch outweighs the occupancy they free;
                                                     //  one warp per word that holds work, from a compacted word list, measured no gain either: the test is bound
                                                     //  by the sector traffic of its index / position gathers, not by the idle padding warps; fetching each 12-byte
                                                     //  triple with one or two aligned 16-byte loads instead of three 4-byte ones made config 3 14% slower.)
constexpr int SB_WORDS = 1024;                       // words per superblock (32768 invocations, 128 workgroups)

struct TriCullParams {
    r3_camera_header cam;
    const uint32_t* mesh; uint64_t mesh_words, mesh_cap_words;   // mesh_cap_words: words of the allocation (>= mesh_words + 4): bulk copies may over-read into the slack
    const r3_object* objects;
    const r3_object_matrices* matrices;
    const r3_batch_data* batches;
    const uint4* wg_info;                // per workgroup: WgRec (two uint4), everything the test needs about its object
    const uint32_t* region_first_inv;    // [n_regions + 1]
    const uint32_t* header;              // job header: [1] n_batches, [2] n_regions, [3] total_invocations (device-side counts)
    uint32_t* idx_pred; uint32_t* idx_resid;
    r3_indirect_call* dc_pred; r3_indirect_call* dc_resid;
    uint32_t* res_out; const uint32_t* res_in; uint64_t res_in_words;
    uint32_t* resid_bits;                // [n_words] residual marks (scratch)
    unsigned long long* sb_counts;       // [n_superblocks + 1] (pred << 32 | resid): counts, then their exclusive prefix
    unsigned long long* region_prefix;   // [n_regions + 1] survivors in front of each region
    float* const* hiz; const uint32_t* hiz_dims; uint32_t hiz_mips;
    // multi-GPU ("triangle cull: shard by batch"): this launch tests the workgroups [n_wg * shard_index / shard_count,
    // n_wg * (shard_index + 1) / shard_count) — workgroups are laid out batch after batch, so a shard is a run of batches — and stores
    // its visibility words into the staging arrays [pred words | resid words] of EVERY rank (peer memory) instead of the local buffers
    uint32_t shard_index, shard_count, ex_n, ex_cap_words;
    uint32_t* ex_peers[R3_MAX_EXCHANGE_RANKS];
};

__device__ __forceinline__ uint32_t mesh_word(const TriCullParams& p, uint64_t i) { return i < p.mesh_words ? __ldg(&p.mesh[i]) : 0u; }

// ceil(log2(max(x,1))) evaluated exactly on the f32 bit pattern (cull.wgsl:314)
__device__ __forceinline__ uint32_t ceil_log2_f32(float x) {
    if (!(x > 1.0f)) return 0u;
    const uint32_t b = __float_as_uint(x), e = (b >> 23) & 0xFFu, m = b & 0x7FFFFFu;
    if (e == 255u) return 128u;
    return (e - 127u) + (m ? 1u : 0u);
}

// textureSampleMin (cull.wgsl:243-262); out-of-range mip / texel coordinates clamp (robust access)
__device__ float hiz_sample_min(const TriCullParams& p, float u, float v, uint32_t mip) {
    if (p.hiz_mips == 0u) return 0.0f;
    if (mip >= p.hiz_mips) mip = p.hiz_mips - 1u;
    const uint32_t w = p.hiz_dims[2 * mip], h = p.hiz_dims[2 * mip + 1];
    const float rw = (float)w, rh = (float)h;
    const float px = sub_rn(mul_rn(u, rw), 0.5f), py = sub_rn(mul_rn(v, rh), 0.5f);
    float lx = fmaxf(floorf(px), 0.0f), ly = fmaxf(floorf(py), 0.0f);
    float hx = fminf(ceilf(px), rw - 1.0f), hy = fminf(ceilf(py), rh - 1.0f);
    lx = fminf(lx, rw - 1.0f); ly = fminf(ly, rh - 1.0f); hx = fmaxf(hx, 0.0f); hy = fmaxf(hy, 0.0f);
    if (!(lx == lx)) lx = 0.f; if (!(ly == ly)) ly = 0.f; if (!(hx == hx)) hx = 0.f; if (!(hy == hy)) hy = 0.f;
    const uint32_t x0 = (uint32_t)lx, y0 = (uint32_t)ly, x1 = (uint32_t)hx, y1 = (uint32_t)hy;
    const float* t = p.hiz[mip];
    float m = t[(size_t)y0 * w + x0];
    m = fminf(m, t[(size_t)y0 * w + x1]);
    m = fminf(m, t[(size_t)y1 * w + x0]);
    m = fminf(m, t[(size_t)y1 * w + x1]);
    return m;
}

// execute_culling (cull.wgsl:264-324), IEEE f32, source order, no FMA
// clip-space x, y, w of M * (v, 1) (mat_point_rn's operation order per component); z is only needed by the hi-Z test of the survivors
__device__ __forceinline__ float3 mat_point_xyw_rn(const float* __restrict__ m, const float3 v) {
    float3 r;
    r.x = add_rn(add_rn(add_rn(mul_rn(m[0], v.x), mul_rn(m[4], v.y)), mul_rn(m[8], v.z)), m[12]);
    r.y = add_rn(add_rn(add_rn(mul_rn(m[1], v.x), mul_rn(m[5], v.y)), mul_rn(m[9], v.z)), m[13]);
    r.z = add_rn(add_rn(add_rn(mul_rn(m[3], v.x), mul_rn(m[7], v.y)), mul_rn(m[11], v.z)), m[15]);   // .z of the result holds clip w
    return r;
}
__device__ __forceinline__ float mat_point_z_rn(const float* __restrict__ m, const float3 v) {
    return add_rn(add_rn(add_rn(mul_rn(m[2], v.x), mul_rn(m[6], v.y)), mul_rn(m[10], v.z)), m[14]);
}
__device__ bool execute_culling(const TriCullParams& p, const float* __restrict__ mvp, const float3 a, const float3 b, const float3 c) {
    struct { float x, y, w; } p0, p1, p2;
    { const float3 q = mat_point_xyw_rn(mvp, a); p0.x = q.x; p0.y = q.y; p0.w = q.z; }
    { const float3 q = mat_point_xyw_rn(mvp, b); p1.x = q.x; p1.y = q.y; p1.w = q.z; }
    { const float3 q = mat_point_xyw_rn(mvp, c); p2.x = q.x; p2.y = q.y; p2.w = q.z; }
    const float t0 = sub_rn(mul_rn(p1.y, p2.w), mul_rn(p2.y, p1.w));
    const float t1 = sub_rn(mul_rn(p0.y, p2.w), mul_rn(p2.y, p0.w));
    const float t2 = sub_rn(mul_rn(p0.y, p1.w), mul_rn(p1.y, p0.w));
    const float det = add_rn(sub_rn(mul_rn(p0.x, t0), mul_rn(p1.x, t1)), mul_rn(p2.x, t2));
    const bool positive = p.cam.flags & R3_PCU_POSITIVE_AREA_VISIBLE;
    if (positive && det <= 0.0f) return false;
    if (!positive && det >= 0.0f) return false;
    float n0x = p0.x, n0y = p0.y, n1x = p1.x, n1y = p1.y, n2x = p2.x, n2y = p2.y;
    const bool unit_w = p0.w == 1.0f && p1.w == 1.0f && p2.w == 1.0f;   // x / 1.0f == x: orthographic (shadow) cameras skip the perspective divide
    if (!unit_w) {
        n0x = div_rn(p0.x, p0.w); n0y = div_rn(p0.y, p0.w);
        n1x = div_rn(p1.x, p1.w); n1y = div_rn(p1.y, p1.w);
        n2x = div_rn(p2.x, p2.w); n2y = div_rn(p2.y, p2.w);
    }
    const float minx = fminf(n0x, fminf(n1x, n2x)), miny = fminf(n0y, fminf(n1y, n2y));
    const float maxx = fmaxf(n0x, fmaxf(n1x, n2x)), maxy = fmaxf(n0y, fmaxf(n1y, n2y));
    const float hrx = mul_rn(p.cam.resolution[0], 0.5f), hry = mul_rn(p.cam.resolution[1], 0.5f);   // x / 2 == x * 0.5 for every float
    const float minsx = mul_rn(add_rn(minx, 1.0f), hrx), minsy = mul_rn(add_rn(miny, 1.0f), hry);
    const float maxsx = mul_rn(add_rn(maxx, 1.0f), hrx), maxsy = mul_rn(add_rn(maxy, 1.0f), hry);
    if (!(p.cam.flags & R3_PCU_MULTISAMPLED)) {
        if (rintf(minsx) == rintf(maxsx) || rintf(minsy) == rintf(maxsy)) return false;   // WGSL round(): ties to even
    }
    if (p.cam.shadow_index != R3_CAMERA_VIEWPORT) return true;
    const float mintx = mul_rn(add_rn(minx, 1.0f), 0.5f), minty = sub_rn(1.0f, mul_rn(add_rn(miny, 1.0f), 0.5f));
    const float maxtx = mul_rn(add_rn(maxx, 1.0f), 0.5f), maxty = sub_rn(1.0f, mul_rn(add_rn(maxy, 1.0f), 0.5f));
    const float u = mul_rn(add_rn(maxtx, mintx), 0.5f), v = mul_rn(add_rn(maxty, minty), 0.5f);
    const float ex = sub_rn(maxsx, minsx), ey = sub_rn(maxsy, minsy);
    const uint32_t mip = ceil_log2_f32(fmaxf(fmaxf(ex, ey), 1.0f));
    // the depths: only the triangles that reach the occlusion test pay for the z row of the transform and its three divisions
    float n0z = mat_point_z_rn(mvp, a), n1z = mat_point_z_rn(mvp, b), n2z = mat_point_z_rn(mvp, c);
    if (!unit_w) { n0z = div_rn(n0z, p0.w); n1z = div_rn(n1z, p1.w); n2z = div_rn(n2z, p2.w); }
    const float depth = fmaxf(fmaxf(n0z, n1z), n2z);
    const float occl = hiz_sample_min(p, u, v, mip);
    return !(depth < occl);
}

// One record per 256-invocation workgroup, written once per frame by expand_wg_info_kernel: the test kernel would otherwise
// walk workgroup -> batch -> object table -> Object record before it can fetch a single index (four dependent misses for
// 1 M workgroups per frame on the 200k-object config).
struct WgRec {
    uint32_t object_id, first_index, pos_off, n_real;             // n_real: invocations of this workgroup that are triangles (rest = 256-padding)
    uint32_t prev_invocation, flags, object_invocation, region;   // flags: bit0 atomic_capable | batch-local object << 8
};
static_assert(sizeof(WgRec) == 32, "WgRec");
constexpr uint32_t WG_ATOMIC = 1u;

__global__ void expand_wg_info_kernel(const r3_batch_data* __restrict__ batches, const uint32_t* __restrict__ header, const r3_object* __restrict__ objects,
                                      uint4* __restrict__ wg_info) {
    const uint32_t b = blockIdx.x, o = threadIdx.x;
    if (b >= header[1]) return;
    const r3_batch_data* job = &batches[b];
    if (o >= job->total_objects) return;
    const r3_object_culling_info info = job->object_culling_information[o];                 // find_object_info (cull.wgsl:181-207), hoisted
    const uint32_t n = info.invocation_end - info.invocation_start;
    const uint32_t first = (job->batch_base_invocation + info.invocation_start) >> 8, count = (n + 255u) >> 8;
    const r3_object* obj = &objects[info.object_id];
    const uint32_t first_index = obj->first_index, pos_off = obj->attr_offset[0] >> 2;
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t done = i << 8;
        wg_info[2 * (size_t)(first + i)] = make_uint4(info.object_id, first_index, pos_off, min(n - done, 256u));
        wg_info[2 * (size_t)(first + i) + 1] = make_uint4(info.previous_global_invocation, (info.atomic_capable == 1u ? WG_ATOMIC : 0u) | (o << 8), done, info.region_id);
    }
}
__device__ __forceinline__ WgRec load_wg(const TriCullParams& p, uint32_t wg) {
    const uint4 a = __ldg(&p.wg_info[2 * (size_t)wg]), b = __ldg(&p.wg_info[2 * (size_t)wg + 1]);
    WgRec r;
    r.object_id = a.x; r.first_index = a.y; r.pos_off = a.z; r.n_real = a.w;
    r.prev_invocation = b.x; r.flags = b.y; r.object_invocation = b.z; r.region = b.w;
    return r;
}

// ---- test: cull.wgsl::cs_main up to the visibility decision
// Persistent grid (R3_SM_COUNT x 8 CTAs of 8 warps = every warp slot of the GPU), each CTA striding over the reference's workgroups;
// warp w of the CTA owns invocations [32 w, 32 w + 32) of the workgroup = one visibility word.  The warps never meet: no shared
// memory, no barrier (the round-2 profile of the barrier version, profiles/README.md: 62-70% issue-active with `barrier` the largest
// stall reason, 7 of 8 warps of a 12-triangle object idling through the block reduction, 2.08 M warps launched for 0.46 M with work).
//   * padding words (past the object's last triangle) cost one predicated store pair and the warp moves on;
//   * survivor counts go to the superblock counters with one 64-bit RED per word that has survivors (most have none);
//   * an orthographic camera (every shadow camera) produces w == 1.0f exactly, where x / w == x bit for bit: the nine IEEE divisions of
//     the perspective divide — a third of the instructions of the test — are skipped; x / 2.0f is evaluated as x * 0.5f (identical
//     for every float, subnormals included); the mesh-buffer bound is checked once per triple instead of once per word.
__device__ __forceinline__ float3 load_position(const TriCullParams& p, uint32_t pos_off, uint32_t vid) {
    const uint64_t f = (uint64_t)pos_off + (uint64_t)vid * 3u;                                 // extract_attribute_vec3_f32
    if (f + 2u < p.mesh_words) return make_float3(__uint_as_float(__ldg(&p.mesh[f])), __uint_as_float(__ldg(&p.mesh[f + 1])), __uint_as_float(__ldg(&p.mesh[f + 2])));
    return make_float3(__uint_as_float(mesh_word(p, f)), __uint_as_float(mesh_word(p, f + 1)), __uint_as_float(mesh_word(p, f + 2)));   // robust access at the buffer end
}
// ---- bulk-async staging of the index pulls (vertex_fetch, cull.wgsl:9-32).  The 32 invocations of a word read 96 consecutive index
// words: one `cp.async.bulk` (the TMA unit's 1-D bulk copy, SASS UBLKCP) moves the run — widened to the enclosing 16-byte aligned 400
// bytes — into the warp's own shared-memory slot and completes on the warp's own mbarrier (SYNCS).  Two slots per warp: the copy for the
// NEXT workgroup is issued while the current one is being tested, so the index latency (and the three strided 4-byte loads per lane it
// replaces) leaves the dependent chain  workgroup record -> indices -> positions -> test.
constexpr uint32_t IDX_RUN_WORDS = 100;                        // 96 index words + up to 3 words of alignment slack, rounded to 16 bytes
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(TC_THREADS, MIN_CTAS) triangle_test_kernel(const __grid_constant__ TriCullParams p) {
    __shared__ __align__(16) uint32_t s_idx[TC_THREADS / 32][2][IDX_RUN_WORDS];
    __shared__ __align__(8) uint64_t s_bar[TC_THREADS / 32][2];
    const uint32_t n_wg = __ldg(&p.header[3]) / TC_THREADS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool shadow = p.cam.shadow_index != R3_CAMERA_VIEWPORT;
    // Work is dealt out per WARP, not per CTA: warp g of the grid takes the workgroups g, g + G, g + 2 G, ... and walks the words of each
    // that hold triangles (1..8) one after the other.  (Dealing a workgroup to a CTA leaves warp 0 busy in every workgroup and warp 7 in
    // few: measured 2.4 ms instead of 2.1 ms per config-3 frame — half the resident warps idle behind the busy ones.)
    const uint32_t total_warps = gridDim.x * (TC_THREADS / 32);
    const uint32_t wg_lo = p.shard_count > 1u ? (uint32_t)((uint64_t)n_wg * p.shard_index / p.shard_count) : 0u;
    const uint32_t wg_end = p.shard_count > 1u ? (uint32_t)((uint64_t)n_wg * (p.shard_index + 1u) / p.shard_count) : n_wg;
    uint32_t wg = wg_lo + blockIdx.x * (TC_THREADS / 32) + warp;
    if (wg >= wg_end) return;
    // one visibility word pair: local buffers + superblock counter, or (sharded) the staging arrays of every rank
    const auto store_words = [&](uint32_t word, uint32_t pred, uint32_t resid) {
        if (p.ex_n) {
#pragma unroll 1
            for (uint32_t r = 0; r < p.ex_n; ++r) { p.ex_peers[r][word] = pred; p.ex_peers[r][p.ex_cap_words + word] = resid; }
        } else {
            p.res_out[word] = pred;                                                            // save_culling_results (cull.wgsl:229-241)
            p.resid_bits[word] = resid;
            // every visibility word is counted (atomic or not) so that the prefix over words is one consistent global scan
            const unsigned long long t = ((unsigned long long)__popc(pred) << 32) | __popc(resid);
            if (t) atomicAdd(&p.sb_counts[word / SB_WORDS], t);
        }
    };
    if (lane == 0) { mbar_init(&s_bar[warp][0], 1u); mbar_init(&s_bar[warp][1], 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    uint32_t uses[2] = {0u, 0u};                                  // completed copies per slot: the mbarrier phase parity
    // first index word of word `w` of the workgroup `r`, or ~0 when the run cannot be staged (it would leave the mesh allocation)
    const auto run_start = [&](const WgRec& r, uint32_t w) -> uint64_t {
        const uint64_t s = (uint64_t)r.first_index + ((uint64_t)r.object_invocation + w * 32u) * 3u;
        return ((s & ~3ull) + IDX_RUN_WORDS <= p.mesh_cap_words) ? s : ~0ull;
    };
    const auto stage_run = [&](uint64_t s, int slot) {             // one elected lane: arm the barrier, start the copy
        if (lane == 0 && s != ~0ull) {
            mbar_expect_tx(&s_bar[warp][slot], IDX_RUN_WORDS * 4u);
            bulk_copy_g2s(&s_idx[warp][slot][0], p.mesh + (s & ~3ull), IDX_RUN_WORDS * 4u, &s_bar[warp][slot]);
        }
    };
    // the words of a workgroup past its last triangle: zero visibility words; non-atomic objects also need their slots INVALID
    const auto pad_words = [&](const WgRec& r, uint32_t g, uint32_t nw) {
        if ((uint32_t)lane >= nw && lane < TC_THREADS / 32) store_words(g * (TC_THREADS / 32) + lane, 0u, 0u);
        if (!(r.flags & WG_ATOMIC))
            for (uint32_t pw = nw; pw < TC_THREADS / 32; ++pw) {
                const uint64_t o = (((uint64_t)g * (TC_THREADS / 32) + pw) * 32u + lane) * 3u;
                p.idx_resid[o] = R3_INVALID_VERTEX; p.idx_resid[o + 1] = R3_INVALID_VERTEX; p.idx_resid[o + 2] = R3_INVALID_VERTEX;
            }
    };
    // of the NEXT workgroup only the start of its first index run is kept (its record is read again, from L1 / L2, when the warp gets there)
    const auto first_run = [&](uint32_t g) -> uint64_t {
        const uint4 a = __ldg(&p.wg_info[2 * (size_t)g]), b = __ldg(&p.wg_info[2 * (size_t)g + 1]);
        const uint64_t s = (uint64_t)a.y + (uint64_t)b.z * 3u;
        return ((s & ~3ull) + IDX_RUN_WORDS <= p.mesh_cap_words) ? s : ~0ull;
    };
    WgRec rec = load_wg(p, wg);
    uint32_t wg_n = wg + total_warps;
    uint64_t run_n0 = wg_n < wg_end ? first_run(wg_n) : ~0ull;
    uint32_t w = 0, nw = (rec.n_real + 31u) >> 5;
    pad_words(rec, wg, nw);
    uint64_t run = run_start(rec, 0u);
    int slot = 0;
    stage_run(run, slot);
    for (;;) {
        // the word after this one: the next word of this workgroup, or the first word of this warp's next workgroup
        const bool same_wg = w + 1u < nw;
        const bool has_next = same_wg || wg_n < wg_end;
        const uint64_t run_next = has_next ? (same_wg ? run_start(rec, w + 1u) : run_n0) : ~0ull;

        const uint32_t word = wg * (TC_THREADS / 32) + w;          // global_invocation >> 5
        const uint32_t local = w * 32u + lane;                     // invocation inside the workgroup
        const bool atomic_capable = rec.flags & WG_ATOMIC;
        const bool real = local < rec.n_real;
        const uint32_t object_invocation = rec.object_invocation + local;
        const uint64_t ib = (uint64_t)rec.first_index + (uint64_t)object_invocation * 3u;   // vertex_fetch (cull.wgsl:9-32)
        bool passes = false, resid = false;
        uint32_t i0 = 0, i1 = 0, i2 = 0;
        if (run != ~0ull) {
            mbar_wait(&s_bar[warp][slot], uses[slot] & 1u);
            uses[slot]++;
            const uint32_t o = (uint32_t)(run & 3ull) + (uint32_t)lane * 3u;
            if (real) { i0 = s_idx[warp][slot][o]; i1 = s_idx[warp][slot][o + 1]; i2 = s_idx[warp][slot][o + 2]; }
            if (real && ib + 2u >= p.mesh_words) { i0 = mesh_word(p, ib); i1 = mesh_word(p, ib + 1); i2 = mesh_word(p, ib + 2); }   // robust access at the buffer end
        } else if (real) {
            i0 = mesh_word(p, ib); i1 = mesh_word(p, ib + 1); i2 = mesh_word(p, ib + 2);
        }
        float3 v0 = make_float3(0.f, 0.f, 0.f), v1 = v0, v2 = v0;
        if (real) { v0 = load_position(p, rec.pos_off, i0); v1 = load_position(p, rec.pos_off, i1); v2 = load_position(p, rec.pos_off, i2); }
        // the other slot was read one word ago by every lane (their position loads depended on it): refill it with the next word's run
        __syncwarp();
        stage_run(run_next, slot ^ 1);
        if (real) {
            float mvp[16];
            const float4* m4 = reinterpret_cast<const float4*>(p.matrices[rec.object_id].model_view_proj);   // 64-byte aligned: four 16-byte loads
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 c4 = __ldg(&m4[q]); mvp[4 * q] = c4.x; mvp[4 * q + 1] = c4.y; mvp[4 * q + 2] = c4.z; mvp[4 * q + 3] = c4.w; }
            passes = execute_culling(p, mvp, v0, v1, v2);
            if (passes && !shadow && atomic_capable) {
                bool prev = false;                                                             // get_previous_culling_result (cull.wgsl:152-160)
                if (rec.prev_invocation != R3_NO_PREVIOUS) {
                    const uint64_t pgi = (uint64_t)object_invocation + rec.prev_invocation;
                    const uint32_t mask = (pgi >> 5) < p.res_in_words ? p.res_in[pgi >> 5] : 0u;
                    prev = (mask >> (pgi & 31)) & 1u;
                }
                resid = !prev;
            }
        }
        const uint32_t word_pred = __ballot_sync(0xFFFFFFFFu, passes);
        const uint32_t word_resid = __ballot_sync(0xFFFFFFFFu, resid);
        if (lane == 0) store_words(word, word_pred, word_resid);
        if (!atomic_capable) {
            // non-atomic (blend) objects keep their slot: survivors in place, everything else INVALID (cull.wgsl:374-380,343-347)
            const uint64_t o = ((uint64_t)word * 32u + lane) * 3u;
            const uint32_t hi = (rec.flags >> 8) << 24;
            p.idx_resid[o] = passes ? (hi | (i0 & 0xFFFFFFu)) : R3_INVALID_VERTEX;
            p.idx_resid[o + 1] = passes ? (hi | (i1 & 0xFFFFFFu)) : R3_INVALID_VERTEX;
            p.idx_resid[o + 2] = passes ? (hi | (i2 & 0xFFFFFFu)) : R3_INVALID_VERTEX;
        }
        if (!has_next) break;
        if (same_wg) {
            w++;
        } else {
            wg = wg_n; rec = load_wg(p, wg); w = 0; nw = (rec.n_real + 31u) >> 5;
            wg_n += total_warps;
            run_n0 = wg_n < wg_end ? first_run(wg_n) : ~0ull;      // in flight while this workgroup is tested
            pad_words(rec, wg, nw);
        }
        run = run_next;
        slot ^= 1;
    }
}

// ---- sharded test: the words every rank stored into this rank's staging arrays -> the local CullingBuffers, plus the superblock counts the
// unsharded test accumulates with atomics.  One CTA per superblock (1024 words).
__global__ void __launch_bounds__(SB_WORDS) shard_unpack_kernel(const __grid_constant__ TriCullParams p) {
    __shared__ unsigned long long s_warp[32];
    const uint32_t n_words = p.header[3] / 32u;
    const uint32_t w = blockIdx.x * SB_WORDS + threadIdx.x;
    if (blockIdx.x * SB_WORDS >= n_words) return;
    const uint32_t* mine = p.ex_peers[p.shard_index];
    uint32_t pred = 0u, resid = 0u;
    if (w < n_words) { pred = __ldcg(&mine[w]); resid = __ldcg(&mine[p.ex_cap_words + w]); p.res_out[w] = pred; p.resid_bits[w] = resid; }   // written by peers: read from L2
    unsigned long long c = ((unsigned long long)__popc(pred) << 32) | __popc(resid);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, s);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 32; ++k) t += s_warp[k];
        p.sb_counts[blockIdx.x] = t;
    }
}

// ---- scan: exclusive prefix of the superblock counts (packed pair), one block
__global__ void __launch_bounds__(1024) superblock_scan_kernel(unsigned long long* __restrict__ counts, const uint32_t* __restrict__ header) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t n = (header[3] / 32u + SB_WORDS - 1) / SB_WORDS;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base <= n; base += 1024) {   // <= n: entry n receives the grand total
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = s_warp[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += t; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[warp] + incl - v;
        if (i <= n) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

// survivors (pred << 32 | resid) in front of word `w`: superblock prefix + the words of the superblock before it
__device__ unsigned long long prefix_before_word(const TriCullParams& p, uint32_t w, unsigned long long* s_red) {
    const uint32_t sb = w / SB_WORDS, first = sb * SB_WORDS;
    unsigned long long acc = 0;
    for (uint32_t i = first + threadIdx.x; i < w; i += blockDim.x) acc += ((unsigned long long)__popc(p.res_out[i]) << 32) | __popc(p.resid_bits[i]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    unsigned long long tot = 0;
    for (uint32_t k = 0; k < blockDim.x / 32; ++k) tot += s_red[k];
    return p.sb_counts[sb] + tot;
}

// ---- regions: survivors in front of every region, and the IndirectCall records (init_draw_calls + final vertex_count)
__global__ void __launch_bounds__(256) region_finish_kernel(const __grid_constant__ TriCullParams p) {
    __shared__ unsigned long long s_red[8];
    const uint32_t r = blockIdx.x, n_regions = p.header[2];
    if (r >= n_regions) return;
    const uint32_t first_inv = p.region_first_inv[r], end_inv = p.region_first_inv[r + 1];
    const unsigned long long before = prefix_before_word(p, first_inv >> 5, s_red);
    const unsigned long long after = prefix_before_word(p, end_inv >> 5, s_red);
    if (threadIdx.x == 0) {
        p.region_prefix[r] = before;
        if (end_inv > first_inv) {
            // the region's first object decides atomic / non-atomic for the whole region (one material key per region)
            const bool atomic_region = load_wg(p, first_inv >> 8).flags & WG_ATOMIC;
            const unsigned long long cnt = after - before;
            r3_indirect_call pc, rc;
            pc.vertex_count = atomic_region ? 3u * (uint32_t)(cnt >> 32) : 0u;
            rc.vertex_count = atomic_region ? 3u * (uint32_t)(cnt & 0xFFFFFFFFull) : 3u * (end_inv - first_inv);
            pc.instance_count = rc.instance_count = 1u;
            pc.base_index = rc.base_index = first_inv * 3u;
            pc.vertex_offset = rc.vertex_offset = 0;
            pc.base_instance = rc.base_instance = 0u;
            p.dc_pred[r] = pc;
            p.dc_resid[r] = rc;
        }
    }
}

// ---- compact: ordered expansion of the surviving triangles (write_predicted/residual_atomic_triangle, cull.wgsl:84-116)
__global__ void __launch_bounds__(SB_WORDS) triangle_compact_kernel(const __grid_constant__ TriCullParams p) {
    __shared__ unsigned long long s_warp[32];
    const uint32_t n_words = p.header[3] / 32u;
    const uint32_t w = blockIdx.x * SB_WORDS + threadIdx.x;
    if (blockIdx.x * SB_WORDS >= n_words) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wp = w < n_words ? p.res_out[w] : 0u, wr = w < n_words ? p.resid_bits[w] : 0u;
    const unsigned long long c = ((unsigned long long)__popc(wp) << 32) | __popc(wr);
    unsigned long long incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long v = s_warp[lane];
        unsigned long long vi = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, vi, d); if (lane >= d) vi += t; }
        s_warp[lane] = vi - v;
    }
    __syncthreads();
    const unsigned long long excl = p.sb_counts[blockIdx.x] + s_warp[warp] + incl - c;   // survivors in front of word w (global)
    // per-word facts needed to re-read the indices and place them
    uint32_t base_pred = 0, base_resid = 0, idx_base = 0, hi = 0;
    bool atomic_word = false;
    if (wp | wr) {
        const WgRec rec = load_wg(p, w >> 3);
        atomic_word = rec.flags & WG_ATOMIC;
        const unsigned long long rp = p.region_prefix[rec.region];
        const uint32_t region_first = p.region_first_inv[rec.region];
        base_pred = region_first + (uint32_t)((excl - rp) >> 32);
        base_resid = region_first + (uint32_t)((excl & 0xFFFFFFFFull) - (rp & 0xFFFFFFFFull));
        // index of the first invocation of this word inside the mesh buffer
        idx_base = rec.first_index + (rec.object_invocation + (w & 7u) * 32u) * 3u;
        hi = (rec.flags >> 8) << 24;
    }
    // the 32 words of a warp are expanded one after the other, one triangle per lane
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, atomic_word && (wp | wr));
#pragma unroll 1
    for (uint32_t m = any; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const uint32_t jp = __shfl_sync(0xFFFFFFFFu, wp, j), jr = __shfl_sync(0xFFFFFFFFu, wr, j);
        const uint32_t jbp = __shfl_sync(0xFFFFFFFFu, base_pred, j), jbr = __shfl_sync(0xFFFFFFFFu, base_resid, j);
        const uint32_t jib = __shfl_sync(0xFFFFFFFFu, idx_base, j), jhi = __shfl_sync(0xFFFFFFFFu, hi, j);
        if ((jp >> lane) & 1u) {   // residual bits are a subset of the predicted bits
            const uint64_t ib = (uint64_t)jib + (uint64_t)lane * 3u;
            const uint32_t k0 = jhi | (mesh_word(p, ib) & 0xFFFFFFu), k1 = jhi | (mesh_word(p, ib + 1) & 0xFFFFFFu), k2 = jhi | (mesh_word(p, ib + 2) & 0xFFFFFFu);
            const uint32_t lt = (1u << lane) - 1u;
            const uint64_t sp = ((uint64_t)jbp + __popc(jp & lt)) * 3u;
            p.idx_pred[sp] = k0; p.idx_pred[sp + 1] = k1; p.idx_pred[sp + 2] = k2;
            if ((jr >> lane) & 1u) {
                const uint64_t sr = ((uint64_t)jbr + __popc(jr & lt)) * 3u;
                p.idx_resid[sr] = k0; p.idx_resid[sr + 1] = k1; p.idx_resid[sr + 2] = k2;
            }
        }
    }
}

}  // namespace

int r3_launch_triangle_cull(r3_ctx* c, r3_camera* cam) {
    r3_jobs& j = cam->jobs[cam->cur];
    const uint64_t inv = j.total_invocations, words = (inv + 31) / 32;   // exact (host batching) or an upper bound (device batching)
    if (!cam->index_buffer.created) {                      // CullingBuffers::new (culler.rs:96-112)
        R3_TRY(r3_iobuf_new(c, &cam->index_buffer, inv * 3, 4, false));
        R3_TRY(r3_iobuf_new(c, &cam->draw_call_buffer, j.n_regions, 20, true));
        R3_TRY(r3_iobuf_new(c, &cam->results_buffer, words, 4, false));
    } else {                                               // update_sizes (culler.rs:114-124)
        R3_TRY(r3_iobuf_swap(c, &cam->index_buffer, inv * 3));
        R3_TRY(r3_iobuf_swap(c, &cam->draw_call_buffer, j.n_regions));
        R3_TRY(r3_iobuf_swap(c, &cam->results_buffer, words));
    }
    R3_CUDA(c, cudaMemsetAsync(cam->draw_call_buffer.d, 0, cam->draw_call_buffer.capacity_elements * 20, c->stream));   // culler.rs:642
    const uint32_t n_wg = (uint32_t)(inv / TC_THREADS);
    cam->has_draw_call_set = true;
    if (n_wg == 0 || j.n_batches == 0 || j.n_regions == 0) return R3_OK;

    const uint32_t n_sb = (uint32_t)((words + SB_WORDS - 1) / SB_WORDS);
    // scratch: wg records [8 words x n_wg] | resid_bits [words]   and   sb_counts [n_sb + 1] | region_prefix [n_regions + 1]
    R3_TRY(r3_reserve_t(c, &cam->d_resid_bits, &cam->resid_bits_cap, (uint64_t)n_wg * 8 + words + 8));
    R3_TRY(r3_reserve_t(c, &cam->d_word_scan, &cam->word_scan_cap, (uint64_t)n_sb + j.n_regions + 4));
    uint4* wg_info = reinterpret_cast<uint4*>(cam->d_resid_bits);
    uint32_t* resid_bits = cam->d_resid_bits + (size_t)n_wg * 8;
    unsigned long long* sb_counts = cam->d_word_scan;
    unsigned long long* region_prefix = sb_counts + n_sb + 1;
    R3_CUDA(c, cudaMemsetAsync(sb_counts, 0, ((size_t)n_sb + 1) * 8, c->stream));
    expand_wg_info_kernel<<<j.n_batches, 256, 0, c->stream>>>(j.d_batches, j.d_header, c->d_objects, wg_info);
    R3_CHECK_LAUNCH(c, "expand_wg_info_kernel");

    TriCullParams p;
    p.cam = cam->header;
    p.mesh = c->d_mesh; p.mesh_words = c->mesh_words; p.mesh_cap_words = c->d_mesh ? c->mesh_cap : 0;
    p.objects = c->d_objects; p.matrices = cam->d_matrices; p.batches = j.d_batches;
    p.wg_info = wg_info; p.region_first_inv = j.d_region_first_inv; p.header = j.d_header;
    p.idx_pred = (uint32_t*)cam->index_buffer.d + cam->index_buffer.out_off();
    p.idx_resid = (uint32_t*)cam->index_buffer.d + cam->index_buffer.in_off();
    p.dc_pred = (r3_indirect_call*)cam->draw_call_buffer.d + cam->draw_call_buffer.out_off();
    p.dc_resid = (r3_indirect_call*)cam->draw_call_buffer.d + cam->draw_call_buffer.in_off();
    p.res_out = (uint32_t*)cam->results_buffer.d + cam->results_buffer.out_off();
    p.res_in = (const uint32_t*)cam->results_buffer.d + cam->results_buffer.in_off();
    p.res_in_words = cam->results_buffer.capacity_elements / 2;
    p.resid_bits = resid_bits; p.sb_counts = sb_counts; p.region_prefix = region_prefix;
    const bool viewport = cam->header.shadow_index == R3_CAMERA_VIEWPORT;
    p.hiz = c->d_hiz_ptrs; p.hiz_dims = c->d_hiz_dims; p.hiz_mips = viewport ? (uint32_t)c->d_hiz.size() : 0u;

    // multi-GPU shard of the viewport's test (r3_set_cull_shard): needs the peers' staging arrays, no non-atomic (blend) objects — their
    // in-place index slots are not exchanged — and a staging capacity that covers this frame; otherwise every rank tests everything
    const bool sharded = viewport && c->tri_shard_count > 1u && c->peer.connected && c->peer.tri_cap_words >= words && !c->any_blend;
    p.shard_index = sharded ? c->tri_shard_index : 0u; p.shard_count = sharded ? c->tri_shard_count : 1u;
    p.ex_n = sharded ? c->peer.n_ranks : 0u; p.ex_cap_words = c->peer.tri_cap_words;
    for (uint32_t r = 0; r < R3_MAX_EXCHANGE_RANKS; ++r) p.ex_peers[r] = sharded ? c->peer.tri_words[r] : nullptr;
    // resident CTAs per SM of the persistent test kernel (a property of the binary).  Two register budgets are compiled: 64 registers / 4 CTAs
    // (default: no spills; config 3 measured 1.63 ms per frame against 1.79 ms) and 48 / 5 (R3_TEST_CTAS=5, for experiments)
    static int test_variant = 0, test_ctas_per_sm = 0;
    if (!test_ctas_per_sm) {
        const char* e = getenv("R3_TEST_CTAS");
        test_variant = (e && e[0] == '5') ? 5 : 4;
        const cudaError_t rc = test_variant == 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&test_ctas_per_sm, triangle_test_kernel<4>, TC_THREADS, 0)
                                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&test_ctas_per_sm, triangle_test_kernel<5>, TC_THREADS, 0);
        if (rc != cudaSuccess || test_ctas_per_sm < 1) test_ctas_per_sm = 4;
    }
    const uint32_t test_grid = (uint32_t)R3_SM_COUNT * (uint32_t)test_ctas_per_sm, need_ctas = (n_wg + TC_THREADS / 32 - 1) / (TC_THREADS / 32);
    r3_stage_begin(c, R3_STAGE_TRIANGLE_TEST);
    if (test_variant == 4) triangle_test_kernel<4><<<need_ctas < test_grid ? need_ctas : test_grid, TC_THREADS, 0, c->stream>>>(p);
    else triangle_test_kernel<5><<<need_ctas < test_grid ? need_ctas : test_grid, TC_THREADS, 0, c->stream>>>(p);
    r3_stage_end(c);
    R3_CHECK_LAUNCH(c, "triangle_test_kernel");
    if (sharded) {
        // publish this rank's words (epoch flag, kind 3), wait for everybody's, then move them into the local buffers
        R3_TRY(r3_peer_signal(c, 3u));
        uint32_t expected[R3_MAX_EXCHANGE_RANKS];
        for (auto& e : expected) e = c->peer.sent[3];
        R3_TRY(r3_peer_wait(c, 3u, expected));
        shard_unpack_kernel<<<n_sb, SB_WORDS, 0, c->stream>>>(p);
        R3_CHECK_LAUNCH(c, "shard_unpack_kernel");
    }
    superblock_scan_kernel<<<1, 1024, 0, c->stream>>>(sb_counts, j.d_header);
    R3_CHECK_LAUNCH(c, "superblock_scan_kernel");
    region_finish_kernel<<<j.n_regions, 256, 0, c->stream>>>(p);
    R3_CHECK_LAUNCH(c, "region_finish_kernel");
    r3_stage_begin(c, R3_STAGE_TRIANGLE_COMPACT);
    triangle_compact_kernel<<<n_sb, SB_WORDS, 0, c->stream>>>(p);
    r3_stage_end(c);
    R3_CHECK_LAUNCH(c, "triangle_compact_kernel");
    return R3_OK;
}
