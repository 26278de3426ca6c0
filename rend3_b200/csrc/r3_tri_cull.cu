// r3_tri_cull.cu — per-triangle cull + ORDERED index compaction (single pass).
//
// Replaces GpuCuller::cull + cull.wgsl::cs_main (rend3-routine/src/culling/culler.rs:531-659,
// rend3-routine/shaders/src/cull.wgsl:264-390): back-face determinant, misses-pixel-centre, hi-Z occlusion,
// predicted/residual index lists, indirect draw records and the per-invocation visibility bits, written into
// the same ping-pong CullingBuffers layout (culler.rs:88-125, suballoc.rs) the forward stage reads.
//
// B200 design.  The reference launches one dispatch per 256-object batch and appends survivors with a global
// atomicAdd per triangle on a handful of contended counters, which also makes the list order nondeterministic.
// Here ONE launch covers every batch; a CTA is one reference workgroup (256 invocations) taken from an atomic
// ticket, each warp's ballot is the 32-bit visibility word the reference assembles with workgroup atomics, and
// the output slot of every surviving triangle comes from a decoupled look-back that is *segmented by region*:
// the first workgroup of a region publishes its count as a finished prefix, later ones look back only until
// they meet one.  Survivors therefore land in ascending invocation order (one legal outcome of the reference's
// atomics), without re-reading anything: the indices are still in registers when the slot is known.  The last
// workgroup of a region writes its two IndirectCall records.
#include "r3_common.cuh"

namespace {

constexpr int TC_THREADS = 256;

struct TriCullParams {
    r3_camera_header cam;
    const uint32_t* mesh; uint64_t mesh_words;
    const r3_object* objects;
    const r3_object_matrices* matrices;
    const r3_batch_data* batches;
    const uint32_t* wg_info;             // per workgroup: (batch << 8) | batch-local object
    const uint32_t* region_first_inv;    // [n_regions + 1]
    const uint32_t* header;              // job header: [1] n_batches, [3] total_invocations (device-side counts)
    uint32_t* idx_pred; uint32_t* idx_resid;
    r3_indirect_call* dc_pred; r3_indirect_call* dc_resid;
    uint32_t* res_out; const uint32_t* res_in; uint64_t res_in_words;
    float* const* hiz; const uint32_t* hiz_dims; uint32_t hiz_mips;
    unsigned long long* state;           // [0] ticket, [1 + wg] descriptors
};

// descriptor: [63:62] flag (0 invalid, 1 aggregate, 2 prefix) | [61:31] predicted count | [30:0] residual count
constexpr unsigned long long TD_AGG = 1ull << 62, TD_PREFIX = 2ull << 62, TD_MASK = (1ull << 31) - 1;
__device__ __forceinline__ unsigned long long td_pack(unsigned long long flag, uint32_t pred, uint32_t resid) {
    return flag | ((unsigned long long)pred << 31) | (unsigned long long)resid;
}
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void st_vol(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }

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
__device__ bool execute_culling(const TriCullParams& p, const float* __restrict__ mvp, const float3 a, const float3 b, const float3 c) {
    const float4 p0 = mat_point_rn(mvp, a.x, a.y, a.z), p1 = mat_point_rn(mvp, b.x, b.y, b.z), p2 = mat_point_rn(mvp, c.x, c.y, c.z);
    const float t0 = sub_rn(mul_rn(p1.y, p2.w), mul_rn(p2.y, p1.w));
    const float t1 = sub_rn(mul_rn(p0.y, p2.w), mul_rn(p2.y, p0.w));
    const float t2 = sub_rn(mul_rn(p0.y, p1.w), mul_rn(p1.y, p0.w));
    const float det = add_rn(sub_rn(mul_rn(p0.x, t0), mul_rn(p1.x, t1)), mul_rn(p2.x, t2));
    const bool positive = p.cam.flags & R3_PCU_POSITIVE_AREA_VISIBLE;
    if (positive && det <= 0.0f) return false;
    if (!positive && det >= 0.0f) return false;
    const float n0x = div_rn(p0.x, p0.w), n0y = div_rn(p0.y, p0.w), n0z = div_rn(p0.z, p0.w);
    const float n1x = div_rn(p1.x, p1.w), n1y = div_rn(p1.y, p1.w), n1z = div_rn(p1.z, p1.w);
    const float n2x = div_rn(p2.x, p2.w), n2y = div_rn(p2.y, p2.w), n2z = div_rn(p2.z, p2.w);
    const float minx = fminf(n0x, fminf(n1x, n2x)), miny = fminf(n0y, fminf(n1y, n2y));
    const float maxx = fmaxf(n0x, fmaxf(n1x, n2x)), maxy = fmaxf(n0y, fmaxf(n1y, n2y));
    const float hrx = div_rn(p.cam.resolution[0], 2.0f), hry = div_rn(p.cam.resolution[1], 2.0f);
    const float minsx = mul_rn(add_rn(minx, 1.0f), hrx), minsy = mul_rn(add_rn(miny, 1.0f), hry);
    const float maxsx = mul_rn(add_rn(maxx, 1.0f), hrx), maxsy = mul_rn(add_rn(maxy, 1.0f), hry);
    if (!(p.cam.flags & R3_PCU_MULTISAMPLED)) {
        if (rintf(minsx) == rintf(maxsx) || rintf(minsy) == rintf(maxsy)) return false;   // WGSL round(): ties to even
    }
    if (p.cam.shadow_index != R3_CAMERA_VIEWPORT) return true;
    const float mintx = div_rn(add_rn(minx, 1.0f), 2.0f), minty = sub_rn(1.0f, div_rn(add_rn(miny, 1.0f), 2.0f));
    const float maxtx = div_rn(add_rn(maxx, 1.0f), 2.0f), maxty = sub_rn(1.0f, div_rn(add_rn(maxy, 1.0f), 2.0f));
    const float u = div_rn(add_rn(maxtx, mintx), 2.0f), v = div_rn(add_rn(maxty, minty), 2.0f);
    const float ex = sub_rn(maxsx, minsx), ey = sub_rn(maxsy, minsy);
    const uint32_t mip = ceil_log2_f32(fmaxf(fmaxf(ex, ey), 1.0f));
    const float depth = fmaxf(fmaxf(n0z, n1z), n2z);
    const float occl = hiz_sample_min(p, u, v, mip);
    return !(depth < occl);
}

__global__ void expand_wg_info_kernel(const r3_batch_data* __restrict__ batches, const uint32_t* __restrict__ header, uint32_t* __restrict__ wg_info) {
    const uint32_t b = blockIdx.x, o = threadIdx.x;
    if (b >= header[1]) return;
    const r3_batch_data* job = &batches[b];
    if (o >= job->total_objects) return;
    const r3_object_culling_info info = job->object_culling_information[o];
    const uint32_t n = info.invocation_end - info.invocation_start;
    const uint32_t first = (job->batch_base_invocation + info.invocation_start) >> 8, count = (n + 255u) >> 8;
    for (uint32_t i = 0; i < count; ++i) wg_info[first + i] = (b << 8) | o;
}

__global__ void __launch_bounds__(TC_THREADS) triangle_cull_kernel(const __grid_constant__ TriCullParams p) {
    __shared__ uint32_t s_wg;
    __shared__ uint32_t s_pred[8], s_resid[8];
    __shared__ uint32_t s_base_pred, s_base_resid;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool shadow = p.cam.shadow_index != R3_CAMERA_VIEWPORT;
    const uint32_t n_workgroups = p.header[3] / TC_THREADS;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_wg = (uint32_t)atomicAdd(&p.state[0], 1ull);
        __syncthreads();
        const uint32_t wg = s_wg;
        if (wg >= n_workgroups) return;

        const uint32_t wi = __ldg(&p.wg_info[wg]);
        const r3_batch_data* job = &p.batches[wi >> 8];
        const uint32_t local_object = wi & 0xFFu;
        const r3_object_culling_info info = job->object_culling_information[local_object];   // find_object_info (cull.wgsl:181-207)
        const uint32_t global_invocation = wg * TC_THREADS + threadIdx.x;
        const uint32_t gid = global_invocation - job->batch_base_invocation;                 // invocation within the batch
        const bool real = gid < info.invocation_end;
        const uint32_t object_invocation = gid - info.invocation_start;

        uint32_t pk0 = R3_INVALID_VERTEX, pk1 = R3_INVALID_VERTEX, pk2 = R3_INVALID_VERTEX;
        bool passes = false, resid = false;
        if (real) {
            const r3_object* obj = &p.objects[info.object_id];
            const uint32_t first_index = obj->first_index, pos_off = obj->attr_offset[0] >> 2;
            const uint64_t ib = (uint64_t)first_index + (uint64_t)object_invocation * 3u;      // vertex_fetch (cull.wgsl:9-32)
            const uint32_t i0 = mesh_word(p, ib), i1 = mesh_word(p, ib + 1), i2 = mesh_word(p, ib + 2);
            float3 v[3];
            const uint32_t ids[3] = {i0, i1, i2};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint64_t f = (uint64_t)pos_off + (uint64_t)ids[k] * 3u;                  // extract_attribute_vec3_f32
                v[k] = make_float3(__uint_as_float(mesh_word(p, f)), __uint_as_float(mesh_word(p, f + 1)), __uint_as_float(mesh_word(p, f + 2)));
            }
            passes = execute_culling(p, p.matrices[info.object_id].model_view_proj, v[0], v[1], v[2]);
            pk0 = (local_object << 24) | (i0 & 0xFFFFFFu); pk1 = (local_object << 24) | (i1 & 0xFFFFFFu); pk2 = (local_object << 24) | (i2 & 0xFFFFFFu);
            if (passes && !shadow && info.atomic_capable == 1u) {
                bool prev = false;                                                            // get_previous_culling_result (cull.wgsl:152-160)
                if (info.previous_global_invocation != R3_NO_PREVIOUS) {
                    const uint64_t pgi = (uint64_t)object_invocation + info.previous_global_invocation;
                    const uint32_t mask = (pgi >> 5) < p.res_in_words ? p.res_in[pgi >> 5] : 0u;
                    prev = (mask >> (pgi & 31)) & 1u;
                }
                resid = !prev;
            }
        }
        const uint32_t word_pred = __ballot_sync(0xFFFFFFFFu, passes);
        const uint32_t word_resid = __ballot_sync(0xFFFFFFFFu, resid);
        if (lane == 0) p.res_out[global_invocation >> 5] = word_pred;                        // save_culling_results (cull.wgsl:229-241)

        if (info.atomic_capable == 0u) {
            // non-atomic (blend) objects keep their slot: survivors in place, everything else INVALID (cull.wgsl:374-380,343-347)
            const uint64_t o = (uint64_t)global_invocation * 3u;
            p.idx_resid[o] = passes ? pk0 : R3_INVALID_VERTEX; p.idx_resid[o + 1] = passes ? pk1 : R3_INVALID_VERTEX;
            p.idx_resid[o + 2] = passes ? pk2 : R3_INVALID_VERTEX;
        }
        if (lane == 0) { s_pred[warp] = __popc(word_pred); s_resid[warp] = __popc(word_resid); }
        __syncthreads();

        const uint32_t region = info.region_id;
        const uint32_t region_first = __ldg(&p.region_first_inv[region]), region_end = __ldg(&p.region_first_inv[region + 1]);
        const bool first_of_region = (region_first >> 8) == wg, last_of_region = (region_end >> 8) == wg + 1;
        if (warp == 0) {
            uint32_t cp = lane < 8 ? s_pred[lane] : 0u, cr = lane < 8 ? s_resid[lane] : 0u;
            uint32_t ip = cp, ir = cr;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const uint32_t np = __shfl_up_sync(0xFFFFFFFFu, ip, d), nr = __shfl_up_sync(0xFFFFFFFFu, ir, d);
                if (lane >= d) { ip += np; ir += nr; }
            }
            const uint32_t tot_p = __shfl_sync(0xFFFFFFFFu, ip, 7), tot_r = __shfl_sync(0xFFFFFFFFu, ir, 7);
            if (lane < 8) { s_pred[lane] = ip - cp; s_resid[lane] = ir - cr; }                // exclusive offsets of the 8 warps
            uint32_t run_p = 0, run_r = 0;
            if (!first_of_region) {
                if (lane == 0) st_vol(&p.state[1 + wg], td_pack(TD_AGG, tot_p, tot_r));
                int pred = (int)wg - 1;
                for (;;) {
                    const int idx = pred - lane;
                    unsigned long long d = (idx >= 0) ? ld_vol(&p.state[1 + idx]) : TD_PREFIX;
                    while (__any_sync(0xFFFFFFFFu, (d >> 62) == 0ull)) {
                        if ((d >> 62) == 0ull) d = ld_vol(&p.state[1 + idx]);
                    }
                    const uint32_t pmask = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2ull);
                    const int first = pmask ? (__ffs(pmask) - 1) : 31;
                    uint32_t vp = (lane <= first) ? (uint32_t)((d >> 31) & TD_MASK) : 0u, vr = (lane <= first) ? (uint32_t)(d & TD_MASK) : 0u;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) { vp += __shfl_xor_sync(0xFFFFFFFFu, vp, s); vr += __shfl_xor_sync(0xFFFFFFFFu, vr, s); }
                    run_p += vp; run_r += vr;
                    if (pmask) break;
                    pred -= 32;
                }
            }
            if (lane == 0) {
                st_vol(&p.state[1 + wg], td_pack(TD_PREFIX, run_p + tot_p, run_r + tot_r));
                s_base_pred = run_p; s_base_resid = run_r;
                if (last_of_region) {
                    // init_draw_calls + the final vertex_count the atomics would have reached (cull.wgsl:47-73)
                    const bool atomic_region = info.atomic_capable == 1u;
                    r3_indirect_call pc, rc;
                    pc.vertex_count = atomic_region ? 3u * (run_p + tot_p) : 0u;
                    rc.vertex_count = atomic_region ? 3u * (run_r + tot_r) : 3u * (region_end - region_first);
                    pc.instance_count = rc.instance_count = 1u;
                    pc.base_index = rc.base_index = region_first * 3u;
                    pc.vertex_offset = rc.vertex_offset = 0;
                    pc.base_instance = rc.base_instance = 0u;
                    p.dc_pred[region] = pc;
                    p.dc_resid[region] = rc;
                }
            }
        }
        __syncthreads();
        if (info.atomic_capable == 1u) {
            const uint32_t lt = (1u << lane) - 1u;
            if (passes) {                                                                     // write_predicted_atomic_triangle (cull.wgsl:84-99)
                const uint64_t slot = (uint64_t)region_first + s_base_pred + s_pred[warp] + __popc(word_pred & lt);
                p.idx_pred[slot * 3] = pk0; p.idx_pred[slot * 3 + 1] = pk1; p.idx_pred[slot * 3 + 2] = pk2;
            }
            if (resid) {                                                                      // write_residual_atomic_triangle (cull.wgsl:101-116)
                const uint64_t slot = (uint64_t)region_first + s_base_resid + s_resid[warp] + __popc(word_resid & lt);
                p.idx_resid[slot * 3] = pk0; p.idx_resid[slot * 3 + 1] = pk1; p.idx_resid[slot * 3 + 2] = pk2;
            }
        }
    }
}

}  // namespace

int r3_launch_triangle_cull(r3_ctx* c, r3_camera* cam) {
    r3_jobs& j = cam->jobs[cam->cur];
    const uint64_t inv = j.total_invocations, words = (inv + 31) / 32;
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
    if (n_wg == 0) return R3_OK;

    // scratch: wg_info [n_wg] + look-back state [1 + n_wg]
    R3_TRY(r3_reserve_t(c, &cam->d_resid_bits, &cam->resid_bits_cap, n_wg));
    R3_TRY(r3_reserve_t(c, &cam->d_word_scan, &cam->word_scan_cap, (uint64_t)n_wg + 1));
    R3_CUDA(c, cudaMemsetAsync(cam->d_word_scan, 0, ((size_t)n_wg + 1) * 8, c->stream));
    expand_wg_info_kernel<<<j.n_batches, 256, 0, c->stream>>>(j.d_batches, j.d_header, cam->d_resid_bits);
    R3_CHECK_LAUNCH(c, "expand_wg_info_kernel");

    TriCullParams p;
    p.cam = cam->header;
    p.mesh = c->d_mesh; p.mesh_words = c->mesh_words;
    p.objects = c->d_objects; p.matrices = cam->d_matrices; p.batches = j.d_batches;
    p.wg_info = cam->d_resid_bits; p.region_first_inv = j.d_region_first_inv; p.header = j.d_header;
    p.idx_pred = (uint32_t*)cam->index_buffer.d + cam->index_buffer.out_off();
    p.idx_resid = (uint32_t*)cam->index_buffer.d + cam->index_buffer.in_off();
    p.dc_pred = (r3_indirect_call*)cam->draw_call_buffer.d + cam->draw_call_buffer.out_off();
    p.dc_resid = (r3_indirect_call*)cam->draw_call_buffer.d + cam->draw_call_buffer.in_off();
    p.res_out = (uint32_t*)cam->results_buffer.d + cam->results_buffer.out_off();
    p.res_in = (const uint32_t*)cam->results_buffer.d + cam->results_buffer.in_off();
    p.res_in_words = cam->results_buffer.capacity_elements / 2;
    const bool viewport = cam->header.shadow_index == R3_CAMERA_VIEWPORT;
    p.hiz = c->d_hiz_ptrs; p.hiz_dims = c->d_hiz_dims; p.hiz_mips = viewport ? (uint32_t)c->d_hiz.size() : 0u;
    p.state = cam->d_word_scan;
    const uint32_t grid = n_wg < (uint32_t)(R3_SM_COUNT * 8) ? n_wg : (uint32_t)(R3_SM_COUNT * 8);
    triangle_cull_kernel<<<grid, TC_THREADS, 0, c->stream>>>(p);
    R3_CHECK_LAUNCH(c, "triangle_cull_kernel");
    return R3_OK;
}
