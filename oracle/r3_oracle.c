/* r3_oracle.c — CPU ORACLE (test infrastructure; never linked into or called by the product).
 *
 * Plain-C restatement of the reference's arithmetic for rend3's GPU-driven hot path, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline to check / time against.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (strict IEEE f32, source order, no FMA:
 * SURVEY.md D7).  Every function cites the reference lines it follows (paths relative to the
 * reference repository root).
 *
 * Pinning: tests/test_oracle_golden.py renders the scenes of rend3-test/tests/{simple,object,
 * shadow,msaa}.rs and examples/src/cube through this file and compares with the reference's own
 * PNG goldens (rend3-test/tests/results/) — see DESIGN.md "oracle".
 *
 * Where the reference defers to the wgpu/Vulkan driver (triangle setup, fill rule, clipping,
 * interpolation, bilinear-compare, f16 store, robust out-of-bounds access) the rule implemented
 * here is stated at the function and in DESIGN.md "raster rules"; the CUDA path implements the
 * same rule independently.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rend3_b200.h" /* status codes + the entry-point list being mirrored (r3_ -> r3o_) */
#include "r3_oracle.h"

#define API __attribute__((visibility("default")))

static int fail(r3o_ctx* c, int code, const char* msg) {
    if (c) snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}
static int cam_slot(uint32_t camera) { return camera == R3_CAMERA_VIEWPORT ? 0 : (int)camera + 1; }
#define CAM_OR_FAIL(ctx, camera)                                                             \
    if ((camera) != R3_CAMERA_VIEWPORT && (camera) >= R3O_MAX_CAMERAS - 1) return fail(ctx, R3_E_INVALID, "bad camera"); \
    r3o_camera* cam = &(ctx)->cams[cam_slot(camera)]

/* ------------------------------------------------------------------ small math, source-order f32 */
typedef struct { float x, y, z, w; } v4;
typedef struct { float x, y, z; } v3;

/* WGSL mat4x4 * vec4: sum of column*component, accumulated x,y,z,w (cull.wgsl:268, opaque.wgsl:126) */
static inline v4 mat_vec(const float* m, v4 v) {
    v4 r;
    r.x = m[0] * v.x; r.y = m[1] * v.x; r.z = m[2] * v.x; r.w = m[3] * v.x;
    r.x = r.x + m[4] * v.y; r.y = r.y + m[5] * v.y; r.z = r.z + m[6] * v.y; r.w = r.w + m[7] * v.y;
    r.x = r.x + m[8] * v.z; r.y = r.y + m[9] * v.z; r.z = r.z + m[10] * v.z; r.w = r.w + m[11] * v.z;
    r.x = r.x + m[12] * v.w; r.y = r.y + m[13] * v.w; r.z = r.z + m[14] * v.w; r.w = r.w + m[15] * v.w;
    return r;
}
/* MVP * vec4(p, 1.0): the last term is the translation column itself */
static inline v4 mat_point(const float* m, float px, float py, float pz) {
    v4 r;
    r.x = m[0] * px; r.y = m[1] * px; r.z = m[2] * px; r.w = m[3] * px;
    r.x = r.x + m[4] * py; r.y = r.y + m[5] * py; r.z = r.z + m[6] * py; r.w = r.w + m[7] * py;
    r.x = r.x + m[8] * pz; r.y = r.y + m[9] * pz; r.z = r.z + m[10] * pz; r.w = r.w + m[11] * pz;
    r.x = r.x + m[12]; r.y = r.y + m[13]; r.z = r.z + m[14]; r.w = r.w + m[15];
    return r;
}
/* WGSL mat4x4 * mat4x4: column j of the result is A * B[j] (uniform_prep.wgsl:22-23) */
static inline void mat_mul(const float* a, const float* b, float* out) {
    for (int j = 0; j < 4; ++j) {
        v4 c = {b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]};
        v4 r = mat_vec(a, c);
        out[4 * j] = r.x; out[4 * j + 1] = r.y; out[4 * j + 2] = r.z; out[4 * j + 3] = r.w;
    }
}
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* WGSL min/max: if one operand is NaN the other is returned (IEEE minNum/maxNum) — the reference relies
 * on this to drop NaNs (opaque.wgsl:545 `max(surface_shading(..), vec3(0.0))`) */
static inline float fmin2(float a, float b) { return fminf(a, b); }
static inline float fmax2(float a, float b) { return fmaxf(a, b); }
static inline float saturate(float v) { return fmin2(fmax2(v, 0.0f), 1.0f); }
static inline v3 normalize3(v3 a) { float l = sqrtf(dot3(a, a)); v3 r = {a.x / l, a.y / l, a.z / l}; return r; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* f32 -> f16 round-to-nearest-even (Rgba16Float colour target, forward.rs:325-329) */
static uint16_t f32_to_f16(float f) {
    uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u, mant = x & 0x7FFFFFu;
    int32_t e = (int32_t)((x >> 23) & 0xFF);
    if (e == 255) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u | (mant >> 13) : 0));
    e = e - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7C00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        mant |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e), half = mant >> shift, rem = mant & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)e << 10) | (mant >> 13), rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}
static float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1F, m = h & 0x3FFu;
    if (e == 0) {
        if (m == 0) return u2f(sign);
        float v = (float)m * 5.9604644775390625e-8f; /* 2^-24 */
        return (h & 0x8000u) ? -v : v;
    }
    if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
    return u2f(sign | ((e + 112) << 23) | (m << 13));
}

/* ------------------------------------------------------------------ context + uploads */
API uint32_t r3o_abi_version(void) { return 1; }
API int r3o_ctx_create(int device, r3o_ctx** out) {
    (void)device;
    if (!out) return R3_E_INVALID;
    r3o_ctx* c = (r3o_ctx*)calloc(1, sizeof(r3o_ctx));
    if (!c) return R3_E_OOM;
    c->samples = 1;
    *out = c;
    return R3_OK;
}
static void iobuf_free(r3o_iobuf* b) { free(b->data); memset(b, 0, sizeof *b); }
API int r3o_ctx_destroy(r3o_ctx* c) {
    if (!c) return R3_E_INVALID;
    free(c->objects); free(c->sort_key); free(c->sort_flags); free(c->sort_loc); free(c->mesh); free(c->materials);
    free(c->dir_lights); free(c->point_lights);
    for (int i = 0; i < R3O_MAX_CAMERAS; ++i) {
        r3o_camera* k = &c->cams[i];
        free(k->matrices); free(k->visible); free(k->flag); free(k->batches); free(k->regions); free(k->prev_batches); free(k->prev_regions);
        free(k->prev_invocation);
        iobuf_free(&k->index_buffer); iobuf_free(&k->draw_call_buffer); iobuf_free(&k->results_buffer);
    }
    free(c->vis); free(c->hdr); free(c->hdr16); free(c->depth); free(c->ldr); free(c->atlas);
    for (uint32_t i = 0; i < c->hiz_mips; ++i) free(c->hiz[i]);
    free(c->hiz); free(c->hiz_w); free(c->hiz_h);
    free(c->tex_descs); free(c->texels); free(c->sky_texels); free(c->tris[0]); free(c->tris[1]); free(c->tris[2]); free(c->tris[3]); free(c->sample_col16); free(c->blended);
    free(c);
    return R3_OK;
}
API const char* r3o_last_error(const r3o_ctx* c) { return c ? c->err : "null context"; }
API int r3o_sync(r3o_ctx* c) { (void)c; return R3_OK; }
API int r3o_get_stream(r3o_ctx* c, void** s) { (void)c; if (s) *s = 0; return R3_OK; }
API int r3o_launch_count(r3o_ctx* c, uint64_t* n) { (void)c; if (n) *n = 0; return R3_OK; }

#define REPLACE(ptr, type, src, count)                                        \
    do {                                                                      \
        free(ptr);                                                            \
        ptr = (type*)malloc((size_t)(count) * sizeof(type) + 16);             \
        if (!ptr) return fail(c, R3_E_OOM, "out of memory");                  \
        memcpy(ptr, src, (size_t)(count) * sizeof(type));                     \
    } while (0)

API int r3o_set_objects(r3o_ctx* c, const r3_object* recs, uint32_t n) {
    if (!c || (!recs && n)) return fail(c, R3_E_INVALID, "set_objects: null");
    REPLACE(c->objects, r3_object, recs, n);
    c->n_slots = n;
    return R3_OK;
}
API int r3o_set_objects_device(r3o_ctx* c, const void* p, uint32_t n) { return r3o_set_objects(c, (const r3_object*)p, n); }
API int r3o_update_objects(r3o_ctx* c, const uint32_t* slots, const r3_object* recs, uint32_t n) {
    if (!c || !slots || !recs) return fail(c, R3_E_INVALID, "update_objects: null");
    for (uint32_t i = 0; i < n; ++i)
        if (slots[i] < c->n_slots) c->objects[slots[i]] = recs[i]; /* out-of-range scatter writes are dropped */
    return R3_OK;
}
API int r3o_set_object_sort_info(r3o_ctx* c, const uint64_t* key, const uint8_t* flags, const float* loc, uint32_t n) {
    if (!c || !key || !flags || !loc) return fail(c, R3_E_INVALID, "sort_info: null");
    REPLACE(c->sort_key, uint64_t, key, n);
    REPLACE(c->sort_flags, uint8_t, flags, n);
    REPLACE(c->sort_loc, float, loc, 3 * (size_t)n);
    c->sort_n = n;
    return R3_OK;
}
API int r3o_set_mesh_buffer(r3o_ctx* c, const void* bytes, uint64_t nbytes) {
    if (!c || (!bytes && nbytes) || (nbytes & 3)) return fail(c, R3_E_INVALID, "mesh buffer");
    REPLACE(c->mesh, uint32_t, bytes, nbytes / 4);
    c->mesh_words = nbytes / 4;
    return R3_OK;
}
API int r3o_set_materials(r3o_ctx* c, const r3_material* recs, uint32_t n) {
    if (!c || (!recs && n)) return fail(c, R3_E_INVALID, "materials");
    REPLACE(c->materials, r3_material, recs, n);
    c->n_materials = n;
    return R3_OK;
}
/* TextureManager<D2> table (rend3/src/managers/texture.rs): validated here so the samplers can index without checks */
API int r3o_set_textures(r3o_ctx* c, const r3_texture_desc* descs, uint32_t n, const void* texels, uint64_t nbytes) {
    if (!c || (!descs && n) || (!texels && nbytes)) return fail(c, R3_E_INVALID, "textures");
    for (uint32_t i = 0; i < n; ++i) {
        const r3_texture_desc* d = &descs[i];
        if (!d->width || !d->height || !d->mip_count || d->mip_count > 32 || d->format >= R3_TEXFMT_COUNT) return fail(c, R3_E_INVALID, "textures: bad descriptor");
        uint64_t total = 0;
        for (uint32_t l = 0; l < d->mip_count; ++l) {
            uint32_t w = (d->width >> l) ? (d->width >> l) : 1, h = (d->height >> l) ? (d->height >> l) : 1;
            total += R3_TEXFMT_LEVEL_BYTES(d->format, w, h);
        }
        if (d->byte_offset % 16 || d->byte_offset + total > nbytes) return fail(c, R3_E_INVALID, "textures: mip chain outside the texel blob");
    }
    REPLACE(c->tex_descs, r3_texture_desc, descs, n);
    free(c->texels);
    c->texels = (uint8_t*)malloc(nbytes + 16);
    if (nbytes) memcpy(c->texels, texels, nbytes);
    c->n_textures = n; c->texel_bytes = nbytes;
    return R3_OK;
}
static uint64_t cube_face_bytes(const r3_texture_desc* d) {
    uint64_t bpp = d->format == R3_TEXFMT_RGBA32_FLOAT ? 16 : 4, total = 0;
    for (uint32_t l = 0; l < d->mip_count; ++l) { uint64_t w = (d->width >> l) ? (d->width >> l) : 1; total += w * w * bpp; }
    return total;
}
API int r3o_set_skybox(r3o_ctx* c, const r3_texture_desc* desc, const void* texels, uint64_t nbytes) {
    if (!c) return R3_E_INVALID;
    free(c->sky_texels); c->sky_texels = NULL; c->has_skybox = 0;
    if (!desc) return R3_OK;
    if (!texels || !desc->width || !desc->mip_count || desc->mip_count > 32 || desc->format > R3_TEXFMT_RGBA32_FLOAT) return fail(c, R3_E_INVALID, "skybox: bad descriptor");
    if (desc->byte_offset % 16 || desc->byte_offset + 6 * cube_face_bytes(desc) > nbytes) return fail(c, R3_E_INVALID, "skybox: faces outside the texel blob");
    c->sky_texels = (uint8_t*)malloc(nbytes + 16);
    memcpy(c->sky_texels, texels, nbytes);
    c->sky_desc = *desc; c->sky_desc.height = desc->width; c->has_skybox = 1;
    return R3_OK;
}
API int r3o_set_directional_lights(r3o_ctx* c, const void* bytes, uint64_t nbytes, uint32_t aw, uint32_t ah) {
    if (!c || !bytes || nbytes < 16) return fail(c, R3_E_INVALID, "directional lights");
    uint32_t n = *(const uint32_t*)bytes;
    if (nbytes < 16 + (uint64_t)n * 128) return fail(c, R3_E_INVALID, "directional lights: short buffer");
    REPLACE(c->dir_lights, r3_directional_light, (const uint8_t*)bytes + 16, n);
    c->n_dir = n;
    if (aw != c->atlas_w || ah != c->atlas_h || !c->atlas) {
        free(c->atlas);
        c->atlas = (float*)calloc((size_t)aw * ah + 1, 4);
        c->atlas_w = aw; c->atlas_h = ah;
    }
    return R3_OK;
}
API int r3o_set_point_lights(r3o_ctx* c, const void* bytes, uint64_t nbytes) {
    if (!c || !bytes || nbytes < 16) return fail(c, R3_E_INVALID, "point lights");
    uint32_t n = *(const uint32_t*)bytes;
    if (nbytes < 16 + (uint64_t)n * 32) return fail(c, R3_E_INVALID, "point lights: short buffer");
    REPLACE(c->point_lights, r3_point_light, (const uint8_t*)bytes + 16, n);
    c->n_point = n;
    return R3_OK;
}
API int r3o_set_frame_uniforms(r3o_ctx* c, const r3_frame_uniforms* u) {
    if (!c || !u) return fail(c, R3_E_INVALID, "frame uniforms");
    c->uniforms = *u;
    return R3_OK;
}

/* ------------------------------------------------------------------ a3: Frustum::contains_sphere
 * rend3/src/util/frustum.rs:148-161 + Plane::distance :79-81 (abc.dot(point) + d, glam scalar dot). */
static inline int frustum_contains_sphere(const float fr[5][4], const float* center, float radius) {
    float neg_radius = -radius;
    for (int p = 0; p < 5; ++p) {
        float dist = ((fr[p][0] * center[0] + fr[p][1] * center[1]) + fr[p][2] * center[2]) + fr[p][3];
        if (!(dist >= neg_radius)) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ a6 + a4(cull): uniform bake + object cull
 * uniform_prep.wgsl:9-27 (MV = view*T, MVP = view_proj*T for idx < object_count, enabled != 0) and
 * the frustum filter of batch_objects (batching.rs:144-148) over enumerated (live) objects. */
API int r3o_object_uniform_upload(r3o_ctx* c, uint32_t camera, const r3_camera_header* h, uint32_t mode) {
    if (!c || !h) return fail(c, R3_E_INVALID, "uniform_upload: null");
    CAM_OR_FAIL(c, camera);
    cam->header = *h;
    cam->header_set = 1;
    uint32_t n = h->object_count;
    if (n > c->n_slots) return fail(c, R3_E_INVALID, "object_count exceeds object buffer");
    if (cam->matrices_cap < n) {
        /* a resized per-camera buffer starts zeroed (culler.rs:459-476) */
        free(cam->matrices);
        cam->matrices = (r3_object_matrices*)calloc((size_t)n + 1, sizeof(r3_object_matrices));
        cam->matrices_cap = n;
    }
    if (cam->visible_cap < n) {
        free(cam->visible);
        cam->visible = (uint32_t*)malloc(((size_t)n + 1) * 4);
        cam->visible_cap = n;
    }
    if (cam->flag_cap < n) {
        free(cam->flag);
        cam->flag = (uint8_t*)malloc((size_t)n + 1);
        cam->flag_cap = n;
    }
    uint8_t* flag = cam->flag;
    if (!flag || !cam->matrices || !cam->visible) return fail(c, R3_E_OOM, "out of memory");
    int have_live = c->sort_flags && c->sort_n >= n;
    /* every host thread owns one contiguous range of slots: bake + test it, count its survivors, then — after an exclusive prefix over
     * the per-thread counts — write them at their final positions.  The visible list stays ascending; nothing is allocated per call. */
    uint32_t counts[R3O_MAX_THREADS + 1];
    int n_threads = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads() < R3O_MAX_THREADS ? omp_get_num_threads() : R3O_MAX_THREADS;
#else
        const int t = 0, nt = 1;
#endif
        if (t == 0) n_threads = nt;
        if (t < nt) {
            const uint64_t lo = (uint64_t)n * t / nt, hi = (uint64_t)n * (t + 1) / nt;
            uint32_t cnt = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                const r3_object* o = &c->objects[i];
                if ((mode & R3_CB_BAKE) && o->enabled != 0) {
                    mat_mul(h->view, o->transform, cam->matrices[i].model_view);
                    mat_mul(h->view_proj, o->transform, cam->matrices[i].model_view_proj);
                }
                int live = have_live ? (c->sort_flags[i] & 1) : (o->enabled != 0);
                uint8_t f = (uint8_t)((mode & R3_CB_CULL) && live && frustum_contains_sphere(h->frustum, o->sphere_center, o->sphere_radius));
                flag[i] = f;
                cnt += f;
            }
            counts[t] = cnt;
        }
#pragma omp barrier
        if (t < nt && (mode & R3_CB_CULL)) {
            uint32_t pos = 0;
            for (int k = 0; k < t; ++k) pos += counts[k];
            const uint64_t lo = (uint64_t)n * t / nt, hi = (uint64_t)n * (t + 1) / nt;
            for (uint64_t i = lo; i < hi; ++i)
                if (flag[i]) cam->visible[pos++] = (uint32_t)i;
        }
    }
    uint32_t cnt = 0;
    for (int k = 0; k < n_threads; ++k) cnt += counts[k];
    cam->visible_count = (mode & R3_CB_CULL) ? cnt : 0;
    return R3_OK;
}
API int r3o_visible_count(r3o_ctx* c, uint32_t camera, uint32_t* count) {
    CAM_OR_FAIL(c, camera);
    if (!count) return fail(c, R3_E_INVALID, "null");
    *count = cam->visible_count;
    return R3_OK;
}
API int r3o_readback_visible(r3o_ctx* c, uint32_t camera, uint32_t* out, uint32_t cap, uint32_t* count) {
    CAM_OR_FAIL(c, camera);
    if (count) *count = cam->visible_count;
    if (out) {
        if (cap < cam->visible_count) return fail(c, R3_E_INVALID, "visible: capacity too small");
        memcpy(out, cam->visible, (size_t)cam->visible_count * 4);
    }
    return R3_OK;
}
API int r3o_readback_object_matrices(r3o_ctx* c, uint32_t camera, r3_object_matrices* out, uint32_t first, uint32_t n) {
    CAM_OR_FAIL(c, camera);
    if (!out || (uint64_t)first + n > cam->matrices_cap) return fail(c, R3_E_INVALID, "matrices: range");
    memcpy(out, cam->matrices + first, (size_t)n * sizeof *out);
    return R3_OK;
}

/* ------------------------------------------------------------------ a4: batch_objects (batching.rs:120-250) */
typedef struct { uint64_t material_key; uint32_t reason; float distance; uint32_t handle; } sort_item;
/* ShaderJobSortingKey::cmp (batching.rs:53-79); bind_group_index is DUMMY (equal) in the GpuDriven
 * profile (batching.rs:151).  OrderedFloat total order on the distance: NaN is greater than every number and
 * equal to itself, -0.0 == +0.0 (ordered-float's Ord).  sort_unstable leaves ties unspecified; ties are resolved
 * by handle so the oracle is deterministic. */
static int sort_cmp(const void* pa, const void* pb) {
    const sort_item* a = (const sort_item*)pa; const sort_item* b = (const sort_item*)pb;
    if (a->material_key != b->material_key) return a->material_key < b->material_key ? -1 : 1;
    if (a->reason != b->reason) return a->reason < b->reason ? -1 : 1;
    const int an = a->distance != a->distance, bn = b->distance != b->distance;
    if (an != bn) return an < bn ? -1 : 1;
    if (!an) {
        if (a->distance < b->distance) return -1;
        if (a->distance > b->distance) return 1;
    }
    return a->handle < b->handle ? -1 : (a->handle > b->handle ? 1 : 0);
}
static uint32_t round_up_u32(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

API int r3o_batch_objects(r3o_ctx* c, uint32_t camera, const float vp_loc[3], uint32_t max_dispatch_count) {
    CAM_OR_FAIL(c, camera);
    if (!vp_loc) return fail(c, R3_E_INVALID, "null location");
    if (!cam->header_set) return fail(c, R3_E_STATE, "batch_objects before object_uniform_upload");
    if (!c->sort_key || c->sort_n < cam->header.object_count) return fail(c, R3_E_STATE, "batch_objects needs r3_set_object_sort_info");
    uint32_t nv = cam->visible_count;
    /* get_and_reset_camera (batching.rs:111-113,129): last frame's map is consumed, a fresh one built */
    uint32_t cap = cam->header.object_count;
    uint32_t* prev_map = cam->prev_invocation; uint32_t prev_cap = cam->prev_invocation_cap;
    uint32_t* cur_map = (uint32_t*)malloc(((size_t)cap + 1) * 4);
    memset(cur_map, 0xFF, ((size_t)cap + 1) * 4);
    sort_item* items = (sort_item*)malloc(((size_t)nv + 1) * sizeof(sort_item));
    for (uint32_t i = 0; i < nv; ++i) {
        uint32_t h = cam->visible[i];
        const float* l = &c->sort_loc[3 * (size_t)h];
        float dx = vp_loc[0] - l[0], dy = vp_loc[1] - l[1], dz = vp_loc[2] - l[2];
        float d2 = (dx * dx + dy * dy) + dz * dz;              /* distance_squared (batching.rs:156-157) */
        if (c->sort_flags[h] & 4) d2 = -d2;                      /* BackToFront (batching.rs:158-160) */
        items[i].material_key = c->sort_key[h];
        items[i].reason = (c->sort_flags[h] & 2) ? 0u : 1u;      /* Optimization < Requirement (rend3-types lib.rs:952-957) */
        items[i].distance = d2;
        items[i].handle = h;
    }
    qsort(items, nv, sizeof(sort_item), sort_cmp);

    free(cam->batches); free(cam->regions);
    cam->batches = NULL; cam->regions = NULL; cam->n_batches = cam->n_regions = 0; cam->total_invocations = 0;
    if (nv) {
        uint32_t bcap = nv / R3_BATCH_SIZE + 2, rcap = nv + 2, nb = 0, nr = 0;
        for (uint32_t i = 0; i < nv; ++i) { /* dispatch-limit splits: at most one extra batch per limit hit */
            uint32_t inv = round_up_u32(c->objects[items[i].handle].index_count / 3, R3_WORKGROUP_SIZE);
            bcap += (uint32_t)(((uint64_t)inv) / ((uint64_t)max_dispatch_count * R3_WORKGROUP_SIZE + 1)) + 0;
        }
        bcap += nv; /* generous: a batch per object worst case */
        r3_batch_data* batches = (r3_batch_data*)calloc(bcap, sizeof(r3_batch_data));
        r3_region* regions = (r3_region*)calloc(rcap, sizeof(r3_region));
        uint32_t cur_region_idx = 0, cur_region_obj = 0, cur_base_inv = 0, cur_region_inv = 0, cur_inv = 0, cur_obj = 0;
        uint64_t cur_key = items[0].material_key;
        r3_batch_data cur; memset(&cur, 0, sizeof cur);
        for (uint32_t i = 0; i < nv; ++i) {
            uint32_t h = items[i].handle;
            uint32_t invocation_count = c->objects[h].index_count / 3;
            int key_difference = items[i].material_key != cur_key;
            int object_limit = cur_obj == R3_BATCH_SIZE;
            int dispatch_limit = ((uint64_t)cur_inv + invocation_count) >= (uint64_t)max_dispatch_count * R3_WORKGROUP_SIZE;
            if (key_difference || object_limit || dispatch_limit) {
                regions[nr].job_index = nb; regions[nr].bind_group_index = 0; regions[nr].material_key = cur_key; nr++;
                cur_region_idx += 1; cur_key = items[i].material_key; cur_region_obj = 0; cur_region_inv = cur_inv;
            }
            if (object_limit || dispatch_limit) {
                cur.total_objects = cur_obj; cur.total_invocations = cur_inv; cur.batch_base_invocation = cur_base_inv;
                batches[nb++] = cur;
                cur_base_inv += cur_inv; cur_inv = 0; cur_region_inv = 0; cur_obj = 0;
            }
            r3_object_culling_info* r = &cur.object_culling_information[cur_obj];
            r->invocation_start = cur_inv;
            r->invocation_end = cur_inv + invocation_count;
            r->region_id = cur_region_idx;
            r->object_id = h;
            r->base_region_invocation = cur_region_inv;
            r->local_region_id = cur_region_obj;
            r->previous_global_invocation = (prev_map && h < prev_cap) ? prev_map[h] : R3_NO_PREVIOUS;
            r->atomic_capable = (c->sort_flags[h] & 2) ? 1u : 0u;
            cur_map[h] = cur_inv + cur_base_inv;
            cur_obj += 1; cur_region_obj += 1;
            cur_inv += round_up_u32(invocation_count, R3_WORKGROUP_SIZE);
        }
        regions[nr].job_index = nb; regions[nr].bind_group_index = 0; regions[nr].material_key = cur_key; nr++;
        cur.total_objects = cur_obj; cur.total_invocations = cur_inv; cur.batch_base_invocation = cur_base_inv;
        batches[nb++] = cur;
        cam->batches = batches; cam->n_batches = nb; cam->regions = regions; cam->n_regions = nr;
        uint64_t tot = 0;
        for (uint32_t b = 0; b < nb; ++b) tot += batches[b].total_invocations;
        cam->total_invocations = (uint32_t)tot;
    }
    free(items);
    free(prev_map);
    cam->prev_invocation = cur_map; cam->prev_invocation_cap = cap;   /* set_camera (batching.rs:247) */
    return R3_OK;
}
API int r3o_batch_counts(r3o_ctx* c, uint32_t camera, uint32_t* nb, uint32_t* nr, uint32_t* tot) {
    CAM_OR_FAIL(c, camera);
    if (nb) *nb = cam->n_batches;
    if (nr) *nr = cam->n_regions;
    if (tot) *tot = cam->total_invocations;
    return R3_OK;
}
API int r3o_batching_info(r3o_ctx* c, uint32_t camera, uint32_t info[4]) {
    CAM_OR_FAIL(c, camera);
    if (!info) return fail(c, R3_E_INVALID, "null");
    info[0] = cam->batches ? 2u : 0u; info[1] = 0; info[2] = cam->n_batches; info[3] = cam->n_regions;   /* the oracle batches on the host */
    return R3_OK;
}
API int r3o_forward_light_evaluations(r3o_ctx* c, uint64_t* n) { if (!c || !n) return R3_E_INVALID; *n = 0; return R3_OK; }   /* statistics of the CUDA path only */
API int r3o_frame_begin(r3o_ctx* c) { return c ? R3_OK : R3_E_INVALID; }   /* submission is a property of the CUDA path */
API int r3o_frame_end(r3o_ctx* c) { return c ? R3_OK : R3_E_INVALID; }
API int r3o_frame_graph_stats(r3o_ctx* c, uint64_t stats[4]) { if (!c || !stats) return R3_E_INVALID; stats[0] = stats[1] = stats[2] = stats[3] = 0; return R3_OK; }
API int r3o_set_stage_timing(r3o_ctx* c, int enabled) { (void)enabled; return c ? R3_OK : R3_E_INVALID; }
API int r3o_stage_times(r3o_ctx* c, double ms[8], uint32_t launches[8]) { if (!c || !ms || !launches) return R3_E_INVALID; for (int k = 0; k < 8; ++k) { ms[k] = 0.0; launches[k] = 0; } return R3_OK; }
API int r3o_set_parity_target(r3o_ctx* c, int enabled) { (void)enabled; return c ? R3_OK : R3_E_INVALID; }   /* the oracle always keeps the f32 result */
API int r3o_readback_batches(r3o_ctx* c, uint32_t camera, r3_batch_data* b, r3_region* r) {
    CAM_OR_FAIL(c, camera);
    if (b && cam->n_batches) memcpy(b, cam->batches, (size_t)cam->n_batches * sizeof *b);
    if (r && cam->n_regions) memcpy(r, cam->regions, (size_t)cam->n_regions * sizeof *r);
    return R3_OK;
}

/* ------------------------------------------------------------------ a7: InputOutputBuffer (suballoc.rs:66-222) */
static uint64_t next_pow2_u64(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }
static uint64_t io_capacity(uint64_t in, uint64_t out) { return next_pow2_u64(in > out ? in : out) * 2; }
static uint64_t io_out_off(const r3o_iobuf* b) { return b->flipped ? b->capacity_elements / 2 : 0; }
static uint64_t io_in_off(const r3o_iobuf* b) { return b->flipped ? 0 : b->capacity_elements / 2; }
static void io_new(r3o_iobuf* b, uint64_t elems, uint64_t elem_size, int clear_on_swap) {
    b->capacity_elements = io_capacity(elems, elems);
    b->out_elems = b->in_elems = elems; b->flipped = 0; b->clear_on_swap = clear_on_swap; b->elem_size = elem_size;
    b->data = (uint8_t*)calloc(b->capacity_elements * elem_size + 16, 1); /* wgpu buffers start zeroed */
    b->created = 1;
}
static void io_swap(r3o_iobuf* b, uint64_t new_elems) {
    uint64_t old_out = io_out_off(b);
    b->in_elems = b->out_elems; b->out_elems = new_elems; b->flipped = !b->flipped;
    uint64_t ncap = io_capacity(b->in_elems, b->out_elems);
    if (ncap != b->capacity_elements) {
        uint8_t* nd = (uint8_t*)calloc(ncap * b->elem_size + 16, 1);
        uint64_t old_cap = b->capacity_elements;
        b->capacity_elements = ncap;
        if (!b->clear_on_swap) {
            uint64_t bytes = b->in_elems * b->elem_size, room = (old_cap - old_out) * b->elem_size;
            memcpy(nd + io_in_off(b) * b->elem_size, b->data + old_out * b->elem_size, bytes < room ? bytes : room);
        }
        free(b->data);
        b->data = nd;
    } else if (b->clear_on_swap) {
        memset(b->data, 0, b->capacity_elements * b->elem_size);
    }
}

/* ------------------------------------------------------------------ a8/a9: vertex fetch (vertex_attributes.wgsl:43-85) */
static inline uint32_t mesh_word(const r3o_ctx* c, uint64_t idx) { return idx < c->mesh_words ? c->mesh[idx] : 0u; /* robust access */ }
static inline void fetch_vec2(const r3o_ctx* c, uint32_t byte_off, uint32_t vertex, float* out) {   /* vertex_attributes.wgsl:43-49 */
    uint64_t first = (uint64_t)(byte_off / 4u) + (uint64_t)vertex * 2u;
    out[0] = u2f(mesh_word(c, first)); out[1] = u2f(mesh_word(c, first + 1));
}
static inline v3 fetch_vec3(const r3o_ctx* c, uint32_t byte_off, uint32_t vertex) {
    uint64_t first = (uint64_t)(byte_off / 4u) + (uint64_t)vertex * 3u;
    v3 r = {u2f(mesh_word(c, first)), u2f(mesh_word(c, first + 1)), u2f(mesh_word(c, first + 2))};
    return r;
}

/* textureSampleMin (cull.wgsl:243-262).  Robust-access rule: the mip index and the texel coordinates
 * are clamped into range (the reference leaves out-of-range loads to the driver). */
static float hiz_sample_min(const r3o_ctx* c, float u, float v, uint32_t mip) {
    if (!c->hiz_mips) return 0.0f;
    if (mip >= c->hiz_mips) mip = c->hiz_mips - 1;
    float rw = (float)c->hiz_w[mip], rh = (float)c->hiz_h[mip];
    float px = u * rw - 0.5f, py = v * rh - 0.5f;
    float lx = fmax2(floorf(px), 0.0f), ly = fmax2(floorf(py), 0.0f);
    float hx = fmin2(ceilf(px), rw - 1.0f), hy = fmin2(ceilf(py), rh - 1.0f);
    lx = fmin2(lx, rw - 1.0f); ly = fmin2(ly, rh - 1.0f); hx = fmax2(hx, 0.0f); hy = fmax2(hy, 0.0f);
    if (!(lx == lx)) lx = 0; if (!(ly == ly)) ly = 0; if (!(hx == hx)) hx = 0; if (!(hy == hy)) hy = 0;
    uint32_t x0 = (uint32_t)lx, y0 = (uint32_t)ly, x1 = (uint32_t)hx, y1 = (uint32_t)hy, w = c->hiz_w[mip];
    const float* t = c->hiz[mip];
    float m = t[(size_t)y0 * w + x0];
    m = fmin2(m, t[(size_t)y0 * w + x1]);
    m = fmin2(m, t[(size_t)y1 * w + x0]);
    m = fmin2(m, t[(size_t)y1 * w + x1]);
    return m;
}
/* ceil(log2(max(x, 1))) evaluated exactly on the f32 bit pattern (cull.wgsl:314) */
static uint32_t ceil_log2_f32(float x) {
    if (!(x > 1.0f)) return 0;
    uint32_t b = f2u(x), e = (b >> 23) & 0xFF, m = b & 0x7FFFFFu;
    if (e == 255) return 128;
    return (e - 127) + (m ? 1u : 0u);
}

/* execute_culling (cull.wgsl:264-324) */
static int execute_culling(const r3o_ctx* c, const r3_camera_header* h, const float* mvp, const v3* p) {
    v4 p0 = mat_point(mvp, p[0].x, p[0].y, p[0].z);
    v4 p1 = mat_point(mvp, p[1].x, p[1].y, p[1].z);
    v4 p2 = mat_point(mvp, p[2].x, p[2].y, p[2].z);
    /* determinant(mat3x3(p0.xyw, p1.xyw, p2.xyw)), cofactor expansion along the first row of columns */
    float det = (p0.x * (p1.y * p2.w - p2.y * p1.w) - p1.x * (p0.y * p2.w - p2.y * p0.w)) + p2.x * (p0.y * p1.w - p1.y * p0.w);
    if ((h->flags & R3_PCU_POSITIVE_AREA_VISIBLE) && det <= 0.0f) return 0;
    if (!(h->flags & R3_PCU_POSITIVE_AREA_VISIBLE) && det >= 0.0f) return 0;
    float n0x = p0.x / p0.w, n0y = p0.y / p0.w, n0z = p0.z / p0.w;
    float n1x = p1.x / p1.w, n1y = p1.y / p1.w, n1z = p1.z / p1.w;
    float n2x = p2.x / p2.w, n2y = p2.y / p2.w, n2z = p2.z / p2.w;
    float minx = fmin2(n0x, fmin2(n1x, n2x)), miny = fmin2(n0y, fmin2(n1y, n2y));
    float maxx = fmax2(n0x, fmax2(n1x, n2x)), maxy = fmax2(n0y, fmax2(n1y, n2y));
    float hrx = h->resolution[0] / 2.0f, hry = h->resolution[1] / 2.0f;
    float minsx = (minx + 1.0f) * hrx, minsy = (miny + 1.0f) * hry;
    float maxsx = (maxx + 1.0f) * hrx, maxsy = (maxy + 1.0f) * hry;
    if (!(h->flags & R3_PCU_MULTISAMPLED)) {
        /* WGSL round() = ties to even = rintf in the default rounding mode */
        if (rintf(minsx) == rintf(maxsx) || rintf(minsy) == rintf(maxsy)) return 0;
    }
    if (h->shadow_index != R3_CAMERA_VIEWPORT) return 1;
    float mintx = (minx + 1.0f) / 2.0f, minty = 1.0f - (miny + 1.0f) / 2.0f;
    float maxtx = (maxx + 1.0f) / 2.0f, maxty = 1.0f - (maxy + 1.0f) / 2.0f;
    float u = (maxtx + mintx) / 2.0f, v = (maxty + minty) / 2.0f;
    float ex = maxsx - minsx, ey = maxsy - minsy;
    uint32_t mip = ceil_log2_f32(fmax2(fmax2(ex, ey), 1.0f));
    float depth = fmax2(fmax2(n0z, n1z), n2z);
    float occl = hiz_sample_min(c, u, v, mip);
    if (depth < occl) return 0;
    return 1;
}

/* ------------------------------------------------------------------ a8: GpuCuller::cull + cull.wgsl::cs_main
 * Host part culler.rs:531-659 (buffer swap, draw-call clear, one dispatch per batch); device part
 * cull.wgsl:326-390 executed in invocation order, which is one legal outcome of the reference's
 * atomic appends and makes the lists reproducible. */
API int r3o_cull(r3o_ctx* c, uint32_t camera, const r3_batch_data* batches, uint32_t n_batches, const r3_region* regions,
                 uint32_t n_regions) {
    CAM_OR_FAIL(c, camera);
    if (!cam->header_set) return fail(c, R3_E_STATE, "cull before object_uniform_upload");
    if (batches) {
        if (!regions) return fail(c, R3_E_INVALID, "cull: batches without regions");
        r3_batch_data* nb = (r3_batch_data*)malloc((size_t)n_batches * sizeof *nb + 16);
        r3_region* nr = (r3_region*)malloc((size_t)n_regions * sizeof *nr + 16);
        memcpy(nb, batches, (size_t)n_batches * sizeof *nb);
        memcpy(nr, regions, (size_t)n_regions * sizeof *nr);
        free(cam->batches); free(cam->regions);
        cam->batches = nb; cam->regions = nr; cam->n_batches = n_batches; cam->n_regions = n_regions;
        uint64_t tot = 0;
        for (uint32_t b = 0; b < n_batches; ++b) tot += nb[b].total_invocations;
        cam->total_invocations = (uint32_t)tot;
    }
    if (cam->n_batches == 0) { cam->has_draw_call_set = 0; return R3_OK; } /* add_culling_to_graph returns early (culler.rs:705-707) */
    uint64_t inv = cam->total_invocations, words = (inv + 31) / 32;
    if (!cam->index_buffer.created) {                     /* CullingBuffers::new (culler.rs:96-112) */
        io_new(&cam->index_buffer, inv * 3, 4, 0);
        io_new(&cam->draw_call_buffer, cam->n_regions, 20, 1);
        io_new(&cam->results_buffer, words, 4, 0);
    } else {                                              /* update_sizes (culler.rs:114-124) */
        io_swap(&cam->index_buffer, inv * 3);
        io_swap(&cam->draw_call_buffer, cam->n_regions);
        io_swap(&cam->results_buffer, words);
    }
    memset(cam->draw_call_buffer.data, 0, cam->draw_call_buffer.capacity_elements * 20); /* clear_buffer(.., 8, None) culler.rs:642 */

    uint32_t* idx_pred = (uint32_t*)cam->index_buffer.data + io_out_off(&cam->index_buffer);
    uint32_t* idx_resid = (uint32_t*)cam->index_buffer.data + io_in_off(&cam->index_buffer);
    r3_indirect_call* dc_pred = (r3_indirect_call*)cam->draw_call_buffer.data + io_out_off(&cam->draw_call_buffer);
    r3_indirect_call* dc_resid = (r3_indirect_call*)cam->draw_call_buffer.data + io_in_off(&cam->draw_call_buffer);
    uint32_t* res_out = (uint32_t*)cam->results_buffer.data + io_out_off(&cam->results_buffer);
    const uint32_t* res_in = (const uint32_t*)cam->results_buffer.data + io_in_off(&cam->results_buffer);
    uint64_t res_in_words = cam->results_buffer.capacity_elements / 2;
    int shadow = cam->header.shadow_index != R3_CAMERA_VIEWPORT;

    for (uint32_t b = 0; b < cam->n_batches; ++b) {
        const r3_batch_data* job = &cam->batches[b];
        for (uint32_t o = 0; o < job->total_objects; ++o) {
            const r3_object_culling_info* info = &job->object_culling_information[o];
            const r3_object* obj = &c->objects[info->object_id];
            const float* mvp = cam->matrices[info->object_id].model_view_proj;
            uint32_t padded_end = info->invocation_start + round_up_u32(info->invocation_end - info->invocation_start, R3_WORKGROUP_SIZE);
            /* result bits of every workgroup this object spans are written by lane 0 (cull.wgsl:229-241) */
            for (uint32_t g = info->invocation_start; g < padded_end; g += 32)
                res_out[((uint64_t)job->batch_base_invocation + g) / 32] = 0;
            for (uint32_t gid = info->invocation_start; gid < padded_end; ++gid) {
                uint32_t global_invocation = job->batch_base_invocation + gid;
                if (gid >= info->invocation_end) {                                      /* cull.wgsl:343-347 */
                    if (info->atomic_capable == 0) {
                        dc_resid[info->region_id].vertex_count += 3;
                        idx_resid[(uint64_t)global_invocation * 3] = R3_INVALID_VERTEX;
                        idx_resid[(uint64_t)global_invocation * 3 + 1] = R3_INVALID_VERTEX;
                        idx_resid[(uint64_t)global_invocation * 3 + 2] = R3_INVALID_VERTEX;
                    }
                    continue;
                }
                uint32_t object_invocation = gid - info->invocation_start;
                if (info->local_region_id == 0 && object_invocation == 0) {             /* init_draw_calls cull.wgsl:47-61 */
                    dc_pred[info->region_id].vertex_offset = 0; dc_pred[info->region_id].instance_count = 1;
                    dc_pred[info->region_id].base_instance = 0; dc_pred[info->region_id].base_index = global_invocation * 3u;
                    dc_resid[info->region_id].vertex_offset = 0; dc_resid[info->region_id].instance_count = 1;
                    dc_resid[info->region_id].base_instance = 0; dc_resid[info->region_id].base_index = global_invocation * 3u;
                }
                uint32_t i0 = mesh_word(c, (uint64_t)obj->first_index + object_invocation * 3u);     /* vertex_fetch cull.wgsl:9-32 */
                uint32_t i1 = mesh_word(c, (uint64_t)obj->first_index + object_invocation * 3u + 1u);
                uint32_t i2 = mesh_word(c, (uint64_t)obj->first_index + object_invocation * 3u + 2u);
                v3 p[3] = {fetch_vec3(c, obj->attr_offset[0], i0), fetch_vec3(c, obj->attr_offset[0], i1), fetch_vec3(c, obj->attr_offset[0], i2)};
                int passes = execute_culling(c, &cam->header, mvp, p);
                uint32_t pk0 = (o << 24) | (i0 & 0xFFFFFFu), pk1 = (o << 24) | (i1 & 0xFFFFFFu), pk2 = (o << 24) | (i2 & 0xFFFFFFu);
                if (info->atomic_capable == 1) {
                    if (passes) {
                        uint32_t slot = dc_pred[info->region_id].vertex_count / 3u;                  /* cull.wgsl:63-67,84-99 */
                        dc_pred[info->region_id].vertex_count += 3;
                        uint64_t gi = (uint64_t)slot + info->base_region_invocation + job->batch_base_invocation;
                        idx_pred[gi * 3] = pk0; idx_pred[gi * 3 + 1] = pk1; idx_pred[gi * 3 + 2] = pk2;
                        if (!shadow) {
                            int prev = 0;                                                            /* cull.wgsl:152-160 */
                            if (info->previous_global_invocation != R3_NO_PREVIOUS) {
                                uint64_t pgi = (uint64_t)object_invocation + info->previous_global_invocation;
                                uint32_t mask = (pgi / 32 < res_in_words) ? res_in[pgi / 32] : 0u;
                                prev = (mask >> (pgi % 32)) & 1u;
                            }
                            if (!prev) {                                                             /* cull.wgsl:101-116 */
                                uint32_t rslot = dc_resid[info->region_id].vertex_count / 3u;
                                dc_resid[info->region_id].vertex_count += 3;
                                uint64_t rgi = (uint64_t)rslot + info->base_region_invocation + job->batch_base_invocation;
                                idx_resid[rgi * 3] = pk0; idx_resid[rgi * 3 + 1] = pk1; idx_resid[rgi * 3 + 2] = pk2;
                            }
                        }
                    }
                } else {                                                                             /* cull.wgsl:374-380 */
                    dc_resid[info->region_id].vertex_count += 3;
                    uint64_t gi = (uint64_t)global_invocation * 3;
                    if (passes) { idx_resid[gi] = pk0; idx_resid[gi + 1] = pk1; idx_resid[gi + 2] = pk2; }
                    else { idx_resid[gi] = R3_INVALID_VERTEX; idx_resid[gi + 1] = R3_INVALID_VERTEX; idx_resid[gi + 2] = R3_INVALID_VERTEX; }
                }
                if (passes) res_out[global_invocation / 32] |= 1u << (global_invocation % 32);       /* cull.wgsl:225-227 */
            }
        }
    }
    cam->has_draw_call_set = 1;
    return R3_OK;
}

static int io_read(r3o_ctx* c, const r3o_iobuf* b, int partition, void* out, uint64_t cap, uint64_t* count) {
    if (!b->created) { if (count) *count = 0; return R3_OK; }
    uint64_t elems = partition ? b->in_elems : b->out_elems, off = partition ? io_in_off(b) : io_out_off(b);
    uint64_t room = b->capacity_elements / 2;
    if (elems > room) elems = room;
    if (count) *count = elems;
    if (out) {
        if (cap < elems) return fail(c, R3_E_INVALID, "readback: capacity too small");
        memcpy(out, b->data + off * b->elem_size, elems * b->elem_size);
    }
    return R3_OK;
}
API int r3o_readback_indices(r3o_ctx* c, uint32_t camera, int partition, uint32_t* out, uint64_t cap, uint64_t* count) {
    CAM_OR_FAIL(c, camera);
    return io_read(c, &cam->index_buffer, partition, out, cap, count);
}
API int r3o_readback_draw_calls(r3o_ctx* c, uint32_t camera, int partition, r3_indirect_call* out, uint32_t cap, uint32_t* count) {
    CAM_OR_FAIL(c, camera);
    uint64_t n = 0;
    int rc = io_read(c, &cam->draw_call_buffer, partition, out, cap, &n);
    if (count) *count = (uint32_t)n;
    return rc;
}
API int r3o_readback_culling_results(r3o_ctx* c, uint32_t camera, int partition, uint32_t* out, uint64_t cap, uint64_t* count) {
    CAM_OR_FAIL(c, camera);
    return io_read(c, &cam->results_buffer, partition, out, cap, count);
}

#include "r3_oracle_forward.inc"
