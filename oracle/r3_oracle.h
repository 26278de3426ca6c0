/* r3_oracle.h — internal declarations of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The oracle is a plain-C restatement of the reference's arithmetic for the hot path.  It exists
 * so tests/ can check the CUDA path; nothing in rend3_b200/ may include, link or call it.
 * It exports the same entry points as include/rend3_b200.h with the prefix r3o_ instead of r3_.
 */
#ifndef R3_ORACLE_H
#define R3_ORACLE_H

#include <stdint.h>
#include "../include/r3_layouts.h"

#define R3O_MAX_CAMERAS 64
#define R3O_MAX_THREADS 1024

/* InputOutputBuffer — rend3-routine/src/culling/suballoc.rs:17-223 (header kept out of `data`) */
typedef struct {
    uint8_t* data;
    uint64_t capacity_elements, out_elems, in_elems, elem_size;
    int flipped, clear_on_swap, created;
} r3o_iobuf;

typedef struct {
    int header_set;
    r3_camera_header header;
    r3_object_matrices* matrices; uint32_t matrices_cap;
    uint32_t* visible; uint32_t visible_count, visible_cap;
    uint8_t* flag; uint32_t flag_cap;          /* per-slot survivor flags of the last cull (scratch kept across calls) */
    /* batch_objects products (this frame) and the cached DrawCallSet of the previous frame (forward.rs:219) */
    r3_batch_data* batches; uint32_t n_batches;
    r3_region* regions; uint32_t n_regions; uint32_t total_invocations;
    r3_batch_data* prev_batches; uint32_t prev_n_batches;
    r3_region* prev_regions; uint32_t prev_n_regions; uint32_t prev_total_invocations;
    int has_draw_call_set, has_prev_draw_call_set;
    uint32_t* prev_invocation; uint32_t prev_invocation_cap;   /* PerCameraPreviousInvocationsMap */
    r3o_iobuf index_buffer, draw_call_buffer, results_buffer;  /* CullingBuffers (culler.rs:88-125) */
} r3o_camera;

typedef struct r3o_ctx {
    char err[256];
    r3_object* objects; uint32_t n_slots;
    uint64_t* sort_key; uint8_t* sort_flags; float* sort_loc; uint32_t sort_n;
    uint32_t* mesh; uint64_t mesh_words;
    r3_material* materials; uint32_t n_materials;
    r3_texture_desc* tex_descs; uint32_t n_textures; uint8_t* texels; uint64_t texel_bytes;   /* bindless d2 texture table */
    int has_skybox; r3_texture_desc sky_desc; uint8_t* sky_texels;                          /* cube map of the skybox routine */
    r3_directional_light* dir_lights; uint32_t n_dir; uint32_t atlas_w, atlas_h;
    r3_point_light* point_lights; uint32_t n_point;
    r3_frame_uniforms uniforms;
    r3o_camera cams[R3O_MAX_CAMERAS];
    /* render targets */
    uint32_t width, height, samples; float clear_color[4];
    uint32_t row_begin, row_end;
    uint64_t* vis;            /* per pixel (depth bits << 32) | (pass << 31) | triangle record */
    float* hdr; uint16_t* hdr16; float* depth; uint8_t* ldr;
    float* atlas;
    float** hiz; uint32_t* hiz_w; uint32_t* hiz_h; uint32_t hiz_mips;
    struct r3o_tri* tris[4]; uint64_t n_tris[4], cap_tris[4];   /* predicted, residual, blend, shadow-pass scratch */
    uint16_t* sample_col16;   /* the rgba16f colour target per SAMPLE (what the blend routine reads and writes) */
    uint8_t* blended;         /* per pixel: touched by the blend routine this frame */
    uint64_t stats[4];
} r3o_ctx;

#endif
