/* r3_layouts.h — byte-exact std430 records shared by rend3's Rust managers and this library.
 *
 * These are the buffers the reference's managers already upload to wgpu; the C ABI in
 * rend3_b200.h consumes the very same bytes, so the Rust side needs no repacking.
 * Every struct cites the reference definition it mirrors (paths relative to the reference
 * repository root) and is checked with static_assert against the encase/std430 offsets.
 *
 * Plain C99 / C++11 / CUDA compatible.  No pointers, no torch types.
 */
#ifndef R3_LAYOUTS_H
#define R3_LAYOUTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define R3_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define R3_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#define R3_INVALID_VERTEX 0x00FFFFFFu        /* rend3/shaders/vertex_attributes.wgsl:15 */
#define R3_CAMERA_VIEWPORT 0xFFFFFFFFu       /* CameraSpecifier::Viewport -> shadow_index u32::MAX (culler.rs:166-169) */
#define R3_ATTR_ABSENT 0xFFFFFFFFu           /* missing vertex attribute (managers/object.rs:263) */
#define R3_BATCH_SIZE 256u                   /* rend3-routine/src/culling/mod.rs:1 */
#define R3_WORKGROUP_SIZE 256u               /* rend3-routine/src/culling/mod.rs:2 */
#define R3_NO_PREVIOUS 0xFFFFFFFFu           /* batching.rs:226 */

/* PerCameraUniform.flags (culler.rs:151-156, structures.wgsl:64-72) */
#define R3_PCU_POSITIVE_AREA_VISIBLE 0x1u
#define R3_PCU_MULTISAMPLED 0x2u

/* PbrMaterial flags (rend3-routine/shaders/src/material.wgsl:1-15) */
#define R3_MAT_ALBEDO_ACTIVE 0x0001u
#define R3_MAT_ALBEDO_BLEND 0x0002u
#define R3_MAT_ALBEDO_VERTEX_SRGB 0x0004u
#define R3_MAT_BICOMPONENT_NORMAL 0x0008u
#define R3_MAT_SWIZZLED_NORMAL 0x0010u
#define R3_MAT_YDOWN_NORMAL 0x0020u
#define R3_MAT_AOMR_COMBINED 0x0040u
#define R3_MAT_AOMR_SWIZZLED_SPLIT 0x0080u
#define R3_MAT_AOMR_SPLIT 0x0100u
#define R3_MAT_AOMR_BW_SPLIT 0x0200u
#define R3_MAT_CC_GLTF_COMBINED 0x0400u
#define R3_MAT_CC_GLTF_SPLIT 0x0800u
#define R3_MAT_CC_BW_SPLIT 0x1000u
#define R3_MAT_UNLIT 0x2000u
#define R3_MAT_NEAREST 0x4000u

/* vertex attribute slot order of PbrMaterial::supported_attributes (pbr/material.rs:486-495) */
enum { R3_ATTR_POSITION = 0, R3_ATTR_NORMAL, R3_ATTR_TANGENT, R3_ATTR_UV0, R3_ATTR_UV1, R3_ATTR_COLOR0, R3_ATTR_COUNT };

/* ShaderObject<PbrMaterial> — rend3/src/managers/object.rs:23-36, structures_object.wgsl:3-12. stride 128 */
typedef struct r3_object {
    float transform[16];          /* @0   model->world, column major */
    float sphere_center[3];       /* @64  world-space bounding sphere (object.rs:269) */
    float sphere_radius;          /* @76 */
    uint32_t first_index;         /* @80  word index into the mesh buffer (object.rs:279) */
    uint32_t index_count;         /* @84 */
    uint32_t material_index;      /* @88 */
    uint32_t attr_offset[6];      /* @92  byte offsets into the mesh buffer, R3_ATTR_ABSENT if missing */
    uint32_t enabled;             /* @116 */
    uint32_t _pad[2];             /* @120 */
} r3_object;
R3_STATIC_ASSERT(sizeof(r3_object) == 128, "Object stride");
R3_STATIC_ASSERT(offsetof(r3_object, sphere_center) == 64, "sphere");
R3_STATIC_ASSERT(offsetof(r3_object, first_index) == 80, "first_index");
R3_STATIC_ASSERT(offsetof(r3_object, attr_offset) == 92, "attr offsets");
R3_STATIC_ASSERT(offsetof(r3_object, enabled) == 116, "enabled");

/* PerCameraUniform header — rend3-routine/src/culling/culler.rs:158-175, structures.wgsl:47-62.
 * `objects[]` (r3_object_matrices, stride 128) follows at byte 240. */
typedef struct r3_camera_header {
    float view[16];               /* @0 */
    float view_proj[16];          /* @64 */
    uint32_t shadow_index;        /* @128 R3_CAMERA_VIEWPORT for the viewport camera */
    uint32_t _pad0[3];
    float frustum[5][4];          /* @144 left,right,top,bottom,near: (abc, d) normalised (util/frustum.rs:96-145) */
    float resolution[2];          /* @224 */
    uint32_t flags;               /* @232 R3_PCU_* */
    uint32_t object_count;        /* @236 capacity of the object buffer (culler.rs:446-447) */
} r3_camera_header;
R3_STATIC_ASSERT(sizeof(r3_camera_header) == 240, "PerCameraUniform header");
R3_STATIC_ASSERT(offsetof(r3_camera_header, frustum) == 144, "frustum");
R3_STATIC_ASSERT(offsetof(r3_camera_header, resolution) == 224, "resolution");
R3_STATIC_ASSERT(offsetof(r3_camera_header, object_count) == 236, "object_count");

/* PerCameraUniformObjectData — culler.rs:177-183 */
typedef struct r3_object_matrices {
    float model_view[16];
    float model_view_proj[16];
} r3_object_matrices;
R3_STATIC_ASSERT(sizeof(r3_object_matrices) == 128, "PerCameraUniformObjectData");

/* ShaderObjectCullingInformation — culling/batching.rs:90-100 */
typedef struct r3_object_culling_info {
    uint32_t invocation_start;
    uint32_t invocation_end;
    uint32_t object_id;
    uint32_t region_id;
    uint32_t base_region_invocation;
    uint32_t local_region_id;
    uint32_t previous_global_invocation;
    uint32_t atomic_capable;
} r3_object_culling_info;
R3_STATIC_ASSERT(sizeof(r3_object_culling_info) == 32, "ObjectCullingInformation");

/* ShaderBatchData — culling/batching.rs:81-88 (#[align(256)] => 8448 bytes) */
typedef struct r3_batch_data {
    uint32_t total_objects;
    uint32_t total_invocations;
    uint32_t batch_base_invocation;
    r3_object_culling_info object_culling_information[256];
    uint32_t _pad[61];
} r3_batch_data;
R3_STATIC_ASSERT(sizeof(r3_batch_data) == 8448, "ShaderBatchData");
R3_STATIC_ASSERT(offsetof(r3_batch_data, object_culling_information) == 12, "batch table");

/* JobSubRegion + ShaderJobKey — culling/batching.rs:22-32 (host side only) */
typedef struct r3_region {
    uint32_t job_index;           /* batch this region's draw call binds (DrawCall::batch_index) */
    uint32_t bind_group_index;    /* TextureBindGroupIndex (DUMMY = 0 in the GpuDriven profile) */
    uint64_t material_key;
} r3_region;
R3_STATIC_ASSERT(sizeof(r3_region) == 16, "region");

/* IndirectCall — structures.wgsl:20-26 (20 bytes) */
typedef struct r3_indirect_call {
    uint32_t vertex_count;
    uint32_t instance_count;
    uint32_t base_index;
    int32_t vertex_offset;
    uint32_t base_instance;
} r3_indirect_call;
R3_STATIC_ASSERT(sizeof(r3_indirect_call) == 20, "IndirectCall");

/* FrameUniforms — rend3-routine/src/uniforms.rs:16-27, structures.wgsl:28-38 (uniform buffer, 496 B) */
typedef struct r3_frame_uniforms {
    float view[16];
    float view_proj[16];
    float origin_view_proj[16];
    float inv_view[16];
    float inv_view_proj[16];
    float inv_origin_view_proj[16];
    float frustum[5][4];          /* @384 */
    float ambient[4];             /* @464 */
    uint32_t resolution[2];       /* @480 */
    uint32_t _pad[2];
} r3_frame_uniforms;
R3_STATIC_ASSERT(sizeof(r3_frame_uniforms) == 496, "FrameUniforms");
R3_STATIC_ASSERT(offsetof(r3_frame_uniforms, ambient) == 464, "ambient");

/* ShaderDirectionalLight — rend3/src/managers/directional.rs:38-53; buffer = u32 count @0, array @16 */
typedef struct r3_directional_light {
    float view_proj[16];          /* @0 */
    float color[3];               /* @64 color*intensity */
    float _pad0;
    float direction[3];           /* @80 un-normalised (directional.rs:145) */
    float _pad1;
    float inv_resolution[2];      /* @96 1/atlas size */
    float atlas_offset[2];        /* @104 */
    float atlas_size[2];          /* @112 */
    float _pad2[2];
} r3_directional_light;
R3_STATIC_ASSERT(sizeof(r3_directional_light) == 128, "DirectionalLight");
R3_STATIC_ASSERT(offsetof(r3_directional_light, atlas_offset) == 104, "atlas_offset");

/* ShaderPointLight — rend3/src/managers/point.rs:21-26; buffer = u32 count @0, array @16 */
typedef struct r3_point_light {
    float position[4];
    float color[3];
    float radius;
} r3_point_light;
R3_STATIC_ASSERT(sizeof(r3_point_light) == 32, "PointLight");

/* GpuPoweredShaderWrapper<PbrMaterial> — managers/material.rs:25-29 + pbr/material.rs:526-543,
 * material.wgsl:21-57 (208 bytes) */
typedef struct r3_material {
    uint32_t textures[10];        /* @0 0 = none */
    uint32_t _pad0[2];
    float uv_transform0[3][4];    /* @48 mat3x3 columns padded to vec4 */
    float uv_transform1[3][4];    /* @96 */
    float albedo[4];              /* @144 */
    float emissive[3];            /* @160 */
    float roughness;              /* @172 */
    float metallic;               /* @176 */
    float reflectance;            /* @180 */
    float clear_coat;             /* @184 */
    float clear_coat_roughness;   /* @188 */
    float anisotropy;             /* @192 */
    float ambient_occlusion;      /* @196 */
    float alpha_cutout;           /* @200 */
    uint32_t flags;               /* @204 */
} r3_material;
R3_STATIC_ASSERT(sizeof(r3_material) == 208, "GpuMaterialData");
R3_STATIC_ASSERT(offsetof(r3_material, albedo) == 144, "albedo");
R3_STATIC_ASSERT(offsetof(r3_material, flags) == 204, "flags");

/* texture slots of r3_material.textures[] (material.wgsl:21-35): value = index into the texture table + 1, 0 = none */
enum { R3_TEX_ALBEDO = 0, R3_TEX_NORMAL, R3_TEX_ROUGHNESS, R3_TEX_METALLIC, R3_TEX_REFLECTANCE, R3_TEX_CLEAR_COAT, R3_TEX_CLEAR_COAT_ROUGHNESS,
       R3_TEX_EMISSIVE, R3_TEX_ANISOTROPY, R3_TEX_AMBIENT_OCCLUSION };

/* One entry of the bindless `textures` array (TextureManager<D2>, rend3/src/managers/texture.rs; binding opaque.wgsl:35):
 * a 2D texture with `mip_count` levels stored tightly one after the other (level l is max(w >> l, 1) x max(h >> l, 1))
 * starting at `byte_offset` of the texel blob handed to r3_set_textures. */
#define R3_TEXFMT_RGBA8_UNORM 0u
#define R3_TEXFMT_RGBA8_UNORM_SRGB 1u   /* rgb decoded to linear before filtering, alpha linear */
#define R3_TEXFMT_RGBA32_FLOAT 2u
#define R3_TEXFMT_R8_UNORM 3u           /* single channel (split AO / metallic / roughness maps): (r, 0, 0, 1) */
#define R3_TEXFMT_RG8_UNORM 4u          /* two channels (bicomponent normal maps): (r, g, 0, 1) */
/* Block-compressed formats (what rend3-gltf's ktx2 / dds loaders hand to add_texture_2d, rend3-gltf/src/lib.rs:1300-1335, 1455-1476,
 * 1556-1602): 4x4-texel blocks, row-major, level l stores ceil(w_l / 4) x ceil(h_l / 4) blocks of 8 (BC1, BC4) or 16 bytes.  The decode is
 * rule R11 of the oracle (oracle/r3_oracle_forward.inc): the ideal palette of the format as ONE IEEE division of two exact integers per
 * channel, e.g. BC1 code 2 red = (2 r0 + r1) / 93 with the 5-bit endpoints r0, r1.  BC7 is integer-exact by its specification (8-bit texels
 * from the interpolation ((64 - w) e0 + w e1 + 32) >> 6, then / 255); a block of the reserved mode reads (0, 0, 0, 0).  BC6H is not
 * implemented (r3_set_textures rejects it like any unknown format). */
#define R3_TEXFMT_BC1_RGBA_UNORM 5u
#define R3_TEXFMT_BC1_RGBA_UNORM_SRGB 6u
#define R3_TEXFMT_BC2_RGBA_UNORM 7u
#define R3_TEXFMT_BC2_RGBA_UNORM_SRGB 8u
#define R3_TEXFMT_BC3_RGBA_UNORM 9u
#define R3_TEXFMT_BC3_RGBA_UNORM_SRGB 10u
#define R3_TEXFMT_BC4_R_UNORM 11u        /* (r, 0, 0, 1) */
#define R3_TEXFMT_BC4_R_SNORM 12u
#define R3_TEXFMT_BC5_RG_UNORM 13u       /* (r, g, 0, 1) */
#define R3_TEXFMT_BC5_RG_SNORM 14u
#define R3_TEXFMT_BC7_RGBA_UNORM 15u
#define R3_TEXFMT_BC7_RGBA_UNORM_SRGB 16u
/* The other filterable uncompressed formats the ktx2 path can hand over (rend3-gltf/src/lib.rs:1195-1285).  Missing channels read (0, 0, 1);
 * snorm: max(v, -127) / 127; unorm16: v / 65535; Rgb10a2: 10-bit channels / 1023 from bit 0, alpha / 3; Bgra8: the bytes are b, g, r, a.
 * The integer (Uint / Sint) formats cannot be bound to a filtering sampler and are rejected. */
#define R3_TEXFMT_R8_SNORM 17u
#define R3_TEXFMT_RG8_SNORM 18u
#define R3_TEXFMT_RGBA8_SNORM 19u
#define R3_TEXFMT_BGRA8_UNORM 20u
#define R3_TEXFMT_BGRA8_UNORM_SRGB 21u
#define R3_TEXFMT_RGB10A2_UNORM 22u
#define R3_TEXFMT_R16_FLOAT 23u
#define R3_TEXFMT_RG16_FLOAT 24u
#define R3_TEXFMT_RGBA16_FLOAT 25u
#define R3_TEXFMT_R32_FLOAT 26u
#define R3_TEXFMT_RG32_FLOAT 27u
#define R3_TEXFMT_R16_UNORM 28u
#define R3_TEXFMT_RG16_UNORM 29u
#define R3_TEXFMT_RGBA16_UNORM 30u
#define R3_TEXFMT_COUNT 31u
#define R3_TEXFMT_IS_BLOCK(f) ((f) >= R3_TEXFMT_BC1_RGBA_UNORM && (f) <= R3_TEXFMT_BC7_RGBA_UNORM_SRGB)
#define R3_TEXFMT_BLOCK_BYTES(f) (((f) <= R3_TEXFMT_BC1_RGBA_UNORM_SRGB || (f) == R3_TEXFMT_BC4_R_UNORM || (f) == R3_TEXFMT_BC4_R_SNORM) ? 8u : 16u)
/* bytes per texel of an uncompressed format */
#define R3_TEXFMT_BPP(f) \
    ((f) == R3_TEXFMT_RGBA32_FLOAT ? 16u : \
     ((f) == R3_TEXFMT_RGBA16_FLOAT || (f) == R3_TEXFMT_RG32_FLOAT || (f) == R3_TEXFMT_RGBA16_UNORM) ? 8u : \
     ((f) == R3_TEXFMT_RG8_UNORM || (f) == R3_TEXFMT_RG8_SNORM || (f) == R3_TEXFMT_R16_FLOAT || (f) == R3_TEXFMT_R16_UNORM) ? 2u : \
     ((f) == R3_TEXFMT_R8_UNORM || (f) == R3_TEXFMT_R8_SNORM) ? 1u : 4u)
/* bytes of one w x h level */
#define R3_TEXFMT_LEVEL_BYTES(f, w, h) \
    (R3_TEXFMT_IS_BLOCK(f) ? (uint64_t)(((w) + 3u) / 4u) * (((h) + 3u) / 4u) * R3_TEXFMT_BLOCK_BYTES(f) : (uint64_t)(w) * (h) * R3_TEXFMT_BPP(f))
typedef struct r3_texture_desc {
    uint32_t width, height, mip_count, format;
    uint64_t byte_offset;
    uint64_t _reserved;
} r3_texture_desc;
R3_STATIC_ASSERT(sizeof(r3_texture_desc) == 32, "r3_texture_desc");

/* GpuSkinningInput — rend3-routine/src/skinning.rs:20-45, skinning.wgsl:3-26 (40 bytes, byte offsets into the mesh buffer,
 * R3_ATTR_ABSENT when an attribute is missing) */
typedef struct r3_skinning_input {
    uint32_t base_position_offset;
    uint32_t base_normal_offset;
    uint32_t base_tangent_offset;
    uint32_t joint_indices_offset;    /* [u16; 4] per vertex */
    uint32_t joint_weight_offset;     /* vec4<f32> per vertex */
    uint32_t updated_position_offset;
    uint32_t updated_normal_offset;
    uint32_t updated_tangent_offset;
    uint32_t joint_matrix_base_offset;
    uint32_t vertex_count;
} r3_skinning_input;
R3_STATIC_ASSERT(sizeof(r3_skinning_input) == 40, "GpuSkinningInput");

#endif /* R3_LAYOUTS_H */
