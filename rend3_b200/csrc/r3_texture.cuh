// r3_texture.cuh — textureSampleGrad on the bindless d2 texture table (r3_set_textures), shared by the shading kernels
// (r3_shade.cu) and the per-fragment alpha cutout of the rasteriser (r3_raster.cu).
#ifndef R3_TEXTURE_CUH
#define R3_TEXTURE_CUH
#include <cuda_fp16.h>

#include "r3_common.cuh"
#define R3_BC7_TABLE __device__ const
#include "../../include/r3_bc7_tables.h"

namespace {

struct TexTable { const r3_texture_desc* tex; uint32_t n_tex; const uint8_t* texels; uint32_t clamp_to_edge; };   // clamp_to_edge: cube-map faces (skybox); 0 = Repeat

__device__ __forceinline__ float srgb_to_linear(float e) { return e > 0.04045f ? powf((e + 0.055f) / 1.055f, 2.4f) : e / 12.92f; }   // math/color.wgsl:3-9

// ------------------------------------------------------------------ material textures (rule R9 of the oracle)
// textureSampleGrad with the linear / nearest Repeat sampler of common/samplers.rs:42-56.  Everything that SELECTS texels or
// levels (coordinates, floor) follows the oracle's order without contraction; the filter weights are continuous.
// Block-compressed texels, rule R11 (include/r3_layouts.h, oracle/r3_oracle_forward.inc): every palette entry is ONE correctly rounded
// division of two exact integers, so a texel has the same bit pattern as in the oracle.  The block is read with one or two 8-byte loads.
__device__ __forceinline__ float3 bc_colour(uint2 blk, uint32_t t, bool bc1, float& alpha) {
    const uint32_t c0 = blk.x & 0xFFFFu, c1 = blk.x >> 16, code = (blk.y >> (2u * t)) & 3u;
    const int r0 = (int)(c0 >> 11), g0 = (int)((c0 >> 5) & 63u), b0 = (int)(c0 & 31u), r1 = (int)(c1 >> 11), g1 = (int)((c1 >> 5) & 63u), b1 = (int)(c1 & 31u);
    const bool four = !bc1 || c0 > c1;
    int wa = 1, wb = 0, den = 1;                                     // numerator wa * e0 + wb * e1, denominator den * full
    if (code == 1u) { wa = 0; wb = 1; }
    else if (code == 2u) { if (four) { wa = 2; wb = 1; den = 3; } else { wa = 1; wb = 1; den = 2; } }
    else if (code == 3u) { if (four) { wa = 1; wb = 2; den = 3; } else { wa = 0; wb = 0; alpha = 0.0f; } }
    const float d5 = (float)(den * 31), d6 = (float)(den * 63);
    return make_float3(div_rn((float)(wa * r0 + wb * r1), d5), div_rn((float)(wa * g0 + wb * g1), d6), div_rn((float)(wa * b0 + wb * b1), d5));
}
__device__ __forceinline__ float bc_channel(uint2 blk, uint32_t t, bool snorm) {
    const unsigned long long bits = (((unsigned long long)blk.y << 32) | blk.x) >> 16;
    const int code = (int)((bits >> (3u * t)) & 7ull);
    int r0 = snorm ? (int)(signed char)(blk.x & 0xFFu) : (int)(blk.x & 0xFFu), r1 = snorm ? (int)(signed char)((blk.x >> 8) & 0xFFu) : (int)((blk.x >> 8) & 0xFFu);
    const bool wide = r0 > r1;
    if (snorm) { r0 = max(r0, -127); r1 = max(r1, -127); }
    const int full = snorm ? 127 : 255;
    if (code == 0) return div_rn((float)r0, (float)full);
    if (code == 1) return div_rn((float)r1, (float)full);
    if (wide) return div_rn((float)((8 - code) * r0 + (code - 1) * r1), (float)(7 * full));
    if (code < 6) return div_rn((float)((6 - code) * r0 + (code - 1) * r1), (float)(5 * full));
    return code == 6 ? (snorm ? -1.0f : 0.0f) : 1.0f;
}
// BC7: the format fixes the 8-bit texel (include/r3_bc7_tables.h, rule R11); instead of unpacking the block field after field like the oracle,
// the one texel asked for reads its fields at offsets computed from the mode row.
__device__ __forceinline__ uint32_t bc7_bits(unsigned long long lo, unsigned long long hi, uint32_t pos, uint32_t width) {   // width 0..8
    const unsigned long long v = pos >= 64u ? hi >> (pos - 64u) : (lo >> pos) | ((hi << 1) << (63u - pos));
    return (uint32_t)v & ((1u << width) - 1u);
}
__device__ __forceinline__ uint32_t bc7_widen(uint32_t v, uint32_t bits) { return ((v << (8u - bits)) | (v >> (2u * bits - 8u))) & 255u; }
__device__ __forceinline__ uint32_t bc7_weight(uint32_t index, uint32_t bits) { const uint32_t m = (1u << bits) - 1u; return (index * 64u + (m >> 1)) / m; }   // {0,21,43,64}, {0,9,...,64}, {0,4,...,64}
__device__ __noinline__ uchar4 bc7_texel(uint4 blk, uint32_t t) {
    const int mode = __ffs((int)(blk.x & 0xFFu)) - 1;
    if (mode < 0) return make_uchar4(0, 0, 0, 0);                                    // reserved mode
    const unsigned long long lo = ((unsigned long long)blk.y << 32) | blk.x, hi = ((unsigned long long)blk.w << 32) | blk.z;
    const uint8_t* m = r3_bc7_modes[mode];
    const uint32_t ns = m[0], pb = m[1], rb = m[2], isb = m[3], cb = m[4], ab = m[5], epb = m[6], spb = m[7], ib = m[8], ib2 = m[9];
    uint32_t pos = (uint32_t)mode + 1u;
    const uint32_t partition = bc7_bits(lo, hi, pos, pb); pos += pb;
    const uint32_t rotation = bc7_bits(lo, hi, pos, rb); pos += rb;
    const uint32_t idxsel = bc7_bits(lo, hi, pos, isb); pos += isb;
    uint32_t subset = 0u, a1 = 16u, a2 = 16u;                                       // a1, a2: the other anchor texels (16 = none)
    if (ns == 2u) { subset = (r3_bc7_partition2[partition] >> t) & 1u; a1 = r3_bc7_anchor2[partition]; }
    else if (ns == 3u) { subset = (r3_bc7_partition3[partition] >> (2u * t)) & 3u; a1 = r3_bc7_anchor3a[partition]; a2 = r3_bc7_anchor3b[partition]; }
    const uint32_t e0 = 2u * subset, e1 = e0 + 1u, alpha_base = pos + 6u * ns * cb, p_base = alpha_base + 2u * ns * ab;
    uint32_t lo_ep[4], hi_ep[4];
#pragma unroll
    for (uint32_t ch = 0; ch < 3u; ++ch) {
        lo_ep[ch] = bc7_bits(lo, hi, pos + (ch * 2u * ns + e0) * cb, cb);
        hi_ep[ch] = bc7_bits(lo, hi, pos + (ch * 2u * ns + e1) * cb, cb);
    }
    lo_ep[3] = bc7_bits(lo, hi, alpha_base + e0 * ab, ab);
    hi_ep[3] = bc7_bits(lo, hi, alpha_base + e1 * ab, ab);
    uint32_t index_base = p_base, cbits = cb, abits = ab;
    if (epb | spb) {
        const uint32_t p0 = bc7_bits(lo, hi, epb ? p_base + e0 : p_base + subset, 1u), p1 = epb ? bc7_bits(lo, hi, p_base + e1, 1u) : p0;
#pragma unroll
        for (uint32_t ch = 0; ch < 4u; ++ch) { lo_ep[ch] = (lo_ep[ch] << 1) | p0; hi_ep[ch] = (hi_ep[ch] << 1) | p1; }
        index_base += epb ? 2u * ns : ns;
        ++cbits; ++abits;
    }
#pragma unroll
    for (uint32_t ch = 0; ch < 3u; ++ch) { lo_ep[ch] = bc7_widen(lo_ep[ch], cbits); hi_ep[ch] = bc7_widen(hi_ep[ch], cbits); }
    if (ab) { lo_ep[3] = bc7_widen(lo_ep[3], abits); hi_ep[3] = bc7_widen(hi_ep[3], abits); } else { lo_ep[3] = 255u; hi_ep[3] = 255u; }
    // the texel's index: every anchor in front of it shortens the offset by one bit, and an anchor's own index is one bit short
    const uint32_t before = (t > 0u) + (t > a1) + (t > a2), is_anchor = (t == 0u) | (t == a1) | (t == a2);
    uint32_t ci = bc7_bits(lo, hi, index_base + t * ib - before, ib - is_anchor), cw = ib, ai = ci, aw = ib;
    if (ib2) {
        const uint32_t second = bc7_bits(lo, hi, index_base + 16u * ib - 1u + t * ib2 - (t > 0u), ib2 - (t == 0u));
        if (idxsel) { ai = ci; aw = ib; ci = second; cw = ib2; } else { ai = second; aw = ib2; }
    }
    const uint32_t wc = bc7_weight(ci, cw), wa = bc7_weight(ai, aw);
    uint32_t r = ((64u - wc) * lo_ep[0] + wc * hi_ep[0] + 32u) >> 6, g = ((64u - wc) * lo_ep[1] + wc * hi_ep[1] + 32u) >> 6,
             b = ((64u - wc) * lo_ep[2] + wc * hi_ep[2] + 32u) >> 6, a = ((64u - wa) * lo_ep[3] + wa * hi_ep[3] + 32u) >> 6;
    if (rotation == 1u) { const uint32_t s = r; r = a; a = s; } else if (rotation == 2u) { const uint32_t s = g; g = a; a = s; } else if (rotation == 3u) { const uint32_t s = b; b = a; a = s; }
    return make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)b, (unsigned char)a);
}
__device__ __noinline__ float4 block_texel_fetch(const uint8_t* level_base, uint32_t f, long long w, long long x, long long y) {
    const uint8_t* b = level_base + (unsigned long long)((y >> 2) * ((w + 3) >> 2) + (x >> 2)) * R3_TEXFMT_BLOCK_BYTES(f);
    const uint32_t t = (uint32_t)((y & 3) * 4 + (x & 3));
    if (f == R3_TEXFMT_BC7_RGBA_UNORM || f == R3_TEXFMT_BC7_RGBA_UNORM_SRGB) {
        const uchar4 c8 = bc7_texel(__ldg(reinterpret_cast<const uint4*>(b)), t);
        float4 o = make_float4(div_rn((float)c8.x, 255.0f), div_rn((float)c8.y, 255.0f), div_rn((float)c8.z, 255.0f), div_rn((float)c8.w, 255.0f));
        if (f == R3_TEXFMT_BC7_RGBA_UNORM_SRGB) { o.x = srgb_to_linear(o.x); o.y = srgb_to_linear(o.y); o.z = srgb_to_linear(o.z); }
        return o;
    }
    const uint2 first = __ldg(reinterpret_cast<const uint2*>(b));
    if (f == R3_TEXFMT_BC4_R_UNORM || f == R3_TEXFMT_BC4_R_SNORM) return make_float4(bc_channel(first, t, f == R3_TEXFMT_BC4_R_SNORM), 0.0f, 0.0f, 1.0f);
    const bool bc1 = f == R3_TEXFMT_BC1_RGBA_UNORM || f == R3_TEXFMT_BC1_RGBA_UNORM_SRGB;
    const uint2 second = bc1 ? first : __ldg(reinterpret_cast<const uint2*>(b + 8));
    if (f == R3_TEXFMT_BC5_RG_UNORM || f == R3_TEXFMT_BC5_RG_SNORM)
        return make_float4(bc_channel(first, t, f == R3_TEXFMT_BC5_RG_SNORM), bc_channel(second, t, f == R3_TEXFMT_BC5_RG_SNORM), 0.0f, 1.0f);
    float alpha = 1.0f;
    float3 c = bc_colour(second, t, bc1, alpha);
    if (f == R3_TEXFMT_BC2_RGBA_UNORM || f == R3_TEXFMT_BC2_RGBA_UNORM_SRGB) alpha = div_rn((float)(((t < 8u ? first.x : first.y) >> (4u * (t & 7u))) & 15u), 15.0f);
    else if (!bc1) alpha = bc_channel(first, t, false);
    if (f == R3_TEXFMT_BC1_RGBA_UNORM_SRGB || f == R3_TEXFMT_BC2_RGBA_UNORM_SRGB || f == R3_TEXFMT_BC3_RGBA_UNORM_SRGB) { c.x = srgb_to_linear(c.x); c.y = srgb_to_linear(c.y); c.z = srgb_to_linear(c.z); }
    return make_float4(c.x, c.y, c.z, alpha);
}
// the uncompressed formats beyond the five common ones (include/r3_layouts.h): one aligned load of the texel, missing channels (0, 0, 1)
__device__ __forceinline__ float snorm8(int v) { return div_rn((float)max(v, -127), 127.0f); }
__device__ __forceinline__ float half_bits(unsigned short h) { return __half2float(__ushort_as_half(h)); }
__device__ __noinline__ float4 wide_texel_fetch(const uint8_t* t, uint32_t f) {
    switch (f) {
    case R3_TEXFMT_R8_SNORM: return make_float4(snorm8((signed char)__ldg(t)), 0.0f, 0.0f, 1.0f);
    case R3_TEXFMT_RG8_SNORM: { const char2 c = __ldg(reinterpret_cast<const char2*>(t)); return make_float4(snorm8(c.x), snorm8(c.y), 0.0f, 1.0f); }
    case R3_TEXFMT_RGBA8_SNORM: { const char4 c = __ldg(reinterpret_cast<const char4*>(t)); return make_float4(snorm8(c.x), snorm8(c.y), snorm8(c.z), snorm8(c.w)); }
    case R3_TEXFMT_BGRA8_UNORM:
    case R3_TEXFMT_BGRA8_UNORM_SRGB: {
        const uchar4 c = __ldg(reinterpret_cast<const uchar4*>(t));
        float4 o = make_float4(div_rn((float)c.z, 255.0f), div_rn((float)c.y, 255.0f), div_rn((float)c.x, 255.0f), div_rn((float)c.w, 255.0f));
        if (f == R3_TEXFMT_BGRA8_UNORM_SRGB) { o.x = srgb_to_linear(o.x); o.y = srgb_to_linear(o.y); o.z = srgb_to_linear(o.z); }
        return o;
    }
    case R3_TEXFMT_RGB10A2_UNORM: {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(t));
        return make_float4(div_rn((float)(v & 1023u), 1023.0f), div_rn((float)((v >> 10) & 1023u), 1023.0f), div_rn((float)((v >> 20) & 1023u), 1023.0f), div_rn((float)(v >> 30), 3.0f));
    }
    case R3_TEXFMT_R16_FLOAT: return make_float4(half_bits(__ldg(reinterpret_cast<const unsigned short*>(t))), 0.0f, 0.0f, 1.0f);
    case R3_TEXFMT_RG16_FLOAT: { const ushort2 h = __ldg(reinterpret_cast<const ushort2*>(t)); return make_float4(half_bits(h.x), half_bits(h.y), 0.0f, 1.0f); }
    case R3_TEXFMT_RGBA16_FLOAT: { const ushort4 h = __ldg(reinterpret_cast<const ushort4*>(t)); return make_float4(half_bits(h.x), half_bits(h.y), half_bits(h.z), half_bits(h.w)); }
    case R3_TEXFMT_R32_FLOAT: return make_float4(__ldg(reinterpret_cast<const float*>(t)), 0.0f, 0.0f, 1.0f);
    case R3_TEXFMT_RG32_FLOAT: { const float2 v = __ldg(reinterpret_cast<const float2*>(t)); return make_float4(v.x, v.y, 0.0f, 1.0f); }
    case R3_TEXFMT_R16_UNORM: return make_float4(div_rn((float)__ldg(reinterpret_cast<const unsigned short*>(t)), 65535.0f), 0.0f, 0.0f, 1.0f);
    case R3_TEXFMT_RG16_UNORM: { const ushort2 h = __ldg(reinterpret_cast<const ushort2*>(t)); return make_float4(div_rn((float)h.x, 65535.0f), div_rn((float)h.y, 65535.0f), 0.0f, 1.0f); }
    default: { const ushort4 h = __ldg(reinterpret_cast<const ushort4*>(t));          // R3_TEXFMT_RGBA16_UNORM (r3_set_textures validated the format)
        return make_float4(div_rn((float)h.x, 65535.0f), div_rn((float)h.y, 65535.0f), div_rn((float)h.z, 65535.0f), div_rn((float)h.w, 65535.0f)); }
    }
}
__device__ __forceinline__ float4 texel_fetch(const TexTable& p, const r3_texture_desc& d, uint32_t level, long long x, long long y) {
    unsigned long long off = d.byte_offset;
    for (uint32_t l = 0; l < level; ++l) off += R3_TEXFMT_LEVEL_BYTES(d.format, max(d.width >> l, 1u), max(d.height >> l, 1u));
    const long long w = max(d.width >> level, 1u), h = max(d.height >> level, 1u);
    if (p.clamp_to_edge) { x = x < 0 ? 0 : (x >= w ? w - 1 : x); y = y < 0 ? 0 : (y >= h ? h - 1 : y); }   // texels clamped to the cube face
    else { x = ((x % w) + w) % w; y = ((y % h) + h) % h; }                          // AddressMode::Repeat
    if (R3_TEXFMT_IS_BLOCK(d.format)) return block_texel_fetch(p.texels + off, d.format, w, x, y);
    const unsigned long long bpp = R3_TEXFMT_BPP(d.format);
    const uint8_t* t = p.texels + off + (unsigned long long)(y * w + x) * bpp;
    if (d.format >= R3_TEXFMT_R8_SNORM) return wide_texel_fetch(t, d.format);
    if (d.format == R3_TEXFMT_RGBA32_FLOAT) return __ldg(reinterpret_cast<const float4*>(t));
    if (d.format == R3_TEXFMT_R8_UNORM) return make_float4((float)__ldg(t) / 255.0f, 0.0f, 0.0f, 1.0f);                     // missing channels read (0, 0, 1)
    if (d.format == R3_TEXFMT_RG8_UNORM) { const uchar2 c2 = __ldg(reinterpret_cast<const uchar2*>(t)); return make_float4((float)c2.x / 255.0f, (float)c2.y / 255.0f, 0.0f, 1.0f); }
    const uchar4 c = __ldg(reinterpret_cast<const uchar4*>(t));
    float4 o = make_float4((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, (float)c.w / 255.0f);
    if (d.format == R3_TEXFMT_RGBA8_UNORM_SRGB) { o.x = srgb_to_linear(o.x); o.y = srgb_to_linear(o.y); o.z = srgb_to_linear(o.z); }
    return o;
}
__device__ __forceinline__ float clamp_coord(float v) { return fminf(fmaxf(v, -1.0e9f), 1.0e9f); }
__device__ __noinline__ float4 sample_level(const TexTable& p, const r3_texture_desc& d, uint32_t level, bool nearest, float u, float v) {
    const float w = (float)max(d.width >> level, 1u), h = (float)max(d.height >> level, 1u);
    if (nearest) {
        float x = floorf(mul_rn(u, w)), y = floorf(mul_rn(v, h));
        if (!(x == x)) x = 0.0f; if (!(y == y)) y = 0.0f;
        return texel_fetch(p, d, level, (long long)clamp_coord(x), (long long)clamp_coord(y));
    }
    const float x = sub_rn(mul_rn(u, w), 0.5f), y = sub_rn(mul_rn(v, h), 0.5f);
    float x0 = floorf(x), y0 = floorf(y), fx = sub_rn(x, x0), fy = sub_rn(y, y0);
    if (!(x0 == x0)) { x0 = 0.0f; fx = 0.0f; } if (!(y0 == y0)) { y0 = 0.0f; fy = 0.0f; }
    const long long ix = (long long)clamp_coord(x0), iy = (long long)clamp_coord(y0);
    const float4 t00 = texel_fetch(p, d, level, ix, iy), t10 = texel_fetch(p, d, level, ix + 1, iy);
    const float4 t01 = texel_fetch(p, d, level, ix, iy + 1), t11 = texel_fetch(p, d, level, ix + 1, iy + 1);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    return make_float4((t00.x * gx + t10.x * fx) * gy + (t01.x * gx + t11.x * fx) * fy, (t00.y * gx + t10.y * fx) * gy + (t01.y * gx + t11.y * fx) * fy,
                       (t00.z * gx + t10.z * fx) * gy + (t01.z * gx + t11.z * fx) * fy, (t00.w * gx + t10.w * fx) * gy + (t01.w * gx + t11.w * fx) * fy);
}
struct TexCoords { float u, v, dudx, dvdx, dudy, dvdy; };
// log2 of the footprint, rule R9: WGSL leaves the LOD arithmetic of textureSampleGrad to the driver; the rule fixes it as a sequence of
// IEEE f32 operations (no MUFU.LG2, no libm) so that the level picked at a footprint of 2^(k+0.5) texels is one bit pattern:
// x = 2^e * m, m in [sqrt(1/2), sqrt(2)), s = (m - 1) / (m + 1), log2 x = e + (2 / ln 2) * s * (1 + s^2/3 + s^4/5 + s^6/7 + s^8/9).
__device__ __forceinline__ float log2_r9(float x) {
    if (!(x > 0.0f)) return x == 0.0f ? -__int_as_float(0x7F800000) : __int_as_float(0x7FC00000);
    const uint32_t b = __float_as_uint(x);
    if (b >= 0x7F800000u) return __int_as_float(0x7F800000);
    int e = (int)(b >> 23) - 127;
    if (e == -127) return -127.0f;
    float m = __uint_as_float((b & 0x7FFFFFu) | 0x3F800000u);
    if (m > 1.41421356f) { m = mul_rn(m, 0.5f); e += 1; }
    const float s = div_rn(sub_rn(m, 1.0f), add_rn(m, 1.0f)), s2 = mul_rn(s, s);
    float q = 0.11111111f;
    q = add_rn(mul_rn(q, s2), 0.14285715f);
    q = add_rn(mul_rn(q, s2), 0.2f);
    q = add_rn(mul_rn(q, s2), 0.33333334f);
    q = add_rn(mul_rn(q, s2), 1.0f);
    return add_rn((float)e, mul_rn(mul_rn(s, q), 2.88539008f));
}
__device__ __noinline__ float4 sample_grad_desc(const TexTable& p, const r3_texture_desc& d, bool nearest, const TexCoords& c) {
    // level SELECTION in the oracle's operation order, never contracted (this header is also compiled into the FMA-enabled shading unit)
    const float w0 = (float)d.width, h0 = (float)d.height;
    const float ax = mul_rn(c.dudx, w0), ay = mul_rn(c.dvdx, h0), bx = mul_rn(c.dudy, w0), by = mul_rn(c.dvdy, h0);
    const float rho = fmaxf(__fsqrt_rn(add_rn(mul_rn(ax, ax), mul_rn(ay, ay))), __fsqrt_rn(add_rn(mul_rn(bx, bx), mul_rn(by, by))));
    const float lambda = log2_r9(rho);
    const uint32_t last = d.mip_count - 1u;
    if (!(lambda > 0.0f)) return sample_level(p, d, 0u, nearest, c.u, c.v);
    if (nearest) {
        const float lv = floorf(add_rn(lambda, 0.5f));
        return sample_level(p, d, lv >= (float)last ? last : (uint32_t)lv, true, c.u, c.v);
    }
    const float l = fminf(lambda, (float)last), lo = floorf(l), fr = sub_rn(l, lo);
    const uint32_t level = (uint32_t)lo;
    const float4 a = sample_level(p, d, level, false, c.u, c.v);
    if (level >= last || fr == 0.0f) return a;
    const float4 b = sample_level(p, d, level + 1u, false, c.u, c.v);
    const float g = 1.0f - fr;
    return make_float4(a.x * g + b.x * fr, a.y * g + b.y * fr, a.z * g + b.z * fr, a.w * g + b.w * fr);
}
// slot value = table index + 1; an index outside the table reads zeros (robust access)
__device__ __forceinline__ float4 texture_sample_grad(const TexTable& p, uint32_t slot_value, bool nearest, const TexCoords& c) {
    if (slot_value == 0u || slot_value > p.n_tex) return make_float4(0.f, 0.f, 0.f, 0.f);
    return sample_grad_desc(p, p.tex[slot_value - 1u], nearest, c);
}

}  // namespace
#endif
