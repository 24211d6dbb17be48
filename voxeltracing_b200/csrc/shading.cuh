// shading.cuh — helpers shared by the material / GI / reflection kernels: ray set-up, face tables,
// UV derivation, the blue-noise sampler and the BRDF terms.  Citations: Core/Shaders/<file>:line.
#pragma once
#include "traverse.cuh"
#include "texture.cuh"

#define VX_PI 3.14159265359f

// GetRayDirectionAt (identical in every pass, e.g. GenerateGBuffer.glsl:110-115)
VXD f3 ray_direction_at(const float* inv_view, const float* inv_proj, f2 ss) {
    f4 clip = F4(ss.x * 2.0f - 1.0f, ss.y * 2.0f - 1.0f, -1.0f, 1.0f);
    f4 e = mat4_mul(inv_proj, clip);
    f4 r = mat4_mul(inv_view, F4(e.x, e.y, -1.0f, 0.0f));
    return F3(r.x, r.y, r.z);
}
VXD f2 pixel_uv(int px, int py, int W, int H) { return F2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H); }

// pixel of this thread: CTA = 32x8 pixels, warp = 8x4 tile
VXD void tile_pixel(int& px, int& py, int row0, int col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = col0 + blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    py = row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
}

VXD f3 face_normal(int i) {
    switch (i) {
        case 0: return F3(0.0f, 0.0f, 1.0f);
        case 1: return F3(0.0f, 0.0f, -1.0f);
        case 2: return F3(0.0f, 1.0f, 0.0f);
        case 3: return F3(0.0f, -1.0f, 0.0f);
        case 4: return F3(-1.0f, 0.0f, 0.0f);
        default: return F3(1.0f, 0.0f, 0.0f);
    }
}
// GetNormalFromID: idx > 5 returns `miss`
VXD f3 normal_from_id(float n, f3 miss) {
    int i = cvt_round(n * 10.0f);
    if (i > 5) return miss;
    return face_normal(i);
}
// CompareVec3, e = 0.0125 (GenerateGBuffer.glsl:125-128)
VXD bool cmp3(f3 a, f3 b) {
    const float e = 0.0125f;
    return fabsf(a.x - b.x) < e && fabsf(a.y - b.y) < e && fabsf(a.z - b.z) < e;
}
VXD bool eq3(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// face index of an axis-aligned normal by CompareVec3, -1 if none matches
VXD int face_of(f3 n) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (cmp3(n, face_normal(i))) return i;
    return -1;
}
// CalculateVectors (GenerateGBuffer.glsl:444-513, ReflectionTraceFrag.glsl:1388-1449)
VXD void calculate_vectors(f3 p, f3 n, f3& tangent, f3& bitangent, f2& uv) {
    int i = face_of(n);
    if (i < 0) return;
    uv = (i < 2) ? F2(gfract(p.x), gfract(p.y)) : ((i < 4) ? F2(gfract(p.x), gfract(p.z)) : F2(gfract(p.z), gfract(p.y)));
    tangent = (i < 4) ? F3(1.0f, 0.0f, 0.0f) : F3(0.0f, 0.0f, -1.0f);
    bitangent = (i < 2) ? F3(0.0f, 1.0f, 0.0f) : ((i < 4) ? F3(0.0f, 0.0f, 1.0f) : F3(0.0f, -1.0f, 0.0f));
}
// CalculateUV (DiffuseRayTraceFrag.glsl:1321-1359)
VXD void calculate_uv(f3 p, f3 n, f2& uv) {
    int i = face_of(n);
    if (i < 0) return;
    uv = (i < 2) ? F2(gfract(p.x), gfract(p.y)) : ((i < 4) ? F2(gfract(p.x), gfract(p.z)) : F2(gfract(p.z), gfract(p.y)));
}
// BasicSaturation (ColorPassFrag.glsl:1228-1233)
VXD f3 basic_saturation(f3 c, float adj) {
    float l = dot(c, F3(0.2125f, 0.7154f, 0.0721f));
    return gmix(F3(l), c, adj);
}
VXD f3 xyz(f4 a) { return F3(a.x, a.y, a.z); }

// samplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_32spp (DiffuseRayTraceFrag.glsl:130-153);
// table = sobol[65536] ++ scramble[131072] ++ ranking[131072]; reads past the end return 0 (pinned)
VXD float blue_noise_1d(const int32_t* __restrict__ table, int px, int py, int sampleIndex, int sampleDimension) {
    const int32_t* sobol = table;
    const int32_t* scramble = table + 256 * 256;
    const int32_t* ranking = scramble + 128 * 128 * 8;
    int pi = px & 127, pj = py & 127;
    sampleIndex &= 255;
    sampleDimension &= 255;
    int ri = sampleDimension + (pi + pj * 128) * 8;
    int rank = (ri < 128 * 128 * 8) ? __ldg(ranking + ri) : 0;
    int rankedSampleIndex = sampleIndex ^ rank;
    int si = sampleDimension + rankedSampleIndex * 256;
    int value = (si >= 0 && si < 256 * 256) ? __ldg(sobol + si) : 0;
    value = value ^ __ldg(scramble + (sampleDimension % 8) + (pi + pj * 128) * 8);
    return (0.5f + (float)value) / 256.0f;
}

// ndfGGX / gaSchlickG1 / gaSchlickGGX (ColorPassFrag.glsl:394-413, ReflectionTraceFrag.glsl:283-305)
VXD float ndf_ggx(float cosLh, float roughness) {
    float alpha = roughness * roughness;
    float alphaSq = alpha * alpha;
    float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    return alphaSq / (VX_PI * denom * denom);
}
VXD float ga_schlick_g1(float c, float k) { return c / (c * (1.0f - k) + k); }
VXD float ga_schlick_ggx(float cosLi, float cosLo, float roughness) {
    float r = roughness + 1.0f;
    float k = (r * r) / 8.0f;
    return ga_schlick_g1(cosLi, k) * ga_schlick_g1(cosLo, k);
}
VXD f3 pix3(const uint16_t* __restrict__ p) { return F3(half_bits_to_float(p[0]), half_bits_to_float(p[1]), half_bits_to_float(p[2])); }
