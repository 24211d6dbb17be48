// lpv_sample.cuh — SampleLPVData (ReflectionTraceFrag.glsl:1516-1528) with SampleLPVColor (:1484-1487) and InterpolateLPVColorDithered
// (:1490-1509): the light the reflection pass takes from the propagation volume at a point in voxel units — trilinear light level
// (R8 unorm, LINEAR, CLAMP_TO_EDGE) times the average colour of the block types at eight dithered taps (R8UI, NEAREST), desaturated
// half way.  Exact float arithmetic only (no contraction), bit-identical to the compiled shader function.  Shared by the batch entry
// point (lpv.cu, vxrt_cuda_lpv_sample) and the consumer inside the reflection shading (ApproximateGILPV, :673-700).
#pragma once
#include "vmath.cuh"

struct LpvSampleArgs {
    const uint8_t* level;
    const uint8_t* type;
    const float4* avg;
    int nx, ny, nz;
    float dx, dy, dz;   // LPVDither (:714-723) of the batch entry point; the reflection pass computes its own per pixel
};
__device__ __forceinline__ int lpv_clampi(int i, int n) { return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); }
__device__ __forceinline__ float lpv_texel(const LpvSampleArgs& a, int i, int j, int k) {
    return (float)__ldg(a.level + i + (size_t)j * a.nx + (size_t)k * a.nx * a.ny) / 255.0f;
}
__device__ __forceinline__ f3 lpv_color_at(const LpvSampleArgs& a, float x, float y, float z) {
    const int i = lpv_clampi(cvt_floor(x * (float)a.nx), a.nx), j = lpv_clampi(cvt_floor(y * (float)a.ny), a.ny), k = lpv_clampi(cvt_floor(z * (float)a.nz), a.nz);
    const unsigned id = __ldg(a.type + i + (size_t)j * a.nx + (size_t)k * a.nx * a.ny);
    if (id > 127u) return F3(0.0f, 0.0f, 0.0f);   // clamp(BlockID, 0u, 128u) indexes past the 128-entry table: read as 0
    const float4 c = __ldg(a.avg + id);
    return F3(c.x, c.y, c.z);
}
// point in voxel units (like HitPosition), dither = LPVDither
__device__ __forceinline__ f3 lpv_sample_data(const LpvSampleArgs& a, f3 point, f3 dither) {
    const float R[3] = {384.0f, 128.0f, 384.0f};   // VolumeResolution, hard-coded in the shader
    const float P[3] = {point.x, point.y, point.z};
    float UV[3], W0[3], W1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        UV[c] = P[c] * (1.0f / R[c]);
        const float t = UV[c] * R[c];
        const float F = t - floorf(t);
        const float L = (F * (F - 1.0f) + 0.5f) / R[c];
        W0[c] = UV[c] - L; W1[c] = UV[c] + L;
    }
    // texture(u_LPV, UV).x
    const float u = UV[0] * (float)a.nx - 0.5f, v = UV[1] * (float)a.ny - 0.5f, w = UV[2] * (float)a.nz - 0.5f;
    const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    const float wa = u - fu, wb = v - fv, wg = w - fw;
    const int i0 = lpv_clampi(cvt_floor(fu), a.nx), i1 = lpv_clampi(cvt_floor(fu) + 1, a.nx);
    const int j0 = lpv_clampi(cvt_floor(fv), a.ny), j1 = lpv_clampi(cvt_floor(fv) + 1, a.ny);
    const int k0 = lpv_clampi(cvt_floor(fw), a.nz), k1 = lpv_clampi(cvt_floor(fw) + 1, a.nz);
    float pl[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = q ? k1 : k0;
        const float top = lpv_texel(a, i0, j0, k) * (1.0f - wa) + lpv_texel(a, i1, j0, k) * wa;
        const float bot = lpv_texel(a, i0, j1, k) * (1.0f - wa) + lpv_texel(a, i1, j1, k) * wa;
        pl[q] = top * (1.0f - wb) + bot * wb;
    }
    const float level = pl[0] * (1.0f - wg) + pl[1] * wg;
    // the eight dithered taps in the order of the shader; DitherWeights = 1, GlobalDitherNoiseWeight = 2
    const float d0 = (dither.x * 1.0f) * 2.0f, d1 = (dither.y * 1.0f) * 2.0f, d2 = (dither.z * 1.0f) * 2.0f;
    f3 sum = F3(0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int sx = (k == 1 || k == 2 || k == 5 || k == 6), sy = (k >= 2 && k <= 5), sz = (k >= 4);
        const float x = sx ? W1[0] : W0[0], y = sy ? W1[1] : W0[1], z = sz ? W1[2] : W0[2];
        const f3 c = (k & 1) ? lpv_color_at(a, x - d0, y - d1, z - d2) : lpv_color_at(a, x + d0, y + d1, z + d2);
        sum = k == 0 ? c : F3(sum.x + c.x, sum.y + c.y, sum.z + c.z);
    }
    const f3 col = F3(gmax(sum.x / 8.0f, 0.00000001f), gmax(sum.y / 8.0f, 0.00000001f), gmax(sum.z / 8.0f, 0.00000001f));
    const float s = level * 325.0f;
    const f3 Fi = F3(s * col.x, s * col.y, s * col.z);
    const float luma = (Fi.x * 0.2125f + Fi.y * 0.7154f) + Fi.z * 0.0721f;
    return F3(luma * (1.0f - 0.5f) + Fi.x * 0.5f, luma * (1.0f - 0.5f) + Fi.y * 0.5f, luma * (1.0f - 0.5f) + Fi.z * 0.5f);
}
