// texture.cuh — device-side texture model of the block texture arrays, FBO attachments and sky cube
// maps.  GL leaves filtering details to the driver; the pinned behaviour is listed in
// REPEAT wrap, texel centres at +0.5, bilinear weights in full float, lod <= 0.5 =>
// magnification filter (NEAREST on level 0), otherwise bilinear in floor(lod) / floor(lod)+1 blended by
// fract(lod); albedo decoded from sRGB before filtering; 8-bit mip levels.
#pragma once
#include "vmath.cuh"

#include "ctx.h"

VXD f4 texarray_texel(const TexArrayDev& t, int level, int layer, int x, int y) {
    int lw = max(t.w >> level, 1), lh = max(t.h >> level, 1);
    const uchar4 p = __ldg(reinterpret_cast<const uchar4*>(t.data + t.level_offset[level]) + (((size_t)layer * lh + y) * lw + x));
    return F4(__ldg(t.decode + p.x), __ldg(t.decode + p.y), __ldg(t.decode + p.z), unorm8_to_float(p.w));
}

VXD f4 lerp4(f4 a, f4 b, float t) {
    return F4(a.x * (1.0f - t) + b.x * t, a.y * (1.0f - t) + b.y * t, a.z * (1.0f - t) + b.z * t, a.w * (1.0f - t) + b.w * t);
}

VXD f4 texarray_bilinear(const TexArrayDev& t, int level, int layer, float s, float tt) {
    int lw = max(t.w >> level, 1), lh = max(t.h >> level, 1);
    float u = s * (float)lw - 0.5f, v = tt * (float)lh - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), lw), j0 = wrap_repeat(cvt_floor(fv), lh);
    int i1 = wrap_repeat(i0 + 1, lw), j1 = wrap_repeat(j0 + 1, lh);
    f4 t00 = texarray_texel(t, level, layer, i0, j0), t10 = texarray_texel(t, level, layer, i1, j0);
    f4 t01 = texarray_texel(t, level, layer, i0, j1), t11 = texarray_texel(t, level, layer, i1, j1);
    return lerp4(lerp4(t00, t10, a), lerp4(t01, t11, a), b);
}

// textureLod(sampler2DArray, vec3(s, t, layer), lod); texture() with undefined derivatives uses lod 0
VXD f4 texarray_sample(const TexArrayDev& t, float s, float tt, float layer_f, float lod) {
    int layer = iclamp(cvt_round(layer_f), 0, t.layers - 1);
    int maxl = t.levels - 1;
    if (!(lod > 0.5f)) {
        int i = wrap_repeat(cvt_floor(s * (float)t.w), t.w), j = wrap_repeat(cvt_floor(tt * (float)t.h), t.h);
        return texarray_texel(t, 0, layer, i, j);
    }
    float l = gmin(lod, (float)maxl);
    int d1 = cvt_floor(l);
    float f = l - (float)d1;
    f4 a = texarray_bilinear(t, d1, layer, s, tt);
    if (f == 0.0f || d1 >= maxl) return a;
    f4 b = texarray_bilinear(t, d1 + 1, layer, s, tt);
    return lerp4(a, b, f);
}

// texture(samplerCube, dir): LINEAR within the selected face, CLAMP_TO_EDGE
VXD f3 texcube_sample(const TexCubeDev& t, f3 d) {
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = d.x >= 0.0f ? 0 : 1; sc = d.x >= 0.0f ? -d.z : d.z; tc = -d.y; ma = ax; }
    else if (ay >= az) { face = d.y >= 0.0f ? 2 : 3; sc = d.x; tc = d.y >= 0.0f ? d.z : -d.z; ma = ay; }
    else { face = d.z >= 0.0f ? 4 : 5; sc = d.z >= 0.0f ? d.x : -d.x; tc = -d.y; ma = az; }
    float s = 0.5f * (sc / ma + 1.0f), tt = 0.5f * (tc / ma + 1.0f);
    float u = s * (float)t.res - 0.5f, v = tt * (float)t.res - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = iclamp(cvt_floor(fu), 0, t.res - 1), j0 = iclamp(cvt_floor(fv), 0, t.res - 1);
    int i1 = iclamp(cvt_floor(fu) + 1, 0, t.res - 1), j1 = iclamp(cvt_floor(fv) + 1, 0, t.res - 1);
    const float* f = t.data + (size_t)face * t.res * t.res * 3;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float t00 = __ldg(f + ((size_t)j0 * t.res + i0) * 3 + c), t10 = __ldg(f + ((size_t)j0 * t.res + i1) * 3 + c);
        float t01 = __ldg(f + ((size_t)j1 * t.res + i0) * 3 + c), t11 = __ldg(f + ((size_t)j1 * t.res + i1) * 3 + c);
        float top = t00 * (1.0f - a) + t10 * a;
        float bot = t01 * (1.0f - a) + t11 * a;
        o[c] = top * (1.0f - b) + bot * b;
    }
    return F3(o[0], o[1], o[2]);
}

// ---- FBO attachment reads (REPEAT wrap; LINEAR or NEAREST per attachment) ----
VXD float att_r16f_bilinear(const uint16_t* __restrict__ img, int w, int h, f2 uv) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
    float t00 = half_bits_to_float(__ldg(img + (size_t)j0 * w + i0)), t10 = half_bits_to_float(__ldg(img + (size_t)j0 * w + i1));
    float t01 = half_bits_to_float(__ldg(img + (size_t)j1 * w + i0)), t11 = half_bits_to_float(__ldg(img + (size_t)j1 * w + i1));
    float top = t00 * (1.0f - a) + t10 * a;
    float bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
VXD float att_r32f_bilinear(const float* __restrict__ img, int w, int h, f2 uv) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
    float t00 = __ldg(img + (size_t)j0 * w + i0), t10 = __ldg(img + (size_t)j0 * w + i1);
    float t01 = __ldg(img + (size_t)j1 * w + i0), t11 = __ldg(img + (size_t)j1 * w + i1);
    float top = t00 * (1.0f - a) + t10 * a;
    float bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
VXD float att_r8_nearest(const uint8_t* __restrict__ img, int w, int h, f2 uv) {
    int i = wrap_repeat(cvt_floor(uv.x * (float)w), w), j = wrap_repeat(cvt_floor(uv.y * (float)h), h);
    return unorm8_to_float(__ldg(img + (size_t)j * w + i));
}
// bilinear read of an interleaved half / unorm8 attachment with CH channels (LINEAR, REPEAT)
template <int CH>
VXD void att_half_bilinear(const uint16_t* __restrict__ img, int w, int h, f2 uv, float* out) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float t00 = half_bits_to_float(__ldg(img + ((size_t)j0 * w + i0) * CH + c)), t10 = half_bits_to_float(__ldg(img + ((size_t)j0 * w + i1) * CH + c));
        float t01 = half_bits_to_float(__ldg(img + ((size_t)j1 * w + i0) * CH + c)), t11 = half_bits_to_float(__ldg(img + ((size_t)j1 * w + i1) * CH + c));
        float top = t00 * (1.0f - a) + t10 * a;
        float bot = t01 * (1.0f - a) + t11 * a;
        out[c] = top * (1.0f - b) + bot * b;
    }
}
template <int CH>
VXD void att_unorm8_bilinear(const uint8_t* __restrict__ img, int w, int h, f2 uv, float* out) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float t00 = unorm8_to_float(__ldg(img + ((size_t)j0 * w + i0) * CH + c)), t10 = unorm8_to_float(__ldg(img + ((size_t)j0 * w + i1) * CH + c));
        float t01 = unorm8_to_float(__ldg(img + ((size_t)j1 * w + i0) * CH + c)), t11 = unorm8_to_float(__ldg(img + ((size_t)j1 * w + i1) * CH + c));
        float top = t00 * (1.0f - a) + t10 * a;
        float bot = t01 * (1.0f - a) + t11 * a;
        out[c] = top * (1.0f - b) + bot * b;
    }
}
