/*
 * vxo_texture.h — CPU ORACLE texture model (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 *
 * The fixed-function sampling behaviour the shaders rely on, pinned where GL leaves
 * it to the driver:
 *   - FBO attachments (Core/GLClasses/Framebuffer.cpp:64-70): REPEAT wrap; LINEAR or NEAREST per
 *     attachment; bilinear weights in full float, texel centres at +0.5.
 *   - block texture arrays (Core/GLClasses/TextureArray.cpp:10-78, BlockDatabase.cpp:519-529):
 *     RGBA8, REPEAT, MIN = LINEAR_MIPMAP_LINEAR, MAG = NEAREST; albedo is sRGB (decoded before
 *     filtering).  lod <= 0.5 selects the magnification filter (nearest texel of level 0), otherwise
 *     bilinear in levels floor(lod), floor(lod)+1 blended by fract(lod).  Mip chain = 2x2 box filter
 *     in linear light, re-quantised to 8 bits per level (what glGenerateMipmap leaves in an RGBA8 /
 *     SRGB8_ALPHA8 texture).  Implicit-derivative texture() calls in divergent flow use lod 0.
 *   - sky cube maps (AtmosphereRenderCubemap.cpp:12-21): LINEAR, CLAMP_TO_EDGE per face (seamless
 *     filtering across faces is not modelled).
 */
#ifndef VXO_TEXTURE_H
#define VXO_TEXTURE_H

#include <math.h>
#include <stdint.h>
#include <vector>

#include "vxo_math.h"

namespace vxo {

/* an FBO colour attachment viewed as floats: channels interleaved */
struct Tex2D {
    const float* data = nullptr;
    int w = 0, h = 0, ch = 1;
    bool linear = true;
};

static inline v4 tex2d_fetch(const Tex2D& t, int x, int y) {
    const float* p = t.data + ((size_t)y * t.w + x) * t.ch;
    v4 r = {p[0], t.ch > 1 ? p[1] : 0.0f, t.ch > 2 ? p[2] : 0.0f, t.ch > 3 ? p[3] : 1.0f};
    return r;
}
static inline int tex_wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

static inline v4 tex2d_sample(const Tex2D& t, float s, float tt) {
    if (!t.linear) {
        int i = tex_wrap(cvt_floor(s * (float)t.w), t.w), j = tex_wrap(cvt_floor(tt * (float)t.h), t.h);
        return tex2d_fetch(t, i, j);
    }
    float u = s * (float)t.w - 0.5f, v = tt * (float)t.h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = tex_wrap(cvt_floor(fu), t.w), j0 = tex_wrap(cvt_floor(fv), t.h);
    int i1 = tex_wrap(i0 + 1, t.w), j1 = tex_wrap(j0 + 1, t.h);
    v4 t00 = tex2d_fetch(t, i0, j0), t10 = tex2d_fetch(t, i1, j0), t01 = tex2d_fetch(t, i0, j1), t11 = tex2d_fetch(t, i1, j1);
    v4 r;
    float* o = &r.x;
    const float *p00 = &t00.x, *p10 = &t10.x, *p01 = &t01.x, *p11 = &t11.x;
    for (int c = 0; c < 4; ++c) {
        float top = p00[c] * (1.0f - a) + p10[c] * a;
        float bot = p01[c] * (1.0f - a) + p11[c] * a;
        o[c] = top * (1.0f - b) + bot * b;
    }
    return r;
}

/* ---- block texture arrays ---- */
struct TexArray {
    /* levels[l] = layers * (w>>l) * (h>>l) * 4 bytes, RGBA8 as stored by GL */
    std::vector<std::vector<uint8_t>> levels;
    int w = 0, h = 0, layers = 0;
    bool srgb = false;
    float decode[256];  /* code -> float for RGB (sRGB or plain /255) */
};

static inline float srgb_to_linear(int c) {
    double cs = (double)c / 255.0;
    double l = cs <= 0.04045 ? cs / 12.92 : pow((cs + 0.055) / 1.055, 2.4);
    return (float)l;
}
static inline uint8_t linear_to_srgb8(float l) {
    double x = l;
    if (!(x > 0.0)) return 0;
    if (x >= 1.0) return 255;
    double s = x <= 0.0031308 ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
    return (uint8_t)nearbyint(s * 255.0);
}

static inline void texarray_build(TexArray& t, const uint8_t* rgba, int layers, int w, int h, bool srgb) {
    t.w = w; t.h = h; t.layers = layers; t.srgb = srgb;
    for (int c = 0; c < 256; ++c) t.decode[c] = srgb ? srgb_to_linear(c) : unorm8_to_float(c);
    t.levels.clear();
    t.levels.emplace_back(rgba, rgba + (size_t)layers * w * h * 4);
    int lw = w, lh = h;
    while (lw > 1 || lh > 1) {
        int nw = lw > 1 ? lw / 2 : 1, nh = lh > 1 ? lh / 2 : 1;
        const std::vector<uint8_t>& src = t.levels.back();
        std::vector<uint8_t> dst((size_t)layers * nw * nh * 4);
        for (int L = 0; L < layers; ++L)
            for (int y = 0; y < nh; ++y)
                for (int x = 0; x < nw; ++x) {
                    int x0 = (2 * x < lw) ? 2 * x : lw - 1, x1 = (2 * x + 1 < lw) ? 2 * x + 1 : lw - 1;
                    int y0 = (2 * y < lh) ? 2 * y : lh - 1, y1 = (2 * y + 1 < lh) ? 2 * y + 1 : lh - 1;
                    const uint8_t* p00 = &src[(((size_t)L * lh + y0) * lw + x0) * 4];
                    const uint8_t* p10 = &src[(((size_t)L * lh + y0) * lw + x1) * 4];
                    const uint8_t* p01 = &src[(((size_t)L * lh + y1) * lw + x0) * 4];
                    const uint8_t* p11 = &src[(((size_t)L * lh + y1) * lw + x1) * 4];
                    uint8_t* o = &dst[(((size_t)L * nh + y) * nw + x) * 4];
                    for (int c = 0; c < 4; ++c) {
                        if (c < 3) {
                            float s = ((t.decode[p00[c]] + t.decode[p10[c]]) + (t.decode[p01[c]] + t.decode[p11[c]])) * 0.25f;
                            o[c] = srgb ? linear_to_srgb8(s) : float_to_unorm8(s);
                        } else {
                            float s = ((unorm8_to_float(p00[c]) + unorm8_to_float(p10[c])) + (unorm8_to_float(p01[c]) + unorm8_to_float(p11[c]))) * 0.25f;
                            o[c] = float_to_unorm8(s);
                        }
                    }
                }
        t.levels.push_back(std::move(dst));
        lw = nw; lh = nh;
    }
}

static inline v4 texarray_texel(const TexArray& t, int level, int layer, int x, int y) {
    int lw = t.w >> level, lh = t.h >> level;
    if (lw < 1) lw = 1;
    if (lh < 1) lh = 1;
    const uint8_t* p = &t.levels[level][(((size_t)layer * lh + y) * lw + x) * 4];
    v4 r = {t.decode[p[0]], t.decode[p[1]], t.decode[p[2]], unorm8_to_float(p[3])};
    return r;
}
static inline v4 texarray_bilinear(const TexArray& t, int level, int layer, float s, float tt) {
    int lw = t.w >> level, lh = t.h >> level;
    if (lw < 1) lw = 1;
    if (lh < 1) lh = 1;
    float u = s * (float)lw - 0.5f, v = tt * (float)lh - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = tex_wrap(cvt_floor(fu), lw), j0 = tex_wrap(cvt_floor(fv), lh);
    int i1 = tex_wrap(i0 + 1, lw), j1 = tex_wrap(j0 + 1, lh);
    v4 t00 = texarray_texel(t, level, layer, i0, j0), t10 = texarray_texel(t, level, layer, i1, j0);
    v4 t01 = texarray_texel(t, level, layer, i0, j1), t11 = texarray_texel(t, level, layer, i1, j1);
    v4 r;
    float* o = &r.x;
    const float *p00 = &t00.x, *p10 = &t10.x, *p01 = &t01.x, *p11 = &t11.x;
    for (int c = 0; c < 4; ++c) {
        float top = p00[c] * (1.0f - a) + p10[c] * a;
        float bot = p01[c] * (1.0f - a) + p11[c] * a;
        o[c] = top * (1.0f - b) + bot * b;
    }
    return r;
}
static inline v4 texarray_sample(const TexArray& t, float s, float tt, float layer_f, float lod) {
    int layer = iclamp(cvt_round(layer_f), 0, t.layers - 1);
    int maxl = (int)t.levels.size() - 1;
    if (!(lod > 0.5f)) { /* magnification: NEAREST on level 0 */
        int i = tex_wrap(cvt_floor(s * (float)t.w), t.w), j = tex_wrap(cvt_floor(tt * (float)t.h), t.h);
        return texarray_texel(t, 0, layer, i, j);
    }
    float l = gmin(lod, (float)maxl);
    int d1 = cvt_floor(l);
    float f = l - (float)d1;
    v4 a = texarray_bilinear(t, d1, layer, s, tt);
    if (f == 0.0f || d1 >= maxl) return a;
    v4 b = texarray_bilinear(t, d1 + 1, layer, s, tt);
    v4 r = {a.x * (1.0f - f) + b.x * f, a.y * (1.0f - f) + b.y * f, a.z * (1.0f - f) + b.z * f, a.w * (1.0f - f) + b.w * f};
    return r;
}

/* ---- sky cube map: 6 faces (+X,-X,+Y,-Y,+Z,-Z) of res*res RGB float ---- */
struct TexCube {
    const float* data = nullptr;
    int res = 0;
};
static inline v4 texcube_sample(const TexCube& t, float x, float y, float z) {
    float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = x >= 0.0f ? 0 : 1; sc = x >= 0.0f ? -z : z; tc = -y; ma = ax; }
    else if (ay >= az) { face = y >= 0.0f ? 2 : 3; sc = x; tc = y >= 0.0f ? z : -z; ma = ay; }
    else { face = z >= 0.0f ? 4 : 5; sc = z >= 0.0f ? x : -x; tc = -y; ma = az; }
    float s = 0.5f * (sc / ma + 1.0f), tt = 0.5f * (tc / ma + 1.0f);
    float u = s * (float)t.res - 0.5f, v = tt * (float)t.res - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = iclamp(cvt_floor(fu), 0, t.res - 1), j0 = iclamp(cvt_floor(fv), 0, t.res - 1);
    int i1 = iclamp(cvt_floor(fu) + 1, 0, t.res - 1), j1 = iclamp(cvt_floor(fv) + 1, 0, t.res - 1);
    const float* f = t.data + (size_t)face * t.res * t.res * 3;
    v4 r;
    float* o = &r.x;
    for (int c = 0; c < 3; ++c) {
        float t00 = f[((size_t)j0 * t.res + i0) * 3 + c], t10 = f[((size_t)j0 * t.res + i1) * 3 + c];
        float t01 = f[((size_t)j1 * t.res + i0) * 3 + c], t11 = f[((size_t)j1 * t.res + i1) * 3 + c];
        float top = t00 * (1.0f - a) + t10 * a;
        float bot = t01 * (1.0f - a) + t11 * a;
        o[c] = top * (1.0f - b) + bot * b;
    }
    r.w = 1.0f;
    return r;
}

}  // namespace vxo
#endif
