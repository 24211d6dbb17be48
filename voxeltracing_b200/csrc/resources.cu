// resources.cu — block texture arrays and the sky cube map: upload, mip-chain construction and the
// host-side evaluation of the per-frame light colours.
//
// TextureArray::CreateArray (Core/GLClasses/TextureArray.cpp:10-69) uploads level 0 and calls
// glGenerateMipmap; the filter GL applies there is implementation defined.  Pinned model:
// 2x2 box filter in linear light (albedo is GL_SRGB_ALPHA: decode, average, re-encode), each level
// re-quantised to 8 bits, alpha averaged linearly.  Built once on the host at set-up time.
#include <math.h>
#include <string.h>

#include "ctx.h"

namespace {

float srgb_decode(int c) {
    double cs = (double)c / 255.0;
    double l = cs <= 0.04045 ? cs / 12.92 : pow((cs + 0.055) / 1.055, 2.4);
    return (float)l;
}
uint8_t srgb_encode(float l) {
    double x = l;
    if (!(x > 0.0)) return 0;
    if (x >= 1.0) return 255;
    double s = x <= 0.0031308 ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
    return (uint8_t)nearbyint(s * 255.0);
}
uint8_t unorm_encode(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 1.0f) return 255;
    return (uint8_t)nearbyintf(f * 255.0f);
}

}  // namespace

int vxrt_set_texture_array(vxrt_ctx* c, int kind, int layers, int w, int h, const uint8_t* rgba8) {
    const bool srgb = kind == VXRT_TEX_ALBEDO;
    float decode[256], lin[256];
    for (int i = 0; i < 256; ++i) { decode[i] = srgb ? srgb_decode(i) : (float)i / 255.0f; lin[i] = (float)i / 255.0f; }

    // level sizes / offsets
    // the validated size limit of vxrt_cuda_set_texture_array is 2048 (12 levels)
    unsigned offsets[12];
    int lw[12], lh[12], levels = 0;
    size_t total = 0;
    for (int cw = w, ch = h;; ) {
        if (levels >= 12) return vxrt_fail(VXRT_E_INVALID, "texture array too large (at most 2048 x 2048)");
        lw[levels] = cw; lh[levels] = ch; offsets[levels] = (unsigned)total;
        total += (size_t)layers * cw * ch * 4;
        ++levels;
        if (cw == 1 && ch == 1) break;
        cw = cw > 1 ? cw / 2 : 1;
        ch = ch > 1 ? ch / 2 : 1;
    }
    if (total > 0xffffffffull) return vxrt_fail(VXRT_E_INVALID, "texture array exceeds 4 GiB");
    std::vector<uint8_t> all;
    try {
        all.resize(total);
    } catch (...) {  // extern "C" callers never see an exception
        return vxrt_fail(VXRT_E_NOMEM, "set_texture_array: out of host memory for the %zu-byte mip chain", total);
    }
    memcpy(all.data(), rgba8, (size_t)layers * w * h * 4);
    for (int l = 1; l < levels; ++l) {
        const uint8_t* src = all.data() + offsets[l - 1];
        uint8_t* dst = all.data() + offsets[l];
        const int sw = lw[l - 1], sh = lh[l - 1], dw = lw[l], dh = lh[l];
        for (int L = 0; L < layers; ++L)
            for (int y = 0; y < dh; ++y)
                for (int x = 0; x < dw; ++x) {
                    const int x0 = (2 * x < sw) ? 2 * x : sw - 1, x1 = (2 * x + 1 < sw) ? 2 * x + 1 : sw - 1;
                    const int y0 = (2 * y < sh) ? 2 * y : sh - 1, y1 = (2 * y + 1 < sh) ? 2 * y + 1 : sh - 1;
                    const uint8_t* p00 = src + (((size_t)L * sh + y0) * sw + x0) * 4;
                    const uint8_t* p10 = src + (((size_t)L * sh + y0) * sw + x1) * 4;
                    const uint8_t* p01 = src + (((size_t)L * sh + y1) * sw + x0) * 4;
                    const uint8_t* p11 = src + (((size_t)L * sh + y1) * sw + x1) * 4;
                    uint8_t* o = dst + (((size_t)L * dh + y) * dw + x) * 4;
                    for (int ch = 0; ch < 3; ++ch) {
                        float s = ((decode[p00[ch]] + decode[p10[ch]]) + (decode[p01[ch]] + decode[p11[ch]])) * 0.25f;
                        o[ch] = srgb ? srgb_encode(s) : unorm_encode(s);
                    }
                    float a = ((lin[p00[3]] + lin[p10[3]]) + (lin[p01[3]] + lin[p11[3]])) * 0.25f;
                    o[3] = unorm_encode(a);
                }
    }
    if (c->d_tex_data[kind]) {
        // the array is unusable from here until the new upload has succeeded
        uint8_t* old = c->d_tex_data[kind];
        c->d_tex_data[kind] = nullptr; c->tex[kind].data = nullptr; c->tex_set[kind] = false;
        VX_CUDA(cudaStreamSynchronize(c->stream));
        VX_CUDA(cudaFree(old));
    }
    if (!c->d_tex_decode[kind]) VX_CUDA(cudaMalloc(&c->d_tex_decode[kind], 256 * sizeof(float)));
    VX_CUDA(cudaMalloc(&c->d_tex_data[kind], total));
    VX_CUDA(cudaMemcpyAsync(c->d_tex_data[kind], all.data(), total, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(c->d_tex_decode[kind], decode, sizeof(decode), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    TexArrayDev& t = c->tex[kind];
    t.data = c->d_tex_data[kind];
    t.decode = c->d_tex_decode[kind];
    for (int l = 0; l < 12; ++l) t.level_offset[l] = l < levels ? offsets[l] : 0u;
    t.w = w; t.h = h; t.layers = layers; t.levels = levels;
    c->tex_set[kind] = true;
    return VXRT_OK;
}

int vxrt_set_skymap(vxrt_ctx* c, int res, const float* f) {
    const size_t n = (size_t)6 * res * res * 3;
    if (c->d_sky) { VX_CUDA(cudaFree(c->d_sky)); c->d_sky = nullptr; }
    VX_CUDA(cudaMalloc(&c->d_sky, n * sizeof(float)));
    VX_CUDA(cudaMemcpyAsync(c->d_sky, f, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->h_sky.assign(f, f + n);
    c->sky.data = c->d_sky;
    c->sky.res = res;
    return VXRT_OK;
}

// texture(u_Skymap, dir) on the host: same face selection / bilinear rule as texcube_sample (texture.cuh)
void vxrt_host_sky_sample(const vxrt_ctx* c, const float d[3], float rgb[3]) {
    const int res = c->sky.res;
    const float x = d[0], y = d[1], z = d[2];
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = x >= 0.0f ? 0 : 1; sc = x >= 0.0f ? -z : z; tc = -y; ma = ax; }
    else if (ay >= az) { face = y >= 0.0f ? 2 : 3; sc = x; tc = y >= 0.0f ? z : -z; ma = ay; }
    else { face = z >= 0.0f ? 4 : 5; sc = z >= 0.0f ? x : -x; tc = -y; ma = az; }
    const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    const float u = s * (float)res - 0.5f, v = t * (float)res - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = u - fu, b = v - fv;
    auto clampi = [res](int i) { return i < 0 ? 0 : (i > res - 1 ? res - 1 : i); };
    const int i0 = clampi((int)fu), j0 = clampi((int)fv), i1 = clampi((int)fu + 1), j1 = clampi((int)fv + 1);
    const float* f = c->h_sky.data() + (size_t)face * res * res * 3;
    for (int ch = 0; ch < 3; ++ch) {
        float t00 = f[((size_t)j0 * res + i0) * 3 + ch], t10 = f[((size_t)j0 * res + i1) * 3 + ch];
        float t01 = f[((size_t)j1 * res + i0) * 3 + ch], t11 = f[((size_t)j1 * res + i1) * 3 + ch];
        float top = t00 * (1.0f - a) + t10 * a;
        float bot = t01 * (1.0f - a) + t11 * a;
        rgb[ch] = top * (1.0f - b) + bot * b;
    }
}

namespace {
// SRGBToLinear / TemperatureToRGB (DiffuseRayTraceFrag.glsl:822-825, 862-886)
float srgb_to_linear_f(float x) { return x > 0.04045f ? powf(x * (1.0f / 1.055f) + 0.0521327f, 2.4f) : x / 12.92f; }
float clamp01(float x) { float lo = (x < 0.0f) ? 0.0f : x; return (1.0f < lo) ? 1.0f : lo; }
void temperature_to_rgb(float K, float out[3]) {
    float r, g, b;
    K = (K < 1000.0f ? 1000.0f : (K > 50000.0f ? 50000.0f : K)) / 100.0f;
    if (K <= 66.0f) {
        r = 1.0f;
        g = clamp01(0.39008157876901960784f * logf(K) - 0.63184144378862745098f);
    } else {
        float t = K - 60.0f;
        r = clamp01(1.29293618606274509804f * powf(t, -0.1332047592f));
        g = clamp01(1.12989086089529411765f * powf(t, -0.0755148492f));
    }
    if (K >= 66.0f) b = 1.0f;
    else if (K <= 19.0f) b = 0.0f;
    else b = clamp01(0.54320678911019607843f * logf(K - 10.0f) - 1.19625408914f);
    out[0] = srgb_to_linear_f(r); out[1] = srgb_to_linear_f(g); out[2] = srgb_to_linear_f(b);
}
}  // namespace

// SampleSunColor (DiffuseRayTraceFrag.glsl:901-908, ReflectionTraceFrag.glsl:649-656)
void vxrt_host_sun_color(const vxrt_ctx* c, const float sun[3], float strength, float rgb[3]) {
    const float PI = 3.14159265359f;
    float sky[3], tm[3];
    vxrt_host_sky_sample(c, sun, sky);
    temperature_to_rgb(5778.0f, tm);
    for (int i = 0; i < 3; ++i) {
        float v = sky[i] * tm[i];
        rgb[i] = v * PI * 2.2f * strength;
    }
}
// SampleMoonColor (ReflectionTraceFrag.glsl:657-665)
void vxrt_host_moon_color(const vxrt_ctx* c, const float moon[3], float strength, float rgb[3]) {
    const float PI = 3.14159265359f;
    float m[3];
    vxrt_host_sky_sample(c, moon, m);
    for (int i = 0; i < 3; ++i) m[i] = m[i] * PI * strength;
    // BasicSaturation(MoonColor, 1.3)
    float lum = (m[0] * 0.2125f + m[1] * 0.7154f) + m[2] * 0.0721f;
    for (int i = 0; i < 3; ++i) {
        float v = lum * (1.0f - 1.3f) + m[i] * 1.3f;
        rgb[i] = v * 0.42525f * strength;
    }
}
