// filter_sampler.cuh — sampling helpers shared by the screen-space filter passes (svgf.cu, shadow_filter.cu): the attachment
// sampler model of texture.cuh (REPEAT, texel centres at +0.5, bilinear weights in full float) with the coordinate set-up done
// once per tap and shared by every image of the same geometry, whole-texel loads, and unorm8 -> float through a table.
#pragma once
#include "texture.cuh"

VXD f3 ray_direction_at(const float* __restrict__ inv_view, const float* __restrict__ inv_proj, f2 ss) {
    f4 clip = F4(ss.x * 2.0f - 1.0f, ss.y * 2.0f - 1.0f, -1.0f, 1.0f);
    f4 e = mat4_mul(inv_proj, clip);
    f4 r = mat4_mul(inv_view, F4(e.x, e.y, -1.0f, 0.0f));
    return F3(r.x, r.y, r.z);
}
VXD bool in_screen_space(f2 v) { return v.x < 1.0f && v.x > 0.0f && v.y < 1.0f && v.y > 0.0f; }
VXD float sh_to_y(float w) { return gmax(0.0f, 3.544905f * w); }

// ---- samplers (the model of texture.cuh: REPEAT, texel centres at +0.5, bilinear weights in full float), with the
// coordinate set-up done once per tap and shared by every image of the same geometry, and texels loaded whole ----
struct Tap {
    int o00, o10, o01, o11;  // pixel offsets of the four texels
    float a, b, ia, ib;      // weights and their complements
};
// REPEAT wrap; every tap of these passes lies in [-1, n] (screen-space tests), anything else takes the general path
VXD int wrap_near(int i, int n) {
    if ((unsigned)(i + 1) <= (unsigned)(n + 1)) return i < 0 ? i + n : (i >= n ? i - n : i);
    return wrap_repeat(i, n);
}
// ---- tolerance mode (set_option "filter_snap", vxrt_cuda_set_option): every tap of these shaders is a texture() read, and where the
// images have the resolution of the pass almost every tap sits on a texel centre up to float rounding, so the bit-faithful path
// blends four texels with weights (1 - eps, eps) - 6 FMUL + 3 FADD and four loads per channel for a result that differs from the
// nearest texel by eps * contrast.  With a snap threshold s > 0 a weight below s (or above 1 - s) becomes exactly 0 (the hardware
// texture unit quantises bilinear weights to 1 / 256 in the same spirit), and a tap whose two weights are both 0 is ONE texel load
// and no arithmetic.  s = 0 (the default) is the bit-faithful parity mode; taps that really lie between texels (reprojection,
// images of another resolution) are blended in either mode.
static __constant__ float g_filter_snap = 0.0f;

// one axis of a tap: the two texel indices and the weight pair
struct Axis {
    int i0, i1;
    float a, ia;
};
VXD Axis make_axis(int n, float s) {
    Axis x;
    const float u = s * (float)n - 0.5f, fu = floorf(u);
    x.a = u - fu; x.ia = 1.0f - x.a;
    x.i0 = wrap_near(cvt_floor(fu), n);
    x.i1 = x.i0 + 1 == n ? 0 : x.i0 + 1;
    const float snap = g_filter_snap;
    if (snap > 0.0f) {
        if (x.a < snap) { x.a = 0.0f; x.ia = 1.0f; }
        else if (x.ia < snap) { x.a = 0.0f; x.ia = 1.0f; x.i0 = x.i1; }
    }
    return x;
}
// a tap that is exactly one texel (both weights 0: a snapped tap, or one that landed on a texel centre exactly - for finite texels
// the blend then returns that texel anyway)
#define VX_TAP_SINGLE(t) ((t).a == 0.0f && (t).b == 0.0f)
VXD Tap join_axes(const Axis& x, const Axis& y, int w) {
    Tap t;
    t.a = x.a; t.ia = x.ia; t.b = y.a; t.ib = y.ia;
    t.o00 = y.i0 * w + x.i0; t.o10 = y.i0 * w + x.i1; t.o01 = y.i1 * w + x.i0; t.o11 = y.i1 * w + x.i1;
    return t;
}
VXD Tap make_tap(int w, int h, f2 uv) { return join_axes(make_axis(w, uv.x), make_axis(h, uv.y), w); }
VXD int nearest_offset(int w, int h, f2 uv) {
    return wrap_near(cvt_floor(uv.y * (float)h), h) * w + wrap_near(cvt_floor(uv.x * (float)w), w);
}
VXD float bl(const Tap& t, float t00, float t10, float t01, float t11) {
    const float top = t00 * t.ia + t10 * t.a;
    const float bot = t01 * t.ia + t11 * t.a;
    return top * t.ib + bot * t.b;
}
VXD float2 h2f(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
VXD float hf(uint16_t v) { return __half2float(__ushort_as_half(v)); }

VXD void sample_rgba16(const uint16_t* __restrict__ img, const Tap& t, float* o) {
    const uint2* p = reinterpret_cast<const uint2*>(img);
    if (VX_TAP_SINGLE(t)) {
        const uint2 q = __ldg(p + t.o00);
        const float2 a = h2f(q.x), b = h2f(q.y);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
        return;
    }
    const uint2 q00 = __ldg(p + t.o00), q10 = __ldg(p + t.o10), q01 = __ldg(p + t.o01), q11 = __ldg(p + t.o11);
    const float2 a00 = h2f(q00.x), a10 = h2f(q10.x), a01 = h2f(q01.x), a11 = h2f(q11.x);
    const float2 b00 = h2f(q00.y), b10 = h2f(q10.y), b01 = h2f(q01.y), b11 = h2f(q11.y);
    o[0] = bl(t, a00.x, a10.x, a01.x, a11.x); o[1] = bl(t, a00.y, a10.y, a01.y, a11.y);
    o[2] = bl(t, b00.x, b10.x, b01.x, b11.x); o[3] = bl(t, b00.y, b10.y, b01.y, b11.y);
}
VXD void sample_rg16(const uint16_t* __restrict__ img, const Tap& t, float* o) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(img);
    if (VX_TAP_SINGLE(t)) {
        const float2 a = h2f(__ldg(p + t.o00));
        o[0] = a.x; o[1] = a.y;
        return;
    }
    const float2 a00 = h2f(__ldg(p + t.o00)), a10 = h2f(__ldg(p + t.o10)), a01 = h2f(__ldg(p + t.o01)), a11 = h2f(__ldg(p + t.o11));
    o[0] = bl(t, a00.x, a10.x, a01.x, a11.x); o[1] = bl(t, a00.y, a10.y, a01.y, a11.y);
}
VXD float sample_r16(const uint16_t* __restrict__ img, const Tap& t) {
    if (VX_TAP_SINGLE(t)) return hf(__ldg(img + t.o00));
    return bl(t, hf(__ldg(img + t.o00)), hf(__ldg(img + t.o10)), hf(__ldg(img + t.o01)), hf(__ldg(img + t.o11)));
}
// one channel of an RGB16F image
VXD float sample_rgb16_ch(const uint16_t* __restrict__ img, const Tap& t, int ch) {
    if (VX_TAP_SINGLE(t)) return hf(__ldg(img + 3 * t.o00 + ch));
    return bl(t, hf(__ldg(img + 3 * t.o00 + ch)), hf(__ldg(img + 3 * t.o10 + ch)), hf(__ldg(img + 3 * t.o01 + ch)), hf(__ldg(img + 3 * t.o11 + ch)));
}
// RG8 through the k / 255 table in shared memory (a float division per texel channel otherwise)
VXD void sample_rg8(const uint8_t* __restrict__ img, const Tap& t, const float* __restrict__ lut, float* o) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(img);
    if (VX_TAP_SINGLE(t)) {
        const uint32_t q = __ldg(p + t.o00);
        o[0] = lut[q & 255]; o[1] = lut[q >> 8];
        return;
    }
    const uint32_t q00 = __ldg(p + t.o00), q10 = __ldg(p + t.o10), q01 = __ldg(p + t.o01), q11 = __ldg(p + t.o11);
    o[0] = bl(t, lut[q00 & 255], lut[q10 & 255], lut[q01 & 255], lut[q11 & 255]);
    o[1] = bl(t, lut[q00 >> 8], lut[q10 >> 8], lut[q01 >> 8], lut[q11 >> 8]);
}
// R8 through the table
VXD float sample_r8(const uint8_t* __restrict__ img, const Tap& t, const float* __restrict__ lut) {
    if (VX_TAP_SINGLE(t)) return lut[__ldg(img + t.o00)];
    return bl(t, lut[__ldg(img + t.o00)], lut[__ldg(img + t.o10)], lut[__ldg(img + t.o01)], lut[__ldg(img + t.o11)]);
}
// dot product of two GetNormalFromID normals given as indices (0..5 = +Z -Z +Y -Y -X +X, 6 = (1, 1, 1)): -1, 0, 1 or 3
VXD float normal_dot(int n0, int n1) {
    if (n0 == 6 && n1 == 6) return 3.0f;
    if (n0 == 6 || n1 == 6) { const int n = n0 == 6 ? n1 : n0; return (n == 0 || n == 2 || n == 5) ? 1.0f : -1.0f; }
    return n0 == n1 ? 1.0f : ((n0 >> 1) == (n1 >> 1) ? -1.0f : 0.0f);
}
// GetNormalFromID (TemporalFilter.glsl:111-123) as an index: 0..5 = the face normals, 6 = (1, 1, 1)
VXD int normal_index(float n) {
    int i = cvt_round(n * 10.0f);
    return i > 5 ? 6 : (i < 0 ? 0 : i);
}
// each translation unit that includes this header has its own copy of g_filter_snap (and of this function and its statics): the
// launchers of the unit call it before every launch; the constant is only rewritten when the context's setting differs from what the
// unit last wrote on that device
static int vx_apply_filter_snap(vxrt_ctx* c) {
    static float applied[16];
    static bool known[16];
    const int d = c->device & 15;
    if (!known[d] || applied[d] != c->filter_snap) {
        const float v = c->filter_snap;
        VX_CUDA(cudaMemcpyToSymbolAsync(g_filter_snap, &v, sizeof(float), 0, cudaMemcpyHostToDevice, c->stream));
        VX_CUDA(cudaStreamSynchronize(c->stream));   // v is a stack variable
        applied[d] = v; known[d] = true;
    }
    return VXRT_OK;
}

// k / 255 for k = 0..255 (unorm8 -> float exactly as the samplers define it); blockDim.x == 256
VXD void fill_unorm_lut(float* lut) {
    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
}

