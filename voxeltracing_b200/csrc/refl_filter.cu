// refl_filter.cu — reflection temporal filter: Core/Shaders/SpecularTemporalFilter.glsl, dispatched at
// Core/Pipeline.cpp:3316-3400.  Hit-distance reprojection of the reflected ray, neighbourhood clipping of the history
// (ReflectionClipping), firefly rejection next to emissive hits, accumulation factor from the screen-space velocity.  One thread
// per pixel, a warp covers an 8 x 4 pixel tile; samplers of filter_sampler.cuh.  The trace images, the temporal images, the
// primary G-buffer and the material G-buffer may each have their own size.  The only transcendental is one expf per pixel:
// R16F outputs agree with the oracle to one half-ulp (tests/test_gpu_refl_filter.py).
#include "ctx.h"
#include "filter_sampler.cuh"

namespace {

struct Img8 { const uint8_t* __restrict__ p; int w, h; };
struct Img16 { const uint16_t* __restrict__ p; int w, h; };

struct SpecTemporalArgs {
    float inv_view[16], inv_proj[16], prev_pv[16];
    float cam_cur[3], cam_prev[3];
    int width, height, row0, row1;
    int temporal_spec, firefly, aggressive, smart_clip, rough_weight, stabilize;
    Img16 cur_color;   // ReflectionTraceFBO[0] RGBA16F
    Img16 cur_hit;     // [1] R16F
    Img8 cur_mask;     // [2] R8
    Img16 prev_hit;    // previous frame's [1]
    Img16 hist_color;  // previous temporal set +0
    Img16 hist_hit;    // previous temporal set +2
    Img16 g_t, prev_t;
    Img8 g_n, prev_n;
    Img8 pbr;          // GeneratedGBuffer[2] RGBA8
    uint16_t* __restrict__ out_color;
    uint16_t* __restrict__ out_frames;
    uint16_t* __restrict__ out_hit;
};

struct c4 { float x, y, z, w; };
VXD c4 sample4(const Img16& im, f2 uv) {
    float o[4];
    sample_rgba16(im.p, make_tap(im.w, im.h, uv), o);
    c4 r; r.x = o[0]; r.y = o[1]; r.z = o[2]; r.w = o[3];
    return r;
}
VXD float sample1(const Img16& im, f2 uv) { return sample_r16(im.p, make_tap(im.w, im.h, uv)); }
VXD float sample1(const Img8& im, f2 uv, const float* lut) { return sample_r8(im.p, make_tap(im.w, im.h, uv), lut); }
VXD int normal_at(const Img8& im, f2 uv, const float* lut) { return normal_index(lut[__ldg(im.p + nearest_offset(im.w, im.h, uv))]); }
// .xy of a bilinear RGBA8 sample
VXD f2 sample_pbr_xy(const Img8& im, f2 uv, const float* lut) {
    const Tap t = make_tap(im.w, im.h, uv);
    const uint32_t* p = reinterpret_cast<const uint32_t*>(im.p);
    if (VX_TAP_SINGLE(t)) {
        const uint32_t q = __ldg(p + t.o00);
        return F2(lut[q & 255], lut[(q >> 8) & 255]);
    }
    const uint32_t q00 = __ldg(p + t.o00), q10 = __ldg(p + t.o10), q01 = __ldg(p + t.o01), q11 = __ldg(p + t.o11);
    return F2(bl(t, lut[q00 & 255], lut[q10 & 255], lut[q01 & 255], lut[q11 & 255]),
              bl(t, lut[(q00 >> 8) & 255], lut[(q10 >> 8) & 255], lut[(q01 >> 8) & 255], lut[(q11 >> 8) & 255]));
}
VXD f3 xyz(const c4& v) { return F3(v.x, v.y, v.z); }
VXD void set_xyz(c4& v, f3 a) { v.x = a.x; v.y = a.y; v.z = a.z; }
VXD c4 add4(c4 a, float s) { a.x += s; a.y += s; a.z += s; a.w += s; return a; }
VXD float dist_sq(f3 a, f3 b) { const f3 c = a - b; return dot(c, c); }
VXD f2 project_prev(const float* pv, f3 pos) {
    const f4 P = mat4_mul(pv, F4(pos.x, pos.y, pos.z, 1.0f));
    return F2((P.x / P.w) * 0.5f + 0.5f, (P.y / P.w) * 0.5f + 0.5f);
}
// ClipToAABB (:126-135)
VXD f3 clip_to_aabb(f3 prev, f3 mn, f3 mx) {
    const f3 pClip = 0.5f * (mx + mn), eClip = 0.5f * (mx - mn);
    const f3 vClip = prev - pClip, vUnit = vClip / eClip;
    const float denom = gmax(fabsf(vUnit.x), gmax(fabsf(vUnit.y), fabsf(vUnit.z)));
    return denom > 1.0f ? pClip + vClip / denom : prev;
}

// SpecularTemporalFilter.glsl main() (:285-417)
__global__ void __launch_bounds__(256) specular_temporal_kernel(const __grid_constant__ SpecTemporalArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const f3 origin = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const float CurDist = sample1(a.g_t, tc);
    const f3 CurPos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * CurDist;   // GetPositionAt (:94-98)
    const int InitialNormal = normal_at(a.g_n, tc, lut);
    c4 CurrentColor = sample4(a.cur_color, tc);
    float oFrames = 0.0f, oHit;
    c4 oColor;
    const float HitDistanceCurrent = sample1(a.cur_hit, tc);
    const f2 TexelSize = F2(1.0f / (float)a.cur_color.w, 1.0f / (float)a.cur_color.h);
    if (a.firefly && sample1(a.cur_mask, tc, lut) > 0.05f) {   // FireflyReject (:246-277)
        const int SampleThreshold = a.aggressive ? 3 : 4;
        c4 NonLit; NonLit.x = NonLit.y = NonLit.z = NonLit.w = 0.0f;
        int Unlit = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float ox = i == 0 ? 1.0f : (i == 2 ? -1.0f : 0.0f), oy = i == 1 ? 1.0f : (i == 3 ? -1.0f : 0.0f);
            const f2 sc = F2(tc.x + ox * TexelSize.x, tc.y + oy * TexelSize.y);
            const float Mask = sample1(a.cur_mask, sc, lut);
            const c4 Color = sample4(a.cur_color, sc);
            if (Mask < 0.01f) { NonLit.x += Color.x; NonLit.y += Color.y; NonLit.z += Color.z; NonLit.w += Color.w; ++Unlit; }
        }
        if (Unlit >= SampleThreshold) {
            const float n = (float)Unlit;
            CurrentColor.x = NonLit.x / n; CurrentColor.y = NonLit.y / n; CurrentColor.z = NonLit.z / n; CurrentColor.w = NonLit.w / n;
        }
    }
    if (CurDist > 0.0f && a.temporal_spec) {
        const bool SkySample = HitDistanceCurrent < 0.0f;
        const f2 pxy = sample_pbr_xy(a.pbr, tc, lut);
        const float RoughnessAt = gmix(0.095f, pxy.x, a.rough_weight ? 1.0f : 0.0f);
        const float MetalnessAt = pxy.y;
        bool LessValid = false;
        f2 R = F2(0.0f, 0.0f);
        if (HitDistanceCurrent > 0.0f && !SkySample && RoughnessAt <= 0.875f + 0.01f) {   // reproject along the reflected ray
            const f3 I = normalize(origin - CurPos);
            R = project_prev(a.prev_pv, CurPos - I * HitDistanceCurrent);
            const float PreviousT = sample1(a.prev_hit, R);
            if (fabsf(PreviousT - HitDistanceCurrent) >= 3.8f) LessValid = true;
        } else if (!SkySample) {
            f3 CameraOffset = F3(a.cam_cur[0], a.cam_cur[1], a.cam_cur[2]) - F3(a.cam_prev[0], a.cam_prev[1], a.cam_prev[2]);
            CameraOffset = CameraOffset * 0.6f;
            R = project_prev(a.prev_pv, CurPos - CameraOffset);
        }
        if (SkySample) {
            const f3 I = normalize(origin - CurPos);
            R = project_prev(a.prev_pv, CurPos - I * 64.0f);
        }
        const float PrevDist = sample1(a.prev_t, R);
        const f3 PrevPos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, R)) * PrevDist;
        const float d = fabsf(distance(PrevPos, CurPos));
        const float Bias = 0.01f;
        const int PrevNormal = normal_at(a.prev_n, R, lut);
        if (R.x > 0.0f + Bias && R.x < 1.0f - Bias && R.y > 0.0f + Bias && R.y < 1.0f - Bias && d < 1.0f && PrevNormal == InitialNormal) {
            c4 PrevColor = sample4(a.hist_color, R);
            const f3 BasePrevColor = xyz(PrevColor);
            const f3 CamCur = F3(a.cam_cur[0], a.cam_cur[1], a.cam_cur[2]), CamPrev = F3(a.cam_prev[0], a.cam_prev[1], a.cam_prev[2]);
            const bool Moved = dist_sq(CamCur, CamPrev) > 0.0001f;
            const bool TryClipping = RoughnessAt < 0.5f + 0.01f;
            if (TryClipping && Moved && a.smart_clip) {   // ReflectionClipping (:143-225), called with v_TexCoords
                const float RoughnessThreshold = 0.275f + 0.01f;
                c4 MinColor, MaxColor;
                MinColor.x = MinColor.y = MinColor.z = MinColor.w = 1000.0f;
                MaxColor.x = MaxColor.y = MaxColor.z = MaxColor.w = -1000.0f;
                float AdditionalMaxBias = 0.0f;
#pragma unroll 1
                for (int x = -1; x <= 1; ++x)
#pragma unroll 1
                    for (int y = -1; y <= 1; ++y) {
                        const f2 sc = F2(tc.x + (float)x * TexelSize.x, tc.y + (float)y * TexelSize.y);
                        if (!(sample1(a.cur_mask, sc, lut) < 0.01f)) AdditionalMaxBias += 0.1f;
                        const c4 S = sample4(a.cur_color, sc);
                        MinColor.x = gmin(S.x, MinColor.x); MinColor.y = gmin(S.y, MinColor.y); MinColor.z = gmin(S.z, MinColor.z); MinColor.w = gmin(S.w, MinColor.w);
                        MaxColor.x = gmax(S.x, MaxColor.x); MaxColor.y = gmax(S.y, MaxColor.y); MaxColor.z = gmax(S.z, MaxColor.z); MaxColor.w = gmax(S.w, MaxColor.w);
                    }
                const f3 OriginalMin = xyz(MinColor), OriginalMax = xyz(MaxColor);
                const bool Smoothish = RoughnessAt < RoughnessThreshold;
                const bool Roughish = RoughnessAt > RoughnessThreshold && RoughnessAt < 0.50f + 0.01f;
                if (Smoothish) {
                    const float Perceived = RoughnessAt * RoughnessAt;
                    const float RT2 = RoughnessThreshold * RoughnessThreshold;
                    const float Remapped = 0.0f + (gclamp((Perceived - 0.0f) / (RT2 - 0.0f), 0.0f, 1.0f) * (1.0f - 0.0f));   // remap (:121-124)
                    float B = gmix(0.01f, 0.085f, Remapped);
                    if (RoughnessAt > 0.235f) B *= 1.55f;
                    MinColor = add4(MinColor, -(B * 0.95f));
                    MaxColor = add4(MaxColor, (B * 0.95f) + AdditionalMaxBias);
                } else if (Roughish) {
                    MinColor = add4(MinColor, -0.37f);
                    MaxColor = add4(MaxColor, 0.37f + (AdditionalMaxBias * 1.1f));
                }
                const float m = gmix(0.05f, 0.25f, MetalnessAt > 0.05f ? 1.0f : 0.0f);
                set_xyz(MinColor, gmix(xyz(MinColor), OriginalMin, m));
                set_xyz(MaxColor, gmix(xyz(MaxColor), OriginalMax, m));
                const float BiasMixer = LessValid ? 0.5f : 0.0f;
                set_xyz(MinColor, gmix(xyz(MinColor), OriginalMin, BiasMixer));
                set_xyz(MaxColor, gmix(xyz(MaxColor), OriginalMax, BiasMixer));
                const f3 Prev3 = xyz(PrevColor);
                const f3 Clamped = clip_to_aabb(Prev3, xyz(MinColor), xyz(MaxColor));
                if (Clamped.x != Prev3.x || Clamped.y != Prev3.y || Clamped.z != Prev3.z)
                    PrevColor = dist_sq(Clamped, xyz(MinColor)) > dist_sq(Clamped, xyz(MaxColor)) ? MaxColor : MinColor;
            }
            // GetAccumulationFactor (:279-283)
            const float vx = (tc.x - R.x) * (float)a.cur_color.w, vy = (tc.y - R.y) * (float)a.cur_color.h;
            float AF = gclamp(expf(-sqrtf(vx * vx + vy * vy)) * 0.9f + 0.750f, 0.00000001f, 0.96f);
            AF = gclamp(AF, 0.001f, 0.95f);
            CurrentColor.x = gmax(CurrentColor.x, 0.0f); CurrentColor.y = gmax(CurrentColor.y, 0.0f);
            CurrentColor.z = gmax(CurrentColor.z, 0.0f); CurrentColor.w = gmax(CurrentColor.w, 0.0f);
            PrevColor.x = gmax(PrevColor.x, 0.0f); PrevColor.y = gmax(PrevColor.y, 0.0f); PrevColor.z = gmax(PrevColor.z, 0.0f); PrevColor.w = gmax(PrevColor.w, 0.0f);
            oColor.x = gmix(CurrentColor.x, PrevColor.x, AF); oColor.y = gmix(CurrentColor.y, PrevColor.y, AF);
            oColor.z = gmix(CurrentColor.z, PrevColor.z, AF); oColor.w = gmix(CurrentColor.w, PrevColor.w, AF);
            oFrames = AF;
            oHit = HitDistanceCurrent;
            if (dist_sq(BasePrevColor, xyz(PrevColor)) < 0.2f && a.stabilize)
                oHit = gmix(HitDistanceCurrent, sample1(a.hist_hit, R), gclamp(AF * 1.1f, 0.0f, 0.9f));
        } else {
            oColor = CurrentColor; oFrames = 0.0f; oHit = HitDistanceCurrent;
        }
    } else {
        oColor = CurrentColor; oHit = HitDistanceCurrent;
    }
    if (!a.temporal_spec) oFrames = -1.0f;
    const size_t i = (size_t)py * a.width + px;
    uint2 packed;
    packed.x = (uint32_t)float_to_half_bits(oColor.x) | ((uint32_t)float_to_half_bits(oColor.y) << 16);
    packed.y = (uint32_t)float_to_half_bits(oColor.z) | ((uint32_t)float_to_half_bits(oColor.w) << 16);
    reinterpret_cast<uint2*>(a.out_color)[i] = packed;
    a.out_frames[i] = float_to_half_bits(oFrames);
    a.out_hit[i] = float_to_half_bits(oHit);
}

// ---- spatial pass: ReflectionDenoiserNew.glsl main() (:97-364), one direction per launch ------------------------------------
struct ReflDenoiseArgs {
    float inv_view[16], inv_proj[16], view[16];
    int width, height, row0, row1;
    int dir, roughness_bias, normal_map_aware, handle_lobe_deviation, derive_from_diffuse_sh, amplify, temporal_weight, radius_bias;
    float normal_map_weight_strength, denoiser_scale, resolution_scale, rnw_bias_strength;
    Img16 in_color;   // u_InputTexture RGBA16F
    Img16 frames;     // u_Frames R16F
    Img16 hit;        // u_SpecularHitData R16F
    Img16 g_t;
    Img8 g_n;         // (u_BlockIDTex only feeds BlockValidity (:251), which nothing reads)
    Img16 gb_normal;  // GeneratedGBuffer[1] RGB16F
    Img8 pbr;         // GeneratedGBuffer[2] RGBA8
    uint16_t* __restrict__ out;
};

__constant__ float c_gauss[33] = {0.004013f, 0.005554f, 0.007527f, 0.00999f, 0.012984f, 0.016524f, 0.020594f, 0.025133f, 0.030036f, 0.035151f, 0.040283f,
                                  0.045207f, 0.049681f, 0.053463f, 0.056341f, 0.058141f, 0.058754f, 0.058141f, 0.056341f, 0.053463f, 0.049681f, 0.045207f,
                                  0.040283f, 0.035151f, 0.030036f, 0.025133f, 0.020594f, 0.016524f, 0.012984f, 0.00999f, 0.007527f, 0.005554f, 0.004013f};

VXD f3 sample_rgb16(const Img16& im, const Tap& t) {
    return F3(sample_rgb16_ch(im.p, t, 0), sample_rgb16_ch(im.p, t, 1), sample_rgb16_ch(im.p, t, 2));
}
VXD float luma(f3 c) { return dot(c, F3(0.299f, 0.587f, 0.114f)); }
// pow(x, n) for the integer exponents of the shader, by multiplication (within the error of the general powf, whose accuracy the
// reference leaves to the driver; inf and 0 behave like pow's)
VXD float pow4(float x) { const float x2 = x * x; return x2 * x2; }
VXD float pow12(float x) { const float x2 = x * x, x4 = x2 * x2; return (x4 * x4) * x4; }
VXD float pow16(float x) { const float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4; return x8 * x8; }

__global__ void __launch_bounds__(256) reflection_denoise_kernel(const __grid_constant__ ReflDenoiseArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const f3 origin = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const float BaseDist = sample1(a.g_t, tc);
    const f3 BasePos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * BaseDist;
    const int BaseNormal = normal_at(a.g_n, tc, lut);
    const bool BaseIsSky = BaseDist < 0.0f;
    const c4 BaseColor = sample4(a.in_color, tc);
    const float BaseLuminance = luma(xyz(BaseColor));
    float TotalWeight = 0.0f;
    const bool Dir = a.dir != 0;
    const float TexelSize = Dir ? 1.0f / (float)a.width : 1.0f / (float)a.height;   // u_Dimensions = the output's size
    const f2 pxy = sample_pbr_xy(a.pbr, tc, lut);
    float BaseRoughness = pxy.x;
    const float RawRoughness = BaseRoughness;
    BaseRoughness *= gmix(1.0f, 0.91f, a.roughness_bias ? 1.0f : 0.0f);
    const f3 NormalMappedBase = sample_rgb16(a.gb_normal, make_tap(a.gb_normal.w, a.gb_normal.h, tc));
    float HitDistanceFetch = sample1(a.hit, tc) + 0.0001f;
    if (HitDistanceFetch < 0.001f) HitDistanceFetch = 1.75f;
    else HitDistanceFetch = powf(HitDistanceFetch, 1.0f / 1.3f);
    if (a.handle_lobe_deviation) {
        if (BaseRoughness <= 0.25f + 0.05f) HitDistanceFetch = gclamp(HitDistanceFetch, 0.0f, 8.0f);
        if (BaseRoughness <= 0.2f) HitDistanceFetch = gclamp(HitDistanceFetch, 0.0f, 5.5f);
    }
    const float SpecularHitDistance = gmax(HitDistanceFetch, 0.01f) * (RawRoughness < 0.51f ? 0.5f : 0.85f);
    const f4 vs = mat4_mul(a.view, F4(BasePos.x, BasePos.y, BasePos.z, 1.0f));
    const float ViewLength = length(F3(vs.x, vs.y, vs.z));
    float ViewLengthWeight = 0.001f + ViewLength;
    if (BaseRoughness > 0.135f) ViewLengthWeight = gmax(ViewLengthWeight, 0.750f);
    else ViewLengthWeight = gmax(ViewLengthWeight, 3.0f);
    if (BaseRoughness < 0.125f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 6.0f);
    else if (BaseRoughness < 0.25f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 8.0f + 1.0f);
    else if (BaseRoughness < 0.5f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 16.0f);
    else if (BaseRoughness < 0.75f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 24.0f);
    else ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 32.0f);
    float TransversalContrib = SpecularHitDistance / gmax((SpecularHitDistance + ViewLengthWeight), 0.00001f);
    if (RawRoughness < 0.535f && a.amplify && BaseDist < 50.0f) {
        const float Remapped = (((RawRoughness - 0.0f) / (0.535f - 0.0f)) * (1.0f - 0.0f)) + 0.0f;   // remap (:93-96)
        const float TransversalExponent = gmix(3.5f, 2.0f, pow4(Remapped));
        TransversalContrib = powf(TransversalContrib, TransversalExponent + 0.8125f);
    }
    const float RadiusExponent = powf((1.0f - BaseRoughness), 1.0f / 1.4f) * 5.0f;
    const float RadiusPow = gclamp(powf(gmix(1.0f * BaseRoughness, 1.0f, TransversalContrib), RadiusExponent), 0.0f, 1.0f);
    const float Radius = RadiusPow, NormalMapRadius = 1.0f - RadiusPow;
    int EffectiveRadius = cvt_floor(Radius * 15.0f);
    EffectiveRadius = iclamp(EffectiveRadius, 1, 15);
    EffectiveRadius = BaseRoughness > 0.897511f ? 15 : EffectiveRadius;
    float Scale = gmix(1.0f, 2.0f, gclamp(a.resolution_scale, 0.0000001f, 1.0f)) + 0.5f;
    int RadiusBias = 0;
    if (pxy.y > 0.1f - 0.001f && BaseRoughness > 0.4f - 0.001f) RadiusBias += 2;
    EffectiveRadius = iclamp(EffectiveRadius + RadiusBias + a.radius_bias, 1, 15);
    if (RawRoughness >= 0.5f - 0.01f) EffectiveRadius += 1;
    if (a.derive_from_diffuse_sh && RawRoughness >= 0.865f) { EffectiveRadius = 4; Scale *= 1.25f; }
    Scale *= a.denoiser_scale;
    float TemporalWeight = 0.0f, AccumulatedFramesClamped = 0.01f;
    if (a.temporal_weight) {
        const float AccumulatedFrames = sample1(a.frames, tc);
        AccumulatedFramesClamped = AccumulatedFrames < -0.1f ? 0.0f : (1.0f - AccumulatedFrames);
        AccumulatedFramesClamped = gclamp(AccumulatedFramesClamped, 0.000001f, 1.0f);
        TemporalWeight = gclamp(AccumulatedFramesClamped * 0.85f, 0.0f, 1.0f);
        float FLT_radius = (float)EffectiveRadius;
        FLT_radius = gmix(FLT_radius, FLT_radius + 2.0f, AccumulatedFramesClamped * 1.05f);
        EffectiveRadius = __float2int_rz(FLT_radius);
    }
    float HF_e = 64.0f * a.normal_map_weight_strength * 1.350f;
    HF_e *= powf(NormalMapRadius, 1.0f / 1.33f);
    float HF_WeightAdder = BaseRoughness > 0.45f ? 0.005f : 0.0f;   // mix(0, c, float(cond)) is exactly c or 0
    HF_WeightAdder += BaseRoughness > 0.525f ? 0.0125f : 0.0f;
    HF_WeightAdder += BaseRoughness > 0.625f ? 0.022f : 0.0f;
    HF_WeightAdder += BaseRoughness > 0.725f ? 0.026f : 0.0f;
    HF_WeightAdder += BaseRoughness > 0.75f ? 0.031f : 0.0f;
    const float HF_bias = HF_WeightAdder * a.rnw_bias_strength * 1.4f;
    EffectiveRadius = iclamp(EffectiveRadius, 1, 15);
    if (RawRoughness < 0.002f) EffectiveRadius = 0;
    const bool hf_pixel = a.normal_map_aware && BaseRoughness < 0.8f && AccumulatedFramesClamped <= 0.185f + 0.001f + 0.001f + 0.0001f;
    // The taps move along one axis only: the other axis of every sampler is set up once per pixel, and images of the same size
    // (G-buffer planes; input colour; the two material planes) share the moving axis too.
    const bool sameI = a.in_color.w == a.g_t.w && a.in_color.h == a.g_t.h, sameM = a.pbr.w == a.g_t.w && a.pbr.h == a.g_t.h;
    const bool sameN = a.gb_normal.w == a.pbr.w && a.gb_normal.h == a.pbr.h;
    const float fixed = Dir ? tc.y : tc.x, moving0 = Dir ? tc.x : tc.y;
    const Axis fG = make_axis(Dir ? a.g_t.h : a.g_t.w, fixed);
    const Axis fI = sameI ? fG : make_axis(Dir ? a.in_color.h : a.in_color.w, fixed);
    const Axis fM = sameM ? fG : make_axis(Dir ? a.pbr.h : a.pbr.w, fixed);
    const Axis fB = sameN ? fM : make_axis(Dir ? a.gb_normal.h : a.gb_normal.w, fixed);
    const int fNrm = wrap_near(cvt_floor(fixed * (float)(Dir ? a.g_n.h : a.g_n.w)), Dir ? a.g_n.h : a.g_n.w);
    const uint32_t* pp = reinterpret_cast<const uint32_t*>(a.pbr.p);
    const float lw_of_one = gclamp(gmix(1.0f, 1.0f, TemporalWeight), 0.0000000001f, 1.0f);   // the luminance weight of a tap whose clamped pow() is 1
    c4 Filtered; Filtered.x = Filtered.y = Filtered.z = Filtered.w = 0.0f;
#pragma unroll 1
    for (int Sample = -EffectiveRadius; Sample <= EffectiveRadius; ++Sample) {
        const float m = moving0 + ((float)Sample * Scale) * TexelSize;
        const float bias = 0.01f;
        if (!(m > 0.0f + bias && m < 1.0f - bias && fixed > 0.0f + bias && fixed < 1.0f - bias)) continue;
        const Axis mG = make_axis(Dir ? a.g_t.w : a.g_t.h, m);
        const float SampleDepth = sample_r16(a.g_t.p, Dir ? join_axes(mG, fG, a.g_t.w) : join_axes(fG, mG, a.g_t.w));
        if ((SampleDepth < 0.0f) != BaseIsSky) continue;
        const Axis mI = sameI ? mG : make_axis(Dir ? a.in_color.w : a.in_color.h, m);
        float sd[4];
        sample_rgba16(a.in_color.p, Dir ? join_axes(mI, fI, a.in_color.w) : join_axes(fI, mI, a.in_color.w), sd);
        const float DepthDifference = fabsf(SampleDepth - BaseDist) * 1.5f;
        const float ed = expf(-DepthDifference);
        const float DepthWeight = ed * ed;   // pow(x, 2.0f)
        // pow(max(dot, 1e-11), 32): 0 (underflow), 1, or 3^32
        const int mNrm = wrap_near(cvt_floor(m * (float)(Dir ? a.g_n.w : a.g_n.h)), Dir ? a.g_n.w : a.g_n.h);
        const int SampleNormal = normal_index(lut[__ldg(a.g_n.p + (Dir ? fNrm * a.g_n.w + mNrm : mNrm * a.g_n.w + fNrm))]);
        const float nd = normal_dot(BaseNormal, SampleNormal);
        const float NormalWeight = nd <= 0.0f ? 0.0f : (nd == 1.0f ? 1.0f : 1853020153315328.0f);
        float LuminanceWeight = 1.0f;
        const Axis mM = sameM ? mG : make_axis(Dir ? a.pbr.w : a.pbr.h, m);
        const Tap tp = Dir ? join_axes(mM, fM, a.pbr.w) : join_axes(fM, mM, a.pbr.w);
        const float SampleRoughness = VX_TAP_SINGLE(tp) ? lut[__ldg(pp + tp.o00) & 255]
                                                        : bl(tp, lut[__ldg(pp + tp.o00) & 255], lut[__ldg(pp + tp.o10) & 255], lut[__ldg(pp + tp.o01) & 255], lut[__ldg(pp + tp.o11) & 255]);
        const bool SampleTooRough = SampleRoughness >= 0.89f;
        if (!SampleTooRough) {
            const float LumaAt = luma(F3(sd[0], sd[1], sd[2]));
            float LuminanceError = 1.0f / fabsf(LumaAt - BaseLuminance);
            if (LuminanceError >= 1.0f) {
                // luminances less than 1 apart (every tap but fireflies): pow(x >= 1, y > 0) >= 1 twice, so the clamp yields exactly 1
                LuminanceWeight = lw_of_one;
            } else {
                LuminanceError = powf(LuminanceError, 1.7f);
                const float LumaWeightExponent = gmix(0.001f, 8.0f, pow16(SampleRoughness));
                LuminanceWeight = powf(LuminanceError, LumaWeightExponent + 0.8f);
                LuminanceWeight = gclamp(LuminanceWeight, 0.0000000001f, 1.0f);
                LuminanceWeight = gmix(LuminanceWeight, 1.0f, TemporalWeight);
                LuminanceWeight = gclamp(LuminanceWeight, 0.0000000001f, 1.0f);
            }
        }
        float HFNormalWeight = 1.0f;
        if (hf_pixel && !SampleTooRough) {
            const Axis mB = sameN ? mM : make_axis(Dir ? a.gb_normal.w : a.gb_normal.h, m);
            const f3 NormalMapAt = sample_rgb16(a.gb_normal, sameN ? tp : (Dir ? join_axes(mB, fB, a.gb_normal.w) : join_axes(fB, mB, a.gb_normal.w)));
            const float Angle = dot(NormalMapAt, NormalMappedBase);
            HFNormalWeight = powf(gclamp(Angle, 0.00000001f, 1.0f), HF_e);
            HFNormalWeight = gclamp(HFNormalWeight + HF_bias, 0.00000000001f, 1.0f);
        }
        const float RoughnessError = fabsf(SampleRoughness - BaseRoughness);
        float RoughnessTransversalWeight = pow12(1.0f / RoughnessError);
        RoughnessTransversalWeight = gclamp(RoughnessTransversalWeight, 0.00000000001f, 1.0f);
        const float CurrentKernelWeight = c_gauss[iclamp(16 + Sample, 0, 32)];
        float CurrentWeight = 1.0f;
        CurrentWeight *= DepthWeight;
        CurrentWeight *= NormalWeight;
        CurrentWeight *= HFNormalWeight;
        CurrentWeight *= LuminanceWeight;
        CurrentWeight *= RoughnessTransversalWeight;
        CurrentWeight *= CurrentKernelWeight;
        CurrentWeight = gclamp(CurrentWeight, 0.000000001f, 1.0f);
        Filtered.x += sd[0] * CurrentWeight; Filtered.y += sd[1] * CurrentWeight;
        Filtered.z += sd[2] * CurrentWeight; Filtered.w += sd[3] * CurrentWeight;
        TotalWeight += CurrentWeight;
    }
    c4 o = BaseColor;
    if (TotalWeight > 0.001f && !(RawRoughness < 0.002f)) {
        Filtered.x /= TotalWeight; Filtered.y /= TotalWeight; Filtered.z /= TotalWeight; Filtered.w /= TotalWeight;
        float Smooth = 1.0f;
        if (BaseRoughness <= 0.1f + 0.007f) {
            Smooth = BaseRoughness * 16.0f;
            Smooth = 1.0f - Smooth;
            Smooth = pow4(Smooth);
            Smooth = gclamp(Smooth, 0.1f, 0.999f);
        }
        o.x = gmix(BaseColor.x, Filtered.x, Smooth); o.y = gmix(BaseColor.y, Filtered.y, Smooth);
        o.z = gmix(BaseColor.z, Filtered.z, Smooth); o.w = gmix(BaseColor.w, Filtered.w, Smooth);
    }
    uint2 packed;
    packed.x = (uint32_t)float_to_half_bits(o.x) | ((uint32_t)float_to_half_bits(o.y) << 16);
    packed.y = (uint32_t)float_to_half_bits(o.z) | ((uint32_t)float_to_half_bits(o.w) << 16);
    reinterpret_cast<uint2*>(a.out)[(size_t)py * a.width + px] = packed;
}

inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}
inline bool is_refl_set(int id) { return id == VXRT_ATT_REFL_TEMPORAL_A || id == VXRT_ATT_REFL_TEMPORAL_B; }

template <typename I>
int image_in(vxrt_ctx* c, const char* fn, int id, int bpp, I* img) {
    const Attachment& a = c->att[id];
    if (!a.ptr || a.width <= 0) return vxrt_fail(VXRT_E_STATE, "%s: attachment %d has not been written", fn, id);
    if (a.bpp != bpp) return vxrt_fail(VXRT_E_STATE, "%s: attachment %d has %d bytes per pixel, expected %d", fn, id, a.bpp, bpp);
    img->p = (decltype(img->p))a.ptr; img->w = a.width; img->h = a.height;
    return VXRT_OK;
}

// an image of the previous frame that does not exist yet (first frame) starts out zero-filled, like the engine's FBOs
int zero_if_missing(vxrt_ctx* c, int id, int w, int h, int bpp) {
    const Attachment& a = c->att[id];
    if (a.ptr && a.width == w && a.height == h && a.bpp == bpp) return VXRT_OK;
    int rc = vxrt_ensure_attachment(c, id, w, h, bpp);
    if (rc) return rc;
    VX_CUDA(cudaMemsetAsync(c->att[id].ptr, 0, (size_t)w * h * bpp, c->stream));
    return VXRT_OK;
}

}  // namespace

int vxrt_launch_specular_temporal(vxrt_ctx* c, const vxrt_specular_temporal_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_specular_temporal";
    if (!is_refl_set(p.history_set) || !is_refl_set(p.out_set) || p.history_set == p.out_set)
        return vxrt_fail(VXRT_E_INVALID, "%s: history_set / out_set must be the two of VXRT_ATT_REFL_TEMPORAL_A / _B", fn);
    SpecTemporalArgs a;
    int rc;
    if ((rc = image_in(c, fn, VXRT_ATT_REFL_COLOR, 8, &a.cur_color))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_REFL_HITDIST, 2, &a.cur_hit))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_REFL_EMISSIVE, 1, &a.cur_mask))) return rc;
    if (a.cur_hit.w != a.cur_color.w || a.cur_hit.h != a.cur_color.h || a.cur_mask.w != a.cur_color.w || a.cur_mask.h != a.cur_color.h)
        return vxrt_fail(VXRT_E_STATE, "%s: the reflection trace images differ in size", fn);
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_T, 2, &a.g_t))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_NORMAL, 1, &a.g_n))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_GBUF_PBR, 4, &a.pbr))) return rc;
    if ((rc = zero_if_missing(c, p.history_set, p.width, p.height, 8))) return rc;
    if ((rc = zero_if_missing(c, p.history_set + 1, p.width, p.height, 2))) return rc;
    if ((rc = zero_if_missing(c, p.history_set + 2, p.width, p.height, 2))) return rc;
    if ((rc = zero_if_missing(c, VXRT_ATT_PREV_INITIAL_T, a.g_t.w, a.g_t.h, 2))) return rc;
    if ((rc = zero_if_missing(c, VXRT_ATT_PREV_INITIAL_NORMAL, a.g_n.w, a.g_n.h, 1))) return rc;
    if ((rc = zero_if_missing(c, VXRT_ATT_PREV_REFL_HITDIST, a.cur_hit.w, a.cur_hit.h, 2))) return rc;
    if ((rc = image_in(c, fn, p.history_set, 8, &a.hist_color))) return rc;
    if ((rc = image_in(c, fn, p.history_set + 2, 2, &a.hist_hit))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_PREV_INITIAL_T, 2, &a.prev_t))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_PREV_INITIAL_NORMAL, 1, &a.prev_n))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_PREV_REFL_HITDIST, 2, &a.prev_hit))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_set, p.width, p.height, 8))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_set + 1, p.width, p.height, 2))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_set + 2, p.width, p.height, 2))) return rc;
    a.out_color = (uint16_t*)c->att[p.out_set].ptr;
    a.out_frames = (uint16_t*)c->att[p.out_set + 1].ptr;
    a.out_hit = (uint16_t*)c->att[p.out_set + 2].ptr;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    for (int j = 0; j < 4; ++j) {   // u_PrevProjection * u_PrevView, column by column (mat4 * vec4 association of vmath.cuh)
        const float* v = p.prev_view + 4 * j;
        const float* m = p.prev_projection;
        for (int r = 0; r < 4; ++r) a.prev_pv[4 * j + r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
    }
    for (int k = 0; k < 3; ++k) { a.cam_cur[k] = p.current_camera_pos[k]; a.cam_prev[k] = p.prev_camera_pos[k]; }
    a.width = p.width; a.height = p.height;
    a.temporal_spec = p.temporal_spec; a.firefly = p.firefly_rejection; a.aggressive = p.aggressive_firefly_rejection;
    a.smart_clip = p.smart_clip; a.rough_weight = p.roughness_weight; a.stabilize = p.stabilize_hit_distance;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    specular_temporal_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_reflection_denoise(vxrt_ctx* c, const vxrt_reflection_denoise_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_reflection_denoise";
    if (p.out_attachment != VXRT_ATT_REFL_DENOISED_A && p.out_attachment != VXRT_ATT_REFL_DENOISED_B)
        return vxrt_fail(VXRT_E_INVALID, "%s: out_attachment must be VXRT_ATT_REFL_DENOISED_A / _B", fn);
    if (p.in_attachment == p.out_attachment) return vxrt_fail(VXRT_E_INVALID, "%s: in_attachment == out_attachment", fn);
    if (!is_refl_set(p.temporal_set)) return vxrt_fail(VXRT_E_INVALID, "%s: temporal_set must be VXRT_ATT_REFL_TEMPORAL_A / _B", fn);
    if (p.in_attachment < 0 || p.in_attachment >= VXRT_ATT_COUNT || p.hit_distance_attachment < 0 || p.hit_distance_attachment >= VXRT_ATT_COUNT)
        return vxrt_fail(VXRT_E_INVALID, "%s: bad attachment id", fn);
    ReflDenoiseArgs a;
    int rc;
    if ((rc = image_in(c, fn, p.in_attachment, 8, &a.in_color))) return rc;
    if ((rc = image_in(c, fn, p.temporal_set + 1, 2, &a.frames))) return rc;
    if ((rc = image_in(c, fn, p.hit_distance_attachment, 2, &a.hit))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_T, 2, &a.g_t))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_NORMAL, 1, &a.g_n))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_GBUF_NORMAL, 6, &a.gb_normal))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_GBUF_PBR, 4, &a.pbr))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_attachment, p.width, p.height, 8))) return rc;
    a.out = (uint16_t*)c->att[p.out_attachment].ptr;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; a.view[i] = p.view[i]; }
    a.width = p.width; a.height = p.height;
    a.dir = p.dir; a.roughness_bias = p.roughness_bias; a.normal_map_aware = p.normal_map_aware; a.handle_lobe_deviation = p.handle_lobe_deviation;
    a.derive_from_diffuse_sh = p.derive_from_diffuse_sh; a.amplify = p.amplify_transversal_weight; a.temporal_weight = p.temporal_weight;
    a.radius_bias = p.radius_bias; a.normal_map_weight_strength = p.normal_map_weight_strength; a.denoiser_scale = p.denoiser_scale;
    a.resolution_scale = p.resolution_scale; a.rnw_bias_strength = p.roughness_normal_weight_bias_strength;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    reflection_denoise_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
