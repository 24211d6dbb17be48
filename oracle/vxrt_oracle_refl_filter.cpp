/*
 * vxrt_oracle_refl_filter.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * Reflection temporal filter: Core/Shaders/SpecularTemporalFilter.glsl (dispatch and bindings: Core/Pipeline.cpp:3316-3400;
 * FBO formats :1189-1192, all LINEAR + REPEAT except the NEAREST R8 normal planes of the primary G-buffer).
 * Pinned against the shader itself compiled through the GLSL shim (oracle/_ref, vxref_specular_temporal):
 * tests/test_oracle_refl_filter.py.
 */
#include "vxrt_oracle.h"
#include "vxo_math.h"
#include "vxo_texture.h"

#include <vector>

using namespace vxo;

namespace {

inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}
/* GetRayDirectionAt (:87-92) */
inline v3 ray_direction_at(const float* inv_view, const float* inv_proj, v2 ss) {
    v4 clip = V4(ss.x * 2.0f - 1.0f, ss.y * 2.0f - 1.0f, -1.0f, 1.0f);
    v4 e = mat4_mul(inv_proj, clip);
    v4 r = mat4_mul(inv_view, V4(e.x, e.y, -1.0f, 0.0f));
    return V3(r.x, r.y, r.z);
}
/* GetNormalFromID (:100-112) as an index: 0..5 the face normals, 6 = (1, 1, 1); vector equality is index equality */
inline int normal_index(float n) { int i = cvt_round(n * 10.0f); return i > 5 ? 6 : i; }
inline Tex2D view(const std::vector<float>& d, int w, int h, int ch, bool linear) { Tex2D t; t.data = d.data(); t.w = w; t.h = h; t.ch = ch; t.linear = linear; return t; }
std::vector<float> from_half(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = half_to_float(h[i]); return o; }
std::vector<float> from_u8(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = unorm8_to_float(h[i]); return o; }
inline v4 tex4(const Tex2D& t, v2 c) { return tex2d_sample(t, c.x, c.y); }
inline float tex1(const Tex2D& t, v2 c) { return tex2d_sample(t, c.x, c.y).x; }
inline v3 xyz(v4 v) { return V3(v.x, v.y, v.z); }
inline v4 add4(v4 a, float s) { return V4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline v4 min4(v4 a, v4 b) { return V4(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z), gmin(a.w, b.w)); }   /* GLSL min(a, b) */
inline v4 max4(v4 a, v4 b) { return V4(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z), gmax(a.w, b.w)); }
inline void set_xyz(v4& v, v3 a) { v.x = a.x; v.y = a.y; v.z = a.z; }
inline float dist_sq(v3 a, v3 b) { v3 c = a - b; return dot(c, c); }                                        /* GetDistSquared (:137-141) */
/* ClipToAABB (:126-135) */
inline v3 clip_to_aabb(v3 prev, v3 mn, v3 mx) {
    v3 pClip = 0.5f * (mx + mn), eClip = 0.5f * (mx - mn);
    v3 vClip = prev - pClip, vUnit = vClip / eClip;
    v3 aUnit = V3(fabsf(vUnit.x), fabsf(vUnit.y), fabsf(vUnit.z));
    float denom = gmax(aUnit.x, gmax(aUnit.y, aUnit.z));
    return denom > 1.0f ? pClip + vClip / denom : prev;
}

}  // namespace

extern "C" void vxo_specular_temporal(const vxrt_specular_temporal_params* p, const uint16_t* cur_color_h4, const uint16_t* cur_hitdist,
                                      const uint8_t* cur_mask, const uint16_t* prev_hitdist, int rw, int rh, const uint16_t* hist_color_h4,
                                      const uint16_t* hist_hitdist, const uint16_t* g_t, const uint8_t* g_normal, const uint16_t* prev_t,
                                      const uint8_t* prev_normal, int gw, int gh, const uint8_t* pbr_u8x4, int mw, int mh,
                                      uint16_t* out_color_h4, uint16_t* out_frames, uint16_t* out_hitdist) {
    const int W = p->width, H = p->height;
    auto fc = from_half(cur_color_h4, (size_t)rw * rh * 4), fh = from_half(cur_hitdist, (size_t)rw * rh), fm = from_u8(cur_mask, (size_t)rw * rh);
    auto fph = from_half(prev_hitdist, (size_t)rw * rh);
    auto hc = from_half(hist_color_h4, (size_t)W * H * 4), hh = from_half(hist_hitdist, (size_t)W * H);
    auto ft = from_half(g_t, (size_t)gw * gh), fn = from_u8(g_normal, (size_t)gw * gh), pt = from_half(prev_t, (size_t)gw * gh), pn = from_u8(prev_normal, (size_t)gw * gh);
    auto fp = from_u8(pbr_u8x4, (size_t)mw * mh * 4);
    const Tex2D tCur = view(fc, rw, rh, 4, true), tHit = view(fh, rw, rh, 1, true), tMask = view(fm, rw, rh, 1, true), tPrevHit = view(fph, rw, rh, 1, true);
    const Tex2D tHist = view(hc, W, H, 4, true), tHistHit = view(hh, W, H, 1, true);
    const Tex2D tT = view(ft, gw, gh, 1, true), tN = view(fn, gw, gh, 1, false), tPT = view(pt, gw, gh, 1, true), tPN = view(pn, gw, gh, 1, false);
    const Tex2D tPBR = view(fp, mw, mh, 4, true);
    const v3 origin = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    float PV[16];   /* u_PrevProjection * u_PrevView */
    for (int j = 0; j < 4; ++j) {
        v4 c = mat4_mul(p->prev_projection, V4(p->prev_view[4 * j], p->prev_view[4 * j + 1], p->prev_view[4 * j + 2], p->prev_view[4 * j + 3]));
        PV[4 * j] = c.x; PV[4 * j + 1] = c.y; PV[4 * j + 2] = c.z; PV[4 * j + 3] = c.w;
    }
    const v2 TexelSize = V2(1.0f / (float)rw, 1.0f / (float)rh);   /* 1 / textureSize(u_CurrentColorTexture, 0) */
    const v3 CamCur = V3(p->current_camera_pos[0], p->current_camera_pos[1], p->current_camera_pos[2]);
    const v3 CamPrev = V3(p->prev_camera_pos[0], p->prev_camera_pos[1], p->prev_camera_pos[2]);
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    auto project_prev = [&](v3 pos) {
        v4 P = mat4_mul(PV, V4(pos.x, pos.y, pos.z, 1.0f));
        return V2((P.x / P.w) * 0.5f + 0.5f, (P.y / P.w) * 0.5f + 0.5f);
    };
    auto position_at = [&](const Tex2D& t, v2 c, float* dist) {   /* GetPositionAt (:94-98) */
        *dist = tex1(t, c);
        return origin + normalize(ray_direction_at(p->inv_view, p->inv_projection, c)) * *dist;
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            float CurDist;
            const v3 CurPos = position_at(tT, tc, &CurDist);
            const int InitialNormal = normal_index(tex1(tN, tc));
            v4 CurrentColor = tex4(tCur, tc);
            float oFrames = 0.0f, oHit;
            v4 oColor;
            const float HitDistanceCurrent = tex1(tHit, tc);
            if (p->firefly_rejection && tex1(tMask, tc) > 0.05f) {   /* FireflyReject (:246-277) */
                const int SampleThreshold = p->aggressive_firefly_rejection ? 3 : 4;
                const v2 Offsets[4] = {{1.0f, 0.0f}, {0.0f, 1.0f}, {-1.0f, 0.0f}, {0.0f, -1.0f}};
                v4 NonLit = V4(0.0f, 0.0f, 0.0f, 0.0f);
                int Unlit = 0;
                for (int i = 0; i < 4; ++i) {
                    const v2 sc = tc + Offsets[i] * TexelSize;
                    const float Mask = tex1(tMask, sc);
                    const v4 Color = tex4(tCur, sc);
                    if (Mask < 0.01f) { NonLit = V4(NonLit.x + Color.x, NonLit.y + Color.y, NonLit.z + Color.z, NonLit.w + Color.w); Unlit++; }
                }
                if (Unlit >= SampleThreshold) {
                    const float n = (float)Unlit;
                    CurrentColor = V4(NonLit.x / n, NonLit.y / n, NonLit.z / n, NonLit.w / n);
                }
            }
            if (CurDist > 0.0f && p->temporal_spec) {
                const bool SkySample = HitDistanceCurrent < 0.0f;
                const v4 pbr = tex4(tPBR, tc);
                const float RoughnessAt = gmix(0.095f, pbr.x, p->roughness_weight ? 1.0f : 0.0f);
                const float MetalnessAt = pbr.y;
                bool LessValid = false;
                v2 R = V2(0.0f, 0.0f);
                if (HitDistanceCurrent > 0.0f && !SkySample && RoughnessAt <= 0.875f + 0.01f) {   /* reproject along the reflected ray */
                    const v3 I = normalize(origin - CurPos);
                    R = project_prev(CurPos - I * HitDistanceCurrent);
                    const float PreviousT = tex1(tPrevHit, R);
                    if (fabsf(PreviousT - HitDistanceCurrent) >= 3.8f) LessValid = true;
                } else if (!SkySample) {
                    v3 CameraOffset = CamCur - CamPrev;
                    CameraOffset = CameraOffset * 0.6f;
                    R = project_prev(CurPos - CameraOffset);
                }
                if (SkySample) {
                    const v3 I = normalize(origin - CurPos);
                    R = project_prev(CurPos - I * 64.0f);
                }
                float PrevDist;
                const v3 PrevPos = position_at(tPT, R, &PrevDist);
                const float d = fabsf(distance(PrevPos, CurPos));
                const float Bias = 0.01f;
                const int PrevNormal = normal_index(tex1(tPN, R));
                if (R.x > 0.0f + Bias && R.x < 1.0f - Bias && R.y > 0.0f + Bias && R.y < 1.0f - Bias && d < 1.0f && PrevNormal == InitialNormal) {
                    v4 PrevColor = tex4(tHist, R);
                    const v3 BasePrevColor = xyz(PrevColor);
                    const bool Moved = dist_sq(CamCur, CamPrev) > 0.0001f;
                    const bool TryClipping = RoughnessAt < 0.5f + 0.01f;
                    if (TryClipping && Moved && p->smart_clip && !(RoughnessAt > 0.5f + 0.01f)) {   /* ReflectionClipping (:143-225), called with v_TexCoords */
                        const float RoughnessThreshold = 0.275f + 0.01f;
                        v4 MinColor = V4(1000.0f, 1000.0f, 1000.0f, 1000.0f), MaxColor = V4(-1000.0f, -1000.0f, -1000.0f, -1000.0f);
                        float AdditionalMaxBias = 0.0f;
                        for (int x = -1; x <= 1; ++x)
                            for (int y = -1; y <= 1; ++y) {
                                const v2 sc = V2(tc.x + (float)x * TexelSize.x, tc.y + (float)y * TexelSize.y);
                                if (!(tex1(tMask, sc) < 0.01f)) AdditionalMaxBias += 0.1f;
                                const v4 SampleColor = tex4(tCur, sc);
                                MinColor = min4(SampleColor, MinColor);
                                MaxColor = max4(SampleColor, MaxColor);
                            }
                        const v3 OriginalMin = xyz(MinColor), OriginalMax = xyz(MaxColor);
                        const bool Smoothish = RoughnessAt < RoughnessThreshold;
                        const bool Roughish = RoughnessAt > RoughnessThreshold && RoughnessAt < 0.50f + 0.01f;
                        if (Smoothish) {
                            const float Perceived = RoughnessAt * RoughnessAt;
                            const float RT2 = RoughnessThreshold * RoughnessThreshold;
                            const float Remapped = 0.0f + (gclamp((Perceived - 0.0f) / (RT2 - 0.0f), 0.0f, 1.0f) * (1.0f - 0.0f));   /* remap (:121-124) */
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
                        const v3 Prev3 = xyz(PrevColor);
                        const v3 Clamped = clip_to_aabb(Prev3, xyz(MinColor), xyz(MaxColor));
                        if (Clamped.x != Prev3.x || Clamped.y != Prev3.y || Clamped.z != Prev3.z)
                            PrevColor = dist_sq(Clamped, xyz(MinColor)) > dist_sq(Clamped, xyz(MaxColor)) ? MaxColor : MinColor;
                    }
                    /* GetAccumulationFactor (:279-283) */
                    const v2 Vel = V2((tc.x - R.x) * (float)rw, (tc.y - R.y) * (float)rh);
                    float AF = gclamp(expf(-sqrtf(dot(Vel, Vel))) * 0.9f + 0.750f, 0.00000001f, 0.96f);
                    AF = gclamp(AF, 0.001f, 0.95f);
                    CurrentColor = V4(gmax(CurrentColor.x, 0.0f), gmax(CurrentColor.y, 0.0f), gmax(CurrentColor.z, 0.0f), gmax(CurrentColor.w, 0.0f));
                    PrevColor = V4(gmax(PrevColor.x, 0.0f), gmax(PrevColor.y, 0.0f), gmax(PrevColor.z, 0.0f), gmax(PrevColor.w, 0.0f));
                    oColor = V4(gmix(CurrentColor.x, PrevColor.x, AF), gmix(CurrentColor.y, PrevColor.y, AF), gmix(CurrentColor.z, PrevColor.z, AF),
                                gmix(CurrentColor.w, PrevColor.w, AF));
                    oFrames = AF;
                    oHit = HitDistanceCurrent;
                    if (dist_sq(BasePrevColor, xyz(PrevColor)) < 0.2f && p->stabilize_hit_distance)
                        oHit = gmix(HitDistanceCurrent, tex1(tHistHit, R), gclamp(AF * 1.1f, 0.0f, 0.9f));
                } else {
                    oColor = CurrentColor; oFrames = 0.0f; oHit = HitDistanceCurrent;
                }
            } else {
                oColor = CurrentColor; oHit = HitDistanceCurrent;
            }
            if (!p->temporal_spec) oFrames = -1.0f;
            const size_t i = (size_t)py * W + px;
            out_color_h4[4 * i] = float_to_half(oColor.x); out_color_h4[4 * i + 1] = float_to_half(oColor.y);
            out_color_h4[4 * i + 2] = float_to_half(oColor.z); out_color_h4[4 * i + 3] = float_to_half(oColor.w);
            out_frames[i] = float_to_half(oFrames);
            out_hitdist[i] = float_to_half(oHit);
        }
}

/* ReflectionDenoiserNew.glsl main() (:97-364): one direction of the separable bilateral filter of the reflection colour.
 * `int Jitter = int((GradientNoise() - 0.5f) * 0.75f)` (:204) truncates a value in (-0.375, 0.375): it is 0 for every pixel and
 * every u_Time, and `SampleCoord += Jitter * TexelSize * 0.75f` adds +0.0 — neither the noise nor u_Time reach the output. */
extern "C" void vxo_reflection_denoise(const vxrt_reflection_denoise_params* p, const uint16_t* in_color_h4, int iw, int ih,
                                       const uint16_t* frames, const uint16_t* hitdist, int tw, int th, int hw, int hh,
                                       const uint16_t* g_t, const uint8_t* g_normal, const uint8_t* g_block, int gw, int gh,
                                       const uint16_t* gb_normal_h3, const uint8_t* pbr_u8x4, int mw, int mh, uint16_t* out_color_h4) {
    const int W = p->width, H = p->height;
    auto fc = from_half(in_color_h4, (size_t)iw * ih * 4), ffr = from_half(frames, (size_t)tw * th), fhd = from_half(hitdist, (size_t)hw * hh);
    auto ft = from_half(g_t, (size_t)gw * gh), fn = from_u8(g_normal, (size_t)gw * gh), fb = from_u8(g_block, (size_t)gw * gh);
    auto fgn = from_half(gb_normal_h3, (size_t)mw * mh * 3), fp = from_u8(pbr_u8x4, (size_t)mw * mh * 4);
    const Tex2D tIn = view(fc, iw, ih, 4, true), tFrames = view(ffr, tw, th, 1, true), tHit = view(fhd, hw, hh, 1, true);
    const Tex2D tT = view(ft, gw, gh, 1, true), tN = view(fn, gw, gh, 1, false), tGN = view(fgn, mw, mh, 3, true), tPBR = view(fp, mw, mh, 4, true);
    const v3 origin = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    static const float Gauss[33] = {0.004013f, 0.005554f, 0.007527f, 0.00999f, 0.012984f, 0.016524f, 0.020594f, 0.025133f, 0.030036f, 0.035151f, 0.040283f,
                                    0.045207f, 0.049681f, 0.053463f, 0.056341f, 0.058141f, 0.058754f, 0.058141f, 0.056341f, 0.053463f, 0.049681f, 0.045207f,
                                    0.040283f, 0.035151f, 0.030036f, 0.025133f, 0.020594f, 0.016524f, 0.012984f, 0.00999f, 0.007527f, 0.005554f, 0.004013f};
    static const v3 Normals[7] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {1, 1, 1}};
    const v3 LumaW = V3(0.299f, 0.587f, 0.114f);
    auto block_id = [&](v2 c) {   /* GetBlockID (:79-83): texelFetch at ivec2(txc * size) */
        const int x = (int)(c.x * (float)gw), y = (int)(c.y * (float)gh);
        const float id = fb[(size_t)y * gw + x];
        return iclamp((int)floorf(id * 255.0f), 0, 127);
    };
    const bool Dir = p->dir != 0;
    const float EPS = 0.001f;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const float BaseDist = tex1(tT, tc);
            const v3 BasePos = origin + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * BaseDist;
            const v3 BaseNormal = Normals[normal_index(tex1(tN, tc))];
            const int BaseBlockID = block_id(tc);
            (void)BaseBlockID;   /* BlockValidity (:251) is computed and never used */
            const bool BaseIsSky = BaseDist < 0.0f;
            const v4 BaseColor = tex4(tIn, tc);
            const float BaseLuminance = dot(xyz(BaseColor), LumaW);
            float TotalWeight = 0.0f;
            const float TexelSize = Dir ? 1.0f / (float)W : 1.0f / (float)H;   /* u_Dimensions = size of the output (Pipeline.cpp:3426) */
            const v4 SampledPBR = tex4(tPBR, tc);
            float BaseRoughness = SampledPBR.x;
            const float RawRoughness = BaseRoughness;
            BaseRoughness *= gmix(1.0f, 0.91f, p->roughness_bias ? 1.0f : 0.0f);
            const v3 NormalMappedBase = xyz(tex4(tGN, tc));
            float HitDistanceFetch = tex1(tHit, tc) + 0.0001f;
            const float LobeDistanceCurveBias = 1.0f / 1.3f;
            if (HitDistanceFetch < 0.001f) HitDistanceFetch = 1.75f;
            else HitDistanceFetch = powf(HitDistanceFetch, LobeDistanceCurveBias);
            if (p->handle_lobe_deviation) {
                if (BaseRoughness <= 0.25f + 0.05f) HitDistanceFetch = gclamp(HitDistanceFetch, 0.0f, 8.0f);
                if (BaseRoughness <= 0.2f) HitDistanceFetch = gclamp(HitDistanceFetch, 0.0f, 5.5f);
            }
            const float SpecularHitDistance = gmax(HitDistanceFetch, 0.01f) * (RawRoughness < 0.51f ? 0.5f : 0.85f);
            const v4 vs = mat4_mul(p->view, V4(BasePos.x, BasePos.y, BasePos.z, 1.0f));
            const float ViewLength = length(V3(vs.x, vs.y, vs.z));
            float ViewLengthWeight = 0.001f + ViewLength;
            if (BaseRoughness > 0.135f) ViewLengthWeight = gmax(ViewLengthWeight, 0.750f);
            else ViewLengthWeight = gmax(ViewLengthWeight, 3.0f);
            if (BaseRoughness < 0.125f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 6.0f);
            else if (BaseRoughness < 0.25f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 8.0f + 1.0f);
            else if (BaseRoughness < 0.5f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 16.0f);
            else if (BaseRoughness < 0.75f) ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 24.0f);
            else ViewLengthWeight = gclamp(ViewLengthWeight, 0.000001f, 32.0f);
            float TransversalContrib = SpecularHitDistance / gmax((SpecularHitDistance + ViewLengthWeight), 0.00001f);
            if (RawRoughness < 0.535f && p->amplify_transversal_weight && BaseDist < 50.0f) {
                const float Remapped = (((RawRoughness - 0.0f) / (0.535f - 0.0f)) * (1.0f - 0.0f)) + 0.0f;   /* remap (:93-96) */
                const float TransversalExponent = gmix(3.5f, 2.0f, powf(Remapped, 4.0f));
                TransversalContrib = powf(TransversalContrib, TransversalExponent + 0.8125f);
            }
            const float RadiusExponent = powf((1.0f - BaseRoughness), 1.0f / 1.4f) * 5.0f;
            const float Radius = gclamp(powf(gmix(1.0f * BaseRoughness, 1.0f, TransversalContrib), RadiusExponent), 0.0f, 1.0f);
            const float NormalMapRadius = 1.0f - gclamp(powf(gmix(1.0f * BaseRoughness, 1.0f, TransversalContrib), RadiusExponent), 0.0f, 1.0f);
            int EffectiveRadius = (int)floorf(Radius * 15.0f);
            EffectiveRadius = iclamp(EffectiveRadius, 1, 15);
            const bool BaseTooRough = BaseRoughness > 0.897511f;
            EffectiveRadius = BaseTooRough ? 15 : EffectiveRadius;
            float Scale = gmix(1.0f, 2.0f, gclamp(p->resolution_scale, 0.0000001f, 1.0f)) + 0.5f;
            int RadiusBias = 0;
            if (SampledPBR.y > 0.1f - EPS && BaseRoughness > 0.4f - EPS) RadiusBias += 2;
            EffectiveRadius = iclamp(EffectiveRadius + RadiusBias + p->radius_bias, 1, 15);
            if (RawRoughness >= 0.5f - 0.01f) EffectiveRadius += 1;
            if (p->derive_from_diffuse_sh && RawRoughness >= 0.865f) { EffectiveRadius = 4; Scale *= 1.25f; }
            Scale *= p->denoiser_scale;
            float TemporalWeight = 0.0f, AccumulatedFramesClamped = 0.01f;
            if (p->temporal_weight) {
                const float AccumulatedFrames = tex1(tFrames, tc);
                AccumulatedFramesClamped = AccumulatedFrames < -0.1f ? 0.0f : (1.0f - AccumulatedFrames);
                AccumulatedFramesClamped = gclamp(AccumulatedFramesClamped, 0.000001f, 1.0f);
                TemporalWeight = gclamp(AccumulatedFramesClamped * 0.85f, 0.0f, 1.0f);
                float FLT_radius = (float)EffectiveRadius;
                FLT_radius = gmix(FLT_radius, FLT_radius + 2.0f, AccumulatedFramesClamped * 1.05f);
                EffectiveRadius = cvt_trunc(FLT_radius);
            }
            float HF_e = 64.0f * p->normal_map_weight_strength * 1.350f;
            HF_e *= powf(NormalMapRadius, 1.0f / 1.33f);
            float HF_WeightAdder = gmix(0.0f, 0.005f, BaseRoughness > 0.45f ? 1.0f : 0.0f);
            HF_WeightAdder += gmix(0.0f, 0.0125f, BaseRoughness > 0.525f ? 1.0f : 0.0f);
            HF_WeightAdder += gmix(0.0f, 0.022f, BaseRoughness > 0.625f ? 1.0f : 0.0f);
            HF_WeightAdder += gmix(0.0f, 0.026f, BaseRoughness > 0.725f ? 1.0f : 0.0f);
            HF_WeightAdder += gmix(0.0f, 0.031f, BaseRoughness > 0.75f ? 1.0f : 0.0f);
            EffectiveRadius = iclamp(EffectiveRadius, 1, 15);
            if (RawRoughness < 0.002f) EffectiveRadius = 0;
            v4 Filtered = V4(0.0f, 0.0f, 0.0f, 0.0f);
            for (int Sample = -EffectiveRadius; Sample <= EffectiveRadius; Sample++) {
                const float SampleOffset = (float)Sample;
                const v2 sc = Dir ? V2(tc.x + (SampleOffset * Scale * TexelSize), tc.y) : V2(tc.x, tc.y + (SampleOffset * Scale * TexelSize));
                const float bias = 0.01f;
                if (!(sc.x > 0.0f + bias && sc.x < 1.0f - bias && sc.y > 0.0f + bias && sc.y < 1.0f - bias)) continue;
                const float SampleDepth = tex1(tT, sc);
                const bool SampleIsSky = SampleDepth < 0.0f;
                if (SampleIsSky != BaseIsSky) continue;
                const v4 SampleData = tex4(tIn, sc);
                const float DepthDifference = fabsf(SampleDepth - BaseDist) * 1.5f;
                const float DepthWeight = powf(expf(-DepthDifference), 2.0f);
                const v3 SampleNormal = Normals[normal_index(tex1(tN, sc))];
                const float NormalWeight = powf(gmax(dot(BaseNormal, SampleNormal), 0.00000000001f), 32.0f);
                float LuminanceWeight = 1.0f;
                const float SampleRoughness = tex4(tPBR, sc).x;
                const bool SampleTooRough = SampleRoughness >= 0.89f;
                if (!SampleTooRough) {
                    const float LumaAt = dot(xyz(SampleData), LumaW);
                    float LuminanceError = 1.0f / fabsf(LumaAt - BaseLuminance);
                    LuminanceError = powf(LuminanceError, 1.7f);
                    const float LumaWeightExponent = gmix(0.001f, 8.0f, powf(SampleRoughness, 16.0f));
                    LuminanceWeight = powf(LuminanceError, LumaWeightExponent + 0.8f);
                    LuminanceWeight = gclamp(LuminanceWeight, 0.0000000001f, 1.0f);
                    LuminanceWeight = gmix(LuminanceWeight, 1.0f, TemporalWeight);
                    LuminanceWeight = gclamp(LuminanceWeight, 0.0000000001f, 1.0f);
                }
                float HFNormalWeight = 1.0f;
                if (p->normal_map_aware && !SampleTooRough && BaseRoughness < 0.8f && AccumulatedFramesClamped <= 0.185f + 0.001f + 0.001f + 0.0001f) {
                    const v3 NormalMapAt = xyz(tex4(tGN, sc));
                    const float Angle = dot(NormalMapAt, NormalMappedBase);
                    HFNormalWeight = powf(gclamp(Angle, 0.00000001f, 1.0f), HF_e);
                    HFNormalWeight = gclamp(HFNormalWeight + (HF_WeightAdder * p->roughness_normal_weight_bias_strength * 1.4f), 0.00000000001f, 1.0f);
                }
                const float RoughnessError = fabsf(SampleRoughness - BaseRoughness);
                float RoughnessTransversalWeight = 1.0f / RoughnessError;
                RoughnessTransversalWeight = powf(RoughnessTransversalWeight, 12.0f);
                RoughnessTransversalWeight = gclamp(RoughnessTransversalWeight, 0.00000000001f, 1.0f);
                const float CurrentKernelWeight = Gauss[iclamp(16 + Sample, 0, 32)];
                float CurrentWeight = 1.0f;
                CurrentWeight *= DepthWeight;
                CurrentWeight *= NormalWeight;
                CurrentWeight *= HFNormalWeight;
                CurrentWeight *= LuminanceWeight;
                CurrentWeight *= RoughnessTransversalWeight;
                CurrentWeight *= CurrentKernelWeight;
                CurrentWeight = gclamp(CurrentWeight, 0.000000001f, 1.0f);
                Filtered = V4(Filtered.x + SampleData.x * CurrentWeight, Filtered.y + SampleData.y * CurrentWeight, Filtered.z + SampleData.z * CurrentWeight,
                              Filtered.w + SampleData.w * CurrentWeight);
                TotalWeight += CurrentWeight;
            }
            const bool DoSpatial = !(RawRoughness < 0.002f);
            v4 o;
            if (TotalWeight > 0.001f && DoSpatial) {
                Filtered = V4(Filtered.x / TotalWeight, Filtered.y / TotalWeight, Filtered.z / TotalWeight, Filtered.w / TotalWeight);
                float Smooth = 1.0f;
                if (BaseRoughness <= 0.1f + 0.007f) {
                    Smooth = BaseRoughness * 16.0f;
                    Smooth = 1.0f - Smooth;
                    Smooth = powf(Smooth, 4.0f);
                    Smooth = gclamp(Smooth, 0.1f, 0.999f);
                }
                o = V4(gmix(BaseColor.x, Filtered.x, Smooth), gmix(BaseColor.y, Filtered.y, Smooth), gmix(BaseColor.z, Filtered.z, Smooth), gmix(BaseColor.w, Filtered.w, Smooth));
            } else {
                o = BaseColor;
            }
            const size_t i = (size_t)py * W + px;
            out_color_h4[4 * i] = float_to_half(o.x); out_color_h4[4 * i + 1] = float_to_half(o.y);
            out_color_h4[4 * i + 2] = float_to_half(o.z); out_color_h4[4 * i + 3] = float_to_half(o.w);
        }
}
