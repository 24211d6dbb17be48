/*
 * vxrt_oracle_shadow_filter.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * Sun-shadow denoiser: Core/Shaders/ShadowTemporalFilter.glsl and ShadowFilter.glsl (dispatch and bindings:
 * Core/Pipeline.cpp:2947-3044; FBO formats :1200-1202, all LINEAR + REPEAT except the NEAREST R8 normal plane).
 * The shadow images are single-channel: every vec3 / vec4 the shaders build from them is consumed through .x only.
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
inline v3 ray_direction_at(const float* inv_view, const float* inv_proj, v2 ss) {
    v4 clip = V4(ss.x * 2.0f - 1.0f, ss.y * 2.0f - 1.0f, -1.0f, 1.0f);
    v4 e = mat4_mul(inv_proj, clip);
    v4 r = mat4_mul(inv_view, V4(e.x, e.y, -1.0f, 0.0f));
    return V3(r.x, r.y, r.z);
}
/* GetNormalFromID as an index (0..5 face normals, 6 = (1, 1, 1)) and the dot product of two of them */
inline int normal_index(float n) { int i = cvt_round(n * 10.0f); return i > 5 ? 6 : i; }
inline v3 normal_of(int i) {
    static const v3 N[7] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {1, 1, 1}};
    return N[i];
}
inline Tex2D view(const std::vector<float>& d, int w, int h, bool linear) { Tex2D t; t.data = d.data(); t.w = w; t.h = h; t.ch = 1; t.linear = linear; return t; }
std::vector<float> from_half(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = half_to_float(h[i]); return o; }
std::vector<float> from_u8(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = unorm8_to_float(h[i]); return o; }
inline float gclampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float tex1(const Tex2D& t, v2 c) { return tex2d_sample(t, c.x, c.y).x; }

}  // namespace

/* ShadowTemporalFilter.glsl main() (:193-267) with GetShadowSpatial (:107-151), ClipShadow / clipAABB (:153-191). */
extern "C" void vxo_shadow_temporal(const vxrt_shadow_temporal_params* p, const uint8_t* raw_shadow, const uint16_t* raw_transversal, int sw, int sh,
                                    const uint8_t* hist_shadow, const uint16_t* hist_frames, const uint16_t* g_t, const uint8_t* g_normal,
                                    const uint16_t* prev_t, int gw, int gh, uint8_t* out_shadow, uint16_t* out_frames) {
    const int W = p->width, H = p->height;
    auto fs = from_u8(raw_shadow, (size_t)sw * sh), ftr = from_half(raw_transversal, (size_t)sw * sh);
    auto hs = from_u8(hist_shadow, (size_t)W * H), hf = from_half(hist_frames, (size_t)W * H);
    auto ft = from_half(g_t, (size_t)gw * gh), fn = from_u8(g_normal, (size_t)gw * gh), pt = from_half(prev_t, (size_t)gw * gh);
    const Tex2D tCur = view(fs, sw, sh, true), tTr = view(ftr, sw, sh, true), tPrev = view(hs, W, H, true), tFrames = view(hf, W, H, true);
    const Tex2D tT = view(ft, gw, gh, true), tN = view(fn, gw, gh, false), tPT = view(pt, gw, gh, true);
    const v3 origin = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    float PV[16];   /* u_PrevProjection * u_PrevView */
    for (int j = 0; j < 4; ++j) {
        v4 c = mat4_mul(p->prev_projection, V4(p->prev_view[4 * j], p->prev_view[4 * j + 1], p->prev_view[4 * j + 2], p->prev_view[4 * j + 3]));
        PV[4 * j] = c.x; PV[4 * j + 1] = c.y; PV[4 * j + 2] = c.z; PV[4 * j + 3] = c.w;
    }
    const v2 TexelSize = V2(1.0f / (float)sw, 1.0f / (float)sh);   /* 1 / textureSize(u_CurrentColorTexture, 0) */
    const float UnitDiagonal = sqrtf(2.0f);
    const bool ST = p->shadow_temporal != 0;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const float Dist = tex1(tT, tc);
            const v3 CurPos = origin + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            float oColor, oFrames = 0.0f;
            if (Dist > 0.0f) {
                v4 Proj = mat4_mul(PV, V4(CurPos.x, CurPos.y, CurPos.z, 1.0f));
                const v2 R = V2((Proj.x / Proj.w) * 0.5f + 0.5f, (Proj.y / Proj.w) * 0.5f + 0.5f);
                const float TransversalAt = tex1(tTr, tc) * 100.0f;
                float CurrentColor;
                if (!ST) CurrentColor = tex1(tCur, tc);
                else if (TransversalAt <= UnitDiagonal * 2.0f) CurrentColor = 1.0f;
                else {   /* GetShadowSpatial */
                    float Total = tex1(tCur, tc);
                    const float Base = Total;
                    float Weight = 1.0f;
                    const int BaseNormal = normal_index(tex1(tN, tc));
                    for (int x = -1; x <= 1; ++x)
                        for (int y = -1; y <= 1; ++y) {
                            if (x == 0 && y == 0) continue;
                            const v2 sc = V2(tc.x + (float)x * TexelSize.x, tc.y + (float)y * TexelSize.y);
                            const float b = 0.03f;
                            if (!(sc.x > b && sc.x < 1.0f - b && sc.y > b && sc.y < 1.0f - b)) continue;
                            const float SampleDepth = tex1(tT, sc);
                            const int SampleNormal = normal_index(tex1(tN, sc));
                            if (SampleNormal == BaseNormal && fabsf(SampleDepth - Dist) < 1.0f) {
                                const float Sample = tex1(tCur, sc);
                                float WeightAt = gclampf(1.0f - gclampf(fabsf(Sample - Base) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                                WeightAt = gclampf(powf(WeightAt, 7.0f), 0.000001f, 1.0f);
                                Total += Sample * WeightAt;
                                Weight += WeightAt;
                            }
                        }
                    CurrentColor = Total / Weight;
                }
                const float PrevColorOrig = tex1(tPrev, R);
                float PrevColor = PrevColorOrig;
                if (ST && TransversalAt < 1.414f * 3.0f) {   /* ClipShadow */
                    float MinColor = 100.0f, MaxColor = -100.0f;
                    const v2 Off[5] = {{-1, 0}, {1, 0}, {0, 0}, {0, -1}, {0, 1}};
                    for (int s = 0; s < 5; ++s) {
                        const float Sample = tex1(tCur, V2(tc.x + Off[s].x * TexelSize.x, tc.y + Off[s].y * TexelSize.y));
                        MinColor = gmin(Sample, MinColor);
                        MaxColor = gmax(Sample, MaxColor);
                    }
                    const float History = tex1(tPrev, R);
                    const float mn = MinColor - 0.125f, mx = MaxColor + 0.125f;   /* clipAABB on equal components */
                    const float pClip = 0.5f * (mx + mn), eClip = 0.5f * (mx - mn), vClip = History - pClip;
                    const float denom = fabsf(vClip / eClip);
                    PrevColor = denom > 1.0f ? pClip + vClip / denom : History;
                }
                const float PrevDist = tex1(tPT, R);
                const v3 PrevPos = origin + normalize(ray_direction_at(p->inv_view, p->inv_projection, R)) * PrevDist;
                const float Bias = ST ? 0.005f : 0.01f;
                const bool Rejected = !(R.x > 0.0f + Bias && R.x < 1.0f - Bias && R.y > 0.0f + Bias && R.y < 1.0f - Bias);
                if (!Rejected) {
                    const float d = distance(PrevPos, CurPos);
                    CurrentColor = gclampf(CurrentColor, 0.0f, 1.0f);
                    PrevColor = gclampf(PrevColor, 0.0f, 1.0f);
                    const v2 Vel = V2((tc.x - R.x) * (float)sw, (tc.y - R.y) * (float)sh);
                    const float ClipError = fabsf(PrevColorOrig - PrevColor);
                    const float FrameIncrement = ClipError < 0.2f ? 1.0f : 0.6f;
                    const float FrameCountFetch = tex1(tFrames, R);
                    const float FrameIncremented = FrameCountFetch + FrameIncrement;
                    float BlendFactor = gclampf((1.0f - (1.0f / FrameIncremented)) * 1.2f, 0.01f, 0.97f);
                    const float VRF = gclampf(expf(-sqrtf(dot(Vel, Vel))) * 0.8f + 0.6f, 0.00000001f, 1.0f);
                    BlendFactor *= VRF;
                    float DepthRejection = 1.0f;
                    if (d > 0.4f) {
                        DepthRejection = powf(expf(-d), 48.0f);
                        BlendFactor *= gclampf(DepthRejection, 0.0f, 1.0f);
                    }
                    oColor = gmix(CurrentColor, PrevColor, gclampf(BlendFactor, 0.0f, 0.97f));
                    const float BFM = DepthRejection * VRF;
                    oFrames = FrameCountFetch + gclampf(BFM * 1.1f, 0.0f, 1.0f);
                    if (BFM < 0.1f) oFrames = 0.0f;
                    else if (BFM <= 0.2f + 0.001f) oFrames = 2.0f;
                    else if (BFM <= 0.3f + 0.001f) oFrames = 3.25f;
                } else {
                    oColor = CurrentColor;
                    oFrames = 0.0f;
                }
            } else {
                oColor = tex1(tCur, tc);
                oFrames = 0.0f;
            }
            oFrames = gclampf(oFrames, 0.0f, 256.0f);
            const size_t i = (size_t)py * W + px;
            out_shadow[i] = float_to_unorm8(oColor);
            out_frames[i] = float_to_half(oFrames);
        }
}

/* ShadowFilter.glsl ShadowSpatial (:68-159): taps are not tested against the screen, REPEAT wraps them. */
extern "C" void vxo_shadow_filter(const vxrt_shadow_filter_params* p, const uint8_t* in_shadow, const uint16_t* in_frames, int iw, int ih,
                                  const uint16_t* raw_transversal, int sw, int sh, const uint16_t* g_t, const uint8_t* g_normal, int gw, int gh,
                                  uint8_t* out_shadow) {
    const int W = p->width, H = p->height;
    auto fs = from_u8(in_shadow, (size_t)iw * ih), ff = from_half(in_frames, (size_t)iw * ih), ftr = from_half(raw_transversal, (size_t)sw * sh);
    auto ft = from_half(g_t, (size_t)gw * gh), fn = from_u8(g_normal, (size_t)gw * gh);
    const Tex2D tIn = view(fs, iw, ih, true), tFrames = view(ff, iw, ih, true), tTr = view(ftr, sw, sh, true);
    const Tex2D tT = view(ft, gw, gh, true), tN = view(fn, gw, gh, false);
    const v2 TexelSize = V2(1.0f / (float)iw, 1.0f / (float)ih);
    const float Cutoff = sqrtf(2.0f);
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const float Frames = tex1(tFrames, tc);
            const bool ApplyLumaWeight = Frames > 7.5f;
            const float CenterDist = tex1(tT, tc);
            const bool Sky = CenterDist < 0.0f;
            const v3 CenterNormal = normal_of(normal_index(tex1(tN, tc)));
            const float CenterShadow = tex1(tIn, tc);
            const float Transversal = tex1(tTr, tc) * 100.0f;
            const bool Sharp = Transversal > 0.0f && Transversal < Cutoff;
            float result = CenterShadow;
            if (!(Sharp || Sky)) {
                const bool Reduced = Transversal < Cutoff * 1.414f;
                const int K = Reduced ? 1 : 3;
                float Scale = 1.0f;
                if (Transversal > 6.0f) Scale = 2.0f;
                if (Transversal > 16.0f) Scale = 2.4f;
                if (Transversal > 32.0f) Scale = 2.6f;
                const float ClampedT = gclampf(Transversal, 0.0f, 10.0f);
                float VarianceEstimate = gmix(20.0f, 6.0f, ClampedT / 10.0f) + (Transversal < 6.0f ? 5.0f : 2.0f);
                VarianceEstimate = gclampf(VarianceEstimate - 1.75f, 0.0000001f, 64.0f);
                float LumaMixer = 1.0f;
                if (!ApplyLumaWeight) LumaMixer = gmix(0.1f, 0.5f, Frames / 7.5f);
                float TotalWeight = 0.0f, TotalShadow = 0.0f;
                for (int x = -K; x <= K; ++x)
                    for (int y = -K; y <= K; ++y) {
                        const v2 sc = V2(tc.x + ((((float)x * TexelSize.x) * 1.2f) * Scale) * p->filter_scale,
                                         tc.y + ((((float)y * TexelSize.y) * 1.2f) * Scale) * p->filter_scale);
                        const float SampleDepth = tex1(tT, sc);
                        const v3 SampleNormal = normal_of(normal_index(tex1(tN, sc)));
                        const float DepthWeight = powf(expf(-(fabsf(CenterDist - SampleDepth))), 3.0f);
                        const float NormalWeight = powf(gmax(dot(CenterNormal, SampleNormal), 0.000000001f), 32.0f);
                        const float ShadowAt = tex1(tIn, sc);
                        const float LuminanceError = gclampf(1.0f - gclampf(fabsf(ShadowAt - CenterShadow) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                        float Weight = 1.0f;
                        Weight *= gclampf(powf(LuminanceError, VarianceEstimate * LumaMixer * 0.9f), 0.0f, 1.0f);
                        Weight *= DepthWeight;
                        Weight *= NormalWeight;
                        Weight = gclampf(Weight, 0.000000001f, 1.0f);
                        TotalShadow += ShadowAt * Weight;
                        TotalWeight += Weight;
                    }
                result = TotalShadow / gmax(TotalWeight, 0.01f);
            }
            out_shadow[(size_t)py * W + px] = float_to_unorm8(result);
        }
}
