/*
 * vxrt_oracle_svgf.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * SVGF denoiser chain of the diffuse GI: Core/Shaders/SVGF/TemporalFilter.glsl, VarianceEstimate.glsl,
 * SpatialFilter.glsl (dispatch and bindings: Core/Pipeline.cpp:2428-2700).  Citations are file:line of the reference.
 * FBO attachments are LINEAR + REPEAT for the float formats and RG8, NEAREST for the R8 G-buffer planes
 * (Core/Pipeline.cpp:1142-1156, Core/GLClasses/Framebuffer.h:16-18).
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
/* GetNormalFromID (TemporalFilter.glsl:111-123): idx > 5 -> (1,1,1) */
inline v3 normal_from_id(float n) {
    static const v3 N[6] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}};
    int i = cvt_round(n * 10.0f);
    if (i > 5) return V3(1.0f, 1.0f, 1.0f);
    return N[i];
}
inline Tex2D view(const std::vector<float>& d, int w, int h, int ch, bool linear) { Tex2D t; t.data = d.data(); t.w = w; t.h = h; t.ch = ch; t.linear = linear; return t; }
std::vector<float> from_half(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = half_to_float(h[i]); return o; }
std::vector<float> from_u8(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = unorm8_to_float(h[i]); return o; }
inline float gclampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline bool in_screen_space(v2 v) { return v.x < 1.0f && v.x > 0.0f && v.y < 1.0f && v.y > 0.0f; }
/* SHToY (VarianceEstimate.glsl:43-46) */
inline float sh_to_y(v4 sh) { return gmax(0.0f, 3.544905f * sh.w); }

struct GBuf {
    Tex2D t, n, b;
    const float* inv_view; const float* inv_proj;
    v3 origin;
    /* GetPositionAt (TemporalFilter.glsl:80-84): current inverse matrices and ray origin for BOTH frames' depth */
    v4 position_at(const Tex2D& pos, v2 txc) const {
        float Dist = tex2d_sample(pos, txc.x, txc.y).x;
        v3 p = origin + normalize(ray_direction_at(inv_view, inv_proj, txc)) * Dist;
        return V4(p.x, p.y, p.z, Dist);
    }
};

}  // namespace

/* one image set as the attachments hold it */
struct vxo_svgf_set {
    const uint16_t* sh;    /* RGBA16F */
    const uint16_t* cocg;  /* RG16F */
    const uint16_t* x;     /* RGB16F utility or R16F variance / luminance */
    const uint8_t* aosky;  /* RG8 */
};
struct vxo_svgf_out {
    uint16_t* sh; uint16_t* cocg; uint16_t* x; uint8_t* aosky;
};

/* TemporalFilter.glsl main() (:133-344).  cur.x = u_NoisyLuminosity (R16F); hist.x = u_PreviousUtility (RGB16F).
 * Jitter = ivec2((GradientNoise() - 0.5) * 1.0) is (0, 0) for every pixel: the noise is in [0, 1) and int() truncates. */
extern "C" void vxo_svgf_temporal(const vxrt_svgf_temporal_params* p, const vxo_svgf_set* cur, const vxo_svgf_set* hist,
                                  const uint16_t* g_t, const uint8_t* g_normal, const uint8_t* g_block,
                                  const uint16_t* prev_t, const uint8_t* prev_normal, const uint8_t* prev_block, const vxo_svgf_out* out) {
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = from_half(cur->sh, 4 * n), fcc = from_half(cur->cocg, 2 * n), flum = from_half(cur->x, n), fao = from_u8(cur->aosky, 2 * n);
    auto hsh = from_half(hist->sh, 4 * n), hcc = from_half(hist->cocg, 2 * n), hut = from_half(hist->x, 3 * n), hao = from_u8(hist->aosky, 2 * n);
    auto ft = from_half(g_t, n), fn = from_u8(g_normal, n), fb = from_u8(g_block, n);
    auto pt = from_half(prev_t, n), pn = from_u8(prev_normal, n), pb = from_u8(prev_block, n);
    const Tex2D tSH = view(fsh, W, H, 4, true), tCC = view(fcc, W, H, 2, true), tLum = view(flum, W, H, 1, true), tAO = view(fao, W, H, 2, true);
    const Tex2D pSH = view(hsh, W, H, 4, true), pCC = view(hcc, W, H, 2, true), pUt = view(hut, W, H, 3, true), pAO = view(hao, W, H, 2, true);
    const Tex2D tT = view(ft, W, H, 1, true), tN = view(fn, W, H, 1, false), tB = view(fb, W, H, 1, false);
    const Tex2D qT = view(pt, W, H, 1, true), qN = view(pn, W, H, 1, false), qB = view(pb, W, H, 1, false);
    GBuf g; g.inv_view = p->inv_view; g.inv_proj = p->inv_projection; g.origin = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    /* u_PrevProjection * u_PrevView (left to right: the matrix product first), columns = M * column */
    float PV[16];
    for (int j = 0; j < 4; ++j) {
        v4 c = mat4_mul(p->prev_projection, V4(p->prev_view[4 * j], p->prev_view[4 * j + 1], p->prev_view[4 * j + 2], p->prev_view[4 * j + 3]));
        PV[4 * j] = c.x; PV[4 * j + 1] = c.y; PV[4 * j + 2] = c.z; PV[4 * j + 3] = c.w;
    }
    const float Weights[5] = {3.0f / 32.0f, 3.0f / 32.0f, 9.0f / 64.0f, 3.0f / 32.0f, 3.0f / 32.0f};
    const v2 Offsets[5] = {{1, 0}, {0, 1}, {0, 0}, {-1, 0}, {0, -1}};
    const v2 TexelSize = V2(1.0f / (float)W, 1.0f / (float)H);
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const v4 BasePosition = g.position_at(tT, tc);
            const v3 BaseNormal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x);
            const v4 BaseSH = tex2d_sample(tSH, tc.x, tc.y);
            const v4 cc = tex2d_sample(tCC, tc.x, tc.y);
            const v2 BaseCoCg = V2(cc.x, cc.y);
            const v4 ao = tex2d_sample(tAO, tc.x, tc.y);
            const v2 BaseAOSky = V2(ao.x, ao.y);
            float TotalWeight = 0.0f, SumLuminosity = 0.0f, SumSPP = 0.0f, SumMoment = 0.0f;
            v4 SumSH = V4(0.0f, 0.0f, 0.0f, 0.0f);
            v2 SumCoCg = V2(0.0f, 0.0f), SumAOSky = V2(0.0f, 0.0f);
            v4 Proj = mat4_mul(PV, V4(BasePosition.x, BasePosition.y, BasePosition.z, 1.0f));
            const v2 ReprojectedCoord = V2((Proj.x / Proj.w) * 0.5f + 0.5f, (Proj.y / Proj.w) * 0.5f + 0.5f);
            const float BaseLuminosity = tex2d_sample(tLum, tc.x, tc.y).x;
            const int BaseBlock = iclamp(cvt_trunc(floorf(tex2d_sample(tB, tc.x, tc.y).x * 255.0f)), 0, 127);
            int SuccessfulSamples = 0;
            bool DoBlockWeight = true, DoNormalWeight = true;
            float Tol = 0.75f;
            const float DistanceToPlayer = distance(V3(BasePosition.x, BasePosition.y, BasePosition.z), g.origin);
            if (DistanceToPlayer < 4.0f) Tol = 0.3f;
            else if (DistanceToPlayer < 6.0f) Tol = 0.65f;
            else if (DistanceToPlayer < 8.0f) Tol = 0.85f;
            else if (DistanceToPlayer < 16.0f) Tol = 1.414f;
            else if (DistanceToPlayer < 32.0f) Tol = 2.4f;
            else if (DistanceToPlayer < 48.0f) { Tol = 3.5f; DoBlockWeight = false; }
            else if (DistanceToPlayer < 64.0f) { Tol = 4.2f; DoBlockWeight = false; }
            else if (DistanceToPlayer < 96.0f) { Tol = 6.25f; DoBlockWeight = false; DoNormalWeight = false; }
            else if (DistanceToPlayer < 128.0f) { Tol = 9.0f; DoBlockWeight = false; DoNormalWeight = false; }
            else if (DistanceToPlayer < 200.0f) { Tol = 14.0f; DoBlockWeight = false; DoNormalWeight = false; }
            for (int i = 0; i < 5; ++i) {
                const v2 sc = V2(ReprojectedCoord.x + (Offsets[i].x + 0.0f) * TexelSize.x, ReprojectedCoord.y + (Offsets[i].y + 0.0f) * TexelSize.y);
                const float b = 0.0035f;  /* InThresholdedScreenSpace (:101-105) */
                if (!(sc.x < 1.0f - b && sc.x > b && sc.y < 1.0f - b && sc.y > b)) continue;
                const v4 Prev = g.position_at(qT, sc);
                const v3 PrevNormal = normal_from_id(tex2d_sample(qN, sc.x, sc.y).x);
                const v3 d = V3(fabsf(BasePosition.x - Prev.x), fabsf(BasePosition.y - Prev.y), fabsf(BasePosition.z - Prev.z));
                const float PositionError = dot(d, d);
                const float CurrentWeight = Weights[i];
                const int SampleBlock = iclamp(cvt_trunc(floorf(tex2d_sample(qB, sc.x, sc.y).x * 255.0f)), 0, 127);
                bool SampleValid = false;
                if (PositionError < Tol && ((Prev.w < 0.0f) == (BasePosition.w < 0.0f))) {
                    SampleValid = true;
                    if (DoNormalWeight && (PrevNormal.x != BaseNormal.x || PrevNormal.y != BaseNormal.y || PrevNormal.z != BaseNormal.z)) SampleValid = false;
                    if (DoBlockWeight && BaseBlock != SampleBlock) SampleValid = false;
                }
                if (SampleValid) {
                    const v4 u = tex2d_sample(pUt, sc.x, sc.y), s = tex2d_sample(pSH, sc.x, sc.y), c2 = tex2d_sample(pCC, sc.x, sc.y), a2 = tex2d_sample(pAO, sc.x, sc.y);
                    SumSH = V4(SumSH.x + s.x * CurrentWeight, SumSH.y + s.y * CurrentWeight, SumSH.z + s.z * CurrentWeight, SumSH.w + s.w * CurrentWeight);
                    SumCoCg = V2(SumCoCg.x + c2.x * CurrentWeight, SumCoCg.y + c2.y * CurrentWeight);
                    SumSPP += u.x * CurrentWeight;
                    SumMoment += u.y * CurrentWeight;
                    SumLuminosity += u.z * CurrentWeight;
                    SumAOSky = V2(SumAOSky.x + a2.x * CurrentWeight, SumAOSky.y + a2.y * CurrentWeight);
                    TotalWeight += CurrentWeight;
                    SuccessfulSamples++;
                }
            }
            if (TotalWeight > 0.001f) {
                SumSH = V4(SumSH.x / TotalWeight, SumSH.y / TotalWeight, SumSH.z / TotalWeight, SumSH.w / TotalWeight);
                SumCoCg = V2(SumCoCg.x / TotalWeight, SumCoCg.y / TotalWeight);
                SumMoment /= TotalWeight; SumSPP /= TotalWeight; SumLuminosity /= TotalWeight;
                SumAOSky = V2(SumAOSky.x / TotalWeight, SumAOSky.y / TotalWeight);
            } else {
                SuccessfulSamples = 0;
            }
            const float Increment = p->be_useful ? 1.0f : 0.0f;
            float SumSPPAndIncrement = SumSPP + Increment;
            if (SuccessfulSamples <= 0) SumSPPAndIncrement = 0.01f;
            float BlendFactor = gmax(1.0f / SumSPPAndIncrement, 0.05f);
            const float MomentFactor = gmax(1.0f / SumSPPAndIncrement, 0.05f);
            if (!p->be_useful) BlendFactor = 0.99f;
            float UtilitySPP = SumSPPAndIncrement;
            if (SuccessfulSamples <= 0) UtilitySPP = 0.0f;
            const float UtilityMoment = (1.0f - MomentFactor) * SumMoment + MomentFactor * (BaseLuminosity * BaseLuminosity);
            const float StoreLuma = gmix(SumLuminosity, BaseLuminosity, BlendFactor);
            v4 oSH = V4(gmix(SumSH.x, BaseSH.x, BlendFactor), gmix(SumSH.y, BaseSH.y, BlendFactor), gmix(SumSH.z, BaseSH.z, BlendFactor), gmix(SumSH.w, BaseSH.w, BlendFactor));
            v2 oCC = V2(gmix(SumCoCg.x, BaseCoCg.x, BlendFactor), gmix(SumCoCg.y, BaseCoCg.y, BlendFactor));
            v2 oAO = V2(gmix(SumAOSky.x, BaseAOSky.x, BlendFactor), gmix(SumAOSky.y, BaseAOSky.y, BlendFactor));
            if (SuccessfulSamples <= 0) { oSH = BaseSH; oCC = BaseCoCg; oAO = BaseAOSky; }
            const size_t i = (size_t)py * W + px;
            out->sh[4 * i] = float_to_half(gclampf(oSH.x, -100.0f, 100.0f)); out->sh[4 * i + 1] = float_to_half(gclampf(oSH.y, -100.0f, 100.0f));
            out->sh[4 * i + 2] = float_to_half(gclampf(oSH.z, -100.0f, 100.0f)); out->sh[4 * i + 3] = float_to_half(gclampf(oSH.w, -100.0f, 100.0f));
            out->cocg[2 * i] = float_to_half(gclampf(oCC.x, -10.0f, 100.0f)); out->cocg[2 * i + 1] = float_to_half(gclampf(oCC.y, -10.0f, 100.0f));
            out->x[3 * i] = float_to_half(gclampf(UtilitySPP, -150.0f, 150.0f)); out->x[3 * i + 1] = float_to_half(gclampf(UtilityMoment, -150.0f, 150.0f));
            out->x[3 * i + 2] = float_to_half(gclampf(StoreLuma, -150.0f, 150.0f));
            out->aosky[2 * i] = float_to_unorm8(gclampf(oAO.x, 0.0f, 1.0f)); out->aosky[2 * i + 1] = float_to_unorm8(gclampf(oAO.y, 0.0f, 1.0f));
        }
}

/* VarianceEstimate.glsl main() (:75-184).  in->x = u_Utility (RGB16F: accumulated frames, second moment, luminance);
 * out->x = o_Variance (R16F); out->aosky is not written (VarianceFBO has three attachments, Pipeline.cpp:1153).
 * GetPositionAt(SampleCoord) is only consumed through .w, the sampled distance. */
extern "C" void vxo_svgf_variance(const vxrt_svgf_variance_params* p, const vxo_svgf_set* in, const uint16_t* g_t, const uint8_t* g_normal,
                                  const vxo_svgf_out* out) {
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = from_half(in->sh, 4 * n), fcc = from_half(in->cocg, 2 * n), fut = from_half(in->x, 3 * n);
    auto ft = from_half(g_t, n), fn = from_u8(g_normal, n);
    const Tex2D tSH = view(fsh, W, H, 4, true), tCC = view(fcc, W, H, 2, true), tUt = view(fut, W, H, 3, true);
    const Tex2D tT = view(ft, W, H, 1, true), tN = view(fn, W, H, 1, false);
    const v2 TexelSize = V2(1.0f / (float)W, 1.0f / (float)H);
    const bool aggressive = p->aggressive_disocclusion != 0;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const float BaseDist = tex2d_sample(tT, tc.x, tc.y).x;
            const v3 BaseNormal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x);
            const v4 BaseUtility = tex2d_sample(tUt, tc.x, tc.y);
            const v4 BaseSH = tex2d_sample(tSH, tc.x, tc.y);
            const v4 bcc = tex2d_sample(tCC, tc.x, tc.y);
            const float BaseLuminosity = sh_to_y(BaseSH);
            const float ACCUMULATED_FRAMES = BaseUtility.x, BaseMoment = BaseUtility.y;
            v4 oSH = BaseSH;
            v2 oCC = V2(bcc.x, bcc.y);
            float Variance = BaseMoment - BaseLuminosity * BaseLuminosity;
            if (p->do_spatial) {
                const float THRESH = aggressive ? 4.0f + 4.0f + 4.0f : 4.0f + 4.0f;
                if (ACCUMULATED_FRAMES < THRESH) {
                    const float ColorPhi = aggressive ? 5.0f : 5.0f * 2.0f;
                    const int K = aggressive ? 4 : 1;
                    float TotalWeight = 0.0f, TotalMoment = 0.0f, TotalLuminosity = 0.0f, TotalWeight2 = 0.0f;
                    v4 TotalSH = V4(0.0f, 0.0f, 0.0f, 0.0f);
                    v2 TotalCoCg = V2(0.0f, 0.0f);
                    for (int x = -K; x <= K; ++x)
                        for (int y = -K; y <= K; ++y) {
                            const v2 sc = V2(tc.x + (float)x * TexelSize.x, tc.y + (float)y * TexelSize.y);
                            if (!in_screen_space(sc)) continue;
                            const float SampleDist = tex2d_sample(tT, sc.x, sc.y).x;
                            const v3 SampleNormal = normal_from_id(tex2d_sample(tN, sc.x, sc.y).x);
                            const float SampleMoment = tex2d_sample(tUt, sc.x, sc.y).y;
                            const v4 SampleSH = tex2d_sample(tSH, sc.x, sc.y);
                            const v4 scc = tex2d_sample(tCC, sc.x, sc.y);
                            const float SampleLuminosity = sh_to_y(SampleSH);
                            const float NormalWeight = powf(gmax(dot(BaseNormal, SampleNormal), 0.0f), 16.0f);
                            const float DepthWeight = powf(expf(-fabsf(SampleDist - BaseDist)), 2.0f);
                            const float LuminosityWeight = fabsf(SampleLuminosity - BaseLuminosity) / ColorPhi;
                            float Weight = expf(-LuminosityWeight) * NormalWeight * DepthWeight;
                            float Weight_2 = Weight;
                            Weight = gmax(Weight, 0.000000015f);
                            Weight_2 = gmax(Weight_2, 0.0000000015f);
                            TotalWeight += Weight;
                            TotalMoment += SampleMoment * Weight_2;
                            TotalSH = V4(TotalSH.x + SampleSH.x * Weight, TotalSH.y + SampleSH.y * Weight, TotalSH.z + SampleSH.z * Weight, TotalSH.w + SampleSH.w * Weight);
                            TotalCoCg = V2(TotalCoCg.x + scc.x * Weight, TotalCoCg.y + scc.y * Weight);
                            TotalLuminosity += SampleLuminosity * Weight_2;
                            TotalWeight2 += Weight_2;
                        }
                    if (TotalWeight > 0.0f) {
                        TotalMoment /= TotalWeight2;
                        TotalLuminosity /= TotalWeight2;
                        TotalCoCg = V2(TotalCoCg.x / TotalWeight, TotalCoCg.y / TotalWeight);
                        TotalSH = V4(TotalSH.x / TotalWeight, TotalSH.y / TotalWeight, TotalSH.z / TotalWeight, TotalSH.w / TotalWeight);
                    }
                    oSH = TotalSH; oCC = TotalCoCg;
                    Variance = TotalMoment - TotalLuminosity * TotalLuminosity;
                    Variance *= 3.0f;
                }
                Variance *= THRESH / ACCUMULATED_FRAMES;
            }
            const size_t i = (size_t)py * W + px;
            if (p->do_spatial) {   /* the !DO_SPATIAL path returns before the clamps (:96-101) */
                oSH = V4(gclampf(oSH.x, -100.0f, 100.0f), gclampf(oSH.y, -100.0f, 100.0f), gclampf(oSH.z, -100.0f, 100.0f), gclampf(oSH.w, -100.0f, 100.0f));
                oCC = V2(gclampf(oCC.x, -10.0f, 100.0f), gclampf(oCC.y, -10.0f, 100.0f));
                Variance = gclampf(Variance, -1.0f, 50.0f);
            }
            out->sh[4 * i] = float_to_half(oSH.x); out->sh[4 * i + 1] = float_to_half(oSH.y);
            out->sh[4 * i + 2] = float_to_half(oSH.z); out->sh[4 * i + 3] = float_to_half(oSH.w);
            out->cocg[2 * i] = float_to_half(oCC.x); out->cocg[2 * i + 1] = float_to_half(oCC.y);
            out->x[i] = float_to_half(Variance);
        }
}

/* TweakVariance (SpatialFilter.glsl:170-176) */
static inline float tweak_variance(float V, float E) {
    float F = gclampf(V, 0.0f, 1.0f);
    const float T = 1.0f - F;
    return F * powf(T, E + 6.0f);
}

/* SpatialFilter.glsl main() (:178-364), one a-trous iteration.  in->x = u_VarianceTexture (R16F), in->aosky = u_AO (the
 * temporal set's for iteration 0, Pipeline.cpp:2620-2636); temporal_utility = u_TemporalMoment (RGB16F, .x = accumulated
 * frames).  u_Utility, u_BlockIDTexture and BaseUtility are bound / sampled by the reference but never used.
 * `BaseDepth < 0.0f == DepthDiff < 0.0f` parses as (BaseDepth < 0) == (DepthDiff < 0); DepthDiff is an abs(). */
extern "C" void vxo_svgf_spatial(const vxrt_svgf_spatial_params* p, const vxo_svgf_set* in, const uint16_t* temporal_utility, const uint16_t* g_t,
                                 const uint8_t* g_normal, const vxo_svgf_out* out) {
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = from_half(in->sh, 4 * n), fcc = from_half(in->cocg, 2 * n), fvar = from_half(in->x, n), fao = from_u8(in->aosky, 2 * n);
    auto fut = from_half(temporal_utility, 3 * n), ft = from_half(g_t, n), fn = from_u8(g_normal, n);
    const Tex2D tSH = view(fsh, W, H, 4, true), tCC = view(fcc, W, H, 2, true), tVar = view(fvar, W, H, 1, true), tAO = view(fao, W, H, 2, true);
    const Tex2D tUt = view(fut, W, H, 3, true), tT = view(ft, W, H, 1, true), tN = view(fn, W, H, 1, false);
    const float AtrousWeights[3] = {1.0f, 2.0f / 3.0f, 1.0f / 6.0f};
    const float Gaussian[2] = {0.60283f, 0.198585f};
    const v2 TexelSize = V2(1.0f / (float)W, 1.0f / (float)H);       /* 1 / u_Dimensions */
    const v2 TexelSizeSH = V2(1.0f / (float)W, 1.0f / (float)H);     /* 1 / textureSize(u_SH, 0) */
    const bool FilterAO = p->step <= 4;
    const bool FilterSky = p->step <= 6 || FilterAO;
    const int K = p->large_kernel ? 2 : 1;
    const float AdditionalScale = gmix(1.0f, 2.4f, p->resolution_scale);
    const float tmod = p->time * 100.493850275f;
    const float toff = tmod - 500.0f * floorf(tmod / 500.0f);        /* mod(u_Time * 100.49.., 500) */
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            /* GradientNoise (:163-168) */
            const float cx = ((float)px + 0.5f) + toff, cy = ((float)py + 0.5f) + toff;
            const float noise = gfract(52.9829189f * gfract(0.06711056f * cx + 0.00583715f * cy));
            const float js = (noise - 0.5f) * ((float)p->step * 0.8f);
            const int Jx = cvt_trunc(js), Jy = Jx;
            const float BaseDepth = tex2d_sample(tT, tc.x, tc.y).x;
            const v3 BaseNormal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x);
            const v4 BaseSH = tex2d_sample(tSH, tc.x, tc.y);
            const v4 bcc = tex2d_sample(tCC, tc.x, tc.y);
            const v2 BaseCoCg = V2(bcc.x, bcc.y);
            const float BaseLuminance = sh_to_y(BaseSH);
            /* GaussianVariance (:98-133) */
            float BaseVariance = 0.0f, VarianceSum = 0.0f, TotalKernel = 0.0f;
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    const v2 sc = V2(tc.x + (float)x * TexelSizeSH.x, tc.y + (float)y * TexelSizeSH.y);
                    if (!in_screen_space(sc)) continue;
                    const float KernelValue = Gaussian[x < 0 ? -x : x] * Gaussian[y < 0 ? -y : y];
                    const float V = tex2d_sample(tVar, sc.x, sc.y).x;
                    if (x == 0 && y == 0) BaseVariance = V;
                    VarianceSum += V * KernelValue;
                    TotalKernel += KernelValue;
                }
            const float VarianceEstimate = VarianceSum / gmax(TotalKernel, 0.01f);
            const v4 bao = tex2d_sample(tAO, tc.x, tc.y);
            const v2 BaseAOSky = V2(bao.x, bao.y);
            v4 oSH = BaseSH; v2 oCC = BaseCoCg; float oVar = BaseVariance; v2 oAO = BaseAOSky;
            if (p->do_spatial) {
                v4 TotalSH = BaseSH; v2 TotalCoCg = BaseCoCg; float TotalWeight = 1.0f, TotalVariance = BaseVariance;
                v2 TotalAOSky = BaseAOSky; float TotalAOWeight = 1.0f;
                const float AccumulatedFrames = tex2d_sample(tUt, tc.x, tc.y).x;
                const bool DoStrongSpatial = AccumulatedFrames <= 8.0f && p->aggressive_disocclusion && p->step <= 8;
                float CurveExponent = 0.0f;
                if (VarianceEstimate < 0.01f) CurveExponent = 128.0f;
                else if (VarianceEstimate < 0.025f) CurveExponent = 112.0f;
                else if (VarianceEstimate < 0.05f) CurveExponent = 96.0f;
                else if (VarianceEstimate < 0.075f) CurveExponent = 84.0f;
                else if (VarianceEstimate < 0.1f) CurveExponent = 70.0f;
                const float TweakedVariance = VarianceEstimate < 0.1f ? tweak_variance(VarianceEstimate, CurveExponent) : VarianceEstimate;
                float PhiColor = sqrtf(gmax(0.0f, 0.000001f + TweakedVariance));
                PhiColor /= gmax(p->color_phi_bias, 0.1f);
                for (int x = -K; x <= K; ++x)
                    for (int y = -K; y <= K; ++y) {
                        if (x == 0 && y == 0) continue;
                        const v2 sc = V2(tc.x + (((float)x * (float)p->step) * AdditionalScale + (float)Jx * 0.5f) * TexelSize.x,
                                         tc.y + (((float)y * (float)p->step) * AdditionalScale + (float)Jy * 0.5f) * TexelSize.y);
                        if (!in_screen_space(sc)) continue;
                        const float SampleDepth = tex2d_sample(tT, sc.x, sc.y).x;
                        const float DepthDiff = fabsf(SampleDepth - BaseDepth);
                        const v3 SampleNormal = normal_from_id(tex2d_sample(tN, sc.x, sc.y).x);
                        if ((BaseDepth < 0.0f) == (DepthDiff < 0.0f)) {
                            const v4 SampleSH = tex2d_sample(tSH, sc.x, sc.y);
                            const v4 scc = tex2d_sample(tCC, sc.x, sc.y);
                            const float SampleLuma = sh_to_y(SampleSH);
                            const float SampleVariance = tex2d_sample(tVar, sc.x, sc.y).x;
                            float NormalWeight = powf(gmax(dot(BaseNormal, SampleNormal), 0.0f), 32.0f);
                            NormalWeight = gclampf(NormalWeight, 0.001f, 1.0f);
                            const float LuminosityWeight = fabsf(SampleLuma - BaseLuminance) / PhiColor;
                            const float DepthWeight = gclampf(powf(expf(-gmax(DepthDiff, 0.00001f)), 2.0f), 0.0001f, 1.0f);
                            float Weight = DoStrongSpatial ? (NormalWeight * DepthWeight) : (expf(-LuminosityWeight) * NormalWeight * DepthWeight);
                            Weight = gclampf(Weight, 0.001f, 1.0f);
                            const float XWeight = AtrousWeights[x < 0 ? -x : x], YWeight = AtrousWeights[y < 0 ? -y : y];
                            Weight = (XWeight * YWeight) * Weight;
                            Weight = gmax(Weight, 0.00000001f);
                            TotalSH = V4(TotalSH.x + SampleSH.x * Weight, TotalSH.y + SampleSH.y * Weight, TotalSH.z + SampleSH.z * Weight, TotalSH.w + SampleSH.w * Weight);
                            TotalCoCg = V2(TotalCoCg.x + scc.x * Weight, TotalCoCg.y + scc.y * Weight);
                            TotalVariance += (Weight * Weight) * SampleVariance;
                            TotalWeight += Weight;
                            if (FilterSky || FilterAO) {
                                const float CurrAOWeight = gclampf((XWeight * YWeight) * NormalWeight * DepthWeight, 0.000001f, 1.0f);
                                const v4 a = tex2d_sample(tAO, sc.x, sc.y);
                                TotalAOSky.x += a.x * CurrAOWeight;
                                TotalAOSky.y += a.y * CurrAOWeight;
                                TotalAOWeight += CurrAOWeight;
                            }
                        }
                    }
                oSH = V4(TotalSH.x / TotalWeight, TotalSH.y / TotalWeight, TotalSH.z / TotalWeight, TotalSH.w / TotalWeight);
                oCC = V2(TotalCoCg.x / TotalWeight, TotalCoCg.y / TotalWeight);
                oVar = TotalVariance / (TotalWeight * TotalWeight);
                oAO = V2(TotalAOSky.x / TotalAOWeight, TotalAOSky.y / TotalAOWeight);
                if (!FilterAO) oAO.x = BaseAOSky.x;
            }
            const size_t i = (size_t)py * W + px;
            if (p->do_spatial) {   /* the !DO_SPATIAL path returns before the clamps (:205-211) */
                oSH = V4(gclampf(oSH.x, -100.0f, 100.0f), gclampf(oSH.y, -100.0f, 100.0f), gclampf(oSH.z, -100.0f, 100.0f), gclampf(oSH.w, -100.0f, 100.0f));
                oCC = V2(gclampf(oCC.x, -10.0f, 100.0f), gclampf(oCC.y, -10.0f, 100.0f));
                oVar = gclampf(oVar, -1.0f, 50.0f);
                oAO = V2(gclampf(oAO.x, 0.0f, 1.0f), gclampf(oAO.y, 0.0f, 1.0f));
            }
            out->sh[4 * i] = float_to_half(oSH.x); out->sh[4 * i + 1] = float_to_half(oSH.y);
            out->sh[4 * i + 2] = float_to_half(oSH.z); out->sh[4 * i + 3] = float_to_half(oSH.w);
            out->cocg[2 * i] = float_to_half(oCC.x); out->cocg[2 * i + 1] = float_to_half(oCC.y);
            out->x[i] = float_to_half(oVar);
            out->aosky[2 * i] = float_to_unorm8(oAO.x); out->aosky[2 * i + 1] = float_to_unorm8(oAO.y);
        }
}

/* Spatial3x3Initial.glsl main() (:103-175), the 3 x 3 pass in front of the temporal filter (Core/Pipeline.cpp:2381-2424).
 * in->x = u_Utility (R16F, passed through).  The eight neighbours are taken x-major (x = -1, 0, 1 outside, y inside); a tap counts
 * when its world position is less than 1 from the centre's (GetPositionAt :34-38 with v_RayOrigin = u_VertInverseView[3],
 * FBOVert.glsl:21).  Weight = exp(-|dY| / 4 - pow(max(n.n', 0), 16)): the normal term is SUBTRACTED in the exponent (:137-139), so
 * equal normals lower the weight; kept as written.  Jitter (:120) is computed and never used. */
extern "C" void vxo_svgf_prespatial(const vxrt_svgf_prespatial_params* p, const vxo_svgf_set* in, const uint16_t* g_t, const uint8_t* g_normal,
                                    int32_t gw, int32_t gh, const vxo_svgf_out* out) {
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H, ng = (size_t)gw * gh;
    auto fsh = from_half(in->sh, 4 * n), fcc = from_half(in->cocg, 2 * n), fao = from_u8(in->aosky, 2 * n), fut = from_half(in->x, n);
    auto ft = from_half(g_t, ng), fn = from_u8(g_normal, ng);
    const Tex2D tSH = view(fsh, W, H, 4, true), tCC = view(fcc, W, H, 2, true), tAO = view(fao, W, H, 2, true), tUt = view(fut, W, H, 1, true);
    const Tex2D tT = view(ft, gw, gh, 1, true), tN = view(fn, gw, gh, 1, false);
    const float AtrousWeights[3] = {1.0f, 2.0f / 3.0f, 1.0f / 6.0f};
    const v2 TexelSize = V2(1.0f / (float)W, 1.0f / (float)H);
    const float PhiColor = 4.0f;
    GBuf g;
    g.inv_view = p->inv_view; g.inv_proj = p->inv_projection;
    g.origin = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            const v4 BasePosition = g.position_at(tT, tc);
            const v3 BaseNormal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x);
            const v4 BaseSH = tex2d_sample(tSH, tc.x, tc.y);
            const v4 bcc = tex2d_sample(tCC, tc.x, tc.y);
            const float BaseLuminance = sh_to_y(BaseSH);
            const v4 bao = tex2d_sample(tAO, tc.x, tc.y);
            v4 TotalSH = BaseSH;
            v2 TotalCoCg = V2(bcc.x, bcc.y), TotalAOSky = V2(bao.x, bao.y);
            float TotalWeight = 1.0f, TotalAOWeight = 1.0f;
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    if (x == 0 && y == 0) continue;
                    const v2 sc = V2(tc.x + ((float)x * 1.0f) * TexelSize.x, tc.y + ((float)y * 1.0f) * TexelSize.y);
                    if (!(sc.x > 0.0f && sc.x < 1.0f && sc.y > 0.0f && sc.y < 1.0f)) continue;
                    const v4 sp = g.position_at(tT, sc);
                    const v3 d = V3(fabsf(sp.x - BasePosition.x), fabsf(sp.y - BasePosition.y), fabsf(sp.z - BasePosition.z));
                    const float DistSqr = dot(d, d);
                    if (!(DistSqr < 1.0f)) continue;
                    const v4 SampleSH = tex2d_sample(tSH, sc.x, sc.y);
                    const v4 scc = tex2d_sample(tCC, sc.x, sc.y);
                    const v3 SampleNormal = normal_from_id(tex2d_sample(tN, sc.x, sc.y).x);
                    const float SampleLuma = sh_to_y(SampleSH);
                    const float NormalWeight = powf(gmax(dot(BaseNormal, SampleNormal), 0.0f), 16.0f);
                    const float LuminosityWeight = fabsf(SampleLuma - BaseLuminance) / PhiColor;
                    float Weight = expf(-LuminosityWeight - NormalWeight);
                    Weight = gmax(Weight, 0.01f);
                    Weight = (AtrousWeights[x < 0 ? -x : x] * AtrousWeights[y < 0 ? -y : y]) * Weight;
                    Weight = gmax(Weight, 0.01f);
                    Weight = gclampf(Weight, 0.0f, 1.0f);
                    TotalSH = V4(TotalSH.x + SampleSH.x * Weight, TotalSH.y + SampleSH.y * Weight, TotalSH.z + SampleSH.z * Weight, TotalSH.w + SampleSH.w * Weight);
                    TotalCoCg = V2(TotalCoCg.x + scc.x * Weight, TotalCoCg.y + scc.y * Weight);
                    TotalWeight += Weight;
                    const v4 sao = tex2d_sample(tAO, sc.x, sc.y);
                    TotalAOSky = V2(TotalAOSky.x + sao.x * Weight, TotalAOSky.y + sao.y * Weight);
                    TotalAOWeight += Weight;
                }
            TotalWeight = gmax(TotalWeight, 0.01f);
            const float aw = gmax(TotalAOWeight, 0.01f);
            const size_t i = (size_t)py * W + px;
            out->sh[4 * i] = float_to_half(TotalSH.x / TotalWeight); out->sh[4 * i + 1] = float_to_half(TotalSH.y / TotalWeight);
            out->sh[4 * i + 2] = float_to_half(TotalSH.z / TotalWeight); out->sh[4 * i + 3] = float_to_half(TotalSH.w / TotalWeight);
            out->cocg[2 * i] = float_to_half(TotalCoCg.x / TotalWeight); out->cocg[2 * i + 1] = float_to_half(TotalCoCg.y / TotalWeight);
            out->x[i] = float_to_half(tex2d_sample(tUt, tc.x, tc.y).x);   /* o_Utility = BaseUtility */
            out->aosky[2 * i] = float_to_unorm8(TotalAOSky.x / aw); out->aosky[2 * i + 1] = float_to_unorm8(TotalAOSky.y / aw);
        }
}
