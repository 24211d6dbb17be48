// svgf.cu — SVGF denoiser chain of the diffuse GI: temporal accumulation, variance estimate and the
// a-trous spatial filter of Core/Shaders/SVGF/{TemporalFilter,VarianceEstimate,SpatialFilter}.glsl, dispatched by
// Core/Pipeline.cpp:2428-2700.  One thread per pixel, a warp covers an 8x4 pixel tile (taps of neighbouring pixels
// share sectors), attachments are read with the sampler model of texture.cuh (LINEAR + REPEAT for the float and RG8
// images, NEAREST for the R8 G-buffer planes) and written in their final formats.
//
// The temporal pass has no transcendental on its path and is bit-identical to the oracle; the variance and spatial
// passes go through expf / powf (CUDA vs libm: <= 2 ulp), tolerance in tests/test_gpu_svgf.py.
#include "ctx.h"
#include "filter_sampler.cuh"

namespace {

struct SetIn {
    const uint16_t* __restrict__ sh;    // RGBA16F
    const uint16_t* __restrict__ cocg;  // RG16F
    const uint16_t* __restrict__ x;     // RGB16F utility | R16F variance / luminance
    const uint8_t* __restrict__ aosky;  // RG8
    int w, h;
};
struct SetOut {
    uint16_t* __restrict__ sh;
    uint16_t* __restrict__ cocg;
    uint16_t* __restrict__ x;
    uint8_t* __restrict__ aosky;
};
struct GBufIn {
    const uint16_t* __restrict__ t;  // R16F, LINEAR
    const uint8_t* __restrict__ n;   // R8, NEAREST
    const uint8_t* __restrict__ b;   // R8, NEAREST
    int w, h;
};

VXD int normal_at(const GBufIn& g, int o, const float* __restrict__ lut) { return normal_index(lut[__ldg(g.n + o)]); }
VXD int block_at(const GBufIn& g, int o, const float* __restrict__ lut) { return iclamp(cvt_trunc(floorf(lut[__ldg(g.b + o)] * 255.0f)), 0, 127); }

#ifndef VX_SVGF_OCC
#define VX_SVGF_OCC 4   // CTAs per SM the temporal / spatial kernels are compiled for (register cap 64 / 48 / 40 at 4 / 5 / 6)
#endif

// ---------------------------------------------------------------------------------------------------------------
// TemporalFilter.glsl main() (:133-344)
struct TemporalArgs {
    float inv_view[16], inv_proj[16], prev_pv[16];
    int width, height, row0, row1, be_useful;
    SetIn cur, hist;   // both width x height
    GBufIn g, pg;
    SetOut out;
};

__global__ void __launch_bounds__(256, VX_SVGF_OCC) svgf_temporal_kernel(const __grid_constant__ TemporalArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const bool same = a.g.w == a.width && a.g.h == a.height;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const f3 origin = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const Tap ts = make_tap(a.width, a.height, tc);
    Tap tg = ts;
    if (!same) tg = make_tap(a.g.w, a.g.h, tc);
    const int ng = nearest_offset(a.g.w, a.g.h, tc);
    const float BaseDist = sample_r16(a.g.t, tg);
    const f3 BasePos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * BaseDist;
    const int BaseNormal = normal_at(a.g, ng, lut);
    const int BaseBlock = block_at(a.g, ng, lut);
    float BaseSH[4], BaseCoCg[2], BaseAO[2];
    sample_rgba16(a.cur.sh, ts, BaseSH);
    sample_rg16(a.cur.cocg, ts, BaseCoCg);
    sample_rg8(a.cur.aosky, ts, lut, BaseAO);
    const float BaseLuminosity = sample_r16(a.cur.x, ts);
    const f4 Proj = mat4_mul(a.prev_pv, F4(BasePos.x, BasePos.y, BasePos.z, 1.0f));
    const f2 Reproj = F2((Proj.x / Proj.w) * 0.5f + 0.5f, (Proj.y / Proj.w) * 0.5f + 0.5f);

    bool DoBlockWeight = true, DoNormalWeight = true;
    float Tol = 0.75f;
    const float d = distance(BasePos, origin);
    if (d < 4.0f) Tol = 0.3f;
    else if (d < 6.0f) Tol = 0.65f;
    else if (d < 8.0f) Tol = 0.85f;
    else if (d < 16.0f) Tol = 1.414f;
    else if (d < 32.0f) Tol = 2.4f;
    else if (d < 48.0f) { Tol = 3.5f; DoBlockWeight = false; }
    else if (d < 64.0f) { Tol = 4.2f; DoBlockWeight = false; }
    else if (d < 96.0f) { Tol = 6.25f; DoBlockWeight = false; DoNormalWeight = false; }
    else if (d < 128.0f) { Tol = 9.0f; DoBlockWeight = false; DoNormalWeight = false; }
    else if (d < 200.0f) { Tol = 14.0f; DoBlockWeight = false; DoNormalWeight = false; }

    float TotalWeight = 0.0f, SumLuminosity = 0.0f, SumSPP = 0.0f, SumMoment = 0.0f;
    float SumSH[4] = {0.0f, 0.0f, 0.0f, 0.0f}, SumCoCg[2] = {0.0f, 0.0f}, SumAO[2] = {0.0f, 0.0f};
    int Successful = 0;
    const f2 Texel = F2(1.0f / (float)a.width, 1.0f / (float)a.height);
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
        // Offsets (1,0) (0,1) (0,0) (-1,0) (0,-1), Weights 3/32 3/32 9/64 3/32 3/32 (:150-156)
        const float ox = i == 0 ? 1.0f : (i == 3 ? -1.0f : 0.0f), oy = i == 1 ? 1.0f : (i == 4 ? -1.0f : 0.0f);
        const float w = i == 2 ? 9.0f / 64.0f : 3.0f / 32.0f;
        const f2 sc = F2(Reproj.x + (ox + 0.0f) * Texel.x, Reproj.y + (oy + 0.0f) * Texel.y);
        const float b = 0.0035f;
        if (!(sc.x < 1.0f - b && sc.x > b && sc.y < 1.0f - b && sc.y > b)) continue;
        const Tap hs = make_tap(a.width, a.height, sc);
        Tap hg = hs;
        if (!same) hg = make_tap(a.pg.w, a.pg.h, sc);
        const float PrevDist = sample_r16(a.pg.t, hg);
        const f3 PrevPos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, sc)) * PrevDist;
        const f3 e = F3(fabsf(BasePos.x - PrevPos.x), fabsf(BasePos.y - PrevPos.y), fabsf(BasePos.z - PrevPos.z));
        const float PositionError = dot(e, e);
        if (!(PositionError < Tol && ((PrevDist < 0.0f) == (BaseDist < 0.0f)))) continue;
        const int np = nearest_offset(a.pg.w, a.pg.h, sc);
        if (DoNormalWeight && normal_at(a.pg, np, lut) != BaseNormal) continue;
        if (DoBlockWeight && block_at(a.pg, np, lut) != BaseBlock) continue;
        float s[4], c2[2], a2[2];
        sample_rgba16(a.hist.sh, hs, s);
        sample_rg16(a.hist.cocg, hs, c2);
        sample_rg8(a.hist.aosky, hs, lut, a2);
#pragma unroll
        for (int k = 0; k < 4; ++k) SumSH[k] += s[k] * w;
        SumCoCg[0] += c2[0] * w; SumCoCg[1] += c2[1] * w;
        SumSPP += sample_rgb16_ch(a.hist.x, hs, 0) * w; SumMoment += sample_rgb16_ch(a.hist.x, hs, 1) * w;
        SumLuminosity += sample_rgb16_ch(a.hist.x, hs, 2) * w;
        SumAO[0] += a2[0] * w; SumAO[1] += a2[1] * w;
        TotalWeight += w;
        Successful++;
    }
    if (TotalWeight > 0.001f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) SumSH[k] /= TotalWeight;
        SumCoCg[0] /= TotalWeight; SumCoCg[1] /= TotalWeight;
        SumMoment /= TotalWeight; SumSPP /= TotalWeight; SumLuminosity /= TotalWeight;
        SumAO[0] /= TotalWeight; SumAO[1] /= TotalWeight;
    } else {
        Successful = 0;
    }
    float SppInc = SumSPP + (a.be_useful ? 1.0f : 0.0f);
    if (Successful <= 0) SppInc = 0.01f;
    float Blend = gmax(1.0f / SppInc, 0.05f);
    const float MomentFactor = Blend;
    if (!a.be_useful) Blend = 0.99f;
    const float UtilitySPP = Successful <= 0 ? 0.0f : SppInc;
    const float UtilityMoment = (1.0f - MomentFactor) * SumMoment + MomentFactor * (BaseLuminosity * BaseLuminosity);
    const float StoreLuma = gmix(SumLuminosity, BaseLuminosity, Blend);
    float oSH[4], oCC[2], oAO[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) oSH[k] = Successful <= 0 ? BaseSH[k] : gmix(SumSH[k], BaseSH[k], Blend);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        oCC[k] = Successful <= 0 ? BaseCoCg[k] : gmix(SumCoCg[k], BaseCoCg[k], Blend);
        oAO[k] = Successful <= 0 ? BaseAO[k] : gmix(SumAO[k], BaseAO[k], Blend);
    }
    const size_t i = (size_t)py * a.width + px;
    ushort4 o4;
    o4.x = float_to_half_bits(gclamp(oSH[0], -100.0f, 100.0f)); o4.y = float_to_half_bits(gclamp(oSH[1], -100.0f, 100.0f));
    o4.z = float_to_half_bits(gclamp(oSH[2], -100.0f, 100.0f)); o4.w = float_to_half_bits(gclamp(oSH[3], -100.0f, 100.0f));
    reinterpret_cast<ushort4*>(a.out.sh)[i] = o4;
    ushort2 o2;
    o2.x = float_to_half_bits(gclamp(oCC[0], -10.0f, 100.0f)); o2.y = float_to_half_bits(gclamp(oCC[1], -10.0f, 100.0f));
    reinterpret_cast<ushort2*>(a.out.cocg)[i] = o2;
    a.out.x[3 * i] = float_to_half_bits(gclamp(UtilitySPP, -150.0f, 150.0f));
    a.out.x[3 * i + 1] = float_to_half_bits(gclamp(UtilityMoment, -150.0f, 150.0f));
    a.out.x[3 * i + 2] = float_to_half_bits(gclamp(StoreLuma, -150.0f, 150.0f));
    uchar2 ao;
    ao.x = float_to_unorm8(gclamp(oAO[0], 0.0f, 1.0f)); ao.y = float_to_unorm8(gclamp(oAO[1], 0.0f, 1.0f));
    reinterpret_cast<uchar2*>(a.out.aosky)[i] = ao;
}

// ---------------------------------------------------------------------------------------------------------------
// VarianceEstimate.glsl main() (:75-184).  GetPositionAt(SampleCoord) is only consumed through .w (the sampled distance).
// The 9 x 9 bilateral pre-filter only runs for pixels with fewer than 12 accumulated frames; a CTA in which some pixel
// needs it stages its 32 x 8 tile plus a 5-texel apron in shared memory (converted to float once), and every tap
// is four shared-memory texels away.
struct VarianceArgs {
    int width, height, row0, row1, do_spatial, aggressive;
    SetIn in;  // temporal set: x = utility RGB16F
    GBufIn g;
    SetOut out;  // x = variance R16F; aosky not written
};

// pow(max(dot(n0, n1), 0), 16) over the values a pair of GetNormalFromID normals can produce (0, 1, 3): exact like powf
VXD float normal_weight16(int n0, int n1) {
    if (n0 == 6 && n1 == 6) return 43046720.0f;                         // powf(3, 16)
    if (n0 == 6) return (n1 == 0 || n1 == 2 || n1 == 5) ? 1.0f : 0.0f;  // (1,1,1) . axis = sign of the axis
    if (n1 == 6) return (n0 == 0 || n0 == 2 || n0 == 5) ? 1.0f : 0.0f;
    return n0 == n1 ? 1.0f : 0.0f;
}

constexpr int VAR_HALO = 5, VAR_TW = 32 + 2 * VAR_HALO, VAR_TH = 8 + 2 * VAR_HALO;

__global__ void __launch_bounds__(256) svgf_variance_kernel(const __grid_constant__ VarianceArgs a) {
    __shared__ float lut[256];
    __shared__ float4 s_sh[VAR_TH * VAR_TW];    // SH
    __shared__ float4 s_aux[VAR_TH * VAR_TW];   // CoCg.x, CoCg.y, second moment, distance
    __shared__ uint8_t s_n[VAR_TH * VAR_TW];    // normal index
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = a.row0 + blockIdx.y * 8;
    const int px = x0 + (warp & 3) * 8 + (lane & 7);
    const int py = y0 + (warp >> 2) * 4 + (lane >> 3);
    const bool active = px < a.width && py < a.row1;
    const bool same = a.g.w == a.in.w && a.g.h == a.in.h;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    float BaseDist = 0.0f, BaseLum = 0.0f, Frames = 0.0f, BaseMoment = 0.0f;
    int BaseNormal = 0;
    float oSH[4] = {0.0f, 0.0f, 0.0f, 0.0f}, oCC[2] = {0.0f, 0.0f};
    if (active) {
        const Tap ts = make_tap(a.in.w, a.in.h, tc);
        Tap tg = ts;
        if (!same) tg = make_tap(a.g.w, a.g.h, tc);
        BaseDist = sample_r16(a.g.t, tg);
        BaseNormal = normal_at(a.g, nearest_offset(a.g.w, a.g.h, tc), lut);
        Frames = sample_rgb16_ch(a.in.x, ts, 0);
        BaseMoment = sample_rgb16_ch(a.in.x, ts, 1);
        sample_rgba16(a.in.sh, ts, oSH);
        sample_rg16(a.in.cocg, ts, oCC);
        BaseLum = sh_to_y(oSH[3]);
    }
    float Variance = BaseMoment - BaseLum * BaseLum;
    const float THRESH = a.aggressive ? 4.0f + 4.0f + 4.0f : 4.0f + 4.0f;
    const bool filter = active && a.do_spatial && Frames < THRESH;
    const bool tiled = same && a.aggressive;   // the apron is sized for the 9 x 9 kernel
    if (__syncthreads_or(filter) && tiled) {
        for (int idx = threadIdx.x; idx < VAR_TH * VAR_TW; idx += 256) {
            const int ty = idx / VAR_TW, tx = idx - ty * VAR_TW;
            const int o = wrap_repeat(y0 - VAR_HALO + ty, a.in.h) * a.in.w + wrap_repeat(x0 - VAR_HALO + tx, a.in.w);
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(a.in.sh) + o);
            const float2 s0 = h2f(q.x), s1 = h2f(q.y), c = h2f(__ldg(reinterpret_cast<const uint32_t*>(a.in.cocg) + o));
            s_sh[idx] = make_float4(s0.x, s0.y, s1.x, s1.y);
            s_aux[idx] = make_float4(c.x, c.y, hf(__ldg(a.in.x + 3 * o + 1)), hf(__ldg(a.g.t + o)));
            s_n[idx] = (uint8_t)normal_at(a.g, o, lut);
        }
        __syncthreads();
    }
    if (filter) {
        const float ColorPhi = a.aggressive ? 5.0f : 5.0f * 2.0f;
        const int K = a.aggressive ? 4 : 1;
        const f2 Texel = F2(1.0f / (float)a.in.w, 1.0f / (float)a.in.h);
        float TotalWeight = 0.0f, TotalMoment = 0.0f, TotalLum = 0.0f, TotalWeight2 = 0.0f;
        float TotalSH[4] = {0.0f, 0.0f, 0.0f, 0.0f}, TotalCC[2] = {0.0f, 0.0f};
#pragma unroll 1
        for (int x = -K; x <= K; ++x) {
            const float scx = tc.x + (float)x * Texel.x;
            if (!(scx < 1.0f && scx > 0.0f)) continue;
            // column set-up of the tap (shared by the 9 taps of this column)
            const float u = scx * (float)a.in.w - 0.5f, fu = floorf(u);
            const float wa = u - fu, wia = 1.0f - wa;
            const int ci = iclamp(cvt_floor(fu) - (x0 - VAR_HALO), 0, VAR_TW - 2);
            const int cn = iclamp(cvt_floor(scx * (float)a.in.w) - (x0 - VAR_HALO), 0, VAR_TW - 1);
#pragma unroll 1
            for (int y = -K; y <= K; ++y) {
                const float scy = tc.y + (float)y * Texel.y;
                if (!(scy < 1.0f && scy > 0.0f)) continue;
                float sh[4], cc[2], SampleMoment, SampleDist;
                int SampleNormal;
                if (tiled) {
                    const float v = scy * (float)a.in.h - 0.5f, fv = floorf(v);
                    const float wb = v - fv, wib = 1.0f - wb;
                    const int rj = iclamp(cvt_floor(fv) - (y0 - VAR_HALO), 0, VAR_TH - 2);
                    const int rn = iclamp(cvt_floor(scy * (float)a.in.h) - (y0 - VAR_HALO), 0, VAR_TH - 1);
                    const int o = rj * VAR_TW + ci;
                    const float4 p00 = s_sh[o], p10 = s_sh[o + 1], p01 = s_sh[o + VAR_TW], p11 = s_sh[o + VAR_TW + 1];
                    const float4 q00 = s_aux[o], q10 = s_aux[o + 1], q01 = s_aux[o + VAR_TW], q11 = s_aux[o + VAR_TW + 1];
#define VX_BL(c00, c10, c01, c11) (((c00) * wia + (c10) * wa) * wib + ((c01) * wia + (c11) * wa) * wb)
                    sh[0] = VX_BL(p00.x, p10.x, p01.x, p11.x); sh[1] = VX_BL(p00.y, p10.y, p01.y, p11.y);
                    sh[2] = VX_BL(p00.z, p10.z, p01.z, p11.z); sh[3] = VX_BL(p00.w, p10.w, p01.w, p11.w);
                    cc[0] = VX_BL(q00.x, q10.x, q01.x, q11.x); cc[1] = VX_BL(q00.y, q10.y, q01.y, q11.y);
                    SampleMoment = VX_BL(q00.z, q10.z, q01.z, q11.z); SampleDist = VX_BL(q00.w, q10.w, q01.w, q11.w);
#undef VX_BL
                    SampleNormal = s_n[rn * VAR_TW + cn];
                } else {
                    const f2 sc = F2(scx, scy);
                    const Tap ss = make_tap(a.in.w, a.in.h, sc);
                    Tap sg = ss;
                    if (!same) sg = make_tap(a.g.w, a.g.h, sc);
                    SampleDist = sample_r16(a.g.t, sg);
                    SampleNormal = normal_at(a.g, nearest_offset(a.g.w, a.g.h, sc), lut);
                    SampleMoment = sample_rgb16_ch(a.in.x, ss, 1);
                    sample_rgba16(a.in.sh, ss, sh);
                    sample_rg16(a.in.cocg, ss, cc);
                }
                const float SampleLum = sh_to_y(sh[3]);
                const float NormalWeight = normal_weight16(BaseNormal, SampleNormal);
                const float ed = expf(-fabsf(SampleDist - BaseDist));
                const float DepthWeight = ed * ed;
                const float LumWeight = fabsf(SampleLum - BaseLum) / ColorPhi;
                float Weight = expf(-LumWeight) * NormalWeight * DepthWeight;
                const float Weight_2 = gmax(Weight, 0.0000000015f);
                Weight = gmax(Weight, 0.000000015f);
                TotalWeight += Weight;
                TotalMoment += SampleMoment * Weight_2;
#pragma unroll
                for (int k = 0; k < 4; ++k) TotalSH[k] += sh[k] * Weight;
                TotalCC[0] += cc[0] * Weight; TotalCC[1] += cc[1] * Weight;
                TotalLum += SampleLum * Weight_2;
                TotalWeight2 += Weight_2;
            }
        }
        if (TotalWeight > 0.0f) {
            TotalMoment /= TotalWeight2;
            TotalLum /= TotalWeight2;
            TotalCC[0] /= TotalWeight; TotalCC[1] /= TotalWeight;
#pragma unroll
            for (int k = 0; k < 4; ++k) TotalSH[k] /= TotalWeight;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) oSH[k] = TotalSH[k];
        oCC[0] = TotalCC[0]; oCC[1] = TotalCC[1];
        Variance = TotalMoment - TotalLum * TotalLum;
        Variance *= 3.0f;
    }
    if (!active) return;
    if (a.do_spatial) {
        Variance *= THRESH / Frames;
#pragma unroll
        for (int k = 0; k < 4; ++k) oSH[k] = gclamp(oSH[k], -100.0f, 100.0f);
        oCC[0] = gclamp(oCC[0], -10.0f, 100.0f); oCC[1] = gclamp(oCC[1], -10.0f, 100.0f);
        Variance = gclamp(Variance, -1.0f, 50.0f);
    }
    const size_t i = (size_t)py * a.width + px;
    ushort4 o4;
    o4.x = float_to_half_bits(oSH[0]); o4.y = float_to_half_bits(oSH[1]); o4.z = float_to_half_bits(oSH[2]); o4.w = float_to_half_bits(oSH[3]);
    reinterpret_cast<ushort4*>(a.out.sh)[i] = o4;
    ushort2 o2;
    o2.x = float_to_half_bits(oCC[0]); o2.y = float_to_half_bits(oCC[1]);
    reinterpret_cast<ushort2*>(a.out.cocg)[i] = o2;
    a.out.x[i] = float_to_half_bits(Variance);
}

// ---------------------------------------------------------------------------------------------------------------
// SpatialFilter.glsl main() (:178-364), one a-trous iteration
struct SpatialArgs {
    int width, height, row0, row1;
    int step, large_kernel, do_spatial, aggressive;
    float phi_bias, time_offset, additional_scale;
    SetIn in;                           // sh, cocg, x = u_VarianceTexture (R16F), aosky = u_AO; width x height
    const uint16_t* __restrict__ temporal_utility;  // u_TemporalMoment RGB16F, width x height
    GBufIn g;
    SetOut out;
};

__global__ void __launch_bounds__(256, VX_SVGF_OCC) svgf_spatial_kernel(const __grid_constant__ SpatialArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const bool same = a.g.w == a.width && a.g.h == a.height;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    // GradientNoise (:163-168); Jitter = ivec2(float) has both components equal
    const float cx = ((float)px + 0.5f) + a.time_offset, cy = ((float)py + 0.5f) + a.time_offset;
    const float noise = gfract(52.9829189f * gfract(0.06711056f * cx + 0.00583715f * cy));
    const float jf = (float)cvt_trunc((noise - 0.5f) * ((float)a.step * 0.8f)) * 0.5f;
    const Tap ts = make_tap(a.width, a.height, tc);
    Tap tg = ts;
    if (!same) tg = make_tap(a.g.w, a.g.h, tc);
    const float BaseDepth = sample_r16(a.g.t, tg);
    const int BaseNormal = normal_at(a.g, nearest_offset(a.g.w, a.g.h, tc), lut);
    float BaseSH[4], BaseCC[2], BaseAO[2];
    sample_rgba16(a.in.sh, ts, BaseSH);
    sample_rg16(a.in.cocg, ts, BaseCC);
    const float BaseLum = sh_to_y(BaseSH[3]);
    // GaussianVariance (:98-133): 3 x 3 taps at whole-texel offsets, set up per column and per row
    float BaseVariance = 0.0f, VarianceSum = 0.0f, TotalKernel = 0.0f;
    const f2 Texel = F2(1.0f / (float)a.width, 1.0f / (float)a.height);   // 1 / u_Dimensions == 1 / textureSize(u_SH, 0)
    {
        Axis gy[3];
        bool oky[3];
#pragma unroll
        for (int y = -1; y <= 1; ++y) {
            const float scy = tc.y + (float)y * Texel.y;
            oky[y + 1] = scy < 1.0f && scy > 0.0f;
            gy[y + 1] = make_axis(a.height, scy);
        }
#pragma unroll
        for (int x = -1; x <= 1; ++x) {
            const float scx = tc.x + (float)x * Texel.x;
            if (!(scx < 1.0f && scx > 0.0f)) continue;
            const Axis gx = make_axis(a.width, scx);
#pragma unroll
            for (int y = -1; y <= 1; ++y) {
                if (!oky[y + 1]) continue;
                const float kx = x == 0 ? 0.60283f : 0.198585f, ky = y == 0 ? 0.60283f : 0.198585f;
                const float KernelValue = kx * ky;
                const float V = sample_r16(a.in.x, join_axes(gx, gy[y + 1], a.width));
                if (x == 0 && y == 0) BaseVariance = V;
                VarianceSum += V * KernelValue;
                TotalKernel += KernelValue;
            }
        }
    }
    const float VarianceEstimate = VarianceSum / gmax(TotalKernel, 0.01f);
    sample_rg8(a.in.aosky, ts, lut, BaseAO);
    float oSH[4] = {BaseSH[0], BaseSH[1], BaseSH[2], BaseSH[3]}, oCC[2] = {BaseCC[0], BaseCC[1]}, oVar = BaseVariance, oAO[2] = {BaseAO[0], BaseAO[1]};
    if (a.do_spatial) {
        const bool FilterAO = a.step <= 4;
        float TotalSH[4] = {BaseSH[0], BaseSH[1], BaseSH[2], BaseSH[3]}, TotalCC[2] = {BaseCC[0], BaseCC[1]};
        float TotalWeight = 1.0f, TotalVariance = BaseVariance, TotalAO[2] = {BaseAO[0], BaseAO[1]}, TotalAOWeight = 1.0f;
        const bool Strong = sample_rgb16_ch(a.temporal_utility, ts, 0) <= 8.0f && a.aggressive && a.step <= 8;
        float CurveExponent = 0.0f;
        if (VarianceEstimate < 0.01f) CurveExponent = 128.0f;
        else if (VarianceEstimate < 0.025f) CurveExponent = 112.0f;
        else if (VarianceEstimate < 0.05f) CurveExponent = 96.0f;
        else if (VarianceEstimate < 0.075f) CurveExponent = 84.0f;
        else if (VarianceEstimate < 0.1f) CurveExponent = 70.0f;
        float Tweaked = VarianceEstimate;
        if (VarianceEstimate < 0.1f) {  // TweakVariance (:170-176)
            const float F = gclamp(VarianceEstimate, 0.0f, 1.0f);
            Tweaked = F * powf(1.0f - F, CurveExponent + 6.0f);
        }
        float PhiColor = sqrtf(gmax(0.0f, 0.000001f + Tweaked));
        PhiColor /= gmax(a.phi_bias, 0.1f);
        const int K = a.large_kernel ? 2 : 1;
        const float fstep = (float)a.step;
#pragma unroll 1
        for (int x = -K; x <= K; ++x) {
            const float scx = tc.x + (((float)x * fstep) * a.additional_scale + jf) * Texel.x;
            if (!(scx > 0.0f && scx < 1.0f)) continue;
            const Axis sx = make_axis(a.width, scx);
            Axis gx = sx;
            if (!same) gx = make_axis(a.g.w, scx);
            const int nx = wrap_near(cvt_floor(scx * (float)a.g.w), a.g.w);
#pragma unroll 1
            for (int y = -K; y <= K; ++y) {
                if (x == 0 && y == 0) continue;
                const float scy = tc.y + (((float)y * fstep) * a.additional_scale + jf) * Texel.y;
                if (!(scy > 0.0f && scy < 1.0f)) continue;
                const Axis sy = make_axis(a.height, scy);
                const Tap ss = join_axes(sx, sy, a.width);
                Tap sg = ss;
                if (!same) sg = join_axes(gx, make_axis(a.g.h, scy), a.g.w);
                const float SampleDepth = sample_r16(a.g.t, sg);
                const float DepthDiff = fabsf(SampleDepth - BaseDepth);
                // `BaseDepth < 0.0f == DepthDiff < 0.0f` parses as (BaseDepth < 0) == (DepthDiff < 0)
                if ((BaseDepth < 0.0f) != (DepthDiff < 0.0f)) continue;
                const int SampleNormal = normal_at(a.g, wrap_near(cvt_floor(scy * (float)a.g.h), a.g.h) * a.g.w + nx, lut);
                float sh[4], cc[2];
                sample_rgba16(a.in.sh, ss, sh);
                sample_rg16(a.in.cocg, ss, cc);
                const float SampleLum = sh_to_y(sh[3]);
                const float SampleVariance = sample_r16(a.in.x, ss);
                // pow(max(dot, 0), 32) in {0, 1, 3^32}, clamped to [0.001, 1]
                const float NormalWeight = normal_weight16(BaseNormal, SampleNormal) > 0.0f ? 1.0f : 0.001f;
                const float ed = expf(-gmax(DepthDiff, 0.00001f));
                const float DepthWeight = gclamp(ed * ed, 0.0001f, 1.0f);
                float Weight = NormalWeight * DepthWeight;
                if (!Strong) Weight = expf(-(fabsf(SampleLum - BaseLum) / PhiColor)) * NormalWeight * DepthWeight;
                Weight = gclamp(Weight, 0.001f, 1.0f);
                const int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
                const float XW = ax == 0 ? 1.0f : (ax == 1 ? 2.0f / 3.0f : 1.0f / 6.0f), YW = ay == 0 ? 1.0f : (ay == 1 ? 2.0f / 3.0f : 1.0f / 6.0f);
                Weight = (XW * YW) * Weight;
                Weight = gmax(Weight, 0.00000001f);
#pragma unroll
                for (int k = 0; k < 4; ++k) TotalSH[k] += sh[k] * Weight;
                TotalCC[0] += cc[0] * Weight; TotalCC[1] += cc[1] * Weight;
                TotalVariance += (Weight * Weight) * SampleVariance;
                TotalWeight += Weight;
                if (a.step <= 6) {  // FilterSky || FilterAO
                    const float AOW = gclamp((XW * YW) * NormalWeight * DepthWeight, 0.000001f, 1.0f);
                    float ao[2];
                    sample_rg8(a.in.aosky, ss, lut, ao);
                    TotalAO[0] += ao[0] * AOW; TotalAO[1] += ao[1] * AOW;
                    TotalAOWeight += AOW;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) oSH[k] = gclamp(TotalSH[k] / TotalWeight, -100.0f, 100.0f);
        oCC[0] = gclamp(TotalCC[0] / TotalWeight, -10.0f, 100.0f); oCC[1] = gclamp(TotalCC[1] / TotalWeight, -10.0f, 100.0f);
        oVar = gclamp(TotalVariance / (TotalWeight * TotalWeight), -1.0f, 50.0f);
        oAO[0] = FilterAO ? TotalAO[0] / TotalAOWeight : BaseAO[0];
        oAO[1] = TotalAO[1] / TotalAOWeight;
        oAO[0] = gclamp(oAO[0], 0.0f, 1.0f); oAO[1] = gclamp(oAO[1], 0.0f, 1.0f);
    }
    const size_t i = (size_t)py * a.width + px;
    ushort4 o4;
    o4.x = float_to_half_bits(oSH[0]); o4.y = float_to_half_bits(oSH[1]); o4.z = float_to_half_bits(oSH[2]); o4.w = float_to_half_bits(oSH[3]);
    reinterpret_cast<ushort4*>(a.out.sh)[i] = o4;
    ushort2 o2;
    o2.x = float_to_half_bits(oCC[0]); o2.y = float_to_half_bits(oCC[1]);
    reinterpret_cast<ushort2*>(a.out.cocg)[i] = o2;
    a.out.x[i] = float_to_half_bits(oVar);
    uchar2 ao;
    ao.x = float_to_unorm8(oAO[0]); ao.y = float_to_unorm8(oAO[1]);
    reinterpret_cast<uchar2*>(a.out.aosky)[i] = ao;
}

// ---------------------------------------------------------------------------------------------------------------
// Spatial3x3Initial.glsl main() (:103-175): the 3 x 3 pass in front of the temporal filter (Core/Pipeline.cpp:2381-2424).
// Eight neighbours, x-major; a tap counts when its world position is less than 1 from the centre's.  The normal term is subtracted
// in the exponent as the shader writes it (:137-139).  13 algorithmic bytes per pixel in (SH, CoCg, utility, AO/sky, hit distance,
// face) and 16 out.
struct PreSpatialArgs {
    float inv_view[16], inv_proj[16];
    int width, height, row0, row1;
    SetIn in;   // raw trace set: x = utility R16F
    GBufIn g;
    SetOut out;
};

__global__ void __launch_bounds__(256, VX_SVGF_OCC) svgf_prespatial_kernel(const __grid_constant__ PreSpatialArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const bool same = a.g.w == a.width && a.g.h == a.height;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const f3 origin = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const Tap ts = make_tap(a.width, a.height, tc);
    Tap tg = ts;
    if (!same) tg = make_tap(a.g.w, a.g.h, tc);
    const f3 BasePos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * sample_r16(a.g.t, tg);
    const int BaseNormal = normal_at(a.g, nearest_offset(a.g.w, a.g.h, tc), lut);
    float TotalSH[4], TotalCoCg[2], TotalAO[2];
    sample_rgba16(a.in.sh, ts, TotalSH);
    sample_rg16(a.in.cocg, ts, TotalCoCg);
    sample_rg8(a.in.aosky, ts, lut, TotalAO);
    const float BaseUtility = sample_r16(a.in.x, ts);
    const float BaseLuminance = gmax(0.0f, 3.544905f * TotalSH[3]);   // SHToY
    float TotalWeight = 1.0f, TotalAOWeight = 1.0f;
    const f2 Texel = F2(1.0f / (float)a.width, 1.0f / (float)a.height);
#pragma unroll 1
    for (int x = -1; x <= 1; ++x) {
        const float scx = tc.x + ((float)x * 1.0f) * Texel.x;
        if (!(scx > 0.0f && scx < 1.0f)) continue;
        const Axis sx = make_axis(a.width, scx);   // the column set-up is shared by the three taps of the column
        Axis gx = sx;
        if (!same) gx = make_axis(a.g.w, scx);
        const int nx = wrap_near(cvt_floor(scx * (float)a.g.w), a.g.w);
        const float wx = x == 0 ? 1.0f : 2.0f / 3.0f;   // AtrousWeights[abs(x)]
#pragma unroll 1
        for (int y = -1; y <= 1; ++y) {
            if (x == 0 && y == 0) continue;
            const f2 sc = F2(scx, tc.y + ((float)y * 1.0f) * Texel.y);
            if (!(sc.y > 0.0f && sc.y < 1.0f)) continue;
            const Tap ss = join_axes(sx, make_axis(a.height, sc.y), a.width);
            Tap sg = ss;
            if (!same) sg = join_axes(gx, make_axis(a.g.h, sc.y), a.g.w);
            const f3 SamplePos = origin + normalize(ray_direction_at(a.inv_view, a.inv_proj, sc)) * sample_r16(a.g.t, sg);
            const f3 e = F3(fabsf(SamplePos.x - BasePos.x), fabsf(SamplePos.y - BasePos.y), fabsf(SamplePos.z - BasePos.z));
            if (!(dot(e, e) < 1.0f)) continue;
            float s[4], c2[2], a2[2];
            sample_rgba16(a.in.sh, ss, s);
            sample_rg16(a.in.cocg, ss, c2);
            const float SampleLuma = gmax(0.0f, 3.544905f * s[3]);
            const float NormalWeight = normal_weight16(BaseNormal, normal_at(a.g, wrap_near(cvt_floor(sc.y * (float)a.g.h), a.g.h) * a.g.w + nx, lut));
            const float LuminosityWeight = fabsf(SampleLuma - BaseLuminance) / 4.0f;
            float Weight = expf(-LuminosityWeight - NormalWeight);
            Weight = gmax(Weight, 0.01f);
            const float wy = y == 0 ? 1.0f : 2.0f / 3.0f;
            Weight = (wx * wy) * Weight;
            Weight = gmax(Weight, 0.01f);
            Weight = gclamp(Weight, 0.0f, 1.0f);
#pragma unroll
            for (int j = 0; j < 4; ++j) TotalSH[j] += s[j] * Weight;
            TotalCoCg[0] += c2[0] * Weight; TotalCoCg[1] += c2[1] * Weight;
            TotalWeight += Weight;
            sample_rg8(a.in.aosky, ss, lut, a2);
            TotalAO[0] += a2[0] * Weight; TotalAO[1] += a2[1] * Weight;
            TotalAOWeight += Weight;
        }
    }
    TotalWeight = gmax(TotalWeight, 0.01f);
    const float aw = gmax(TotalAOWeight, 0.01f);
    const size_t i = (size_t)py * a.width + px;
    ushort4 o4;
    o4.x = float_to_half_bits(TotalSH[0] / TotalWeight); o4.y = float_to_half_bits(TotalSH[1] / TotalWeight);
    o4.z = float_to_half_bits(TotalSH[2] / TotalWeight); o4.w = float_to_half_bits(TotalSH[3] / TotalWeight);
    reinterpret_cast<ushort4*>(a.out.sh)[i] = o4;
    ushort2 o2;
    o2.x = float_to_half_bits(TotalCoCg[0] / TotalWeight); o2.y = float_to_half_bits(TotalCoCg[1] / TotalWeight);
    reinterpret_cast<ushort2*>(a.out.cocg)[i] = o2;
    a.out.x[i] = float_to_half_bits(BaseUtility);
    uchar2 ao;
    ao.x = float_to_unorm8(TotalAO[0] / aw); ao.y = float_to_unorm8(TotalAO[1] / aw);
    reinterpret_cast<uchar2*>(a.out.aosky)[i] = ao;
}

inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}

// bytes per pixel of image k (0..3) of a set: temporal sets carry the RGB16F utility, the others an R16F plane
inline int set_bpp(int k, bool temporal) { return k == 0 ? 8 : (k == 1 ? 4 : (k == 2 ? (temporal ? 6 : 2) : 2)); }
inline bool is_temporal_set(int id) { return id == VXRT_ATT_SVGF_TEMPORAL_A || id == VXRT_ATT_SVGF_TEMPORAL_B; }
inline bool is_svgf_set(int id) {
    return id == VXRT_ATT_SVGF_TEMPORAL_A || id == VXRT_ATT_SVGF_TEMPORAL_B || id == VXRT_ATT_SVGF_VARIANCE || id == VXRT_ATT_SVGF_DENOISE_A ||
           id == VXRT_ATT_SVGF_DENOISE_B;
}

int set_in(vxrt_ctx* c, const char* fn, int id, int x_bpp, bool need_ao, SetIn* s) {
    if (!(is_svgf_set(id) || id == VXRT_ATT_GI_SH || id == VXRT_ATT_SVGF_PRESPATIAL)) return vxrt_fail(VXRT_E_INVALID, "%s: %d does not name an image set", fn, id);
    const Attachment& a0 = c->att[id];
    if (!a0.ptr || a0.width <= 0) return vxrt_fail(VXRT_E_STATE, "%s: image set %d has not been written", fn, id);
    for (int k = 0; k < 4; ++k) {
        if (k == 3 && !need_ao) continue;
        const Attachment& a = c->att[id + k];
        const int want = k == 2 ? x_bpp : set_bpp(k, false);
        if (!a.ptr || a.width != a0.width || a.height != a0.height || a.bpp != want)
            return vxrt_fail(VXRT_E_STATE, "%s: image %d of set %d is missing or has the wrong format (%d bytes per pixel, expected %d)", fn, k, id, a.bpp, want);
    }
    s->sh = (const uint16_t*)c->att[id].ptr; s->cocg = (const uint16_t*)c->att[id + 1].ptr; s->x = (const uint16_t*)c->att[id + 2].ptr;
    s->aosky = need_ao ? (const uint8_t*)c->att[id + 3].ptr : nullptr;
    s->w = a0.width; s->h = a0.height;
    return VXRT_OK;
}

int set_out(vxrt_ctx* c, const char* fn, int id, int w, int h, bool with_ao, SetOut* s) {
    if (!is_svgf_set(id)) return vxrt_fail(VXRT_E_INVALID, "%s: %d does not name an SVGF image set", fn, id);
    int rc;
    for (int k = 0; k < (with_ao ? 4 : 3); ++k)
        if ((rc = vxrt_ensure_attachment(c, id + k, w, h, set_bpp(k, is_temporal_set(id))))) return rc;
    s->sh = (uint16_t*)c->att[id].ptr; s->cocg = (uint16_t*)c->att[id + 1].ptr; s->x = (uint16_t*)c->att[id + 2].ptr;
    s->aosky = with_ao ? (uint8_t*)c->att[id + 3].ptr : nullptr;
    return VXRT_OK;
}

int gbuf_in(vxrt_ctx* c, const char* fn, int t_id, int n_id, int b_id, GBufIn* g) {
    const Attachment& t = c->att[t_id];
    if (!t.ptr || t.width <= 0) return vxrt_fail(VXRT_E_STATE, "%s: G-buffer attachment %d has not been written", fn, t_id);
    const Attachment& n = c->att[n_id];
    const Attachment& b = c->att[b_id];
    if (!n.ptr || !b.ptr || n.width != t.width || n.height != t.height || b.width != t.width || b.height != t.height)
        return vxrt_fail(VXRT_E_STATE, "%s: G-buffer attachments %d / %d / %d do not form one frame", fn, t_id, n_id, b_id);
    g->t = (const uint16_t*)t.ptr; g->n = (const uint8_t*)n.ptr; g->b = (const uint8_t*)b.ptr; g->w = t.width; g->h = t.height;
    return VXRT_OK;
}

}  // namespace

int vxrt_launch_svgf_temporal(vxrt_ctx* c, const vxrt_svgf_temporal_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_svgf_temporal";
    if (p.out_set == p.history_set || p.out_set == p.in_set) return vxrt_fail(VXRT_E_INVALID, "%s: out_set aliases an input set", fn);
    if (!is_temporal_set(p.out_set) || !is_temporal_set(p.history_set)) return vxrt_fail(VXRT_E_INVALID, "%s: history_set / out_set must be VXRT_ATT_SVGF_TEMPORAL_A / _B", fn);
    TemporalArgs a;
    int rc;
    if ((rc = set_in(c, fn, p.in_set, 2, true, &a.cur))) return rc;
    if (a.cur.w != p.width || a.cur.h != p.height) return vxrt_fail(VXRT_E_STATE, "%s: image set is %dx%d, the pass runs at %dx%d", fn, a.cur.w, a.cur.h, p.width, p.height);
    if ((rc = gbuf_in(c, fn, VXRT_ATT_INITIAL_T, VXRT_ATT_INITIAL_NORMAL, VXRT_ATT_INITIAL_BLOCK, &a.g))) return rc;
    // first frame: no history yet.  The engine's FBOs start out zero-filled, so do these.
    if (!c->att[p.history_set].ptr || c->att[p.history_set].width != p.width || c->att[p.history_set].height != p.height) {
        SetOut z;
        if ((rc = set_out(c, fn, p.history_set, p.width, p.height, true, &z))) return rc;
        for (int k = 0; k < 4; ++k) VX_CUDA(cudaMemsetAsync(c->att[p.history_set + k].ptr, 0, (size_t)p.width * p.height * c->att[p.history_set + k].bpp, c->stream));
    }
    if (!c->att[VXRT_ATT_PREV_INITIAL_T].ptr || c->att[VXRT_ATT_PREV_INITIAL_T].width != a.g.w || c->att[VXRT_ATT_PREV_INITIAL_T].height != a.g.h) {
        const int ids[3] = {VXRT_ATT_PREV_INITIAL_T, VXRT_ATT_PREV_INITIAL_NORMAL, VXRT_ATT_PREV_INITIAL_BLOCK}, bpp[3] = {2, 1, 1};
        for (int k = 0; k < 3; ++k) {
            if ((rc = vxrt_ensure_attachment(c, ids[k], a.g.w, a.g.h, bpp[k]))) return rc;
            VX_CUDA(cudaMemsetAsync(c->att[ids[k]].ptr, 0, (size_t)a.g.w * a.g.h * bpp[k], c->stream));
        }
    }
    if ((rc = set_in(c, fn, p.history_set, 6, true, &a.hist))) return rc;
    if (a.hist.w != p.width || a.hist.h != p.height) return vxrt_fail(VXRT_E_STATE, "%s: image set is %dx%d, the pass runs at %dx%d", fn, a.hist.w, a.hist.h, p.width, p.height);
    if ((rc = gbuf_in(c, fn, VXRT_ATT_PREV_INITIAL_T, VXRT_ATT_PREV_INITIAL_NORMAL, VXRT_ATT_PREV_INITIAL_BLOCK, &a.pg))) return rc;
    if ((rc = set_out(c, fn, p.out_set, p.width, p.height, true, &a.out))) return rc;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    // u_PrevProjection * u_PrevView: column j = P * (column j of V), with the mat4 * vec4 association of vmath.cuh
    for (int j = 0; j < 4; ++j) {
        const float* v = p.prev_view + 4 * j;
        const float* m = p.prev_projection;
        for (int r = 0; r < 4; ++r) a.prev_pv[4 * j + r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
    }
    a.width = p.width; a.height = p.height; a.be_useful = p.be_useful;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    svgf_temporal_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_svgf_prespatial(vxrt_ctx* c, const vxrt_svgf_prespatial_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_svgf_prespatial";
    if (p.in_set != VXRT_ATT_GI_SH) return vxrt_fail(VXRT_E_INVALID, "%s: in_set must be VXRT_ATT_GI_SH (the raw trace)", fn);
    PreSpatialArgs a;
    int rc;
    if ((rc = set_in(c, fn, p.in_set, 2, true, &a.in))) return rc;
    if (a.in.w != p.width || a.in.h != p.height) return vxrt_fail(VXRT_E_STATE, "%s: image set is %dx%d, the pass runs at %dx%d", fn, a.in.w, a.in.h, p.width, p.height);
    if ((rc = gbuf_in(c, fn, VXRT_ATT_INITIAL_T, VXRT_ATT_INITIAL_NORMAL, VXRT_ATT_INITIAL_BLOCK, &a.g))) return rc;
    for (int k = 0; k < 4; ++k)   // DiffusePreTemporal_SpatialFBO (Pipeline.cpp:1153): RGBA16F, RG16F, R16F, RG8
        if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_SVGF_PRESPATIAL + k, p.width, p.height, set_bpp(k, false)))) return rc;
    a.out.sh = (uint16_t*)c->att[VXRT_ATT_SVGF_PRESPATIAL].ptr; a.out.cocg = (uint16_t*)c->att[VXRT_ATT_SVGF_PRESPATIAL + 1].ptr;
    a.out.x = (uint16_t*)c->att[VXRT_ATT_SVGF_PRESPATIAL + 2].ptr; a.out.aosky = (uint8_t*)c->att[VXRT_ATT_SVGF_PRESPATIAL + 3].ptr;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    svgf_prespatial_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_svgf_variance(vxrt_ctx* c, const vxrt_svgf_variance_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_svgf_variance";
    VarianceArgs a;
    int rc;
    if (!is_temporal_set(p.in_set)) return vxrt_fail(VXRT_E_INVALID, "%s: in_set must be a temporal set", fn);
    if ((rc = set_in(c, fn, p.in_set, 6, false, &a.in))) return rc;
    if (a.in.w != p.width || a.in.h != p.height) return vxrt_fail(VXRT_E_STATE, "%s: image set is %dx%d, the pass runs at %dx%d", fn, a.in.w, a.in.h, p.width, p.height);
    if ((rc = gbuf_in(c, fn, VXRT_ATT_INITIAL_T, VXRT_ATT_INITIAL_NORMAL, VXRT_ATT_INITIAL_BLOCK, &a.g))) return rc;
    if ((rc = set_out(c, fn, VXRT_ATT_SVGF_VARIANCE, p.width, p.height, false, &a.out))) return rc;
    a.width = p.width; a.height = p.height; a.do_spatial = p.do_spatial; a.aggressive = p.aggressive_disocclusion;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    svgf_variance_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_svgf_spatial(vxrt_ctx* c, const vxrt_svgf_spatial_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_svgf_spatial";
    if (p.out_set == p.in_set || p.out_set == p.ao_set || p.out_set == p.temporal_set) return vxrt_fail(VXRT_E_INVALID, "%s: out_set aliases an input set", fn);
    if (p.out_set != VXRT_ATT_SVGF_DENOISE_A && p.out_set != VXRT_ATT_SVGF_DENOISE_B) return vxrt_fail(VXRT_E_INVALID, "%s: out_set must be VXRT_ATT_SVGF_DENOISE_A / _B", fn);
    if (!is_temporal_set(p.temporal_set)) return vxrt_fail(VXRT_E_INVALID, "%s: temporal_set must be a temporal set", fn);
    if (p.step < 1) return vxrt_fail(VXRT_E_INVALID, "%s: step %d", fn, p.step);
    SpatialArgs a;
    SetIn ao, tmp;
    int rc;
    if ((rc = set_in(c, fn, p.in_set, 2, false, &a.in))) return rc;
    if (a.in.w != p.width || a.in.h != p.height) return vxrt_fail(VXRT_E_STATE, "%s: image set is %dx%d, the pass runs at %dx%d", fn, a.in.w, a.in.h, p.width, p.height);
    if ((rc = set_in(c, fn, p.ao_set, is_temporal_set(p.ao_set) ? 6 : 2, true, &ao))) return rc;
    if ((rc = set_in(c, fn, p.temporal_set, 6, false, &tmp))) return rc;
    if (ao.w != a.in.w || ao.h != a.in.h) return vxrt_fail(VXRT_E_STATE, "%s: ao_set and in_set differ in size", fn);
    a.in.aosky = ao.aosky;
    if (tmp.w != a.in.w || tmp.h != a.in.h) return vxrt_fail(VXRT_E_STATE, "%s: temporal_set and in_set differ in size", fn);
    a.temporal_utility = tmp.x;
    if ((rc = gbuf_in(c, fn, VXRT_ATT_INITIAL_T, VXRT_ATT_INITIAL_NORMAL, VXRT_ATT_INITIAL_BLOCK, &a.g))) return rc;
    if ((rc = set_out(c, fn, p.out_set, p.width, p.height, true, &a.out))) return rc;
    a.width = p.width; a.height = p.height;
    a.step = p.step; a.large_kernel = p.large_kernel; a.do_spatial = p.do_spatial; a.aggressive = p.aggressive_disocclusion;
    a.phi_bias = p.color_phi_bias;
    const float tm = p.time * 100.493850275f;
    a.time_offset = tm - 500.0f * floorf(tm / 500.0f);  // mod(u_Time * 100.493850275, 500)
    a.additional_scale = 1.0f * (1.0f - p.resolution_scale) + 2.4f * p.resolution_scale;  // mix(1, 2.4, u_ResolutionScale)
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    svgf_spatial_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

// end of frame: this frame's primary G-buffer becomes the previous one (the engine swaps InitialTraceFBO_1 / _2,
// Core/Pipeline.cpp:2046-2048).  4 bytes per pixel are copied on the stream (pointers handed out by
// vxrt_cuda_attachment_device stay valid, which a swap would break).
int vxrt_launch_svgf_end_frame(vxrt_ctx* c) {
    const int cur[3] = {VXRT_ATT_INITIAL_T, VXRT_ATT_INITIAL_NORMAL, VXRT_ATT_INITIAL_BLOCK};
    const int prev[3] = {VXRT_ATT_PREV_INITIAL_T, VXRT_ATT_PREV_INITIAL_NORMAL, VXRT_ATT_PREV_INITIAL_BLOCK};
    for (int k = 0; k < 3; ++k)
        if (!c->att[cur[k]].ptr || c->att[cur[k]].width <= 0) return vxrt_fail(VXRT_E_STATE, "vxrt_cuda_svgf_end_frame: no primary G-buffer (vxrt_cuda_initial_trace)");
    for (int k = 0; k < 3; ++k) {
        Attachment& a = c->att[cur[k]];
        Attachment& b = c->att[prev[k]];
        int rc;
        if ((rc = vxrt_ensure_attachment(c, prev[k], a.width, a.height, a.bpp))) return rc;
        VX_CUDA(cudaMemcpyAsync(b.ptr, a.ptr, (size_t)a.width * a.height * a.bpp, cudaMemcpyDeviceToDevice, c->stream));
    }
    // the reflection trace's hit distance of this frame is what the reflection temporal filter reprojects into next frame
    // (PrevReflectionTraceFBO.GetTexture(1), Core/Pipeline.cpp:1864-1865, 3381)
    const Attachment& h = c->att[VXRT_ATT_REFL_HITDIST];
    if (h.ptr && h.width > 0) {
        int rc;
        if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_PREV_REFL_HITDIST, h.width, h.height, h.bpp))) return rc;
        VX_CUDA(cudaMemcpyAsync(c->att[VXRT_ATT_PREV_REFL_HITDIST].ptr, h.ptr, (size_t)h.width * h.height * h.bpp, cudaMemcpyDeviceToDevice, c->stream));
    }
    return VXRT_OK;
}
