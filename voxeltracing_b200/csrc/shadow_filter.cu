// shadow_filter.cu — sun-shadow denoiser: Core/Shaders/ShadowTemporalFilter.glsl (reprojection, 3 x 3
// pre-filter of the binary trace, history clipping, frame counter) and ShadowFilter.glsl (bilateral 3 x 3 / 7 x 7 filter
// steered by the occluder distance), dispatched at Core/Pipeline.cpp:2947-3044.  One thread per pixel, a warp covers an
// 8 x 4 pixel tile; samplers of filter_sampler.cuh.  The raw trace, the temporal images and the G-buffer may each have
// their own size (ShadowTraceResolution / ShadowSupersampleRes / full resolution).  Both passes weight with expf / powf:
// R8 outputs agree with the oracle to one code (tests/test_gpu_shadow_filter.py).
#include "ctx.h"
#include "filter_sampler.cuh"

namespace {

struct Img8 { const uint8_t* __restrict__ p; int w, h; };
struct Img16 { const uint16_t* __restrict__ p; int w, h; };

struct ShadowTemporalArgs {
    float inv_view[16], inv_proj[16], prev_pv[16];
    int width, height, row0, row1, shadow_temporal;
    Img8 raw;            // ShadowRawTrace[0]
    Img16 transversal;   // ShadowRawTrace[1], same size
    Img8 hist;           // previous temporal shadow, width x height
    Img16 hist_frames;
    Img16 g_t, prev_t;   // hit distance of this / the previous frame
    Img8 g_n;
    uint8_t* __restrict__ out;
    uint16_t* __restrict__ out_frames;
};

// pow(x, n) for the small integer exponents the shaders use, by multiplication (<= 3 ulp for n = 48, below the error of the
// general powf; the reference's own pow() is implementation defined)
VXD float pow3(float x) { return (x * x) * x; }
VXD float pow7(float x) { const float x2 = x * x, x4 = x2 * x2; return (x4 * x2) * x; }
VXD float pow48(float x) { const float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8; return (x16 * x16) * x16; }

VXD f3 position_at(const float* inv_view, const float* inv_proj, f3 origin, f2 uv, float dist) {
    return origin + normalize(ray_direction_at(inv_view, inv_proj, uv)) * dist;
}

// ShadowTemporalFilter.glsl main() (:193-267)
__global__ void __launch_bounds__(256) shadow_temporal_kernel(const __grid_constant__ ShadowTemporalArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const f3 origin = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const float Dist = sample_r16(a.g_t.p, make_tap(a.g_t.w, a.g_t.h, tc));
    const Tap tr = make_tap(a.raw.w, a.raw.h, tc);
    float oColor, oFrames = 0.0f;
    if (Dist > 0.0f) {
        const f3 CurPos = position_at(a.inv_view, a.inv_proj, origin, tc, Dist);
        const f4 Proj = mat4_mul(a.prev_pv, F4(CurPos.x, CurPos.y, CurPos.z, 1.0f));
        const f2 R = F2((Proj.x / Proj.w) * 0.5f + 0.5f, (Proj.y / Proj.w) * 0.5f + 0.5f);
        const float TransversalAt = sample_r16(a.transversal.p, tr) * 100.0f;
        const f2 Texel = F2(1.0f / (float)a.raw.w, 1.0f / (float)a.raw.h);
        const bool ST = a.shadow_temporal != 0;
        float CurrentColor;
        if (!ST) CurrentColor = sample_r8(a.raw.p, tr, lut);
        else if (TransversalAt <= 1.41421354f * 2.0f) CurrentColor = 1.0f;
        else {  // GetShadowSpatial (:107-151)
            float Total = sample_r8(a.raw.p, tr, lut);
            const float Base = Total;
            float Weight = 1.0f;
            const int BaseNormal = normal_index(lut[__ldg(a.g_n.p + nearest_offset(a.g_n.w, a.g_n.h, tc))]);
#pragma unroll 1
            for (int x = -1; x <= 1; ++x)
#pragma unroll 1
                for (int y = -1; y <= 1; ++y) {
                    if (x == 0 && y == 0) continue;
                    const f2 sc = F2(tc.x + (float)x * Texel.x, tc.y + (float)y * Texel.y);
                    const float b = 0.03f;
                    if (!(sc.x > b && sc.x < 1.0f - b && sc.y > b && sc.y < 1.0f - b)) continue;
                    const float SampleDepth = sample_r16(a.g_t.p, make_tap(a.g_t.w, a.g_t.h, sc));
                    const int SampleNormal = normal_index(lut[__ldg(a.g_n.p + nearest_offset(a.g_n.w, a.g_n.h, sc))]);
                    if (SampleNormal == BaseNormal && fabsf(SampleDepth - Dist) < 1.0f) {
                        const float Sample = sample_r8(a.raw.p, make_tap(a.raw.w, a.raw.h, sc), lut);
                        float WeightAt = gclamp(1.0f - gclamp(fabsf(Sample - Base) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                        WeightAt = gclamp(pow7(WeightAt), 0.000001f, 1.0f);
                        Total += Sample * WeightAt;
                        Weight += WeightAt;
                    }
                }
            CurrentColor = Total / Weight;
        }
        const Tap th = make_tap(a.hist.w, a.hist.h, R);
        const float PrevColorOrig = sample_r8(a.hist.p, th, lut);
        float PrevColor = PrevColorOrig;
        if (ST && TransversalAt < 1.414f * 3.0f) {  // ClipShadow / clipAABB (:153-191), all components equal
            float MinColor = 100.0f, MaxColor = -100.0f;
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const float ox = s == 0 ? -1.0f : (s == 1 ? 1.0f : 0.0f), oy = s == 3 ? -1.0f : (s == 4 ? 1.0f : 0.0f);
                const float Sample = sample_r8(a.raw.p, make_tap(a.raw.w, a.raw.h, F2(tc.x + ox * Texel.x, tc.y + oy * Texel.y)), lut);
                MinColor = gmin(Sample, MinColor);
                MaxColor = gmax(Sample, MaxColor);
            }
            const float mn = MinColor - 0.125f, mx = MaxColor + 0.125f;
            const float pClip = 0.5f * (mx + mn), eClip = 0.5f * (mx - mn), vClip = PrevColorOrig - pClip;
            const float denom = fabsf(vClip / eClip);
            PrevColor = denom > 1.0f ? pClip + vClip / denom : PrevColorOrig;
        }
        const float PrevDist = sample_r16(a.prev_t.p, make_tap(a.prev_t.w, a.prev_t.h, R));
        const f3 PrevPos = position_at(a.inv_view, a.inv_proj, origin, R, PrevDist);
        const float Bias = ST ? 0.005f : 0.01f;
        const bool Rejected = !(R.x > 0.0f + Bias && R.x < 1.0f - Bias && R.y > 0.0f + Bias && R.y < 1.0f - Bias);
        if (!Rejected) {
            const float d = distance(PrevPos, CurPos);
            CurrentColor = gclamp(CurrentColor, 0.0f, 1.0f);
            PrevColor = gclamp(PrevColor, 0.0f, 1.0f);
            const float vx = (tc.x - R.x) * (float)a.raw.w, vy = (tc.y - R.y) * (float)a.raw.h;
            const float ClipError = fabsf(PrevColorOrig - PrevColor);
            const float FrameIncrement = ClipError < 0.2f ? 1.0f : 0.6f;
            const float FrameCountFetch = sample_r16(a.hist_frames.p, th);
            const float FrameIncremented = FrameCountFetch + FrameIncrement;
            float BlendFactor = gclamp((1.0f - (1.0f / FrameIncremented)) * 1.2f, 0.01f, 0.97f);
            const float VRF = gclamp(expf(-sqrtf(vx * vx + vy * vy)) * 0.8f + 0.6f, 0.00000001f, 1.0f);
            BlendFactor *= VRF;
            float DepthRejection = 1.0f;
            if (d > 0.4f) {
                DepthRejection = pow48(expf(-d));
                BlendFactor *= gclamp(DepthRejection, 0.0f, 1.0f);
            }
            oColor = gmix(CurrentColor, PrevColor, gclamp(BlendFactor, 0.0f, 0.97f));
            const float BFM = DepthRejection * VRF;
            oFrames = FrameCountFetch + gclamp(BFM * 1.1f, 0.0f, 1.0f);
            if (BFM < 0.1f) oFrames = 0.0f;
            else if (BFM <= 0.2f + 0.001f) oFrames = 2.0f;
            else if (BFM <= 0.3f + 0.001f) oFrames = 3.25f;
        } else {
            oColor = CurrentColor;
        }
    } else {
        oColor = sample_r8(a.raw.p, tr, lut);
    }
    oFrames = gclamp(oFrames, 0.0f, 256.0f);
    const size_t i = (size_t)py * a.width + px;
    a.out[i] = float_to_unorm8(oColor);
    a.out_frames[i] = float_to_half_bits(oFrames);
}

struct ShadowFilterArgs {
    int width, height, row0, row1;
    float filter_scale;
    Img8 in;             // temporal shadow
    Img16 in_frames;     // same size
    Img16 transversal;   // ShadowRawTrace[1]
    Img16 g_t;
    Img8 g_n;
    uint8_t* __restrict__ out;
};

// ShadowFilter.glsl ShadowSpatial (:68-159); taps are not tested against the screen, REPEAT wraps them
__global__ void __launch_bounds__(256) shadow_filter_kernel(const __grid_constant__ ShadowFilterArgs a) {
    __shared__ float lut[256];
    fill_unorm_lut(lut);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int py = a.row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= a.width || py >= a.row1) return;
    const f2 tc = F2(((float)px + 0.5f) / (float)a.width, ((float)py + 0.5f) / (float)a.height);
    const Tap ti = make_tap(a.in.w, a.in.h, tc);
    const float Frames = sample_r16(a.in_frames.p, ti);
    const float CenterDist = sample_r16(a.g_t.p, make_tap(a.g_t.w, a.g_t.h, tc));
    const int CenterNormal = normal_index(lut[__ldg(a.g_n.p + nearest_offset(a.g_n.w, a.g_n.h, tc))]);
    const float CenterShadow = sample_r8(a.in.p, ti, lut);
    const float Transversal = sample_r16(a.transversal.p, make_tap(a.transversal.w, a.transversal.h, tc)) * 100.0f;
    const float Cutoff = 1.41421354f;  // sqrt(2.0f)
    float result = CenterShadow;
    if (!((Transversal > 0.0f && Transversal < Cutoff) || CenterDist < 0.0f)) {
        const int K = Transversal < Cutoff * 1.414f ? 1 : 3;
        float Scale = 1.0f;
        if (Transversal > 6.0f) Scale = 2.0f;
        if (Transversal > 16.0f) Scale = 2.4f;
        if (Transversal > 32.0f) Scale = 2.6f;
        const float ClampedT = gclamp(Transversal, 0.0f, 10.0f);
        float VarianceEstimate = gmix(20.0f, 6.0f, ClampedT / 10.0f) + (Transversal < 6.0f ? 5.0f : 2.0f);
        VarianceEstimate = gclamp(VarianceEstimate - 1.75f, 0.0000001f, 64.0f);
        float LumaMixer = 1.0f;
        if (!(Frames > 7.5f)) LumaMixer = gmix(0.1f, 0.5f, Frames / 7.5f);
        const float LumaExponent = VarianceEstimate * LumaMixer * 0.9f;
        const f2 Texel = F2(1.0f / (float)a.in.w, 1.0f / (float)a.in.h);
        const bool same = a.g_t.w == a.in.w && a.g_t.h == a.in.h;
        float TotalWeight = 0.0f, TotalShadow = 0.0f;
#pragma unroll 1
        for (int x = -K; x <= K; ++x) {
            const float scx = tc.x + ((((float)x * Texel.x) * 1.2f) * Scale) * a.filter_scale;
            const Axis ix = make_axis(a.in.w, scx);
            Axis gx = ix;
            if (!same) gx = make_axis(a.g_t.w, scx);
            const int nx = wrap_near(cvt_floor(scx * (float)a.g_n.w), a.g_n.w);
#pragma unroll 1
            for (int y = -K; y <= K; ++y) {
                const float scy = tc.y + ((((float)y * Texel.y) * 1.2f) * Scale) * a.filter_scale;
                const Axis iy = make_axis(a.in.h, scy);
                const Tap si = join_axes(ix, iy, a.in.w);
                Tap sg = si;
                if (!same) sg = join_axes(gx, make_axis(a.g_t.h, scy), a.g_t.w);
                const float SampleDepth = sample_r16(a.g_t.p, sg);
                const int SampleNormal = normal_index(lut[__ldg(a.g_n.p + wrap_near(cvt_floor(scy * (float)a.g_n.h), a.g_n.h) * a.g_n.w + nx)]);
                const float ed = expf(-(fabsf(CenterDist - SampleDepth)));
                const float DepthWeight = pow3(ed);
                // pow(max(dot, 1e-9), 32): 1e-288 underflows to 0, 1, or powf(3, 32)
                const float nd = normal_dot(CenterNormal, SampleNormal);
                const float NormalWeight = nd <= 0.0f ? 0.0f : (nd == 1.0f ? 1.0f : 1853020153315328.0f);
                const float ShadowAt = sample_r8(a.in.p, si, lut);
                const float LuminanceError = gclamp(1.0f - gclamp(fabsf(ShadowAt - CenterShadow) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                float Weight = 1.0f;
                Weight *= LuminanceError == 1.0f ? 1.0f : gclamp(powf(LuminanceError, LumaExponent), 0.0f, 1.0f);   // pow(1, y) == 1 exactly
                Weight *= DepthWeight;
                Weight *= NormalWeight;
                Weight = gclamp(Weight, 0.000000001f, 1.0f);
                TotalShadow += ShadowAt * Weight;
                TotalWeight += Weight;
            }
        }
        result = TotalShadow / gmax(TotalWeight, 0.01f);
    }
    a.out[(size_t)py * a.width + px] = float_to_unorm8(result);
}

inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}
inline bool is_shadow_set(int id) { return id == VXRT_ATT_SHADOW_TEMPORAL_A || id == VXRT_ATT_SHADOW_TEMPORAL_B; }

template <typename I>
int image_in(vxrt_ctx* c, const char* fn, int id, int bpp, I* img) {
    const Attachment& a = c->att[id];
    if (!a.ptr || a.width <= 0) return vxrt_fail(VXRT_E_STATE, "%s: attachment %d has not been written", fn, id);
    if (a.bpp != bpp) return vxrt_fail(VXRT_E_STATE, "%s: attachment %d has %d bytes per pixel, expected %d", fn, id, a.bpp, bpp);
    img->p = (decltype(img->p))a.ptr; img->w = a.width; img->h = a.height;
    return VXRT_OK;
}

}  // namespace

int vxrt_launch_shadow_temporal(vxrt_ctx* c, const vxrt_shadow_temporal_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_shadow_temporal";
    if (!is_shadow_set(p.history_set) || !is_shadow_set(p.out_set) || p.history_set == p.out_set)
        return vxrt_fail(VXRT_E_INVALID, "%s: history_set / out_set must be the two of VXRT_ATT_SHADOW_TEMPORAL_A / _B", fn);
    ShadowTemporalArgs a;
    int rc;
    if ((rc = image_in(c, fn, VXRT_ATT_SHADOW, 1, &a.raw))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_SHADOW_TRANSVERSAL, 2, &a.transversal))) return rc;
    if (a.transversal.w != a.raw.w || a.transversal.h != a.raw.h) return vxrt_fail(VXRT_E_STATE, "%s: SHADOW and SHADOW_TRANSVERSAL differ in size", fn);
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_T, 2, &a.g_t))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_NORMAL, 1, &a.g_n))) return rc;
    // first frame: no history and no previous G-buffer yet.  The engine's FBOs start out zero-filled, so do these.
    const Attachment& h0 = c->att[p.history_set];
    if (!h0.ptr || h0.width != p.width || h0.height != p.height) {
        const int bpp[2] = {1, 2};
        for (int k = 0; k < 2; ++k) {
            if ((rc = vxrt_ensure_attachment(c, p.history_set + k, p.width, p.height, bpp[k]))) return rc;
            VX_CUDA(cudaMemsetAsync(c->att[p.history_set + k].ptr, 0, (size_t)p.width * p.height * bpp[k], c->stream));
        }
    }
    const Attachment& pt = c->att[VXRT_ATT_PREV_INITIAL_T];
    if (!pt.ptr || pt.width != a.g_t.w || pt.height != a.g_t.h) {
        if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_PREV_INITIAL_T, a.g_t.w, a.g_t.h, 2))) return rc;
        VX_CUDA(cudaMemsetAsync(c->att[VXRT_ATT_PREV_INITIAL_T].ptr, 0, (size_t)a.g_t.w * a.g_t.h * 2, c->stream));
    }
    if ((rc = image_in(c, fn, p.history_set, 1, &a.hist))) return rc;
    if ((rc = image_in(c, fn, p.history_set + 1, 2, &a.hist_frames))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_PREV_INITIAL_T, 2, &a.prev_t))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_set, p.width, p.height, 1))) return rc;
    if ((rc = vxrt_ensure_attachment(c, p.out_set + 1, p.width, p.height, 2))) return rc;
    a.out = (uint8_t*)c->att[p.out_set].ptr; a.out_frames = (uint16_t*)c->att[p.out_set + 1].ptr;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    for (int j = 0; j < 4; ++j) {   // u_PrevProjection * u_PrevView, column by column (mat4 * vec4 association of vmath.cuh)
        const float* v = p.prev_view + 4 * j;
        const float* m = p.prev_projection;
        for (int r = 0; r < 4; ++r) a.prev_pv[4 * j + r] = (m[r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
    }
    a.width = p.width; a.height = p.height; a.shadow_temporal = p.shadow_temporal;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    shadow_temporal_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_shadow_filter(vxrt_ctx* c, const vxrt_shadow_filter_params& p) {
    { const int rc_snap = vx_apply_filter_snap(c); if (rc_snap != VXRT_OK) return rc_snap; }
    static const char* fn = "vxrt_cuda_shadow_filter";
    if (!is_shadow_set(p.in_set)) return vxrt_fail(VXRT_E_INVALID, "%s: in_set must be VXRT_ATT_SHADOW_TEMPORAL_A / _B", fn);
    ShadowFilterArgs a;
    int rc;
    if ((rc = image_in(c, fn, p.in_set, 1, &a.in))) return rc;
    if ((rc = image_in(c, fn, p.in_set + 1, 2, &a.in_frames))) return rc;
    if (a.in_frames.w != a.in.w || a.in_frames.h != a.in.h) return vxrt_fail(VXRT_E_STATE, "%s: the temporal set's images differ in size", fn);
    if ((rc = image_in(c, fn, VXRT_ATT_SHADOW_TRANSVERSAL, 2, &a.transversal))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_T, 2, &a.g_t))) return rc;
    if ((rc = image_in(c, fn, VXRT_ATT_INITIAL_NORMAL, 1, &a.g_n))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_SHADOW_FILTERED, p.width, p.height, 1))) return rc;
    a.out = (uint8_t*)c->att[VXRT_ATT_SHADOW_FILTERED].ptr;
    a.width = p.width; a.height = p.height; a.filter_scale = p.filter_scale;
    tile_rows(p.tile, p.height, &a.row0, &a.row1);
    if (a.row1 <= a.row0) return VXRT_OK;
    dim3 grid((p.width + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    shadow_filter_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
