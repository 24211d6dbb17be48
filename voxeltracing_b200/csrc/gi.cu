// gi.cu — diffuse GI pass: DiffuseRayTraceFrag.glsl (main :910-1021, CalculateDiffuse :535-664),
// dispatched at Core/Pipeline.cpp:2267-2374; attachments Core/Pipeline.cpp:1150.
//
// One thread per GI pixel walks SPP cosine-weighted 2-bounce paths (<= trace_length iterations per
// bounce ray, <= shadow_trace_length per sun-shadow ray), shading each hit from the albedo / PBR /
// emissive arrays and the BlockData table, and writes SH(L1) + CoCg + luminance + AO/sky in the
// attachment formats.  Per-pass constants that need libm transcendentals (LIGHT_COLOR) come from the host.
#include "gi_common.cuh"

namespace {

// CalculateDiffuse (:535-664)
template <bool STATS>
VXD f4 calculate_diffuse(const GridView& g, const GiArgs& a, GiState& st, f3 initial_origin, f3 input_normal, f3& odir, bool& Skyhit,
                         LaneStats* ls) {
    Skyhit = false;
    const float bias = 0.06f;
    const f3 light = a.sun_stronger ? F3(a.sun[0], a.sun[1], a.sun[2]) : F3(a.moon[0], a.moon[1], a.moon[2]);
    const f3 LIGHT_COLOR = F3(a.light_color[0], a.light_color[1], a.light_color[2]);
    f3 rayO = initial_origin + input_normal * bias;
    f3 rayD = cos_weighted_hemisphere(a, st, input_normal);
    float ao = 1.0f;
    f3 RayContribution = F3(0.0f), RayThroughput = F3(1.0f);
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
        if (i == 0) odir = rayD;
        TraceResult h = traverse_df<STATS>(g, rayO, rayD, a.trace_length, ls);
        const float T = h.t;
        const f3 HitNormal = h.normal;
        const int tex_ref = iclamp(h.block, 0, 127);
        const f3 IntersectionPosition = rayO + (rayD * T);
        if (T > 0.0f && h.block > 0) {
            f2 txc = F2(0.0f, 0.0f);
            calculate_uv(IntersectionPosition, HitNormal, txc);
            const float TexA = (float)__ldg(a.block_data + tex_ref), TexE = (float)__ldg(a.block_data + 384 + tex_ref);
            const f3 Albedo = xyz(texarray_sample(a.tex[VXRT_TEX_ALBEDO], txc.x, txc.y, TexA, 3.0f));
            const f3 PBR = xyz(texarray_sample(a.tex[VXRT_TEX_PBR], txc.x, txc.y, TexA, 2.0f));  // sic: albedo layer id (:578)
            float Emmisivity = 0.0f;
            if (TexE >= 0.0f) {
                float SampledEmmisivity = texarray_sample(a.tex[VXRT_TEX_EMISSIVE], txc.x, txc.y, TexE, 0.0f).x;
                Emmisivity = SampledEmmisivity * 12.0f * a.diffuse_light_intensity;
            }
            const float NDotL = gmax(dot(HitNormal, light), 0.0f);
            float ShadowAt;
            if (!a.sun_stronger) {
                ShadowAt = 1.0f;
            } else if (NDotL < 0.001f) {
                ShadowAt = 0.0f;
            } else {
                // GetShadowAt (:1290-1308)
                const f3 spos = IntersectionPosition + HitNormal * 0.045f;
                bool player = false;
                if (a.apply_player_shadow) {
                    const f3 vp = F3(a.viewer[0], a.viewer[1], a.viewer[2]);
                    player = ray_box_intersect(vp + F3(0.2f, 0.0f, 0.2f), vp - F3(0.75f, 1.75f, 0.75f), spos, light);
                }
                if (player) ShadowAt = 1.0f;
                else {
                    TraceResult sh = traverse_df<STATS>(g, spos, light, a.shadow_trace_length, ls);
                    ShadowAt = sh.t > 0.0f ? 1.0f : 0.0f;
                }
            }
            const f3 EmmisivityColor = (Emmisivity * gmix(1.0f, 1.0f, a.sun_visibility)) * Albedo;
            const f3 SUNBRDF = Albedo * diffuse_hammon(HitNormal, -rayD, light, PBR.x) * (LIGHT_COLOR * 3.5f) * (1.0f - ShadowAt) * VX_PI;
            const f3 NewDirection = cos_weighted_hemisphere(a, st, HitNormal);
            const float CosTheta = gclamp(dot(HitNormal, NewDirection), 0.0f, 1.0f);
            const float PDF = gmax(CosTheta / VX_PI, 0.00001f);
            const f3 Attenuation = F3(1.0f) * diffuse_hammon(HitNormal, -rayD, NewDirection, PBR.x);
            RayContribution = RayContribution + RayThroughput * SUNBRDF;
            RayContribution = RayContribution + EmmisivityColor * RayThroughput;
            RayThroughput = RayThroughput * (Albedo * Attenuation / PDF);
            rayD = NewDirection;
            rayO = IntersectionPosition + HitNormal * bias;
        } else {
            float x = gmix(1.0f, 1.05f, a.sun_visibility);
            x = gclamp(x * 1.0f * a.gi_sky_strength, 0.0f, 5.0f);
            f3 rd = rayD;
            rd.y = gclamp(rd.y, 0.125f, 1.5f);  // GetSkyColorAt (:1070-1074)
            const f3 sky = texcube_sample(a.sky, rd) * x;
            RayContribution = RayContribution + sky * RayThroughput;
            Skyhit = true;
            break;
        }
        if (i == 0) {
            const float dao = 2.0f;
            if (T < dao && T > 0.0f) ao = gmax(T / dao, 0.0f);
        }
    }
    return F4(RayContribution.x, RayContribution.y, RayContribution.z, ao);
}

template <bool STATS>
__global__ void __launch_bounds__(256) diffuse_trace_kernel(GridView g, const __grid_constant__ GiArgs a, TraceStatsDev* stats) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    const bool active = px < a.col1 && py < a.row1;
    LaneStats ls = {0u, 0u, 0u, 0u};
    if (active) {
        const size_t i = (size_t)py * a.width + px;
        GiState st;
        st.px = px; st.py = py; st.CurrentBLSample = 0;
        const f2 vtc = pixel_uv(px, py, a.width, a.height);
        f2 tc = vtc;
        if (a.supersample) {
            f2 h = F2(a.halton[0] * 0.75f, a.halton[1] * 0.75f);
            tc = F2(tc.x + h.x / (float)a.width, tc.y + h.y / (float)a.height);
        }
        const float Dist = att_r16f_bilinear(a.g_t, a.gw, a.gh, tc);
        const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
        const f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
        const f3 Normal = normal_from_id(att_r8_nearest(a.g_normal, a.gw, a.gh, tc), F3(0.5f));
        float oSH[4], oCoCg[2], oUtil = 0.0f, oAO = 1.0f, oSky = 0.0f;
        if (Dist < 0.0f) {
            const f3 rd = normalize(ray_direction_at(a.inv_view, a.inv_proj, vtc));
            float SH[6];
            irradiance_to_sh(texcube_sample(a.sky, rd) * 2.66f, Normal, SH);
            oSH[0] = SH[0]; oSH[1] = SH[1]; oSH[2] = SH[2]; oSH[3] = SH[3]; oCoCg[0] = SH[4]; oCoCg[1] = SH[5];
        } else {
            int SPP = iclamp(a.spp, 1, 32);
            if (a.checkerboard) {
                bool CheckerStep = cvt_trunc(((float)px + 0.5f) + ((float)py + 0.5f)) % 2 == a.frame % 2;
                SPP = cvt_trunc(gmix((float)a.spp, (float)a.checker_spp, CheckerStep ? 1.0f : 0.0f));
            }
            SPP = iclamp(SPP, 1, 32);
            if (!a.sun_stronger) SPP *= 2;
            f4 TotalSHy = F4(0.0f, 0.0f, 0.0f, 0.0f);
            f2 CoCg = F2(0.0f, 0.0f);
            f3 radiance = F3(0.0f);
            float Skyhits = 0.0f, AccumulatedAO = 0.0f;
#pragma unroll 1
            for (int s = 0; s < SPP; ++s) {
                f3 d = F3(0.0f);
                bool ss = false;
                f4 x = calculate_diffuse<STATS>(g, a, st, P, Normal, d, ss, &ls);
                f3 xc = gclamp(F3(x.x, x.y, x.z), 0.0f, 8.0f);
                radiance = radiance + xc;
                AccumulatedAO += x.w;
                float SH[6];
                irradiance_to_sh(xc, d, SH);
                TotalSHy = F4(TotalSHy.x + SH[0], TotalSHy.y + SH[1], TotalSHy.z + SH[2], TotalSHy.w + SH[3]);
                CoCg = F2(CoCg.x + SH[4], CoCg.y + SH[5]);
                Skyhits += ss ? 1.0f : 0.0f;
            }
            const float n = (float)SPP;
            AccumulatedAO /= n;
            TotalSHy = F4(TotalSHy.x / n, TotalSHy.y / n, TotalSHy.z / n, TotalSHy.w / n);
            CoCg = F2(CoCg.x / n, CoCg.y / n);
            radiance = radiance / n;
            Skyhits /= n;
            oUtil = gmax(dot(radiance, F3(0.299f, 0.587f, 0.114f)), 0.01f);
            oAO = gclamp(AccumulatedAO, 0.0f, 1.0f);
            oSky = gclamp(Skyhits, 0.0f, 1.0f);
            oSH[0] = gclamp(TotalSHy.x, -100.0f, 100.0f); oSH[1] = gclamp(TotalSHy.y, -100.0f, 100.0f);
            oSH[2] = gclamp(TotalSHy.z, -100.0f, 100.0f); oSH[3] = gclamp(TotalSHy.w, -100.0f, 100.0f);
            oCoCg[0] = gclamp(CoCg.x, -100.0f, 100.0f); oCoCg[1] = gclamp(CoCg.y, -100.0f, 100.0f);
            oUtil = gclamp(oUtil, 0.001f, 64.0f);
        }
        reinterpret_cast<ushort4*>(a.sh)[i] = make_ushort4(float_to_half_bits(oSH[0]), float_to_half_bits(oSH[1]), float_to_half_bits(oSH[2]), float_to_half_bits(oSH[3]));
        reinterpret_cast<ushort2*>(a.cocg)[i] = make_ushort2(float_to_half_bits(oCoCg[0]), float_to_half_bits(oCoCg[1]));
        a.utility[i] = float_to_half_bits(oUtil);
        reinterpret_cast<uchar2*>(a.aosky)[i] = make_uchar2(float_to_unorm8(oAO), float_to_unorm8(oSky));
    }
    if (STATS) flush_stats(stats, ls);
}

}  // namespace

int vxrt_launch_diffuse_trace(vxrt_ctx* c, const vxrt_gi_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GI_SH, p.width, p.height, 8))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GI_COCG, p.width, p.height, 4))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GI_UTILITY, p.width, p.height, 2))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GI_AOSKY, p.width, p.height, 2))) return rc;
    GiArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    a.spp = p.spp; a.checker_spp = p.checker_spp; a.checkerboard = p.checkerboard; a.trace_length = p.trace_length;
    a.shadow_trace_length = p.shadow_trace_length; a.frame = p.current_frame; a.frame_mod128 = p.current_frame_mod128;
    a.supersample = p.supersample;
    a.halton[0] = p.halton[0]; a.halton[1] = p.halton[1];
    for (int i = 0; i < 3; ++i) { a.sun[i] = p.sun_direction[i]; a.moon[i] = p.moon_direction[i]; a.viewer[i] = p.viewer_position[i]; }
    a.sun_visibility = p.sun_visibility; a.gi_sky_strength = p.gi_sky_strength; a.diffuse_light_intensity = p.diffuse_light_intensity;
    a.apply_player_shadow = p.apply_player_shadow;
    // main() prologue (:910-925): SunStronger, LIGHT_COLOR = SunStronger ? SampleSunColor() : vec3(1)
    a.sun_stronger = (-p.sun_direction[1] < 0.01f) ? 1 : 0;
    if (a.sun_stronger) vxrt_host_sun_color(c, p.sun_direction, p.gi_sun_strength, a.light_color);
    else a.light_color[0] = a.light_color[1] = a.light_color[2] = 1.0f;
    const Attachment& gt = c->att[VXRT_ATT_INITIAL_T];
    a.g_t = (const uint16_t*)gt.ptr; a.g_normal = (const uint8_t*)c->att[VXRT_ATT_INITIAL_NORMAL].ptr; a.gw = gt.width; a.gh = gt.height;
    for (int k = 0; k < 4; ++k) a.tex[k] = c->tex[k];
    a.sky = c->sky;
    a.block_data = c->d_block_data;
    a.blue = c->d_blue_noise;
    a.sh = (uint16_t*)c->att[VXRT_ATT_GI_SH].ptr; a.cocg = (uint16_t*)c->att[VXRT_ATT_GI_COCG].ptr;
    a.utility = (uint16_t*)c->att[VXRT_ATT_GI_UTILITY].ptr; a.aosky = (uint8_t*)c->att[VXRT_ATT_GI_AOSKY].ptr;
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    if (c->wavefront) return vxrt_run_bands(c, a.row0, a.row1, [&](int r0, int r1) { GiArgs b = a; b.row0 = r0; b.row1 = r1; return vxrt_launch_diffuse_trace_wavefront(c, &b); });
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    if (c->stats_on) diffuse_trace_kernel<true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats);
    else diffuse_trace_kernel<false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
