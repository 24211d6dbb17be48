// reflect.cu — reflection pass: ReflectionTraceFrag.glsl (main :717-1038), dispatched at
// Core/Pipeline.cpp:3096-3257; attachments Core/Pipeline.cpp:1189.
// One thread per pixel: GGX-importance-sampled (best of 3) reflection ray per sample, Cook-Torrance hit
// shading with the specular lobe zeroed exactly as the reference does, optional screen-space reuse of
// the diffuse SH / shadow attachments, <= 150-iteration shadow ray for the first max(SPP/4,1) hits.
// Not modelled (out-of-scope subsystems, must be off): LPV ambient, projected clouds, player
// reflection, lava UV distortion.
#include "reflect_common.cuh"

namespace {

template <bool STATS>
__global__ void __launch_bounds__(256) reflection_trace_kernel(GridView g, const __grid_constant__ ReflArgs a, TraceStatsDev* stats) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    const bool active = px < a.col1 && py < a.row1;
    LaneStats ls = {0u, 0u, 0u, 0u};
    if (active) {
        const size_t i = (size_t)py * a.width + px;
        RfState st;
        st.px = px; st.py = py; st.CurrentBLSample = 0;
        const f2 vtc = pixel_uv(px, py, a.width, a.height);
        const bool CheckerStep = cvt_trunc(((float)px + 0.5f) + ((float)py + 0.5f)) % 2 == (a.frame % 2);
        int SPP = iclamp(a.spp, 1, 16);
        if (a.checkerboard) SPP = cvt_trunc(gmix((float)a.spp, (float)((a.spp + a.spp % 2) / 2), CheckerStep ? 1.0f : 0.0f));
        SPP = iclamp(SPP, 1, 16);
        const f2 Jitter = F2(gclamp(a.halton[0] * 1.0f, -2.0f, 2.0f), gclamp(a.halton[1] * 1.0f, -2.0f, 2.0f));
        const float tf = a.temporal ? 1.0f : 0.0f;
        const f2 tc = F2(vtc.x + (Jitter.x / (float)a.width) * tf, vtc.y + (Jitter.y / (float)a.height) * tf);
        const float Dist = att_r16f_bilinear(a.g_t, a.gw, a.gh, tc);
        const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
        const f3 viewer = F3(a.viewer[0], a.viewer[1], a.viewer[2]);
        const f3 strong = F3(a.strong[0], a.strong[1], a.strong[2]);
        f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
        f4 oColor = F4(0.0f, 0.0f, 0.0f, 0.0f);
        float oHit = -1.0f, oMask = 0.0f;
        if (!(Dist < 0.0f)) {
            const f3 N0 = normal_from_id(att_r8_nearest(a.g_normal, a.gw, a.gh, tc), F3(1.0f));
            const int mi = wrap_repeat(cvt_floor(vtc.x * (float)a.mw), a.mw), mj = wrap_repeat(cvt_floor(vtc.y * (float)a.mh), a.mh);
            const uchar4 pb = __ldg(reinterpret_cast<const uchar4*>(a.gb_pbr) + ((size_t)mj * a.mw + mi));
            const f4 PBRMap = F4(unorm8_to_float(pb.x), unorm8_to_float(pb.y), unorm8_to_float(pb.z), unorm8_to_float(pb.w));
            const f3 I = normalize(P - viewer);
            P = P + N0 * 0.035f;
            float nv[3], shv[4], ccv[2];
            att_half_bilinear<3>(a.gb_normal, a.mw, a.mh, vtc, nv);
            att_half_bilinear<4>(a.gi_sh, a.iw, a.ih, vtc, shv);
            att_half_bilinear<2>(a.gi_cocg, a.iw, a.ih, vtc, ccv);
            const f3 NormalMappedInitial = F3(nv[0], nv[1], nv[2]);
            const f4 DiffuseSH = F4(shv[0], shv[1], shv[2], shv[3]);
            const f2 DiffuseCoCg = F2(ccv[0], ccv[1]);
            const f3 BaseIndirectDiffuse = sh_to_irradiance_a(DiffuseSH, DiffuseCoCg);
            if (PBRMap.x >= 0.865f && a.derive_sh) {
                f3 r = derive_specular_from_diffuse_sh(DiffuseSH, sh_to_irradiance(DiffuseSH, DiffuseCoCg, NormalMappedInitial), I, NormalMappedInitial);
                oColor = F4(r.x, r.y, r.z, 0.0f); oHit = 0.5f; oMask = 0.0f;
            } else {
                const float RoughnessAt = PBRMap.x;
                const float RoughnessBias = gmix(1.0f, 0.85f, a.roughness_bias ? 1.0f : 0.0f);
                float ComputedShadow = 0.0f;
                int ShadowItr = 0;
                float AveragedHitDistance = 0.001f, TotalMeaningfulHits = 0.0f, EmissivityMask = 0.0f;
                int total_hits = 0;
                f4 TotalColor = F4(0.0f, 0.0f, 0.0f, 0.0f);
                const f3 MIXED = F3(a.color_mixed[0], a.color_mixed[1], a.color_mixed[2]);
#pragma unroll 1
                for (int s = 0; s < SPP; ++s) {
                    const f3 ReflectionNormal = a.rough ? get_reflection_direction(a, st, NormalMappedInitial, gclamp(RoughnessAt * RoughnessBias, 0.01f, 1.0f)) : NormalMappedInitial;
                    const f3 R = reflect(I, ReflectionNormal);
                    TraceResult h = traverse_df<STATS>(g, P, R, a.trace_length, &ls);
                    const float T = h.t;
                    const f3 Normal = h.normal;
                    const f3 HitPosition = P + (R * T);
                    if (T > 0.0f) {
                        f2 UV = F2(0.0f, 0.0f);
                        f3 Tangent = F3(0.0f), Bitangent = F3(0.0f);
                        calculate_vectors(HitPosition, Normal, Tangent, Bitangent, UV);
                        UV.y = 1.0f - UV.y;
                        const int reference_id = iclamp(h.block, 0, 127);
                        bool ReprojectionSuccessful = false;
                        f2 SS = F2(-1.0f, -1.0f);
                        f3 Ambient = BaseIndirectDiffuse;
                        if (a.reproject) {
                            // ReprojectReflectionToScreenSpace (:574-586)
                            f4 pp = mat4_mul(a.proj_view, F4(HitPosition.x, HitPosition.y, HitPosition.z, 1.0f));
                            f3 q = F3(pp.x / pp.w, pp.y / pp.w, pp.z / pp.w);
                            SS = F2(q.x * 0.5f + 0.5f, q.y * 0.5f + 0.5f);
                            const float d2 = att_r16f_bilinear(a.g_t, a.gw, a.gh, SS);
                            const f3 PosAt = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, SS)) * d2;
                            const f3 NormalAt = normal_from_id(att_r8_nearest(a.g_normal, a.gw, a.gh, SS), F3(1.0f));
                            const f3 df = PosAt - HitPosition;
                            const f3 diff = F3(fabsf(df.x), fabsf(df.y), fabsf(df.z));
                            const float Error = dot(diff, diff);
                            ReprojectionSuccessful = Error < 0.095f && eq3(NormalAt, Normal) && in_thresholded_screen_space(SS);
                            if (ReprojectionSuccessful) {
                                float rs[4], rc[2], ra[2];
                                att_half_bilinear<4>(a.gi_sh, a.iw, a.ih, SS, rs);
                                att_half_bilinear<2>(a.gi_cocg, a.iw, a.ih, SS, rc);
                                Ambient = sh_to_irradiance_a(F4(rs[0], rs[1], rs[2], rs[3]), F2(rc[0], rc[1]));
                                att_unorm8_bilinear<2>(a.gi_aosky, a.iw, a.ih, SS, ra);
                                const float ReprojectedVXAO = powf(ra[0], 0.75f);
                                if (d2 > 0.0f) {
                                    if (distance(PosAt, cam) < 40.0f) Ambient = Ambient * ReprojectedVXAO;
                                }
                            }
                        }
                        if (a.lpv_gi && !ReprojectionSuccessful) Ambient = rf_approximate_gi_lpv(a, px, py, HitPosition + Normal * 0.5f, BaseIndirectDiffuse);
                        f4 ids = F4((float)__ldg(a.block_data + reference_id), (float)__ldg(a.block_data + 128 + reference_id),
                                    (float)__ldg(a.block_data + 256 + reference_id), (float)__ldg(a.block_data + 384 + reference_id));
                        if (reference_id == a.grass[0]) {
                            if (eq3(Normal, face_normal(4)) || eq3(Normal, face_normal(5)) || eq3(Normal, face_normal(0)) || eq3(Normal, face_normal(1))) { ids.x = (float)a.grass[4]; ids.y = (float)a.grass[5]; ids.z = (float)a.grass[6]; }
                            else if (eq3(Normal, face_normal(2))) { ids.x = (float)a.grass[1]; ids.y = (float)a.grass[2]; ids.z = (float)a.grass[3]; }
                            else if (eq3(Normal, face_normal(3))) { ids.x = (float)a.grass[7]; ids.y = (float)a.grass[8]; ids.z = (float)a.grass[9]; }
                        }
                        const f3 Albedo = xyz(texarray_sample(a.tex[VXRT_TEX_ALBEDO], UV.x, UV.y, ids.x, 0.0f));
                        const f3 Radiance = MIXED * 0.6f;
                        const f4 SampledPBR = texarray_sample(a.tex[VXRT_TEX_PBR], UV.x, UV.y, ids.z, 0.0f);
                        const float AO = pow2_mul(SampledPBR.w);
                        const bool PlayerInShadow = get_player_intersect(viewer, HitPosition + Normal * 0.035f, strong);
                        if (ShadowItr < (SPP / 4 > 1 ? SPP / 4 : 1)) {
                            if (!PlayerInShadow) {
                                if (ReprojectionSuccessful && a.reproject && in_thresholded_screen_space(SS)) {
                                    float sv[1];
                                    att_unorm8_bilinear<1>(a.shadow, a.sw, a.sh, SS, sv);
                                    ComputedShadow = sv[0];
                                } else {
                                    // GetShadowAt (:1327-1346)
                                    const f3 spos = HitPosition + Normal * 0.055f;
                                    if (get_player_intersect(viewer, spos, strong)) ComputedShadow = 1.0f;
                                    else {
                                        TraceResult sh = traverse_df<STATS>(g, spos, strong, a.shadow_trace_length, &ls);
                                        ComputedShadow = sh.t > 0.0f ? 1.0f : 0.0f;
                                    }
                                }
                            } else {
                                ComputedShadow = 1.0f;
                            }
                            ShadowItr = ShadowItr + 1;
                        }
                        Ambient = (Ambient * 1.0f * gclamp(AO, 0.1f, 1.0f)) * Albedo;
                        const f3 nm = xyz(texarray_sample(a.tex[VXRT_TEX_NORMAL], UV.x, UV.y, ids.y, 3.0f)) * 2.0f - F3(1.0f);
                        const f3 NormalMapped = mat3_mul(Tangent, Bitangent, Normal, nm);
                        f3 DirectLighting = Ambient + rf_directional_light(viewer, HitPosition, strong, Radiance, Albedo, NormalMapped,
                                                                           F3(SampledPBR.x, SampledPBR.y, SampledPBR.z), ComputedShadow);
                        if (ids.w > -0.5f) {
                            float Emissivity = texarray_sample(a.tex[VXRT_TEX_EMISSIVE], UV.x, UV.y, ids.w, 2.0f).x;
                            if (Emissivity > 0.1f) {
                                const float m = 19.0f, lbiasx = 0.02501f, lbiasy = 0.03001f;
                                Emissivity *= (UV.x > lbiasx && UV.x < 1.0f - lbiasx && UV.y > lbiasy && UV.y < 1.0f - lbiasy) ? 1.0f : 0.0f;
                                const float Flicker = 1.0f;
                                DirectLighting = Albedo * gmax(Emissivity * m * Flicker, 2.0f);
                                EmissivityMask = 1.0f;
                            }
                        }
                        TotalColor = F4(TotalColor.x + DirectLighting.x, TotalColor.y + DirectLighting.y, TotalColor.z + DirectLighting.z, TotalColor.w + 1.0f);
                        AveragedHitDistance += T;
                        TotalMeaningfulHits += 1.0f;
                    } else {
                        const f3 Atmos = texcube_sample(a.sky, normalize(R));
                        const f3 am = Atmos * gmix(1.0f, 1.175f, (PBRMap.y > 0.05f) ? 1.0f : 0.0f);
                        TotalColor = F4(TotalColor.x + am.x, TotalColor.y + am.y, TotalColor.z + am.z, TotalColor.w + 1.0f);
                    }
                    total_hits++;
                }
                AveragedHitDistance /= gmax(TotalMeaningfulHits, 0.01f);
                const float th = (float)total_hits;
                TotalColor = F4(TotalColor.x / th, TotalColor.y / th, TotalColor.z / th, TotalColor.w / th);
                oColor = F4(gclamp(TotalColor.x, 0.0000001f, 100.0f), gclamp(TotalColor.y, 0.0000001f, 100.0f), gclamp(TotalColor.z, 0.0000001f, 100.0f),
                            gclamp(TotalColor.w, 0.0000001f, 100.0f));
                oHit = gclamp(TotalMeaningfulHits > 0.01f ? AveragedHitDistance : -1.0f, -10.0f, 200.0f);
                oMask = gclamp(EmissivityMask, 0.0f, 1.0f);
            }
        }
        reinterpret_cast<ushort4*>(a.color)[i] = make_ushort4(float_to_half_bits(oColor.x), float_to_half_bits(oColor.y), float_to_half_bits(oColor.z), float_to_half_bits(oColor.w));
        a.hitdist[i] = float_to_half_bits(oHit);
        a.emissive[i] = float_to_unorm8(oMask);
    }
    if (STATS) flush_stats(stats, ls);
}

}  // namespace

int vxrt_launch_reflection_trace(vxrt_ctx* c, const vxrt_reflection_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_REFL_COLOR, p.width, p.height, 8))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_REFL_HITDIST, p.width, p.height, 2))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_REFL_EMISSIVE, p.width, p.height, 1))) return rc;
    ReflArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    // u_Projection * u_View: mat4 * mat4 evaluated column by column with the pinned mat4*vec4 association
    for (int col = 0; col < 4; ++col)
        for (int r = 0; r < 4; ++r) {
            const float* v = p.view + 4 * col;
            a.proj_view[4 * col + r] = (p.projection[r] * v[0] + p.projection[4 + r] * v[1]) + (p.projection[8 + r] * v[2] + p.projection[12 + r] * v[3]);
        }
    a.width = p.width; a.height = p.height;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    a.spp = p.spp; a.checkerboard = p.checkerboard; a.trace_length = p.trace_length; a.shadow_trace_length = p.shadow_trace_length;
    a.frame = p.current_frame; a.frame_mod128 = p.current_frame_mod128;
    a.rough = p.rough_reflections; a.roughness_bias = p.roughness_bias; a.temporal = p.temporal; a.reproject = p.reproject_to_screen_space;
    a.derive_sh = p.derive_from_diffuse_sh;
    a.halton[0] = p.halton[0]; a.halton[1] = p.halton[1];
    for (int i = 0; i < 3; ++i) { a.sun[i] = p.sun_direction[i]; a.moon[i] = p.moon_direction[i]; a.strong[i] = p.stronger_light_direction[i]; a.viewer[i] = p.viewer_position[i]; }
    // main() prologue (:722-727): SAMPLED_COLOR_MIXED = mix(SampleSunColor(), SampleMoonColor(), SunVisibility)
    float sc[3], mc[3];
    vxrt_host_sun_color(c, p.sun_direction, p.sun_strength_modifier, sc);
    vxrt_host_moon_color(c, p.moon_direction, p.moon_strength_modifier, mc);
    float sdot = (p.sun_direction[0] * 0.0f + p.sun_direction[1] * 1.0f) + p.sun_direction[2] * 0.0f;
    float sv = sdot + 0.05f;
    sv = sv < 0.0f ? 0.0f : sv; sv = (0.1f < sv) ? 0.1f : sv;
    sv = sv * 12.0f;
    sv = 1.0f - sv;
    for (int i = 0; i < 3; ++i) a.color_mixed[i] = sc[i] * (1.0f - sv) + mc[i] * sv;
    for (int i = 0; i < 10; ++i) a.grass[i] = p.grass_props[i];
    const Attachment& gt = c->att[VXRT_ATT_INITIAL_T];
    a.g_t = (const uint16_t*)gt.ptr; a.g_normal = (const uint8_t*)c->att[VXRT_ATT_INITIAL_NORMAL].ptr; a.gw = gt.width; a.gh = gt.height;
    const Attachment& gn = c->att[VXRT_ATT_GBUF_NORMAL];
    a.gb_normal = (const uint16_t*)gn.ptr; a.gb_pbr = (const uint8_t*)c->att[VXRT_ATT_GBUF_PBR].ptr; a.mw = gn.width; a.mh = gn.height;
    const Attachment& gs = c->att[VXRT_ATT_GI_SH];
    a.gi_sh = (const uint16_t*)gs.ptr; a.gi_cocg = (const uint16_t*)c->att[VXRT_ATT_GI_COCG].ptr; a.gi_aosky = (const uint8_t*)c->att[VXRT_ATT_GI_AOSKY].ptr;
    a.iw = gs.width; a.ih = gs.height;
    const Attachment& sh = c->att[c->shadow_source];
    a.shadow = (const uint8_t*)sh.ptr; a.sw = sh.width; a.sh = sh.height;
    for (int k = 0; k < 4; ++k) a.tex[k] = c->tex[k];
    a.sky = c->sky;
    a.block_data = c->d_block_data;
    a.blue = c->d_blue_noise;
    a.color = (uint16_t*)c->att[VXRT_ATT_REFL_COLOR].ptr; a.hitdist = (uint16_t*)c->att[VXRT_ATT_REFL_HITDIST].ptr;
    a.emissive = (uint8_t*)c->att[VXRT_ATT_REFL_EMISSIVE].ptr;
    // ApproximateGILPV: the propagation volume and the average block colours must be there (vxrt_cuda_lpv_repropagate,
    // vxrt_cuda_lpv_average_colors / _set_average_colors)
    a.lpv_gi = p.lpv_gi != 0; a.decoupled_gi = p.use_decoupled_gi != 0; a.ss_sky_valid = p.screen_space_skylighting_valid != 0;
    a.sun_stronger = p.stronger_light_direction[0] == p.sun_direction[0] && p.stronger_light_direction[1] == p.sun_direction[1] &&
                     p.stronger_light_direction[2] == p.sun_direction[2];   // SunStronger (:809)
    a.lpv.level = nullptr; a.lpv.type = nullptr; a.lpv.avg = nullptr; a.lpv.nx = c->nx; a.lpv.ny = c->ny; a.lpv.nz = c->nz;
    a.lpv.dx = a.lpv.dy = a.lpv.dz = 0.0f;
    a.sky_ambient_g[0] = a.sky_ambient_g[1] = a.sky_ambient_g[2] = 0.0f;
    if (a.lpv_gi) {
        if (!c->lpv_valid || !c->d_lpv) return vxrt_fail(VXRT_E_STATE, "reflection_trace: lpv_gi needs the light propagation volume (vxrt_cuda_lpv_repropagate)");
        if (!c->d_lpv_avg) return vxrt_fail(VXRT_E_STATE, "reflection_trace: lpv_gi needs the average block colours (vxrt_cuda_lpv_average_colors)");
        a.lpv.level = c->d_lpv; a.lpv.type = c->d_lpv + c->nvox; a.lpv.avg = reinterpret_cast<const float4*>(c->d_lpv_avg);
        const float up[3] = {0.0f, 1.0f, 0.0f};
        vxrt_host_sky_sample(c, up, a.sky_ambient_g);
    }
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    if (c->wavefront) return vxrt_run_bands(c, a.row0, a.row1, [&](int r0, int r1) { ReflArgs b = a; b.row0 = r0; b.row1 = r1; return vxrt_launch_reflection_trace_wavefront(c, &b); });
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    if (c->stats_on) reflection_trace_kernel<true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats);
    else reflection_trace_kernel<false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
