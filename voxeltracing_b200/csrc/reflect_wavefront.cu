// reflect_wavefront.cu — the reflection pass (ReflectionTraceFrag.glsl main :717-1038) as a wavefront
// pipeline; same arithmetic as reflect.cu, bit-identical attachments.
//
//   gen      sample 0: everything main() does before the sample loop (G-buffer reads, incident vector,
//            diffuse-SH ambient, early outs); every sample: best-of-3 GGX microfacet normal -> reflection ray
//   trace    reflection rays                                                   <= trace_length iterations
//   shadeA   hit: material fetch, ambient (optionally the screen-space reuse), the shadow-independent part
//            of CalculateDirectionalLight, emissive override; decides whether this sample casts the
//            (<= 150 iteration) shadow ray and appends it to a compacted queue.  miss: sky colour.
//   trace    shadow queue (single light direction)
//   shadeB   combine with the shadow term, accumulate colour / hit distance
//   resolve  averages, clamps, attachment formats
#include "reflect_common.cuh"
#include "trace_queue.cuh"

#ifndef VX_SHADE_OCC
#define VX_SHADE_OCC 5  // minimum CTAs per SM asked of the gen / shade kernels: 48 registers; 4 / 5 / 6 measured, GI 1.122 / 1.089 / 1.090 ms
#endif

namespace {

struct RfWf {
    // per pixel, persistent across samples
    float4* P;        // biased hit position (xyz), roughness (w)
    float4* I;        // incident direction (xyz), metal flag PBRMap.y > 0.05 (w)
    float4* Nmap;     // normal-mapped G-buffer normal (xyz)
    float4* Total;    // TotalColor
    float4* misc;     // AveragedHitDistance, TotalMeaningfulHits, EmissivityMask, ComputedShadow
    int4* cnt;        // ShadowItr, total_hits, SPP, CurrentBLSample (-1: pixel has no paths)
    // per sample
    float4* rayD;     // reflection direction (xyz), active flag (w)
    float* hitT;
    unsigned* hitInfo;
    float4* Amb;      // Ambient (xyz); w: 0 = sample already accumulated, 1 = waits for shadeB, 2 = emissive override (no sun term)
                      // DEFER (ctx.h refl_defer_gi), w == 1: xyz = Albedo, the ambient product is formed in shadeB
    float4* Res;      // max(Result, 0) of CalculateDirectionalLight (xyz); w: 1 = take ComputedShadow from shadowRes
                      // DEFER: |w| = clamp(AO, 0.1, 1), w < 0 = take ComputedShadow from shadowRes
    float* shadowRes;
    float4* qShadowO; // compacted shadow rays
    int* counters;
};

VXD unsigned pack_hit(const TraceResult& r) {
    unsigned face = 7u;
    if (r.normal.z == 1.0f) face = 0u; else if (r.normal.z == -1.0f) face = 1u;
    else if (r.normal.y == 1.0f) face = 2u; else if (r.normal.y == -1.0f) face = 3u;
    else if (r.normal.x == -1.0f) face = 4u; else if (r.normal.x == 1.0f) face = 5u;
    return (unsigned)(r.block & 0xff) | (face << 8);
}
VXD f3 unpack_normal(unsigned info) {
    unsigned face = (info >> 8) & 7u;
    return face == 7u ? F3(0.0f) : face_normal((int)face);
}

struct ReflRays {
    RfWf w;
    VXD bool fetch(int idx, f3& o, f3& d) const {
        const float4 d4 = w.rayD[idx];
        if (d4.w == 0.0f) return false;
        const float4 o4 = w.P[idx];
        o = F3(o4.x, o4.y, o4.z); d = F3(d4.x, d4.y, d4.z);
        return true;
    }
    VXD void store(int idx, const TraceResult& r) const {
        w.hitT[idx] = r.t;
        w.hitInfo[idx] = pack_hit(r);
    }
};
template <bool STATS>
__global__ void VX_TRACE_BOUNDS rf_wf_trace_kernel(GridView g, RfWf w, int n, int max_iter, TraceStatsDev* stats) {
    LaneStats ls = {0u, 0u, 0u, 0u};
    ReflRays pol = {w};
    trace_queue<STATS>(g, pol, n, max_iter, &ls);
    if (STATS) flush_stats(stats, ls);
}
struct ReflShadowRays {
    RfWf w;
    f3 light;
    VXD bool fetch(int idx, f3& o, f3& d) const {
        const float4 o4 = w.qShadowO[idx];
        o = F3(o4.x, o4.y, o4.z); d = light;
        return true;
    }
    VXD void store(int idx, const TraceResult& r) const {
        w.shadowRes[__float_as_int(w.qShadowO[idx].w)] = r.t > 0.0f ? 1.0f : 0.0f;
    }
};
template <bool STATS>
__global__ void VX_TRACE_BOUNDS rf_wf_trace_shadow_kernel(GridView g, RfWf w, f3 light, int max_iter, TraceStatsDev* stats) {
    const int count = w.counters[0];
    LaneStats ls = {0u, 0u, 0u, 0u};
    ReflShadowRays pol = {w, light};
    trace_queue<STATS>(g, pol, count, max_iter, &ls);
    if (STATS) flush_stats(stats, ls);
}

__global__ void __launch_bounds__(256, VX_SHADE_OCC) rf_wf_gen_kernel(const __grid_constant__ ReflArgs a, RfWf w, int sample) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    // sample 0 keeps what it has just computed in registers; the incident direction and the normal are only stored for a pixel that takes
    // further samples, its metal flag travels in rayD.w (2 instead of 1)
    float4 p4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), i4 = p4, n4 = p4;
    int4 cnt = make_int4(0, 0, 0, 0);
    if (sample == 0) {
        const size_t pi = (size_t)py * a.width + px;
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
        f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
        bool done = false;
        f4 oColor = F4(0.0f, 0.0f, 0.0f, 0.0f);
        float oHit = -1.0f, oMask = 0.0f;
        if (Dist < 0.0f) {
            done = true;
        } else {
            const f3 N0 = normal_from_id(att_r8_nearest(a.g_normal, a.gw, a.gh, tc), F3(1.0f));
            const int mi = wrap_repeat(cvt_floor(vtc.x * (float)a.mw), a.mw), mj = wrap_repeat(cvt_floor(vtc.y * (float)a.mh), a.mh);
            const uchar4 pb = __ldg(reinterpret_cast<const uchar4*>(a.gb_pbr) + ((size_t)mj * a.mw + mi));
            const f4 PBRMap = F4(unorm8_to_float(pb.x), unorm8_to_float(pb.y), unorm8_to_float(pb.z), unorm8_to_float(pb.w));
            const f3 I = normalize(P - viewer);
            P = P + N0 * 0.035f;
            float nv[3];
            att_half_bilinear<3>(a.gb_normal, a.mw, a.mh, vtc, nv);
            const f3 Nmap = F3(nv[0], nv[1], nv[2]);
            // The diffuse SH of the GI (u_DiffuseSHy / u_DiffuseCoCg at v_TexCoords) is read HERE only for the DeriveFromDiffuseSH shortcut; the
            // ambient base of a hit (BaseIndirectDiffuse) is evaluated from the same texels in shade_a, so that without the shortcut gen and
            // the first trace do not depend on the GI pass (api.cu runs them beside it).
            if (PBRMap.x >= 0.865f && a.derive_sh) {
                float shv[4], ccv[2];
                att_half_bilinear<4>(a.gi_sh, a.iw, a.ih, vtc, shv);
                att_half_bilinear<2>(a.gi_cocg, a.iw, a.ih, vtc, ccv);
                const f4 DiffuseSH = F4(shv[0], shv[1], shv[2], shv[3]);
                const f2 DiffuseCoCg = F2(ccv[0], ccv[1]);
                f3 r = derive_specular_from_diffuse_sh(DiffuseSH, sh_to_irradiance(DiffuseSH, DiffuseCoCg, Nmap), I, Nmap);
                oColor = F4(r.x, r.y, r.z, 0.0f); oHit = 0.5f; oMask = 0.0f;
                done = true;
            } else {
                p4 = make_float4(P.x, P.y, P.z, PBRMap.x);
                i4 = make_float4(I.x, I.y, I.z, (PBRMap.y > 0.05f) ? 1.0f : 0.0f);
                n4 = make_float4(Nmap.x, Nmap.y, Nmap.z, 0.0f);
                w.P[i] = p4;
                if (SPP > 1) { w.I[i] = i4; w.Nmap[i] = n4; }
                // TotalColor = 0 and (AveragedHitDistance, TotalMeaningfulHits, EmissivityMask, ComputedShadow) = (0.001, 0, 0, 0) are not stored:
                // the first sample's shading starts from these constants instead of loading them (`first` in shade_a / shade_b / final)
                cnt = make_int4(0, 0, SPP, 0);
            }
        }
        if (done) {
            reinterpret_cast<ushort4*>(a.color)[pi] = make_ushort4(float_to_half_bits(oColor.x), float_to_half_bits(oColor.y), float_to_half_bits(oColor.z), float_to_half_bits(oColor.w));
            a.hitdist[pi] = float_to_half_bits(oHit);
            a.emissive[pi] = float_to_unorm8(oMask);
            w.cnt[i] = make_int4(0, 0, 0, -1);
            w.rayD[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            return;
        }
    }
    if (sample != 0) {
        cnt = w.cnt[i];
        if (cnt.w < 0 || sample >= cnt.z) {
            w.rayD[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            return;
        }
        p4 = w.P[i]; i4 = w.I[i]; n4 = w.Nmap[i];
    }
    const f3 Nmap = F3(n4.x, n4.y, n4.z);
    RfState st;
    st.px = px; st.py = py; st.CurrentBLSample = cnt.w;
    const float RoughnessBias = gmix(1.0f, 0.85f, a.roughness_bias ? 1.0f : 0.0f);
    const f3 ReflectionNormal = a.rough ? get_reflection_direction(a, st, Nmap, gclamp(p4.w * RoughnessBias, 0.01f, 1.0f)) : Nmap;
    const f3 R = reflect(F3(i4.x, i4.y, i4.z), ReflectionNormal);
    cnt.w = st.CurrentBLSample;
    w.cnt[i] = cnt;
    w.rayD[i] = make_float4(R.x, R.y, R.z, i4.w != 0.0f ? 2.0f : 1.0f);
}

// BaseIndirectDiffuse = SHToIrridiance(texture(u_DiffuseSHy, v_TexCoords), texture(u_DiffuseCoCg, v_TexCoords)) of main(): a function of the
// PIXEL, not of the hit
VXD f3 rf_base_indirect_diffuse(const ReflArgs& a, int px, int py) {
    float bsh[4], bcc[2];
    const f2 vtc = pixel_uv(px, py, a.width, a.height);
    att_half_bilinear<4>(a.gi_sh, a.iw, a.ih, vtc, bsh);
    att_half_bilinear<2>(a.gi_cocg, a.iw, a.ih, vtc, bcc);
    return sh_to_irradiance_a(F4(bsh[0], bsh[1], bsh[2], bsh[3]), F2(bcc[0], bcc[1]));
}

// LPVGI: ApproximateGILPV for hits whose reprojection failed (a.lpv_gi); a template flag so the default path compiles without it.
// DEFER (only without a.reproject / a.lpv_gi, where the ambient term of a hit is ((BaseIndirectDiffuse * 1) * clamp(AO)) * Albedo): the kernel
// does not read the GI attachments at all; it leaves Albedo and the AO factor in Amb / Res and shade_b forms the product from the same
// operands in the same order.
template <bool LPVGI, bool DEFER>
__global__ void __launch_bounds__(256, VX_SHADE_OCC) rf_wf_shade_a_kernel(const __grid_constant__ ReflArgs a, RfWf w, int first) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    const bool inside = px < a.col1 && py < a.row1;
    const int i = inside ? (py - a.row0) * (a.col1 - a.col0) + (px - a.col0) : 0;
    bool push_shadow = false;
    f3 shadow_o = F3(0.0f);
    const float4 d4 = inside ? w.rayD[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (inside && d4.w != 0.0f) {
        const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
        const f3 viewer = F3(a.viewer[0], a.viewer[1], a.viewer[2]);
        const f3 strong = F3(a.strong[0], a.strong[1], a.strong[2]);
        const float4 p4 = w.P[i];
        const f3 P = F3(p4.x, p4.y, p4.z), R = F3(d4.x, d4.y, d4.z);
        const float T = w.hitT[i];
        const unsigned info = w.hitInfo[i];
        int4 cnt = w.cnt[i];
        float4 misc = first ? make_float4(0.001f, 0.0f, 0.0f, 0.0f) : w.misc[i];
        float4 Total = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : w.Total[i];
        const int SPP = cnt.z;
        if (T > 0.0f) {
            const f3 Normal = unpack_normal(info);
            const f3 HitPosition = P + (R * T);
            f2 UV = F2(0.0f, 0.0f);
            f3 Tangent = F3(0.0f), Bitangent = F3(0.0f);
            calculate_vectors(HitPosition, Normal, Tangent, Bitangent, UV);
            UV.y = 1.0f - UV.y;
            const int reference_id = iclamp((int)(info & 0xffu), 0, 127);
            bool ReprojectionSuccessful = false;
            f2 SS = F2(-1.0f, -1.0f);
            const f3 b4 = DEFER ? F3(0.0f) : rf_base_indirect_diffuse(a, px, py);
            f3 Ambient = b4;
            if (!DEFER && a.reproject) {
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
            if (LPVGI && !ReprojectionSuccessful) Ambient = rf_approximate_gi_lpv(a, px, py, HitPosition + Normal * 0.5f, F3(b4.x, b4.y, b4.z));
            f4 ids = F4((float)__ldg(a.block_data + reference_id), (float)__ldg(a.block_data + 128 + reference_id),
                        (float)__ldg(a.block_data + 256 + reference_id), (float)__ldg(a.block_data + 384 + reference_id));
            if (reference_id == a.grass[0]) {
                if (eq3(Normal, face_normal(4)) || eq3(Normal, face_normal(5)) || eq3(Normal, face_normal(0)) || eq3(Normal, face_normal(1))) { ids.x = (float)a.grass[4]; ids.y = (float)a.grass[5]; ids.z = (float)a.grass[6]; }
                else if (eq3(Normal, face_normal(2))) { ids.x = (float)a.grass[1]; ids.y = (float)a.grass[2]; ids.z = (float)a.grass[3]; }
                else if (eq3(Normal, face_normal(3))) { ids.x = (float)a.grass[7]; ids.y = (float)a.grass[8]; ids.z = (float)a.grass[9]; }
            }
            const f3 Albedo = xyz(texarray_sample(a.tex[VXRT_TEX_ALBEDO], UV.x, UV.y, ids.x, 0.0f));
            const f3 Radiance = F3(a.color_mixed[0], a.color_mixed[1], a.color_mixed[2]) * 0.6f;
            const f4 SampledPBR = texarray_sample(a.tex[VXRT_TEX_PBR], UV.x, UV.y, ids.z, 0.0f);
            const float AO = pow2_mul(SampledPBR.w);
            const bool PlayerInShadow = get_player_intersect(viewer, HitPosition + Normal * 0.035f, strong);
            float from_ray = 0.0f;
            if (cnt.x < (SPP / 4 > 1 ? SPP / 4 : 1)) {
                if (!PlayerInShadow) {
                    if (ReprojectionSuccessful && a.reproject && in_thresholded_screen_space(SS)) {
                        float sv[1];
                        att_unorm8_bilinear<1>(a.shadow, a.sw, a.sh, SS, sv);
                        misc.w = sv[0];
                    } else {
                        shadow_o = HitPosition + Normal * 0.055f;
                        if (get_player_intersect(viewer, shadow_o, strong)) misc.w = 1.0f;
                        else { push_shadow = true; from_ray = 1.0f; }
                    }
                } else {
                    misc.w = 1.0f;
                }
                cnt.x = cnt.x + 1;
            }
            const float AOFactor = gclamp(AO, 0.1f, 1.0f);
            if (!DEFER) Ambient = (Ambient * 1.0f * AOFactor) * Albedo;
            const f3 nm = xyz(texarray_sample(a.tex[VXRT_TEX_NORMAL], UV.x, UV.y, ids.y, 3.0f)) * 2.0f - F3(1.0f);
            const f3 NormalMapped = mat3_mul(Tangent, Bitangent, Normal, nm);
            // CalculateDirectionalLight = max(Result, 0) * clamp(1 - min(shadow, 1), 0, 1): evaluate the first factor now
            f3 Res = rf_directional_light(viewer, HitPosition, strong, Radiance, Albedo, NormalMapped, F3(SampledPBR.x, SampledPBR.y, SampledPBR.z), 0.0f);
            float override_flag = 0.0f;
            if (ids.w > -0.5f) {
                float Emissivity = texarray_sample(a.tex[VXRT_TEX_EMISSIVE], UV.x, UV.y, ids.w, 2.0f).x;
                if (Emissivity > 0.1f) {
                    const float m = 19.0f, lbiasx = 0.02501f, lbiasy = 0.03001f;
                    Emissivity *= (UV.x > lbiasx && UV.x < 1.0f - lbiasx && UV.y > lbiasy && UV.y < 1.0f - lbiasy) ? 1.0f : 0.0f;
                    const float Flicker = 1.0f;
                    Ambient = Albedo * gmax(Emissivity * m * Flicker, 2.0f);  // DirectLighting is replaced outright
                    override_flag = 1.0f;
                    misc.z = 1.0f;
                }
            }
            if (DEFER) {
                const f3 av = override_flag != 0.0f ? Ambient : Albedo;
                w.Amb[i] = make_float4(av.x, av.y, av.z, 1.0f + override_flag);
                w.Res[i] = make_float4(Res.x, Res.y, Res.z, from_ray != 0.0f ? -AOFactor : AOFactor);
            } else {
                w.Amb[i] = make_float4(Ambient.x, Ambient.y, Ambient.z, 1.0f + override_flag);
                w.Res[i] = make_float4(Res.x, Res.y, Res.z, from_ray);
            }
            misc.x += T;
            misc.y += 1.0f;
        } else {
            const f3 Atmos = texcube_sample(a.sky, normalize(R));
            const f3 am = Atmos * gmix(1.0f, 1.175f, d4.w == 2.0f ? 1.0f : 0.0f);   // the metal flag (PBRMap.y > 0.05) rides in rayD.w
            Total = make_float4(Total.x + am.x, Total.y + am.y, Total.z + am.z, Total.w + 1.0f);
            w.Amb[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            w.Total[i] = Total;
        }
        cnt.y = cnt.y + 1;
        w.cnt[i] = cnt;
        w.misc[i] = misc;
    } else if (inside) {
        w.Amb[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ms = __ballot_sync(0xffffffffu, push_shadow);
    if (ms) {
        int base = 0;
        if (lane == (unsigned)(__ffs(ms) - 1)) base = atomicAdd(w.counters + 0, __popc(ms));
        base = __shfl_sync(0xffffffffu, base, __ffs(ms) - 1);
        if (push_shadow) w.qShadowO[base + __popc(ms & ((1u << lane) - 1u))] = make_float4(shadow_o.x, shadow_o.y, shadow_o.z, __int_as_float(i));
    }
}

// shade_b of one path: the sample's sun-shadow term once the shadow ray is back; leaves the updated totals in Total / misc (written back
// only when `store`)
template <bool DEFER>
VXD bool rf_shade_b_path(const ReflArgs& a, const RfWf& w, int px, int py, int i, float4& Total, float4& misc, bool store, bool first) {
    const float4 amb = w.Amb[i];
    if (amb.w == 0.0f) return false;
    const float4 res = w.Res[i];
    misc = w.misc[i];
    if (DEFER ? res.w < 0.0f : res.w != 0.0f) { misc.w = w.shadowRes[i]; if (store) w.misc[i] = misc; }   // ComputedShadow = GetShadowAt(...)
    f3 Direct = F3(amb.x, amb.y, amb.z);
    if (amb.w < 2.0f) {
        if (DEFER) {   // shade_a's `Ambient = (Ambient * 1.0f * clamp(AO, 0.1, 1)) * Albedo` with Ambient = BaseIndirectDiffuse
            const f3 Albedo = Direct;
            f3 Ambient = rf_base_indirect_diffuse(a, px, py);
            Ambient = (Ambient * 1.0f * fabsf(res.w)) * Albedo;
            Direct = Ambient;
        }
        const float Shadow = gmin(misc.w, 1.0f);
        Direct = Direct + F3(res.x, res.y, res.z) * gclamp(1.0f - Shadow, 0.0f, 1.0f);
    }
    Total = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : w.Total[i];   // a path that hit something has not stored its total yet (shade_a only stores sky samples)
    Total = make_float4(Total.x + Direct.x, Total.y + Direct.y, Total.z + Direct.z, Total.w + 1.0f);
    if (store) w.Total[i] = Total;
    return true;
}
// averages, clamps and attachment formats of main()
VXD void rf_resolve_pixel(const ReflArgs& a, int px, int py, const int4 cnt, float4 Total, const float4 misc) {
    const size_t pi = (size_t)py * a.width + px;
    float AveragedHitDistance = misc.x / gmax(misc.y, 0.01f);
    const float th = (float)cnt.y;
    Total = make_float4(Total.x / th, Total.y / th, Total.z / th, Total.w / th);
    const float oHit = gclamp(misc.y > 0.01f ? AveragedHitDistance : -1.0f, -10.0f, 200.0f);
    reinterpret_cast<ushort4*>(a.color)[pi] = make_ushort4(float_to_half_bits(gclamp(Total.x, 0.0000001f, 100.0f)), float_to_half_bits(gclamp(Total.y, 0.0000001f, 100.0f)),
                                                           float_to_half_bits(gclamp(Total.z, 0.0000001f, 100.0f)), float_to_half_bits(gclamp(Total.w, 0.0000001f, 100.0f)));
    a.hitdist[pi] = float_to_half_bits(oHit);
    a.emissive[pi] = float_to_unorm8(gclamp(misc.z, 0.0f, 1.0f));
}

template <bool DEFER>
__global__ void __launch_bounds__(256) rf_wf_shade_b_kernel(const __grid_constant__ ReflArgs a, RfWf w, int first) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    float4 Total, misc;
    rf_shade_b_path<DEFER>(a, w, px, py, i, Total, misc, true, first != 0);
}

__global__ void __launch_bounds__(256) rf_wf_resolve_kernel(const __grid_constant__ ReflArgs a, RfWf w) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    const int4 cnt = w.cnt[i];
    if (cnt.w < 0) return;
    rf_resolve_pixel(a, px, py, cnt, w.Total[i], w.misc[i]);
}

// the LAST sample's shade_b fused with resolve (set_option "gi_fuse_final" governs both wavefronts): the totals of a path whose sample was still
// waiting for its shadow ray go from registers into the attachments instead of through Total / misc and back; one launch less.  Same
// arithmetic in the same order, bit-identical.
template <bool DEFER>
__global__ void __launch_bounds__(256) rf_wf_final_kernel(const __grid_constant__ ReflArgs a, RfWf w, int first) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    const int4 cnt = w.cnt[i];
    if (cnt.w < 0) return;
    float4 Total, misc;
    if (!rf_shade_b_path<DEFER>(a, w, px, py, i, Total, misc, false, first != 0)) { Total = w.Total[i]; misc = w.misc[i]; }
    rf_resolve_pixel(a, px, py, cnt, Total, misc);
}

template <typename T>
T* carve(uint8_t*& p, size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 255) / 256 * 256;
    return r;
}

}  // namespace

int vxrt_launch_reflection_trace_wavefront(vxrt_ctx* c, const void* args_blob) {
    const ReflArgs& a = *reinterpret_cast<const ReflArgs*>(args_blob);
    const int rows = a.row1 - a.row0, cols = a.col1 - a.col0;   // the tile rectangle; path state is indexed inside it
    if (rows <= 0 || cols <= 0) return VXRT_OK;
    const size_t n = (size_t)rows * cols;
    const size_t need = n * (16 * 10 + 4 * 3) + 256 * 32;
    if (need > c->wf_cap) {
        if (c->d_wf) VX_CUDA(cudaFree(c->d_wf));
        c->d_wf = nullptr; c->wf_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_wf, need));
        c->wf_cap = need;
    }
    uint8_t* p = (uint8_t*)c->d_wf;
    RfWf w;
    w.P = carve<float4>(p, n); w.I = carve<float4>(p, n); w.Nmap = carve<float4>(p, n);
    w.Total = carve<float4>(p, n); w.misc = carve<float4>(p, n); w.cnt = carve<int4>(p, n); w.rayD = carve<float4>(p, n);
    w.Amb = carve<float4>(p, n); w.Res = carve<float4>(p, n); w.qShadowO = carve<float4>(p, n);
    w.hitT = carve<float>(p, n); w.hitInfo = carve<unsigned>(p, n); w.shadowRes = carve<float>(p, n);
    w.counters = carve<int>(p, 16);

    const dim3 pgrid((cols + 31) / 32, (rows + 7) / 8);
    const int lgrid = trace_queue_grid(n);
    const GridView g = c->grid();
    f3 strong;
    strong.x = a.strong[0]; strong.y = a.strong[1]; strong.z = a.strong[2];
    int max_spp = a.spp < 1 ? 1 : (a.spp > 16 ? 16 : a.spp);
    const bool st = c->stats_on;
    cudaStream_t s = c->stream;
    // without reprojection and without the LPV term the GI enters a sample only where it is accumulated (shade_b / final): see ctx.h refl_defer_gi
    const bool defer = c->refl_defer_gi && !a.reproject && !a.lpv_gi;
    for (int sample = 0; sample < max_spp; ++sample) {
        VX_CUDA(cudaMemsetAsync(w.counters, 0, sizeof(int), s));
        rf_wf_gen_kernel<<<pgrid, 256, 0, s>>>(a, w, sample);
        if (c->trace_caps | c->trace_spill) {
            const ReflRays pol = {w};
            const int rc = launch_trace_capped(c, g, pol, nullptr, n, a.trace_length, c->d_stats);
            if (rc != VXRT_OK) return rc;
            c->launches -= 1;
        } else if (st) rf_wf_trace_kernel<true><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, (int)n, a.trace_length, c->d_stats);
        else rf_wf_trace_kernel<false><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, (int)n, a.trace_length, c->d_stats);
        // the shading reads the GI attachments: on lane 1 of the pass-level concurrency this is where the pass meets the GI (ctx.h)
        if (sample == 0 && c->refl_gi_event && !defer) VX_CUDA(cudaStreamWaitEvent(s, c->refl_gi_event, 0));
        if (defer) rf_wf_shade_a_kernel<false, true><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
        else if (a.lpv_gi) rf_wf_shade_a_kernel<true, false><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
        else rf_wf_shade_a_kernel<false, false><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
        if (c->trace_caps | c->trace_spill) {
            const ReflShadowRays pol = {w, strong};
            const int rc = launch_trace_capped(c, g, pol, w.counters + 0, n, a.shadow_trace_length, c->d_stats);
            if (rc != VXRT_OK) return rc;
            c->launches -= 1;
        } else if (st) rf_wf_trace_shadow_kernel<true><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, strong, a.shadow_trace_length, c->d_stats);
        else rf_wf_trace_shadow_kernel<false><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, strong, a.shadow_trace_length, c->d_stats);
        if (sample == 0 && c->refl_gi_event && defer) VX_CUDA(cudaStreamWaitEvent(s, c->refl_gi_event, 0));   // ... and with `defer` here
        if (sample + 1 < max_spp || !c->gi_fuse_final) {
            if (defer) rf_wf_shade_b_kernel<true><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
            else rf_wf_shade_b_kernel<false><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
        } else {   // the last sample's shade_b + resolve in one pass
            if (defer) rf_wf_final_kernel<true><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
            else rf_wf_final_kernel<false><<<pgrid, 256, 0, s>>>(a, w, sample == 0);
        }
        c->launches += 5;
    }
    if (!c->gi_fuse_final) {
        rf_wf_resolve_kernel<<<pgrid, 256, 0, s>>>(a, w);
        c->launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    return VXRT_OK;
}
