// gi_wavefront.cu — diffuse GI as a wavefront pipeline (DiffuseRayTraceFrag.glsl, same arithmetic as
// gi.cu, reorganised for the machine).
//
// The one-thread-per-pixel kernel (gi.cu) spends its time diverged: lanes of a warp sit in different
// phases of CalculateDiffuse (bounce-0 ray, sun-shadow ray, bounce-1 ray, texture shading) and ncu shows
// 12 of 32 lanes active per issued instruction at 25 % occupancy (profiles/r1_b_*).  Here every phase is
// its own kernel over dense, compacted work lists:
//
//   gen      per pixel: reconstruct P / N from the G-buffer, first cosine-weighted direction
//   trace    bounce-0 rays (pixel order: coherent origins)                       <= trace_length iterations
//   shade<0> material fetch, emissive, sun term; enqueue sun-shadow ray + bounce-1 ray (warp-aggregated
//            compaction: survivors of a warp stay contiguous, so the queues keep screen coherence)
//   trace    shadow queue (one direction for all rays), bounce-1 queue
//   shade<1> apply bounce-0 terms with the shadow result; shade bounce-1 hit; enqueue its sun-shadow ray
//   trace    shadow queue
//   shade<2> apply bounce-1 terms; clamp, SH/CoCg encode, accumulate the sample
//   resolve  per pixel: average over SPP, clamps, attachment formats
//
// The lean trace kernels run at the occupancy of the primary pass, the shading kernels run converged.
// Path state lives in HBM as SoA float4 arrays (~150 B per pixel); results are bit-identical to gi.cu.
#include "gi_common.cuh"
#include "trace_queue.cuh"

#ifndef VX_SHADE_OCC
#define VX_SHADE_OCC 5  // minimum CTAs per SM asked of the gen / shade kernels: 48 registers; 4 / 5 / 6 measured, GI 1.122 / 1.089 / 1.090 ms
#endif

namespace {

struct GiWf {
    // per pixel, persistent across samples
    float4* pixP;       // P.xyz, w = face id of the G-buffer normal as float
    int* bl;            // CurrentBLSample counter; -1 = sky pixel (no paths)
    int* spp;           // number of samples this pixel takes
    float4* accSH;      // sum of SH[0..3]
    float4* accRadAO;   // sum of clamped radiance (xyz), sum of ao (w)
    float4* accCoCgSky; // sum CoCg (xy), sum sky hits (z)
    // per sample
    float4* odirAo;     // first-bounce direction (xyz), ao (w)
    float4* contrib;    // RayContribution (xyz), w = 1 while the path still has a ray in flight
    float4* thr;        // RayThroughput (xyz), w = sky-hit flag of the sample
    // After shade<1> the two hold the sample's FINAL contribution instead: the pending bounce-1 terms only wait for one bit, the result of the
    // last sun-shadow ray, so shade<1> applies them for both outcomes with the shader's own expression (gi_apply_pending) -
    // contrib = (lit result, w = 1: pick by shadowRes / 2: shadow known, take this one), thr = (shadowed result, sky-hit flag) - and shade<2> /
    // final select.  32 instead of 80 bytes per path written by shade<1> and read back (Apend, EmmisivityColor and the throughput factor no
    // longer travel), bit-identical: the selected value is the one the old sequence computed.
    float4* A;          // bounce 0 -> 1 only: (Albedo*DiffuseHammon)*(LIGHT_COLOR*3.5) (xyz); w = ShadowAt if known, -1 = from shadowRes
    float4* rayO;       // current bounce ray
    float4* rayD;
    float* hitT;
    unsigned* hitInfo;  // bits 0-7 block id, bits 8-10 face (7 = zero normal)
    float* shadowRes;   // 1 if the sun-shadow ray of the current bounce hit
    float4* qShadowO;   // compacted shadow rays: origin (xyz), path index bits (w)
    int* qBounce;       // compacted path indices whose bounce-1 ray is traced
    int* counters;      // [0] shadow rays, [1] bounce rays
};

// the tail of a bounce iteration once its sun-shadow term is known (:611-614): RayContribution += RayThroughput * SUNBRDF + EmmisivityColor * RayThroughput
VXD f3 gi_apply_pending(f3 contrib, f3 thr, f3 Apend, float ShadowAt, f3 Em) {
    const f3 SUNBRDF = Apend * (1.0f - ShadowAt) * VX_PI;
    contrib = contrib + thr * SUNBRDF;
    return contrib + Em * thr;
}

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
VXD f3 ld3(const float4* p) { float4 v = *p; return F3(v.x, v.y, v.z); }

// ---- trace kernels (trace_queue.cuh) -----------------------------------------------------------------
// closest hit for the path rays; `list` == nullptr walks all paths whose ray is flagged active
struct PathRays {
    GiWf w;
    const int* list;
    VXD bool fetch(int idx, f3& o, f3& d) const {
        const int i = list ? list[idx] : idx;
        const float4 d4 = w.rayD[i];
        if (!list && d4.w == 0.0f) return false;
        const float4 o4 = w.rayO[i];
        o = F3(o4.x, o4.y, o4.z); d = F3(d4.x, d4.y, d4.z);
        return true;
    }
    VXD void store(int idx, const TraceResult& r) const {
        const int i = list ? list[idx] : idx;
        w.hitT[i] = r.t;
        w.hitInfo[i] = pack_hit(r);
    }
};
template <bool STATS>
__global__ void VX_TRACE_BOUNDS wf_trace_paths_kernel(GridView g, GiWf w, const int* __restrict__ list, const int* __restrict__ count_ptr,
                                                             int n, int max_iter, TraceStatsDev* stats) {
    const int count = list ? *count_ptr : n;
    LaneStats ls = {0u, 0u, 0u, 0u};
    PathRays pol = {w, list};
    trace_queue<STATS>(g, pol, count, max_iter, &ls);
    if (STATS) flush_stats(stats, ls);
}
// any hit along the (single) light direction for the compacted shadow queue
struct ShadowRays {
    GiWf w;
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
__global__ void VX_TRACE_BOUNDS wf_trace_shadow_kernel(GridView g, GiWf w, f3 light, int max_iter, TraceStatsDev* stats) {
    const int count = w.counters[0];
    LaneStats ls = {0u, 0u, 0u, 0u};
    ShadowRays pol = {w, light};
    trace_queue<STATS>(g, pol, count, max_iter, &ls);
    if (STATS) flush_stats(stats, ls);
}

// ---- gen: main() prologue per pixel (:910-969) + first direction of sample `sample` ------------------
__global__ void __launch_bounds__(256, VX_SHADE_OCC) gi_wf_gen_kernel(const __grid_constant__ GiArgs a, GiWf w, int sample) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    // sample 0 keeps what it has just computed in registers; the pixel's position is only stored when it takes further samples
    float4 p4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int bl = 0;
    if (sample == 0) {
        const size_t pi = (size_t)py * a.width + px;
        const f2 vtc = pixel_uv(px, py, a.width, a.height);
        f2 tc = vtc;
        if (a.supersample) {
            f2 h = F2(a.halton[0] * 0.75f, a.halton[1] * 0.75f);
            tc = F2(tc.x + h.x / (float)a.width, tc.y + h.y / (float)a.height);
        }
        const float Dist = att_r16f_bilinear(a.g_t, a.gw, a.gh, tc);
        const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
        const f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
        const float nid = att_r8_nearest(a.g_normal, a.gw, a.gh, tc);
        if (Dist < 0.0f) {
            // sky pixel: SH of the sky, no paths (:952-958)
            const f3 Normal = normal_from_id(nid, F3(0.5f));
            const f3 rd = normalize(ray_direction_at(a.inv_view, a.inv_proj, vtc));
            float SH[6];
            irradiance_to_sh(texcube_sample(a.sky, rd) * 2.66f, Normal, SH);
            reinterpret_cast<ushort4*>(a.sh)[pi] = make_ushort4(float_to_half_bits(SH[0]), float_to_half_bits(SH[1]), float_to_half_bits(SH[2]), float_to_half_bits(SH[3]));
            reinterpret_cast<ushort2*>(a.cocg)[pi] = make_ushort2(float_to_half_bits(SH[4]), float_to_half_bits(SH[5]));
            a.utility[pi] = float_to_half_bits(0.0f);
            reinterpret_cast<uchar2*>(a.aosky)[pi] = make_uchar2(float_to_unorm8(1.0f), float_to_unorm8(0.0f));
            w.bl[i] = -1;
            w.rayD[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // no path (shade<0> marks it dead for the later stages)
            return;
        }
        int SPP = iclamp(a.spp, 1, 32);
        if (a.checkerboard) {
            bool CheckerStep = cvt_trunc(((float)px + 0.5f) + ((float)py + 0.5f)) % 2 == a.frame % 2;
            SPP = cvt_trunc(gmix((float)a.spp, (float)a.checker_spp, CheckerStep ? 1.0f : 0.0f));
        }
        SPP = iclamp(SPP, 1, 32);
        if (!a.sun_stronger) SPP *= 2;
        p4 = make_float4(P.x, P.y, P.z, (float)cvt_round(nid * 10.0f));
        if (SPP > 1) w.pixP[i] = p4;
        w.spp[i] = SPP;
        // the accumulators are not cleared here: the first sample's finish_sample writes 0 + x instead of reading them
    } else {
        bl = w.bl[i];
        if (bl < 0 || sample >= w.spp[i]) {
            w.rayD[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            return;
        }
        p4 = w.pixP[i];
    }
    const int face = (int)p4.w;
    const f3 Normal = face > 5 ? F3(0.5f) : face_normal(face);
    GiState st;
    st.px = px; st.py = py; st.CurrentBLSample = bl;
    const f3 d = cos_weighted_hemisphere(a, st, Normal);
    w.bl[i] = st.CurrentBLSample;
    const f3 o = F3(p4.x, p4.y, p4.z) + Normal * 0.06f;
    w.rayO[i] = make_float4(o.x, o.y, o.z, 0.0f);
    w.rayD[i] = make_float4(d.x, d.y, d.z, 1.0f);
    w.odirAo[i] = make_float4(d.x, d.y, d.z, 1.0f);
    // RayContribution = 0, RayThroughput = 1, no sky hit: shade<0> starts from these constants instead of loading them
}

// sample epilogue of main() (:985-1000): clamp, SH encode, accumulate
VXD void finish_sample(const GiWf& w, int i, f3 contrib, float skyhit, bool first) {
    const float4 oa = w.odirAo[i];
    const f3 xc = gclamp(contrib, 0.0f, 8.0f);
    float SH[6];
    irradiance_to_sh(xc, F3(oa.x, oa.y, oa.z), SH);
    const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // 0 + x, not x: keeps -0 -> +0 like the shader's running sum
    const float4 s = first ? zero : w.accSH[i];
    w.accSH[i] = make_float4(s.x + SH[0], s.y + SH[1], s.z + SH[2], s.w + SH[3]);
    const float4 r = first ? zero : w.accRadAO[i];
    w.accRadAO[i] = make_float4(r.x + xc.x, r.y + xc.y, r.z + xc.z, r.w + oa.w);
    const float4 c = first ? zero : w.accCoCgSky[i];
    w.accCoCgSky[i] = make_float4(c.x + SH[4], c.y + SH[5], c.z + skyhit, 0.0f);
}

// ---- shade<BOUNCE>: the body of the bounce loop of CalculateDiffuse (:547-655) ----------------------
// BOUNCE 0/1: consume the closest-hit result of that bounce; BOUNCE 2: only apply the pending bounce-1 terms.
template <int BOUNCE>
__global__ void __launch_bounds__(256, VX_SHADE_OCC) gi_wf_shade_kernel(const __grid_constant__ GiArgs a, GiWf w, int sample) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    const bool inside = px < a.col1 && py < a.row1;
    const int i = inside ? (py - a.row0) * (a.col1 - a.col0) + (px - a.col0) : 0;
    bool push_shadow = false, push_bounce = false;
    f3 shadow_o = F3(0.0f);
    // a path is alive at bounce 0 iff gen gave it a ray (rayD.w), later iff the previous stage left contrib.w set
    float4 c4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t4 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    bool alive = false;
    if (inside) {
        if (BOUNCE == 0) {
            alive = w.rayD[i].w != 0.0f;
            if (!alive) w.contrib[i] = c4;  // dead for shade<1>, shade<2> (the arena may hold a stale flag)
        } else {
            c4 = w.contrib[i];
            alive = c4.w != 0.0f;
            if (alive) t4 = w.thr[i];
        }
    }
    if (alive) {
        f3 contrib = F3(c4.x, c4.y, c4.z);
        f3 thr = F3(t4.x, t4.y, t4.z);
        float skyhit = t4.w;
        bool still_alive = false;
        float alive_flag_out = 1.0f;
        if (BOUNCE == 1) {
            // the tail of the previous iteration: RayContribution += ..., RayThroughput *= ... (:611-614)
            const float4 A4 = w.A[i];
            const float ShadowAt = A4.w >= 0.0f ? A4.w : w.shadowRes[i];
            const f3 SUNBRDF = F3(A4.x, A4.y, A4.z) * (1.0f - ShadowAt) * VX_PI;
            // bounce 0 left contrib = EmmisivityColor and thr = Albedo*Attenuation/PDF (see below): with RayContribution = 0 and
            // RayThroughput = 1 the shader's (0 + 1 * SUNBRDF) + Em * 1 is SUNBRDF + Em and 1 * F is F, bit for bit
            contrib = SUNBRDF + contrib;
        }
        if (BOUNCE == 2) {
            // shade<1> left both outcomes of the last sun-shadow ray (GiWf): c4 = lit, t4 = shadowed
            if (!(c4.w == 2.0f || w.shadowRes[i] == 0.0f)) contrib = thr;
        }
        if (BOUNCE < 2) {
            const f3 light = a.sun_stronger ? F3(a.sun[0], a.sun[1], a.sun[2]) : F3(a.moon[0], a.moon[1], a.moon[2]);
            const f3 LIGHT_COLOR = F3(a.light_color[0], a.light_color[1], a.light_color[2]);
            const f3 rayO = ld3(w.rayO + i), rayD = ld3(w.rayD + i);
            const float T = w.hitT[i];
            const unsigned info = w.hitInfo[i];
            const int block = (int)(info & 0xffu);
            const f3 HitNormal = unpack_normal(info);
            const int tex_ref = iclamp(block, 0, 127);
            const f3 IntersectionPosition = rayO + (rayD * T);
            if (T > 0.0f && block > 0) {
                f2 txc = F2(0.0f, 0.0f);
                calculate_uv(IntersectionPosition, HitNormal, txc);
                const float TexA = (float)__ldg(a.block_data + tex_ref), TexE = (float)__ldg(a.block_data + 384 + tex_ref);
                const f3 Albedo = xyz(texarray_sample(a.tex[VXRT_TEX_ALBEDO], txc.x, txc.y, TexA, 3.0f));
                const f3 PBR = xyz(texarray_sample(a.tex[VXRT_TEX_PBR], txc.x, txc.y, TexA, 2.0f));
                float Emmisivity = 0.0f;
                if (TexE >= 0.0f) {
                    float SampledEmmisivity = texarray_sample(a.tex[VXRT_TEX_EMISSIVE], txc.x, txc.y, TexE, 0.0f).x;
                    Emmisivity = SampledEmmisivity * 12.0f * a.diffuse_light_intensity;
                }
                const float NDotL = gmax(dot(HitNormal, light), 0.0f);
                float ShadowAt = -1.0f;
                if (!a.sun_stronger) ShadowAt = 1.0f;
                else if (NDotL < 0.001f) ShadowAt = 0.0f;
                else {
                    shadow_o = IntersectionPosition + HitNormal * 0.045f;
                    bool player = false;
                    if (a.apply_player_shadow) {
                        const f3 vp = F3(a.viewer[0], a.viewer[1], a.viewer[2]);
                        player = ray_box_intersect(vp + F3(0.2f, 0.0f, 0.2f), vp - F3(0.75f, 1.75f, 0.75f), shadow_o, light);
                    }
                    if (player) ShadowAt = 1.0f; else push_shadow = true;
                }
                const f3 EmmisivityColor = (Emmisivity * gmix(1.0f, 1.0f, a.sun_visibility)) * Albedo;
                const f3 Apend = Albedo * diffuse_hammon(HitNormal, -rayD, light, PBR.x) * (LIGHT_COLOR * 3.5f);
                GiState st;
                st.px = px; st.py = py; st.CurrentBLSample = w.bl[i];
                const f3 NewDirection = cos_weighted_hemisphere(a, st, HitNormal);
                w.bl[i] = st.CurrentBLSample;
                const float CosTheta = gclamp(dot(HitNormal, NewDirection), 0.0f, 1.0f);
                const float PDF = gmax(CosTheta / VX_PI, 0.00001f);
                const f3 Attenuation = F3(1.0f) * diffuse_hammon(HitNormal, -rayD, NewDirection, PBR.x);
                const f3 F = Albedo * Attenuation / PDF;
                float alive_flag = 1.0f;
                if (BOUNCE == 0) {
                    // the pending terms of bounce 0 travel in contrib / thr themselves (16 + 16 bytes per path less to write and to read
                    // back): at this point RayContribution is exactly 0 and RayThroughput exactly 1
                    w.A[i] = make_float4(Apend.x, Apend.y, Apend.z, ShadowAt);
                    contrib = EmmisivityColor;
                    thr = F;
                } else {
                    // last bounce: nothing is left to do but these pending terms, and they wait for one bit.  Both outcomes now (the
                    // throughput factor F only fed a further bounce: RayThroughput is not read after the loop, :655-664)
                    const f3 c_in = contrib, t_in = thr;
                    if (ShadowAt >= 0.0f) {
                        contrib = gi_apply_pending(c_in, t_in, Apend, ShadowAt, EmmisivityColor);
                        thr = contrib;
                        alive_flag = 2.0f;
                    } else {
                        contrib = gi_apply_pending(c_in, t_in, Apend, 0.0f, EmmisivityColor);
                        thr = gi_apply_pending(c_in, t_in, Apend, 1.0f, EmmisivityColor);
                    }
                }
                alive_flag_out = alive_flag;
                if (BOUNCE == 0) {
                    const f3 no = IntersectionPosition + HitNormal * 0.06f;
                    w.rayO[i] = make_float4(no.x, no.y, no.z, 0.0f);
                    w.rayD[i] = make_float4(NewDirection.x, NewDirection.y, NewDirection.z, 1.0f);
                    push_bounce = true;
                    const float dao = 2.0f;
                    if (T < dao && T > 0.0f)   // AO from the first bounce (:637-650): only the scalar is rewritten
                        reinterpret_cast<float*>(w.odirAo + i)[3] = gmax(T / dao, 0.0f);
                }
                still_alive = true;
            } else {
                float x = gmix(1.0f, 1.05f, a.sun_visibility);
                x = gclamp(x * 1.0f * a.gi_sky_strength, 0.0f, 5.0f);
                f3 rd = rayD;
                rd.y = gclamp(rd.y, 0.125f, 1.5f);
                const f3 sky = texcube_sample(a.sky, rd) * x;
                contrib = contrib + sky * thr;
                skyhit = 1.0f;
            }
        }
        if (still_alive) {
            w.contrib[i] = make_float4(contrib.x, contrib.y, contrib.z, alive_flag_out);
            w.thr[i] = make_float4(thr.x, thr.y, thr.z, skyhit);
        } else {
            finish_sample(w, i, contrib, skyhit, sample == 0);
            w.contrib[i] = make_float4(contrib.x, contrib.y, contrib.z, 0.0f);
        }
    }
    // warp-aggregated compaction: one atomic per warp per queue; survivors of a warp stay contiguous
    const unsigned lane = threadIdx.x & 31u;
    if (BOUNCE < 2) {
        const unsigned ms = __ballot_sync(0xffffffffu, push_shadow);
        if (ms) {
            int base = 0;
            if (lane == (unsigned)(__ffs(ms) - 1)) base = atomicAdd(w.counters + 0, __popc(ms));
            base = __shfl_sync(0xffffffffu, base, __ffs(ms) - 1);
            if (push_shadow) w.qShadowO[base + __popc(ms & ((1u << lane) - 1u))] = make_float4(shadow_o.x, shadow_o.y, shadow_o.z, __int_as_float(i));
        }
    }
    if (BOUNCE == 0) {
        const unsigned mb = __ballot_sync(0xffffffffu, push_bounce);
        if (mb) {
            int base = 0;
            if (lane == (unsigned)(__ffs(mb) - 1)) base = atomicAdd(w.counters + 1, __popc(mb));
            base = __shfl_sync(0xffffffffu, base, __ffs(mb) - 1);
            if (push_bounce) w.qBounce[base + __popc(mb & ((1u << lane) - 1u))] = i;
        }
    }
}

// ---- resolve: averages, clamps and attachment formats of main() (:1003-1020) ------------------------
__global__ void __launch_bounds__(256) gi_wf_resolve_kernel(const __grid_constant__ GiArgs a, GiWf w) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    if (w.bl[i] < 0) return;  // sky pixel, written by gen
    const size_t pi = (size_t)py * a.width + px;
    const float n = (float)w.spp[i];
    const float4 s = w.accSH[i], r = w.accRadAO[i], c = w.accCoCgSky[i];
    const float AccumulatedAO = r.w / n;
    const f4 TotalSHy = F4(s.x / n, s.y / n, s.z / n, s.w / n);
    const f2 CoCg = F2(c.x / n, c.y / n);
    const f3 radiance = F3(r.x, r.y, r.z) / n;
    const float Skyhits = c.z / n;
    float oUtil = gmax(dot(radiance, F3(0.299f, 0.587f, 0.114f)), 0.01f);
    oUtil = gclamp(oUtil, 0.001f, 64.0f);
    reinterpret_cast<ushort4*>(a.sh)[pi] = make_ushort4(float_to_half_bits(gclamp(TotalSHy.x, -100.0f, 100.0f)), float_to_half_bits(gclamp(TotalSHy.y, -100.0f, 100.0f)),
                                                        float_to_half_bits(gclamp(TotalSHy.z, -100.0f, 100.0f)), float_to_half_bits(gclamp(TotalSHy.w, -100.0f, 100.0f)));
    reinterpret_cast<ushort2*>(a.cocg)[pi] = make_ushort2(float_to_half_bits(gclamp(CoCg.x, -100.0f, 100.0f)), float_to_half_bits(gclamp(CoCg.y, -100.0f, 100.0f)));
    a.utility[pi] = float_to_half_bits(oUtil);
    reinterpret_cast<uchar2*>(a.aosky)[pi] = make_uchar2(float_to_unorm8(gclamp(AccumulatedAO, 0.0f, 1.0f)), float_to_unorm8(gclamp(Skyhits, 0.0f, 1.0f)));
}

// ---- final: shade<2> of the LAST sample fused with resolve ------------------------------------------------------------------------
// The last sample's pending bounce-1 terms, its epilogue (finish_sample) and the averages / clamps / attachment formats of resolve in
// one pass: a path that is still alive never writes its sample into the accumulators nor reads them back (at 1 spp the accumulators
// of such a pixel are not touched at all), and one launch less.  Same arithmetic in the same order, bit-identical.
__global__ void __launch_bounds__(256, VX_SHADE_OCC) gi_wf_final_kernel(const __grid_constant__ GiArgs a, GiWf w, int sample) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const int i = (py - a.row0) * (a.col1 - a.col0) + (px - a.col0);
    if (w.bl[i] < 0) return;  // sky pixel, written by gen
    float4 s, r, c;
    const float4 c4 = w.contrib[i];
    if (c4.w != 0.0f) {
        const float4 t4 = w.thr[i];
        // shade<1> left both outcomes of the last sun-shadow ray (GiWf): c4 = lit (or the known one), t4 = shadowed
        const f3 contrib = (c4.w == 2.0f || w.shadowRes[i] == 0.0f) ? F3(c4.x, c4.y, c4.z) : F3(t4.x, t4.y, t4.z);
        // finish_sample
        const float4 oa = w.odirAo[i];
        const f3 xc = gclamp(contrib, 0.0f, 8.0f);
        float SH[6];
        irradiance_to_sh(xc, F3(oa.x, oa.y, oa.z), SH);
        const bool first = sample == 0;
        const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float4 s0 = first ? zero : w.accSH[i], r0 = first ? zero : w.accRadAO[i], c0 = first ? zero : w.accCoCgSky[i];
        s = make_float4(s0.x + SH[0], s0.y + SH[1], s0.z + SH[2], s0.w + SH[3]);
        r = make_float4(r0.x + xc.x, r0.y + xc.y, r0.z + xc.z, r0.w + oa.w);
        c = make_float4(c0.x + SH[4], c0.y + SH[5], c0.z + t4.w, 0.0f);
    } else {
        s = w.accSH[i]; r = w.accRadAO[i]; c = w.accCoCgSky[i];
    }
    const size_t pi = (size_t)py * a.width + px;
    const float n = (float)w.spp[i];
    const float AccumulatedAO = r.w / n;
    const f4 TotalSHy = F4(s.x / n, s.y / n, s.z / n, s.w / n);
    const f2 CoCg = F2(c.x / n, c.y / n);
    const f3 radiance = F3(r.x, r.y, r.z) / n;
    const float Skyhits = c.z / n;
    float oUtil = gmax(dot(radiance, F3(0.299f, 0.587f, 0.114f)), 0.01f);
    oUtil = gclamp(oUtil, 0.001f, 64.0f);
    reinterpret_cast<ushort4*>(a.sh)[pi] = make_ushort4(float_to_half_bits(gclamp(TotalSHy.x, -100.0f, 100.0f)), float_to_half_bits(gclamp(TotalSHy.y, -100.0f, 100.0f)),
                                                        float_to_half_bits(gclamp(TotalSHy.z, -100.0f, 100.0f)), float_to_half_bits(gclamp(TotalSHy.w, -100.0f, 100.0f)));
    reinterpret_cast<ushort2*>(a.cocg)[pi] = make_ushort2(float_to_half_bits(gclamp(CoCg.x, -100.0f, 100.0f)), float_to_half_bits(gclamp(CoCg.y, -100.0f, 100.0f)));
    a.utility[pi] = float_to_half_bits(oUtil);
    reinterpret_cast<uchar2*>(a.aosky)[pi] = make_uchar2(float_to_unorm8(gclamp(AccumulatedAO, 0.0f, 1.0f)), float_to_unorm8(gclamp(Skyhits, 0.0f, 1.0f)));
}

template <typename T>
T* carve(uint8_t*& p, size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 255) / 256 * 256;
    return r;
}

}  // namespace

// `a` is the fully populated argument block built by vxrt_launch_diffuse_trace (gi.cu)
int vxrt_launch_diffuse_trace_wavefront(vxrt_ctx* c, const void* args_blob) {
    const GiArgs& a = *reinterpret_cast<const GiArgs*>(args_blob);
    const int rows = a.row1 - a.row0, cols = a.col1 - a.col0;   // the tile rectangle; path state is indexed inside it
    if (rows <= 0 || cols <= 0) return VXRT_OK;
    const size_t n = (size_t)rows * cols;
    const size_t need = n * (16 * 11 + 4 * 6) + 256 * 32;
    if (need > c->wf_cap) {
        if (c->d_wf) VX_CUDA(cudaFree(c->d_wf));
        c->d_wf = nullptr; c->wf_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_wf, need));
        c->wf_cap = need;
    }
    uint8_t* p = (uint8_t*)c->d_wf;
    GiWf w;
    w.pixP = carve<float4>(p, n); w.accSH = carve<float4>(p, n); w.accRadAO = carve<float4>(p, n); w.accCoCgSky = carve<float4>(p, n);
    w.odirAo = carve<float4>(p, n); w.contrib = carve<float4>(p, n); w.thr = carve<float4>(p, n); w.A = carve<float4>(p, n);
    w.rayO = carve<float4>(p, n); w.rayD = carve<float4>(p, n);
    w.qShadowO = carve<float4>(p, n);
    w.bl = carve<int>(p, n); w.spp = carve<int>(p, n); w.hitT = carve<float>(p, n); w.hitInfo = carve<unsigned>(p, n);
    w.shadowRes = carve<float>(p, n); w.qBounce = carve<int>(p, n);
    w.counters = carve<int>(p, 16);

    const dim3 pgrid((cols + 31) / 32, (rows + 7) / 8);
    const int lgrid = trace_queue_grid(n);
    const GridView g = c->grid();
    f3 light;
    light.x = a.sun_stronger ? a.sun[0] : a.moon[0]; light.y = a.sun_stronger ? a.sun[1] : a.moon[1]; light.z = a.sun_stronger ? a.sun[2] : a.moon[2];
    int max_spp = a.spp < 1 ? 1 : (a.spp > 32 ? 32 : a.spp);
    if (a.checkerboard) { int cs = a.checker_spp < 1 ? 1 : (a.checker_spp > 32 ? 32 : a.checker_spp); if (cs > max_spp) max_spp = cs; }
    if (!a.sun_stronger) max_spp *= 2;
    const bool st = c->stats_on;
    cudaStream_t s = c->stream;
    // the path-ray kernel is the probed kernel: its own stats slot, event pairs around each launch when the probe is on
    // (with trace_caps set, a "launch" of the probed kernel is the sequence of its capped passes)
#define TRACE_PATHS(list, cnt, iters)                                                                                     \
    do {                                                                                                                  \
        cudaEvent_t e0 = vxrt_probe_event(c), e1 = vxrt_probe_event(c);                                                   \
        if (e0 && e1) cudaEventRecord(e0, s);                                                                             \
        if (c->trace_caps | c->trace_spill) {                                                                                              \
            const PathRays pol = {w, list};                                                                               \
            const int rc_ = launch_trace_capped(c, g, pol, cnt, n, iters, c->d_stats + 1);                                \
            if (rc_ != VXRT_OK) return rc_;                                                                               \
            c->launches -= 1;                                                                                             \
        } else if (st) wf_trace_paths_kernel<true><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, list, cnt, (int)n, iters, c->d_stats + 1); \
        else wf_trace_paths_kernel<false><<<lgrid, VX_TRACE_CTA, 0, s>>>(g, w, list, cnt, (int)n, iters, c->d_stats + 1);   \
        if (e0 && e1) cudaEventRecord(e1, s);                                                                             \
    } while (0)
#define TRACE_SHADOW(ss)                                                                                                  \
    do {                                                                                                                  \
        if (c->trace_caps | c->trace_spill) {                                                                             \
            const ShadowRays pol = {w, light};                                                                            \
            const int rc_ = launch_trace_capped(c, g, pol, w.counters + 0, n, a.shadow_trace_length, c->d_stats);         \
            if (rc_ != VXRT_OK) return rc_;                                                                               \
            c->launches -= 1;                                                                                             \
        } else if (st) wf_trace_shadow_kernel<true><<<lgrid, VX_TRACE_CTA, 0, ss>>>(g, w, light, a.shadow_trace_length, c->d_stats);      \
        else wf_trace_shadow_kernel<false><<<lgrid, VX_TRACE_CTA, 0, ss>>>(g, w, light, a.shadow_trace_length, c->d_stats);        \
    } while (0)
    // the shadow-queue trace of bounce 0 beside the bounce-ray trace: both read queues shade<0> has written, the first writes shadowRes,
    // the second hitT / hitInfo; shade<1> needs both (the capped / hand-over variants keep their continuation storage on c->stream)
    const bool overlap = c->gi_overlap && !c->probe_on && !(c->trace_caps | c->trace_spill);
    if (overlap && !c->aux_stream) {
        VX_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        VX_CUDA(cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
        VX_CUDA(cudaEventCreateWithFlags(&c->aux_join, cudaEventDisableTiming));
    }
    for (int sample = 0; sample < max_spp; ++sample) {
        VX_CUDA(cudaMemsetAsync(w.counters, 0, 2 * sizeof(int), s));
        gi_wf_gen_kernel<<<pgrid, 256, 0, s>>>(a, w, sample);
        TRACE_PATHS(nullptr, nullptr, a.trace_length);
        gi_wf_shade_kernel<0><<<pgrid, 256, 0, s>>>(a, w, sample);
        if (overlap) {
            VX_CUDA(cudaEventRecord(c->aux_fork, s));
            VX_CUDA(cudaStreamWaitEvent(c->aux_stream, c->aux_fork, 0));
            TRACE_SHADOW(c->aux_stream);
            VX_CUDA(cudaEventRecord(c->aux_join, c->aux_stream));
            TRACE_PATHS(w.qBounce, w.counters + 1, a.trace_length);
            VX_CUDA(cudaStreamWaitEvent(s, c->aux_join, 0));
        } else {
            TRACE_SHADOW(s);
            TRACE_PATHS(w.qBounce, w.counters + 1, a.trace_length);
        }
        VX_CUDA(cudaMemsetAsync(w.counters, 0, sizeof(int), s));
        gi_wf_shade_kernel<1><<<pgrid, 256, 0, s>>>(a, w, sample);
        TRACE_SHADOW(s);
        if (sample + 1 < max_spp || !c->gi_fuse_final) gi_wf_shade_kernel<2><<<pgrid, 256, 0, s>>>(a, w, sample);
        else gi_wf_final_kernel<<<pgrid, 256, 0, s>>>(a, w, sample);   // the last sample's shade<2> + resolve in one pass
        c->launches += 8;
    }
#undef TRACE_PATHS
#undef TRACE_SHADOW
    if (!c->gi_fuse_final) {
        gi_wf_resolve_kernel<<<pgrid, 256, 0, s>>>(a, w);
        c->launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    return VXRT_OK;
}
