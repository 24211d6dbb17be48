// reflect_common.cuh — argument block and shading terms of the reflection pass (ReflectionTraceFrag.glsl),
// shared by the one-thread-per-pixel kernel (reflect.cu) and the wavefront pipeline (reflect_wavefront.cu).
#pragma once
#include "lpv_sample.cuh"
#include "shading.cuh"

namespace {

struct ReflArgs {
    float inv_view[16], inv_proj[16], proj_view[16];
    int width, height, row0, row1, col0, col1;
    int spp, checkerboard, trace_length, shadow_trace_length, frame, frame_mod128;
    int rough, roughness_bias, temporal, reproject, derive_sh;
    float halton[2];
    float sun[3], moon[3], strong[3], viewer[3];
    float color_mixed[3];
    int grass[10];
    const uint16_t* g_t; const uint8_t* g_normal; int gw, gh;
    const uint16_t* gb_normal; const uint8_t* gb_pbr; int mw, mh;
    const uint16_t* gi_sh; const uint16_t* gi_cocg; const uint8_t* gi_aosky; int iw, ih;
    const uint8_t* shadow; int sw, sh;
    TexArrayDev tex[4];
    TexCubeDev sky;
    const int32_t* block_data;
    const int32_t* blue;
    uint16_t* color; uint16_t* hitdist; uint8_t* emissive;
    // ApproximateGILPV (:673-700)
    int lpv_gi, decoupled_gi, ss_sky_valid, sun_stronger;
    float sky_ambient_g[3];   // SkyAmbientG = texture(u_Skymap, vec3(0, 1, 0)).xyz (:719), evaluated once per pass on the host
    LpvSampleArgs lpv;
};

struct RfState { int px, py, CurrentBLSample; };

// bayer2 .. bayer32 (:21-29): fract(dot(floor(a), vec2(0.5, floor(a).y * 0.75))), each level adding a quarter of the level below at half
// the coordinate
VXD float rf_bayer2(float ax, float ay) {
    ax = floorf(ax); ay = floorf(ay);
    const float d = ax * 0.5f + ay * (ay * 0.75f);
    return d - floorf(d);
}
VXD float rf_bayer4(float ax, float ay) { return rf_bayer2(0.5f * ax, 0.5f * ay) * 0.25f + rf_bayer2(ax, ay); }
VXD float rf_bayer8(float ax, float ay) { return rf_bayer4(0.5f * ax, 0.5f * ay) * 0.25f + rf_bayer2(ax, ay); }
VXD float rf_bayer16(float ax, float ay) { return rf_bayer8(0.5f * ax, 0.5f * ay) * 0.25f + rf_bayer2(ax, ay); }
VXD float rf_bayer32(float ax, float ay) { return rf_bayer16(0.5f * ax, 0.5f * ay) * 0.25f + rf_bayer2(ax, ay); }

// ApproximateGILPV(P, B) (:673-700) for the pixel (px, py); LPVDither as main() sets it (:721-723)
VXD f3 rf_approximate_gi_lpv(const ReflArgs& a, int px, int py, f3 P, f3 B) {
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;   // gl_FragCoord.xy
    const float tf = a.temporal ? 1.0f : 0.0f;
    const float b32 = rf_bayer32(fx + ((float)a.frame * 0.75f) * tf, fy + ((float)a.frame * 0.5f) * tf);
    const f3 dither = F3(b32 / 384.0f, b32 / 128.0f, b32 / 384.0f);
    const f3 LPV = lpv_sample_data(a.lpv, P, dither);
    if (a.decoupled_gi) {
        f3 Sky = F3(a.sky_ambient_g[0], a.sky_ambient_g[1], a.sky_ambient_g[2]);
        const float L = dot(Sky, F3(0.2125f, 0.7154f, 0.0721f));
        Sky = gmix(F3(L), Sky, a.sun_stronger ? 0.3f : 0.6f);
        float ao[2];
        att_unorm8_bilinear<2>(a.gi_aosky, a.iw, a.ih, pixel_uv(px, py, a.width, a.height), ao);
        f3 Skylighting = Sky * (ao[1] * (a.sun_stronger ? 3.5f : 4.0f));
        Skylighting = Skylighting + F3(rf_bayer16(fx, fy) / 512.0f);
        const f3 hi = B + F3(0.075f);
        Skylighting = F3(gclamp(Skylighting.x * 13.0f, 0.0f, hi.x), gclamp(Skylighting.y * 13.0f, 0.0f, hi.y), gclamp(Skylighting.z * 13.0f, 0.0f, hi.z));
        if (!a.ss_sky_valid) Skylighting = B;
        return Skylighting + LPV;
    }
    const f3 BaseAmbient = B * 0.9f;
    return LPV + BaseAmbient;
}

// SampleBlueNoise2D (:606-614)
VXD f2 rf_blue_noise_2d(const ReflArgs& a, RfState& st, int Index) {
    f2 n;
    n.x = blue_noise_1d(a.blue, st.px, st.py, Index, 1 + st.CurrentBLSample);
    n.y = blue_noise_1d(a.blue, st.px, st.py, Index, 2 + st.CurrentBLSample);
    st.CurrentBLSample += 2;
    st.CurrentBLSample = st.CurrentBLSample % 128;
    return n;
}
// ImportanceSampleGGX (:345-365)
VXD f3 importance_sample_ggx(f3 N, float roughness, f2 Xi) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float phi = 2.0f * VX_PI * Xi.x;
    float cosTheta = sqrtf((1.0f - Xi.y) / (1.0f + (alpha2 - 1.0f) * Xi.y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    f3 H = F3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    f3 up = fabsf(N.z) < 0.999f ? F3(0.0f, 0.0f, 1.0f) : F3(1.0f, 0.0f, 0.0f);
    f3 tangent = normalize(cross(up, N));
    f3 bitangent = cross(N, tangent);
    f3 sampleVec = tangent * H.x + bitangent * H.y + N * H.z;
    return normalize(sampleVec);
}
// GetReflectionDirection (:621-644)
VXD f3 get_reflection_direction(const ReflArgs& a, RfState& st, f3 N, float R) {
    R = gmax(R, 0.05f);
    float NearestDot = -100.0f;
    f3 Best = F3(0.0f);
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        f2 Xi = rf_blue_noise_2d(a, st, a.temporal ? a.frame_mod128 : 100);
        Xi = Xi * F2(0.9f, 0.65f);
        f3 H = importance_sample_ggx(N, R, Xi);
        float d = dot(H, N);
        if (d > NearestDot) { Best = H; NearestDot = d; }
    }
    return Best;
}
// SHToIrradianceA (:452-462), SHToIrridiance (:437-449)
VXD f3 sh_to_irradiance_a(f4 shY, f2 CoCg) {
    float Y = gmax(0.0f, 3.544905f * shY.w);
    CoCg = CoCg * (Y * 0.282095f / (shY.w + 1e-6f));
    float T = Y - CoCg.y * 0.5f;
    float G = CoCg.y + T;
    float B = T - CoCg.x * 0.5f;
    float R = B + CoCg.x;
    return F3(gmax(R, 0.0f), gmax(G, 0.0f), gmax(B, 0.0f));
}
VXD f3 sh_to_irradiance(f4 shY, f2 CoCg, f3 v) {
    float x = dot(F3(shY.x, shY.y, shY.z), v);
    float Y = 2.0f * (1.023326f * x + 0.886226f * shY.w);
    Y = gmax(Y, 0.0f);
    CoCg = CoCg * (Y * 0.282095f / (shY.w + 1e-6f));
    float T = Y - CoCg.y * 0.5f;
    float G = CoCg.y + T;
    float B = T - CoCg.x * 0.5f;
    float R = B + CoCg.x;
    return F3(gmax(R, 0.0f), gmax(G, 0.0f), gmax(B, 0.0f));
}
VXD float sq(float x) { return x * x; }
// G_Smith_over_NdotV, SpecularGGX (:410-435), DeriveSpecularFromDiffuseSH (:521-546)
VXD float g_smith_over_ndotv(float roughness, float NdotV, float NdotL) {
    float alpha = sq(roughness);
    float g1 = NdotV * sqrtf(sq(alpha) + (1.0f - sq(alpha)) * sq(NdotL));
    float g2 = NdotL * sqrtf(sq(alpha) + (1.0f - sq(alpha)) * sq(NdotV));
    return 2.0f * NdotL / (g1 + g2);
}
VXD float specular_ggx(f3 V, f3 L, f3 N, float roughness, float NoH_offset) {
    f3 H = normalize(L - V);
    float NoL = gmax(0.0f, dot(N, L));
    float NoV = gmax(0.0f, -dot(N, V));
    float NoH = gclamp(dot(N, H) + NoH_offset, 0.0f, 1.0f);
    if (NoL > 0.0f) {
        float G = g_smith_over_ndotv(roughness, NoV, NoL);
        float alpha = sq(gmax(roughness, 0.02f));
        float D = sq(alpha) / (VX_PI * sq(sq(NoH) * sq(alpha) + (1.0f - sq(NoH))));
        return D * G / 4.0f;
    }
    return 0.0f;
}
VXD f3 derive_specular_from_diffuse_sh(f4 SHy, f3 IndirectDiffuse, f3 Eye, f3 Normal) {
    float Roughness = 0.4f;
    f3 IncomingDir = F3(SHy.x, SHy.y, SHy.z) / SHy.w * (0.282095f / 0.488603f);
    f3 RawSpecularDir = reflect(Eye, Normal);
    float IncomingLen = length(IncomingDir);
    float Directionality = IncomingLen;
    float Scale = 1.0f;
    if (Directionality >= 1.0f) {
        IncomingDir = IncomingDir / IncomingLen;
    } else {
        f3 q = IncomingDir / (IncomingLen + 0.00001f);
        IncomingDir = F3(gmix(RawSpecularDir.x, q.x, Directionality), gmix(RawSpecularDir.y, q.y, Directionality), gmix(RawSpecularDir.z, q.z, Directionality));
        Scale = pow3_mul(Roughness + 1.0f);
    }
    float Sp = specular_ggx(Eye, IncomingDir, Normal, gmax(Roughness, 0.39f), 0.0f);
    f3 Integrated = powf(Sp, 1.2f) * IndirectDiffuse * 18.0f * Scale;
    if (Integrated.x != Integrated.x || isinf(Integrated.x) || Integrated.y != Integrated.y || isinf(Integrated.y) || Integrated.z != Integrated.z || isinf(Integrated.z))
        Integrated = F3(0.0f);
    return gmax(Integrated, 0.00001f);
}
// capIntersect (:1264-1291), GetPlayerIntersect (:1301-1307)
VXD float cap_intersect(f3 ro, f3 rd, f3 pa, f3 pb, float r) {
    f3 ba = pb - pa, oa = ro - pa;
    float baba = dot(ba, ba), bard = dot(ba, rd), baoa = dot(ba, oa), rdoa = dot(rd, oa), oaoa = dot(oa, oa);
    float a = baba - bard * bard;
    float b = baba * rdoa - baoa * bard;
    float cc = baba * oaoa - baoa * baoa - r * r * baba;
    float h = b * b - a * cc;
    if (h >= 0.0f) {
        float t = (-b - sqrtf(h)) / a;
        float y = baoa + t * bard;
        if (y > 0.0f && y < baba) return t;
        f3 oc = (y <= 0.0f) ? oa : ro - pb;
        b = dot(rd, oc);
        cc = dot(oc, oc) - r * r;
        h = b * b - cc;
        if (h > 0.0f) return -b - sqrtf(h);
    }
    return -1.0f;
}
VXD bool get_player_intersect(f3 viewer, f3 WorldPos, f3 d) {
    float x = 0.4f;
    f3 VP = viewer + F3(-x, -x, +x);
    return cap_intersect(WorldPos, d, VP, VP + F3(0.0f, 1.0f, 0.0f), 0.5f) > 0.0f;
}
// the reflection pass' own CalculateDirectionalLight (:313-336)
VXD f3 rf_directional_light(f3 viewer, f3 world_pos, f3 light_dir, f3 radiance, f3 albedo, f3 normal, f3 pbr, float shadow) {
    const float Epsilon = 0.00001f;
    float Shadow = gmin(shadow, 1.0f);
    f3 Lo = normalize(viewer - world_pos);
    f3 N = normal;
    float cosLo = gmax(0.0f, dot(N, Lo));
    f3 F0 = gmix(F3(0.04f), albedo, pbr.y);
    f3 Li = light_dir;
    f3 Lh = normalize(Li + Lo);
    float cosLi = gmax(0.0f, dot(N, Li));
    float cosLh = gmax(0.0f, dot(N, Lh));
    float fc = pow5_mul(1.0f - gmax(0.0f, dot(Lh, Lo)));
    f3 F = F0 + (F3(1.0f) - F0) * fc;
    float D = ndf_ggx(cosLh, pbr.x);
    float G = ga_schlick_ggx(cosLi, cosLo, pbr.x);
    f3 kd = gmix(F3(1.0f) - F, F3(0.0f), pbr.y);
    f3 diffuseBRDF = kd * albedo;
    f3 specularBRDF = (F * D * G) / gmax(Epsilon, 4.0f * cosLi * cosLo);
    f3 radiance_s = radiance * 0.05f * 0.0f;
    f3 Result = (diffuseBRDF * radiance * cosLi) + (specularBRDF * radiance_s * cosLi);
    return gmax(Result, 0.0f) * gclamp(1.0f - Shadow, 0.0f, 1.0f);
}
VXD bool in_thresholded_screen_space(f2 v) {
    float b = 0.032593f;
    return v.x > b && v.x < 1.0f - b && v.y > b && v.y < 1.0f - b;
}

}  // namespace
