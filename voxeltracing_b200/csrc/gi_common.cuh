// gi_common.cuh — argument block and shading terms of the diffuse GI pass (DiffuseRayTraceFrag.glsl),
// shared by the one-thread-per-pixel kernel (gi.cu) and the wavefront pipeline (gi_wavefront.cu).
#pragma once
#include "shading.cuh"

namespace {

struct GiArgs {
    float inv_view[16], inv_proj[16];
    int width, height, row0, row1, col0, col1;
    int spp, checker_spp, checkerboard, trace_length, shadow_trace_length, frame, frame_mod128, supersample;
    float halton[2];
    float sun[3], moon[3], viewer[3], light_color[3];
    float sun_visibility, gi_sky_strength, diffuse_light_intensity;
    int apply_player_shadow, sun_stronger;
    const uint16_t* g_t; const uint8_t* g_normal; int gw, gh;
    TexArrayDev tex[4];
    TexCubeDev sky;
    const int32_t* block_data;
    const int32_t* blue;
    uint16_t* sh; uint16_t* cocg; uint16_t* utility; uint8_t* aosky;
};

struct GiState {
    int px, py;
    int CurrentBLSample;
};

// SampleBlueNoise2D (:807-820) + cosWeightedRandomHemisphereDirection (:1031-1053)
VXD f3 cos_weighted_hemisphere(const GiArgs& a, GiState& st, f3 n) {
    f2 r;
    r.x = blue_noise_1d(a.blue, st.px, st.py, a.frame_mod128, 1 + st.CurrentBLSample);
    r.y = blue_noise_1d(a.blue, st.px, st.py, a.frame_mod128, 2 + st.CurrentBLSample);
    st.CurrentBLSample += 2;
    float PI2 = 2.0f * VX_PI;
    f3 uu = normalize(cross(n, F3(0.0f, 1.0f, 1.0f)));
    f3 vv = cross(uu, n);
    float ra = sqrtf(r.y);
    float rx = ra * cosf(PI2 * r.x);
    float ry = ra * sinf(PI2 * r.x);
    float rz = sqrtf(1.0f - r.y);
    f3 rr = rx * uu + ry * vv + rz * n;
    return normalize(rr);
}
// InverseSchlick / DiffuseHammon (:1391-1418)
VXD float inverse_schlick(float f0, float VoH) { return 1.0f - gclamp(f0 + (1.0f - f0) * pow5_mul(1.0f - VoH), 0.0f, 1.0f); }
VXD float diffuse_hammon(f3 normal, f3 viewDir, f3 lightDir, float roughness) {
    float nDotL = gmax(dot(normal, lightDir), 0.0f);
    if (nDotL <= 0.0f) return 0.0f;
    float nDotV = gmax(dot(normal, viewDir), 0.0f);
    float lDotV = gmax(dot(lightDir, viewDir), 0.0f);
    f3 halfWay = normalize(viewDir + lightDir);
    float nDotH = gmax(dot(normal, halfWay), 0.0f);
    float facing = lDotV * 0.5f + 0.5f;
    float singleRough = facing * (0.9f - 0.4f * facing) * ((0.5f + nDotH) * (1.0f / gmax(nDotH, 0.02f)));
    float singleSmooth = 1.05f * inverse_schlick(0.0f, nDotL) * inverse_schlick(0.0f, gmax(nDotV, 0.0f));
    float single = gclamp(gmix(singleSmooth, singleRough, roughness) * (1.0f / VX_PI), 0.0f, 1.0f);
    float multi = 0.1159f * roughness;
    return gclamp((multi + single) * nDotL, 0.0f, 1.0f);
}
// RayBoxIntersect (:1274-1288)
VXD bool ray_box_intersect(f3 boxMin, f3 boxMax, f3 r0, f3 rD) {
    f3 inv = F3(1.0f / rD.x, 1.0f / rD.y, 1.0f / rD.z);
    f3 tbot = inv * (boxMin - r0), ttop = inv * (boxMax - r0);
    f3 tmin = F3(gmin(ttop.x, tbot.x), gmin(ttop.y, tbot.y), gmin(ttop.z, tbot.z));
    f3 tmax = F3(gmax(ttop.x, tbot.x), gmax(ttop.y, tbot.y), gmax(ttop.z, tbot.z));
    float t0 = gmax(gmax(tmin.x, tmin.y), gmax(tmin.x, tmin.z));
    float t1 = gmin(gmin(tmax.x, tmax.y), gmin(tmax.x, tmax.z));
    return t1 > gmax(t0, 0.0f);
}
// IrridianceToSH (:766-784)
VXD void irradiance_to_sh(f3 Radiance, f3 Direction, float* o) {
    float Co = Radiance.x - Radiance.z;
    float T = Radiance.z + Co * 0.5f;
    float Cg = Radiance.y - T;
    float Y = gmax(T + Cg * 0.5f, 0.0f);
    float L00 = 0.282095f;
    float L1_1 = 0.488603f * Direction.y, L10 = 0.488603f * Direction.z, L11 = 0.488603f * Direction.x;
    o[0] = gmax(L11 * Y, -100.0f); o[1] = gmax(L1_1 * Y, -100.0f); o[2] = gmax(L10 * Y, -100.0f); o[3] = gmax(L00 * Y, -100.0f);
    o[4] = Co; o[5] = Cg;
}


}  // namespace
