// shade.cu — hit-material fetch (GenerateGBuffer.glsl:351-423, dispatched at Core/Pipeline.cpp:2147-2229)
// and the Cook-Torrance direct term of the colour pass (ColorPassFrag.glsl:394-451, 776, 812-816, 886-899).
// Both are one thread per pixel, HBM/L2 gather bound (4-5 texel fetches from ~100 MB arrays per pixel).
#include "shading.cuh"

namespace {

struct GBufferArgs {
    float inv_view[16], inv_proj[16];
    int width, height, row0, row1, col0, col1;
    int grass[10], cactus[10];
    const float* g_inv_t; const uint8_t* g_normal; const uint8_t* g_block; int gw, gh;
    TexArrayDev tex[4];
    const int32_t* block_data;
    uint16_t* albedo; uint16_t* normal; uint8_t* pbr; uint8_t* texao;
};

// GetTextureIDs (GenerateGBuffer.glsl:522-578)
VXD f4 gbuffer_texture_ids(const GBufferArgs& a, int id, f3 n) {
    f4 d = F4((float)__ldg(a.block_data + id), (float)__ldg(a.block_data + 128 + id), (float)__ldg(a.block_data + 256 + id),
              (float)__ldg(a.block_data + 384 + id));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int* q = k == 0 ? a.grass : a.cactus;
        if (id == q[0]) {
            if (eq3(n, face_normal(4)) || eq3(n, face_normal(5)) || eq3(n, face_normal(0)) || eq3(n, face_normal(1))) { d.x = (float)q[4]; d.y = (float)q[5]; d.z = (float)q[6]; }
            else if (eq3(n, face_normal(2))) { d.x = (float)q[1]; d.y = (float)q[2]; d.z = (float)q[3]; }
            else if (eq3(n, face_normal(3))) { d.x = (float)q[7]; d.y = (float)q[8]; d.z = (float)q[9]; }
        }
    }
    return d;
}

__global__ void __launch_bounds__(256) generate_gbuffer_kernel(const __grid_constant__ GBufferArgs a) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const size_t i = (size_t)py * a.width + px;
    const f2 tc = pixel_uv(px, py, a.width, a.height);
    const int BaseID = iclamp(cvt_floor(att_r8_nearest(a.g_block, a.gw, a.gh, tc) * 255.0f), 0, 127);
    const float Dist = 1.0f / att_r32f_bilinear(a.g_inv_t, a.gw, a.gh, tc);
    const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
    f3 oA, oN; f4 oP; float oAO;
    if (Dist < 0.0f) {
        oA = F3(0.0f); oN = F3(1.0f); oP = F4(0.0f, 0.0f, 0.0f, 0.0f); oAO = 0.0f;
    } else {
        const f3 FlatNormal = normal_from_id(att_r8_nearest(a.g_normal, a.gw, a.gh, tc), F3(1.0f));
        const f4 data = gbuffer_texture_ids(a, BaseID, FlatNormal);
        f2 UV = F2(1.0f, 1.0f), tUV = F2(1.0f, 1.0f);
        f3 T = F3(0.0f), B = F3(0.0f), tT = F3(0.0f), tB = F3(0.0f);
        calculate_vectors(P, FlatNormal, T, B, UV);
        calculate_vectors(P, F3(fabsf(FlatNormal.x), fabsf(FlatNormal.y), fabsf(FlatNormal.z)), tT, tB, tUV);
        UV = F2(1.0f - tUV.x, 1.0f - tUV.y);  // Parallax() with u_POM == false returns FlatUV; then UV = 1 - UV
        f3 nm = xyz(texarray_sample(a.tex[VXRT_TEX_NORMAL], UV.x, UV.y, data.y, 0.0f));
        nm = nm * 2.0f - F3(1.0f);
        nm = mat3_mul(T, B, FlatNormal, nm);
        const f4 PBRMap = texarray_sample(a.tex[VXRT_TEX_PBR], UV.x, UV.y, data.z, 0.0f);
        const float Emissivity = data.w > -0.5f ? texarray_sample(a.tex[VXRT_TEX_EMISSIVE], UV.x, UV.y, data.w, 0.0f).x : 0.0f;
        oN = nm;
        oP = F4(gclamp(PBRMap.x, 0.0f, 1.0f), gclamp(PBRMap.y, 0.0f, 1.0f), gclamp(PBRMap.z, 0.0f, 1.0f), gclamp(Emissivity, 0.0f, 1.0f));
        oAO = gclamp(PBRMap.w, 0.00000001f, 1.0f);
        oA = xyz(texarray_sample(a.tex[VXRT_TEX_ALBEDO], UV.x, UV.y, data.x, 0.0f));
        const float lb = 0.02f;
        oP.w *= (UV.x > lb && UV.x < 1.0f - lb && UV.y > lb && UV.y < 1.0f - lb) ? 1.0f : 0.0f;
    }
    a.albedo[3 * i] = float_to_half_bits(oA.x); a.albedo[3 * i + 1] = float_to_half_bits(oA.y); a.albedo[3 * i + 2] = float_to_half_bits(oA.z);
    a.normal[3 * i] = float_to_half_bits(oN.x); a.normal[3 * i + 1] = float_to_half_bits(oN.y); a.normal[3 * i + 2] = float_to_half_bits(oN.z);
    reinterpret_cast<uchar4*>(a.pbr)[i] = make_uchar4(float_to_unorm8(oP.x), float_to_unorm8(oP.y), float_to_unorm8(oP.z), float_to_unorm8(oP.w));
    a.texao[i] = float_to_unorm8(oAO);
}

struct DirectArgs {
    float inv_view[16], inv_proj[16];
    int width, height, row0, row1, col0, col1;
    float viewer[3], sun[3], moon[3], sun_color[3], moon_color[3];
    float desat; int amplify;
    const float* g_inv_t; int gw, gh;
    const uint16_t* albedo; const uint16_t* normal; const uint8_t* pbr; const uint8_t* texao; int mw, mh;
    const uint8_t* shadow; int sw, sh;
    uint16_t* direct;
};

// FresnelSchlickRoughness (ColorPassFrag.glsl:1206-1210)
VXD f3 fresnel_schlick_roughness(f3 Eye, f3 norm, f3 F0, float roughness) {
    float cosTheta = gclamp(dot(Eye, norm), 0.00001f, 1.0f);
    float pw = pow5_mul(1.0f - cosTheta);
    f3 m = F3(gmax(1.0f - roughness, F0.x), gmax(1.0f - roughness, F0.y), gmax(1.0f - roughness, F0.z));
    return F0 + (m - F0) * pw;
}
// CalculateDirectionalLight (ColorPassFrag.glsl:419-451)
VXD f3 color_directional_light(f3 viewer, f3 world_pos, f3 light_dir, f3 radiance, f3 radiance_s, f3 albedo, f3 normal, f3 pbr, float shadow) {
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
    f3 F = fresnel_schlick_roughness(Lo, normal, F0, pbr.x);
    float D = ndf_ggx(cosLh, pbr.x);
    float G = ga_schlick_ggx(cosLi, cosLo, pbr.x);
    f3 kd = gmix(F3(1.0f) - F, F3(0.0f), pbr.y);
    f3 diffuseBRDF = kd * albedo;
    f3 specularBRDF = (F * D * G) / gmax(Epsilon, 4.0f * cosLi * cosLo);
    specularBRDF = gclamp(specularBRDF, 0.0f, 2.0f);
    f3 Result = (diffuseBRDF * radiance * cosLi) + (specularBRDF * radiance_s * cosLi);
    return gclamp(Result, 0.0f, 2.5f) * gclamp(1.0f - Shadow, 0.0f, 1.0f);
}

__global__ void __launch_bounds__(256) shade_direct_kernel(const __grid_constant__ DirectArgs a) {
    int px, py;
    tile_pixel(px, py, a.row0, a.col0);
    if (px >= a.col1 || py >= a.row1) return;
    const size_t i = (size_t)py * a.width + px;
    const f2 tc = pixel_uv(px, py, a.width, a.height);
    const float Dist = 1.0f / att_r32f_bilinear(a.g_inv_t, a.gw, a.gh, tc);
    const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
    const f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
    f3 out = F3(0.0f);
    if (Dist > 0.0f) {
        float av[3], nv[3];
        att_half_bilinear<3>(a.albedo, a.mw, a.mh, tc, av);
        att_half_bilinear<3>(a.normal, a.mw, a.mh, tc, nv);
        f3 Albedo = F3(av[0], av[1], av[2]), N = F3(nv[0], nv[1], nv[2]);
        const int mi = wrap_repeat(cvt_floor(tc.x * (float)a.mw), a.mw), mj = wrap_repeat(cvt_floor(tc.y * (float)a.mh), a.mh);
        const uchar4 pb = __ldg(reinterpret_cast<const uchar4*>(a.pbr) + ((size_t)mj * a.mw + mi));
        const f3 pbr = F3(unorm8_to_float(pb.x), unorm8_to_float(pb.y), unorm8_to_float(pb.z));
        const float Emissivity = unorm8_to_float(pb.w);
        Albedo = basic_saturation(Albedo, 1.0f - a.desat);
        if (pbr.y >= 0.1f - 0.01f) Albedo = basic_saturation(Albedo, 0.9f);
        if (a.amplify) {
            N.x *= 1.64f; N.z *= 1.85f;
            N = N + F3(1e-4f);
            N = normalize(N);
        }
        float sv[1];
        att_unorm8_bilinear<1>(a.shadow, a.sw, a.sh, tc, sv);
        const float shadow = gclamp(sv[0], 0.0f, 1.0f);
        const f3 viewer = F3(a.viewer[0], a.viewer[1], a.viewer[2]);
        const f3 sun = F3(a.sun[0], a.sun[1], a.sun[2]), moon = F3(a.moon[0], a.moon[1], a.moon[2]);
        const f3 SunColor = F3(a.sun_color[0], a.sun_color[1], a.sun_color[2]), MoonColor = F3(a.moon_color[0], a.moon_color[1], a.moon_color[2]);
        float SunVisibility = gclamp(dot(sun, F3(0.0f, 1.0f, 0.0f)) + 0.05f, 0.0f, 0.1f) * 12.0f;
        SunVisibility = 1.0f - SunVisibility;
        f3 SunDirect = color_directional_light(viewer, P, sun, SunColor, SunColor, Albedo, N, pbr, shadow);
        f3 MoonDirect = color_directional_light(viewer, P, moon, MoonColor, MoonColor, Albedo, N, pbr, shadow);
        const float sv1 = SunVisibility * 1.0f;
        f3 Direct = F3(gmix(SunDirect.x, MoonDirect.x, sv1), gmix(SunDirect.y, MoonDirect.y, sv1), gmix(SunDirect.z, MoonDirect.z, sv1));
        Direct = ((!(Emissivity > 0.05f)) ? 1.0f : 0.0f) * Direct;
        out = gmax(Direct, 0.000001f);
    }
    a.direct[3 * i] = float_to_half_bits(out.x); a.direct[3 * i + 1] = float_to_half_bits(out.y); a.direct[3 * i + 2] = float_to_half_bits(out.z);
}

}  // namespace

int vxrt_launch_generate_gbuffer(vxrt_ctx* c, const vxrt_gbuffer_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GBUF_ALBEDO, p.width, p.height, 6))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GBUF_NORMAL, p.width, p.height, 6))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GBUF_PBR, p.width, p.height, 4))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_GBUF_TEXAO, p.width, p.height, 1))) return rc;
    GBufferArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    for (int i = 0; i < 10; ++i) { a.grass[i] = p.grass_props[i]; a.cactus[i] = p.cactus_props[i]; }
    const Attachment& gi = c->att[VXRT_ATT_INITIAL_INVT];
    a.g_inv_t = (const float*)gi.ptr; a.g_normal = (const uint8_t*)c->att[VXRT_ATT_INITIAL_NORMAL].ptr;
    a.g_block = (const uint8_t*)c->att[VXRT_ATT_INITIAL_BLOCK].ptr; a.gw = gi.width; a.gh = gi.height;
    for (int k = 0; k < 4; ++k) a.tex[k] = c->tex[k];
    a.block_data = c->d_block_data;
    a.albedo = (uint16_t*)c->att[VXRT_ATT_GBUF_ALBEDO].ptr; a.normal = (uint16_t*)c->att[VXRT_ATT_GBUF_NORMAL].ptr;
    a.pbr = (uint8_t*)c->att[VXRT_ATT_GBUF_PBR].ptr; a.texao = (uint8_t*)c->att[VXRT_ATT_GBUF_TEXAO].ptr;
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    generate_gbuffer_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_shade_direct(vxrt_ctx* c, const vxrt_direct_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_DIRECT, p.width, p.height, 6))) return rc;
    DirectArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    for (int i = 0; i < 3; ++i) {
        a.viewer[i] = p.viewer_position[i]; a.sun[i] = p.sun_direction[i]; a.moon[i] = p.moon_direction[i];
        a.sun_color[i] = p.sun_color[i]; a.moon_color[i] = p.moon_color[i];
    }
    a.desat = p.texture_desat_amount; a.amplify = p.amplify_normal_map;
    const Attachment& gi = c->att[VXRT_ATT_INITIAL_INVT];
    a.g_inv_t = (const float*)gi.ptr; a.gw = gi.width; a.gh = gi.height;
    const Attachment& ga = c->att[VXRT_ATT_GBUF_ALBEDO];
    a.albedo = (const uint16_t*)ga.ptr; a.normal = (const uint16_t*)c->att[VXRT_ATT_GBUF_NORMAL].ptr;
    a.pbr = (const uint8_t*)c->att[VXRT_ATT_GBUF_PBR].ptr; a.texao = (const uint8_t*)c->att[VXRT_ATT_GBUF_TEXAO].ptr;
    a.mw = ga.width; a.mh = ga.height;
    const Attachment& sh = c->att[c->shadow_source];
    a.shadow = (const uint8_t*)sh.ptr; a.sw = sh.width; a.sh = sh.height;
    a.direct = (uint16_t*)c->att[VXRT_ATT_DIRECT].ptr;
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    shade_direct_kernel<<<grid, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
