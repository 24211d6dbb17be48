/*
 * ref_driver.cpp — host harness that runs the reference's shaders, compiled for the CPU by
 * oracle/build_ref.py, the way Core/Pipeline.cpp / Core/World.cpp dispatch them: set the uniforms,
 * bind the resources, invoke main() once per fragment / compute invocation, collect the outputs in
 * the attachment formats of Core/Pipeline.cpp:1142-1202.   TEST INFRASTRUCTURE ONLY.
 *
 * VXREF_HAVE_<shader> is defined by the build for every shader that compiled through the shim.
 */
#include "glsl_shim.h"

namespace glsl {
thread_local vec4 gl_FragCoord;
thread_local uvec3 gl_GlobalInvocationID;
}

#ifdef VXREF_HAVE_ManhattanDistanceX
#include "ManhattanDistanceX.cpp"
#endif
#ifdef VXREF_HAVE_ManhattanDistanceY
#include "ManhattanDistanceY.cpp"
#endif
#ifdef VXREF_HAVE_ManhattanDistanceZ
#include "ManhattanDistanceZ.cpp"
#endif
#ifdef VXREF_HAVE_InitialRayTraceFrag
#include "InitialRayTraceFrag.cpp"
#endif
#ifdef VXREF_HAVE_ShadowRayTraceFrag
#include "ShadowRayTraceFrag.cpp"
#endif

#ifdef VXREF_HAVE_GenerateGBuffer
#include "GenerateGBuffer.cpp"
#endif
#ifdef VXREF_HAVE_DiffuseRayTraceFrag
#include "DiffuseRayTraceFrag.cpp"
#endif
#ifdef VXREF_HAVE_ReflectionTraceFrag
#include "ReflectionTraceFrag.cpp"
#endif
#ifdef VXREF_HAVE_SVGFTemporal
#include "SVGFTemporal.cpp"
#endif
#ifdef VXREF_HAVE_SVGFVariance
#include "SVGFVariance.cpp"
#endif
#ifdef VXREF_HAVE_SVGFSpatial
#undef sqr   /* a function in SpatialFilter.glsl, a macro in ReflectionTraceFrag.glsl (one translation unit here) */
#include "SVGFSpatial.cpp"
#endif
#ifdef VXREF_HAVE_SVGFPreSpatial
#undef sqr
#include "SVGFPreSpatial.cpp"
#endif
#ifdef VXREF_HAVE_LPVAverageColor
#include "LPVAverageColor.cpp"
#endif
#ifdef VXREF_HAVE_ShadowTemporal
#include "ShadowTemporal.cpp"
#endif
#ifdef VXREF_HAVE_ShadowFilter
#include "ShadowFilter.cpp"
#endif
#ifdef VXREF_HAVE_SpecularTemporal
#include "SpecularTemporal.cpp"
#endif
#ifdef VXREF_HAVE_ReflectionDenoise
#undef sqr
#undef EPS
#include "ReflectionDenoise.cpp"
#endif
#ifdef VXREF_HAVE_ColorPassDirect
#include "ColorPassDirect.cpp"
#endif

#include "../include/vxrt_cuda.h"
#include <omp.h>

using namespace glsl;

extern "C" {

/* bitmask of the shaders this build contains */
int32_t vxref_available(void) {
    int32_t m = 0;
#ifdef VXREF_HAVE_ManhattanDistanceX
    m |= 1;
#endif
#ifdef VXREF_HAVE_ManhattanDistanceY
    m |= 2;
#endif
#ifdef VXREF_HAVE_ManhattanDistanceZ
    m |= 4;
#endif
#ifdef VXREF_HAVE_InitialRayTraceFrag
    m |= 8;
#endif
#ifdef VXREF_HAVE_ShadowRayTraceFrag
    m |= 16;
#endif
#ifdef VXREF_HAVE_GenerateGBuffer
    m |= 32;
#endif
#ifdef VXREF_HAVE_DiffuseRayTraceFrag
    m |= 64;
#endif
#ifdef VXREF_HAVE_ReflectionTraceFrag
    m |= 128;
#endif
#ifdef VXREF_HAVE_ColorPassDirect
    m |= 256;
#endif
#ifdef VXREF_HAVE_SVGFTemporal
    m |= 1024;
#endif
#ifdef VXREF_HAVE_SVGFVariance
    m |= 2048;
#endif
#ifdef VXREF_HAVE_SVGFSpatial
    m |= 4096;
#endif
#ifdef VXREF_HAVE_SVGFPreSpatial
    m |= 131072;
#endif
#ifdef VXREF_HAVE_LPVAverageColor
    m |= 262144;
#endif
#ifdef VXREF_HAVE_ShadowTemporal
    m |= 8192;
#endif
#ifdef VXREF_HAVE_ShadowFilter
    m |= 16384;
#endif
#ifdef VXREF_HAVE_SpecularTemporal
    m |= 32768;
#endif
#ifdef VXREF_HAVE_ReflectionDenoise
    m |= 65536;
#endif
#ifdef VXREF_HAVE_RaycastDetect
    m |= 512;   /* World::RaycastDetect, host C++ lifted from Core/World.cpp (vxref_raycast_detect, generated unit) */
#endif
    return m;
}

#if defined(VXREF_HAVE_ManhattanDistanceX) && defined(VXREF_HAVE_ManhattanDistanceY) && defined(VXREF_HAVE_ManhattanDistanceZ)
/* World::GenerateDistanceField (Core/World.cpp:69-113): the three compute dispatches over the fixed
 * 384x128x384 grid (the shaders hard-code WORLD_SIZE_*).                                           */
void vxref_distance_field(const uint8_t* blocks, uint8_t* df) {
    const int NX = 384, NY = 128, NZ = 384;
    {
        namespace S = shader_ManhattanDistanceX;
        S::u_BlockData.data = blocks; S::u_BlockData.w = NX; S::u_BlockData.h = NY; S::u_BlockData.d = NZ;
        S::o_DistanceBuffer.data = df; S::o_DistanceBuffer.w = NX; S::o_DistanceBuffer.h = NY; S::o_DistanceBuffer.d = NZ;
#pragma omp parallel for collapse(2)
        for (int z = 0; z < NZ; ++z)
            for (int y = 0; y < NY; ++y) { gl_GlobalInvocationID = uvec3(0u, (uint)y, (uint)z); S::shader_reset(); S::shader_main(); }
    }
    {
        namespace S = shader_ManhattanDistanceY;
        S::o_DistanceBuffer.data = df; S::o_DistanceBuffer.w = NX; S::o_DistanceBuffer.h = NY; S::o_DistanceBuffer.d = NZ;
#pragma omp parallel for collapse(2)
        for (int z = 0; z < NZ; ++z)
            for (int x = 0; x < NX; ++x) { gl_GlobalInvocationID = uvec3((uint)x, 0u, (uint)z); S::shader_reset(); S::shader_main(); }
    }
    {
        namespace S = shader_ManhattanDistanceZ;
        S::o_DistanceBuffer.data = df; S::o_DistanceBuffer.w = NX; S::o_DistanceBuffer.h = NY; S::o_DistanceBuffer.d = NZ;
#pragma omp parallel for collapse(2)
        for (int y = 0; y < NY; ++y)
            for (int x = 0; x < NX; ++x) { gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u); S::shader_reset(); S::shader_main(); }
    }
}
#endif

/* ---- scene resources shared by the alpha-tested traversal and the material / GI / reflection shaders ---- */
static struct RefScene {
    const uint8_t* blocks = nullptr; const uint8_t* df = nullptr;
    int32_t block_data[6 * 128];
    std::vector<int32_t> blue;
    vxo::TexArray tex[4];
    std::vector<float> sky; vxo::TexCube cube;
} g_scene;

void vxref_set_world(const uint8_t* blocks, const uint8_t* df) { g_scene.blocks = blocks; g_scene.df = df; }
void vxref_set_block_data(const int32_t* t) { memcpy(g_scene.block_data, t, sizeof(g_scene.block_data)); }
void vxref_set_blue_noise(const int32_t* d, int32_t n) { g_scene.blue.assign(d, d + n); }
void vxref_set_texture_array(int32_t kind, int32_t layers, int32_t w, int32_t h, const uint8_t* rgba) {
    vxo::texarray_build(g_scene.tex[kind], rgba, layers, w, h, kind == 0);
}
void vxref_set_skymap(int32_t res, const float* f) {
    g_scene.sky.assign(f, f + (size_t)6 * res * res * 3);
    g_scene.cube.data = g_scene.sky.data(); g_scene.cube.res = res;
}

static inline void rows_of(const vxrt_tile& t, int H, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = H; } else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > H) *r1 = H; }
}

#ifdef VXREF_HAVE_InitialRayTraceFrag
/* Pipeline.cpp:2051-2094 + InitialRayTraceVert.glsl (v_TexCoords = quad texcoord) */
void vxref_initial_trace(const uint8_t* blocks, const uint8_t* df, const vxrt_primary_params* p, uint16_t* t_half,
                         uint8_t* normal_u8, uint8_t* block_u8, float* inv_t, float* t32) {
    namespace S = shader_InitialRayTraceFrag;
    const int NX = 384, NY = 128, NZ = 384;
    S::u_VoxelDataTexture.data = blocks; S::u_VoxelDataTexture.w = NX; S::u_VoxelDataTexture.h = NY; S::u_VoxelDataTexture.d = NZ;
    S::u_DistanceFieldTexture.data = df; S::u_DistanceFieldTexture.w = NX; S::u_DistanceFieldTexture.h = NY; S::u_DistanceFieldTexture.d = NZ;
    S::u_InverseView.load(p->inv_view);
    S::u_InverseProjection.load(p->inv_projection);
    S::u_Dimensions = vec2((float)p->width, (float)p->height);
    S::u_ShouldAlphaTest = p->alpha_test != 0;   /* Pipeline.cpp:2076 */
    if (p->alpha_test) {
        S::u_AlbedoTextures.t = &g_scene.tex[0];
        S::BlockAlbedoData.data = g_scene.block_data; S::BlockTransparentData.data = g_scene.block_data + 512;
    }
    S::u_RenderDistance = p->render_distance;
    S::u_JitterSceneForTAA = p->jitter_on != 0;
    S::u_CurrentTAAJitter = vec2(p->jitter[0], p->jitter[1]);
    S::u_FOV = p->alpha_test ? p->fov : 90.0f;   /* Pipeline.cpp:2073; only the alpha test reads it */
    S::u_Time = 0.0f;
    S::u_CurrentFrame = 0;
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 2)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::shader_reset(); S::shader_main();
            size_t i = (size_t)py * W + px;
            t_half[i] = vxo::float_to_half(S::o_HitDistance);
            if (t32) t32[i] = S::o_HitDistance;
            normal_u8[i] = vxo::float_to_unorm8(S::o_Normal);
            block_u8[i] = vxo::float_to_unorm8(S::o_BlockID);
            inv_t[i] = S::o_DepthNonLinear;
        }
}
#endif

#ifdef VXREF_HAVE_ShadowRayTraceFrag
/* Pipeline.cpp:2888-2945 + FBOVert.glsl (v_RayOrigin = u_VertInverseView[3]) */
void vxref_shadow_trace(const uint8_t* blocks, const uint8_t* df, const vxrt_shadow_params* p, const uint16_t* g_t_half,
                        const uint8_t* g_normal_u8, int32_t gw, int32_t gh, const uint8_t* blue_rgba, int32_t bw, int32_t bh,
                        uint8_t* shadow_u8, uint16_t* transversal_half) {
    namespace S = shader_ShadowRayTraceFrag;
    const int NX = 384, NY = 128, NZ = 384;
    S::u_VoxelData.data = blocks; S::u_VoxelData.w = NX; S::u_VoxelData.h = NY; S::u_VoxelData.d = NZ;
    S::u_DistanceFieldTexture.data = df; S::u_DistanceFieldTexture.w = NX; S::u_DistanceFieldTexture.h = NY; S::u_DistanceFieldTexture.d = NZ;
    std::vector<float> tf((size_t)gw * gh), nf((size_t)gw * gh), bf((size_t)bw * bh * 4);
    for (size_t i = 0; i < tf.size(); ++i) { tf[i] = vxo::half_to_float(g_t_half[i]); nf[i] = vxo::unorm8_to_float(g_normal_u8[i]); }
    for (size_t i = 0; i < bf.size(); ++i) bf[i] = vxo::unorm8_to_float(blue_rgba[i]);
    S::u_PositionTexture.data = tf.data(); S::u_PositionTexture.w = gw; S::u_PositionTexture.h = gh; S::u_PositionTexture.ch = 1; S::u_PositionTexture.linear = true;   /* R16F LINEAR */
    S::u_NormalTexture.data = nf.data(); S::u_NormalTexture.w = gw; S::u_NormalTexture.h = gh; S::u_NormalTexture.ch = 1; S::u_NormalTexture.linear = false;             /* R8 NEAREST */
    S::u_BlueNoiseTexture.data = bf.data(); S::u_BlueNoiseTexture.w = bw; S::u_BlueNoiseTexture.h = bh; S::u_BlueNoiseTexture.ch = 4; S::u_BlueNoiseTexture.linear = false;
    S::u_LightDirection = vec3(p->light_direction[0], p->light_direction[1], p->light_direction[2]);
    S::u_InverseView.load(p->inv_view);
    S::u_InverseProjection.load(p->inv_projection);
    S::u_CurrentFrame = p->current_frame;
    S::u_ContactHardeningShadows = p->soft_shadows != 0;
    S::u_ShouldAlphaTest = p->alpha_test != 0;   /* Pipeline.cpp:2913 */
    if (p->alpha_test) {
        S::u_AlbedoTextures.t = &g_scene.tex[0];
        S::BlockAlbedoData.data = g_scene.block_data; S::BlockTransparentData.data = g_scene.block_data + 512;
    }
    S::u_Dimensions = vec2((float)p->width, (float)p->height);
    S::u_Halton = vec2(p->halton[0], p->halton[1]);
    S::u_Time = 0.0f;
    S::u_FOV = p->alpha_test ? p->fov : 90.0f;   /* Pipeline.cpp:2914 */
    S::u_DoFullTrace = true;
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
#pragma omp parallel for schedule(dynamic, 2)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam;
            S::shader_reset(); S::shader_main();
            size_t i = (size_t)py * W + px;
            shadow_u8[i] = vxo::float_to_unorm8(S::o_Shadow);
            transversal_half[i] = vxo::float_to_half(S::o_IntersectionTransversal);
        }
}
#endif


static inline void bind3d(sampler3D& s, const uint8_t* d) { s.data = d; s.w = 384; s.h = 128; s.d = 384; }
static inline void bind2d(sampler2D& s, const float* d, int w, int h, int ch, bool linear) { s.data = d; s.w = w; s.h = h; s.ch = ch; s.linear = linear; }
static std::vector<float> h2f(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = vxo::half_to_float(h[i]); return o; }
static std::vector<float> u2f(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = vxo::unorm8_to_float(h[i]); return o; }
#define BIND_BLOCK_SSBO(S)                                                                                    \
    S::BlockAlbedoData.data = g_scene.block_data; S::BlockNormalData.data = g_scene.block_data + 128;           \
    S::BlockPBRData.data = g_scene.block_data + 256; S::BlockEmissiveData.data = g_scene.block_data + 384;      \
    S::BlockTransparentData.data = g_scene.block_data + 512;
#define BIND_BLUE_SSBO(S)                                                                                     \
    S::sobol_256spp_256d.data = g_scene.blue.data(); S::scramblingTile.data = g_scene.blue.data() + 65536;      \
    S::rankingTile.data = g_scene.blue.data() + 65536 + 131072;

#ifdef VXREF_HAVE_GenerateGBuffer
/* Pipeline.cpp:2147-2229 */
void vxref_generate_gbuffer(const vxrt_gbuffer_params* p, const float* g_inv_t, const uint8_t* g_normal, const uint8_t* g_block,
                            int32_t gw, int32_t gh, uint16_t* albedo_h3, uint16_t* normal_h3, uint8_t* pbr_u8x4, uint8_t* texao_u8) {
    namespace S = shader_GenerateGBuffer;
    std::vector<float> nf = u2f(g_normal, (size_t)gw * gh), bf = u2f(g_block, (size_t)gw * gh);
    bind2d(S::u_NonLinearDepth, g_inv_t, gw, gh, 1, true);
    bind2d(S::u_Normals, nf.data(), gw, gh, 1, false);
    bind2d(S::u_BlockIDs, bf.data(), gw, gh, 1, false);
    S::u_BlockAlbedos.t = &g_scene.tex[0]; S::u_BlockNormals.t = &g_scene.tex[1]; S::u_BlockPBR.t = &g_scene.tex[2]; S::u_BlockEmissive.t = &g_scene.tex[3];
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    for (int i = 0; i < 10; ++i) { S::u_GrassBlockProps[i] = p->grass_props[i]; S::u_CactusBlockProps[i] = p->cactus_props[i]; }
    S::u_LavaBlockID = -1;  /* lava's animated 3-D textures are out of scope */
    S::u_Time = 0.0f; S::uTime = 0.0f; S::u_UpdateGBufferThisFrame = true; S::u_Frame = 0;
    S::u_POM = false; S::u_HighQualityPOM = false; S::u_DitherPOM = false; S::u_POMHeight = 1.0f; S::u_POMExp = 1.0f;
    BIND_BLOCK_SSBO(S)
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::shader_reset(); S::shader_main();
            size_t i = (size_t)py * W + px;
            for (int c = 0; c < 3; ++c) { albedo_h3[3 * i + c] = vxo::float_to_half(S::o_Albedo[c]); normal_h3[3 * i + c] = vxo::float_to_half(S::o_Normal[c]); }
            for (int c = 0; c < 4; ++c) pbr_u8x4[4 * i + c] = vxo::float_to_unorm8(S::o_PBR[c]);
            texao_u8[i] = vxo::float_to_unorm8(S::o_TextureAO);
        }
}
#endif

#ifdef VXREF_HAVE_DiffuseRayTraceFrag
/* Pipeline.cpp:2267-2374 + FBOVert.glsl */
void vxref_diffuse_trace(const vxrt_gi_params* p, const uint16_t* g_t_half, const uint8_t* g_normal, int32_t gw, int32_t gh,
                         uint16_t* sh_h4, uint16_t* cocg_h2, uint16_t* utility_h, uint8_t* aosky_u8x2) {
    namespace S = shader_DiffuseRayTraceFrag;
    std::vector<float> tf = h2f(g_t_half, (size_t)gw * gh), nf = u2f(g_normal, (size_t)gw * gh);
    bind3d(S::u_VoxelData, g_scene.blocks); bind3d(S::u_DistanceFieldTexture, g_scene.df);
    bind2d(S::u_PositionTexture, tf.data(), gw, gh, 1, true);
    bind2d(S::u_NormalTexture, nf.data(), gw, gh, 1, false);
    S::u_Skymap = g_scene.cube;
    S::u_BlockNormalTextures.t = &g_scene.tex[1]; S::u_BlockAlbedoTextures.t = &g_scene.tex[0];
    S::u_BlockPBRTextures.t = &g_scene.tex[2]; S::u_BlockEmissiveTextures.t = &g_scene.tex[3];
    S::CHECKERBOARD_SPP = p->checkerboard != 0;
    S::u_Dimensions = vec2((float)p->width, (float)p->height);
    S::u_Halton = vec2(p->halton[0], p->halton[1]);
    S::u_Time = 0.0f; S::u_Supersample = p->supersample != 0;
    S::u_APPLY_PLAYER_SHADOW = p->apply_player_shadow != 0;
    S::u_UseDirectSampling = false;
    S::u_SPP = p->spp; S::u_DiffuseTraceLength = p->trace_length; S::u_CheckerSPP = p->checker_spp;
    S::u_CurrentFrame = p->current_frame; S::u_CurrentFrameMod512 = p->current_frame % 512; S::u_CurrentFrameMod128 = p->current_frame_mod128;
    S::u_UseBlueNoise = p->use_blue_noise != 0;
    S::u_GISunStrength = p->gi_sun_strength; S::u_GISkyStrength = p->gi_sky_strength; S::u_SunVisibility = p->sun_visibility;
    S::u_ViewerPosition = vec3(p->viewer_position[0], p->viewer_position[1], p->viewer_position[2]);
    S::u_SunDirection = vec3(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]);
    S::u_MoonDirection = vec3(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    S::u_DiffuseLightIntensity = p->diffuse_light_intensity;
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    BIND_BLOCK_SSBO(S)
    BIND_BLUE_SSBO(S)
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
#pragma omp parallel for schedule(dynamic, 2)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam;
            S::v_RayDirection = S::GetRayDirectionAt(S::v_TexCoords);  /* FBOVert.glsl:15-20, interpolated */
            S::shader_reset(); S::shader_main();
            size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) sh_h4[4 * i + c] = vxo::float_to_half(S::o_SH[c]);
            cocg_h2[2 * i] = vxo::float_to_half(S::o_CoCg.x); cocg_h2[2 * i + 1] = vxo::float_to_half(S::o_CoCg.y);
            utility_h[i] = vxo::float_to_half(S::o_Utility);
            aosky_u8x2[2 * i] = vxo::float_to_unorm8(S::o_AOAndSkyLighting.x); aosky_u8x2[2 * i + 1] = vxo::float_to_unorm8(S::o_AOAndSkyLighting.y);
        }
}
#endif


#ifdef VXREF_HAVE_ColorPassDirect
/* The direct-lighting term of ColorPassFrag.glsl main(): the BRDF functions are the reference's own
 * (cut out of the file by name); the call site below restates :776, :802-816, :886-899.             */
void vxref_shade_direct(const vxrt_direct_params* p, const float* g_inv_t, int32_t gw, int32_t gh, const uint16_t* albedo_h3,
                        const uint16_t* normal_h3, const uint8_t* pbr_u8x4, const uint8_t* texao_u8, int32_t mw, int32_t mh,
                        const uint8_t* shadow_u8, int32_t sw, int32_t sh, uint16_t* direct_h3) {
    namespace S = shader_ColorPassDirect;
    std::vector<float> af = h2f(albedo_h3, (size_t)mw * mh * 3), nf = h2f(normal_h3, (size_t)mw * mh * 3);
    std::vector<float> pf = u2f(pbr_u8x4, (size_t)mw * mh * 4), sf = u2f(shadow_u8, (size_t)sw * sh);
    sampler2D tInvT, tA, tN, tP, tS;
    bind2d(tInvT, g_inv_t, gw, gh, 1, true);
    bind2d(tA, af.data(), mw, mh, 3, true); bind2d(tN, nf.data(), mw, mh, 3, true);
    bind2d(tP, pf.data(), mw, mh, 4, false); bind2d(tS, sf.data(), sw, sh, 1, true);
    S::u_ViewerPosition = vec3(p->viewer_position[0], p->viewer_position[1], p->viewer_position[2]);
    mat4 inv_view, inv_proj;
    inv_view.load(p->inv_view); inv_proj.load(p->inv_projection);
    const vec3 u_SunDirection(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]);
    const vec3 u_MoonDirection(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    const vec3 SunColor(p->sun_color[0], p->sun_color[1], p->sun_color[2]), MoonColor(p->moon_color[0], p->moon_color[1], p->moon_color[2]);
    float SunVisibility = clamp(dot(u_SunDirection, vec3(0.0f, 1.0f, 0.0f)) + 0.05f, 0.0f, 0.1f) * 12.0f; SunVisibility = 1.0f - SunVisibility;
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            vec2 g_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            vec4 clip = vec4(g_TexCoords * 2.0f - 1.0f, -1.0f, 1.0f);
            vec4 eye = vec4(vec2(inv_proj * clip), -1.0f, 0.0f);
            vec3 rd = vec3(inv_view * eye);
            float Dist = 1.0f / texture(tInvT, g_TexCoords).r;
            vec4 WorldPosition = vec4(vec3(inv_view[3]) + normalize(rd) * Dist, Dist);
            vec3 o_Direct = vec3(0.0f);
            if (WorldPosition.w > 0.0f) {
                vec3 AlbedoColor = vec3(texture(tA, g_TexCoords));
                vec3 NormalMapped = vec3(texture(tN, g_TexCoords));
                vec4 PBRMap = texture(tP, g_TexCoords);
                float Emissivity = PBRMap.w;
                AlbedoColor = S::BasicSaturation(AlbedoColor, 1.0f - p->texture_desat_amount);
                if (PBRMap.y >= 0.1f - 0.01f) AlbedoColor = S::BasicSaturation(AlbedoColor, 0.9f);
                if (p->amplify_normal_map) {
                    NormalMapped.x *= 1.64f; NormalMapped.z *= 1.85f; NormalMapped += 1e-4f; NormalMapped = normalize(NormalMapped);
                }
                float RayTracedShadow = clamp(texture(tS, g_TexCoords).r, 0.0f, 1.0f);
                vec3 SunDirectLighting = S::CalculateDirectionalLight(vec3(WorldPosition), u_SunDirection, SunColor, SunColor, AlbedoColor, NormalMapped, vec3(PBRMap), RayTracedShadow);
                vec3 MoonDirectLighting = S::CalculateDirectionalLight(vec3(WorldPosition), u_MoonDirection, MoonColor, MoonColor, AlbedoColor, NormalMapped, vec3(PBRMap), RayTracedShadow);
                vec3 DirectLighting = mix(SunDirectLighting, MoonDirectLighting, SunVisibility * vec3(1.0f));
                DirectLighting = (float(!(Emissivity > 0.05f)) * DirectLighting);
                DirectLighting = max(DirectLighting, 0.000001f);
                o_Direct = DirectLighting;
            }
            size_t i = (size_t)py * W + px;
            for (int c = 0; c < 3; ++c) direct_h3[3 * i + c] = vxo::float_to_half(o_Direct[c]);
        }
}
#endif


#ifdef VXREF_HAVE_ReflectionTraceFrag
/* Pipeline.cpp:3096-3257 + FBOVert.glsl.  LPV, projected clouds, player reflection and lava are off. */
struct vxref_reflection_inputs {
    const uint16_t* g_t_half; const uint8_t* g_normal; int32_t gw, gh;
    const uint16_t* gb_normal_h3; const uint8_t* gb_pbr_u8x4; int32_t mw, mh;
    const uint16_t* gi_sh_h4; const uint16_t* gi_cocg_h2; const uint8_t* gi_aosky_u8x2; int32_t iw, ih;
    const uint8_t* shadow_u8; int32_t sw, sh;
};
static const uint8_t* g_lpv_level = nullptr;
static const uint8_t* g_lpv_type = nullptr;
static const float* g_lpv_avg512 = nullptr;
/* the propagation volume + average block colours the reflection shader binds when u_LPVGI is on (borrowed) */
void vxref_set_lpv(const uint8_t* level, const uint8_t* block_type, const float* avg512) {
    g_lpv_level = level; g_lpv_type = block_type; g_lpv_avg512 = avg512;
}
void vxref_reflection_trace(const vxrt_reflection_params* p, const vxref_reflection_inputs* in, uint16_t* color_h4, uint16_t* hitdist_h,
                            uint8_t* emissive_u8) {
    namespace S = shader_ReflectionTraceFrag;
    std::vector<float> tf = h2f(in->g_t_half, (size_t)in->gw * in->gh), nf = u2f(in->g_normal, (size_t)in->gw * in->gh);
    std::vector<float> gnf = h2f(in->gb_normal_h3, (size_t)in->mw * in->mh * 3), gpf = u2f(in->gb_pbr_u8x4, (size_t)in->mw * in->mh * 4);
    std::vector<float> shf = h2f(in->gi_sh_h4, (size_t)in->iw * in->ih * 4), ccf = h2f(in->gi_cocg_h2, (size_t)in->iw * in->ih * 2);
    std::vector<float> aof = u2f(in->gi_aosky_u8x2, (size_t)in->iw * in->ih * 2), sf = u2f(in->shadow_u8, (size_t)in->sw * in->sh);
    bind3d(S::u_VoxelData, g_scene.blocks); bind3d(S::u_DistanceFieldTexture, g_scene.df);
    bind2d(S::u_PositionTexture, tf.data(), in->gw, in->gh, 1, true);
    bind2d(S::u_InitialTraceNormalTexture, nf.data(), in->gw, in->gh, 1, false);
    bind2d(S::u_GBufferNormals, gnf.data(), in->mw, in->mh, 3, true);
    bind2d(S::u_GBufferPBR, gpf.data(), in->mw, in->mh, 4, false);
    bind2d(S::u_DiffuseSH, shf.data(), in->iw, in->ih, 4, true);
    bind2d(S::u_DiffuseCoCg, ccf.data(), in->iw, in->ih, 2, true);
    bind2d(S::u_IndirectAO, aof.data(), in->iw, in->ih, 2, true);
    bind2d(S::u_ShadowTrace, sf.data(), in->sw, in->sh, 1, true);
    S::u_Skymap = g_scene.cube;
    S::u_BlockNormalTextures.t = &g_scene.tex[1]; S::u_BlockAlbedoTextures.t = &g_scene.tex[0];
    S::u_BlockPBRTextures.t = &g_scene.tex[2]; S::u_BlockEmissiveTextures.t = &g_scene.tex[3];
    S::u_SunStrengthModifier = p->sun_strength_modifier; S::u_MoonStrengthModifier = p->moon_strength_modifier;
    S::u_CloudReflections = false; S::u_TemporalFilterReflections = p->temporal != 0; S::TEMPORAL_SPEC = p->temporal != 0;
    S::u_RoughReflections = p->rough_reflections != 0; S::CHECKERBOARD_SPEC_SPP = p->checkerboard != 0;
    S::u_ScreenSpaceSkylightingValid = false; S::u_ReflectionTraceRes = 1.0f;
    S::u_Dimensions = vec2((float)p->width, (float)p->height);
    S::u_SunDirection = vec3(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]);
    S::u_MoonDirection = vec3(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    S::u_StrongerLightDirection = vec3(p->stronger_light_direction[0], p->stronger_light_direction[1], p->stronger_light_direction[2]);
    S::u_Time = 0.0f;
    for (int i = 0; i < 10; ++i) S::u_GrassBlockProps[i] = p->grass_props[i];
    S::u_ViewerPosition = vec3(p->viewer_position[0], p->viewer_position[1], p->viewer_position[2]);
    S::u_SPP = p->spp; S::u_LavaBlockID = -1; S::u_ReflectionTraceLength = p->trace_length;
    S::u_CurrentFrame = p->current_frame; S::u_CurrentFrameMod128 = p->current_frame_mod128;
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_View.load(p->view); S::u_Projection.load(p->projection);
    S::u_ReprojectToScreenSpace = p->reproject_to_screen_space != 0; S::u_UseBlueNoise = p->use_blue_noise != 0;
    S::u_ReflectPlayer = false; S::u_DeriveFromDiffuseSH = p->derive_from_diffuse_sh != 0;
    S::u_LPVGI = p->lpv_gi != 0; S::u_QualityLPVGI = false; S::u_Halton = vec2(p->halton[0], p->halton[1]);
    S::u_RoughnessBias = p->roughness_bias != 0; S::u_UseDecoupledGI = p->use_decoupled_gi != 0;
    S::u_ScreenSpaceSkylightingValid = p->screen_space_skylighting_valid != 0;
    if (p->lpv_gi) {   /* Pipeline.cpp:3115-3116,3240-3251: u_LPV (R8, LINEAR), u_LPVBlocks (R8UI, NEAREST), BlockAverageColorData (binding 4) */
        S::u_LPV.data = g_lpv_level; S::u_LPV.w = 384; S::u_LPV.h = 128; S::u_LPV.d = 384;
        S::u_LPVBlocks.data = g_lpv_type; S::u_LPVBlocks.w = 384; S::u_LPVBlocks.h = 128; S::u_LPVBlocks.d = 384;
        S::BlockAverageColorData.data = reinterpret_cast<const vec4*>(g_lpv_avg512);
    }
    BIND_BLOCK_SSBO(S)
    BIND_BLUE_SSBO(S)
    const int W = p->width, H = p->height;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
#pragma omp parallel for schedule(dynamic, 2)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam;
            S::v_RayDirection = S::GetRayDirectionAt(S::v_TexCoords);
            S::shader_reset(); S::shader_main();
            size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) color_h4[4 * i + c] = vxo::float_to_half(S::o_Color[c]);
            hitdist_h[i] = vxo::float_to_half(S::o_HitDistance);
            emissive_u8[i] = vxo::float_to_unorm8(S::o_EmissivityHitMask);
        }
}
#endif

}  // extern "C"


/* ---- SVGF chain of the diffuse GI (Core/Pipeline.cpp:2428-2700) ---- */
struct vxref_svgf_set { const uint16_t* sh; const uint16_t* cocg; const uint16_t* x; const uint8_t* aosky; };
struct vxref_svgf_out { uint16_t* sh; uint16_t* cocg; uint16_t* x; uint8_t* aosky; };
static std::vector<float> svgf_half(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = vxo::half_to_float(h[i]); return o; }
static std::vector<float> svgf_u8(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = vxo::unorm8_to_float(h[i]); return o; }

#ifdef VXREF_HAVE_SVGFTemporal
extern "C" void vxref_svgf_temporal(const vxrt_svgf_temporal_params* p, const vxref_svgf_set* cur, const vxref_svgf_set* hist,
                                    const uint16_t* g_t, const uint8_t* g_normal, const uint8_t* g_block, const uint16_t* prev_t,
                                    const uint8_t* prev_normal, const uint8_t* prev_block, const vxref_svgf_out* out) {
    namespace S = shader_SVGFTemporal;
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = svgf_half(cur->sh, 4 * n), fcc = svgf_half(cur->cocg, 2 * n), flum = svgf_half(cur->x, n), fao = svgf_u8(cur->aosky, 2 * n);
    auto hsh = svgf_half(hist->sh, 4 * n), hcc = svgf_half(hist->cocg, 2 * n), hut = svgf_half(hist->x, 3 * n), hao = svgf_u8(hist->aosky, 2 * n);
    auto ft = svgf_half(g_t, n), fn = svgf_u8(g_normal, n), fb = svgf_u8(g_block, n);
    auto pt = svgf_half(prev_t, n), pn = svgf_u8(prev_normal, n), pb = svgf_u8(prev_block, n);
    bind2d(S::u_CurrentSH, fsh.data(), W, H, 4, true); bind2d(S::u_CurrentCoCg, fcc.data(), W, H, 2, true);
    bind2d(S::u_NoisyLuminosity, flum.data(), W, H, 1, true); bind2d(S::u_CurrentAO, fao.data(), W, H, 2, true);
    bind2d(S::u_PreviousSH, hsh.data(), W, H, 4, true); bind2d(S::u_PrevCoCg, hcc.data(), W, H, 2, true);
    bind2d(S::u_PreviousUtility, hut.data(), W, H, 3, true); bind2d(S::u_PreviousAO, hao.data(), W, H, 2, true);
    bind2d(S::u_CurrentPositionTexture, ft.data(), W, H, 1, true); bind2d(S::u_CurrentNormalTexture, fn.data(), W, H, 1, false);
    bind2d(S::u_CurrentBlockIDTexture, fb.data(), W, H, 1, false);
    bind2d(S::u_PreviousPositionTexture, pt.data(), W, H, 1, true); bind2d(S::u_PreviousNormalTexture, pn.data(), W, H, 1, false);
    bind2d(S::u_PrevBlockIDTexture, pb.data(), W, H, 1, false);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_PrevView.load(p->prev_view); S::u_PrevProjection.load(p->prev_projection);
    S::u_MinimumMix = 0.0f; S::u_MaximumMix = 0.96f;   /* Pipeline.cpp:2464-2465 (unused by the shader body) */
    S::u_BeUseful = p->be_useful != 0;
    S::u_Time = 0.0f; S::u_DeltaTime = 0.0f;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam;
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out->sh[4 * i + c] = vxo::float_to_half(S::o_SH[c]);
            for (int c = 0; c < 2; ++c) out->cocg[2 * i + c] = vxo::float_to_half(S::o_CoCg[c]);
            for (int c = 0; c < 3; ++c) out->x[3 * i + c] = vxo::float_to_half(S::o_Utility[c]);
            for (int c = 0; c < 2; ++c) out->aosky[2 * i + c] = vxo::float_to_unorm8(S::o_AOAndSkyLighting[c]);
        }
}
#endif

#ifdef VXREF_HAVE_SVGFVariance
extern "C" void vxref_svgf_variance(const vxrt_svgf_variance_params* p, const vxref_svgf_set* in, const uint16_t* g_t, const uint8_t* g_normal,
                                    const vxref_svgf_out* out) {
    namespace S = shader_SVGFVariance;
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = svgf_half(in->sh, 4 * n), fcc = svgf_half(in->cocg, 2 * n), fut = svgf_half(in->x, 3 * n);
    auto ft = svgf_half(g_t, n), fn = svgf_u8(g_normal, n);
    bind2d(S::u_PositionTexture, ft.data(), W, H, 1, true); bind2d(S::u_NormalTexture, fn.data(), W, H, 1, false);
    bind2d(S::u_SH, fsh.data(), W, H, 4, true); bind2d(S::u_CoCg, fcc.data(), W, H, 2, true); bind2d(S::u_Utility, fut.data(), W, H, 3, true);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::DO_SPATIAL = p->do_spatial != 0; S::AGGRESSIVE_DISOCCLUSION_HANDLING = p->aggressive_disocclusion != 0;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam;
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out->sh[4 * i + c] = vxo::float_to_half(S::o_SH[c]);
            for (int c = 0; c < 2; ++c) out->cocg[2 * i + c] = vxo::float_to_half(S::o_CoCg[c]);
            out->x[i] = vxo::float_to_half(S::o_Variance);
        }
}
#endif

#ifdef VXREF_HAVE_SVGFSpatial
extern "C" void vxref_svgf_spatial(const vxrt_svgf_spatial_params* p, const vxref_svgf_set* in, const uint16_t* temporal_utility, const uint16_t* g_t,
                                   const uint8_t* g_normal, const vxref_svgf_out* out) {
    namespace S = shader_SVGFSpatial;
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = svgf_half(in->sh, 4 * n), fcc = svgf_half(in->cocg, 2 * n), fvar = svgf_half(in->x, n), fao = svgf_u8(in->aosky, 2 * n);
    auto fut = svgf_half(temporal_utility, 3 * n), ft = svgf_half(g_t, n), fn = svgf_u8(g_normal, n);
    bind2d(S::u_SH, fsh.data(), W, H, 4, true); bind2d(S::u_CoCg, fcc.data(), W, H, 2, true);
    bind2d(S::u_VarianceTexture, fvar.data(), W, H, 1, true); bind2d(S::u_AO, fao.data(), W, H, 2, true);
    bind2d(S::u_Utility, fao.data(), W, H, 2, true);   /* Pipeline.cpp:2693 binds the temporal set's AO image here; never used */
    bind2d(S::u_TemporalMoment, fut.data(), W, H, 3, true);
    bind2d(S::u_PositionTexture, ft.data(), W, H, 1, true); bind2d(S::u_NormalTexture, fn.data(), W, H, 1, false);
    bind2d(S::u_BlockIDTexture, fn.data(), W, H, 1, false);   /* declared, never sampled by main() */
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_Step = p->step; S::u_ShouldDetailWeight = true; S::DO_SPATIAL = p->do_spatial != 0; S::u_LargeKernel = p->large_kernel != 0;
    S::AGGRESSIVE_DISOCCLUSION_HANDLING = p->aggressive_disocclusion != 0;
    S::u_ColorPhiBias = p->color_phi_bias; S::u_DeltaTime = 0.0f; S::u_Time = p->time; S::u_ResolutionScale = p->resolution_scale;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam; S::v_RayDirection = vec3(0.0f, 0.0f, 1.0f);
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out->sh[4 * i + c] = vxo::float_to_half(S::o_SH[c]);
            for (int c = 0; c < 2; ++c) out->cocg[2 * i + c] = vxo::float_to_half(S::o_CoCg[c]);
            out->x[i] = vxo::float_to_half(S::o_Variance);
            for (int c = 0; c < 2; ++c) out->aosky[2 * i + c] = vxo::float_to_unorm8(S::o_AOAndSkylighting[c]);
        }
}
#endif

#ifdef VXREF_HAVE_SVGFPreSpatial
/* Spatial3x3Initial.glsl as dispatched at Core/Pipeline.cpp:2381-2424: in = DiffuseRawTraceFBO (x = R16F utility), out = DiffusePreTemporal_SpatialFBO */
extern "C" void vxref_svgf_prespatial(const vxrt_svgf_prespatial_params* p, const vxref_svgf_set* in, const uint16_t* g_t, const uint8_t* g_normal,
                                      int gw, int gh, const vxref_svgf_out* out) {
    namespace S = shader_SVGFPreSpatial;
    const int W = p->width, H = p->height;
    const size_t n = (size_t)W * H;
    auto fsh = svgf_half(in->sh, 4 * n), fcc = svgf_half(in->cocg, 2 * n), fut = svgf_half(in->x, n), fao = svgf_u8(in->aosky, 2 * n);
    auto ft = svgf_half(g_t, (size_t)gw * gh), fn = svgf_u8(g_normal, (size_t)gw * gh);
    bind2d(S::u_SH, fsh.data(), W, H, 4, true); bind2d(S::u_CoCg, fcc.data(), W, H, 2, true);
    bind2d(S::u_Utility, fut.data(), W, H, 1, true); bind2d(S::u_AO, fao.data(), W, H, 2, true);
    bind2d(S::u_PositionTexture, ft.data(), gw, gh, 1, true); bind2d(S::u_NormalTexture, fn.data(), gw, gh, 1, false);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_DeltaTime = 0.0f; S::u_Time = p->time;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam; S::v_RayDirection = vec3(0.0f, 0.0f, 1.0f);
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out->sh[4 * i + c] = vxo::float_to_half(S::o_SH[c]);
            for (int c = 0; c < 2; ++c) out->cocg[2 * i + c] = vxo::float_to_half(S::o_CoCg[c]);
            out->x[i] = vxo::float_to_half(S::o_Utility);
            for (int c = 0; c < 2; ++c) out->aosky[2 * i + c] = vxo::float_to_unorm8(S::o_AOSky[c]);
        }
}
#endif

#ifdef VXREF_HAVE_ReflectionTraceFrag
/* SampleLPVData of ReflectionTraceFrag.glsl (:1484-1528; the light the engine's reflections take from the propagation volume), called
 * as a function on caller points: u_LPV = level volume (R8, LINEAR), u_LPVBlocks = block-type volume (R8UI, NEAREST), BlockAverageColorData =
 * avg512 (binding 4, Pipeline.cpp:3240-3251); LPVDither (:714-723) is the caller's.  Points are in voxel units like HitPosition. */
extern "C" void vxref_lpv_sample(const uint8_t* level, const uint8_t* block_type, const float* avg512, const float* points, int32_t n,
                                 const float dither[3], float* rgb_out) {
    namespace S = shader_ReflectionTraceFrag;
    S::u_LPV.data = level; S::u_LPV.w = 384; S::u_LPV.h = 128; S::u_LPV.d = 384;
    S::u_LPVBlocks.data = block_type; S::u_LPVBlocks.w = 384; S::u_LPVBlocks.h = 128; S::u_LPVBlocks.d = 384;
    S::BlockAverageColorData.data = reinterpret_cast<const vec4*>(avg512);
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        S::LPVDither = vec3(dither[0], dither[1], dither[2]);   /* thread-local, like every mutable global of a shader */
        const vec3 r = S::SampleLPVData(vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
        rgb_out[3 * i] = r.x; rgb_out[3 * i + 1] = r.y; rgb_out[3 * i + 2] = r.z;
    }
    S::u_LPV.data = nullptr; S::u_LPVBlocks.data = nullptr; S::BlockAverageColorData.data = nullptr;
}
#endif

#ifdef VXREF_HAVE_LPVAverageColor
/* PrecomputeAverageBlockColor.comp as dispatched by Volumetrics::CreateVolume (VolumetricFloodFill.cpp:102-123): one invocation, the scene's
 * albedo array and block table; out = BlockAverageColorData, 128 x vec4 */
extern "C" void vxref_lpv_average_colors(float* out512) {
    namespace S = shader_LPVAverageColor;
    S::u_BlockAlbedo.t = &g_scene.tex[0];
    BIND_BLOCK_SSBO(S)
    S::shader_reset(); S::shader_main();
    for (int i = 0; i < 128; ++i)
        for (int c = 0; c < 4; ++c) out512[4 * i + c] = S::BlockAverageColorData[i][c];
}
#endif

/* ---- sun-shadow denoiser (Core/Pipeline.cpp:2947-3044).  raw = ShadowRawTrace (R8 shadow + R16F transversals, sw x sh),
 * hist / out = a ShadowTemporalFBO (R8 shadow + R16F frame count, width x height), G-buffer gw x gh ---- */
#ifdef VXREF_HAVE_ShadowTemporal
extern "C" void vxref_shadow_temporal(const vxrt_shadow_temporal_params* p, const uint8_t* raw_shadow, const uint16_t* raw_transversal, int sw, int sh,
                                      const uint8_t* hist_shadow, const uint16_t* hist_frames, const uint16_t* g_t, const uint8_t* g_normal,
                                      const uint16_t* prev_t, int gw, int gh, uint8_t* out_shadow, uint16_t* out_frames) {
    namespace S = shader_ShadowTemporal;
    const int W = p->width, H = p->height;
    auto fs = svgf_u8(raw_shadow, (size_t)sw * sh), ftr = svgf_half(raw_transversal, (size_t)sw * sh);
    auto hs = svgf_u8(hist_shadow, (size_t)W * H), hf = svgf_half(hist_frames, (size_t)W * H);
    auto ft = svgf_half(g_t, (size_t)gw * gh), fn = svgf_u8(g_normal, (size_t)gw * gh), pt = svgf_half(prev_t, (size_t)gw * gh);
    bind2d(S::u_CurrentColorTexture, fs.data(), sw, sh, 1, true); bind2d(S::u_ShadowTransversals, ftr.data(), sw, sh, 1, true);
    bind2d(S::u_DenoisedTransversals, ftr.data(), sw, sh, 1, true);   /* declared, never sampled */
    bind2d(S::u_PreviousColorTexture, hs.data(), W, H, 1, true); bind2d(S::u_FrameCount, hf.data(), W, H, 1, true);
    bind2d(S::u_CurrentPositionTexture, ft.data(), gw, gh, 1, true); bind2d(S::u_PreviousFramePositionTexture, pt.data(), gw, gh, 1, true);
    bind2d(S::u_NormalTexture, fn.data(), gw, gh, 1, false);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_PrevView.load(p->prev_view); S::u_PrevProjection.load(p->prev_projection);
    S::u_ShadowTemporal = p->shadow_temporal != 0; S::u_ShouldFilterShadows = true;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam; S::v_RayDirection = vec3(0.0f, 0.0f, 1.0f);
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            out_shadow[i] = vxo::float_to_unorm8(S::o_Color[0]);
            out_frames[i] = vxo::float_to_half(S::o_Frames);
        }
}
#endif

#ifdef VXREF_HAVE_ShadowFilter
extern "C" void vxref_shadow_filter(const vxrt_shadow_filter_params* p, const uint8_t* in_shadow, const uint16_t* in_frames, int iw, int ih,
                                    const uint16_t* raw_transversal, int sw, int sh, const uint16_t* g_t, const uint8_t* g_normal, int gw, int gh,
                                    uint8_t* out_shadow) {
    namespace S = shader_ShadowFilter;
    const int W = p->width, H = p->height;
    auto fs = svgf_u8(in_shadow, (size_t)iw * ih), ff = svgf_half(in_frames, (size_t)iw * ih), ftr = svgf_half(raw_transversal, (size_t)sw * sh);
    auto ft = svgf_half(g_t, (size_t)gw * gh), fn = svgf_u8(g_normal, (size_t)gw * gh);
    bind2d(S::u_InputTexture, fs.data(), iw, ih, 1, true); bind2d(S::u_FrameCount, ff.data(), iw, ih, 1, true);
    bind2d(S::u_IntersectionTransversals, ftr.data(), sw, sh, 1, true);
    bind2d(S::u_PositionTexture, ft.data(), gw, gh, 1, true); bind2d(S::u_NormalTexture, fn.data(), gw, gh, 1, false);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_ShadowFilterScale = p->filter_scale;
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::shader_reset(); S::shader_main();
            out_shadow[(size_t)py * W + px] = vxo::float_to_unorm8(S::o_Color);
        }
}
#endif

/* ---- reflection temporal filter (Core/Pipeline.cpp:3316-3400).  cur = ReflectionTraceFBO (colour RGBA16F, hit distance R16F,
 * emissive mask R8, rw x rh) + the previous frame's hit distance; hist = PrevReflectionTemporalFBO images 0 and 2 (width x height);
 * primary G-buffer of this and the previous frame (gw x gh); GeneratedGBuffer PBR (mw x mh) ---- */
#ifdef VXREF_HAVE_SpecularTemporal
extern "C" void vxref_specular_temporal(const vxrt_specular_temporal_params* p, const uint16_t* cur_color_h4, const uint16_t* cur_hitdist,
                                        const uint8_t* cur_mask, const uint16_t* prev_hitdist, int rw, int rh, const uint16_t* hist_color_h4,
                                        const uint16_t* hist_hitdist, const uint16_t* g_t, const uint8_t* g_normal, const uint16_t* prev_t,
                                        const uint8_t* prev_normal, int gw, int gh, const uint8_t* pbr_u8x4, int mw, int mh,
                                        uint16_t* out_color_h4, uint16_t* out_frames, uint16_t* out_hitdist) {
    namespace S = shader_SpecularTemporal;
    const int W = p->width, H = p->height;
    auto fc = svgf_half(cur_color_h4, (size_t)rw * rh * 4), fh = svgf_half(cur_hitdist, (size_t)rw * rh), fm = svgf_u8(cur_mask, (size_t)rw * rh);
    auto fph = svgf_half(prev_hitdist, (size_t)rw * rh);
    auto hc = svgf_half(hist_color_h4, (size_t)W * H * 4), hh = svgf_half(hist_hitdist, (size_t)W * H);
    auto ft = svgf_half(g_t, (size_t)gw * gh), fn = svgf_u8(g_normal, (size_t)gw * gh), pt = svgf_half(prev_t, (size_t)gw * gh), pn = svgf_u8(prev_normal, (size_t)gw * gh);
    auto fp = svgf_u8(pbr_u8x4, (size_t)mw * mh * 4);
    bind2d(S::u_CurrentColorTexture, fc.data(), rw, rh, 4, true); bind2d(S::u_SpecularHitDist, fh.data(), rw, rh, 1, true);
    bind2d(S::u_EmissivityIntersectionMask, fm.data(), rw, rh, 1, true); bind2d(S::u_PrevSpecularHitDist, fph.data(), rw, rh, 1, true);
    bind2d(S::u_PreviousColorTexture, hc.data(), W, H, 4, true); bind2d(S::u_TemporalHitDist, hh.data(), W, H, 1, true);
    bind2d(S::u_CurrentPositionTexture, ft.data(), gw, gh, 1, true); bind2d(S::u_PreviousFramePositionTexture, pt.data(), gw, gh, 1, true);
    bind2d(S::u_NormalTexture, fn.data(), gw, gh, 1, false); bind2d(S::u_PreviousNormalTexture, pn.data(), gw, gh, 1, false);
    bind2d(S::u_PBRTex, fp.data(), mw, mh, 4, true);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection);
    S::u_PrevView.load(p->prev_view); S::u_PrevProjection.load(p->prev_projection);
    S::u_MinimumMix = 0.0f; S::u_MaximumMix = 0.95f; S::u_TemporalQuality = 1; S::u_ReflectionTemporal = true;   /* Pipeline.cpp:3342-3345 */
    S::TEMPORAL_SPEC = p->temporal_spec != 0; S::u_FireflyRejection = p->firefly_rejection != 0;
    S::u_AggressiveFireflyRejection = p->aggressive_firefly_rejection != 0; S::u_SmartClip = p->smart_clip != 0;
    S::u_RoughnessWeight = p->roughness_weight != 0; S::u_TemporallyStabializeHitDistance = p->stabilize_hit_distance != 0;
    S::u_PrevCameraPos = vec3(p->prev_camera_pos[0], p->prev_camera_pos[1], p->prev_camera_pos[2]);
    S::u_CurrentCameraPos = vec3(p->current_camera_pos[0], p->current_camera_pos[1], p->current_camera_pos[2]);
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam; S::v_RayDirection = vec3(0.0f, 0.0f, 1.0f);
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out_color_h4[4 * i + c] = vxo::float_to_half(S::o_Color[c]);
            out_frames[i] = vxo::float_to_half(S::o_AccumulatedFrames);
            out_hitdist[i] = vxo::float_to_half(S::o_HitDistanceStable);
        }
}
#endif

/* ---- reflection spatial denoiser (Core/Pipeline.cpp:3404-3560): one direction per call; `time` feeds u_Time, which the shader's
 * jitter term consumes and then truncates away ---- */
#ifdef VXREF_HAVE_ReflectionDenoise
extern "C" void vxref_reflection_denoise(const vxrt_reflection_denoise_params* p, const uint16_t* in_color_h4, int iw, int ih, const uint16_t* frames,
                                         const uint16_t* hitdist, int tw, int th, int hw, int hh, const uint16_t* g_t, const uint8_t* g_normal,
                                         const uint8_t* g_block, int gw, int gh, const uint16_t* gb_normal_h3, const uint8_t* pbr_u8x4, int mw, int mh,
                                         uint16_t* out_color_h4, float time) {
    namespace S = shader_ReflectionDenoise;
    const int W = p->width, H = p->height;
    auto fc = svgf_half(in_color_h4, (size_t)iw * ih * 4), ffr = svgf_half(frames, (size_t)tw * th), fhd = svgf_half(hitdist, (size_t)hw * hh);
    auto ft = svgf_half(g_t, (size_t)gw * gh), fn = svgf_u8(g_normal, (size_t)gw * gh), fb = svgf_u8(g_block, (size_t)gw * gh);
    auto fgn = svgf_half(gb_normal_h3, (size_t)mw * mh * 3), fp = svgf_u8(pbr_u8x4, (size_t)mw * mh * 4);
    bind2d(S::u_InputTexture, fc.data(), iw, ih, 4, true); bind2d(S::u_Frames, ffr.data(), tw, th, 1, true);
    bind2d(S::u_SpecularHitData, fhd.data(), hw, hh, 1, true);
    bind2d(S::u_PositionTexture, ft.data(), gw, gh, 1, true); bind2d(S::u_NormalTexture, fn.data(), gw, gh, 1, false);
    bind2d(S::u_BlockIDTex, fb.data(), gw, gh, 1, false);
    bind2d(S::u_GBufferNormals, fgn.data(), mw, mh, 3, true); bind2d(S::u_GBufferPBR, fp.data(), mw, mh, 4, true);
    S::u_InverseView.load(p->inv_view); S::u_InverseProjection.load(p->inv_projection); S::u_View.load(p->view);
    S::u_Dir = p->dir != 0; S::u_RoughnessBias = p->roughness_bias != 0; S::u_NormalMapAware = p->normal_map_aware != 0;
    S::u_HandleLobeDeviation = p->handle_lobe_deviation != 0; S::u_DeriveFromDiffuseSH = p->derive_from_diffuse_sh != 0;
    S::u_AmplifyReflectionTransversalWeight = p->amplify_transversal_weight != 0; S::u_TemporalWeight = p->temporal_weight != 0;
    S::u_Dimensions = vec2((float)W, (float)H); S::u_Step = 1; S::u_Time = time;
    S::u_ReflectionDenoisingRadiusBias = p->radius_bias; S::u_NormalMapWeightStrength = p->normal_map_weight_strength;
    S::u_ReflectionDenoiserScale = p->denoiser_scale; S::u_ResolutionScale = p->resolution_scale;
    S::u_RoughnessNormalWeightBiasStrength = p->roughness_normal_weight_bias_strength;
    const vec3 cam = vec3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    int r0, r1;
    rows_of(p->tile, H, &r0, &r1);
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, 0.5f, 1.0f);
            S::v_TexCoords = vec2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            S::v_RayOrigin = cam; S::v_RayDirection = vec3(0.0f, 0.0f, 1.0f);
            S::shader_reset(); S::shader_main();
            const size_t i = (size_t)py * W + px;
            for (int c = 0; c < 4; ++c) out_color_h4[4 * i + c] = vxo::float_to_half(S::o_SpatialResult[c]);
        }
}
#endif
