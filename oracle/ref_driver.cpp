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
    S::u_ShouldAlphaTest = false;
    S::u_RenderDistance = p->render_distance;
    S::u_JitterSceneForTAA = p->jitter_on != 0;
    S::u_CurrentTAAJitter = vec2(p->jitter[0], p->jitter[1]);
    S::u_FOV = 90.0f;
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
    S::u_ShouldAlphaTest = false;
    S::u_Dimensions = vec2((float)p->width, (float)p->height);
    S::u_Halton = vec2(p->halton[0], p->halton[1]);
    S::u_Time = 0.0f;
    S::u_FOV = 90.0f;
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

}  // extern "C"
