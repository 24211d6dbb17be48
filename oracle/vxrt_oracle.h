/*
 * vxrt_oracle.h — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement of the reference's GLSL hot path (swr06/VoxelTracing, Core/Shaders/*).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (voxeltracing_b200/libvxrt_cuda.so) never links or calls it.
 *
 * Parity pinning: the reference has no tests and no golden vectors.  The
 * oracle is pinned against the reference's own shader sources compiled for the CPU through a
 * GLSL-as-C++ shim (oracle/_ref, built by oracle/build_ref.py when /root/reference is mounted)
 * and against golden fixtures generated from that build (tests/golden/).
 *
 * Pinned implementation-defined behaviour (GL leaves these to the driver; documented in DESIGN.md):
 *   - unorm8 -> float is k/255.0f; float -> unorm8 is round-half-even(clamp(f,0,1)*255)
 *   - float -> R16F is round-to-nearest-even
 *   - no FMA contraction anywhere; dot(a,b) = (ax*bx + ay*by) + az*bz; normalize(v) = v * (1/sqrt(dot(v,v)))
 *   - mat4*vec4 uses glm's association  (m0*x + m1*y) + (m2*z + m3*w)
 *   - float -> int conversions saturate and map NaN to 0
 *   - min/max follow GLSL/glm: max(a,b) = (a<b)?b:a, min(a,b) = (b<a)?b:a
 */
#ifndef VXRT_ORACLE_H
#define VXRT_ORACLE_H

#include "../include/vxrt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vxo_world {
    const uint8_t* blocks; /* x-fastest block ids */
    const uint8_t* df;     /* distance field, same layout */
    int32_t nx, ny, nz;
} vxo_world;

/* ManhattanDistance{X,Y,Z}.comp in dispatch order of World.cpp:75-110 */
void vxo_distance_field(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, uint8_t* df);
/* literal float-carrying version (imageLoad/imageStore through val/255, floor(r*255)); slow,
 * used on small grids to prove the integer version is an exact restatement                  */
void vxo_distance_field_literal(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, uint8_t* df);
/* brute-force min(254, L1 distance to nearest solid) — self-check */
void vxo_distance_field_brute(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, uint8_t* df);
/* 256-entry step table E[k] computed exactly as the shader does (float path) */
void vxo_step_table(int32_t* out256);

typedef struct vxo_hit {
    float t;          /* return value of VoxelTraversalDF */
    float normal[3];  /* valid when intersection != 0 */
    float end[3];     /* final ray position */
    int32_t block;    /* block id at the end position (0..255), 0 if none */
    int32_t intersection;
    int32_t min_idx;
    int32_t iterations;
    int32_t dda_steps;
} vxo_hit;

/* VoxelTraversalDF (InitialRayTraceFrag.glsl:307-374) */
float vxo_traverse(const vxo_world* w, const float origin[3], const float dir[3], int32_t max_iter,
                   vxo_hit* hit);
void vxo_traverse_batch(const vxo_world* w, const float* origins, const float* dirs, int32_t n, int32_t max_iter, vxo_hit* hits);
/* VoxelTraversalDF_AlphaTest + StopRay + CalculateUV (InitialRayTraceFrag.glsl:189-305,498-540 with shadow_variant == 0,
 * ShadowRayTraceFrag.glsl:105-220 with shadow_variant == 1).  viewer = u_InverseView[3].xyz, g_K as main() computes it.
 * Needs the scene's block table and albedo array.                                             */
struct vxo_scene;
float vxo_traverse_alpha(const struct vxo_scene* s, const float origin[3], const float dir[3], int32_t max_iter,
                         const float viewer[3], float g_K, int32_t shadow_variant, vxo_hit* hit);
/* the scene the primary / shadow passes take their alpha-test resources from while params->alpha_test != 0
 * (process-wide; the reference binds them as GL state the same way) */
void vxo_bind_alpha_scene(const struct vxo_scene* s);
/* g_K = 1 / (tan(radians(fov) / (2 * width)) * 2)  (InitialRayTraceFrag.glsl:438, ShadowRayTraceFrag.glsl:419) */
float vxo_alpha_g_k(float fov_degrees, int32_t width);
/* World::RaycastDetect (Core/World.cpp:496-546), the CPU picking ray behind block placement / removal
 * (World::Raycast, Core/World.cpp:214-262, uses the same march and adds the face normal).  out8 = hit voxel x, y, z,
 * block id, face normal x, y, z (what World::Raycast derives: -/+1 on every axis whose slab was hit), found (1 / 0).
 * No hit within the 48-step reach: the reference falls off the end of the function (undefined); here found = 0 and
 * x = y = z = block = -1.                                                                                    */
void vxo_raycast_detect(const vxo_world* w, const float pos[3], const float dir[3], int32_t out8[8]);
void vxo_raycast_detect_batch(const vxo_world* w, const float* pos, const float* dir, int32_t n, int32_t* out8);
/* plain Amanatides-Woo DDA returning the first solid voxel (self-check).
 * returns 1 and fills voxel[3] on hit, 0 on leaving the volume / max_steps.                 */
int32_t vxo_plain_dda(const vxo_world* w, const float origin[3], const float dir[3], int32_t max_steps,
                      int32_t voxel[3]);

/* InitialRayTraceFrag.glsl main() — outputs as the attachments of Pipeline.cpp:1142 hold them.
 * t32 (optional) receives the float value written to the R16F target before rounding.
 * stats (optional) accumulates.                                                              */
void vxo_initial_trace(const vxo_world* w, const vxrt_primary_params* p, uint16_t* t_half,
                       uint8_t* normal_u8, uint8_t* block_u8, float* inv_t, float* t32,
                       vxrt_trace_stats* stats);

/* ShadowRayTraceFrag.glsl main(); gbuffer = attachments of the primary pass at gw x gh. */
void vxo_shadow_trace(const vxo_world* w, const vxrt_shadow_params* p, const uint16_t* g_t_half,
                      const uint8_t* g_normal_u8, int32_t gw, int32_t gh, const uint8_t* blue_rgba,
                      int32_t bw, int32_t bh, uint8_t* shadow_u8, uint16_t* transversal_half,
                      vxrt_trace_stats* stats);

/* ---- scene: tables, texture arrays, sky map (the GL resources the material / GI / reflection
 * shaders bind) ---- */
typedef struct vxo_scene vxo_scene;
vxo_scene* vxo_scene_create(const vxo_world* w);
void vxo_scene_destroy(vxo_scene* s);
void vxo_scene_set_block_data(vxo_scene* s, const int32_t* table6x128);
void vxo_scene_set_blue_noise(vxo_scene* s, const int32_t* data, int32_t count);
void vxo_scene_set_texture_array(vxo_scene* s, int32_t kind, int32_t layers, int32_t w, int32_t h, const uint8_t* rgba);
void vxo_scene_set_skymap(vxo_scene* s, int32_t res, const float* rgb_faces);
/* mip level `level` of array `kind` as built by the pinned glGenerateMipmap model (for tests) */
void vxo_scene_lpv_average_colors(const vxo_scene* s, float* out512);
/* the two byte volumes of the light propagation volume + BlockAverageColorData (128 x 4 floats), borrowed: ApproximateGILPV of the
 * reflection pass (vxrt_reflection_params.lpv_gi) */
void vxo_scene_set_lpv(vxo_scene* s, const uint8_t* level, const uint8_t* block_type, const float* avg512);
int32_t vxo_scene_texture_level(const vxo_scene* s, int32_t kind, int32_t level, uint8_t* out, int64_t out_bytes);

/* GenerateGBuffer.glsl main(); inputs = primary attachments (1/t R32F, face R8, block R8) at gw x gh */
void vxo_generate_gbuffer(const vxo_scene* s, const vxrt_gbuffer_params* p, const float* g_inv_t, const uint8_t* g_normal,
                          const uint8_t* g_block, int32_t gw, int32_t gh, uint16_t* albedo_h3, uint16_t* normal_h3,
                          uint8_t* pbr_u8x4, uint8_t* texao_u8);
/* Cook-Torrance direct term (ColorPassFrag.glsl:419-451, 776, 812-816, 886-899) */
void vxo_shade_direct(const vxrt_direct_params* p, const float* g_inv_t, int32_t gw, int32_t gh, const uint16_t* albedo_h3,
                      const uint16_t* normal_h3, const uint8_t* pbr_u8x4, const uint8_t* texao_u8, int32_t mw, int32_t mh,
                      const uint8_t* shadow_u8, int32_t sw, int32_t sh, uint16_t* direct_h3);
/* DiffuseRayTraceFrag.glsl main() */
void vxo_diffuse_trace(const vxo_scene* s, const vxrt_gi_params* p, const uint16_t* g_t_half, const uint8_t* g_normal,
                       int32_t gw, int32_t gh, uint16_t* sh_h4, uint16_t* cocg_h2, uint16_t* utility_h, uint8_t* aosky_u8x2,
                       vxrt_trace_stats* stats);
/* ReflectionTraceFrag.glsl main() */
typedef struct vxo_reflection_inputs {
    const uint16_t* g_t_half; const uint8_t* g_normal; int32_t gw, gh;          /* primary G-buffer */
    const uint16_t* gb_normal_h3; const uint8_t* gb_pbr_u8x4; int32_t mw, mh;   /* generated G-buffer */
    const uint16_t* gi_sh_h4; const uint16_t* gi_cocg_h2; const uint8_t* gi_aosky_u8x2; int32_t iw, ih; /* GI outputs */
    const uint8_t* shadow_u8; int32_t sw, sh;                                   /* shadow trace */
} vxo_reflection_inputs;
void vxo_reflection_trace(const vxo_scene* s, const vxrt_reflection_params* p, const vxo_reflection_inputs* in,
                          uint16_t* color_h4, uint16_t* hitdist_h, uint8_t* emissive_u8, vxrt_trace_stats* stats);

/* SpecularTemporalFilter.glsl main() (vxrt_oracle_refl_filter.cpp): reflection trace images (rw x rh), previous temporal
 * set (width x height), primary G-buffers of this and the previous frame (gw x gh), GeneratedGBuffer PBR (mw x mh) */
void vxo_specular_temporal(const vxrt_specular_temporal_params* p, const uint16_t* cur_color_h4, const uint16_t* cur_hitdist,
                           const uint8_t* cur_mask, const uint16_t* prev_hitdist, int rw, int rh, const uint16_t* hist_color_h4,
                           const uint16_t* hist_hitdist, const uint16_t* g_t, const uint8_t* g_normal, const uint16_t* prev_t,
                           const uint8_t* prev_normal, int gw, int gh, const uint8_t* pbr_u8x4, int mw, int mh,
                           uint16_t* out_color_h4, uint16_t* out_frames, uint16_t* out_hitdist);

/* ReflectionDenoiserNew.glsl main(): input colour (iw x ih), u_Frames (tw x th), u_SpecularHitData (hw x hh), primary G-buffer (gw x gh),
 * GeneratedGBuffer normals RGB16F + PBR RGBA8 (mw x mh) */
void vxo_reflection_denoise(const vxrt_reflection_denoise_params* p, const uint16_t* in_color_h4, int iw, int ih, const uint16_t* frames,
                            const uint16_t* hitdist, int tw, int th, int hw, int hh, const uint16_t* g_t, const uint8_t* g_normal,
                            const uint8_t* g_block, int gw, int gh, const uint16_t* gb_normal_h3, const uint8_t* pbr_u8x4, int mw, int mh,
                            uint16_t* out_color_h4);

/* ---- world producers (vxrt_oracle_world.cpp) ---- */
/* FastNoise::GetNoise(x, y) for the Simplex (fractal = 0) / SimplexFractal FBM (fractal = 1) types; xy = 2*n floats */
void vxo_fastnoise_2d(int32_t seed, int32_t fractal, float frequency, int32_t octaves, const float* xy, int32_t n, float* out);
/* VoxelRT::GenerateWorld without structures (Core/WorldGenerator.cpp:208-313) */
void vxo_generate_world(uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const vxrt_worldgen_params* p);
/* MCWorldImporter::ImportRegionFile voxel loop + WriteVoxel (Core/NBT/Importer.cpp:67-83, 112-141) over inflated sections */
void vxo_import_sections(uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const uint8_t* ids, const uint8_t* nibbles,
                         const uint8_t* has_data, const int32_t* origins, int32_t n, const int32_t* import_origin, const uint8_t* lut,
                         int32_t clear_first);
/* LightLocations of LoadWorld (Core/WorldFileHandler.cpp:53-69); returns the number found, writes min(found, capacity) */
int32_t vxo_collect_lights(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const int32_t* table6x128, int32_t* xyz_out,
                           int32_t capacity);

/* format helpers */
uint16_t vxo_float_to_half(float f);
float vxo_half_to_float(uint16_t h);
uint8_t vxo_float_to_unorm8(float f);

void vxo_set_threads(int32_t n); /* 0 = all cores */
int32_t vxo_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
