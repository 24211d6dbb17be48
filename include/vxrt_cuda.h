/*
 * vxrt_cuda.h — C ABI of the B200-native VXRT voxel ray-traversal hot path.
 *
 * Drop-in boundary for the GLSL dispatches of the reference engine (swr06/VoxelTracing).
 * The reference has no plugin/FFI layer: its de-facto pass interface is "bind inputs, set
 * uniforms, draw a full-screen quad / dispatch compute, later passes read the FBO attachments"
 * (Core/Pipeline.cpp, Core/World.cpp).  Each export below replaces one such call site; the
 * file:line it replaces is cited next to it (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0 (VXRT_OK) or a negative vxrt_status; it never throws or aborts.
 *    vxrt_cuda_last_error() returns a thread-local human-readable string for the last failure.
 *  - host pointers are caller-owned and only borrowed for the duration of the call.
 *  - device memory (grids, tables, attachments) is owned by the opaque vxrt_ctx.
 *  - matrices are 16 floats, column-major, exactly what glm::value_ptr() yields.
 *  - images use the GL convention: row 0 is the BOTTOM row, pixel (x,y) has gl_FragCoord
 *    (x+0.5, y+0.5), linear index y*width + x.
 *  - block grid: uint8 block ids, x fastest: idx = x + y*nx + z*nx*ny  (Core/World.h:46-49).
 *  - all calls on one ctx must come from one thread at a time (the reference is single-threaded);
 *    work is enqueued on the ctx stream and calls that return host data synchronise it.
 */
#ifndef VXRT_CUDA_H
#define VXRT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXRT_WORLD_SIZE_X 384 /* Core/Macros.h:3 */
#define VXRT_WORLD_SIZE_Y 128 /* Core/Macros.h:4 */
#define VXRT_WORLD_SIZE_Z 384 /* Core/Macros.h:5 */

typedef enum vxrt_status {
    VXRT_OK = 0,
    VXRT_E_INVALID = -1,     /* bad argument */
    VXRT_E_CUDA = -2,        /* CUDA runtime error (see last_error) */
    VXRT_E_STATE = -3,       /* call order violated (e.g. trace before upload_world) */
    VXRT_E_NOMEM = -4,
    VXRT_E_UNSUPPORTED = -5
} vxrt_status;

typedef struct vxrt_ctx vxrt_ctx;

/* ---- context (replaces GL context + resource creation, Core/Pipeline.cpp:1212-1540) ---- */
/* dims = {nx,ny,nz}; NULL means the engine's 384x128x384.  Constraints: nx % 16 == 0,
 * nx*ny <= 65536 (one z-slice is staged in shared memory), every dim in [16, 1024].       */
int vxrt_cuda_create(vxrt_ctx** out, int device, const int32_t* dims);
int vxrt_cuda_destroy(vxrt_ctx* ctx);
const char* vxrt_cuda_last_error(void);
/* run all subsequent work of this ctx on an existing cudaStream_t (NULL = ctx-owned stream). */
int vxrt_cuda_set_stream(vxrt_ctx* ctx, void* cuda_stream);
int vxrt_cuda_synchronize(vxrt_ctx* ctx);
/* number of kernels this library has launched on ctx since creation (bench "gpu_launches"). */
int64_t vxrt_cuda_launch_count(vxrt_ctx* ctx);

/* ---- world grid + distance field ---- */
/* World::Buffer -> Texture3D::CreateTexture  (Core/World.h:167-171, Core/Texture3D.cpp:8-28) */
int vxrt_cuda_upload_world(vxrt_ctx* ctx, const uint8_t* blocks);
int vxrt_cuda_download_world(vxrt_ctx* ctx, uint8_t* blocks_out);
/* glTexSubImage3D single-voxel edits (Core/World.cpp:372-373, 458-459); xyz_id = n x {x,y,z,id}.
 * Out-of-range coordinates are rejected with VXRT_E_INVALID and nothing is applied.            */
int vxrt_cuda_edit_blocks(vxrt_ctx* ctx, const int32_t* xyz_id, int32_t n);
/* World::GenerateDistanceField (Core/World.cpp:69-113) + ManhattanDistance{X,Y,Z}.comp */
int vxrt_cuda_generate_distance_field(vxrt_ctx* ctx);
int vxrt_cuda_download_distance_field(vxrt_ctx* ctx, uint8_t* df_out);
/* test hook: overwrite the distance field (lets tests feed an oracle-made field to the tracers). */
int vxrt_cuda_upload_distance_field(vxrt_ctx* ctx, const uint8_t* df);

/* ---- tables ---- */
/* BlockDataSSBO::CreateBuffers (Core/BlockDataSSBO.cpp:5-40): 6 x int[128] =
 * albedo, normal, pbr, emissive layer ids, transparent flag, sss flag.                      */
int vxrt_cuda_set_block_data(vxrt_ctx* ctx, const int32_t* table /*6*128*/);
/* BlueNoiseDataSSBO ctor (Core/BlueNoiseDataSSBO.cpp:16-31): sobol[65536] ++ scramble[131072]
 * ++ ranking[131072].                                                                        */
int vxrt_cuda_set_blue_noise(vxrt_ctx* ctx, const int32_t* data, int32_t count /*327680*/);
/* BluenoiseTexture (Core/Pipeline.cpp:1532): RGBA8, w x h (256 x 256), file row 0 first. */
int vxrt_cuda_set_blue_noise_texture(vxrt_ctx* ctx, const uint8_t* rgba8, int32_t w, int32_t h);

/* ---- attachments (the FBO colour attachments of Core/Pipeline.cpp:1142-1202) ---- */
typedef enum vxrt_attachment {
    VXRT_ATT_INITIAL_T = 0,        /* R16F  hit distance, -1 = miss   InitialTraceFBO[0] */
    VXRT_ATT_INITIAL_NORMAL = 1,   /* R8    face id/10 as unorm8, 255 = miss         [1] */
    VXRT_ATT_INITIAL_BLOCK = 2,    /* R8    block id (id/255 as unorm8)              [2] */
    VXRT_ATT_INITIAL_INVT = 3,     /* R32F  1/t                                       [3] */
    VXRT_ATT_SHADOW = 4,           /* R8    0 / 255                     ShadowRawTrace[0] */
    VXRT_ATT_SHADOW_TRANSVERSAL = 5, /* R16F                                          [1] */
    VXRT_ATT_GBUF_ALBEDO = 6,      /* RGB16F                           GeneratedGBuffer[0] */
    VXRT_ATT_GBUF_NORMAL = 7,      /* RGB16F                                          [1] */
    VXRT_ATT_GBUF_PBR = 8,         /* RGBA8 rough, metal, displacement, emissive      [2] */
    VXRT_ATT_GBUF_TEXAO = 9,       /* R8                                              [3] */
    VXRT_ATT_DIRECT = 10,          /* RGB16F Cook-Torrance direct radiance (ColorPass term) */
    VXRT_ATT_GI_SH = 11,           /* RGBA16F                        DiffuseRawTraceFBO[0] */
    VXRT_ATT_GI_COCG = 12,         /* RG16F                                           [1] */
    VXRT_ATT_GI_UTILITY = 13,      /* R16F                                            [2] */
    VXRT_ATT_GI_AOSKY = 14,        /* RG8                                             [3] */
    VXRT_ATT_REFL_COLOR = 15,      /* RGBA16F                      ReflectionTraceFBO[0] */
    VXRT_ATT_REFL_HITDIST = 16,    /* R16F                                            [1] */
    VXRT_ATT_REFL_EMISSIVE = 17,   /* R8                                              [2] */
    VXRT_ATT_COUNT = 18
} vxrt_attachment;

/* glGetTexImage equivalent: copies the whole attachment (width*height*bytes_per_pixel). */
int vxrt_cuda_read_attachment(vxrt_ctx* ctx, int32_t attachment, void* host_dst, size_t bytes);
/* device pointer + geometry of an attachment (valid until the pass that owns it is re-run at a
 * different size).  Used by the host side for NCCL tile gathers.                              */
int vxrt_cuda_attachment_device(vxrt_ctx* ctx, int32_t attachment, void** dev_ptr, int32_t* width,
                                int32_t* height, int32_t* bytes_per_pixel);

/* Screen-tile sharding: a pass only shades rows [row0, row0+rows) of the frame;
 * rows == 0 means the whole frame.  Attachments always have full-frame geometry.             */
typedef struct vxrt_tile {
    int32_t row0;
    int32_t rows;
} vxrt_tile;

/* ---- primary G-buffer pass: InitialRayTraceFrag.glsl, Core/Pipeline.cpp:2051-2094 ---- */
typedef struct vxrt_primary_params {
    float inv_view[16];        /* u_InverseView */
    float inv_projection[16];  /* u_InverseProjection */
    int32_t width, height;     /* u_Dimensions */
    float jitter[2];           /* u_CurrentTAAJitter */
    int32_t jitter_on;         /* u_JitterSceneForTAA */
    int32_t render_distance;   /* u_RenderDistance (iteration cap, 350) */
    int32_t alpha_test;        /* u_ShouldAlphaTest (must be 0: off by default, Pipeline.cpp:146) */
    vxrt_tile tile;
} vxrt_primary_params;
int vxrt_cuda_initial_trace(vxrt_ctx* ctx, const vxrt_primary_params* p);

/* ---- sun-shadow pass: ShadowRayTraceFrag.glsl, Core/Pipeline.cpp:2888-2945 ----
 * consumes INITIAL_T (bilinear, REPEAT) and INITIAL_NORMAL of the last initial_trace.        */
typedef struct vxrt_shadow_params {
    float inv_view[16];
    float inv_projection[16];
    int32_t width, height;     /* u_Dimensions of the shadow target */
    float light_direction[3];  /* u_LightDirection (StrongerLightDirection) */
    int32_t current_frame;     /* u_CurrentFrame */
    float halton[2];           /* u_Halton */
    int32_t soft_shadows;      /* u_ContactHardeningShadows */
    int32_t alpha_test;        /* u_ShouldAlphaTest (must be 0) */
    int32_t max_iterations;    /* loop cap; the shader uses u_RenderDistance-less constant 350 */
    vxrt_tile tile;
} vxrt_shadow_params;
int vxrt_cuda_shadow_trace(vxrt_ctx* ctx, const vxrt_shadow_params* p);

/* traversal statistics of the most recent pass run with stats enabled */
typedef struct vxrt_trace_stats {
    uint64_t rays;        /* VoxelTraversalDF invocations */
    uint64_t iterations;  /* loop iterations executed = distance-field fetches */
    uint64_t dda_steps;   /* of which single-voxel DDA steps */
    uint64_t hits;
} vxrt_trace_stats;
/* enable/disable (re)counting; when enabled every trace pass accumulates into the counters. */
int vxrt_cuda_stats_enable(vxrt_ctx* ctx, int32_t on);
int vxrt_cuda_stats_read(vxrt_ctx* ctx, vxrt_trace_stats* out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* VXRT_CUDA_H */
