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
/* implementation options (no reference counterpart).  "wavefront" = 1 (default; env VXRT_WAVEFRONT) runs
 * the GI / reflection passes as wavefront pipelines (compacted ray queues, one kernel per phase); 0 runs
 * the one-thread-per-pixel kernels.  Both produce bit-identical attachments.                          */
int vxrt_cuda_set_option(vxrt_ctx* ctx, const char* name, int32_t value);
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

/* ---- multi-GPU z-slab distance field (new, the reference is single-GPU) ----
 * The grid is replicated; rank s owns planes [slab_z0[s], slab_z0[s+1]).
 *   phase A: X and Y sweeps plus slab-local Z sweeps on the rank's planes.
 *   exchange (caller, NCCL): all-gather every rank's first and last plane (nx*ny bytes each), obtained
 *            with vxrt_cuda_df_plane_device().
 *   phase B: apply the carries of the other slabs from the gathered planes (device pointers to nslabs
 *            consecutive planes each).
 *   the caller then all-gathers the slabs through vxrt_cuda_grid_device() and calls vxrt_cuda_df_commit().
 * nslabs <= 64.                                                                                     */
int vxrt_cuda_df_slab_phase_a(vxrt_ctx* ctx, int32_t slab, int32_t nslabs, const int32_t* slab_z0 /*nslabs+1*/);
int vxrt_cuda_df_slab_phase_b(vxrt_ctx* ctx, int32_t slab, int32_t nslabs, const int32_t* slab_z0,
                              const void* dev_first_planes, const void* dev_last_planes);
/* device pointer to plane z of the distance field (nx*ny bytes) */
int vxrt_cuda_df_plane_device(vxrt_ctx* ctx, int32_t z, void** dev_ptr);
/* device pointers to the block grid and the distance field (nx*ny*nz bytes each) */
int vxrt_cuda_grid_device(vxrt_ctx* ctx, void** dev_blocks, void** dev_df);
/* declare the distance field valid after it was completed through device pointers (slab all-gather) */
int vxrt_cuda_df_commit(vxrt_ctx* ctx);

/* ---- tables ---- */
/* BlockDataSSBO::CreateBuffers (Core/BlockDataSSBO.cpp:5-40): 6 x int[128] =
 * albedo, normal, pbr, emissive layer ids, transparent flag, sss flag.                      */
int vxrt_cuda_set_block_data(vxrt_ctx* ctx, const int32_t* table /*6*128*/);
/* BlueNoiseDataSSBO ctor (Core/BlueNoiseDataSSBO.cpp:16-31): sobol[65536] ++ scramble[131072]
 * ++ ranking[131072].                                                                        */
int vxrt_cuda_set_blue_noise(vxrt_ctx* ctx, const int32_t* data, int32_t count /*327680*/);
/* BluenoiseTexture (Core/Pipeline.cpp:1532): RGBA8, w x h (256 x 256), file row 0 first. */
int vxrt_cuda_set_blue_noise_texture(vxrt_ctx* ctx, const uint8_t* rgba8, int32_t w, int32_t h);

typedef enum vxrt_texture_kind {
    VXRT_TEX_ALBEDO = 0,  /* GL_SRGB_ALPHA: decoded to linear before filtering */
    VXRT_TEX_NORMAL = 1,
    VXRT_TEX_PBR = 2,
    VXRT_TEX_EMISSIVE = 3
} vxrt_texture_kind;
/* TextureArray::CreateArray (Core/GLClasses/TextureArray.cpp:10-69): level-0 texels
 * layers*h*w*4 bytes, file row 0 first; the mip chain is built inside (2x2 box filter,
 * albedo averaged in linear light).  w == h, power of two, at most 2048.                     */
int vxrt_cuda_set_texture_array(vxrt_ctx* ctx, int32_t kind, int32_t layers, int32_t w, int32_t h,
                                const uint8_t* rgba8);
/* Sky cubemap consumed by GI / reflections (Core/Pipeline.cpp:1468-1470, 2339, 3207):
 * 6 faces (+X,-X,+Y,-Y,+Z,-Z) of res*res RGB float32, rows as uploaded with glTexImage2D.    */
int vxrt_cuda_set_skymap(vxrt_ctx* ctx, int32_t res, const float* rgb_faces);

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
    /* SVGF chain of the diffuse GI (Core/Pipeline.cpp:1151-1156): an image set is four consecutive ids
     * (SH RGBA16F, CoCg RG16F, X, AO/sky RG8); X is the utility triple RGB16F (accumulated frames, second
     * moment, luminance) of the temporal sets and the variance R16F of the others */
    VXRT_ATT_SVGF_TEMPORAL_A = 18, /* DiffuseTemporalFBO1: +0 SH, +1 CoCg, +2 utility RGB16F, +3 AO/sky */
    VXRT_ATT_SVGF_TEMPORAL_B = 22, /* DiffuseTemporalFBO2 */
    VXRT_ATT_SVGF_VARIANCE = 26,   /* VarianceFBO: +0 SH, +1 CoCg, +2 variance R16F (+3 unused) */
    VXRT_ATT_SVGF_DENOISE_A = 30,  /* DiffuseDenoiseFBO: +0 SH, +1 CoCg, +2 variance R16F, +3 AO/sky */
    VXRT_ATT_SVGF_DENOISE_B = 34,  /* DiffuseDenoisedFBO2 */
    VXRT_ATT_PREV_INITIAL_T = 38,      /* previous frame's primary G-buffer (the engine ping-pongs InitialTraceFBO_1/_2, */
    VXRT_ATT_PREV_INITIAL_NORMAL = 39, /*   Core/Pipeline.cpp:2046-2048); filled by vxrt_cuda_svgf_end_frame */
    VXRT_ATT_PREV_INITIAL_BLOCK = 40,
    /* sun-shadow denoiser (Core/Pipeline.cpp:1201-1202): a temporal set is two consecutive ids */
    VXRT_ATT_SHADOW_TEMPORAL_A = 41, /* ShadowTemporalFBO_1: +0 shadow R8, +1 accumulated frames R16F */
    VXRT_ATT_SHADOW_TEMPORAL_B = 43, /* ShadowTemporalFBO_2 */
    VXRT_ATT_SHADOW_FILTERED = 45,   /* ShadowFiltered: R8 */
    /* reflection temporal filter (Core/Pipeline.cpp:1191-1192): a temporal set is three consecutive ids */
    VXRT_ATT_REFL_TEMPORAL_A = 46,   /* ReflectionTemporalFBO_1: +0 colour RGBA16F, +1 accumulation factor R16F, +2 stabilised hit distance R16F */
    VXRT_ATT_REFL_TEMPORAL_B = 49,   /* ReflectionTemporalFBO_2 */
    VXRT_ATT_PREV_REFL_HITDIST = 52, /* previous frame's REFL_HITDIST (the engine ping-pongs ReflectionTraceFBO_1 / _2, :1864-1865); filled by vxrt_cuda_end_frame */
    VXRT_ATT_REFL_DENOISED_A = 53,   /* ReflectionDenoised_1: RGBA16F (x pass, :1193) */
    VXRT_ATT_REFL_DENOISED_B = 54,   /* ReflectionDenoised_2: RGBA16F (y pass: the denoised reflections) */
    VXRT_ATT_SVGF_PRESPATIAL = 55,   /* DiffusePreTemporal_SpatialFBO (:1153): +0 SH, +1 CoCg, +2 utility R16F, +3 AO/sky */
    VXRT_ATT_COUNT = 59
} vxrt_attachment;

/* glGetTexImage equivalent: copies the whole attachment (width*height*bytes_per_pixel). */
int vxrt_cuda_read_attachment(vxrt_ctx* ctx, int32_t attachment, void* host_dst, size_t bytes);
/* glTexImage2D equivalent: (re)defines the attachment as width x height pixels of bytes_per_pixel and fills it from
 * HOST memory (borrowed for the call).  Lets a caller seed history images (the SVGF sets, the previous G-buffer) or
 * feed a pass with inputs produced elsewhere; formats are the ones listed above.                     */
int vxrt_cuda_write_attachment(vxrt_ctx* ctx, int32_t attachment, int32_t width, int32_t height, int32_t bytes_per_pixel,
                               const void* host_src);
/* Asynchronous read-back, the glGetTexImage-into-a-pixel-pack-buffer + fence pattern: the copy is queued on the
 * context's copy stream behind everything issued so far and the call returns at once; the next pass that writes
 * the attachment waits for the copy on the device, later passes that only read it do not.  dst should be
 * page-locked host memory (with pageable memory the copy degrades to a synchronous one); it is owned by the
 * library until vxrt_cuda_wait_reads returns. */
int vxrt_cuda_read_attachment_async(vxrt_ctx* ctx, int32_t id, void* dst, size_t bytes);
int vxrt_cuda_wait_reads(vxrt_ctx* ctx);

/* ---- multi-GPU export (the reference is single-GPU, this is the gather of its render targets to one GPU) ----
 * One process per GPU: the gathering rank allocates a device buffer every other rank can WRITE over NVLink and hands its 64-byte
 * handle to them (cudaIpc*; any byte transport will do - torch.distributed in bench.py); a rendering rank opens the handle and queues
 * copies of its attachments - whole, or the rows of its screen band - into the buffer.  The copies are DMA transfers on the context's
 * copy stream (the copy engines: no SM, no kernel on either GPU), ordered behind the pass that produced the attachment exactly
 * like vxrt_cuda_read_attachment_async, and a later pass that overwrites the attachment waits for them on the device.
 * vxrt_cuda_wait_reads returns when every queued copy has landed. */
int vxrt_cuda_shared_alloc(vxrt_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t handle[64]);
int vxrt_cuda_shared_free(vxrt_ctx* ctx, void* dev_ptr);
int vxrt_cuda_shared_open(vxrt_ctx* ctx, const uint8_t handle[64], void** dev_ptr);   /* maps the peer's buffer, enabling peer access */
int vxrt_cuda_shared_close(vxrt_ctx* ctx, void* dev_ptr);
/* rows [row0, row0 + rows) of an attachment (rows == 0: the whole attachment) to `dst`, the address those rows have in the
 * destination image: device memory of this GPU, of a peer (vxrt_cuda_shared_open) or page-locked host memory. */
int vxrt_cuda_copy_attachment_rows_async(vxrt_ctx* ctx, int32_t id, int32_t row0, int32_t rows, void* dst);
/* the same for a rectangle (rows == 0: every row, cols == 0: every column): `dst_image` is the address of pixel (0, 0) of the
 * destination image, which has the attachment's geometry; one copy on the copy engines (strided - cudaMemcpy2DAsync - for a column range,
 * linear for whole rows). */
int vxrt_cuda_copy_attachment_rect_async(vxrt_ctx* ctx, int32_t id, int32_t row0, int32_t rows, int32_t col0, int32_t cols, void* dst_image);
/* makes the context's stream wait (on the device) for every copy queued so far: an event recorded on the stream afterwards
 * marks the moment the frame has left the GPU (device-side timing of the export; the host does not block). */
int vxrt_cuda_join_reads(vxrt_ctx* ctx);
/* Pass-level concurrency, opted into with vxrt_cuda_set_option(ctx, "pass_overlap", 1): the sun-shadow trace, the reflection pass and the direct term
 * are then queued on a second stream of the context and wait only for the passes queued before the frame's diffuse_trace, i.e. they run beside
 * the GI wavefront like independent draw calls do on the reference's GL queue; the reflection pass meets the GI where it first reads its SH
 * attachments (its shading; at once with derive_from_diffuse_sh).  Every other entry point first makes the context's stream wait for that
 * work, so results never depend on the option; a caller that orders its OWN work after a frame on the context's stream (an event, a kernel)
 * calls vxrt_cuda_join_passes first.  Refinements, each its own option, on by default, none changes a bit of any attachment:
 *   "lane2_direct"   the direct term (reads the G-buffers and the sun shadow only) does not queue behind a pending reflection pass: it runs on
 *                    a third stream that waits for the frame's fork point and for what the second stream held before that pass;
 *   "refl_defer_gi"  without reproject_to_screen_space and lpv_gi the reflection pass reads the GI only as the per-pixel ambient base, which
 *                    is then applied where a sample is accumulated (its last shading kernel), so the pass meets the GI there;
 *   "copy_lanes"     the asynchronous copies (read_attachment_async, copy_attachment_rows / rect_async) of an attachment written by a pass
 *                    still pending on the second / third stream wait for that stream alone, on a copy stream of its own, instead of
 *                    joining the streams first. */
int vxrt_cuda_join_passes(vxrt_ctx* ctx);
/* device pointer + geometry of an attachment (valid until the pass that owns it is re-run at a
 * different size).  Used by the host side for NCCL tile gathers.                              */
int vxrt_cuda_attachment_device(vxrt_ctx* ctx, int32_t attachment, void** dev_ptr, int32_t* width,
                                int32_t* height, int32_t* bytes_per_pixel);

/* Bind caller-owned device memory as the storage of an attachment — the analogue of attaching a texture
 * to an FBO; lets a caller ping-pong output sets like InitialTraceFBO_1 / _2 (Core/Pipeline.cpp:2046-2048)
 * or render straight into a communication buffer.  `capacity` bytes must cover width*height*bpp of the
 * pass that writes it.  dev_ptr == NULL restores context-owned storage.  Rebinding keeps the geometry, so
 * later passes of the frame can consume a set rendered earlier.                                        */
int vxrt_cuda_bind_attachment(vxrt_ctx* ctx, int32_t attachment, void* dev_ptr, size_t capacity);

/* Screen-tile sharding: a pass only shades the rectangle rows [row0, row0+rows) x columns [col0, col0+cols) of
 * the frame; rows == 0 means every row, cols == 0 every column (a zeroed tile is the whole frame).  Attachments always have
 * full-frame geometry.  Column bands keep sky and ground in every rank's tile, so one launch per pass per rank stays balanced.
 * The ray passes (initial / shadow trace, G-buffer, direct term, GI, reflections) take any rectangle; the screen-space
 * filters (SVGF, shadow and reflection denoisers) take row bands only and reject cols != 0.                                  */
typedef struct vxrt_tile {
    int32_t row0;
    int32_t rows;
    int32_t col0;
    int32_t cols;
} vxrt_tile;

/* ---- primary G-buffer pass: InitialRayTraceFrag.glsl, Core/Pipeline.cpp:2051-2094 ---- */
typedef struct vxrt_primary_params {
    float inv_view[16];        /* u_InverseView */
    float inv_projection[16];  /* u_InverseProjection */
    int32_t width, height;     /* u_Dimensions */
    float jitter[2];           /* u_CurrentTAAJitter */
    int32_t jitter_on;         /* u_JitterSceneForTAA */
    int32_t render_distance;   /* u_RenderDistance (iteration cap, 350) */
    int32_t alpha_test;        /* u_ShouldAlphaTest: VoxelTraversalDF_AlphaTest (InitialRayTraceFrag.glsl:189-305); off by
                                  default (Pipeline.cpp:146); needs set_block_data and the albedo texture array */
    vxrt_tile tile;
    float fov;                 /* u_FOV in degrees (Pipeline.cpp:2073); only read when alpha_test != 0 */
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
    int32_t alpha_test;        /* u_ShouldAlphaTest (ShadowRayTraceFrag.glsl:105-220, Pipeline.cpp:2913); off by default */
    int32_t max_iterations;    /* loop cap; the shader uses u_RenderDistance-less constant 350 */
    vxrt_tile tile;
    float fov;                 /* u_FOV in degrees (Pipeline.cpp:2914); only read when alpha_test != 0 */
} vxrt_shadow_params;
int vxrt_cuda_shadow_trace(vxrt_ctx* ctx, const vxrt_shadow_params* p);

/* ---- hit-material fetch: GenerateGBuffer.glsl, Core/Pipeline.cpp:2147-2229 ----
 * consumes INITIAL_INVT (R32F, bilinear), INITIAL_NORMAL, INITIAL_BLOCK; writes GBUF_*.
 * Parallax mapping (u_POM) is off by default and not implemented; lava's animated 3-D textures are
 * not modelled (a lava block is shaded from its ordinary array layers).                            */
typedef struct vxrt_gbuffer_params {
    float inv_view[16];
    float inv_projection[16];
    int32_t width, height;
    int32_t grass_props[10];   /* u_GrassBlockProps  (Pipeline.cpp:2177-2186) */
    int32_t cactus_props[10];  /* u_CactusBlockProps (Pipeline.cpp:2166-2175) */
    vxrt_tile tile;
} vxrt_gbuffer_params;
int vxrt_cuda_generate_gbuffer(vxrt_ctx* ctx, const vxrt_gbuffer_params* p);

/* ---- Cook-Torrance direct term of the colour pass: ColorPassFrag.glsl:394-451, 776, 812-816,
 * 886-899, 1201-1210, dispatched at Core/Pipeline.cpp:3702-3918 ----
 * consumes INITIAL_INVT, GBUF_* and SHADOW (raw trace; the shadow denoiser is out of scope);
 * writes DIRECT = max(mix(SunDirect, MoonDirect, SunVisibility) * !(emissive > 0.05), 1e-6).        */
typedef struct vxrt_direct_params {
    float inv_view[16];
    float inv_projection[16];
    int32_t width, height;
    float viewer_position[3];  /* u_ViewerPosition */
    float sun_direction[3];    /* u_SunDirection */
    float moon_direction[3];   /* u_MoonDirection */
    float sun_color[3];        /* SampleSunColor()  — caller supplied, the sky model is out of scope */
    float moon_color[3];       /* SampleMoonColor() */
    float texture_desat_amount; /* u_TextureDesatAmount (0.1, Pipeline.cpp:245) */
    int32_t amplify_normal_map; /* u_AmplifyNormalMap (false, Pipeline.cpp:276) */
    vxrt_tile tile;
} vxrt_direct_params;
int vxrt_cuda_shade_direct(vxrt_ctx* ctx, const vxrt_direct_params* p);

/* ---- diffuse GI: DiffuseRayTraceFrag.glsl, Core/Pipeline.cpp:2267-2374 ----
 * consumes INITIAL_T (bilinear) + INITIAL_NORMAL, grids, BlockData, blue-noise tables, albedo / PBR /
 * emissive arrays, sky cube map; writes GI_*.  Direct light sampling / MIS (off by default,
 * Pipeline.cpp:80) and the hash RNG (u_UseBlueNoise = false) are not implemented.                  */
typedef struct vxrt_gi_params {
    float inv_view[16];
    float inv_projection[16];
    int32_t width, height;           /* u_Dimensions of the GI target */
    int32_t spp;                     /* u_SPP (3) */
    int32_t checker_spp;             /* u_CheckerSPP = (spp + spp%2)/2 */
    int32_t checkerboard;            /* CHECKERBOARD_SPP */
    int32_t trace_length;            /* u_DiffuseTraceLength (48) */
    int32_t shadow_trace_length;     /* GetShadowAt loop cap (128, DiffuseRayTraceFrag.glsl:1306) */
    int32_t current_frame;           /* u_CurrentFrame */
    int32_t current_frame_mod128;    /* u_CurrentFrameMod128 */
    int32_t use_blue_noise;          /* u_UseBlueNoise (must be 1) */
    int32_t supersample;             /* u_Supersample */
    float halton[2];                 /* u_Halton */
    float sun_direction[3];          /* u_SunDirection */
    float moon_direction[3];         /* u_MoonDirection */
    float sun_visibility;            /* u_SunVisibility */
    float gi_sun_strength;           /* u_GISunStrength (1.0) */
    float gi_sky_strength;           /* u_GISkyStrength (1.125) */
    float diffuse_light_intensity;   /* u_DiffuseLightIntensity (1.25) */
    float viewer_position[3];        /* u_ViewerPosition */
    int32_t apply_player_shadow;     /* u_APPLY_PLAYER_SHADOW (false) */
    vxrt_tile tile;
} vxrt_gi_params;
int vxrt_cuda_diffuse_trace(vxrt_ctx* ctx, const vxrt_gi_params* p);

/* ---- reflections: ReflectionTraceFrag.glsl, Core/Pipeline.cpp:3096-3257 ----
 * consumes INITIAL_T/NORMAL, GBUF_NORMAL/PBR, GI_SH/COCG/AOSKY, SHADOW; writes REFL_*.
 * LPV ambient (u_LPVGI), projected clouds (u_CloudReflections), player reflection and lava UV
 * distortion depend on out-of-scope subsystems and must be off.                                    */
typedef struct vxrt_reflection_params {
    float inv_view[16];
    float inv_projection[16];
    float view[16];                  /* u_View */
    float projection[16];            /* u_Projection */
    int32_t width, height;
    int32_t spp;                     /* u_SPP (2) */
    int32_t checkerboard;            /* CHECKERBOARD_SPEC_SPP */
    int32_t trace_length;            /* u_ReflectionTraceLength (64) */
    int32_t shadow_trace_length;     /* 150 (ReflectionTraceFrag.glsl:1098) */
    int32_t current_frame;
    int32_t current_frame_mod128;
    int32_t use_blue_noise;          /* must be 1 */
    int32_t rough_reflections;       /* u_RoughReflections */
    int32_t roughness_bias;          /* u_RoughnessBias */
    int32_t temporal;                /* TEMPORAL_SPEC / u_TemporalFilterReflections */
    int32_t reproject_to_screen_space; /* u_ReprojectToScreenSpace */
    int32_t derive_from_diffuse_sh;  /* u_DeriveFromDiffuseSH */
    float halton[2];
    float sun_direction[3];
    float moon_direction[3];
    float stronger_light_direction[3];
    float viewer_position[3];
    float sun_strength_modifier;     /* u_SunStrengthModifier (0.85) */
    float moon_strength_modifier;    /* u_MoonStrengthModifier */
    int32_t grass_props[10];
    vxrt_tile tile;
    /* ApproximateGILPV (ReflectionTraceFrag.glsl:673-700, call site :881-883): where the screen-space reprojection of a reflection hit
     * fails, its ambient term comes from the light propagation volume (vxrt_cuda_lpv_repropagate / _edit + the average block colours of
     * vxrt_cuda_lpv_average_colors) instead of the pixel's own diffuse SH.  On by default in the engine (Pipeline.cpp:116,3162). */
    int32_t lpv_gi;                          /* u_LPVGI */
    int32_t use_decoupled_gi;                /* u_UseDecoupledGI (Pipeline.cpp:117,3168): sky light from the AO image's sky-hit channel + LPV */
    int32_t screen_space_skylighting_valid;  /* u_ScreenSpaceSkylightingValid = USE_SVGF (Pipeline.cpp:3167) */
} vxrt_reflection_params;
int vxrt_cuda_reflection_trace(vxrt_ctx* ctx, const vxrt_reflection_params* p);

/* ---- VoxelTraversalDF as a function: a batch of arbitrary rays --------------------------------------------
 * Replaces a direct call of VoxelTraversalDF(origin, direction, normal, block, iterations)
 * (InitialRayTraceFrag.glsl:307-374 and its clones ShadowRayTraceFrag.glsl:222-289,
 * DiffuseRayTraceFrag.glsl:1129-1196, ReflectionTraceFrag.glsl:1088-1155, PostProcessingVert.glsl:103-170) for
 * callers that hold their own rays (picking, probes, the lens-flare visibility ray).  origins / directions:
 * 3*n floats each, HOST memory, borrowed for the call; hits: n records, HOST memory.  Directions are used
 * as given (the shaders pass unit vectors). */
typedef struct vxrt_ray_hit {
    float t;              /* return value: distance(end, origin) or -1 */
    float normal[3];      /* face normal, valid when intersection != 0 */
    float end[3];         /* final ray position */
    int32_t block;        /* block id at the end position, 0 if none */
    int32_t intersection; /* the sticky Intersection flag */
    int32_t iterations;   /* loop iterations executed (distance-field fetches) */
} vxrt_ray_hit;
int vxrt_cuda_trace_rays(vxrt_ctx* ctx, const float* origins, const float* directions, int32_t n, int32_t max_iterations,
                         vxrt_ray_hit* hits);

/* ---- picking ray: World::RaycastDetect (Core/World.cpp:496-546; call site Core/Pipeline.cpp:2044) and the march of
 * World::Raycast (Core/World.cpp:214-262), which places / removes a block next to / at the hit ----
 * positions / directions: 3*n floats each, HOST memory; out: 8 int32 per ray = hit voxel x, y, z, block id, face normal
 * x, y, z (as World::Raycast derives it: -1 / +1 on every axis whose slab plane the last step crossed), found (1 / 0).
 * The march is 48 steps of the block-id grid itself (no distance field); voxel index 0 counts as outside, like the
 * reference's `<= 0` test.  Nothing hit within reach: found = 0 and x = y = z = block = -1 (the reference function has
 * no return statement on that path).  The engine follows a hit with vxrt_cuda_edit_blocks + generate_distance_field. */
int vxrt_cuda_raycast_detect(vxrt_ctx* ctx, const float* positions, const float* directions, int32_t n, int32_t* out);

/* ---- SVGF denoiser chain of the diffuse GI: Core/Shaders/SVGF/{TemporalFilter,VarianceEstimate,
 * SpatialFilter}.glsl, orchestrated by Core/Pipeline.cpp:2377-2710 ----
 * Every pass reads and writes image sets named by their first attachment id (VXRT_ATT_GI_SH for the raw trace output,
 * VXRT_ATT_SVGF_*); the caller sequences them like the engine: temporal (current raw set + previous temporal set ->
 * current temporal set, ping-ponged by frame parity), variance, five spatial iterations with steps 16, 8, 4, 2, 1
 * ping-ponging DENOISE_A / DENOISE_B, then vxrt_cuda_svgf_end_frame.  The optional 3x3 pre-pass
 * (Spatial3x3Initial.glsl, PreTemporalSpatialPass) is vxrt_cuda_svgf_prespatial below: it writes VXRT_ATT_SVGF_PRESPATIAL, which
 * the temporal pass then takes as its in_set instead of the raw trace output. */
typedef struct vxrt_svgf_temporal_params {   /* Pipeline.cpp:2428-2528 */
    float inv_view[16], inv_projection[16];  /* u_InverseView, u_InverseProjection (v_RayOrigin = u_InverseView[3]) */
    float prev_view[16], prev_projection[16];/* u_PrevView, u_PrevProjection */
    int32_t width, height;                   /* size of the GI / temporal images */
    int32_t in_set;                          /* VXRT_ATT_GI_SH */
    int32_t history_set;                     /* previous frame's temporal set */
    int32_t out_set;                         /* this frame's temporal set */
    int32_t be_useful;                       /* u_BeUseful (DO_SVGF_TEMPORAL) */
    vxrt_tile tile;
} vxrt_svgf_temporal_params;
int vxrt_cuda_svgf_temporal(vxrt_ctx* ctx, const vxrt_svgf_temporal_params* p);

/* The 3 x 3 edge-stopping pass in front of the temporal filter (Core/Shaders/Spatial3x3Initial.glsl, dispatched at
 * Core/Pipeline.cpp:2381-2424 when PreTemporalSpatialPass is on, the engine's default): reads the raw trace set and the primary
 * G-buffer, writes VXRT_ATT_SVGF_PRESPATIAL, which the temporal filter then takes as its in_set (:2488-2520).                */
typedef struct vxrt_svgf_prespatial_params {
    float inv_view[16], inv_projection[16];  /* u_InverseView, u_InverseProjection (v_RayOrigin = u_VertInverseView[3]) */
    int32_t width, height;                   /* size of the GI images */
    int32_t in_set;                          /* VXRT_ATT_GI_SH */
    float time;                              /* u_Time: only feeds a jitter the shader computes and never uses */
    vxrt_tile tile;
} vxrt_svgf_prespatial_params;               /* writes VXRT_ATT_SVGF_PRESPATIAL */
int vxrt_cuda_svgf_prespatial(vxrt_ctx* ctx, const vxrt_svgf_prespatial_params* p);

typedef struct vxrt_svgf_variance_params {   /* Pipeline.cpp:2532-2567 */
    float inv_view[16], inv_projection[16];
    int32_t width, height;
    int32_t in_set;                          /* this frame's temporal set */
    int32_t do_spatial;                      /* DO_SPATIAL (DO_VARIANCE_SPATIAL, true) */
    int32_t aggressive_disocclusion;         /* AGGRESSIVE_DISOCCLUSION_HANDLING (true) */
    vxrt_tile tile;
} vxrt_svgf_variance_params;                 /* writes VXRT_ATT_SVGF_VARIANCE */
int vxrt_cuda_svgf_variance(vxrt_ctx* ctx, const vxrt_svgf_variance_params* p);

typedef struct vxrt_svgf_spatial_params {    /* Pipeline.cpp:2592-2700 */
    float inv_view[16], inv_projection[16];
    int32_t width, height;
    int32_t in_set;                          /* SH, CoCg, variance and AO/sky of the previous iteration (VARIANCE set first) */
    int32_t ao_set;                          /* set whose +3 image is u_AO: the temporal set for iteration 0, else in_set */
    int32_t temporal_set;                    /* this frame's temporal set: u_Utility (+3!, Pipeline.cpp:2693) and u_TemporalMoment (+2) */
    int32_t out_set;
    int32_t step;                            /* u_Step */
    int32_t large_kernel;                    /* u_LargeKernel (false) */
    int32_t do_spatial;                      /* DO_SPATIAL (true) */
    int32_t aggressive_disocclusion;         /* AGGRESSIVE_DISOCCLUSION_HANDLING (true) */
    float color_phi_bias;                    /* u_ColorPhiBias (2.8) */
    float time;                              /* u_Time (glfwGetTime) */
    float resolution_scale;                  /* u_ResolutionScale (DiffuseIndirectSuperSampleRes, 0.25) */
    vxrt_tile tile;
} vxrt_svgf_spatial_params;
int vxrt_cuda_svgf_spatial(vxrt_ctx* ctx, const vxrt_svgf_spatial_params* p);
/* end of frame: this frame's primary G-buffer (INITIAL_T / NORMAL / BLOCK) becomes VXRT_ATT_PREV_INITIAL_* */
int vxrt_cuda_svgf_end_frame(vxrt_ctx* ctx);

/* same hand-over under its general name: every temporal filter (SVGF, shadow, reflection) reprojects into VXRT_ATT_PREV_INITIAL_*;
 * when a reflection trace exists its hit distance also becomes VXRT_ATT_PREV_REFL_HITDIST */
int vxrt_cuda_end_frame(vxrt_ctx* ctx);

/* ---- sun-shadow denoiser: Core/Shaders/ShadowTemporalFilter.glsl and ShadowFilter.glsl, dispatched at
 * Core/Pipeline.cpp:2947-3044 when SoftShadows (and DenoiseSunShadows) are on ----
 * temporal: consumes SHADOW + SHADOW_TRANSVERSAL of vxrt_cuda_shadow_trace, INITIAL_T / INITIAL_NORMAL, PREV_INITIAL_T and
 * the previous frame's temporal set (ping-ponged by frame parity like ShadowTemporalFBO_1 / _2, Pipeline.cpp:1862-1863;
 * zero-filled on first use); writes out_set.  The raw trace may be smaller than the temporal images
 * (ShadowTraceResolution <= ShadowSupersampleRes, Pipeline.cpp:1691).                                                  */
typedef struct vxrt_shadow_temporal_params {   /* Pipeline.cpp:2949-3005 */
    float inv_view[16], inv_projection[16];
    float prev_view[16], prev_projection[16];  /* u_PrevView, u_PrevProjection */
    int32_t width, height;                     /* size of the temporal images */
    int32_t history_set;                       /* VXRT_ATT_SHADOW_TEMPORAL_A / _B: previous frame's */
    int32_t out_set;                           /* the other one */
    int32_t shadow_temporal;                   /* u_ShadowTemporal (true at the only call site) */
    vxrt_tile tile;
} vxrt_shadow_temporal_params;
int vxrt_cuda_shadow_temporal(vxrt_ctx* ctx, const vxrt_shadow_temporal_params* p);

typedef struct vxrt_shadow_filter_params {     /* Pipeline.cpp:3009-3044 */
    float inv_view[16], inv_projection[16];
    int32_t width, height;                     /* size of VXRT_ATT_SHADOW_FILTERED */
    int32_t in_set;                            /* this frame's temporal set (shadow + frame count) */
    float filter_scale;                        /* u_ShadowFilterScale (1.0, Pipeline.cpp:137) */
    vxrt_tile tile;
} vxrt_shadow_filter_params;                   /* writes VXRT_ATT_SHADOW_FILTERED */
int vxrt_cuda_shadow_filter(vxrt_ctx* ctx, const vxrt_shadow_filter_params* p);
/* Which R8 image vxrt_cuda_reflection_trace and vxrt_cuda_shade_direct sample as the shadow texture — the engine binds
 * `SoftShadows ? (DenoiseSunShadows ? ShadowFiltered : ShadowTemporalFBO) : ShadowRawTrace` (Pipeline.cpp:3231, 3838):
 * VXRT_ATT_SHADOW (default), VXRT_ATT_SHADOW_TEMPORAL_A / _B or VXRT_ATT_SHADOW_FILTERED.  Sticky until changed.     */
int vxrt_cuda_select_shadow(vxrt_ctx* ctx, int32_t attachment);

/* ---- world producers: the grid is produced in device memory instead of on the host + vxrt_cuda_upload_world.
 * Each of them leaves the context as after upload_world (distance field invalid). ---- */

/* VoxelRT::GenerateWorld (Core/WorldGenerator.cpp:208-313, called from Core/Pipeline.cpp:1276) without structures
 * (gen_structures = false; trees / cacti / cobblestone patches are drawn from rand() and a random_device-seeded
 * mt19937 in column order and are not reproducible in the reference itself).  One FastNoise simplex-fractal height per
 * column (frequency 0.00385, 6 octaves, lacunarity 2, gain 0.5, :237-239), one simplex biome value (frequency 0.01 at
 * x/2, z/2, :253-256), columns filled like SetVerticalBlocks (:49-88).  The permutation tables are FastNoise::SetSeed's
 * (std::mt19937_64), built on the host.  Heights and biomes are bit-identical to FastNoise compiled without contraction. */
typedef struct vxrt_worldgen_params {
    int32_t gen_type;                 /* 1: plains (:230-298); 0: flat world of height 50, all biome 1 (:301-312) */
    int32_t noise_seed, biome_seed;   /* the reference draws them as rand() % 50000 (:217-218) */
    int32_t grass_id, dirt_id, stone_id, sand_id;   /* BlockDatabase::GetBlockID of Grass / Dirt / Stone / Sand (:224-227) */
} vxrt_worldgen_params;
int vxrt_cuda_generate_world(vxrt_ctx* ctx, const vxrt_worldgen_params* p);

/* MCWorldImporter::ImportWorld (Core/NBT/Importer.cpp:85-166, called from Core/Pipeline.cpp:1294): the host side
 * (vxh_mca_*, voxeltracing_b200/host) inflates the Anvil region files and hands over every chunk section as it is stored —
 * 4096 block ids in YZX order, 2048 bytes of 4-bit data values (low nibble first) and its world-space origin — and this
 * scatters them into the grid: a voxel is written when its data value is 0, its id maps (lut[id], GetIDFromMCID,
 * Core/BlockDatabase.cpp:599-612) to a non-zero block and (position - import_origin + (nx/2, 0, nz/2)) lies inside the
 * grid (WriteVoxel, Importer.cpp:67-83).  All arrays are HOST memory.  clear_first = 1 zero-fills the grid first, as
 * ImportWorld does; 0 adds to the existing world.  has_data[i] = 0: section i carries no Data array (values read as 0). */
int vxrt_cuda_import_sections(vxrt_ctx* ctx, const uint8_t* block_ids /* n*4096 */, const uint8_t* data_nibbles /* n*2048 */,
                              const uint8_t* has_data /* n */, const int32_t* section_origins /* 3*n */, int32_t n,
                              const int32_t import_origin[3], const uint8_t lut[256], int32_t clear_first);

/* LightLocations of LoadWorld (Core/WorldFileHandler.cpp:53-69, called from Core/Pipeline.cpp:1254): the voxels whose block has an emissive texture
 * (BlockDataSSBO emissive row >= 0, vxrt_cuda_set_block_data; -1 everywhere until it is set), in ascending order of x + y*nx + z*nx*ny like the
 * reference's scan.  xyz_out: HOST memory for 3*capacity ints (may be NULL when capacity = 0); *count receives the
 * number found, of which min(count, capacity) are written. */
int vxrt_cuda_collect_lights(vxrt_ctx* ctx, int32_t* xyz_out, int32_t capacity, int32_t* count);

/* ---- light propagation volume: Core/VolumetricFloodFill.cpp ----
 * Two byte volumes of the grid's size, x-fastest like the grid: the light level (VolumetricFloodFillVolume, R8) and the block type of
 * the lamp that lit the voxel (ColorDataFloodFillVolume, R8UI).  The reference floods them on the host with FIFO queues and mirrors
 * every voxel with a 1-byte glTexSubImage3D (UploadLight :172-188); here they live in HBM and the flood fill runs on the device,
 * reproducing the order of the reference's queues (the block type of a voxel depends on it).
 *
 * lpv_repropagate = the start-up sequence Core/Pipeline.cpp:1602-1611 and World::RepropogateLPV_ Core/World.cpp:554-572: both volumes
 * cleared, every light location seeded with min(distance_limit, 8) (VoxelRT_FloodFillDistanceLimit, Pipeline.cpp:53) and the block at
 * it, PropogateVolume.  lights_xyz: HOST memory, 3*n ints in queue order; NULL = the LightLocations scan of the device grid
 * (vxrt_cuda_collect_lights order, at most 2^20 lights), consumed on the device without a round trip. */
int vxrt_cuda_lpv_repropagate(vxrt_ctx* ctx, const int32_t* lights_xyz, int32_t n_lights, int32_t distance_limit);
/* The light-volume half of a block edit in World::Raycast: op 1 = `block` was placed at (x, y, z) (Core/World.cpp:273-333), op 0 =
 * `block` was broken there (:395-446); then 4 x DepropogateVolume + PropogateVolume (:482-485).  Call it after vxrt_cuda_edit_blocks
 * has applied the edit to the grid (SetBlock precedes the propagation in the reference).  Whether `block` is a lamp comes from the
 * emissive row of vxrt_cuda_set_block_data.  (x, y, z) must lie strictly inside the grid (:267-271), else VXRT_E_INVALID.          */
int vxrt_cuda_lpv_edit(vxrt_ctx* ctx, int32_t op, int32_t x, int32_t y, int32_t z, int32_t block, int32_t distance_limit);
/* BlockAverageColorData of Core/Shaders/Volumetrics/PrecomputeAverageBlockColor.comp (dispatched once by Volumetrics::CreateVolume,
 * VolumetricFloodFill.cpp:102-123): per block id the average colour of its albedo layer (ten trilinear samples at LOD 5.5 ... 8, / 10,
 * pow 1.8; zero for ids without an albedo layer), from the albedo array of vxrt_cuda_set_texture_array and the table of
 * vxrt_cuda_set_block_data.  Kept on the device for the consumers of the block-type volume; rgba_out: HOST memory for 128*4 floats or NULL. */
int vxrt_cuda_lpv_average_colors(vxrt_ctx* ctx, float* rgba_out);
/* the same table handed over by the caller (glBufferData on AverageColorSSBO): 128*4 floats, HOST memory */
int vxrt_cuda_lpv_set_average_colors(vxrt_ctx* ctx, const float* rgba);
/* SampleLPVData of Core/Shaders/ReflectionTraceFrag.glsl:1516-1528 (with SampleLPVColor :1484-1487 and InterpolateLPVColorDithered
 * :1490-1509) as a function on caller points: the light the engine's reflections take from the volumes (ApproximateGILPV :673-700 adds the
 * base ambient term to it).  points: 3*n floats in voxel units (HitPosition + Normal * 0.5 at the call site :882), dither: LPVDither
 * (:714-723), rgb_out: 3*n floats; all HOST memory.  Needs the volumes (lpv_repropagate / lpv_upload) and lpv_average_colors.  The shader
 * scales coordinates by its hard-coded 384 x 128 x 384 resolution; the volumes are addressed with the context's dimensions.        */
int vxrt_cuda_lpv_sample(vxrt_ctx* ctx, const float* points, int32_t n, const float dither[3], float* rgb_out);
/* The volumes to / from HOST memory (nx*ny*nz bytes each; download: either may be NULL).  Upload = Volumetrics::Reupload (:207-216). */
int vxrt_cuda_lpv_download(vxrt_ctx* ctx, uint8_t* level, uint8_t* block_type);
int vxrt_cuda_lpv_upload(vxrt_ctx* ctx, const uint8_t* level, const uint8_t* block_type);

/* ---- reflection temporal filter: Core/Shaders/SpecularTemporalFilter.glsl, dispatched at
 * Core/Pipeline.cpp:3316-3400 ----
 * Consumes REFL_COLOR / REFL_HITDIST / REFL_EMISSIVE of vxrt_cuda_reflection_trace, INITIAL_T / INITIAL_NORMAL, GBUF_PBR,
 * PREV_INITIAL_T / PREV_INITIAL_NORMAL, PREV_REFL_HITDIST and the previous frame's temporal set (ping-ponged by frame parity
 * like ReflectionTemporalFBO_1 / _2, :1858-1859; the history images start out zero-filled like the engine's FBOs).  Reprojects
 * along the reflected ray (hit-distance reprojection), clips the history to the 3 x 3 neighbourhood of the current trace
 * (ReflectionClipping), rejects fireflies next to emissive hits and writes out_set.                                        */
typedef struct vxrt_specular_temporal_params {   /* Pipeline.cpp:3319-3359 */
    float inv_view[16], inv_projection[16];      /* u_InverseView, u_InverseProjection (v_RayOrigin = u_InverseView[3]) */
    float prev_view[16], prev_projection[16];    /* u_PrevView, u_PrevProjection */
    float current_camera_pos[3], prev_camera_pos[3];   /* u_CurrentCameraPos, u_PrevCameraPos */
    int32_t width, height;                       /* size of the temporal images (ReflectionSuperSampleResolution) */
    int32_t history_set, out_set;                /* VXRT_ATT_REFL_TEMPORAL_A / _B */
    int32_t temporal_spec;                       /* TEMPORAL_SPEC (true, :129) */
    int32_t firefly_rejection;                   /* u_FireflyRejection (ReflectionFireflyRejection, true) */
    int32_t aggressive_firefly_rejection;        /* u_AggressiveFireflyRejection (true) */
    int32_t smart_clip;                          /* u_SmartClip (SmartReflectionClip, true) */
    int32_t roughness_weight;                    /* u_RoughnessWeight (RoughReflections, true) */
    int32_t stabilize_hit_distance;              /* u_TemporallyStabializeHitDistance (true) */
    vxrt_tile tile;
} vxrt_specular_temporal_params;
int vxrt_cuda_specular_temporal(vxrt_ctx* ctx, const vxrt_specular_temporal_params* p);

/* ---- reflection spatial denoiser: Core/Shaders/ReflectionDenoiserNew.glsl, the x and the y pass of
 * Core/Pipeline.cpp:3404-3560 (DenoiseReflections && RoughReflections, USE_NEW_SPECULAR_SPATIAL) ----
 * One direction of a separable bilateral filter (up to 33 taps) whose radius follows roughness, the reflected ray's length
 * and the accumulation factor; weights from depth, face normal, normal-mapped normal, luminance and roughness.  Consumes
 * INITIAL_T / INITIAL_NORMAL, GBUF_NORMAL / GBUF_PBR and the reflection temporal set of this frame (the block-id plane the
 * shader binds only feeds a value it never reads).  The
 * shader's jitter term (:204) truncates to 0 for every pixel and time, so u_Time is not a parameter.                      */
typedef struct vxrt_reflection_denoise_params {
    float inv_view[16], inv_projection[16];      /* u_InverseView, u_InverseProjection (v_RayOrigin = u_InverseView[3]) */
    float view[16];                              /* u_View */
    int32_t width, height;                       /* size of the output = u_Dimensions (:3426) */
    int32_t in_attachment;                       /* u_InputTexture: the temporal set's colour (x pass), VXRT_ATT_REFL_DENOISED_A (y pass) */
    int32_t out_attachment;                      /* VXRT_ATT_REFL_DENOISED_A (x pass) / _B (y pass) */
    int32_t temporal_set;                        /* this frame's temporal set: u_Frames is its +1 image */
    int32_t hit_distance_attachment;             /* u_SpecularHitData: temporal_set + 2 if TEMPORAL_SPEC && TemporallyStabializeHitDistance,
                                                    else VXRT_ATT_REFL_HITDIST (:3466-3471) */
    int32_t dir;                                 /* u_Dir: 1 = x pass, 0 = y pass */
    int32_t roughness_bias;                      /* u_RoughnessBias (ReflectionRoughnessBias, true) */
    int32_t normal_map_aware;                    /* u_NormalMapAware (ReflectionNormalMapWeight, true) */
    int32_t handle_lobe_deviation;               /* u_HandleLobeDeviation (true) */
    int32_t derive_from_diffuse_sh;              /* u_DeriveFromDiffuseSH (false) */
    int32_t amplify_transversal_weight;          /* u_AmplifyReflectionTransversalWeight (true) */
    int32_t temporal_weight;                     /* u_TemporalWeight (ReflectionTemporalWeight && TEMPORAL_SPEC, true) */
    int32_t radius_bias;                         /* u_ReflectionDenoisingRadiusBias (0) */
    float normal_map_weight_strength;            /* u_NormalMapWeightStrength (0.75) */
    float denoiser_scale;                        /* u_ReflectionDenoiserScale (1.0) */
    float resolution_scale;                      /* u_ResolutionScale (ReflectionSuperSampleResolution, 0.25) */
    float roughness_normal_weight_bias_strength; /* u_RoughnessNormalWeightBiasStrength (1.075) */
    vxrt_tile tile;
} vxrt_reflection_denoise_params;
int vxrt_cuda_reflection_denoise(vxrt_ctx* ctx, const vxrt_reflection_denoise_params* p);

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
/* Kernel probe (measurement only): after vxrt_cuda_set_option(ctx, "probe", 1) every launch of the GI path-ray
 * trace kernel (the dominant kernel of a GI frame) is bracketed by a CUDA-event pair on the context's stream.
 * probe_read synchronises and returns the summed duration, the number of launches and, when statistics are
 * enabled, the traversal statistics of those launches alone. */
int vxrt_cuda_probe_read(vxrt_ctx* ctx, double* total_ms, int64_t* launches, vxrt_trace_stats* stats, int32_t reset);
/* Measurement only: the rate (32-byte sectors per second) at which this GPU serves independent 1-byte loads at
 * random addresses of the L2-resident distance field, all SMs loaded: the gather roof the traversal is compared
 * with.  Each thread issues rounds * 8 loads. */
int vxrt_cuda_gather_peak(vxrt_ctx* ctx, int32_t rounds, double* sectors_per_second);

#ifdef __cplusplus
}
#endif
#endif /* VXRT_CUDA_H */
