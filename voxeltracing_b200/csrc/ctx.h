// ctx.h — context object behind the vxrt_cuda_* C ABI (include/vxrt_cuda.h).
// Owns every device allocation: the block-id grid, the distance field, the material / blue-noise
// tables and the pass attachments (the reference keeps these as GL textures, SSBOs and FBO
// attachments: Core/World.h:167-171, Core/BlockDataSSBO.cpp, Core/Pipeline.cpp:1142-1202).
#pragma once
#include <utility>

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/vxrt_cuda.h"
#include <vector>

struct GridView {
    const uint8_t* __restrict__ df;   // distance field, x-fastest
    const uint8_t* __restrict__ blk;  // block ids, x-fastest
    int nx, ny, nz;
    int sy;  // nx
    int sz;  // nx*ny
};

struct Attachment {
    void* ptr = nullptr;
    size_t capacity = 0;  // bytes allocated
    int width = 0, height = 0, bpp = 0;
    bool external = false;  // storage bound by the caller (vxrt_cuda_bind_attachment), never freed here
};

// device view of one block texture array (see texture.cuh)
struct TexArrayDev {
    const uint8_t* data;      // all mip levels; level l starts at data + level_offset[l]; RGBA8, layer-major
    const float* decode;      // 256-entry code -> float table for RGB (sRGB decode for albedo, k/255 otherwise)
    unsigned level_offset[12];
    int w, h, layers, levels;
};
struct TexCubeDev {
    const float* data;  // 6 faces (+X,-X,+Y,-Y,+Z,-Z) x res x res x RGB float
    int res;
};

struct TraceStatsDev {
    unsigned long long rays, iterations, dda_steps, hits;
};

struct vxrt_ctx {
    int device = 0;
    int nx = 0, ny = 0, nz = 0;
    size_t nvox = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int sm_count = 0;
    int64_t launches = 0;

    uint8_t* d_df = nullptr;      // one allocation: distance field, then block ids
    uint8_t* d_blocks = nullptr;
    bool l2_persist = false;      // L2 access-policy window over the grids (set_option "l2_persist", VXRT_L2_PERSIST); see api.cu
    size_t l2_persist_max = 0;
    bool world_uploaded = false;
    bool df_valid = false;
    int df_sx = 0, df_sy = 0;  // measurement aid: override the segment counts of the XY kernel (0 = automatic)
    int df_zver = 2;   // measurement aid (set_option "df_zver"): 1 = df_z_reg_kernel, 2 = df_z_reg2_kernel (u16 lanes kept unpacked)
    int df_dbg = 0;    // measurement aid (set_option "df_dbg"): phases of df_xy2_kernel to skip (bit 0: Y down, 1: Y up, 2: X)
    int df_xyver = 2;  // set_option "df_xyver": 1 = df_xy_slice_kernel, 2 = df_xy2_kernel (the engine's 384 x 128 slice only)
    int df_stage = 0;  // measurement aid (set_option "df_stage"): 1 = XY kernel only, 2 = Z kernel only, 0 = both

    int32_t* d_block_data = nullptr;      // 6*128
    int32_t* d_blue_noise = nullptr;      // sobol ++ scramble ++ ranking
    int32_t blue_noise_count = 0;
    uint8_t* d_blue_tex = nullptr;        // rgba8
    int blue_w = 0, blue_h = 0;

    // block texture arrays (Core/GLClasses/TextureArray.cpp) and the sky cube map
    uint8_t* d_tex_data[4] = {nullptr, nullptr, nullptr, nullptr};
    float* d_tex_decode[4] = {nullptr, nullptr, nullptr, nullptr};
    TexArrayDev tex[4] = {};
    bool tex_set[4] = {false, false, false, false};
    float* d_sky = nullptr;
    TexCubeDev sky = {nullptr, 0};
    std::vector<float> h_sky;  // host copy: per-frame sun / moon colours are evaluated on the host

    // wavefront path-state arena (gi_wavefront.cu, reflect_wavefront.cu) and pipeline selection
    void* d_wf = nullptr;
    size_t wf_cap = 0;
    bool wavefront = true;  // VXRT_WAVEFRONT=0 selects the one-thread-per-pixel GI / reflection kernels
    float filter_snap = 0.0f;   // tolerance mode of the screen-space filters (filter_sampler.cuh); set_option "filter_snap" in units of 1 / 65536
    // GI wavefront: the shadow-queue trace of a bounce runs on a side stream beside the bounce-ray trace (both consume queues shade<0> wrote,
    // neither reads what the other writes); set_option "gi_overlap", on by default, off while the probe brackets the path-ray kernel
    bool gi_overlap = true;
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    // Pass-level concurrency (set_option "pass_overlap", off by default: it is the caller's opt-in, because work then runs on a stream the
    // caller does not see until the next join).  Lane 0 is `stream`; the sun-shadow trace, the reflection pass and the direct term are issued on
    // lane 1 and wait only for what was queued before the frame's diffuse_trace, so they run beside the GI wavefront; the reflection pass
    // waits for the GI itself where it first reads its attachments (refl_gi_event below).  Every other entry point first makes `stream` wait
    // for lane 1 (api.cu REQUIRE_CTX); vxrt_cuda_join_passes does only that.
    bool pass_overlap = false;
    cudaStream_t lane1 = nullptr;
    cudaEvent_t gi_fork = nullptr, gi_done = nullptr, lane0_mark = nullptr, lane1_tail = nullptr;
    bool gi_fork_valid = false, lane1_pending = false, lane1_reads_gi = false;
    // The reflection pass on lane 1: ray generation and the first closest-hit trace need only the G-buffer, the shading needs the GI's SH
    // attachments (its ambient base).  When set, the wavefront launcher makes its stream wait for this event before the first shading kernel
    // instead of the whole pass following the GI.  Own path-state arena (the GI wavefront is using d_wf at the same time).
    cudaEvent_t refl_gi_event = nullptr;
    void* d_wf1 = nullptr;
    size_t wf1_cap = 0;
    // Lane 2 (set_option "lane2_direct", on by default): the direct term reads the primary / material G-buffer and the sun shadow, never the
    // GI or the reflections.  Queued on lane 1 behind a reflection pass it would sit behind that pass's wait for the GI; when such a pass is
    // pending it goes to a stream of its own that waits for the frame's fork point and for what lane 1 held BEFORE the reflection pass
    // (lane1_pre_refl, recorded by the reflection pass's begin).  Every later lane pass and every join wait for lane 2 as well.
    bool lane2_direct = true;
    // set_option "lane1_gbuffer" (off by default): GenerateGBuffer on lane 1.  Nothing on lane 0 reads the material G-buffer between the primary
    // pass and the end of the GI, so the GI can start right behind the primary pass with the material fetch beside its first kernels.
    // Bit-identical, and measured to change nothing (1.7092 vs 1.7098 ms per 1080p frame, profiles/r2_zo_ab_lanes.txt): with the lanes the
    // frame is bound by the issue slots its instructions need, not by its dependency chain any more.
    bool lane1_gbuffer = false;
    cudaStream_t lane2 = nullptr;
    cudaEvent_t lane1_pre_refl = nullptr, lane2_tail = nullptr;
    bool lane1_pre_refl_valid = false, lane2_pending = false;
    // The reflection pass without screen-space reprojection and without the LPV term reads the GI only as the per-PIXEL ambient base
    // (BaseIndirectDiffuse), which enters a hit's colour as ((base * 1) * clamp(AO)) * albedo.  With set_option "refl_defer_gi" (default on)
    // shade_a leaves albedo and the AO factor in the path state and the product is formed where the sample is accumulated (shade_b / final):
    // same operands, same order, bit-identical - and the pass meets the GI only at its last kernel instead of at its first shading kernel.
    bool refl_defer_gi = true;
    // Band pipelining of the wavefront passes (set_option "wf_bands", default 1 = off): the tile of a GI / reflection pass is cut into row
    // bands, each band's whole kernel sequence is queued on its own stream with its own path-state arena, so while one band is in a shading
    // kernel (waiting on memory, issue slots 58 % busy) another is in a trace kernel (issue bound) and the two fill each other's gaps and tails.
    // Per-pixel arithmetic and sample order are untouched: bit-identical by construction (tests/test_gpu_bench_configs.py).
    int wf_bands = 1;
    struct BandSlot {
        cudaStream_t stream = nullptr, aux = nullptr;
        cudaEvent_t done = nullptr, aux_fork = nullptr, aux_join = nullptr;
        void* wf = nullptr;
        size_t wf_cap = 0;
    } band[2][3];                       // [lane the pass runs on][band - 1]
    cudaEvent_t band_fork[2] = {nullptr, nullptr};
    bool gi_fuse_final = true;  // last sample's shade<2> fused with resolve (set_option "gi_fuse_final"; 0 = the separate kernels)

    int32_t* d_slab_z0 = nullptr;  // slab boundaries of the sharded distance-field regeneration (<= 65 ints)

    int32_t* d_edit_buf = nullptr;
    size_t edit_cap = 0;

    uint8_t* d_lpv = nullptr;       // light propagation volume (lpv.cu): light level [nvox], then block type [nvox]
    void* d_lpv_work = nullptr;     // claim keys, the two frontiers / edit queues, scan scratch
    float* d_lpv_avg = nullptr;     // BlockAverageColorData: 128 x vec4 (PrecomputeAverageBlockColor.comp)
    int* h_lpv_flag = nullptr;      // pinned, device-mapped: queue-overflow flag of the edit kernel
    int* d_lpv_flag = nullptr;
    bool lpv_valid = false;
    bool lpv_coop = true;           // one cooperative kernel for the repropagation (set_option "lpv_coop"); 0 = one kernel per phase

    // iteration-capped passes of the queue trace kernels (trace_queue.cuh): caps as bytes, low byte first (0 = end of list);
    // 0 = one uncapped pass per queue.  set_option "trace_caps".  Off by default: bit-identical, but measured no faster
    // (profiles/r2_k_sweep_caps.txt).
    int trace_caps = 0;
    // adaptive hand-over (trace_queue.cuh): a warp appends its stragglers to the continuation queue once at most T of its lanes still
    // have a ray; thresholds per pass as bytes, low byte first.  set_option "trace_spill"; takes precedence over trace_caps.
    int trace_spill = 0;
    void* d_trace_cont = nullptr;   // 2 continuation queues + counters
    size_t trace_cont_cap = 0;      // rays each queue holds

    void* d_ray_buf = nullptr;  // staging of vxrt_cuda_trace_rays: origins | directions | hits
    size_t ray_cap = 0;

    Attachment att[VXRT_ATT_COUNT];
    int shadow_source = VXRT_ATT_SHADOW;  // the image the reflection / colour passes read as the shadow texture (vxrt_cuda_select_shadow)
    // asynchronous read-back (vxrt_cuda_read_attachment_async): copies run on their own stream, ordered against
    // the passes by one event pair per attachment
    cudaStream_t copy_stream = nullptr;
    // set_option "copy_lanes" (default on): a copy of an attachment that a pending lane-1 / lane-2 pass produced waits for THAT lane (not for
    // a join of the lanes into `stream`, which would put it behind the whole GI) and runs on that lane's own copy stream, so it does not
    // queue behind the copies of another lane's attachments either
    cudaStream_t copy_stream_lane[2] = {nullptr, nullptr};
    bool copy_lanes = true;
    uint8_t att_lane[VXRT_ATT_COUNT] = {};           // lane whose pass last (re)wrote the attachment (vxrt_ensure_attachment)
    cudaEvent_t att_ready[VXRT_ATT_COUNT] = {};      // recorded on `stream` when a read is requested
    cudaEvent_t att_read_done[VXRT_ATT_COUNT] = {};  // recorded on `copy_stream` after the copy
    bool att_read_pending[VXRT_ATT_COUNT] = {};
    cudaEvent_t copies_joined = nullptr;             // vxrt_cuda_join_reads

    TraceStatsDev* d_stats = nullptr;  // [0] every trace kernel except the probed one, [1] the probed kernel
    bool stats_on = false;

    // kernel probe (vxrt_cuda_set_option "probe", vxrt_cuda_probe_read): CUDA-event pairs around every launch of the
    // GI path-ray trace kernel, so bench.py can report the dominant kernel's own duration, live, outside a profiler
    bool probe_on = false;
    std::vector<cudaEvent_t> probe_ev;  // pairs
    size_t probe_used = 0;              // events consumed since the last read
    TraceStatsDev probe_acc = {0, 0, 0, 0};  // statistics of the probed kernel folded in by stats_read(reset) since the last probe_read

    GridView grid() const {
        GridView g;
        g.df = d_df; g.blk = d_blocks; g.nx = nx; g.ny = ny; g.nz = nz; g.sy = nx; g.sz = nx * ny;
        return g;
    }
};

// error plumbing (api.cu)
int vxrt_fail(int code, const char* fmt, ...);
int vxrt_check_cuda(cudaError_t e, const char* what);
#define VX_CUDA(call)                                              \
    do {                                                           \
        int _rc = vxrt_check_cuda((call), #call);                  \
        if (_rc != VXRT_OK) return _rc;                            \
    } while (0)

// rectangle of a vxrt_tile clipped to the frame (rows == 0: every row, cols == 0: every column)
inline void vxrt_tile_rect(const vxrt_tile& t, int width, int height, int* r0, int* r1, int* c0, int* c1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
    if (t.cols <= 0) { *c0 = 0; *c1 = width; }
    else { *c0 = t.col0; *c1 = t.col0 + t.cols; if (*c1 > width) *c1 = width; }
}

// Runs `launch(row0, row1)` (a wavefront launcher that queues on c->stream, carves its state from c->d_wf and uses c->aux_*) once per row band
// of [row0, row1): band 0 as is, band k > 0 with the context's stream / arena / side stream swapped for the band's own (ctx.h wf_bands).
template <class F>
int vxrt_run_bands(vxrt_ctx* c, int row0, int row1, F launch) {
    const int rows = row1 - row0;
    int S = c->wf_bands < 1 ? 1 : (c->wf_bands > 4 ? 4 : c->wf_bands);
    while (S > 1 && rows < 64 * S) --S;
    if (S <= 1 || c->probe_on || (c->trace_caps | c->trace_spill)) return launch(row0, row1);
    const int lane = (c->lane1 && c->stream == c->lane1) ? 1 : 0;
    if (!c->band_fork[lane]) VX_CUDA(cudaEventCreateWithFlags(&c->band_fork[lane], cudaEventDisableTiming));
    VX_CUDA(cudaEventRecord(c->band_fork[lane], c->stream));
    int rc = VXRT_OK;
    int edge[5];
    for (int k = 0; k <= S; ++k) edge[k] = k == S ? row1 : row0 + ((rows * k / S) & ~7);   // bands on the 8-row CTA grid
    for (int k = 0; k < S && rc == VXRT_OK; ++k) {
        if (k == 0) { rc = launch(edge[0], edge[1]); continue; }
        vxrt_ctx::BandSlot& b = c->band[lane][k - 1];
        if (!b.stream) {
            VX_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
            VX_CUDA(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
        }
        VX_CUDA(cudaStreamWaitEvent(b.stream, c->band_fork[lane], 0));
        std::swap(c->stream, b.stream); std::swap(c->d_wf, b.wf); std::swap(c->wf_cap, b.wf_cap);
        std::swap(c->aux_stream, b.aux); std::swap(c->aux_fork, b.aux_fork); std::swap(c->aux_join, b.aux_join);
        rc = launch(edge[k], edge[k + 1]);
        std::swap(c->stream, b.stream); std::swap(c->d_wf, b.wf); std::swap(c->wf_cap, b.wf_cap);
        std::swap(c->aux_stream, b.aux); std::swap(c->aux_fork, b.aux_fork); std::swap(c->aux_join, b.aux_join);
        if (rc != VXRT_OK) break;
        VX_CUDA(cudaEventRecord(b.done, b.stream));
        VX_CUDA(cudaStreamWaitEvent(c->stream, b.done, 0));
    }
    return rc;
}

// continuation storage of the iteration-capped trace passes (trace_queue.cuh), allocated on demand (api.cu)
struct TraceCont {
    float4* q[2];      // ping-pong continuation queues (capacity = rays of the largest pass)
    unsigned* meta[2]; // per entry of q: loop state | iterations done << 3 (adaptive hand-over only)
    int* count;        // [pass]: entries appended for pass + 1
};
int vxrt_ensure_trace_cont(vxrt_ctx* c, size_t rays, TraceCont* out);

cudaEvent_t vxrt_probe_event(vxrt_ctx* c);
int vxrt_apply_l2_policy(vxrt_ctx* c);  // next pooled event (nullptr when the probe is off or on error)

// kernel launchers (one per .cu)
int vxrt_launch_distance_field(vxrt_ctx* c);
int vxrt_launch_edit_blocks(vxrt_ctx* c, const int32_t* d_edits, int n);
int vxrt_launch_df_slab_phase_a(vxrt_ctx* c, int z0, int z1);
int vxrt_launch_df_slab_phase_b(vxrt_ctx* c, int slab, int nslabs, const int* d_slab_z0, int z0, int z1, const void* first_planes,
                                const void* last_planes);
int vxrt_launch_initial_trace(vxrt_ctx* c, const vxrt_primary_params& p);
int vxrt_launch_shadow_trace(vxrt_ctx* c, const vxrt_shadow_params& p);
int vxrt_launch_raycast_detect(vxrt_ctx* c, const float* d_pos, const float* d_dir, int n, int32_t* d_out);
int vxrt_launch_gather_peak(vxrt_ctx* c, int rounds, double* sectors_per_second);
int vxrt_launch_trace_rays(vxrt_ctx* c, const float* d_o, const float* d_d, int n, int max_iter, vxrt_ray_hit* d_hits);
int vxrt_ensure_attachment(vxrt_ctx* c, int id, int w, int h, int bpp);
int vxrt_set_texture_array(vxrt_ctx* c, int kind, int layers, int w, int h, const uint8_t* rgba8);
int vxrt_set_skymap(vxrt_ctx* c, int res, const float* rgb_faces);
int vxrt_launch_generate_gbuffer(vxrt_ctx* c, const vxrt_gbuffer_params& p);
int vxrt_launch_shade_direct(vxrt_ctx* c, const vxrt_direct_params& p);
int vxrt_launch_diffuse_trace(vxrt_ctx* c, const vxrt_gi_params& p);
int vxrt_launch_reflection_trace(vxrt_ctx* c, const vxrt_reflection_params& p);
int vxrt_launch_svgf_temporal(vxrt_ctx* c, const vxrt_svgf_temporal_params& p);
int vxrt_launch_svgf_prespatial(vxrt_ctx* c, const vxrt_svgf_prespatial_params& p);
int vxrt_launch_svgf_variance(vxrt_ctx* c, const vxrt_svgf_variance_params& p);
int vxrt_launch_svgf_spatial(vxrt_ctx* c, const vxrt_svgf_spatial_params& p);
int vxrt_launch_svgf_end_frame(vxrt_ctx* c);
int vxrt_launch_shadow_temporal(vxrt_ctx* c, const vxrt_shadow_temporal_params& p);
int vxrt_launch_shadow_filter(vxrt_ctx* c, const vxrt_shadow_filter_params& p);
int vxrt_launch_specular_temporal(vxrt_ctx* c, const vxrt_specular_temporal_params& p);
int vxrt_launch_reflection_denoise(vxrt_ctx* c, const vxrt_reflection_denoise_params& p);
int vxrt_launch_generate_world(vxrt_ctx* c, const vxrt_worldgen_params& p);
int vxrt_launch_import_sections(vxrt_ctx* c, const uint8_t* d_ids, const uint8_t* d_nibbles, const uint8_t* d_has_data,
                                const int32_t* d_origins, int n, const int32_t origin[3], const uint8_t lut[256]);
int vxrt_lights_chunks(const vxrt_ctx* c);
int vxrt_launch_collect_lights(vxrt_ctx* c, unsigned* d_counts, int32_t* d_out, int capacity);
int vxrt_lpv_ensure(vxrt_ctx* c);
int vxrt_launch_lpv_repropagate(vxrt_ctx* c, const int32_t* d_lights, const unsigned* d_count, int capacity, int limit);
int vxrt_launch_lpv_repropagate_coop(vxrt_ctx* c, const int32_t* d_lights, int n_lights, int limit);
int vxrt_launch_lpv_average_colors(vxrt_ctx* c);
int vxrt_launch_lpv_sample(vxrt_ctx* c, const float* d_points, int n, const float dither[3], float* d_out);
int vxrt_launch_lpv_edit(vxrt_ctx* c, int op, int x, int y, int z, int block, int limit, int* overflowed);
int vxrt_launch_diffuse_trace_wavefront(vxrt_ctx* c, const void* gi_args);
int vxrt_launch_reflection_trace_wavefront(vxrt_ctx* c, const void* refl_args);
// host-side evaluation of texture(u_Skymap, dir) on the context's copy of the sky (resources.cu)
void vxrt_host_sky_sample(const vxrt_ctx* c, const float dir[3], float rgb[3]);
// SampleSunColor() / SampleMoonColor() of the GI / reflection shaders, evaluated once per pass on the host
void vxrt_host_sun_color(const vxrt_ctx* c, const float sun_dir[3], float strength, float rgb[3]);
void vxrt_host_moon_color(const vxrt_ctx* c, const float moon_dir[3], float strength, float rgb[3]);
