// trace.cu — primary G-buffer pass and sun-shadow pass (sm_100a, compiled with --fmad=false).
//
// primary: InitialRayTraceFrag.glsl:398-496 (GetRayStuff, IntersectBox, main), dispatched from
//          Core/Pipeline.cpp:2051-2094; attachments Core/Pipeline.cpp:1142.
// shadow:  ShadowRayTraceFrag.glsl:303-331,388-398,414-513, dispatched from Pipeline.cpp:2888-2945;
//          attachments Core/Pipeline.cpp:1200.
// One ray per thread; a warp covers an 8x4 pixel tile so the rays of a warp walk neighbouring
// voxels (coherent 32-byte sectors of the x-fastest grids, which are L2/L1 resident).
#include "ctx.h"
#include "traverse.cuh"
#include "traverse_alpha.cuh"

namespace {

struct PrimaryArgs {
    float inv_view[16];
    float inv_proj[16];
    int width, height;
    float jitter_x, jitter_y;
    int jitter_on;
    int max_iter;
    int row0, row1, col0, col1;   // tile rectangle
    uint16_t* t_half;
    uint8_t* normal;
    uint8_t* block;
    float* inv_t;
};

struct ShadowArgs {
    float inv_view[16];
    float inv_proj[16];
    int width, height;
    float light[3];
    int frame;
    float halton_x, halton_y;
    int soft;
    int max_iter;
    int row0, row1, col0, col1;   // tile rectangle
    const uint16_t* g_t;
    const uint8_t* g_normal;
    int gw, gh;
    const uint8_t* blue;
    int bw, bh;
    uint8_t* shadow;
    uint16_t* transversal;
};

// GetRayStuff (InitialRayTraceFrag.glsl:398-416) / GetRayDirectionAt (ShadowRayTraceFrag.glsl:303-308): ray_direction_at of shading.cuh

// IntersectBox (InitialRayTraceFrag.glsl:418-432)
VXD f2 intersect_box(f3 ro, f3 invrd, f3 rad) {
    f3 n = invrd * ro;
    f3 k = F3(fabsf(invrd.x), fabsf(invrd.y), fabsf(invrd.z)) * rad;
    f3 t1 = -n - k;
    f3 t2 = -n + k;
    float tN = gmax(gmax(t1.x, t1.y), t1.z);
    float tF = gmin(gmin(t2.x, t2.y), t2.z);
    if (tN > tF || tF < 0.0f) return F2(-1.0f, -1.0f);
    return F2(tN, tF);
}

// GetNormalID (InitialRayTraceFrag.glsl:140-185)
VXD float normal_id(f3 n) {
    if (n.x == 0.0f && n.y == 0.0f && n.z == 1.0f) return 0.0f / 10.0f;
    if (n.x == 0.0f && n.y == 0.0f && n.z == -1.0f) return 1.0f / 10.0f;
    if (n.x == 0.0f && n.y == 1.0f && n.z == 0.0f) return 2.0f / 10.0f;
    if (n.x == 0.0f && n.y == -1.0f && n.z == 0.0f) return 3.0f / 10.0f;
    if (n.x == -1.0f && n.y == 0.0f && n.z == 0.0f) return 4.0f / 10.0f;
    if (n.x == 1.0f && n.y == 0.0f && n.z == 0.0f) return 5.0f / 10.0f;
    return 0.0f;
}

// register allocation of the two per-pixel trace kernels: like the queue trace kernels (trace_queue.cuh VX_TRACE_OCC) a bare
// __launch_bounds__(256) makes ptxas stop at 40 - 44 registers and reload the grid dimensions from the constant bank inside the loop
#ifndef VX_PRIMARY_OCC
#define VX_PRIMARY_OCC 4   // 44 / 48 registers: primary pass 0.103 -> 0.099 ms, sun shadow unchanged (profiles/r2_x_tocc_sweep.txt, pocc4)
#endif
#if VX_PRIMARY_OCC > 0
#define VX_PRIMARY_BOUNDS __launch_bounds__(256, VX_PRIMARY_OCC)
#else
#define VX_PRIMARY_BOUNDS __launch_bounds__(256)
#endif
// pixel of this thread: CTA = 32x8 pixels, warp = 8x4 tile
VXD void pixel_of_thread(int& px, int& py, int row0, int col0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = col0 + blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    py = row0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
}

template <bool STATS, bool ALPHA>
__global__ void VX_PRIMARY_BOUNDS initial_trace_kernel(GridView g, const __grid_constant__ PrimaryArgs a,
                                                            TraceStatsDev* stats, const __grid_constant__ AlphaCtx alpha) {
    int px, py;
    pixel_of_thread(px, py, a.row0, a.col0);
    const bool active = px < a.col1 && py < a.row1;
    LaneStats ls = {0u, 0u, 0u, 0u};
    if (active) {
        const float W = (float)a.width, H = (float)a.height;
        f2 ss = F2(((float)px + 0.5f) / W, ((float)py + 0.5f) / H);
        if (a.jitter_on) {
            ss.x -= a.jitter_x * (1.0f / W);
            ss.y -= a.jitter_y * (1.0f / H);
        }
        f3 dir = normalize(ray_direction_at(a.inv_view, a.inv_proj, ss));
        f3 ro = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
        const f3 half = F3((float)g.nx / 2.0f, (float)g.ny / 2.0f, (float)g.nz / 2.0f);
        float AddT = 0.0f;
        f2 box = intersect_box(ro - half, F3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z), half);
        if (box.x > 0.0f) {
            AddT = box.x + 0.5f;
            ro = ro + dir * AddT;
        }
        TraceResult r;
        if (ALPHA) r = traverse_df_alpha<STATS>(g, alpha, ro, dir, a.max_iter, &ls);
        else r = traverse_df<STATS>(g, ro, dir, a.max_iter, &ls);
        const bool intersect = r.t > 0.0f && r.block > 0;
        float t = r.t + AddT * (intersect ? 1.0f : 0.0f);
        const size_t i = (size_t)py * a.width + px;
        a.t_half[i] = float_to_half_bits(t);
        a.inv_t[i] = 1.0f / t;
        a.normal[i] = float_to_unorm8(intersect ? normal_id(r.normal) : 1.0f);
        a.block[i] = intersect ? (uint8_t)r.block : (uint8_t)0;
    }
    if (STATS) flush_stats(stats, ls);
}

// texture(R16F, uv): LINEAR + REPEAT (Core/GLClasses/Framebuffer.cpp:64-68), weights in full float
VXD float sample_r16f_bilinear(const uint16_t* __restrict__ img, int w, int h, f2 uv) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
    float t00 = half_bits_to_float(__ldg(img + (size_t)j0 * w + i0)), t10 = half_bits_to_float(__ldg(img + (size_t)j0 * w + i1));
    float t01 = half_bits_to_float(__ldg(img + (size_t)j1 * w + i0)), t11 = half_bits_to_float(__ldg(img + (size_t)j1 * w + i1));
    float top = t00 * (1.0f - a) + t10 * a;
    float bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
VXD float sample_r8_nearest(const uint8_t* __restrict__ img, int w, int h, f2 uv) {
    int i = wrap_repeat(cvt_floor(uv.x * (float)w), w), j = wrap_repeat(cvt_floor(uv.y * (float)h), h);
    return unorm8_to_float(__ldg(img + (size_t)j * w + i));
}

// GetNormalFromID (ShadowRayTraceFrag.glsl:317-328)
VXD f3 normal_from_id(float n) {
    int i = cvt_round(n * 10.0f);
    switch (i) {
        case 0: return F3(0.0f, 0.0f, 1.0f);
        case 1: return F3(0.0f, 0.0f, -1.0f);
        case 2: return F3(0.0f, 1.0f, 0.0f);
        case 3: return F3(0.0f, -1.0f, 0.0f);
        case 4: return F3(-1.0f, 0.0f, 0.0f);
        case 5: return F3(1.0f, 0.0f, 0.0f);
        default: return F3(1.0f, 1.0f, 1.0f);
    }
}

// SampleCone (ShadowRayTraceFrag.glsl:388-398)
VXD f3 sample_cone(f2 Xi, float CosThetaMax) {
    const float PI = 3.14159265359f;
    float CosTheta = (1.0f - Xi.x) + Xi.x * CosThetaMax;
    float SinTheta = sqrtf(1.0f - CosTheta * CosTheta);
    float phi = Xi.y * PI * 2.0f;
    return F3(SinTheta * cosf(phi), SinTheta * sinf(phi), CosTheta);
}

template <bool STATS, bool ALPHA>
__global__ void VX_PRIMARY_BOUNDS shadow_trace_kernel(GridView g, const __grid_constant__ ShadowArgs a,
                                                           TraceStatsDev* stats, const __grid_constant__ AlphaCtx alpha) {
    int px, py;
    pixel_of_thread(px, py, a.row0, a.col0);
    const bool active = px < a.col1 && py < a.row1;
    LaneStats ls = {0u, 0u, 0u, 0u};
    if (active) {
        const size_t i = (size_t)py * a.width + px;
        const float W = (float)a.width, H = (float)a.height;
        f2 tc = F2(((float)px + 0.5f) / W, ((float)py + 0.5f) / H);
        tc = tc + F2(a.halton_x, a.halton_y) * F2(1.0f / W, 1.0f / H);
        const float Dist = sample_r16f_bilinear(a.g_t, a.gw, a.gh, tc);
        uint8_t o_shadow;
        float o_trans;
        if (Dist < 0.0f) {
            o_shadow = 0;
            o_trans = 64.0f;
        } else {
            const f3 cam = F3(a.inv_view[12], a.inv_view[13], a.inv_view[14]);
            const f3 P = cam + normalize(ray_direction_at(a.inv_view, a.inv_proj, tc)) * Dist;
            const f3 L = F3(a.light[0], a.light[1], a.light[2]);
            f3 rd = L;
            if (a.soft) {
                int n = a.frame % 1024;
                float qx = (float)(int)((unsigned)n * 12664745u), qy = (float)(int)((unsigned)n * 9560333u);
                float offx = gfract(qx / 16777216.0f) * 1024.0f, offy = gfract(qy / 16777216.0f) * 1024.0f;
                int sx = cvt_trunc(((float)px + 0.5f) + (float)cvt_trunc(floorf(offx))) % a.bw;
                int sy = cvt_trunc(((float)py + 0.5f) + (float)cvt_trunc(floorf(offy))) % a.bh;
                const uchar4 tx = __ldg(reinterpret_cast<const uchar4*>(a.blue) + ((size_t)sy * a.bw + sx));
                f2 Xi = F2(unorm8_to_float(tx.x), unorm8_to_float(tx.y));
                f3 T = normalize(cross(L, F3(0.0f, 1.0f, 1.0f)));
                f3 B = cross(T, L);
                rd = mat3_mul(T, B, L, sample_cone(Xi, 0.9999505604617f));
            }
            const f3 N = normal_from_id(sample_r8_nearest(a.g_normal, a.gw, a.gh, tc));
            const float NDotL = dot(N, rd);
            if (NDotL <= 0.01f) {
                o_shadow = 255;
                o_trans = 1.0f / 100.0f;
            } else {
                const f3 o = P + N * F3(0.1f);
                const int block_at = get_voxel(g, cvt_floor(o.x), cvt_floor(o.y), cvt_floor(o.z));
                float T = -1.0f;
                if (Dist > 0.0f) {
                    TraceResult r;
                    if (ALPHA) r = traverse_df_alpha<STATS>(g, alpha, o, rd, a.max_iter, &ls);
                    else r = traverse_df<STATS>(g, o, rd, a.max_iter, &ls);
                    T = r.t;
                }
                o_shadow = (T > 0.0f || block_at > 0) ? 255 : 0;
                o_trans = gclamp(T / 100.0f, 0.00001f, 196.0f);
                if (T < 0.0f) o_trans = 4.25f / 100.0f;
            }
        }
        a.shadow[i] = o_shadow;
        a.transversal[i] = float_to_half_bits(o_trans);
    }
    if (STATS) flush_stats(stats, ls);
}

}  // namespace

// u_AlbedoTextures / SSBO 0 / u_FOV of the alpha-tested traversal.  g_K (InitialRayTraceFrag.glsl:438,
// ShadowRayTraceFrag.glsl:419) is evaluated here on the host: radians(x) = x * pi/180 as a float constant, libm tanf.
static AlphaCtx make_alpha_ctx(const vxrt_ctx* c, const float* inv_view, float fov, int width, int shadow_variant) {
    AlphaCtx x;
    x.albedo = c->tex[VXRT_TEX_ALBEDO];
    x.block_data = c->d_block_data;
    x.viewer.x = inv_view[12]; x.viewer.y = inv_view[13]; x.viewer.z = inv_view[14];
    x.g_K = 1.0f / (tanf((fov * 0.01745329251994329577f) / (2.0f * (float)width)) * 2.0f);
    x.shadow_variant = shadow_variant;
    return x;
}

int vxrt_launch_initial_trace(vxrt_ctx* c, const vxrt_primary_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_INITIAL_T, p.width, p.height, 2))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_INITIAL_NORMAL, p.width, p.height, 1))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_INITIAL_BLOCK, p.width, p.height, 1))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_INITIAL_INVT, p.width, p.height, 4))) return rc;
    PrimaryArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    a.jitter_x = p.jitter[0]; a.jitter_y = p.jitter[1];
    a.jitter_on = p.jitter_on;
    a.max_iter = p.render_distance;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    a.t_half = (uint16_t*)c->att[VXRT_ATT_INITIAL_T].ptr;
    a.normal = (uint8_t*)c->att[VXRT_ATT_INITIAL_NORMAL].ptr;
    a.block = (uint8_t*)c->att[VXRT_ATT_INITIAL_BLOCK].ptr;
    a.inv_t = (float*)c->att[VXRT_ATT_INITIAL_INVT].ptr;
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    const AlphaCtx ax = make_alpha_ctx(c, p.inv_view, p.fov, p.width, 0);
    if (p.alpha_test) {
        if (c->stats_on) initial_trace_kernel<true, true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
        else initial_trace_kernel<false, true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
    } else {
        if (c->stats_on) initial_trace_kernel<true, false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
        else initial_trace_kernel<false, false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
    }
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_shadow_trace(vxrt_ctx* c, const vxrt_shadow_params& p) {
    int rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_SHADOW, p.width, p.height, 1))) return rc;
    if ((rc = vxrt_ensure_attachment(c, VXRT_ATT_SHADOW_TRANSVERSAL, p.width, p.height, 2))) return rc;
    ShadowArgs a;
    for (int i = 0; i < 16; ++i) { a.inv_view[i] = p.inv_view[i]; a.inv_proj[i] = p.inv_projection[i]; }
    a.width = p.width; a.height = p.height;
    for (int i = 0; i < 3; ++i) a.light[i] = p.light_direction[i];
    a.frame = p.current_frame;
    a.halton_x = p.halton[0]; a.halton_y = p.halton[1];
    a.soft = p.soft_shadows;
    a.max_iter = p.max_iterations;
    vxrt_tile_rect(p.tile, p.width, p.height, &a.row0, &a.row1, &a.col0, &a.col1);
    const Attachment& gt = c->att[VXRT_ATT_INITIAL_T];
    const Attachment& gn = c->att[VXRT_ATT_INITIAL_NORMAL];
    a.g_t = (const uint16_t*)gt.ptr; a.g_normal = (const uint8_t*)gn.ptr;
    a.gw = gt.width; a.gh = gt.height;
    a.blue = c->d_blue_tex; a.bw = c->blue_w; a.bh = c->blue_h;
    a.shadow = (uint8_t*)c->att[VXRT_ATT_SHADOW].ptr;
    a.transversal = (uint16_t*)c->att[VXRT_ATT_SHADOW_TRANSVERSAL].ptr;
    if (a.row1 <= a.row0 || a.col1 <= a.col0) return VXRT_OK;
    dim3 grid((a.col1 - a.col0 + 31) / 32, (a.row1 - a.row0 + 7) / 8);
    const AlphaCtx ax = make_alpha_ctx(c, p.inv_view, p.fov, p.width, 1);
    if (p.alpha_test) {
        if (c->stats_on) shadow_trace_kernel<true, true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
        else shadow_trace_kernel<false, true><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
    } else {
        if (c->stats_on) shadow_trace_kernel<true, false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
        else shadow_trace_kernel<false, false><<<grid, 256, 0, c->stream>>>(c->grid(), a, c->d_stats, ax);
    }
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

// ---- VoxelTraversalDF for caller-supplied rays (vxrt_cuda_trace_rays) -----------------------------------
namespace {
__global__ void __launch_bounds__(128) trace_rays_kernel(GridView g, const float* __restrict__ o, const float* __restrict__ d, int n, int max_iter,
                                                         vxrt_ray_hit* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LaneStats ls = {0u, 0u, 0u, 0u};
    const TraceResult r = traverse_df<true>(g, F3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), F3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), max_iter, &ls);
    vxrt_ray_hit h;
    h.t = r.t;
    h.normal[0] = r.normal.x; h.normal[1] = r.normal.y; h.normal[2] = r.normal.z;
    h.end[0] = r.end.x; h.end[1] = r.end.y; h.end[2] = r.end.z;
    h.block = r.block; h.intersection = r.intersection ? 1 : 0; h.iterations = (int)ls.iterations;
    out[i] = h;
}
}  // namespace

int vxrt_launch_trace_rays(vxrt_ctx* c, const float* d_o, const float* d_d, int n, int max_iter, vxrt_ray_hit* d_hits) {
    trace_rays_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->grid(), d_o, d_d, n, max_iter, d_hits);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

// ---- picking ray (vxrt_cuda_raycast_detect): World::RaycastDetect, Core/World.cpp:496-546 ----------------------------
namespace {
__global__ void __launch_bounds__(128) raycast_detect_kernel(const uint8_t* __restrict__ blocks, int nx, int ny, int nz, const float* __restrict__ pos,
                                                            const float* __restrict__ dir, int n, int32_t* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    f3 position = F3(pos[3 * r], pos[3 * r + 1], pos[3 * r + 2]);
    const f3 direction = F3(dir[3 * r], dir[3 * r + 1], dir[3 * r + 2]);
    const f3 sign = F3(direction.x > 0.0f ? 1.0f : 0.0f, direction.y > 0.0f ? 1.0f : 0.0f, direction.z > 0.0f ? 1.0f : 0.0f);
    int32_t res[8] = {-1, -1, -1, -1, 0, 0, 0, 0};
    for (int i = 0; i < 48; ++i) {  // block reach
        const f3 tvec = (F3(floorf(position.x + sign.x), floorf(position.y + sign.y), floorf(position.z + sign.z)) - position) / direction;
        const float t = fminf(tvec.x, fminf(tvec.y, tvec.z));  // std::min chain: NaN handling differs only when a component is NaN
        position = position + direction * (t + 0.001f);
        const int fx = cvt_floor(position.x), fy = cvt_floor(position.y), fz = cvt_floor(position.z);
        if (!(fx >= nx || fy >= ny || fz >= nz || fx <= 0 || fy <= 0 || fz <= 0)) {
            const size_t at = (size_t)fx + (size_t)fy * nx + (size_t)fz * nx * ny;  // (int)position == floor inside the bounds
            const int b = __ldg(blocks + at);
            if (b != 0) {
                res[0] = fx; res[1] = fy; res[2] = fz; res[3] = b;
                const float tv[3] = {tvec.x, tvec.y, tvec.z}, sg[3] = {sign.x, sign.y, sign.z};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int nj = (t == tv[j]) ? 1 : 0;
                    if (sg[j] != 0.0f) nj = -nj;
                    res[4 + j] = nj;
                }
                res[7] = 1;
                break;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) out[8 * (size_t)r + k] = res[k];
}
}  // namespace

int vxrt_launch_raycast_detect(vxrt_ctx* c, const float* d_pos, const float* d_dir, int n, int32_t* d_out) {
    raycast_detect_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d_blocks, c->nx, c->ny, c->nz, d_pos, d_dir, n, d_out);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

// ---- gather roof (vxrt_cuda_gather_peak) -----------------------------------------------------------------
// The operative roof of the traversal is the rate at which the memory system serves independent 1-byte loads
// from an L2-resident grid (each moves one 32-byte sector).  This kernel measures it on the distance field
// itself: every thread issues batches of 8 independent loads at hashed addresses (no reuse: L1 cannot help).
namespace {
__global__ void __launch_bounds__(256) gather_peak_kernel(const uint8_t* __restrict__ buf, unsigned n, int rounds, unsigned* __restrict__ sink) {
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned acc = 0;
    for (int r = 0; r < rounds; ++r) {
        unsigned v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x = x * 1664525u + 1013904223u;
            v[i] = __ldg(buf + __umulhi(x, n));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i];
    }
    if (acc == 0xffffffffu) *sink = acc;  // keeps the loads alive
}
}  // namespace

int vxrt_launch_gather_peak(vxrt_ctx* c, int rounds, double* sectors_per_second) {
    cudaEvent_t e0, e1;
    VX_CUDA(cudaEventCreate(&e0));
    VX_CUDA(cudaEventCreate(&e1));
    const int grid = c->sm_count * 8;
    unsigned* sink = reinterpret_cast<unsigned*>(c->d_stats);  // never written: the guard value cannot occur
    gather_peak_kernel<<<grid, 256, 0, c->stream>>>(c->d_df, (unsigned)c->nvox, 4, sink);  // warm-up: pull the grid into L2
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        VX_CUDA(cudaEventRecord(e0, c->stream));
        gather_peak_kernel<<<grid, 256, 0, c->stream>>>(c->d_df, (unsigned)c->nvox, rounds, sink);
        VX_CUDA(cudaEventRecord(e1, c->stream));
        VX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f;
        VX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)grid * 256.0 * rounds * 8.0 / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *sectors_per_second = best;
    c->launches += 4;
    return VXRT_OK;
}
