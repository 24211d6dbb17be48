// traverse_alpha.cuh — VoxelTraversalDF_AlphaTest + StopRay (InitialRayTraceFrag.glsl:189-305 for the primary pass,
// ShadowRayTraceFrag.glsl:105-220 for the shadow pass), the traversal variant that looks through texels with
// alpha <= 0.975 of blocks flagged Transparent in the block table.  Off by default in the engine
// (Core/Pipeline.cpp:146-147, "has known artifacts"), so this is the literal loop, not the tuned one of
// traverse.cuh.  Reference behaviour kept on purpose: after four unresolved DDA steps inside a cut-out block the
// `else` of the `Euclidean == 1` test runs with Euclidean == 0 and moves the ray one unit BACKWARDS (:285-288).
#pragma once
#include "shading.cuh"

struct AlphaCtx {
    TexArrayDev albedo;                      // u_AlbedoTextures
    const int32_t* __restrict__ block_data;  // BlockAlbedoData[128] ... BlockTransparentData[128] at +4*128
    f3 viewer;                               // u_InverseView[3].xyz
    float g_K;                               // 1 / (tan(radians(u_FOV) / (2 * u_Dimensions.x)) * 2), evaluated on the host
    int shadow_variant;                      // ShadowRayTraceFrag.glsl: only v is flipped, LOD biased by -2
};

// StopRay (:189-203 / Shadow :105-118).  An N matching no axis leaves uv undefined in the shader: pinned to (0, 0).
VXD bool stop_ray(const AlphaCtx& c, f3 P, f3 N, int block) {
    const int id = iclamp(block, 0, 127);
    if (__ldg(c.block_data + 4 * 128 + id) == 0) return true;
    f2 uv = F2(0.0f, 0.0f);
    calculate_uv(P, N, uv);
    uv.y = 1.0f - uv.y;
    if (!c.shadow_variant) uv.x = 1.0f - uv.x;
    const float D = distance(P, c.viewer);
    const int LOD = cvt_trunc(log2f(512.0f / (1.0f / D * c.g_K)));
    const float lod = c.shadow_variant ? gclamp((float)LOD - 2.0f, 0.0f, 8.0f) : gclamp((float)LOD, 0.0f, 8.0f);
    const float Alpha = texarray_sample(c.albedo, uv.x, uv.y, (float)__ldg(c.block_data + id), lod).w;
    return Alpha > 0.975f;
}

// one DDA step (:233-243 == :268-282), the arithmetic of traverse_df_tail
VXD void alpha_dda_step(f3& origin, f3 direction, f3 inv, int sx, int sy, int sz, int& MinIdx) {
    const int px = (1 + sx) >> 1, py = (1 + sy) >> 1, pz = (1 + sz) >> 1;
    int gx = cvt_trunc(origin.x), gy = cvt_trunc(origin.y), gz = cvt_trunc(origin.z);
    f3 W = origin - F3((float)gx, (float)gy, (float)gz);
    f3 DF = (F3((float)px, (float)py, (float)pz) - W) * inv;
    MinIdx = (DF.x < DF.y && sx != 0) ? ((DF.x < DF.z || sz == 0) ? 0 : 2) : ((DF.y < DF.z || sz == 0) ? 1 : 2);
    W = W + direction * comp(DF, MinIdx);
    if (MinIdx == 0) { gx += sx; W.x = (float)(1 - px); }
    else if (MinIdx == 1) { gy += sy; W.y = (float)(1 - py); }
    else { gz += sz; W.z = (float)(1 - pz); }
    origin = F3((float)gx, (float)gy, (float)gz) + W;
    if (MinIdx == 0) origin.x += (float)sx * 0.0001f;
    else if (MinIdx == 1) origin.y += (float)sy * 0.0001f;
    else origin.z += (float)sz * 0.0001f;
}

VXD f3 axis_normal(int MinIdx, int sx, int sy, int sz) {
    f3 n = F3(0.0f);
    set_comp(n, MinIdx, (float)(-(MinIdx == 0 ? sx : (MinIdx == 1 ? sy : sz))));
    return n;
}

template <bool STATS>
__device__ __noinline__ TraceResult traverse_df_alpha(const GridView& g, const AlphaCtx& c, f3 origin, f3 direction, int max_iter, LaneStats* st) {
    const f3 initial_origin = origin;
    const int sx = gsign(direction.x), sy = gsign(direction.y), sz = gsign(direction.z);
    const f3 inv = F3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    bool Intersection = false, returned = false;
    int MinIdx = 0;
    TraceResult r;
    r.t = -1.0f; r.block = 0; r.normal = F3(0.0f);

    for (int itr = 0; itr < max_iter; ++itr) {
        const int lx = cvt_floor(origin.x), ly = cvt_floor(origin.y), lz = cvt_floor(origin.z);
        if (!in_volume(g, lx, ly, lz)) {
            Intersection = false;
            break;
        }
        const int k = __ldg(g.df + (lx + ly * g.sy + lz * g.sz));
        if (STATS) st->iterations++;
        const int E = euclidean_step(k);
        if (E == 0) {
            const int bt = get_voxel(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
            if (stop_ray(c, origin, axis_normal(MinIdx, sx, sy, sz), bt)) break;
            for (int i = 0; i < 4; ++i) {
                alpha_dda_step(origin, direction, inv, sx, sy, sz, MinIdx);
                if (STATS) st->dda++;
                const int b2 = get_voxel(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
                if (b2 > 0 && stop_ray(c, origin, axis_normal(MinIdx, sx, sy, sz), b2)) {
                    r.normal = axis_normal(MinIdx, sx, sy, sz);
                    r.block = b2;
                    r.t = distance(origin, initial_origin);  // b2 > 0
                    returned = true;  // `return` inside the loop (:251-256)
                    break;
                }
            }
            if (returned) break;
        }
        if (E == 1) {
            alpha_dda_step(origin, direction, inv, sx, sy, sz, MinIdx);
            if (STATS) st->dda++;
            Intersection = true;
        } else {
            origin = origin + (float)(E - 1) * direction;  // E == 0 ends up here too: one unit backwards
        }
    }
    if (!returned && Intersection) {
        r.normal = axis_normal(MinIdx, sx, sy, sz);
        r.block = get_voxel(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
        r.t = r.block > 0 ? distance(origin, initial_origin) : -1.0f;
    }
    r.intersection = returned || Intersection;
    r.end = origin;
    if (STATS) { st->rays++; st->hits += (r.t > 0.0f) ? 1u : 0u; }
    return r;
}
