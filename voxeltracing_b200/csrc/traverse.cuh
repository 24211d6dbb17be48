// traverse.cuh — DF-skipping DDA traversal, the hot loop shared by every ray pass.
//
// Restates VoxelTraversalDF (Core/Shaders/InitialRayTraceFrag.glsl:307-374; clones in
// ShadowRayTraceFrag.glsl:222-289, DiffuseRayTraceFrag.glsl:1129-1196,
// ReflectionTraceFrag.glsl:1088-1155) for one ray per thread.  Semantics preserved on purpose
// (SURVEY.md A.2): sticky Intersection flag, unguarded dir.y == 0, 1e-4 nudges included in t, a ray
// that starts inside a solid voxel misses, hitting the iteration cap in empty space misses.
#pragma once
#include "ctx.h"
#include "vmath.cuh"

struct TraceResult {
    float t;        // distance(origin_end, origin_start) or -1
    f3 normal;      // valid iff intersection
    f3 end;         // final position
    int block;      // block id at end (0 if outside / none)
    bool intersection;
};

struct LaneStats {
    unsigned rays, iterations, dda, hits;
};

VXD bool in_volume(const GridView& g, int x, int y, int z) {
    return ((unsigned)x < (unsigned)g.nx) & ((unsigned)y < (unsigned)g.ny) & ((unsigned)z < (unsigned)g.nz);
}
VXD int get_voxel(const GridView& g, int x, int y, int z) {
    if (in_volume(g, x, y, z)) return __ldg(g.blk + (x + y * g.sy + z * g.sz));
    return 0;
}

// int(floor(ToConservativeEuclidean(GetDistance()*255)))  (InitialRayTraceFrag.glsl:89-102,331-333).
// (k/255.0f)*255.0f == k exactly for every unorm8 code (tests/test_oracle_df.py checks the float
// path), so the step is a pure function of the byte: k==1 ? 1 : floor(k * 0.57735026918f).
VXD int euclidean_step(int k) {
    float ce = (k == 1) ? 1.0f : (float)k * 0.57735026918f;
    return __float2int_rd(ce);
}

template <bool STATS>
VXD TraceResult traverse_df(const GridView& g, f3 origin, f3 direction, int max_iter, LaneStats* st) {
    const f3 initial_origin = origin;
    const int sx = gsign(direction.x), sy = gsign(direction.y), sz = gsign(direction.z);
    const int px = (1 + sx) >> 1, py = (1 + sy) >> 1, pz = (1 + sz) >> 1;
    const f3 inv = F3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    bool Intersection = false;
    int MinIdx = 0;

    for (int itr = 0; itr < max_iter; ++itr) {
        int lx = cvt_floor(origin.x), ly = cvt_floor(origin.y), lz = cvt_floor(origin.z);
        if (!in_volume(g, lx, ly, lz)) {
            Intersection = false;
            break;
        }
        int k = __ldg(g.df + (lx + ly * g.sy + lz * g.sz));
        if (STATS) st->iterations++;
        int E = euclidean_step(k);
        if (E == 0) break;
        if (E == 1) {
            if (STATS) st->dda++;
            int gx = cvt_trunc(origin.x), gy = cvt_trunc(origin.y), gz = cvt_trunc(origin.z);
            f3 W = origin - F3((float)gx, (float)gy, (float)gz);
            f3 DF = (F3((float)px, (float)py, (float)pz) - W) * inv;
            MinIdx = (DF.x < DF.y && sx != 0) ? ((DF.x < DF.z || sz == 0) ? 0 : 2)
                                               : ((DF.y < DF.z || sz == 0) ? 1 : 2);
            float dmin = comp(DF, MinIdx);
            W = W + direction * dmin;
            if (MinIdx == 0) { gx += sx; W.x = (float)(1 - px); }
            else if (MinIdx == 1) { gy += sy; W.y = (float)(1 - py); }
            else { gz += sz; W.z = (float)(1 - pz); }
            origin = F3((float)gx, (float)gy, (float)gz) + W;
            if (MinIdx == 0) origin.x += (float)sx * 0.0001f;
            else if (MinIdx == 1) origin.y += (float)sy * 0.0001f;
            else origin.z += (float)sz * 0.0001f;
            Intersection = true;
        } else {
            origin = origin + (float)(E - 1) * direction;
        }
    }

    TraceResult r;
    r.t = -1.0f;
    r.block = 0;
    r.normal = F3(0.0f);
    r.intersection = Intersection;
    r.end = origin;
    if (Intersection) {
        int s = MinIdx == 0 ? sx : (MinIdx == 1 ? sy : sz);
        set_comp(r.normal, MinIdx, (float)(-s));
        r.block = get_voxel(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
        r.t = r.block > 0 ? distance(origin, initial_origin) : -1.0f;
    }
    if (STATS) { st->rays++; st->hits += (r.t > 0.0f) ? 1u : 0u; }
    return r;
}

// warp-aggregated flush of per-lane counters
__device__ __forceinline__ void flush_stats(TraceStatsDev* out, const LaneStats& s) {
    unsigned r = s.rays, i = s.iterations, d = s.dda, h = s.hits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        i += __shfl_xor_sync(0xffffffffu, i, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
        h += __shfl_xor_sync(0xffffffffu, h, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out->rays, (unsigned long long)r);
        atomicAdd(&out->iterations, (unsigned long long)i);
        atomicAdd(&out->dda_steps, (unsigned long long)d);
        atomicAdd(&out->hits, (unsigned long long)h);
    }
}
