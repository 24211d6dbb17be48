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

// ---- conversion-free iteration -------------------------------------------------------------------------
// ncu on the first version of this loop (profiles/r1_b_*): the XU pipe (F2I / I2F, 16 lanes/clk/SM) was the
// busiest pipe of the primary and shadow kernels (54 %), ahead of the ALU.  The loop below produces the same
// bits without a single conversion instruction:
//   floor(x)            FADD.RM x + 1.5*2^23: for x in [-2^22, 2^22) the sum lies in [2^23, 2^24) where floats
//                       are the integers, so rounding down gives floor(x) + 1.5*2^23 exactly; the integer is
//                       the mantissa (bits - 0x4B400000), the float floor is F - 1.5*2^23 (exact).
//                       Anything else (|x| >= 2^22, inf, NaN) lands outside the in-volume window of bit
//                       patterns, and only then the saturating conversion decides (NaN -> 0 as pinned in
//                       DESIGN.md §4); a ray with a NaN coordinate that is still "inside" continues in
//                       traverse_df_tail, the literal loop.
//   ivec3(origin)       == floor(origin) inside the volume (every component >= 0).
//   E(k)                floor(float(k) * 0.57735026918f) == (k * 9459) >> 14 for every byte k
//                       (tests/test_oracle_traverse.py::test_step_table); k == 1 -> 1.  So k == 0 stops,
//                       k in 1..3 is a DDA step, k >= 4 skips E - 1 voxels.
//   float(E - 1)        (2^23 + n) - 2^23 built from the bit pattern 0x4B000000 + n.
//   float(G + s) + (1 - p)   on the stepped axis: floor_float + float(s + 1 - p), exact on integers, then the nudge is
//                       added with one rounding exactly like the shader's `origin[MinIdx] += RaySign[MinIdx] * 0.0001f`;
//                       the other two axes are floor_float + (W + dir * DistanceFactor[MinIdx]).
//   Intersection / MinIdx    one register: `state` = MinIdx | 4 once a DDA step has been taken.
#define VX_FLOOR_MAGIC 12582912.0f
#define VX_FLOOR_MAGIC_BITS 0x4B400000

struct RaySetup {
    f3 d, inv;
    f3 fp;      // float((1 + RaySign) >> 1)
    f3 c;       // float(RaySign) + float(1 - ((1 + RaySign) >> 1)): what a DDA step adds to the floor on its axis (exact)
    f3 nudge;   // float(RaySign) * 0.0001f
    int sx, sy, sz;
    unsigned nx, ny, nz, nxy;  // grid size kept in registers: the loop otherwise reloads it from the constant bank
};
VXD RaySetup ray_setup(const GridView& g, f3 direction) {
    RaySetup r;
    r.nx = (unsigned)g.nx; r.ny = (unsigned)g.ny; r.nz = (unsigned)g.nz; r.nxy = (unsigned)g.sz;
    r.d = direction;
    r.sx = gsign(direction.x); r.sy = gsign(direction.y); r.sz = gsign(direction.z);
    const int px = (1 + r.sx) >> 1, py = (1 + r.sy) >> 1, pz = (1 + r.sz) >> 1;
    r.inv = F3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    r.fp = F3((float)px, (float)py, (float)pz);
    r.c = F3((float)(r.sx + 1 - px), (float)(r.sy + 1 - py), (float)(r.sz + 1 - pz));
    r.nudge = F3((float)r.sx * 0.0001f, (float)r.sy * 0.0001f, (float)r.sz * 0.0001f);
    return r;
}

// the literal loop (one F2I per floor, I2F per ivec -> vec), continuing a ray from iteration `itr`
struct TailState {
    f3 origin;
    int MinIdx;
    bool Intersection;
    unsigned iterations, dda;  // stats deltas
};
template <bool STATS>
__device__ __noinline__ TailState traverse_df_tail(const uint8_t* __restrict__ df, int nx, int ny, int nz, f3 origin, f3 direction, int itr,
                                                   int max_iter, bool Intersection, int MinIdx) {
    GridView g;
    g.df = df; g.blk = nullptr; g.nx = nx; g.ny = ny; g.nz = nz; g.sy = nx; g.sz = nx * ny;
    TailState ts;
    ts.iterations = 0u; ts.dda = 0u;
    const int sx = gsign(direction.x), sy = gsign(direction.y), sz = gsign(direction.z);
    const int px = (1 + sx) >> 1, py = (1 + sy) >> 1, pz = (1 + sz) >> 1;
    const f3 inv = F3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    for (; itr < max_iter; ++itr) {
        int lx = cvt_floor(origin.x), ly = cvt_floor(origin.y), lz = cvt_floor(origin.z);
        if (!in_volume(g, lx, ly, lz)) {
            Intersection = false;
            break;
        }
        int k = __ldg(g.df + (lx + ly * g.sy + lz * g.sz));
        if (STATS) ts.iterations++;
        int E = euclidean_step(k);
        if (E == 0) break;
        if (E == 1) {
            if (STATS) ts.dda++;
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
    ts.origin = origin; ts.MinIdx = MinIdx; ts.Intersection = Intersection;
    return ts;
}
// continues a ray in the literal loop and folds the result back into the caller's state
template <bool STATS>
VXD void run_tail(const GridView& g, f3& origin, f3 direction, int itr, int max_iter, bool& Intersection, int& MinIdx, LaneStats* st) {
    const TailState ts = traverse_df_tail<STATS>(g.df, g.nx, g.ny, g.nz, origin, direction, itr, max_iter, Intersection, MinIdx);
    origin = ts.origin; MinIdx = ts.MinIdx; Intersection = ts.Intersection;
    if (STATS) { st->iterations += ts.iterations; st->dda += ts.dda; }
}

enum { VX_ITER_CONTINUE = 0, VX_ITER_STOP = 1, VX_ITER_TAIL = 2 };

// one iteration of VoxelTraversalDF (InitialRayTraceFrag.glsl:320-371)
template <bool STATS>
VXD int df_iteration(const GridView& g, const RaySetup& r, f3& origin, int& state, LaneStats* st) {
    const float Fx = __fadd_rd(origin.x, VX_FLOOR_MAGIC), Fy = __fadd_rd(origin.y, VX_FLOOR_MAGIC), Fz = __fadd_rd(origin.z, VX_FLOOR_MAGIC);
    // floor as unsigned: bits - 0x4B400000; a negative or non-finite coordinate wraps far above any grid size
    const unsigned lx = __float_as_uint(Fx) - VX_FLOOR_MAGIC_BITS, ly = __float_as_uint(Fy) - VX_FLOOR_MAGIC_BITS,
                   lz = __float_as_uint(Fz) - VX_FLOOR_MAGIC_BITS;
    if (!((lx < r.nx) & (ly < r.ny) & (lz < r.nz))) {
        if (!in_volume(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z))) {
            state &= 3;  // Intersection = false
            return VX_ITER_STOP;
        }
        return VX_ITER_TAIL;  // NaN coordinate converted to 0: leave the fast path
    }
    const int k = __ldg(g.df + (lx + ly * r.nx + lz * r.nxy));
    if (STATS) st->iterations++;
    if (k < 4) {
        if (k == 0) return VX_ITER_STOP;
        if (STATS) st->dda++;
        const f3 fl = F3(Fx - VX_FLOOR_MAGIC, Fy - VX_FLOOR_MAGIC, Fz - VX_FLOOR_MAGIC);  // vec3(ivec3(origin))
        f3 W = origin - fl;
        const f3 DF = (r.fp - W) * r.inv;
        // MinIdx (:345-347) as predicate logic: x iff (DF.x < DF.y && sx != 0) && (DF.x < DF.z || sz == 0), ...
        const bool first = (DF.x < DF.y) & (r.sx != 0), z_off = r.sz == 0;
        const bool ax = first & ((DF.x < DF.z) | z_off), ay = !first & ((DF.y < DF.z) | z_off), az = !(ax | ay);
        state = ax ? 4 : (ay ? 5 : 6);  // Intersection = true, MinIdx
        W = W + r.d * (ax ? DF.x : (ay ? DF.y : DF.z));
        origin.x = ax ? (fl.x + r.c.x) + r.nudge.x : fl.x + W.x;
        origin.y = ay ? (fl.y + r.c.y) + r.nudge.y : fl.y + W.y;
        origin.z = az ? (fl.z + r.c.z) + r.nudge.z : fl.z + W.z;
    } else {
        const float skip = __int_as_float(0x4B000000 + ((k * 9459) >> 14) - 1) - 8388608.0f;  // float(E - 1)
        origin = origin + skip * r.d;
    }
    return VX_ITER_CONTINUE;
}

template <bool STATS>
VXD TraceResult trace_result(const GridView& g, const RaySetup& rs, f3 origin, f3 initial_origin, int state, LaneStats* st) {
    const bool Intersection = (state & 4) != 0;
    const int MinIdx = state & 3;
    TraceResult r;
    r.t = -1.0f;
    r.block = 0;
    r.normal = F3(0.0f);
    r.intersection = Intersection;
    r.end = origin;
    if (Intersection) {
        int s = MinIdx == 0 ? rs.sx : (MinIdx == 1 ? rs.sy : rs.sz);
        set_comp(r.normal, MinIdx, (float)(-s));
        r.block = get_voxel(g, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
        r.t = r.block > 0 ? distance(origin, initial_origin) : -1.0f;
    }
    if (STATS) { st->rays++; st->hits += (r.t > 0.0f) ? 1u : 0u; }
    return r;
}

template <bool STATS>
VXD TraceResult traverse_df(const GridView& g, f3 origin, f3 direction, int max_iter, LaneStats* st) {
    const f3 initial_origin = origin;
    const RaySetup rs = ray_setup(g, direction);
    int state = 0;  // MinIdx = 0, Intersection = false
    for (int itr = 0; itr < max_iter; ++itr) {
        const int c = df_iteration<STATS>(g, rs, origin, state, st);
        if (c == VX_ITER_CONTINUE) continue;
        if (c == VX_ITER_TAIL) {
            bool Intersection = (state & 4) != 0;
            int MinIdx = state & 3;
            run_tail<STATS>(g, origin, direction, itr, max_iter, Intersection, MinIdx, st);
            state = MinIdx | (Intersection ? 4 : 0);
        }
        break;
    }
    return trace_result<STATS>(g, rs, origin, initial_origin, state, st);
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
