/*
 * vxrt_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * Distance field, DF-skipping DDA traversal, primary G-buffer pass, sun-shadow pass.
 * Every function cites the reference file:line it restates (paths relative to the reference
 * root, Core/Shaders/<name> unless noted).  Build: g++ -O2 -ffp-contract=off -fopenmp.
 */
#include "vxrt_oracle.h"
#include <algorithm>
#include "vxo_math.h"

#include <stdlib.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace vxo;

static int g_threads = 0;
extern "C" void vxo_set_threads(int32_t n) { g_threads = n; }
extern "C" int32_t vxo_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
static inline int nthreads() { return vxo_get_threads(); }

extern "C" uint16_t vxo_float_to_half(float f) { return float_to_half(f); }
extern "C" float vxo_half_to_float(uint16_t h) { return half_to_float(h); }
extern "C" uint8_t vxo_float_to_unorm8(float f) { return float_to_unorm8(f); }

/* ------------------------------------------------------------------------------------------
 * Distance field
 * ---------------------------------------------------------------------------------------- */

/* ManhattanDistanceX.comp:45-68, ManhattanDistanceY.comp:26-50, ManhattanDistanceZ.comp:22-47,
 * dispatched X, Y, Z (World.cpp:75-110).  One shader invocation owns one grid line, so lines
 * are independent inside a pass.  MaxDistance = min(254, nx+ny+nz) (X.comp:51).              */
extern "C" void vxo_distance_field(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, uint8_t* df) {
    const int maxd = (nx + ny + nz) < 254 ? (nx + ny + nz) : 254;
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    /* X pass (X.comp:53-68) */
#pragma omp parallel for collapse(2) num_threads(nthreads())
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y) {
            const uint8_t* b = blocks + z * sz + y * sy;
            uint8_t* d = df + z * sz + y * sy;
            d[0] = b[0] > 0 ? 0 : (uint8_t)maxd;
            for (int x = 1; x < nx; ++x) {
                int v = 1 + d[x - 1];
                d[x] = b[x] > 0 ? 0 : (uint8_t)(v < maxd ? v : maxd);
            }
            for (int x = nx - 2; x >= 0; --x)
                if (d[x + 1] < d[x]) d[x] = (uint8_t)(1 + d[x + 1]);
        }
    /* Y pass (Y.comp:35-49) */
#pragma omp parallel for collapse(2) num_threads(nthreads())
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) {
            uint8_t* d = df + z * sz + x;
            for (int y = 1; y < ny; ++y)
                if (d[(y - 1) * sy] < d[y * sy]) d[y * sy] = (uint8_t)(1 + d[(y - 1) * sy]);
            for (int y = ny - 2; y >= 0; --y)
                if (d[(y + 1) * sy] < d[y * sy]) d[y * sy] = (uint8_t)(1 + d[(y + 1) * sy]);
        }
    /* Z pass (Z.comp:31-46) */
#pragma omp parallel for collapse(2) num_threads(nthreads())
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            uint8_t* d = df + y * sy + x;
            for (int z = 1; z < nz; ++z)
                if (d[(z - 1) * sz] < d[z * sz]) d[z * sz] = (uint8_t)(1 + d[(z - 1) * sz]);
            for (int z = nz - 2; z >= 0; --z)
                if (d[(z + 1) * sz] < d[z * sz]) d[z * sz] = (uint8_t)(1 + d[(z + 1) * sz]);
        }
}

/* Same passes carried through the r8 image exactly as the shaders do it:
 * imageStore(vec4(val/255.0f)) -> unorm8; imageLoad -> floor(r*255.0f)  (X.comp:25-38).      */
namespace {
struct R8Image {
    uint8_t* p;
    size_t sy, sz;
    float load(int x, int y, int z) const { return floorf(unorm8_to_float(p[x + y * sy + z * sz]) * 255.0f); }
    void store(int x, int y, int z, float val) { p[x + y * sy + z * sz] = float_to_unorm8(val / 255.0f); }
};
}  // namespace
extern "C" void vxo_distance_field_literal(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz,
                                           uint8_t* df) {
    R8Image img = {df, (size_t)nx, (size_t)nx * ny};
    const int MaxDistance = (nx + ny + nz) < 254 ? (nx + ny + nz) : 254;
    auto solid = [&](int x, int y, int z) { return unorm8_to_float(blocks[x + y * img.sy + z * img.sz]) > 0; };
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y) {
            img.store(0, y, z, solid(0, y, z) ? 0 : MaxDistance);
            for (int x = 1; x < nx; ++x)
                img.store(x, y, z, solid(x, y, z) ? 0 : gmin((float)MaxDistance, 1 + img.load(x - 1, y, z)));
            for (int x = nx - 2; x >= 0; --x)
                if (img.load(x + 1, y, z) < img.load(x, y, z)) img.store(x, y, z, 1 + img.load(x + 1, y, z));
        }
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) {
            for (int y = 1; y < ny; ++y)
                if (img.load(x, y - 1, z) < img.load(x, y, z)) img.store(x, y, z, 1 + img.load(x, y - 1, z));
            for (int y = ny - 2; y >= 0; --y)
                if (img.load(x, y + 1, z) < img.load(x, y, z)) img.store(x, y, z, 1 + img.load(x, y + 1, z));
        }
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            for (int z = 1; z < nz; ++z)
                if (img.load(x, y, z - 1) < img.load(x, y, z)) img.store(x, y, z, 1 + img.load(x, y, z - 1));
            for (int z = nz - 2; z >= 0; --z)
                if (img.load(x, y, z + 1) < img.load(x, y, z)) img.store(x, y, z, 1 + img.load(x, y, z + 1));
        }
}

extern "C" void vxo_distance_field_brute(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz,
                                         uint8_t* df) {
    const int maxd = (nx + ny + nz) < 254 ? (nx + ny + nz) : 254;
    std::vector<int> sx, sy_, sz_;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                if (blocks[x + (size_t)y * nx + (size_t)z * nx * ny]) { sx.push_back(x); sy_.push_back(y); sz_.push_back(z); }
    const size_t ns = sx.size();
#pragma omp parallel for num_threads(nthreads())
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                int best = maxd;
                for (size_t s = 0; s < ns; ++s) {
                    int d = abs(x - sx[s]) + abs(y - sy_[s]) + abs(z - sz_[s]);
                    if (d < best) best = d;
                }
                df[x + (size_t)y * nx + (size_t)z * nx * ny] = (uint8_t)best;
            }
}

/* ------------------------------------------------------------------------------------------
 * Traversal
 * ---------------------------------------------------------------------------------------- */

#include "vxo_grid.h"
extern "C" void vxo_step_table(int32_t* out256) {
    for (int k = 0; k < 256; ++k) out256[k] = euclidean_step(k);
}

/* VoxelTraversalDF (InitialRayTraceFrag.glsl:307-374; identical clones in
 * ShadowRayTraceFrag.glsl:222-289, DiffuseRayTraceFrag.glsl:1129-1196,
 * ReflectionTraceFrag.glsl:1088-1155 with the iteration cap as the only difference).        */
extern "C" float vxo_traverse(const vxo_world* w, const float origin0[3], const float dir[3],
                              int32_t max_iter, vxo_hit* hit) {
    v3 initial_origin = V3(origin0[0], origin0[1], origin0[2]);
    v3 origin = initial_origin;
    v3 direction = V3(dir[0], dir[1], dir[2]);
    bool Intersection = false;
    int MinIdx = 0;
    i3 RaySign = {gsign(direction.x), gsign(direction.y), gsign(direction.z)};
    i3 Step01 = {(1 + RaySign.x) >> 1, (1 + RaySign.y) >> 1, (1 + RaySign.z) >> 1};
    int iters = 0, dda = 0;

    for (int itr = 0; itr < max_iter; ++itr) {
        int lx = cvt_floor(origin.x), ly = cvt_floor(origin.y), lz = cvt_floor(origin.z);
        if (!in_volume(w, lx, ly, lz)) {
            Intersection = false;
            break;
        }
        iters++;
        int k = w->df[lx + (size_t)ly * w->nx + (size_t)lz * w->nx * w->ny];
        int Euclidean = euclidean_step(k);
        if (Euclidean == 0) break;
        if (Euclidean == 1) {
            dda++;
            i3 G = {cvt_trunc(origin.x), cvt_trunc(origin.y), cvt_trunc(origin.z)};
            v3 W = origin - V3((float)G.x, (float)G.y, (float)G.z);
            v3 inv = V3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
            v3 DF = (V3((float)Step01.x, (float)Step01.y, (float)Step01.z) - W) * inv;
            MinIdx = (DF.x < DF.y && RaySign.x != 0) ? ((DF.x < DF.z || RaySign.z == 0) ? 0 : 2)
                                                     : ((DF.y < DF.z || RaySign.z == 0) ? 1 : 2);
            idx(G, MinIdx) += idx(RaySign, MinIdx);
            W = W + direction * idx(DF, MinIdx);
            idx(W, MinIdx) = (float)(1 - idx(Step01, MinIdx));
            origin = V3((float)G.x, (float)G.y, (float)G.z) + W;
            idx(origin, MinIdx) += (float)idx(RaySign, MinIdx) * 0.0001f;
            Intersection = true;
        } else {
            origin = origin + (float)(Euclidean - 1) * direction;
        }
    }

    float t = -1.0f;
    int block = 0;
    v3 normal = V3(0.0f);
    if (Intersection) {
        idx(normal, MinIdx) = (float)(-idx(RaySign, MinIdx));
        block = get_voxel(w, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
        t = block > 0 ? distance(origin, initial_origin) : -1.0f;
    }
    if (hit) {
        hit->t = t;
        hit->normal[0] = normal.x; hit->normal[1] = normal.y; hit->normal[2] = normal.z;
        hit->end[0] = origin.x; hit->end[1] = origin.y; hit->end[2] = origin.z;
        hit->block = block;
        hit->intersection = Intersection;
        hit->min_idx = MinIdx;
        hit->iterations = iters;
        hit->dda_steps = dda;
    }
    return t;
}

/* vxo_traverse over n rays (origins / dirs: 3*n floats), scanline-parallel like the passes */
extern "C" void vxo_traverse_batch(const vxo_world* w, const float* origins, const float* dirs, int32_t n, int32_t max_iter,
                                   vxo_hit* hits) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t i = 0; i < n; ++i) vxo_traverse(w, origins + 3 * (size_t)i, dirs + 3 * (size_t)i, max_iter, hits + i);
}

/* World::RaycastDetect (Core/World.cpp:496-546).  The bounds test excludes index 0 as well as the far faces
 * (`<= 0`, :512-513); GetBlock takes uint16_t coordinates (World.h:46-49) but is only reached inside the bounds. */
extern "C" void vxo_raycast_detect(const vxo_world* w, const float pos[3], const float dir[3], int32_t out8[8]) {
    v3 position = V3(pos[0], pos[1], pos[2]);
    const v3 direction = V3(dir[0], dir[1], dir[2]);
    const v3 sign = V3(direction.x > 0.0f ? 1.0f : 0.0f, direction.y > 0.0f ? 1.0f : 0.0f, direction.z > 0.0f ? 1.0f : 0.0f);
    for (int k = 0; k < 8; ++k) out8[k] = k < 4 ? -1 : 0;
    for (int i = 0; i < 48; ++i) {  /* block reach */
        v3 tvec = (V3(floorf(position.x + sign.x), floorf(position.y + sign.y), floorf(position.z + sign.z)) - position) / direction;
        float t = std::min(tvec.x, std::min(tvec.y, tvec.z));
        position = position + direction * (t + 0.001f);
        int fx = (int)floorf(position.x), fy = (int)floorf(position.y), fz = (int)floorf(position.z);
        if (!(fx >= w->nx || fy >= w->ny || fz >= w->nz || fx <= 0 || fy <= 0 || fz <= 0)) {
            int b = w->blocks[(int)position.x + (size_t)(int)position.y * w->nx + (size_t)(int)position.z * w->nx * w->ny];
            if (b != 0) {
                float n[3];
                for (int j = 0; j < 3; ++j) {
                    n[j] = (t == idx(tvec, j)) ? 1.0f : 0.0f;
                    if (idx(sign, j) != 0.0f) n[j] = -n[j];
                }
                position = V3(floorf(position.x), floorf(position.y), floorf(position.z));
                /* the second bounds test (:533-537) repeats the first on the floored position: never fails here */
                out8[0] = (int)position.x; out8[1] = (int)position.y; out8[2] = (int)position.z;
                out8[3] = w->blocks[out8[0] + (size_t)out8[1] * w->nx + (size_t)out8[2] * w->nx * w->ny];
                out8[4] = (int)n[0]; out8[5] = (int)n[1]; out8[6] = (int)n[2];
                out8[7] = 1;
                return;
            }
        }
    }
}
extern "C" void vxo_raycast_detect_batch(const vxo_world* w, const float* pos, const float* dir, int32_t n, int32_t* out8) {
#pragma omp parallel for schedule(static, 256)
    for (int32_t i = 0; i < n; ++i) vxo_raycast_detect(w, pos + 3 * (size_t)i, dir + 3 * (size_t)i, out8 + 8 * (size_t)i);
}

/* Plain Amanatides–Woo grid walk (the style of Shaders/Implementations/DDA/DDA.glsl:134-253),
 * double precision, used only as an independent cross-check of vxo_traverse.                  */
extern "C" int32_t vxo_plain_dda(const vxo_world* w, const float o[3], const float d[3], int32_t max_steps,
                                 int32_t voxel[3]) {
    double p[3] = {o[0], o[1], o[2]};
    int v[3], step[3];
    double tmax[3], tdelta[3];
    const int dims[3] = {w->nx, w->ny, w->nz};
    for (int i = 0; i < 3; ++i) {
        v[i] = (int)floor(p[i]);
        step[i] = d[i] > 0 ? 1 : (d[i] < 0 ? -1 : 0);
        if (step[i] != 0) {
            double next = step[i] > 0 ? (v[i] + 1.0) : (double)v[i];
            tmax[i] = (next - p[i]) / d[i];
            tdelta[i] = fabs(1.0 / d[i]);
        } else {
            tmax[i] = INFINITY;
            tdelta[i] = INFINITY;
        }
    }
    for (int s = 0; s < max_steps; ++s) {
        int a = (tmax[0] < tmax[1]) ? ((tmax[0] < tmax[2]) ? 0 : 2) : ((tmax[1] < tmax[2]) ? 1 : 2);
        v[a] += step[a];
        tmax[a] += tdelta[a];
        for (int i = 0; i < 3; ++i)
            if (v[i] < 0 || v[i] >= dims[i]) return 0;
        if (w->blocks[v[0] + (size_t)v[1] * w->nx + (size_t)v[2] * w->nx * w->ny]) {
            voxel[0] = v[0]; voxel[1] = v[1]; voxel[2] = v[2];
            return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Primary pass
 * ---------------------------------------------------------------------------------------- */

/* GetNormalID (InitialRayTraceFrag.glsl:140-185): exact vec3 compare, fall-through 0.0 */
static inline float normal_id(v3 n) {
    if (n.x == 0.0f && n.y == 0.0f && n.z == 1.0f) return 0.0f / 10.0f;
    if (n.x == 0.0f && n.y == 0.0f && n.z == -1.0f) return 1.0f / 10.0f;
    if (n.x == 0.0f && n.y == 1.0f && n.z == 0.0f) return 2.0f / 10.0f;
    if (n.x == 0.0f && n.y == -1.0f && n.z == 0.0f) return 3.0f / 10.0f;
    if (n.x == -1.0f && n.y == 0.0f && n.z == 0.0f) return 4.0f / 10.0f;
    if (n.x == 1.0f && n.y == 0.0f && n.z == 0.0f) return 5.0f / 10.0f;
    return 0.0f;
}

/* GetRayStuff (InitialRayTraceFrag.glsl:398-416) / GetRayDirectionAt (ShadowRayTraceFrag.glsl:303-308) */
static inline v3 ray_direction_at(const float* inv_view, const float* inv_proj, v2 screenspace) {
    v4 clip = V4(screenspace.x * 2.0f - 1.0f, screenspace.y * 2.0f - 1.0f, -1.0f, 1.0f);
    v4 e = mat4_mul(inv_proj, clip);
    v4 eye = V4(e.x, e.y, -1.0f, 0.0f);
    v4 r = mat4_mul(inv_view, eye);
    return V3(r.x, r.y, r.z);
}

/* IntersectBox (InitialRayTraceFrag.glsl:418-432) */
static inline v2 intersect_box(v3 ro, v3 invrd, v3 rad) {
    v3 m = invrd;
    v3 n = m * ro;
    v3 k = V3(fabsf(m.x), fabsf(m.y), fabsf(m.z)) * rad;
    v3 t1 = -n - k;
    v3 t2 = -n + k;
    float tN = gmax(gmax(t1.x, t1.y), t1.z);
    float tF = gmin(gmin(t2.x, t2.y), t2.z);
    if (tN > tF || tF < 0.0f) return V2(-1.0f, -1.0f);
    return V2(tN, tF);
}

static inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}

static const vxo_scene* g_alpha_scene = nullptr;
extern "C" void vxo_bind_alpha_scene(const vxo_scene* s) { g_alpha_scene = s; }
extern "C" float vxo_alpha_g_k(float fov_degrees, int32_t width) {
    /* radians(x) = x * pi/180 as a float constant; tanf from libm (the CUDA library computes g_K on the host with
     * the same expression, so the device never evaluates tan) */
    return 1.0f / (tanf((fov_degrees * 0.01745329251994329577f) / (2.0f * (float)width)) * 2.0f);
}

/* main() (InitialRayTraceFrag.glsl:434-496); attachment formats Pipeline.cpp:1142 */
extern "C" void vxo_initial_trace(const vxo_world* w, const vxrt_primary_params* p, uint16_t* t_half,
                                  uint8_t* normal_u8, uint8_t* block_u8, float* inv_t, float* t32,
                                  vxrt_trace_stats* stats) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    const v3 half = V3((float)w->nx / 2.0f, (float)w->ny / 2.0f, (float)w->nz / 2.0f);
    uint64_t s_rays = 0, s_it = 0, s_dda = 0, s_hits = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(nthreads()) reduction(+ : s_rays, s_it, s_dda, s_hits)
    for (int py = r0; py < r1; ++py) {
        for (int px = 0; px < W; ++px) {
            /* v_TexCoords: interpolated quad texcoord == pixel centre / dims */
            v2 screenspace = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            v2 TexelSize = V2(1.0f / (float)W, 1.0f / (float)H);
            if (p->jitter_on) {
                screenspace.x -= p->jitter[0] * TexelSize.x;
                screenspace.y -= p->jitter[1] * TexelSize.y;
            }
            v3 rD = ray_direction_at(p->inv_view, p->inv_projection, screenspace);
            v3 ro = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
            v3 dir = normalize(rD);
            float AddT = 0.0f;
            v2 IntBox = intersect_box(ro - half, V3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z), half);
            if (IntBox.x > 0.0f) {
                AddT = IntBox.x + 0.5f;
                ro = ro + dir * AddT;
            }
            vxo_hit h;
            float o[3] = {ro.x, ro.y, ro.z}, d[3] = {dir.x, dir.y, dir.z};
            float t;
            if (p->alpha_test && g_alpha_scene) {
                const float viewer[3] = {p->inv_view[12], p->inv_view[13], p->inv_view[14]};
                t = vxo_traverse_alpha(g_alpha_scene, o, d, p->render_distance, viewer, vxo_alpha_g_k(p->fov, W), 0, &h);
            } else {
                t = vxo_traverse(w, o, d, p->render_distance, &h);
            }
            /* id is the texel value block/255; `id > 0` <=> block > 0.  On a miss id/normal are
             * uninitialised in the shader but t = -1 makes `intersect` false regardless.        */
            bool intersect = t > 0.0f && h.block > 0;
            t += AddT * (intersect ? 1.0f : 0.0f);
            float o_Normal = intersect ? normal_id(V3(h.normal[0], h.normal[1], h.normal[2])) : 1.0f;
            size_t i = (size_t)py * W + px;
            t_half[i] = float_to_half(t);
            if (t32) t32[i] = t;
            inv_t[i] = 1.0f / t;
            normal_u8[i] = float_to_unorm8(o_Normal);
            block_u8[i] = intersect ? float_to_unorm8(unorm8_to_float(h.block)) : 0;
            s_rays += 1; s_it += h.iterations; s_dda += h.dda_steps; s_hits += intersect ? 1 : 0;
        }
    }
    if (stats) { stats->rays += s_rays; stats->iterations += s_it; stats->dda_steps += s_dda; stats->hits += s_hits; }
}

/* ------------------------------------------------------------------------------------------
 * Texture reads of FBO attachments
 * ---------------------------------------------------------------------------------------- */

static inline int wrap_repeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

/* texture(sampler2D R16F, uv): GL_LINEAR, GL_REPEAT (Framebuffer.cpp:64-68). Weights in full float. */
static inline float sample_r16f_bilinear(const uint16_t* img, int w, int h, v2 uv) {
    float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = wrap_repeat(cvt_floor(fu), w), j0 = wrap_repeat(cvt_floor(fv), h);
    int i1 = wrap_repeat(i0 + 1, w), j1 = wrap_repeat(j0 + 1, h);
    float t00 = half_to_float(img[(size_t)j0 * w + i0]), t10 = half_to_float(img[(size_t)j0 * w + i1]);
    float t01 = half_to_float(img[(size_t)j1 * w + i0]), t11 = half_to_float(img[(size_t)j1 * w + i1]);
    float top = t00 * (1.0f - a) + t10 * a;
    float bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
/* texture(sampler2D R8, uv): GL_NEAREST, GL_REPEAT */
static inline float sample_r8_nearest(const uint8_t* img, int w, int h, v2 uv) {
    int i = wrap_repeat(cvt_floor(uv.x * (float)w), w), j = wrap_repeat(cvt_floor(uv.y * (float)h), h);
    return unorm8_to_float(img[(size_t)j * w + i]);
}

/* GetNormalFromID (ShadowRayTraceFrag.glsl:317-328) */
static inline v3 normal_from_id(float n) {
    static const v3 Normals[6] = {{0.0f, 0.0f, 1.0f}, {0.0f, 0.0f, -1.0f}, {0.0f, 1.0f, 0.0f},
                                  {0.0f, -1.0f, 0.0f}, {-1.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f}};
    int i = cvt_round(n * 10.0f);
    if (i > 5) return V3(1.0f, 1.0f, 1.0f);
    return Normals[i];
}

/* SampleCone (ShadowRayTraceFrag.glsl:388-398) */
static inline v3 sample_cone(v2 Xi, float CosThetaMax) {
    const float PI = 3.14159265359f;
    float CosTheta = (1.0f - Xi.x) + Xi.x * CosThetaMax;
    float SinTheta = sqrtf(1.0f - CosTheta * CosTheta);
    float phi = Xi.y * PI * 2.0f;
    return V3(SinTheta * cosf(phi), SinTheta * sinf(phi), CosTheta);
}

/* main() (ShadowRayTraceFrag.glsl:414-513); v_RayOrigin = u_VertInverseView[3] (FBOVert.glsl:21) */
extern "C" void vxo_shadow_trace(const vxo_world* w, const vxrt_shadow_params* p, const uint16_t* g_t_half,
                                 const uint8_t* g_normal_u8, int32_t gw, int32_t gh,
                                 const uint8_t* blue_rgba, int32_t bw, int32_t bh, uint8_t* shadow_u8,
                                 uint16_t* transversal_half, vxrt_trace_stats* stats) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    uint64_t s_rays = 0, s_it = 0, s_dda = 0, s_hits = 0;
    const v3 cam = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
#pragma omp parallel for schedule(dynamic, 2) num_threads(nthreads()) reduction(+ : s_rays, s_it, s_dda, s_hits)
    for (int py = r0; py < r1; ++py) {
        for (int px = 0; px < W; ++px) {
            size_t i = (size_t)py * W + px;
            v2 tc = V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
            tc = tc + V2(p->halton[0], p->halton[1]) * V2(1.0f / (float)W, 1.0f / (float)H);
            /* GetPositionAt (:310-314) */
            float Dist = sample_r16f_bilinear(g_t_half, gw, gh, tc);
            v3 P = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            if (Dist < 0.0f) {
                shadow_u8[i] = float_to_unorm8(0.0f);
                transversal_half[i] = float_to_half(64.0f);
                continue;
            }
            v3 L = V3(p->light_direction[0], p->light_direction[1], p->light_direction[2]);
            v3 RayDirection = L;
            if (p->soft_shadows) {
                /* :456-462 */
                int n = p->current_frame % 1024;
                /* GLSL int products wrap (two's complement), then convert to float */
                v2 q = V2((float)(int32_t)((uint32_t)n * 12664745u), (float)(int32_t)((uint32_t)n * 9560333u));
                v2 off = V2(gfract(q.x / 16777216.0f) * 1024.0f, gfract(q.y / 16777216.0f) * 1024.0f);
                /* gl_FragCoord.xy + ivec2(floor(off)) is a vec2, converted to ivec2 (trunc), then % size */
                int sx = cvt_trunc(((float)px + 0.5f) + (float)cvt_trunc(floorf(off.x))) % bw;
                int sy = cvt_trunc(((float)py + 0.5f) + (float)cvt_trunc(floorf(off.y))) % bh;
                const uint8_t* tx = blue_rgba + 4 * ((size_t)sy * bw + sx);
                v2 Xi = V2(unorm8_to_float(tx[0]), unorm8_to_float(tx[1]));
                v3 T = normalize(cross(L, V3(0.0f, 1.0f, 1.0f)));
                v3 B = cross(T, L);
                RayDirection = mat3_mul(T, B, L, sample_cone(Xi, 0.9999505604617f));
            }
            v3 N = normal_from_id(sample_r8_nearest(g_normal_u8, gw, gh, tc));
            float NDotL = dot(N, RayDirection);
            if (NDotL <= 0.01f) {
                shadow_u8[i] = float_to_unorm8(1.0f);
                transversal_half[i] = float_to_half(1.0f / 100.0f);
                continue;
            }
            v3 Bias = N * V3(0.1f);
            v3 o = P + Bias;
            int block_at = get_voxel(w, cvt_floor(o.x), cvt_floor(o.y), cvt_floor(o.z));
            float T = -1.0f;
            if (Dist > 0.0f) {
                vxo_hit h;
                float oo[3] = {o.x, o.y, o.z}, dd[3] = {RayDirection.x, RayDirection.y, RayDirection.z};
                if (p->alpha_test && g_alpha_scene) {
                    const float viewer[3] = {p->inv_view[12], p->inv_view[13], p->inv_view[14]};
                    T = vxo_traverse_alpha(g_alpha_scene, oo, dd, p->max_iterations, viewer, vxo_alpha_g_k(p->fov, W), 1, &h);
                } else {
                    T = vxo_traverse(w, oo, dd, p->max_iterations, &h);
                }
                s_rays += 1; s_it += h.iterations; s_dda += h.dda_steps; s_hits += (T > 0.0f) ? 1 : 0;
            }
            shadow_u8[i] = float_to_unorm8((T > 0.0f || block_at > 0) ? 1.0f : 0.0f);
            float tr = gclamp(T / 100.0f, 0.00001f, 196.0f);
            if (T < 0.0f) tr = 4.25f / 100.0f;
            transversal_half[i] = float_to_half(tr);
        }
    }
    if (stats) { stats->rays += s_rays; stats->iterations += s_it; stats->dda_steps += s_dda; stats->hits += s_hits; }
}
