// TEST INFRASTRUCTURE ONLY — CPU restatement of the light propagation volume flood fill.
// Follows Core/VolumetricFloodFill.cpp (FIFO queues of light / removal nodes over two byte volumes: light level and the block type that
// lit the voxel) and the call sequences around it: start-up Core/Pipeline.cpp:1602-1611 and World::RepropogateLPV_ Core/World.cpp:554-572
// (clear, seed every light location, propagate), and the LPV half of the block edit in World::Raycast Core/World.cpp:273-333 (place),
// :395-446 (break), :482-485 (4 x depropagate + propagate).  Pinned against that code compiled in oracle/_ref/libvxrt_ref_world.so
// (tests/test_oracle_lpv.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
#include <stdint.h>
#include <string.h>

#include <deque>

#include "vxo_math.h"

namespace {

struct Node { int x, y, z, light; };

struct Lpv {
    const uint8_t* blocks;
    uint8_t *level, *color;
    int nx, ny, nz;
    std::deque<Node> light_q, removal_q;

    // InVoxelVolume (VolumetricFloodFill.cpp:22-30): the planes x = 0, y = 0, z = 0 are outside
    bool inside(int x, int y, int z) const { return x > 0 && y > 0 && z > 0 && x < nx && y < ny && z < nz; }
    size_t at(int x, int y, int z) const { return (size_t)x + (size_t)y * nx + (size_t)z * nx * ny; }
    int get_level(int x, int y, int z) const { return inside(x, y, z) ? level[at(x, y, z)] : 0; }   // GetLightValue :125-138
    int get_color(int x, int y, int z) const { return inside(x, y, z) ? color[at(x, y, z)] : 0; }   // GetBlockTypeLightValue :140-154
    void set(int x, int y, int z, int v, int b) {                                                   // SetLightValue :156-170
        if (!inside(x, y, z)) return;
        color[at(x, y, z)] = (uint8_t)b;
        level[at(x, y, z)] = (uint8_t)v;
    }
    // AddLightToVolume :190-205
    void add_light(int x, int y, int z, int block, int limit) {
        set(x, y, z, limit > 8 ? 8 : limit, block);
        light_q.push_back({x, y, z, 0});
    }
    // PropogateVolume :247-326: neighbours in the order +x -x +y -y -z +z; the node's level and block type are read when it is popped
    void propagate() {
        static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}};
        while (!light_q.empty()) {
            const Node n = light_q.front();
            light_q.pop_front();
            const int cur = get_level(n.x, n.y, n.z), type = get_color(n.x, n.y, n.z);
            for (int k = 0; k < 6; ++k) {
                const int x = n.x + d[k][0], y = n.y + d[k][1], z = n.z + d[k][2];
                if (!inside(x, y, z)) continue;
                if (blocks[at(x, y, z)] == 0 && get_level(x, y, z) + 2 < cur) {
                    set(x, y, z, cur - 1, type);
                    light_q.push_back({x, y, z, 0});
                }
            }
        }
    }
    // DepropogateVolume :328-468: neighbours in the order +x -x +y -y +z -z; the node carries the level it had when it was queued
    void depropagate() {
        static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        while (!removal_q.empty()) {
            const Node n = removal_q.front();
            removal_q.pop_front();
            for (int k = 0; k < 6; ++k) {
                const int x = n.x + d[k][0], y = n.y + d[k][1], z = n.z + d[k][2];
                if (!inside(x, y, z)) continue;
                const int nl = get_level(x, y, z), nb = get_color(x, y, z);
                if (nl != 0 && nl < n.light) {
                    set(x, y, z, 0, nb);
                    removal_q.push_back({x, y, z, nl});
                } else if (nl >= n.light) {
                    light_q.push_back({x, y, z, 0});
                }
            }
        }
    }
};

}  // namespace

extern "C" {

// Start-up (Pipeline.cpp:1602-1611) and World::RepropogateLPV_ (World.cpp:554-572): both volumes cleared, every light location seeded with
// min(limit, 8) and the block at it, then PropogateVolume (the reference calls it 3 or 4 times; the queue is empty after the first).
void vxo_lpv_repropagate(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const int32_t* lights_xyz, int32_t n_lights,
                         int32_t limit, uint8_t* level, uint8_t* color) {
    const size_t n = (size_t)nx * ny * nz;
    memset(level, 0, n);
    memset(color, 0, n);
    Lpv v{blocks, level, color, nx, ny, nz, {}, {}};
    for (int32_t i = 0; i < n_lights; ++i) {
        const int x = lights_xyz[3 * i], y = lights_xyz[3 * i + 1], z = lights_xyz[3 * i + 2];
        // World::GetBlock is unchecked; light locations come from the scan of the grid, so they are inside the array
        const int block = (x >= 0 && y >= 0 && z >= 0 && x < nx && y < ny && z < nz) ? blocks[v.at(x, y, z)] : 0;
        v.add_light(x, y, z, block, limit);
    }
    v.propagate();
}

// The LPV half of a block edit (World.cpp:273-333 place, :395-446 break, then :482-485).  `blocks` is the grid after the edit (neither
// AddLightToVolume nor DepropogateVolume read the grid, and SetBlock precedes the propagation), `block` the block placed (op 1) or the
// block that was broken (op 0), `emissive` whether that block has an emissive texture.
void vxo_lpv_edit(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, int32_t op, int32_t x, int32_t y, int32_t z, int32_t block,
                  int32_t emissive, int32_t limit, uint8_t* level, uint8_t* color) {
    Lpv v{blocks, level, color, nx, ny, nz, {}, {}};
    static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    if (op == 1) {
        v.removal_q.push_back({x, y, z, v.get_level(x, y, z)});
        for (int k = 0; k < 6; ++k) v.removal_q.push_back({x + d[k][0], y + d[k][1], z + d[k][2], v.get_level(x + d[k][0], y + d[k][1], z + d[k][2])});
        if (emissive) v.add_light(x, y, z, block, limit);
    } else {
        if (emissive) {
            v.removal_q.push_back({x, y, z, v.get_level(x, y, z)});
            v.set(x, y, z, 0, 0);
        }
        for (int k = 0; k < 6; ++k) v.light_q.push_back({x + d[k][0], y + d[k][1], z + d[k][2], 0});
    }
    for (int it = 0; it < 4; ++it) {
        v.depropagate();
        v.propagate();
    }
}

}  // extern "C"

/* ---- SampleLPVData (ReflectionTraceFrag.glsl:1516-1528) with SampleLPVColor (:1484-1487) and InterpolateLPVColorDithered (:1490-1509): the
 * light the reflection pass takes from the propagation volume at a point given in voxel units.  u_LPV is R8 unorm LINEAR CLAMP_TO_EDGE
 * (VolumetricFloodFill.cpp:41-48), u_LPVBlocks R8UI NEAREST CLAMP_TO_EDGE (:51-58), BlockAverageColorData the table of
 * PrecomputeAverageBlockColor.comp; the shader hard-codes the 384 x 128 x 384 volume resolution for the coordinate scale.  Block types above
 * 127 index past the table (clamp(BlockID, 0u, 128u), :1486): read as 0 here, like the shim's SSBO view. ---- */
namespace {
using namespace vxo;
struct LpvTex {
    const uint8_t *level, *type;
    const float* avg;   /* 128 x 4 */
    int nx, ny, nz;
    static int clampi(int i, int n) { return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); }
    float texel(int i, int j, int k) const { return unorm8_to_float(level[i + (size_t)j * nx + (size_t)k * nx * ny]); }
    float level_at(float x, float y, float z) const {   /* texture(u_LPV, UV).x */
        const float u = x * (float)nx - 0.5f, v = y * (float)ny - 0.5f, w = z * (float)nz - 0.5f;
        const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
        const float a = u - fu, b = v - fv, g = w - fw;
        const int i0 = clampi(cvt_floor(fu), nx), i1 = clampi(cvt_floor(fu) + 1, nx);
        const int j0 = clampi(cvt_floor(fv), ny), j1 = clampi(cvt_floor(fv) + 1, ny);
        const int k0 = clampi(cvt_floor(fw), nz), k1 = clampi(cvt_floor(fw) + 1, nz);
        float p[2];
        const int ks[2] = {k0, k1};
        for (int q = 0; q < 2; ++q) {
            const float top = texel(i0, j0, ks[q]) * (1.0f - a) + texel(i1, j0, ks[q]) * a;
            const float bot = texel(i0, j1, ks[q]) * (1.0f - a) + texel(i1, j1, ks[q]) * a;
            p[q] = top * (1.0f - b) + bot * b;
        }
        return p[0] * (1.0f - g) + p[1] * g;
    }
    v3 color_at(float x, float y, float z) const {   /* SampleLPVColor */
        const int i = clampi(cvt_floor(x * (float)nx), nx), j = clampi(cvt_floor(y * (float)ny), ny), k = clampi(cvt_floor(z * (float)nz), nz);
        const unsigned id = type[i + (size_t)j * nx + (size_t)k * nx * ny];
        if (id > 127u) return V3(0.0f, 0.0f, 0.0f);
        return V3(avg[4 * id], avg[4 * id + 1], avg[4 * id + 2]);
    }
};
}  // namespace

extern "C" void vxo_lpv_sample(const uint8_t* level, const uint8_t* block_type, int32_t nx, int32_t ny, int32_t nz, const float* avg512,
                               const float* points, int32_t n, const float dither[3], float* rgb_out) {
    const LpvTex t{level, block_type, avg512, nx, ny, nz};
    const float R[3] = {384.0f, 128.0f, 384.0f};   /* VolumeResolution, hard-coded in the shader */
    for (int32_t p = 0; p < n; ++p) {
        float UV[3], W0[3], W1[3];
        for (int c = 0; c < 3; ++c) {
            UV[c] = points[3 * p + c] * (1.0f / R[c]);
            const float F = gfract(UV[c] * R[c]);
            const float L = (F * (F - 1.0f) + 0.5f) / R[c];
            W0[c] = UV[c] - L; W1[c] = UV[c] + L;
        }
        const float level_v = t.level_at(UV[0], UV[1], UV[2]);
        /* the eight dithered taps in the order of the shader; DitherWeights = 1, GlobalDitherNoiseWeight = 2 */
        const float d[3] = {(dither[0] * 1.0f) * 2.0f, (dither[1] * 1.0f) * 2.0f, (dither[2] * 1.0f) * 2.0f};
        static const int sel[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 0, 1}, {0, 0, 1}};
        v3 sum = V3(0.0f, 0.0f, 0.0f);
        for (int k = 0; k < 8; ++k) {
            const float sgn = (k & 1) ? -1.0f : 1.0f;
            const float x = (sel[k][0] ? W1[0] : W0[0]), y = (sel[k][1] ? W1[1] : W0[1]), z = (sel[k][2] ? W1[2] : W0[2]);
            const v3 c = (k & 1) ? t.color_at(x - d[0], y - d[1], z - d[2]) : t.color_at(x + d[0], y + d[1], z + d[2]);
            (void)sgn;
            sum = k == 0 ? c : V3(sum.x + c.x, sum.y + c.y, sum.z + c.z);
        }
        const v3 col = V3(gmax(sum.x / 8.0f, 0.00000001f), gmax(sum.y / 8.0f, 0.00000001f), gmax(sum.z / 8.0f, 0.00000001f));
        const float s = level_v * 325.0f;
        const v3 Fi = V3(s * col.x, s * col.y, s * col.z);
        const float luma = (Fi.x * 0.2125f + Fi.y * 0.7154f) + Fi.z * 0.0721f;   /* dot(x, vec3(0.2125, 0.7154, 0.0721)) */
        rgb_out[3 * p] = gmix(luma, Fi.x, 0.5f); rgb_out[3 * p + 1] = gmix(luma, Fi.y, 0.5f); rgb_out[3 * p + 2] = gmix(luma, Fi.z, 0.5f);
    }
}
