// TEST INFRASTRUCTURE ONLY — CPU restatement of the light propagation volume flood fill.
// Follows Core/VolumetricFloodFill.cpp (FIFO queues of light / removal nodes over two byte volumes: light level and the block type that
// lit the voxel) and the call sequences around it: start-up Core/Pipeline.cpp:1602-1611 and World::RepropogateLPV_ Core/World.cpp:554-572
// (clear, seed every light location, propagate), and the LPV half of the block edit in World::Raycast Core/World.cpp:273-333 (place),
// :395-446 (break), :482-485 (4 x depropagate + propagate).  Pinned against that code compiled in oracle/_ref/libvxrt_ref_world.so
// (tests/test_oracle_lpv.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
#include <stdint.h>
#include <string.h>

#include <deque>

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
