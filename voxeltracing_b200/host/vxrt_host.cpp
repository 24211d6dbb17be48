// vxrt_host.cpp — host-side producers of the hot path's inputs (see vxrt_host.h).
#include "vxrt_host.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <random>
#include <vector>

namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 crs(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float dt(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 nrm(V3 a) { float s = 1.0f / sqrtf(dt(a, a)); return {a.x * s, a.y * s, a.z * s}; }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }

// ---- block ids = 1-based order of records in blockdb.txt (BlockDatabaseParser.cpp:31-42,371) ----
enum : uint8_t {
    B_AIR = 0, B_GRASS = 1, B_DIRT = 2, B_STONE = 3, B_COBBLE = 4, B_SAND = 5, B_OAK_LOG = 6, B_OAK_LEAVES = 7,
    B_BRICK = 8, B_GRAVEL = 11, B_LAMP = 12, B_METAL = 13, B_PLANKS = 20, B_MARBLE = 24, B_SEALANTERN = 26,
    B_GLOWSTONE = 27, B_IRON = 31, B_GOLD = 32
};

struct Grid {
    uint8_t* b;
    int nx, ny, nz;
    bool valid(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < nx && y < ny && z < nz; }
    void set(int x, int y, int z, uint8_t id) { if (valid(x, y, z)) b[(size_t)x + (size_t)y * nx + (size_t)z * nx * ny] = id; }
    uint8_t get(int x, int y, int z) const { return valid(x, y, z) ? b[(size_t)x + (size_t)y * nx + (size_t)z * nx * ny] : 0; }
    void box(int x0, int y0, int z0, int x1, int y1, int z1, uint8_t id) {
        for (int z = z0; z <= z1; ++z) for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x) set(x, y, z, id);
    }
};

// ---- seeded lattice value noise (our own; the reference seeds FastNoise from time(0), so its
// terrain is not reproducible anyway — WorldGenerator.cpp:213-219) -------------------------------
inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
    uint32_t h = x * 0x8da6b343u ^ y * 0xd8163841u ^ s * 0xcb1ab31fu;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
    return h;
}
inline float lattice(int x, int y, uint32_t s) { return (float)(hash3((uint32_t)x, (uint32_t)y, s) & 0xffffffu) / 8388607.5f - 1.0f; }
inline float smooth(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
float value_noise(float x, float y, uint32_t s) {
    float fx = floorf(x), fy = floorf(y);
    int ix = (int)fx, iy = (int)fy;
    float tx = smooth(x - fx), ty = smooth(y - fy);
    float a = lattice(ix, iy, s), b = lattice(ix + 1, iy, s), c = lattice(ix, iy + 1, s), d = lattice(ix + 1, iy + 1, s);
    float top = a + (b - a) * tx, bot = c + (d - c) * tx;
    return top + (bot - top) * ty;
}
float fbm(float x, float y, int octaves, uint32_t s) {
    float sum = 0.0f, amp = 1.0f, norm = 0.0f;
    for (int o = 0; o < octaves; ++o) {
        sum += amp * value_noise(x, y, s + 101u * (uint32_t)o);
        norm += amp;
        amp *= 0.5f; x *= 2.0f; y *= 2.0f;
    }
    return sum / norm;
}

void column(Grid& g, int x, int z, int top, bool sand) {
    for (int y = 0; y < top && y < g.ny; ++y) {
        uint8_t id;
        if (sand) id = (y >= top - 8) ? B_SAND : B_STONE;
        else id = (y >= top - 1) ? B_GRASS : ((y >= top - 5) ? B_DIRT : B_STONE);
        g.set(x, y, z, id);
    }
}
void tree(Grid& g, int x, int y, int z) {
    for (int i = 0; i < 6; ++i) g.set(x, y + i, z, B_OAK_LOG);
    const int cy = y + 7;
    for (int dz = -4; dz <= 4; ++dz) for (int dy = -4; dy <= 4; ++dy) for (int dx = -4; dx <= 4; ++dx) {
        float d2 = (float)(dx * dx + dy * dy + dz * dz);
        if (d2 <= 3.5f * 3.5f && g.get(x + dx, cy + dy, z + dz) == 0) g.set(x + dx, cy + dy, z + dz, B_OAK_LEAVES);
    }
}

}  // namespace

extern "C" {

void vxh_perspective(float fov_deg, float aspect, float zn, float zf, float* m) {
    const float t = tanf(radians(fov_deg) / 2.0f);
    memset(m, 0, 16 * sizeof(float));
    m[0] = 1.0f / (aspect * t);
    m[5] = 1.0f / t;
    m[10] = -(zf + zn) / (zf - zn);
    m[11] = -1.0f;
    m[14] = -(2.0f * zf * zn) / (zf - zn);
}

void vxh_look_at(const float* eye, const float* center, const float* up, float* m) {
    V3 e = {eye[0], eye[1], eye[2]}, c = {center[0], center[1], center[2]}, u0 = {up[0], up[1], up[2]};
    V3 f = nrm(sub(c, e)), s = nrm(crs(f, u0)), u = crs(s, f);
    m[0] = s.x; m[4] = s.y; m[8] = s.z;
    m[1] = u.x; m[5] = u.y; m[9] = u.z;
    m[2] = -f.x; m[6] = -f.y; m[10] = -f.z;
    m[3] = 0; m[7] = 0; m[11] = 0;
    m[12] = -dt(s, e); m[13] = -dt(u, e); m[14] = dt(f, e); m[15] = 1.0f;
}

// adjugate / determinant inverse of a column-major 4x4
void vxh_inverse(const float* a, float* o) {
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    float r = 1.0f / det;
    for (int i = 0; i < 16; ++i) o[i] = inv[i] * r;
}

void vxh_camera(const float* pos, float yaw, float pitch, float fov, float aspect, float* view, float* proj,
                float* inv_view, float* inv_proj) {
    // FpsCamera.cpp:66-70
    float front[3] = {cosf(radians(pitch)) * cosf(radians(yaw)), sinf(radians(pitch)), cosf(radians(pitch)) * sinf(radians(yaw))};
    float center[3] = {front[0] + pos[0], front[1] + pos[1], front[2] + pos[2]};
    const float up[3] = {0.0f, 1.0f, 0.0f};
    float v[16], p[16];
    vxh_look_at(pos, center, up, v);
    vxh_perspective(fov, aspect, 0.1f, 1000.0f, p);
    if (view) memcpy(view, v, sizeof(v));
    if (proj) memcpy(proj, p, sizeof(p));
    if (inv_view) vxh_inverse(v, inv_view);
    if (inv_proj) vxh_inverse(p, inv_proj);
}

static float halton(int prime, int index) {
    float r = 0.0f, f = 1.0f;
    int i = index;
    while (i > 0) {
        f /= (float)prime;
        r += f * (float)(i % prime);
        i = (int)floorf((float)i / (float)prime);
    }
    return r;
}
void vxh_taa_jitter(int32_t frame, float* out2) {
    int i = frame % 64;
    out2[0] = halton(2, i + 1);
    out2[1] = halton(3, i + 1);
}

void vxh_sun_direction(float sun_tick, float* sun3, float* moon3, float* stronger3) {
    // rotate (1,1,1) about +Z by 2*tick degrees (glm::rotate(mat4(1), angle, (0,0,1)) * vec4(1))
    float ang = radians(sun_tick * 2.0f);
    float c = cosf(ang), s = sinf(ang);
    V3 sun = {c * 1.0f - s * 1.0f, s * 1.0f + c * 1.0f, 1.0f};
    V3 moon = {-sun.x, -sun.y, sun.z};
    V3 strong = (-sun.y < 0.01f) ? sun : moon;
    sun = nrm(sun); moon = nrm(moon); strong = nrm(strong);
    if (sun3) { sun3[0] = sun.x; sun3[1] = sun.y; sun3[2] = sun.z; }
    if (moon3) { moon3[0] = moon.x; moon3[1] = moon.y; moon3[2] = moon.z; }
    if (stronger3) { stronger3[0] = strong.x; stronger3[1] = strong.y; stronger3[2] = strong.z; }
}

// plains: heightfield 8 + floor(40 * (0.5 + 0.5 * fbm)) with grass/dirt/stone or sand columns and
// optional trees — the shape of GenerateWorld's "plains" type (WorldGenerator.cpp:233-298).
void vxh_gen_plains(uint32_t seed, int32_t structures, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks) {
    Grid g = {blocks, nx, ny, nz};
    memset(blocks, 0, (size_t)nx * ny * nz);
    std::vector<int> tops((size_t)nx * nz);
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) {
            float h = fbm((float)x * 0.0154f, (float)z * 0.0154f, 6, seed * 7919u + 1u);
            int top = 8 + (int)floorf(40.0f * (0.5f + 0.5f * h));
            if (top > ny - 24) top = ny - 24;
            if (top < 2) top = 2;
            bool sand = value_noise((float)x * 0.011f, (float)z * 0.011f, seed * 7919u + 2u) > 0.45f;
            column(g, x, z, top, sand);
            tops[(size_t)z * nx + x] = sand ? -top : top;
        }
    if (structures) {
        std::mt19937 rng(seed * 2654435761u + 17u);
        std::vector<std::pair<int, int>> placed;
        int want = (nx * nz) / 700;
        for (int tries = 0; tries < want * 8 && (int)placed.size() < want; ++tries) {
            int x = 5 + (int)(rng() % (uint32_t)(nx - 10)), z = 5 + (int)(rng() % (uint32_t)(nz - 10));
            int top = tops[(size_t)z * nx + x];
            if (top <= 0) continue;  // no trees on sand
            bool ok = true;
            for (auto& p : placed) if ((p.first - x) * (p.first - x) + (p.second - z) * (p.second - z) <= 36) { ok = false; break; }
            if (!ok) continue;
            tree(g, x, top, z);
            placed.push_back({x, z});
        }
    }
}

// rooms: an enclosed stone building made of rooms with doorways, ceiling lamps, and metal / marble
// panels — the stand-in for 'Test Worlds/gi' (interior GI + rough reflections).
void vxh_gen_rooms(uint32_t seed, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks) {
    Grid g = {blocks, nx, ny, nz};
    memset(blocks, 0, (size_t)nx * ny * nz);
    std::mt19937 rng(seed * 2246822519u + 3u);
    const int floor_y = ny / 2 - 12, room = 24, wall_h = 14;
    g.box(0, 0, 0, nx - 1, floor_y, nz - 1, B_STONE);
    const int bx0 = nx / 2 - 3 * room, bx1 = nx / 2 + 3 * room, bz0 = nz / 2 - 3 * room, bz1 = nz / 2 + 3 * room;
    g.box(bx0, floor_y, bz0, bx1, floor_y, bz1, B_PLANKS);
    g.box(bx0, floor_y + wall_h + 1, bz0, bx1, floor_y + wall_h + 1, bz1, B_BRICK);  // ceiling
    for (int i = 0; i <= 6; ++i) {
        int x = bx0 + i * room, z = bz0 + i * room;
        g.box(x, floor_y + 1, bz0, x, floor_y + wall_h, bz1, (i % 2) ? B_BRICK : B_MARBLE);
        g.box(bx0, floor_y + 1, z, bx1, floor_y + wall_h, z, (i % 2) ? B_MARBLE : B_BRICK);
    }
    for (int rz = 0; rz < 6; ++rz)
        for (int rx = 0; rx < 6; ++rx) {
            int x0 = bx0 + rx * room, z0 = bz0 + rz * room;
            // doorways (interior walls only) and a window on the outer walls
            if (rx > 0) g.box(x0, floor_y + 1, z0 + room / 2 - 2, x0, floor_y + 7, z0 + room / 2 + 2, B_AIR);
            if (rz > 0) g.box(x0 + room / 2 - 2, floor_y + 1, z0, x0 + room / 2 + 2, floor_y + 7, z0, B_AIR);
            if (rx == 0) g.box(x0, floor_y + 5, z0 + 6, x0, floor_y + 10, z0 + room - 6, B_AIR);
            if (rz == 5) g.box(x0 + 6, floor_y + 5, z0 + room, x0 + room - 6, floor_y + 10, z0 + room, B_AIR);
            // ceiling lamp
            uint8_t lamp = (rng() % 3 == 0) ? B_SEALANTERN : ((rng() % 2) ? B_LAMP : B_GLOWSTONE);
            int lx = x0 + 6 + (int)(rng() % (uint32_t)(room - 12)), lz = z0 + 6 + (int)(rng() % (uint32_t)(room - 12));
            g.box(lx, floor_y + wall_h, lz, lx + 1, floor_y + wall_h, lz + 1, lamp);
            // furniture: metal / gold / iron blocks and a pillar
            uint8_t mats[4] = {B_METAL, B_IRON, B_GOLD, B_MARBLE};
            int n = 1 + (int)(rng() % 3);
            for (int k = 0; k < n; ++k) {
                int fx = x0 + 3 + (int)(rng() % (uint32_t)(room - 6)), fz = z0 + 3 + (int)(rng() % (uint32_t)(room - 6));
                int fh = 1 + (int)(rng() % 5);
                g.box(fx, floor_y + 1, fz, fx + (int)(rng() % 3), floor_y + fh, fz + (int)(rng() % 3), mats[rng() % 4]);
            }
        }
    // skylight in the central room pair
    g.box(nx / 2 - 6, floor_y + wall_h + 1, nz / 2 - 6, nx / 2 + 6, floor_y + wall_h + 1, nz / 2 + 6, B_AIR);
}

// town: plains terrain + a street grid of brick / cobble / plank houses with lamps and towers —
// the stand-in for the 'Medival' Minecraft import.
void vxh_gen_town(uint32_t seed, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks) {
    vxh_gen_plains(seed, 1, nx, ny, nz, blocks);
    Grid g = {blocks, nx, ny, nz};
    std::mt19937 rng(seed * 3266489917u + 5u);
    const int lot = 28;
    for (int lz = 1; lz * lot + lot < nz - 4; ++lz)
        for (int lx = 1; lx * lot + lot < nx - 4; ++lx) {
            if (rng() % 4 == 0) continue;
            int x0 = lx * lot + 3, z0 = lz * lot + 3;
            int w = 10 + (int)(rng() % 12), d = 10 + (int)(rng() % 12), h = 6 + (int)(rng() % 14);
            if (rng() % 9 == 0) h += 24;  // tower
            // ground level = max terrain height under the footprint
            int base = 0;
            for (int z = z0; z <= z0 + d; ++z) for (int x = x0; x <= x0 + w; ++x) {
                int y = ny - 1;
                while (y > 0 && g.get(x, y, z) == 0) --y;
                uint8_t b = g.get(x, y, z);
                if (b == B_OAK_LEAVES || b == B_OAK_LOG) continue;
                if (y > base) base = y;
            }
            if (base + h + 2 >= ny) continue;
            uint8_t wall = (uint8_t[]){B_BRICK, B_COBBLE, B_PLANKS, B_STONE, B_MARBLE}[rng() % 5];
            g.box(x0, 0, z0, x0 + w, base, z0 + d, B_COBBLE);                    // foundation
            g.box(x0, base + 1, z0, x0 + w, base + h, z0 + d, wall);             // shell
            g.box(x0 + 1, base + 1, z0 + 1, x0 + w - 1, base + h - 1, z0 + d - 1, B_AIR);  // interior
            g.box(x0 + w / 2 - 1, base + 1, z0, x0 + w / 2 + 1, base + 3, z0, B_AIR);      // door
            for (int wy = base + 3; wy + 2 < base + h; wy += 5) {                          // windows
                g.box(x0, wy, z0 + 3, x0, wy + 1, z0 + d - 3, B_AIR);
                g.box(x0 + w, wy, z0 + 3, x0 + w, wy + 1, z0 + d - 3, B_AIR);
            }
            g.set(x0 + w / 2, base + h - 1, z0 + d / 2, B_LAMP);                            // ceiling lamp
            g.set(x0 + w / 2, base + 4, z0 - 1, B_GLOWSTONE);                               // door lantern
            if (rng() % 2) g.box(x0 + 2, base + 1, z0 + 2, x0 + 3, base + 2, z0 + 3, (rng() % 2) ? B_IRON : B_GOLD);
        }
    // streets
    for (int lz = 1; lz * lot < nz; ++lz) for (int x = 0; x < nx; ++x) for (int dz = 0; dz < 3; ++dz) {
        int z = lz * lot + dz, y = ny - 1;
        if (z >= nz) continue;
        while (y > 0 && g.get(x, y, z) == 0) --y;
        if (g.get(x, y, z) == B_GRASS || g.get(x, y, z) == B_SAND) g.set(x, y, z, B_GRAVEL);
    }
}

void vxh_random_edits(uint32_t seed, int32_t n, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks, int32_t* e) {
    std::mt19937 rng(seed);
    for (int i = 0; i < n; ++i) {
        int x = 1 + (int)(rng() % (uint32_t)(nx - 2)), y = 1 + (int)(rng() % (uint32_t)(ny - 2)), z = 1 + (int)(rng() % (uint32_t)(nz - 2));
        size_t idx = (size_t)x + (size_t)y * nx + (size_t)z * nx * ny;
        uint8_t id = blocks[idx] ? 0 : B_STONE;
        blocks[idx] = id;
        e[4 * i] = x; e[4 * i + 1] = y; e[4 * i + 2] = z; e[4 * i + 3] = id;
    }
}

int32_t vxh_world_save(const char* path, const uint8_t* blocks, int64_t nbytes) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    size_t w = fwrite(blocks, 1, (size_t)nbytes, f);
    fclose(f);
    return w == (size_t)nbytes ? 0 : -2;
}
int32_t vxh_world_load(const char* path, uint8_t* blocks, int64_t nbytes) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    size_t r = fread(blocks, 1, (size_t)nbytes, f);
    fclose(f);
    return r == (size_t)nbytes ? 0 : -2;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Block database: same record grammar and precedence as the engine's parser
// (Core/BlockDatabaseParser.cpp:44-417): records are `{` ... `}` line blocks of `Key : value` fields;
// face-specific keys assign, the generic key fills whatever is still empty; ids count records from 1.
// ---------------------------------------------------------------------------------------------------
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

namespace {
struct FaceSet { std::string f[6]; };  // front, back, top, bottom, left, right
struct BlockRec {
    std::string name, emissive;
    FaceSet maps[3];  // albedo, normal, pbr
    bool transparent = false, sss = false;
    std::vector<int> mc_ids;
    int id = 0;
};
std::string field_value(const std::string& field) {
    size_t loc = field.find(':');
    std::string s = loc == std::string::npos ? field : field.substr(loc + 1);
    s.erase(std::remove_if(s.begin(), s.end(), [](unsigned char c) { return isspace(c); }), s.end());
    return s;
}
bool has(const std::string& f, const char* k) { return f.find(k) != std::string::npos; }
}  // namespace

struct vxh_blockdb {
    std::vector<BlockRec> blocks;                  // index = id - 1
    std::map<std::string, int> by_name;
    std::vector<std::string> layers[4];            // sorted unique paths per kind
    int layer_of(int kind, const std::string& p) const {
        auto it = std::lower_bound(layers[kind].begin(), layers[kind].end(), p);
        return (it != layers[kind].end() && *it == p) ? (int)(it - layers[kind].begin()) : -1;
    }
};

extern "C" {

vxh_blockdb* vxh_blockdb_parse(const char* path) {
    std::ifstream in(path);
    if (!in.good()) return nullptr;
    vxh_blockdb* db = new vxh_blockdb();
    static const char* kind_key[3] = {"Albedo", "Normal", "PBR"};
    static const char* face_key[6] = {"_front", "_back", "_top", "_bottom", "_left", "_right"};
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line != "{") continue;
        BlockRec r;
        std::vector<std::string> fields;
        std::string f;
        while (std::getline(in, f)) {
            if (!f.empty() && f.back() == '\r') f.pop_back();
            if (f == "}") break;
            fields.push_back(f);
        }
        for (const std::string& field : fields) {
            if (has(field, "Name")) { r.name = field_value(field); continue; }
            bool done = false;
            for (int k = 0; k < 3 && !done; ++k) {
                for (int face = 0; face < 6 && !done; ++face)
                    if (has(field, (std::string(kind_key[k]) + face_key[face]).c_str())) { r.maps[k].f[face] = field_value(field); done = true; }
                if (!done && has(field, kind_key[k])) {  // "<Kind>_default" or plain "<Kind>": fill the empty faces
                    std::string v = field_value(field);
                    for (int face = 0; face < 6; ++face) if (r.maps[k].f[face].empty()) r.maps[k].f[face] = v;
                    done = true;
                }
            }
            if (done) continue;
            if (has(field, "Transparent")) r.transparent = true;
            else if (has(field, "sss") || has(field, "SSS") || has(field, "SUBSURFACE")) r.sss = true;
            else if (has(field, "Emissive")) r.emissive = field_value(field);
            else if (has(field, "SND_STEP") || has(field, "SND_MODIFY")) {}
            else if (has(field, "MC_ID") || has(field, "MCID") || has(field, "mc_id") || has(field, "mcid")) {
                size_t loc = field.find(':');
                std::stringstream ss(loc == std::string::npos ? "" : field.substr(loc + 1));
                std::string tok;
                while (std::getline(ss, tok, ',')) { try { r.mc_ids.push_back(std::stoi(tok)); } catch (...) {} }
            }
        }
        if (db->blocks.size() >= 127) break;  // GenerateBlockID throws at 128 (BlockDatabaseParser.cpp:31-42)
        r.id = (int)db->blocks.size() + 1;
        auto it = db->by_name.find(r.name);
        if (it != db->by_name.end()) {  // a repeated name overwrites the map entry but still burns an id
            db->blocks.push_back(r);
            it->second = r.id;
        } else {
            db->by_name[r.name] = r.id;
            db->blocks.push_back(r);
        }
    }
    for (const BlockRec& b : db->blocks) {
        if (db->by_name[b.name] != b.id) continue;
        for (int k = 0; k < 3; ++k) for (int face = 0; face < 6; ++face) db->layers[k].push_back(b.maps[k].f[face]);
        if (!b.emissive.empty()) db->layers[3].push_back(b.emissive);
    }
    for (int k = 0; k < 4; ++k) {
        std::sort(db->layers[k].begin(), db->layers[k].end());
        db->layers[k].erase(std::unique(db->layers[k].begin(), db->layers[k].end()), db->layers[k].end());
    }
    return db;
}
void vxh_blockdb_free(vxh_blockdb* db) { delete db; }
int32_t vxh_blockdb_block_count(const vxh_blockdb* db) { return (int32_t)db->by_name.size(); }
int32_t vxh_blockdb_block_id(const vxh_blockdb* db, const char* name) {
    auto it = db->by_name.find(name);
    return it == db->by_name.end() ? 0 : it->second;
}
static const BlockRec* rec_by_id(const vxh_blockdb* db, int id) {
    if (id < 1 || id > (int)db->blocks.size()) return nullptr;
    const BlockRec& b = db->blocks[id - 1];
    auto it = db->by_name.find(b.name);
    return (it != db->by_name.end() && it->second == id) ? &b : nullptr;
}
const char* vxh_blockdb_block_name(const vxh_blockdb* db, int32_t id) {
    const BlockRec* b = rec_by_id(db, id);
    return b ? b->name.c_str() : "???";
}
int32_t vxh_blockdb_layer_count(const vxh_blockdb* db, int32_t kind) { return (kind < 0 || kind > 3) ? 0 : (int32_t)db->layers[kind].size(); }
const char* vxh_blockdb_layer_path(const vxh_blockdb* db, int32_t kind, int32_t layer) {
    if (kind < 0 || kind > 3 || layer < 0 || layer >= (int)db->layers[kind].size()) return "";
    return db->layers[kind][layer].c_str();
}
int32_t vxh_blockdb_texture(const vxh_blockdb* db, int32_t kind, int32_t id, int32_t face) {
    const BlockRec* b = rec_by_id(db, id);
    if (!b) return kind == 2 ? 0 : -1;  // unknown id: PBR lookup returns 0 (BlockDatabase.cpp:398-401), others -1
    if (kind == 3) return b->emissive.empty() ? -1 : db->layer_of(3, b->emissive);
    if (kind < 0 || kind > 2 || face < 0 || face > 5) return -1;
    return db->layer_of(kind, b->maps[kind].f[face]);
}
void vxh_blockdb_table(const vxh_blockdb* db, int32_t* t) {
    for (int i = 0; i < 128; ++i) {
        t[0 * 128 + i] = vxh_blockdb_texture(db, 0, i, 0);
        t[1 * 128 + i] = vxh_blockdb_texture(db, 1, i, 0);
        t[2 * 128 + i] = vxh_blockdb_texture(db, 2, i, 0);
        t[3 * 128 + i] = vxh_blockdb_texture(db, 3, i, 0);
        const BlockRec* b = rec_by_id(db, i);
        t[4 * 128 + i] = (b && b->transparent) ? 1 : 0;
        t[5 * 128 + i] = (b && b->sss) ? 1 : 0;
    }
}
void vxh_blockdb_face_props(const vxh_blockdb* db, const char* name, int32_t* o) {
    int id = vxh_blockdb_block_id(db, name);
    o[0] = id;
    const int faces[3] = {2, 0, 3};  // top, front, bottom
    for (int g = 0; g < 3; ++g)
        for (int k = 0; k < 3; ++k) o[1 + g * 3 + k] = id ? vxh_blockdb_texture(db, k, id, faces[g]) : -1;
}
void vxh_blockdb_minecraft_lut(const vxh_blockdb* db, uint8_t* out256) {
    // unlisted ids -> GetBlockID("INVALID_BLOCK") (BlockDatabase.cpp:605-608); id 0 -> 0 (:601-603)
    memset(out256, (uint8_t)vxh_blockdb_block_id(db, "INVALID_BLOCK"), 256);
    for (const BlockRec& b : db->blocks) {
        if (!rec_by_id(db, b.id)) continue;
        for (int mc : b.mc_ids) out256[(uint8_t)mc] = (uint8_t)b.id;
    }
    out256[0] = 0;
}

void vxh_gen_texture_array(uint32_t seed, int32_t kind, int32_t layers, int32_t size, uint8_t* rgba) {
    for (int L = 0; L < layers; ++L) {
        uint32_t ls = seed * 977u + (uint32_t)L * 131u + (uint32_t)kind * 7u;
        float base[3] = {0.25f + 0.7f * (float)(hash3(ls, 1, 1) & 255) / 255.0f, 0.25f + 0.7f * (float)(hash3(ls, 2, 1) & 255) / 255.0f,
                         0.25f + 0.7f * (float)(hash3(ls, 3, 1) & 255) / 255.0f};
        bool metal = (hash3(ls, 4, 1) % 5) == 0;
        float rough0 = 0.1f + 0.85f * (float)(hash3(ls, 5, 1) & 255) / 255.0f;
        for (int y = 0; y < size; ++y)
            for (int x = 0; x < size; ++x) {
                float u = (float)x / (float)size, v = (float)y / (float)size;
                float n = fbm(u * 8.0f, v * 8.0f, 4, ls), m = value_noise(u * 32.0f, v * 32.0f, ls + 9u);
                uint8_t* p = rgba + (((size_t)L * size + y) * size + x) * 4;
                auto q = [](float f) { f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f); return (uint8_t)(f * 255.0f + 0.5f); };
                if (kind == 0) {
                    float s = 0.75f + 0.25f * n + 0.08f * m;
                    bool brick = ((x / 64 + y / 32) & 1) != 0;
                    p[0] = q(base[0] * s * (brick ? 0.9f : 1.0f)); p[1] = q(base[1] * s); p[2] = q(base[2] * s * (brick ? 1.0f : 0.92f));
                    p[3] = (L % 7 == 6 && m > 0.3f) ? 0 : 255;  // a few cut-out (leaf-like) layers
                } else if (kind == 1) {
                    float dx = value_noise(u * 16.0f + 0.37f, v * 16.0f, ls) * 0.35f, dy = value_noise(u * 16.0f, v * 16.0f + 0.61f, ls + 3u) * 0.35f;
                    float nz = sqrtf(fmaxf(0.05f, 1.0f - dx * dx - dy * dy));
                    p[0] = q(0.5f + 0.5f * dx); p[1] = q(0.5f + 0.5f * dy); p[2] = q(0.5f + 0.5f * nz); p[3] = 255;
                } else if (kind == 2) {
                    p[0] = q(rough0 + 0.15f * n); p[1] = metal ? q(0.85f + 0.15f * m) : q(0.02f * (m + 1.0f));
                    p[2] = q(0.5f + 0.5f * n); p[3] = q(0.8f + 0.2f * m);
                } else {
                    float e = (fabsf(u - 0.5f) < 0.3f && fabsf(v - 0.5f) < 0.3f) ? 0.7f + 0.3f * n : 0.05f;
                    p[0] = q(e); p[1] = q(e * 0.8f); p[2] = q(e * 0.5f); p[3] = 255;
                }
            }
    }
}

}  // extern "C"
