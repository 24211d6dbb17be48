// vxrt_oracle_world.cpp — CPU ORACLE of the world producers.  TEST INFRASTRUCTURE ONLY (see vxrt_oracle.h).
//
// Restates, function by function:
//   * VoxelRT::GenerateWorld with gen_structures = false (Core/WorldGenerator.cpp:208-313), GetBiome (:33-47),
//     SetVerticalBlocks (:49-88), over FastNoise (Dependencies/fast_noise/FastNoise.cpp — vendored third-party code, the
//     reference's copy is the pinned version): SetSeed :197-215, CalculateFractalBounding :217-227, Index2D_12 :244-247,
//     GradCoord2D :312-317 with GRAD_X / GRAD_Y :37-48, GetNoise(x, y) :410-470, SingleSimplexFractalFBM :1191-1208,
//     SingleSimplex(offset, x, y) :1268-1335, FastFloor :184.
//   * MCWorldImporter::ImportRegionFile's voxel loop and WriteVoxel (Core/NBT/Importer.cpp:67-83, 112-141) with
//     enkiGetChunkSectionVoxelData's byte / nibble addressing (Dependencies/enkiMI/enkimi.c:2300, 2391-2403) and
//     BlockDatabase::GetIDFromMCID (Core/BlockDatabase.cpp:599-612, as a 256-entry table).
//   * the LightLocations scan of LoadWorld (Core/WorldFileHandler.cpp:53-69).
// Pinned against the reference's own sources compiled in oracle/_ref/libvxrt_ref_world.so (oracle/build_ref_world.py):
// tests/test_oracle_world.py; golden outputs of that build: tests/golden/world_ref.npz.
// Compiled with -ffp-contract=off like the rest of the oracle.
#include "vxrt_oracle.h"

#include <string.h>

#include <random>

namespace {

struct Noise {
    unsigned char perm[512], perm12[512];
    float frequency = 0.01f, lacunarity = 2.0f, gain = 0.5f, bounding = 1.0f;
    int octaves = 3;

    explicit Noise(int seed) {
        std::mt19937_64 gen(seed);
        for (int i = 0; i < 256; i++) perm[i] = (unsigned char)i;
        for (int j = 0; j < 256; j++) {
            int rng = (int)(gen() % (256 - j));
            int k = rng + j;
            int l = perm[j];
            perm[j] = perm[j + 256] = perm[k];
            perm[k] = (unsigned char)l;
            perm12[j] = perm12[j + 256] = perm[j] % 12;
        }
        bound();
    }
    void bound() {
        float amp = gain, ampFractal = 1.0f;
        for (int i = 1; i < octaves; i++) { ampFractal += amp; amp *= gain; }
        bounding = 1.0f / ampFractal;
    }
    static int fast_floor(float f) { return f >= 0 ? (int)f : (int)f - 1; }
    float grad(unsigned char offset, int x, int y, float xd, float yd) const {
        static const float GX[12] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
        static const float GY[12] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
        unsigned char lut = perm12[(x & 0xff) + perm[(y & 0xff) + offset]];
        return xd * GX[lut] + yd * GY[lut];
    }
    float simplex(unsigned char offset, float x, float y) const {
        const float SQRT3 = 1.7320508075688772935274463415059f;
        const float F2 = 0.5f * (SQRT3 - 1.0f);
        const float G2 = (3.0f - SQRT3) / 6.0f;
        float t = (x + y) * F2;
        int i = fast_floor(x + t), j = fast_floor(y + t);
        t = (i + j) * G2;
        float X0 = i - t, Y0 = j - t;
        float x0 = x - X0, y0 = y - Y0;
        int i1, j1;
        if (x0 > y0) { i1 = 1; j1 = 0; } else { i1 = 0; j1 = 1; }
        float x1 = x0 - (float)i1 + G2, y1 = y0 - (float)j1 + G2;
        float x2 = x0 - 1 + 2 * G2, y2 = y0 - 1 + 2 * G2;
        float n0, n1, n2;
        t = 0.5f - x0 * x0 - y0 * y0;
        if (t < 0) n0 = 0; else { t *= t; n0 = t * t * grad(offset, i, j, x0, y0); }
        t = 0.5f - x1 * x1 - y1 * y1;
        if (t < 0) n1 = 0; else { t *= t; n1 = t * t * grad(offset, i + i1, j + j1, x1, y1); }
        t = 0.5f - x2 * x2 - y2 * y2;
        if (t < 0) n2 = 0; else { t *= t; n2 = t * t * grad(offset, i + 1, j + 1, x2, y2); }
        return 70 * (n0 + n1 + n2);
    }
    float fbm(float x, float y) const {
        float sum = simplex(perm[0], x, y), amp = 1;
        int i = 0;
        while (++i < octaves) {
            x *= lacunarity; y *= lacunarity;
            amp *= gain;
            sum += simplex(perm[i], x, y) * amp;
        }
        return sum * bounding;
    }
    float get(float x, float y, bool fractal) const {
        x *= frequency; y *= frequency;
        return fractal ? fbm(x, y) : simplex(0, x, y);
    }
};

}  // namespace

extern "C" {

void vxo_fastnoise_2d(int32_t seed, int32_t fractal, float frequency, int32_t octaves, const float* xy, int32_t n, float* out) {
    Noise g(seed);
    g.frequency = frequency;
    if (fractal) { g.octaves = octaves; g.bound(); }
    for (int32_t i = 0; i < n; ++i) out[i] = g.get(xy[2 * i], xy[2 * i + 1], fractal != 0);
}

void vxo_generate_world(uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const vxrt_worldgen_params* p) {
    memset(blocks, 0, (size_t)nx * ny * nz);
    Noise height(p->noise_seed), biome(p->biome_seed);
    height.frequency = (float)0.00385;  // :238
    height.octaves = 6;                 // :239
    height.bound();
    for (int x = 0; x < nx; x++)
        for (int z = 0; z < nz; z++) {
            int Yc = 50, b = 1;  // flat world (:309)
            if (p->gen_type) {
                float real_x = (float)x, real_z = (float)z;
                float h = height.get(real_x, real_z, true);
                float hgt = ((h + 1.0f) / 2.0f) * 40.0f;
                float column_noise = biome.get(real_x / 2.0f, real_z / 2.0f, false);
                column_noise = ((column_noise + 1.0f) / 2) * 240;
                b = column_noise < 90 ? 0 : 1;
                Yc = (int)(hgt + 8);
            }
            for (int y = 0; y < Yc && y < ny; y++) {  // SetVerticalBlocks, swapstone = false
                int id;
                if (b == 1) id = y >= Yc - 1 ? p->grass_id : (y >= Yc - 5 ? p->dirt_id : p->stone_id);
                else id = y >= Yc - 8 ? p->sand_id : p->stone_id;
                blocks[(size_t)x + (size_t)y * nx + (size_t)z * nx * ny] = (uint8_t)id;
            }
        }
}

void vxo_import_sections(uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const uint8_t* ids, const uint8_t* nibbles,
                         const uint8_t* has_data, const int32_t* origins, int32_t n, const int32_t* import_origin, const uint8_t* lut,
                         int32_t clear_first) {
    if (clear_first) memset(blocks, 0, (size_t)nx * ny * nz);
    for (int32_t s = 0; s < n; ++s)
        for (int sy = 0; sy < 16; ++sy)
            for (int sz = 0; sz < 16; ++sz)
                for (int sx = 0; sx < 16; ++sx) {
                    const uint32_t pos = (uint32_t)(sy * 256 + sz * 16 + sx);
                    uint8_t voxel = ids[(size_t)s * 4096 + pos];
                    uint8_t dataval = 0;
                    if (has_data[s]) dataval = 0xF & (nibbles[(size_t)s * 2048 + pos / 2] >> (4 * (pos & 1)));
                    if (dataval != 0) continue;
                    voxel = lut[voxel];
                    const int X = origins[3 * s] + sx - import_origin[0] + nx / 2;
                    const int Y = origins[3 * s + 1] + sy - import_origin[1];
                    const int Z = origins[3 * s + 2] + sz - import_origin[2] + nz / 2;
                    if (Y >= ny || X >= nx || Z >= nz || Y < 0 || X < 0 || Z < 0 || voxel == 0) continue;
                    blocks[(size_t)X + (size_t)Y * nx + (size_t)Z * nx * ny] = voxel;
                }
}

int32_t vxo_collect_lights(const uint8_t* blocks, int32_t nx, int32_t ny, int32_t nz, const int32_t* table6x128, int32_t* xyz_out,
                           int32_t capacity) {
    int32_t found = 0;
    const size_t n = (size_t)nx * ny * nz;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t b = blocks[i];
        if (b < 128 && table6x128[3 * 128 + b] >= 0) {
            if (found < capacity) {
                size_t idx = i;
                const int z = (int)(idx / ((size_t)nx * ny));
                idx -= (size_t)z * nx * ny;
                xyz_out[3 * found] = (int)(idx % nx);
                xyz_out[3 * found + 1] = (int)(idx / nx);
                xyz_out[3 * found + 2] = z;
            }
            ++found;
        }
    }
    return found;
}

}  // extern "C"
