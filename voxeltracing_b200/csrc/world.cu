// world.cu — world producers: the block-id grid is produced in device memory instead of on the host.
//   worldgen_kernel         VoxelRT::GenerateWorld without structures (Core/WorldGenerator.cpp:208-313): FastNoise 2-D simplex
//                           (Dependencies/fast_noise/FastNoise.cpp:1191-1208, 1268-1335) per column, SetVerticalBlocks (:49-88)
//   import_sections_kernel  MCWorldImporter::ImportRegionFile / WriteVoxel (Core/NBT/Importer.cpp:67-146) over chunk sections
//                           the host side has already inflated (voxeltracing_b200/host/vxrt_mca.cpp)
//   light_count / light_write kernels   the LightLocations scan of LoadWorld (Core/WorldFileHandler.cpp:53-69), in the
//                           reference's order (ascending linear index)
// All three are streaming kernels over the 18.9 MB grid (HBM-bound; the noise is ~400 flops per column of 128 voxels).
// The file is compiled with --fmad=false like the ray path, so the noise rounds like FastNoise built without contraction.
#include "ctx.h"

#include <random>

namespace {

// ---- FastNoise 2-D simplex ------------------------------------------------------------------------------------------
struct NoiseTables {
    uint8_t perm[512];
    uint8_t perm12[512];
};

// FastNoise::SetSeed (FastNoise.cpp:197-215)
void build_tables(int seed, NoiseTables* t) {
    std::mt19937_64 gen(seed);
    for (int i = 0; i < 256; ++i) t->perm[i] = (uint8_t)i;
    for (int j = 0; j < 256; ++j) {
        const int k = (int)(gen() % (uint64_t)(256 - j)) + j;
        const uint8_t l = t->perm[j];
        t->perm[j] = t->perm[j + 256] = t->perm[k];
        t->perm[k] = l;
        t->perm12[j] = t->perm12[j + 256] = t->perm[j] % 12;
    }
}

// FastFloor (FastNoise.cpp:184): note the quirk for negative integers (-2.0 -> -3)
__device__ __forceinline__ int fast_floor(float f) { return f >= 0.0f ? __float2int_rz(f) : __float2int_rz(f) - 1; }

// GradCoord2D (:312-317) with GRAD_X / GRAD_Y (:37-48)
__device__ __forceinline__ float grad2(const NoiseTables& t, unsigned offset, int x, int y, float xd, float yd) {
    const unsigned lut = t.perm12[(x & 0xff) + t.perm[(y & 0xff) + offset]];
    const float gx = lut < 8u ? ((lut & 1u) ? -1.0f : 1.0f) : 0.0f;
    const float gy = lut < 4u ? ((lut & 2u) ? -1.0f : 1.0f) : (lut < 8u ? 0.0f : ((lut & 1u) ? -1.0f : 1.0f));
    return xd * gx + yd * gy;
}

// SingleSimplex(offset, x, y) (:1275-1335)
__device__ float simplex2(const NoiseTables& tb, unsigned offset, float x, float y) {
    constexpr float SQRT3 = 1.7320508075688772935274463415059f;
    constexpr float F2 = 0.5f * (SQRT3 - 1.0f);
    constexpr float G2 = (3.0f - SQRT3) / 6.0f;
    float t = (x + y) * F2;
    const int i = fast_floor(x + t), j = fast_floor(y + t);
    t = (float)(i + j) * G2;
    const float X0 = (float)i - t, Y0 = (float)j - t;
    const float x0 = x - X0, y0 = y - Y0;
    const int i1 = x0 > y0 ? 1 : 0, j1 = 1 - i1;
    const float x1 = x0 - (float)i1 + G2, y1 = y0 - (float)j1 + G2;
    const float x2 = x0 - 1.0f + 2.0f * G2, y2 = y0 - 1.0f + 2.0f * G2;
    float n0 = 0.0f, n1 = 0.0f, n2 = 0.0f;
    t = 0.5f - x0 * x0 - y0 * y0;
    if (!(t < 0.0f)) { t *= t; n0 = t * t * grad2(tb, offset, i, j, x0, y0); }
    t = 0.5f - x1 * x1 - y1 * y1;
    if (!(t < 0.0f)) { t *= t; n1 = t * t * grad2(tb, offset, i + i1, j + j1, x1, y1); }
    t = 0.5f - x2 * x2 - y2 * y2;
    if (!(t < 0.0f)) { t *= t; n2 = t * t * grad2(tb, offset, i + 1, j + 1, x2, y2); }
    return 70.0f * (n0 + n1 + n2);
}

struct WorldGenArgs {
    uint8_t* blocks;
    int nx, ny, nz;
    int gen_type;
    float frequency, biome_frequency, lacunarity, gain, bounding;
    int octaves;
    unsigned grass, dirt, stone, sand;
    NoiseTables height_tab, biome_tab;
};

constexpr int WG_THREADS = 256;
constexpr int WG_MAX_NX = 1024;
constexpr int WG_MAX_OCTAVES = 8;

// One CTA per z-plane (a contiguous nx * ny byte block).
// Phase 1: one task per (column, noise evaluation) — the octaves of the height noise and the biome value are independent
//   SingleSimplex calls, so the plane's nx * (octaves + 1) evaluations are spread over the CTA instead of one thread walking
//   a column's 7 evaluations in series; the results wait in shared memory.
// Phase 2: per column, the octaves are summed in FastNoise's order (same roundings) -> surface level + biome; per 16-column
//   quad the highest level and the lowest stone top.
// Phase 3: the plane is written in memory order with 16-byte stores; a quad above all of its columns is zeros, one below
//   all of their stone tops is a stone splat, only the rows in between look at the columns (heights span 8..48 of 128 rows).
__global__ void __launch_bounds__(WG_THREADS) worldgen_kernel(const __grid_constant__ WorldGenArgs a) {
    __shared__ NoiseTables tabs[2];
    __shared__ float octave[(WG_MAX_OCTAVES + 1) * WG_MAX_NX];   // [evaluation][column]; the last row is the biome value
    __shared__ short level[WG_MAX_NX];
    __shared__ unsigned char biome_of[WG_MAX_NX];
    __shared__ short quad_top[WG_MAX_NX / 16], quad_stone[WG_MAX_NX / 16];
    {
        const unsigned* src0 = reinterpret_cast<const unsigned*>(&a.height_tab);
        const unsigned* src1 = reinterpret_cast<const unsigned*>(&a.biome_tab);
        unsigned* dst = reinterpret_cast<unsigned*>(tabs);
        for (int i = threadIdx.x; i < 256; i += WG_THREADS) { dst[i] = src0[i]; dst[256 + i] = src1[i]; }
    }
    __syncthreads();
    const int z = blockIdx.x;
    if (a.gen_type) {
        const int evals = a.octaves + 1;
        for (int task = threadIdx.x; task < evals * a.nx; task += WG_THREADS) {
            const int e = task / a.nx, x = task - e * a.nx;
            const float real_x = (float)x, real_z = (float)z;
            float v;
            if (e < a.octaves) {   // octave e of SingleSimplexFractalFBM (:1191-1208): coordinates scaled e times by the lacunarity
                float fx = real_x * a.frequency, fz = real_z * a.frequency;
                for (int i = 0; i < e; ++i) { fx *= a.lacunarity; fz *= a.lacunarity; }
                v = simplex2(tabs[0], tabs[0].perm[e], fx, fz);
            } else {               // BiomeGenerator.GetNoise(real_x / 2, real_z / 2) (:249)
                v = simplex2(tabs[1], 0u, (real_x / 2.0f) * a.biome_frequency, (real_z / 2.0f) * a.biome_frequency);
            }
            octave[e * WG_MAX_NX + x] = v;
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < a.nx; x += WG_THREADS) {
        int Yc = 50, biome = 1;  // flat world: SetVerticalBlocks(world, x, z, 50, 1, false) (:309)
        if (a.gen_type) {
            float sum = octave[x], amp = 1.0f;
            for (int i = 1; i < a.octaves; ++i) {
                amp *= a.gain;
                sum += octave[i * WG_MAX_NX + x] * amp;
            }
            const float h = sum * a.bounding;
            const float height = ((h + 1.0f) / 2.0f) * 40.0f;                       // :246-247
            const float column_noise = ((octave[a.octaves * WG_MAX_NX + x] + 1.0f) / 2.0f) * 240.0f;   // :250
            biome = column_noise < 90.0f ? 0 : 1;                                     // GetBiome (:33-47)
            Yc = __float2int_rz(height + 8.0f);                                       // :257
        }
        level[x] = (short)max(-32000, min(Yc, 32000));
        biome_of[x] = (unsigned char)biome;
    }
    __syncthreads();
    const int qpr = a.nx >> 4;
    for (int q = threadIdx.x; q < qpr; q += WG_THREADS) {
        int top = -32768, stone = 32767;
        for (int k = 0; k < 16; ++k) {
            const int L = level[q * 16 + k];
            top = max(top, L);
            stone = min(stone, L - (biome_of[q * 16 + k] ? 5 : 8));   // rows below level - 5 (biome 1) / level - 8 (biome 0) are stone
        }
        quad_top[q] = (short)top; quad_stone[q] = (short)max(stone, -32768);
    }
    __syncthreads();
    uint4* plane = reinterpret_cast<uint4*>(a.blocks + (size_t)z * a.nx * a.ny);
    const unsigned stone4 = a.stone * 0x01010101u;
    const int items = qpr * a.ny;                               // <= 4096 (nx * ny <= 65536)
    const unsigned rdiv = ((1u << 20) + qpr - 1) / qpr;         // it / qpr == (it * rdiv) >> 20 for it < 4096, qpr <= 64
    for (int it = threadIdx.x; it < items; it += WG_THREADS) {
        const int y = (int)(((unsigned)it * rdiv) >> 20), q = it - y * qpr;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (y < quad_stone[q]) v = make_uint4(stone4, stone4, stone4, stone4);
        else if (y < quad_top[q]) {
            unsigned w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                w[j] = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = q * 16 + j * 4 + k, L = level[x];
                    const bool b1 = biome_of[x] != 0;
                    unsigned id = 0;   // SetVerticalBlocks (:49-88), swapstone = false
                    if (y < L) id = y >= L - 1 ? (b1 ? a.grass : a.sand) : (y >= L - (b1 ? 5 : 8) ? (b1 ? a.dirt : a.sand) : a.stone);
                    w[j] |= id << (8 * k);
                }
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        plane[it] = v;
    }
}

// ---- Minecraft section scatter --------------------------------------------------------------------------------------
struct ImportArgs {
    uint8_t* blocks;
    int nx, ny, nz;
    const uint8_t* ids;       // n * 4096, YZX
    const uint8_t* nibbles;   // n * 2048
    const uint8_t* has_data;  // n
    const int32_t* origins;   // 3 * n
    int n;
    int ox, oy, oz;           // import origin
    uint8_t lut[256];
};

// One CTA per section, a thread per group of 4 voxels along x (one 32-bit word of ids, one 16-bit word of nibbles).
// Sections never overlap in space, and within the batch the host keeps only the last section per (chunk, Y) like the
// reference's chunk.sections[] table does, so the scatter is order independent.
__global__ void __launch_bounds__(256) import_sections_kernel(const __grid_constant__ ImportArgs a) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = a.lut[threadIdx.x];
    __syncthreads();
    const int s = blockIdx.x;
    const unsigned* ids = reinterpret_cast<const unsigned*>(a.ids + (size_t)s * 4096);
    const unsigned short* nib = reinterpret_cast<const unsigned short*>(a.nibbles + (size_t)s * 2048);
    const bool has_data = a.has_data[s] != 0;
    // WriteVoxel (:67-83): Position -= ivec3(ImportOrigin); x += HALF_WORLD_X; z += HALF_WORLD_Z
    const int bx = a.origins[3 * s] - a.ox + (a.nx >> 1), by = a.origins[3 * s + 1] - a.oy, bz = a.origins[3 * s + 2] - a.oz + (a.nz >> 1);
    for (int g = threadIdx.x; g < 1024; g += 256) {
        const int sy = g >> 6, sz = (g >> 2) & 15, sx = (g & 3) << 2;
        const int y = by + sy, z = bz + sz;
        if (y < 0 || y >= a.ny || z < 0 || z >= a.nz) continue;
        const unsigned w = ids[g];
        const unsigned d = has_data ? nib[g] : 0u;
        uint8_t* row = a.blocks + (size_t)y * a.nx + (size_t)z * a.nx * a.ny;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = bx + sx + k;
            const unsigned voxel = lut[(w >> (8 * k)) & 0xffu];
            if (((d >> (4 * k)) & 0xfu) == 0u && voxel != 0u && x >= 0 && x < a.nx) row[x] = (uint8_t)voxel;
        }
    }
}

// ---- light locations ------------------------------------------------------------------------------------------------
constexpr int LC_THREADS = 256;
constexpr int LC_QUADS = 4;                               // 16-byte loads per thread
constexpr int LC_CHUNK = LC_THREADS * LC_QUADS * 16;      // voxels per CTA

__device__ __forceinline__ unsigned emissive_mask16(const uint4& v, const unsigned char* em) {
    unsigned m = 0;
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) m |= (unsigned)em[(w[k >> 2] >> (8 * (k & 3))) & 0xffu] << k;
    return m;
}

// phase 0 counts the emissive voxels of each CTA's chunk, phase 1 (after an exclusive scan of the counts) writes their
// coordinates in ascending linear order: a thread owns 64 consecutive voxels, a block-wide scan orders the threads.
template <int PHASE>
__global__ void __launch_bounds__(LC_THREADS) lights_kernel(const uint8_t* __restrict__ blocks, size_t nvox, int nx, int ny,
                                                            const int32_t* __restrict__ block_data, unsigned* __restrict__ counts,
                                                            int32_t* __restrict__ out, int capacity) {
    __shared__ unsigned char em[256];
    __shared__ unsigned warp_sum[LC_THREADS / 32];
    // BlockEmissiveData[id] >= 0 (row 3 of the table); ids >= 128 have no entry: GetBlockEmissiveTexture returns -1 for them
    em[threadIdx.x] = threadIdx.x < 128 ? (block_data[3 * 128 + threadIdx.x] >= 0 ? 1 : 0) : 0;
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * LC_CHUNK + (size_t)threadIdx.x * (LC_QUADS * 16);
    unsigned mask[LC_QUADS];
    unsigned mine = 0;
#pragma unroll
    for (int q = 0; q < LC_QUADS; ++q) {
        const size_t at = first + (size_t)q * 16;
        mask[q] = 0;
        if (at < nvox) mask[q] = emissive_mask16(__ldg(reinterpret_cast<const uint4*>(blocks + at)), em);  // nvox % 16 == 0
        mine += __popc(mask[q]);
    }
    // block-wide exclusive scan of `mine`
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < LC_THREADS / 32; ++w) {
        const unsigned s = warp_sum[w];
        if (w < warp) before += s;
        total += s;
    }
    if (PHASE == 0) {
        if (threadIdx.x == 0) counts[blockIdx.x] = total;
        return;
    }
    unsigned at_out = counts[blockIdx.x] + before + incl - mine;  // counts[] holds the exclusive scan by now
#pragma unroll
    for (int q = 0; q < LC_QUADS; ++q) {
        unsigned m = mask[q];
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            if (at_out < (unsigned)capacity) {
                size_t idx = first + (size_t)q * 16 + k;
                const int z = (int)(idx / ((size_t)nx * ny));
                idx -= (size_t)z * nx * ny;
                const int y = (int)(idx / nx), x = (int)(idx - (size_t)y * nx);
                out[3 * at_out] = x; out[3 * at_out + 1] = y; out[3 * at_out + 2] = z;
            }
            ++at_out;
        }
    }
}

// exclusive scan of the per-CTA counts (a few thousand entries) by one CTA; counts[n] receives the total
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned* __restrict__ counts, int n) {
    __shared__ unsigned part[1024];
    const int per = (n + 1023) / 1024, a = min(threadIdx.x * per, n), b = min(a + per, n);
    unsigned s = 0;
    for (int i = a; i < b; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - s;
    for (int i = a; i < b; ++i) { const unsigned c = counts[i]; counts[i] = run; run += c; }
    if (threadIdx.x == 1023) counts[n] = part[1023];
}

}  // namespace

int vxrt_launch_generate_world(vxrt_ctx* c, const vxrt_worldgen_params& p) {
    WorldGenArgs a;
    a.blocks = c->d_blocks; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz;
    a.gen_type = p.gen_type ? 1 : 0;
    a.frequency = (float)0.00385;   // NoiseGenerator.SetFrequency(0.00385) (WorldGenerator.cpp:238)
    a.biome_frequency = 0.01f;      // FastNoise default m_frequency (FastNoise.h:221)
    a.lacunarity = 2.0f; a.gain = 0.5f; a.octaves = 6;
    {   // CalculateFractalBounding (FastNoise.cpp:217-227)
        float amp = a.gain, amp_fractal = 1.0f;
        for (int i = 1; i < a.octaves; ++i) { amp_fractal += amp; amp *= a.gain; }
        a.bounding = 1.0f / amp_fractal;
    }
    a.grass = (unsigned)p.grass_id & 0xffu; a.dirt = (unsigned)p.dirt_id & 0xffu;
    a.stone = (unsigned)p.stone_id & 0xffu; a.sand = (unsigned)p.sand_id & 0xffu;
    build_tables(p.noise_seed, &a.height_tab);
    build_tables(p.biome_seed, &a.biome_tab);
    worldgen_kernel<<<c->nz, WG_THREADS, 0, c->stream>>>(a);   // nx <= 1024, nx % 16 == 0 (vxrt_cuda_create)
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_import_sections(vxrt_ctx* c, const uint8_t* d_ids, const uint8_t* d_nibbles, const uint8_t* d_has_data,
                                const int32_t* d_origins, int n, const int32_t origin[3], const uint8_t lut[256]) {
    ImportArgs a;
    a.blocks = c->d_blocks; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz;
    a.ids = d_ids; a.nibbles = d_nibbles; a.has_data = d_has_data; a.origins = d_origins; a.n = n;
    a.ox = origin[0]; a.oy = origin[1]; a.oz = origin[2];
    for (int i = 0; i < 256; ++i) a.lut[i] = lut[i];
    import_sections_kernel<<<n, 256, 0, c->stream>>>(a);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_lights_chunks(const vxrt_ctx* c) { return (int)((c->nvox + LC_CHUNK - 1) / LC_CHUNK); }

// d_counts: vxrt_lights_chunks() + 1 unsigned; d_out: 3 * capacity ints (may be null when capacity == 0)
int vxrt_launch_collect_lights(vxrt_ctx* c, unsigned* d_counts, int32_t* d_out, int capacity) {
    const int chunks = vxrt_lights_chunks(c);
    lights_kernel<0><<<chunks, LC_THREADS, 0, c->stream>>>(c->d_blocks, c->nvox, c->nx, c->ny, c->d_block_data, d_counts, nullptr, 0);
    scan_counts_kernel<<<1, 1024, 0, c->stream>>>(d_counts, chunks);
    c->launches += 2;
    if (capacity > 0) {
        lights_kernel<1><<<chunks, LC_THREADS, 0, c->stream>>>(c->d_blocks, c->nvox, c->nx, c->ny, c->d_block_data, d_counts, d_out, capacity);
        c->launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    return VXRT_OK;
}
