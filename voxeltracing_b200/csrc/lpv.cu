// lpv.cu — light propagation volume flood fill: Core/VolumetricFloodFill.cpp on the device-resident grid.
//
// The reference floods light from the lamps with a FIFO queue on the host (PropogateVolume :247-326) and mirrors every voxel it
// touches into two 3-D textures with a 1-byte glTexSubImage3D each.  Two byte volumes result: the light level (seed min(limit, 8),
// minus one per step) and the block type of the lamp that lit the voxel.  A neighbour is only taken when its level + 2 < the node's
// level, so the *block type* volume depends on the order of the queue.  Reproduced here exactly:
//
//  * full repropagation (start-up Pipeline.cpp:1602-1611, World::RepropogateLPV_ World.cpp:554-572): every seed has the same level,
//    so the FIFO order is level-synchronous.  One wave per level; node i of the wave (its FIFO rank) claims each neighbour it would
//    take with atomicMin(claim[n], 6 i + direction): the smallest key is the node the sequential queue would have popped first.
//    The winners, written in key order by an ordered compaction (per-CTA contiguous chunks, counts, one-CTA scan), are the next
//    wave in FIFO order.  4 launches per level, no host synchronisation (sizes stay on the device).
//  * block edits (World.cpp:273-333, :395-446, :482-485: DepropogateVolume + PropogateVolume from mixed-level queues, whose result
//    depends on the pop order node by node): the exact FIFO, run by one warp.  A lane per neighbour direction (the six neighbours of
//    a node are distinct voxels, so handling them together is the sequential order), pushes in direction order by ballot, queue
//    entries fetched 32 at a time.  An edit touches at most a few thousand voxels around it.
#include "ctx.h"
#include "texture.cuh"
#include "lpv_sample.cuh"

#include <stdlib.h>

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int LPV_THREADS = 256;
constexpr unsigned NO_CLAIM = 0xffffffffu;

struct LpvGrid {
    const uint8_t* __restrict__ blocks;
    uint8_t* level;
    uint8_t* color;
    int nx, ny, nz;
};

// InVoxelVolume (VolumetricFloodFill.cpp:22-30): the planes x = 0, y = 0, z = 0 are outside
__device__ __forceinline__ bool lpv_inside(const LpvGrid& g, int x, int y, int z) {
    return x > 0 && y > 0 && z > 0 && x < g.nx && y < g.ny && z < g.nz;
}
__device__ __forceinline__ int lpv_index(const LpvGrid& g, int x, int y, int z) { return x + g.nx * (y + g.ny * z); }

// PropogateVolume visits +x -x +y -y -z +z (:262-324), DepropogateVolume +x -x +y -y +z -z (:344-465)
__device__ __forceinline__ void lpv_dir(int k, bool propagate, int& dx, int& dy, int& dz) {
    dx = k == 0 ? 1 : (k == 1 ? -1 : 0);
    dy = k == 2 ? 1 : (k == 3 ? -1 : 0);
    const int zs = propagate ? -1 : 1;
    dz = k == 4 ? zs : (k == 5 ? -zs : 0);
}

// ---- full repropagation ----------------------------------------------------------------------------------------------------

// AddLightToVolume (:190-205) for every light location: level min(limit, 8), block type = the block at the location
__global__ void __launch_bounds__(LPV_THREADS) lpv_seed_kernel(LpvGrid g, const int32_t* __restrict__ xyz, const unsigned* __restrict__ n_ptr,
                                                               int capacity, int seed_level, int* __restrict__ front, unsigned* __restrict__ n_out) {
    const unsigned n = min(*n_ptr, (unsigned)capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        int idx = -1;   // a location outside the volume is queued by the reference but reads back level 0: it never spreads
        if (lpv_inside(g, x, y, z)) {
            idx = lpv_index(g, x, y, z);
            g.level[idx] = (uint8_t)seed_level;
            g.color[idx] = g.blocks[idx];
        }
        front[i] = idx;
    }
}

__device__ __forceinline__ void lpv_chunk(unsigned n, unsigned& begin, unsigned& end) {
    unsigned chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + LPV_THREADS - 1) / LPV_THREADS * LPV_THREADS;
    begin = min(n, blockIdx.x * chunk);
    end = min(n, begin + chunk);
}

__global__ void __launch_bounds__(LPV_THREADS) lpv_claim_kernel(LpvGrid g, unsigned* __restrict__ claim, const int* __restrict__ front,
                                                                const unsigned* __restrict__ n_ptr) {
    const unsigned n = *n_ptr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int idx = front[i];
        if (idx < 0) continue;
        const int cur = g.level[idx];
        if (cur < 3) continue;   // level + 2 < cur has no solution
        const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            int dx, dy, dz;
            lpv_dir(k, true, dx, dy, dz);
            if (!lpv_inside(g, x + dx, y + dy, z + dz)) continue;
            const int q = lpv_index(g, x + dx, y + dy, z + dz);
            if (g.blocks[q] == 0 && (int)g.level[q] + 2 < cur) atomicMin(claim + q, i * 6u + (unsigned)k);
        }
    }
}

// which of its six claims node i won (a key is unique to (i, direction), so equality alone decides); per-CTA totals for the scan
__global__ void __launch_bounds__(LPV_THREADS) lpv_count_kernel(LpvGrid g, const unsigned* __restrict__ claim, const int* __restrict__ front,
                                                                const unsigned* __restrict__ n_ptr, uint8_t* __restrict__ wins,
                                                                unsigned* __restrict__ counts) {
    __shared__ unsigned warp_sum[LPV_THREADS / 32];
    unsigned begin, end;
    lpv_chunk(*n_ptr, begin, end);
    unsigned mine = 0;
    for (unsigned i = begin + threadIdx.x; i < end; i += LPV_THREADS) {
        const int idx = front[i];
        unsigned m = 0;
        if (idx >= 0 && g.level[idx] >= 3) {
            const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                int dx, dy, dz;
                lpv_dir(k, true, dx, dy, dz);
                if (lpv_inside(g, x + dx, y + dy, z + dz) && claim[lpv_index(g, x + dx, y + dy, z + dz)] == i * 6u + (unsigned)k) m |= 1u << k;
            }
        }
        wins[i] = (uint8_t)m;
        mine += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < LPV_THREADS / 32; ++w) t += warp_sum[w];
        counts[blockIdx.x] = t;
    }
}

// exclusive scan of the per-CTA totals (gridDim of the wave kernels <= 1024); the grand total is the size of the next wave
__global__ void __launch_bounds__(1024) lpv_scan_kernel(unsigned* __restrict__ counts, int n, unsigned* __restrict__ n_next) {
    __shared__ unsigned part[1024];
    const unsigned v = (int)threadIdx.x < n ? counts[threadIdx.x] : 0u;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned a = (int)threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += a;
        __syncthreads();
    }
    if ((int)threadIdx.x < n) counts[threadIdx.x] = part[threadIdx.x] - v;
    if (threadIdx.x == 1023) *n_next = part[1023];
}

// SetLightValue(neighbour, cur - 1, type) + LightBFS.push for the winners, in key order
__global__ void __launch_bounds__(LPV_THREADS) lpv_write_kernel(LpvGrid g, unsigned* __restrict__ claim, const int* __restrict__ front,
                                                                const unsigned* __restrict__ n_ptr, const uint8_t* __restrict__ wins,
                                                                const unsigned* __restrict__ counts, int* __restrict__ front_next) {
    __shared__ unsigned warp_sum[LPV_THREADS / 32];
    unsigned begin, end;
    lpv_chunk(*n_ptr, begin, end);
    unsigned running = counts[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned base = begin; base < end; base += LPV_THREADS) {
        const unsigned i = base + threadIdx.x;
        unsigned m = i < end ? wins[i] : 0u;
        const unsigned cnt = __popc(m);
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < LPV_THREADS / 32; ++w) {
            const unsigned s = warp_sum[w];
            if (w < warp) before += s;
            total += s;
        }
        __syncthreads();
        if (m) {
            unsigned at = running + before + incl - cnt;
            const int idx = front[i];
            const uint8_t next_level = (uint8_t)(g.level[idx] - 1), type = g.color[idx];
            const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                int dx, dy, dz;
                lpv_dir(k, true, dx, dy, dz);
                const int q = lpv_index(g, x + dx, y + dy, z + dz);
                g.color[q] = type;
                g.level[q] = next_level;
                claim[q] = NO_CLAIM;
                front_next[at++] = q;
            }
        }
        running += total;
    }
}

// ---- full repropagation as one persistent cooperative kernel --------------------------------------------------------------------
// The multi-kernel version above runs 13 (limit 4) to 29 (limit 8) small kernels back to back, i.e. at launch latency.  Here one grid of
// co-resident CTAs walks the same phases with grid-wide barriers: clear both volumes + count the lamps of the grid (LoadWorld's scan,
// WorldFileHandler.cpp:53-69) | seed the lamps in scan order | per level: claim | count wins | ordered write.  Frontier sizes and
// prefix sums are recomputed by every CTA from the per-CTA totals, so nothing but the totals crosses a barrier.
constexpr int LPV_SCAN_VOX = 64;                               // consecutive voxels a thread owns in the lamp scan
constexpr int LPV_SCAN_CHUNK = LPV_THREADS * LPV_SCAN_VOX;     // voxels per chunk (a chunk is scanned by one CTA)

struct LpvCoopArgs {
    LpvGrid g;
    unsigned* claim;
    int* front0;
    int* front1;
    uint8_t* wins;
    unsigned* counts;         // per CTA (wave phases)
    unsigned* chunk_counts;   // per chunk (lamp scan)
    const int32_t* block_data;
    const int32_t* lights;    // device list of a caller-provided queue order, or nullptr: scan the grid
    int n_lights;
    int seed_level;
    unsigned nvox;
};

// sums of a and b over the CTA, returned to every thread
__device__ __forceinline__ void lpv_block_sum2(unsigned& a, unsigned& b, unsigned (*sh)[LPV_THREADS / 32]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    __syncthreads();   // sh may still be read from a previous call
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    a = 0; b = 0;
#pragma unroll
    for (int w = 0; w < LPV_THREADS / 32; ++w) { a += sh[0][w]; b += sh[1][w]; }
}
// exclusive prefix of v over the threads of the CTA; total to every thread
__device__ __forceinline__ unsigned lpv_block_excl(unsigned v, unsigned& total, unsigned (*sh)[LPV_THREADS / 32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    __syncthreads();
    if (lane == 31) sh[0][warp] = incl;
    __syncthreads();
    unsigned before = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < LPV_THREADS / 32; ++w) {
        const unsigned s = sh[0][w];
        if (w < warp) before += s;
        total += s;
    }
    return before + incl - v;
}
// lamp masks of the 64 voxels a thread owns in chunk c (bit k of m[q] = voxel 16 q + k has an emissive texture)
__device__ __forceinline__ unsigned lpv_lamp_masks(const LpvCoopArgs& a, int c, const unsigned char* em, unsigned m[4]) {
    const size_t first = (size_t)c * LPV_SCAN_CHUNK + (size_t)threadIdx.x * LPV_SCAN_VOX;
    unsigned n = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        m[q] = 0;
        if (first + 16 * q < a.nvox) {   // nvox % 16 == 0
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.g.blocks + first + 16 * q));
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) m[q] |= (unsigned)em[(w[k >> 2] >> (8 * (k & 3))) & 0xffu] << k;
        }
        n += __popc(m[q]);
    }
    return n;
}

__global__ void __launch_bounds__(LPV_THREADS) lpv_repropagate_coop_kernel(const LpvCoopArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned char em[256];
    __shared__ unsigned sh[2][LPV_THREADS / 32];
    const LpvGrid& g = a.g;
    const int tid = threadIdx.x, G = gridDim.x;
    const bool scan = a.lights == nullptr;
    const int chunks = (int)((a.nvox + LPV_SCAN_CHUNK - 1) / LPV_SCAN_CHUNK), cpc = (chunks + G - 1) / G;
    const int c0 = min(chunks, (int)blockIdx.x * cpc), c1 = min(chunks, c0 + cpc);
    // BlockEmissiveData[id] >= 0 (row 3 of the table); ids >= 128 have no entry
    em[tid] = tid < 128 ? (a.block_data[3 * 128 + tid] >= 0 ? 1 : 0) : 0;
    __syncthreads();

    // ---- clear both volumes (ClearEntireVolume :228-232); count the lamps of this CTA's chunks ----
    {
        uint4* vol = reinterpret_cast<uint4*>(g.level);   // level, then block type: 2 * nvox bytes
        const unsigned nq = a.nvox / 8, stride = G * LPV_THREADS;
        unsigned i = blockIdx.x * LPV_THREADS + tid;
        for (; i + 3 * stride < nq; i += 4 * stride) {   // four stores in flight per thread: few CTAs have to cover the store latency
            vol[i] = make_uint4(0u, 0u, 0u, 0u); vol[i + stride] = make_uint4(0u, 0u, 0u, 0u);
            vol[i + 2 * stride] = make_uint4(0u, 0u, 0u, 0u); vol[i + 3 * stride] = make_uint4(0u, 0u, 0u, 0u);
        }
        for (; i < nq; i += stride) vol[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (scan)
        for (int c = c0; c < c1; ++c) {
            unsigned m[4], mine = lpv_lamp_masks(a, c, em, m), zero = 0;
            lpv_block_sum2(mine, zero, sh);
            if (tid == 0) a.chunk_counts[c] = mine;
        }
    grid.sync();

    // ---- seeds (AddLightToVolume :190-205) in queue order: the scan's ascending voxel order, or the caller's list ----
    unsigned n;
    if (scan) {
        unsigned before = 0, total = 0;
        for (int i = tid; i < chunks; i += LPV_THREADS) {
            const unsigned v = a.chunk_counts[i];
            total += v;
            if (i < c0) before += v;
        }
        lpv_block_sum2(before, total, sh);
        n = total;
        unsigned running = before;
        for (int c = c0; c < c1; ++c) {
            unsigned m[4], chunk_total;
            const unsigned mine = lpv_lamp_masks(a, c, em, m);
            unsigned at = running + lpv_block_excl(mine, chunk_total, sh);
            const size_t first = (size_t)c * LPV_SCAN_CHUNK + (size_t)tid * LPV_SCAN_VOX;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned mq = m[q];
                while (mq) {
                    const int k = __ffs(mq) - 1;
                    mq &= mq - 1;
                    const int idx = (int)(first + 16 * q + k);
                    const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
                    const bool in = lpv_inside(g, x, y, z);
                    if (in) { g.level[idx] = (uint8_t)a.seed_level; g.color[idx] = g.blocks[idx]; }
                    a.front0[at++] = in ? idx : -1;
                }
            }
            running += chunk_total;
        }
    } else {
        n = (unsigned)a.n_lights;
        for (unsigned i = blockIdx.x * LPV_THREADS + tid; i < n; i += G * LPV_THREADS) {
            const int x = a.lights[3 * i], y = a.lights[3 * i + 1], z = a.lights[3 * i + 2];
            int idx = -1;
            if (lpv_inside(g, x, y, z)) {
                idx = lpv_index(g, x, y, z);
                g.level[idx] = (uint8_t)a.seed_level;
                g.color[idx] = g.blocks[idx];
            }
            a.front0[i] = idx;
        }
    }
    grid.sync();

    // ---- one wave per level ----
    const int* front = a.front0;
    int* next = a.front1;
    for (int wave = 0; wave + 3 <= a.seed_level && n > 0; ++wave) {
        // every node of a wave carries the same level (the seeds min(limit, 8), each wave one less): no need to load it
        const int cur = a.seed_level - wave;
        // claim
        for (unsigned i = blockIdx.x * LPV_THREADS + tid; i < n; i += G * LPV_THREADS) {
            const int idx = front[i];
            if (idx < 0) continue;
            const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                int dx, dy, dz;
                lpv_dir(k, true, dx, dy, dz);
                if (!lpv_inside(g, x + dx, y + dy, z + dz)) continue;
                const int q = lpv_index(g, x + dx, y + dy, z + dz);
                if (g.blocks[q] == 0 && (int)g.level[q] + 2 < cur) atomicMin(a.claim + q, i * 6u + (unsigned)k);
            }
        }
        grid.sync();
        // which claims were won; per-CTA totals
        unsigned begin, end;
        lpv_chunk(n, begin, end);
        {
            unsigned mine = 0, zero = 0;
            for (unsigned i = begin + tid; i < end; i += LPV_THREADS) {
                const int idx = front[i];
                unsigned m = 0;
                if (idx >= 0) {
                    const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        int dx, dy, dz;
                        lpv_dir(k, true, dx, dy, dz);
                        if (lpv_inside(g, x + dx, y + dy, z + dz) && __ldcg(a.claim + lpv_index(g, x + dx, y + dy, z + dz)) == i * 6u + (unsigned)k) m |= 1u << k;
                    }
                }
                a.wins[i] = (uint8_t)m;
                mine += __popc(m);
            }
            lpv_block_sum2(mine, zero, sh);
            if (tid == 0) a.counts[blockIdx.x] = mine;
        }
        grid.sync();
        // ordered write of the winners: this CTA's offset and the size of the next wave from the per-CTA totals
        unsigned before = 0, total = 0;
        for (int i = tid; i < G; i += LPV_THREADS) {
            const unsigned v = __ldcg(a.counts + i);
            total += v;
            if (i < (int)blockIdx.x) before += v;
        }
        lpv_block_sum2(before, total, sh);
        unsigned running = before;
        for (unsigned base = begin; base < end; base += LPV_THREADS) {
            const unsigned i = base + tid;
            unsigned m = i < end ? a.wins[i] : 0u, tile_total;
            const unsigned cnt = __popc(m);
            unsigned at = running + lpv_block_excl(cnt, tile_total, sh);
            if (m) {
                const int idx = front[i];
                const uint8_t next_level = (uint8_t)(cur - 1), type = g.color[idx];
                const int z = idx / (g.nx * g.ny), r = idx - z * g.nx * g.ny, y = r / g.nx, x = r - y * g.nx;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    int dx, dy, dz;
                    lpv_dir(k, true, dx, dy, dz);
                    const int q = lpv_index(g, x + dx, y + dy, z + dz);
                    g.color[q] = type;
                    g.level[q] = next_level;
                    a.claim[q] = NO_CLAIM;
                    next[at++] = q;
                }
            }
            running += tile_total;
        }
        n = total;
        const int* t = front; front = next; next = const_cast<int*>(t);
        grid.sync();
    }
}

// ---- PrecomputeAverageBlockColor.comp main() (:23-54), dispatched once by Volumetrics::CreateVolume (VolumetricFloodFill.cpp:102-123) ----
// One thread per block id: ten trilinear samples of the block's albedo layer (five at LOD 8, five at LOD 6 / 6.5 / 6.5 / 5.5 / 5.5),
// summed in the shader's order, / 10, pow 1.8.  The consumers (u_LPVGI of the reflection trace, the colour pass) index it with the block type
// volume.
__global__ void __launch_bounds__(128) lpv_average_colors_kernel(TexArrayDev albedo, const int32_t* __restrict__ block_data, float4* __restrict__ out) {
    const int id = threadIdx.x;
    const int layer = block_data[id];   // BlockAlbedoData
    float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (layer >= 0) {
        const float at[5] = {0.5f, 0.25f, 0.75f, 1.0f, 0.0f}, lod2[5] = {6.0f, 6.5f, 6.5f, 5.5f, 5.5f};
        f3 a = F3(0.0f, 0.0f, 0.0f), b = a;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const f4 t = texarray_sample(albedo, at[k], at[k], (float)layer, 8.0f);
            a = k == 0 ? F3(t.x, t.y, t.z) : F3(a.x + t.x, a.y + t.y, a.z + t.z);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const f4 t = texarray_sample(albedo, at[k], at[k], (float)layer, lod2[k]);
            b = k == 0 ? F3(t.x, t.y, t.z) : F3(b.x + t.x, b.y + t.y, b.z + t.z);
        }
        const float d = 5.0f * 2.0f;
        o = make_float4(powf((a.x + b.x) / d, 1.8f), powf((a.y + b.y) / d, 1.8f), powf((a.z + b.z) / d, 1.8f), 0.0f);
    }
    out[id] = o;
}

// ---- SampleLPVData as a batch function on caller points (lpv_sample.cuh holds the arithmetic): one thread per point ----
__global__ void __launch_bounds__(256) lpv_sample_kernel(const LpvSampleArgs a, const float* __restrict__ points, int n, float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const f3 r = lpv_sample_data(a, F3(points[3 * p], points[3 * p + 1], points[3 * p + 2]), F3(a.dx, a.dy, a.dz));
    out[3 * p] = r.x; out[3 * p + 1] = r.y; out[3 * p + 2] = r.z;
}

// ---- block edits: the exact FIFO of DepropogateVolume / PropogateVolume, one warp -------------------------------------------------

// a queue entry: x + 1, y + 1, z + 1 in 16 bits each (the edit queues the six neighbours of a voxel unchecked, so -1 .. n occur) and,
// for the removal queue, the level the node had when it was queued
__device__ __forceinline__ unsigned long long lpv_pack(int x, int y, int z, int light) {
    return (unsigned long long)(unsigned)(x + 1) | ((unsigned long long)(unsigned)(y + 1) << 16) | ((unsigned long long)(unsigned)(z + 1) << 32) |
           ((unsigned long long)(unsigned)light << 48);
}
__device__ __forceinline__ void lpv_unpack(unsigned long long e, int& x, int& y, int& z, int& light) {
    x = (int)(e & 0xffffu) - 1; y = (int)((e >> 16) & 0xffffu) - 1; z = (int)((e >> 32) & 0xffffu) - 1; light = (int)(e >> 48);
}

struct LpvQueue {
    unsigned long long* q;
    unsigned mask;          // capacity - 1 (power of two)
    unsigned head, tail;    // warp-uniform
    // entries of the lanes with `pred`, in lane order (the reference pushes in direction order)
    __device__ __forceinline__ void push(bool pred, unsigned long long e, int& overflow) {
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (pred) __stcg(q + ((tail + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))) & mask), e);
        tail += __popc(m);
        if (tail - head > mask) overflow = 1;
    }
};

__device__ __forceinline__ int lpv_get_level(const LpvGrid& g, int x, int y, int z) {   // GetLightValue :125-138
    return lpv_inside(g, x, y, z) ? (int)__ldcg(g.level + lpv_index(g, x, y, z)) : 0;
}

__global__ void __launch_bounds__(32) lpv_edit_kernel(LpvGrid g, const int32_t* __restrict__ block_data, int op, int px, int py, int pz, int block,
                                                      int seed_level, unsigned long long* light_q, unsigned long long* removal_q, unsigned mask,
                                                      int* overflow_out) {
    const int lane = threadIdx.x;
    int overflow = 0;   // warp-uniform; reported through host-mapped memory at the end
    LpvQueue lq{light_q, mask, 0u, 0u}, rq{removal_q, mask, 0u, 0u};
    // GetBlockEmissiveTexture(block) >= 0 / HasEmissiveTexture(block): the emissive row of the block table; ids without an entry have none
    const bool emissive = block >= 0 && block < 128 && block_data[3 * 128 + block] >= 0;
    if (op == 1) {
        // World.cpp:273-333: the voxel and its six neighbours (+x -x +y -y +z -z) join the removal queue with their present levels,
        // then a placed lamp is seeded
        int dx = 0, dy = 0, dz = 0;
        if (lane >= 1 && lane < 7) lpv_dir(lane - 1, false, dx, dy, dz);
        rq.push(lane < 7, lpv_pack(px + dx, py + dy, pz + dz, lpv_get_level(g, px + dx, py + dy, pz + dz)), overflow);
        if (emissive) {
            if (lane == 0 && lpv_inside(g, px, py, pz)) {
                __stcg(g.color + lpv_index(g, px, py, pz), (uint8_t)block);
                __stcg(g.level + lpv_index(g, px, py, pz), (uint8_t)seed_level);
            }
            lq.push(lane == 0, lpv_pack(px, py, pz, 0), overflow);
        }
    } else {
        // World.cpp:395-446: a broken lamp queues its own removal and is cleared; the six neighbours are queued for propagation
        if (emissive) {
            rq.push(lane == 0, lpv_pack(px, py, pz, lpv_get_level(g, px, py, pz)), overflow);
            if (lane == 0 && lpv_inside(g, px, py, pz)) {
                __stcg(g.color + lpv_index(g, px, py, pz), (uint8_t)0);
                __stcg(g.level + lpv_index(g, px, py, pz), (uint8_t)0);
            }
        }
        int dx = 0, dy = 0, dz = 0;
        if (lane < 6) lpv_dir(lane, false, dx, dy, dz);
        lq.push(lane < 6, lpv_pack(px + dx, py + dy, pz + dz, 0), overflow);
    }
    __syncwarp();
    for (int it = 0; it < 4; ++it) {   // World.cpp:482-485
        // DepropogateVolume :328-468
        while (rq.head != rq.tail) {
            const unsigned batch = min(32u, rq.tail - rq.head);
            const unsigned long long mine = (unsigned)lane < batch ? __ldcg(rq.q + ((rq.head + lane) & rq.mask)) : 0ull;
            for (unsigned j = 0; j < batch; ++j) {
                int x, y, z, light;
                lpv_unpack(__shfl_sync(0xffffffffu, mine, j), x, y, z, light);
                int dx = 0, dy = 0, dz = 0;
                if (lane < 6) lpv_dir(lane, false, dx, dy, dz);
                x += dx; y += dy; z += dz;
                const bool in = lane < 6 && lpv_inside(g, x, y, z);
                const int nl = in ? (int)__ldcg(g.level + lpv_index(g, x, y, z)) : 0;
                const bool zero = in && nl != 0 && nl < light, relight = in && !zero && nl >= light;
                if (zero) __stcg(g.level + lpv_index(g, x, y, z), (uint8_t)0);   // the block type stays (SetLightValue(p, 0, CurrentBlock))
                rq.push(zero, lpv_pack(x, y, z, nl), overflow);
                lq.push(relight, lpv_pack(x, y, z, 0), overflow);
                __syncwarp();
            }
            rq.head += batch;
        }
        // PropogateVolume :247-326
        while (lq.head != lq.tail) {
            const unsigned batch = min(32u, lq.tail - lq.head);
            const unsigned long long mine = (unsigned)lane < batch ? __ldcg(lq.q + ((lq.head + lane) & lq.mask)) : 0ull;
            for (unsigned j = 0; j < batch; ++j) {
                int x, y, z, unused;
                lpv_unpack(__shfl_sync(0xffffffffu, mine, j), x, y, z, unused);
                const bool node_in = lpv_inside(g, x, y, z);
                const int cur = node_in ? (int)__ldcg(g.level + lpv_index(g, x, y, z)) : 0;     // read when popped
                const int type = node_in ? (int)__ldcg(g.color + lpv_index(g, x, y, z)) : 0;
                int dx = 0, dy = 0, dz = 0;
                if (lane < 6) lpv_dir(lane, true, dx, dy, dz);
                x += dx; y += dy; z += dz;
                const bool in = lane < 6 && lpv_inside(g, x, y, z);
                bool take = false;
                if (in) {
                    const int q = lpv_index(g, x, y, z);
                    take = __ldcg(g.blocks + q) == 0 && (int)__ldcg(g.level + q) + 2 < cur;
                    if (take) {
                        __stcg(g.color + q, (uint8_t)type);
                        __stcg(g.level + q, (uint8_t)(cur - 1));
                    }
                }
                lq.push(take, lpv_pack(x, y, z, 0), overflow);
                __syncwarp();
            }
            lq.head += batch;
        }
    }
    if (lane == 0) { *overflow_out = overflow; __threadfence_system(); }
}

}  // namespace

// work memory: claim[N] u32 | front[2][N] i32 | wins[N] u8 | counts[1024] | sizes[16] | overflow[16] | chunk_counts[N / 16384]
static int lpv_ensure(vxrt_ctx* c) {
    const size_t n = c->nvox;
    if (!c->d_lpv) {
        VX_CUDA(cudaMalloc(&c->d_lpv, 2 * n));
        VX_CUDA(cudaMemsetAsync(c->d_lpv, 0, 2 * n, c->stream));   // CreateVolume clears both volumes (:76-91, :99-100)
    }
    if (!c->d_lpv_work) {
        const size_t bytes = 4 * n + 8 * n + n + (1024 + 16 + 16 + (n + LPV_SCAN_CHUNK - 1) / LPV_SCAN_CHUNK) * sizeof(unsigned);
        VX_CUDA(cudaMalloc(&c->d_lpv_work, bytes));
        VX_CUDA(cudaMemsetAsync(c->d_lpv_work, 0xff, 4 * n, c->stream));   // claim[] = NO_CLAIM; every wave leaves it that way
    }
    return VXRT_OK;
}

namespace {
struct LpvWork {
    unsigned* claim;
    int* front[2];
    uint8_t* wins;
    unsigned* counts;
    unsigned* sizes;
    int* overflow;
    unsigned* chunk_counts;
};
LpvWork lpv_work(const vxrt_ctx* c) {
    LpvWork w;
    uint8_t* p = (uint8_t*)c->d_lpv_work;
    const size_t n = c->nvox;
    w.claim = (unsigned*)p;
    w.front[0] = (int*)(p + 4 * n);
    w.front[1] = (int*)(p + 8 * n);
    w.wins = p + 12 * n;
    w.counts = (unsigned*)(p + 13 * n);   // nvox % 16 == 0 (vxrt_cuda_create)
    w.sizes = w.counts + 1024;
    w.overflow = (int*)(w.sizes + 16);
    w.chunk_counts = (unsigned*)(w.overflow + 16);
    return w;
}
LpvGrid lpv_grid(const vxrt_ctx* c) {
    LpvGrid g;
    g.blocks = c->d_blocks; g.level = c->d_lpv; g.color = c->d_lpv + c->nvox; g.nx = c->nx; g.ny = c->ny; g.nz = c->nz;
    return g;
}
}  // namespace

int vxrt_lpv_ensure(vxrt_ctx* c) { return lpv_ensure(c); }

// d_lights: device list of 3 ints per light, *d_count of them (at most `capacity`); both may live in the caller's staging memory
int vxrt_launch_lpv_repropagate(vxrt_ctx* c, const int32_t* d_lights, const unsigned* d_count, int capacity, int limit) {
    int rc = lpv_ensure(c);
    if (rc) return rc;
    const LpvGrid g = lpv_grid(c);
    const LpvWork w = lpv_work(c);
    const int seed_level = limit > 8 ? 8 : limit;   // AddLightToVolume :190-205
    VX_CUDA(cudaMemsetAsync(c->d_lpv, 0, 2 * c->nvox, c->stream));   // ClearEntireVolume :228-232
    int grid = c->sm_count * 4;
    if (grid > 1024) grid = 1024;
    lpv_seed_kernel<<<grid, LPV_THREADS, 0, c->stream>>>(g, d_lights, d_count, capacity, seed_level, w.front[0], w.sizes);
    c->launches += 1;
    for (int wave = 0; wave + 3 <= seed_level; ++wave) {   // nodes of level cur spread while cur >= 3
        const int* front = w.front[wave & 1];
        int* next = w.front[(wave + 1) & 1];
        lpv_claim_kernel<<<grid, LPV_THREADS, 0, c->stream>>>(g, w.claim, front, w.sizes + wave);
        lpv_count_kernel<<<grid, LPV_THREADS, 0, c->stream>>>(g, w.claim, front, w.sizes + wave, w.wins, w.counts);
        lpv_scan_kernel<<<1, 1024, 0, c->stream>>>(w.counts, grid, w.sizes + wave + 1);
        lpv_write_kernel<<<grid, LPV_THREADS, 0, c->stream>>>(g, w.claim, front, w.sizes + wave, w.wins, w.counts, next);
        c->launches += 4;
    }
    VX_CUDA(cudaGetLastError());
    return VXRT_OK;
}

// d_lights == nullptr: the lamps are scanned from the grid inside the kernel.  Returns VXRT_E_UNSUPPORTED when the device cannot
// launch cooperatively (the caller falls back to the multi-kernel path).
int vxrt_launch_lpv_repropagate_coop(vxrt_ctx* c, const int32_t* d_lights, int n_lights, int limit) {
    static int coop_grid = -1;   // 0: unsupported
    if (coop_grid < 0) {
        int supported = 0, per_sm = 0;
        if (cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, c->device) != cudaSuccess) supported = 0;
        if (supported && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lpv_repropagate_coop_kernel, LPV_THREADS, 0) != cudaSuccess) per_sm = 0;
        int cap = 2;                  // the cost of a grid-wide barrier grows with the number of CTAs (tools/debug/time_lpv.py)
        if (const char* e = getenv("VXRT_LPV_CTAS_PER_SM")) cap = atoi(e) > 0 ? atoi(e) : cap;
        if (per_sm > cap) per_sm = cap;
        coop_grid = supported ? per_sm * c->sm_count : 0;
        if (coop_grid > 1024) coop_grid = 1024;
    }
    if (coop_grid <= 0) return VXRT_E_UNSUPPORTED;
    int rc = lpv_ensure(c);
    if (rc) return rc;
    const LpvWork w = lpv_work(c);
    LpvCoopArgs a;
    a.g = lpv_grid(c); a.claim = w.claim; a.front0 = w.front[0]; a.front1 = w.front[1]; a.wins = w.wins; a.counts = w.counts;
    a.chunk_counts = w.chunk_counts; a.block_data = c->d_block_data; a.lights = d_lights; a.n_lights = n_lights;
    a.seed_level = limit > 8 ? 8 : limit; a.nvox = (unsigned)c->nvox;
    void* args[] = {&a};
    VX_CUDA(cudaLaunchCooperativeKernel((void*)lpv_repropagate_coop_kernel, dim3(coop_grid), dim3(LPV_THREADS), args, 0, c->stream));
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_lpv_average_colors(vxrt_ctx* c) {
    if (!c->d_lpv_avg) VX_CUDA(cudaMalloc(&c->d_lpv_avg, 128 * sizeof(float4)));
    lpv_average_colors_kernel<<<1, 128, 0, c->stream>>>(c->tex[0], c->d_block_data, (float4*)c->d_lpv_avg);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_lpv_sample(vxrt_ctx* c, const float* d_points, int n, const float dither[3], float* d_out) {
    LpvSampleArgs a;
    a.level = c->d_lpv; a.type = c->d_lpv + c->nvox; a.avg = (const float4*)c->d_lpv_avg; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz;
    a.dx = dither[0]; a.dy = dither[1]; a.dz = dither[2];
    lpv_sample_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a, d_points, n, d_out);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_lpv_edit(vxrt_ctx* c, int op, int x, int y, int z, int block, int limit, int* overflowed) {
    int rc = lpv_ensure(c);
    if (rc) return rc;
    const LpvWork w = lpv_work(c);
    // the two queues share the frontier buffers: 4 N bytes each = N / 2 entries, rounded down to a power of two
    unsigned cap = 1;
    while ((size_t)cap * 2 * sizeof(unsigned long long) <= 4 * c->nvox) cap *= 2;
    // the kernel reports queue overflow through pinned host memory it writes directly: one launch + one stream synchronisation per edit
    if (!c->h_lpv_flag) {
        VX_CUDA(cudaHostAlloc((void**)&c->h_lpv_flag, sizeof(int), cudaHostAllocMapped));
        VX_CUDA(cudaHostGetDevicePointer((void**)&c->d_lpv_flag, c->h_lpv_flag, 0));
    }
    lpv_edit_kernel<<<1, 32, 0, c->stream>>>(lpv_grid(c), c->d_block_data, op, x, y, z, block, limit > 8 ? 8 : limit,
                                             (unsigned long long*)w.front[0], (unsigned long long*)w.front[1], cap - 1, c->d_lpv_flag);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    VX_CUDA(cudaStreamSynchronize(c->stream));
    *overflowed = *(volatile int*)c->h_lpv_flag;
    return VXRT_OK;
}
