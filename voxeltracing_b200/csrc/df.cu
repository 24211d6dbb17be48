// df.cu — Manhattan distance-field generation over the block-id grid (sm_100a).
//
// Replaces World::GenerateDistanceField (Core/World.cpp:69-113) and the three compute shaders
// ManhattanDistanceX/Y/Z.comp.  Result: df = min(254, L1 distance to nearest non-zero block),
// bit-exact with the shaders (integer semantics).
//
// B200 design (not the reference's one-invocation-per-line walk through an r8 image): two kernels, each
// reading its input once and writing its output once, every sweep in shared memory.
//   kernel 1 (xy_slice): one CTA per z-slice.  The slice (nx*ny bytes) is read with 16-byte loads, all of a
//     thread's loads in flight before the first use, converted to the solid?0:254 seed and staged in shared
//     memory as rows of uint4 with an odd row stride (in uint4 units), which makes the per-row X sweep
//     (LDS.128, one thread per row), the per-column Y sweep (LDS.32, one thread per 4-voxel word column) and the
//     16-byte load / store phases all bank-conflict free.  X sweep: running value in a register, one VIADDMNMX
//     per voxel.  Y sweep: two VIADDMNMX.U16x2 per 4 voxels per direction (bytes unpacked to u16 lanes with PRMT).
//   kernel 2 (z_tile): one CTA per tile of 32 word columns (128 bytes of x) x all planes.  The tile is staged in
//     shared memory (128-byte rows are full sectors; the field is L2 resident after kernel 1), the z range is
//     cut into 8 segments, one warp each: local forward + backward min-plus sweeps per segment, then the carries
//     of the other segments (their boundary planes + distance) are folded in while the tile streams back out.
//     The first version of this pass (one thread per column walking all planes) exposed 12,288 threads with a
//     768-step dependent chain each and took 67 us of the 77 us regeneration.
// Algorithmic HBM traffic: read N block bytes + write N distance bytes = 2N.
#include "ctx.h"

namespace {

__device__ __forceinline__ unsigned even_lanes(unsigned w) { return __byte_perm(w, 0u, 0x4240); }  // (b0, b2)
__device__ __forceinline__ unsigned odd_lanes(unsigned w) { return __byte_perm(w, 0u, 0x4341); }   // (b1, b3)
__device__ __forceinline__ unsigned pack_lanes(unsigned e, unsigned o) { return __byte_perm(e, o, 0x6240); }

// solid ? 0 : maxd for the four bytes of w
__device__ __forceinline__ unsigned seed_word(unsigned w, unsigned maxd4) {
    unsigned nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;  // bit7 set where byte != 0
    unsigned mask = (nz >> 7) * 0xffu;                                   // 0xff where byte != 0
    return ~mask & maxd4;
}

// forward min-plus step through the four bytes of w (low byte first); c = running value
__device__ __forceinline__ unsigned sweep_word_up(unsigned w, unsigned& c) {
    const unsigned b0 = __viaddmin_u32(c, 1u, w & 0xffu);
    const unsigned b1 = __viaddmin_u32(b0, 1u, (w >> 8) & 0xffu);
    const unsigned b2 = __viaddmin_u32(b1, 1u, (w >> 16) & 0xffu);
    const unsigned b3 = __viaddmin_u32(b2, 1u, w >> 24);
    c = b3;
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
__device__ __forceinline__ unsigned sweep_word_down(unsigned w, unsigned& c) {
    const unsigned b3 = __viaddmin_u32(c, 1u, w >> 24);
    const unsigned b2 = __viaddmin_u32(b3, 1u, (w >> 16) & 0xffu);
    const unsigned b1 = __viaddmin_u32(b2, 1u, (w >> 8) & 0xffu);
    const unsigned b0 = __viaddmin_u32(b1, 1u, w & 0xffu);
    c = b0;
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// ---- kernel 1: X and Y sweeps of one z-slice in shared memory ---------------------------------
// ManhattanDistanceX.comp:53-68 and ManhattanDistanceY.comp:35-49.
constexpr int XY_THREADS = 256;
constexpr int XY_BATCH = 6;  // 16-byte loads in flight per thread

__global__ void __launch_bounds__(XY_THREADS) df_xy_slice_kernel(const uint8_t* __restrict__ blocks,
                                                                 uint8_t* __restrict__ df, int nx, int ny,
                                                                 int z_begin, unsigned maxd) {
    extern __shared__ uint4 smem4[];
    unsigned* smem = reinterpret_cast<unsigned*>(smem4);
    const int qpr = nx >> 4;       // 16-byte quads per row
    const int sq = qpr | 1;        // row stride in quads (odd)
    const int sw = sq << 2;        // row stride in words
    const int nxw = nx >> 2;       // words per row
    const int nq = qpr * ny;       // quads in the slice (<= 4096)
    const unsigned rdiv = ((1u << 20) + qpr - 1) / qpr;  // q / qpr == (q * rdiv) >> 20 for q < 4096, qpr <= 64
    const int z = z_begin + blockIdx.x;
    const size_t slice_off = (size_t)z * nx * ny;
    const uint4* src = reinterpret_cast<const uint4*>(blocks + slice_off);
    const unsigned maxd4 = maxd * 0x01010101u;

    for (int base = 0; base < nq; base += XY_THREADS * XY_BATCH) {
        uint4 v[XY_BATCH];
#pragma unroll
        for (int i = 0; i < XY_BATCH; ++i) {
            const int q = base + i * XY_THREADS + threadIdx.x;
            if (q < nq) v[i] = __ldg(src + q);
        }
#pragma unroll
        for (int i = 0; i < XY_BATCH; ++i) {
            const int q = base + i * XY_THREADS + threadIdx.x;
            if (q < nq) {
                const int row = (int)(((unsigned)q * rdiv) >> 20), col = q - row * qpr;
                smem4[row * sq + col] = make_uint4(seed_word(v[i].x, maxd4), seed_word(v[i].y, maxd4), seed_word(v[i].z, maxd4),
                                                   seed_word(v[i].w, maxd4));
            }
        }
    }
    __syncthreads();

    // X sweep: d[x] = min(seed[x], d[x-1]+1) forward, then d[x] = min(d[x], d[x+1]+1) backward.
    for (int row = threadIdx.x; row < ny; row += XY_THREADS) {
        uint4* r = smem4 + row * sq;
        unsigned c = 255u;  // min(seed, 256) == seed for the first voxel
#pragma unroll 2
        for (int j = 0; j < qpr; ++j) {
            uint4 w = r[j];
            w.x = sweep_word_up(w.x, c); w.y = sweep_word_up(w.y, c); w.z = sweep_word_up(w.z, c); w.w = sweep_word_up(w.w, c);
            r[j] = w;
        }
        c = 255u;
#pragma unroll 2
        for (int j = qpr - 1; j >= 0; --j) {
            uint4 w = r[j];
            w.w = sweep_word_down(w.w, c); w.z = sweep_word_down(w.z, c); w.y = sweep_word_down(w.y, c); w.x = sweep_word_down(w.x, c);
            r[j] = w;
        }
    }
    __syncthreads();

    // Y sweep on 4-voxel word columns, two u16x2 lane pairs per word.
    for (int col = threadIdx.x; col < nxw; col += XY_THREADS) {
        unsigned* cptr = smem + col;
        unsigned w = cptr[0];
        unsigned e = even_lanes(w), o = odd_lanes(w);
#pragma unroll 8
        for (int y = 1; y < ny; ++y) {
            w = cptr[y * sw];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * sw] = pack_lanes(e, o);
        }
#pragma unroll 8
        for (int y = ny - 2; y >= 0; --y) {
            w = cptr[y * sw];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * sw] = pack_lanes(e, o);
        }
    }
    __syncthreads();

    uint4* dst = reinterpret_cast<uint4*>(df + slice_off);
    for (int q = threadIdx.x; q < nq; q += XY_THREADS) {
        const int row = (int)(((unsigned)q * rdiv) >> 20), col = q - row * qpr;
        dst[q] = smem4[row * sq + col];
    }
}

// ---- kernel 2: Z sweeps (ManhattanDistanceZ.comp:31-46) ----------------------------------------
// Planes [z0, z1) of one tile of Z_TILE_WORDS word columns.  Warp s owns the planes [s*seg, (s+1)*seg) of the
// range (local index); lane = word column.
constexpr int Z_TILE_WORDS = 32;
constexpr int Z_TILE_QUADS = Z_TILE_WORDS / 4;
constexpr int Z_SEGS = 8;
constexpr int Z_THREADS = Z_SEGS * 32;
constexpr int Z_BATCH = 6;

__global__ void __launch_bounds__(Z_THREADS) df_z_tile_kernel(uint8_t* __restrict__ df, int words_per_plane, int z0, int z1, int seg) {
    extern __shared__ uint4 smem4[];
    const int nzr = z1 - z0;
    uint4* tile4 = smem4;                          // [nzr][Z_TILE_QUADS]
    uint4* carry = smem4 + nzr * Z_TILE_QUADS;     // [Z_SEGS][4][Z_TILE_QUADS]: (fwd.e, fwd.o, bwd.e, bwd.o) per column
    unsigned* tile = reinterpret_cast<unsigned*>(tile4);
    const int col0 = blockIdx.x * Z_TILE_WORDS;    // first word column of the tile
    const size_t qpp = (size_t)(words_per_plane >> 2);
    uint4* g4 = reinterpret_cast<uint4*>(df) + (size_t)z0 * qpp + (col0 >> 2);
    const int nitems = nzr * Z_TILE_QUADS;
    const int my_q = threadIdx.x & (Z_TILE_QUADS - 1);
    const bool q_ok = col0 + my_q * 4 < words_per_plane;  // words_per_plane % 4 == 0: a quad never straddles the end

    // ---- stage the tile ----
    for (int base = 0; base < nitems; base += Z_THREADS * Z_BATCH) {
        uint4 v[Z_BATCH];
#pragma unroll
        for (int i = 0; i < Z_BATCH; ++i) {
            const int it = base + i * Z_THREADS + threadIdx.x;
            v[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (it < nitems && q_ok) v[i] = g4[(size_t)(it >> 3) * qpp + my_q];
        }
#pragma unroll
        for (int i = 0; i < Z_BATCH; ++i) {
            const int it = base + i * Z_THREADS + threadIdx.x;
            if (it < nitems) tile4[it] = v[i];
        }
    }
    __syncthreads();

    // ---- local sweeps of this warp's segment ----
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = min(s * seg, nzr), b = min(a + seg, nzr);
    if (a < b) {
        unsigned* cptr = tile + lane;
        unsigned e = 0x00ff00ffu, o = 0x00ff00ffu;  // min(v, 256) == v for the first plane
#pragma unroll 8
        for (int z = a; z < b; ++z) {
            const unsigned w = cptr[z * Z_TILE_WORDS];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[z * Z_TILE_WORDS] = pack_lanes(e, o);
        }
#pragma unroll 8
        for (int z = b - 2; z >= a; --z) {
            const unsigned w = cptr[z * Z_TILE_WORDS];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[z * Z_TILE_WORDS] = pack_lanes(e, o);
        }
    }
    __syncthreads();

    // ---- carries into this segment: boundary planes of the other segments + distance ----
    // After the local sweeps the last plane of segment t holds min over its planes z' of v[z'] + (last_t - z') and the
    // first plane holds min v[z'] + (z' - first_t); a voxel of segment s at plane z then sees
    //   min( local, fwd + (z - a), bwd + (b - 1 - z) ),  fwd = min_{t<s} last_t-value + (a - last_t),
    //                                                     bwd = min_{t>s} first_t-value + (first_t - (b - 1)).
    if (a < b) {
        unsigned fe = 0x03ff03ffu, fo = 0x03ff03ffu, be = 0x03ff03ffu, bo = 0x03ff03ffu;  // "no carry": above every distance
        for (int t = 0; t < Z_SEGS; ++t) {
            const int ta = min(t * seg, nzr), tb = min(ta + seg, nzr);
            if (t == s || ta >= tb) continue;
            if (t < s) {
                const unsigned w = tile[(tb - 1) * Z_TILE_WORDS + lane];
                const unsigned d = (unsigned)(a - (tb - 1)), d2 = d | (d << 16);
                fe = __viaddmin_u16x2(even_lanes(w), d2, fe);
                fo = __viaddmin_u16x2(odd_lanes(w), d2, fo);
            } else {
                const unsigned w = tile[ta * Z_TILE_WORDS + lane];
                const unsigned d = (unsigned)(ta - (b - 1)), d2 = d | (d << 16);
                be = __viaddmin_u16x2(even_lanes(w), d2, be);
                bo = __viaddmin_u16x2(odd_lanes(w), d2, bo);
            }
        }
        carry[s * Z_TILE_WORDS + (lane & 3) * Z_TILE_QUADS + (lane >> 2)] = make_uint4(fe, fo, be, bo);
    }
    __syncthreads();

    // ---- fold the carries in while the tile streams out ----
    for (int it = threadIdx.x; it < nitems; it += Z_THREADS) {
        const int z = it >> 3;
        const int t = z / seg, ta = t * seg, tb = min(ta + seg, nzr);
        const unsigned df_ = (unsigned)(z - ta), db_ = (unsigned)(tb - 1 - z);
        const unsigned df2 = df_ | (df_ << 16), db2 = db_ | (db_ << 16);
        uint4 v = tile4[it];
        unsigned* vw = reinterpret_cast<unsigned*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 c = carry[t * Z_TILE_WORDS + k * Z_TILE_QUADS + my_q];
            unsigned e = even_lanes(vw[k]), o = odd_lanes(vw[k]);
            e = __viaddmin_u16x2(c.x, df2, e);
            o = __viaddmin_u16x2(c.y, df2, o);
            e = __viaddmin_u16x2(c.z, db2, e);
            o = __viaddmin_u16x2(c.w, db2, o);
            vw[k] = pack_lanes(e, o);
        }
        if (q_ok) g4[(size_t)z * qpp + my_q] = v;
    }
}

// ---- z-slab sharding (multi-GPU regeneration) -------------------------------------
// After the slab-local sweeps every rank holds L[z] = min over its own planes z' of xy[z'] + |z - z'|.
// With B_t = L on the last plane of slab t and F_t = L on the first plane of slab t (all-gathered, one
// nx*ny plane each), the global transform on slab s is
//   D[z] = min( L[z],  min_{t<s} B_t + (z - (z1_t - 1)),  min_{t>s} F_t + (z0_t - z) ),  clamped to 254.
__global__ void __launch_bounds__(256) df_slab_apply_kernel(uint8_t* __restrict__ df, int words_per_plane, int slab, int nslabs,
                                                            const int* __restrict__ slab_z0, const uint8_t* __restrict__ first_planes,
                                                            const uint8_t* __restrict__ last_planes) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int z0 = slab_z0[slab], z1 = slab_z0[slab + 1];
    const int z = z0 + blockIdx.y;
    if (col >= words_per_plane || z >= z1) return;
    const unsigned* F = reinterpret_cast<const unsigned*>(first_planes);
    const unsigned* B = reinterpret_cast<const unsigned*>(last_planes);
    unsigned* p = reinterpret_cast<unsigned*>(df) + (size_t)z * words_per_plane + col;
    const unsigned w = *p;
    unsigned e = even_lanes(w), o = odd_lanes(w);
    for (int t = 0; t < nslabs; ++t) {
        if (t == slab) continue;
        unsigned add, src;
        if (t < slab) { src = B[(size_t)t * words_per_plane + col]; add = (unsigned)(z - (slab_z0[t + 1] - 1)); }
        else { src = F[(size_t)t * words_per_plane + col]; add = (unsigned)(slab_z0[t] - z); }
        const unsigned add2 = add | (add << 16);  // add <= 1023, lanes hold <= 254: no carry between the u16 lanes
        e = __viaddmin_u16x2(even_lanes(src), add2, e);
        o = __viaddmin_u16x2(odd_lanes(src), add2, o);
    }
    *p = pack_lanes(e, o);  // e, o <= their previous value <= 254
}

// glTexSubImage3D single-voxel edits (Core/World.cpp:372-373, 458-459)
__global__ void edit_blocks_kernel(uint8_t* __restrict__ blocks, const int32_t* __restrict__ e, int n, int nx,
                                   int ny) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = e[4 * i + 0], y = e[4 * i + 1], z = e[4 * i + 2], id = e[4 * i + 3];
    blocks[(size_t)x + (size_t)y * nx + (size_t)z * nx * ny] = (uint8_t)id;
}

}  // namespace

static int set_smem_attrs() {
    static bool attr_set = false;
    if (!attr_set) {
        VX_CUDA(cudaFuncSetAttribute(df_xy_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VX_CUDA(cudaFuncSetAttribute(df_z_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    return VXRT_OK;
}

// X + Y sweeps of slices [z0, z1), then the Z sweeps of the same plane range
static int launch_df_range(vxrt_ctx* c, int z0, int z1) {
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    const unsigned maxd = (unsigned)((nx + ny + nz) < 254 ? (nx + ny + nz) : 254);
    const size_t smem_xy = (size_t)ny * ((nx >> 4) | 1) * sizeof(uint4);
    int rc = set_smem_attrs();
    if (rc) return rc;
    df_xy_slice_kernel<<<z1 - z0, XY_THREADS, smem_xy, c->stream>>>(c->d_blocks, c->d_df, nx, ny, z0, maxd);
    VX_CUDA(cudaGetLastError());
    const int wpp = (nx * ny) >> 2, nzr = z1 - z0;
    const int seg = (nzr + Z_SEGS - 1) / Z_SEGS;
    const size_t smem_z = ((size_t)nzr * Z_TILE_QUADS + (size_t)Z_SEGS * Z_TILE_WORDS) * sizeof(uint4);
    df_z_tile_kernel<<<(wpp + Z_TILE_WORDS - 1) / Z_TILE_WORDS, Z_THREADS, smem_z, c->stream>>>(c->d_df, wpp, z0, z1, seg);
    VX_CUDA(cudaGetLastError());
    c->launches += 2;
    return VXRT_OK;
}

int vxrt_launch_distance_field(vxrt_ctx* c) { return launch_df_range(c, 0, c->nz); }

// phase A of the sharded regeneration: X, Y and slab-local Z sweeps on planes [z0, z1)
int vxrt_launch_df_slab_phase_a(vxrt_ctx* c, int z0, int z1) { return launch_df_range(c, z0, z1); }

// phase B: apply the carries of the other slabs (boundary planes are device pointers, nslabs planes each)
int vxrt_launch_df_slab_phase_b(vxrt_ctx* c, int slab, int nslabs, const int* d_slab_z0, int z0, int z1, const void* first_planes,
                                const void* last_planes) {
    const int wpp = (c->nx * c->ny) >> 2;
    dim3 grid((wpp + 255) / 256, z1 - z0);
    df_slab_apply_kernel<<<grid, 256, 0, c->stream>>>(c->d_df, wpp, slab, nslabs, d_slab_z0, (const uint8_t*)first_planes,
                                                      (const uint8_t*)last_planes);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}

int vxrt_launch_edit_blocks(vxrt_ctx* c, const int32_t* d_edits, int n) {
    if (n <= 0) return VXRT_OK;
    edit_blocks_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_blocks, d_edits, n, c->nx, c->ny);
    VX_CUDA(cudaGetLastError());
    c->launches += 1;
    return VXRT_OK;
}
