// df.cu — Manhattan distance-field generation over the block-id grid (sm_100a).
//
// Replaces World::GenerateDistanceField (Core/World.cpp:69-113) and the three compute shaders
// ManhattanDistanceX/Y/Z.comp.  Result: df = min(254, L1 distance to nearest non-zero block),
// bit-exact with the shaders (integer semantics).
//
// B200 design (not the reference's one-invocation-per-line walk through an r8 image):
//   kernel 1 (xy_slice): one CTA per z-slice.  The slice (nx*ny bytes) is read once from HBM
//     with 16-byte loads, converted to the solid?0:254 seed, and staged in shared memory as
//     packed words (row stride padded to an odd number of words => conflict-free for both the
//     per-row X sweep and the per-column Y sweep).  X sweep: one thread per row, running value
//     in a register, one VIADDMNMX per voxel.  Y sweep: one thread per 4-voxel word column,
//     two VIADDMNMX.U16x2 per word (4 voxels) per direction.  The slice is written back once.
//   kernel 2 (z_columns): one thread per 4-voxel word column, software-pipelined (batches of
//     independent 4-byte loads that hit L2: the whole 18.9 MB field is L2 resident after kernel 1)
//     forward and backward min-plus sweeps with VIADDMNMX.U16x2.
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

// ---- kernel 1: X and Y sweeps of one z-slice in shared memory ---------------------------------
// ManhattanDistanceX.comp:53-68 and ManhattanDistanceY.comp:35-49.
__global__ void __launch_bounds__(128) df_xy_slice_kernel(const uint8_t* __restrict__ blocks,
                                                          uint8_t* __restrict__ df, int nx, int ny,
                                                          int z_begin, unsigned maxd) {
    extern __shared__ unsigned smem[];
    const int nxw = nx >> 2;      // words per row
    const int stride = nxw | 1;   // odd stride in words
    const int z = z_begin + blockIdx.x;
    const size_t slice_off = (size_t)z * nx * ny;
    const uint4* src = reinterpret_cast<const uint4*>(blocks + slice_off);
    const unsigned maxd4 = maxd * 0x01010101u;
    const int nq = (nx * ny) >> 4;  // 16-byte quads in the slice
    const int qpr = nx >> 4;        // quads per row

    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        uint4 v = __ldg(src + q);
        int row = q / qpr, col = (q - row * qpr) << 2;
        unsigned* d = smem + row * stride + col;
        d[0] = seed_word(v.x, maxd4);
        d[1] = seed_word(v.y, maxd4);
        d[2] = seed_word(v.z, maxd4);
        d[3] = seed_word(v.w, maxd4);
    }
    __syncthreads();

    // X sweep: d[x] = min(seed[x], d[x-1]+1) forward, then d[x] = min(d[x], d[x+1]+1) backward.
    for (int row = threadIdx.x; row < ny; row += blockDim.x) {
        unsigned* r = smem + row * stride;
        unsigned c = 255u;  // min(seed, 256) == seed for the first voxel
#pragma unroll 4
        for (int j = 0; j < nxw; ++j) {
            unsigned w = r[j];
            unsigned b0 = __viaddmin_u32(c, 1u, w & 0xffu);
            unsigned b1 = __viaddmin_u32(b0, 1u, (w >> 8) & 0xffu);
            unsigned b2 = __viaddmin_u32(b1, 1u, (w >> 16) & 0xffu);
            unsigned b3 = __viaddmin_u32(b2, 1u, w >> 24);
            c = b3;
            r[j] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        c = 255u;
#pragma unroll 4
        for (int j = nxw - 1; j >= 0; --j) {
            unsigned w = r[j];
            unsigned b3 = __viaddmin_u32(c, 1u, w >> 24);
            unsigned b2 = __viaddmin_u32(b3, 1u, (w >> 16) & 0xffu);
            unsigned b1 = __viaddmin_u32(b2, 1u, (w >> 8) & 0xffu);
            unsigned b0 = __viaddmin_u32(b1, 1u, w & 0xffu);
            c = b0;
            r[j] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
    }
    __syncthreads();

    // Y sweep on 4-voxel word columns, two u16x2 lanes-pairs per word.
    for (int col = threadIdx.x; col < nxw; col += blockDim.x) {
        unsigned* cptr = smem + col;
        unsigned w = cptr[0];
        unsigned e = even_lanes(w), o = odd_lanes(w);
#pragma unroll 4
        for (int y = 1; y < ny; ++y) {
            w = cptr[y * stride];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * stride] = pack_lanes(e, o);
        }
#pragma unroll 4
        for (int y = ny - 2; y >= 0; --y) {
            w = cptr[y * stride];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * stride] = pack_lanes(e, o);
        }
    }
    __syncthreads();

    uint4* dst = reinterpret_cast<uint4*>(df + slice_off);
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        int row = q / qpr, col = (q - row * qpr) << 2;
        const unsigned* d = smem + row * stride + col;
        dst[q] = make_uint4(d[0], d[1], d[2], d[3]);
    }
}

// ---- kernel 2: Z sweeps (ManhattanDistanceZ.comp:31-46) ----------------------------------------
// One thread per 4-voxel word column; planes [z0, z1).  Loads are issued in independent batches
// of ZB planes so the serial min-plus chain never waits on L2 latency.
constexpr int ZB = 8;

__global__ void __launch_bounds__(128) df_z_columns_kernel(uint8_t* __restrict__ df, int words_per_plane,
                                                           int z0, int z1) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= words_per_plane) return;
    unsigned* base = reinterpret_cast<unsigned*>(df) + col;
    const size_t ps = (size_t)words_per_plane;

    unsigned e = 0x00ff00ffu, o = 0x00ff00ffu;  // min(v, 256) == v for the first plane
    int z = z0;
    for (; z + ZB <= z1; z += ZB) {
        unsigned w[ZB];
#pragma unroll
        for (int i = 0; i < ZB; ++i) w[i] = base[(size_t)(z + i) * ps];
#pragma unroll
        for (int i = 0; i < ZB; ++i) {
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w[i]));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w[i]));
            base[(size_t)(z + i) * ps] = pack_lanes(e, o);
        }
    }
    for (; z < z1; ++z) {
        unsigned w = base[(size_t)z * ps];
        e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
        o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
        base[(size_t)z * ps] = pack_lanes(e, o);
    }
    // backward: the last plane is final after the forward sweep
    z = z1 - 2;
    for (; z - (ZB - 1) >= z0; z -= ZB) {
        unsigned w[ZB];
#pragma unroll
        for (int i = 0; i < ZB; ++i) w[i] = base[(size_t)(z - i) * ps];
#pragma unroll
        for (int i = 0; i < ZB; ++i) {
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w[i]));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w[i]));
            base[(size_t)(z - i) * ps] = pack_lanes(e, o);
        }
    }
    for (; z >= z0; --z) {
        unsigned w = base[(size_t)z * ps];
        e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
        o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
        base[(size_t)z * ps] = pack_lanes(e, o);
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

static int set_xy_smem_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        VX_CUDA(cudaFuncSetAttribute(df_xy_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    return VXRT_OK;
}

int vxrt_launch_distance_field(vxrt_ctx* c) {
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    const unsigned maxd = (unsigned)((nx + ny + nz) < 254 ? (nx + ny + nz) : 254);
    const int stride = (nx >> 2) | 1;
    const size_t smem = (size_t)ny * stride * sizeof(unsigned);
    int rc = set_xy_smem_attr();
    if (rc) return rc;
    df_xy_slice_kernel<<<nz, 128, smem, c->stream>>>(c->d_blocks, c->d_df, nx, ny, 0, maxd);
    VX_CUDA(cudaGetLastError());
    const int wpp = (nx * ny) >> 2;
    df_z_columns_kernel<<<(wpp + 127) / 128, 128, 0, c->stream>>>(c->d_df, wpp, 0, nz);
    VX_CUDA(cudaGetLastError());
    c->launches += 2;
    return VXRT_OK;
}

// phase A of the sharded regeneration: X, Y and slab-local Z sweeps on planes [z0, z1)
int vxrt_launch_df_slab_phase_a(vxrt_ctx* c, int z0, int z1) {
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    const unsigned maxd = (unsigned)((nx + ny + nz) < 254 ? (nx + ny + nz) : 254);
    const int stride = (nx >> 2) | 1;
    const size_t smem = (size_t)ny * stride * sizeof(unsigned);
    int rc = set_xy_smem_attr();
    if (rc) return rc;
    df_xy_slice_kernel<<<z1 - z0, 128, smem, c->stream>>>(c->d_blocks, c->d_df, nx, ny, z0, maxd);
    VX_CUDA(cudaGetLastError());
    const int wpp = (nx * ny) >> 2;
    df_z_columns_kernel<<<(wpp + 127) / 128, 128, 0, c->stream>>>(c->d_df, wpp, z0, z1);
    VX_CUDA(cudaGetLastError());
    c->launches += 2;
    return VXRT_OK;
}

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
