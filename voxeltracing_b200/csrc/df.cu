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
//     16-byte load / store phases all bank-conflict free.  X: from the solid mask of the row (quad carries + nibble
//     table, see the kernel).  Y sweep: two VIADDMNMX.U16x2 per 4 voxels per direction (bytes unpacked to u16 lanes with PRMT).
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
__device__ __forceinline__ unsigned even_lanes_lop(unsigned w) { return w & 0x00ff00ffu; }          // LOP3 issues at twice the rate of PRMT
// lanes hold values < 256, so the pack is e + 256 * o: one IMAD on the FMA pipe.  Both kernels are bound by the
// half-rate integer (ALU) pipe that VIADDMNMX / PRMT / LOP3 share (ncu: pipe_alu 62 %, math-pipe throttle the top
// stall), so everything that can be a multiply-add is one.
__device__ __forceinline__ unsigned pack_lanes(unsigned e, unsigned o) { return o * 256u + e; }
__device__ __forceinline__ unsigned pack_bytes(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {
    return (b3 * 256u + b2) * 65536u + (b1 * 256u + b0);
}

// byte 3 of the result = the solid mask of the four bytes of w (bit i = byte i non-zero), its high nibble 0; the other bytes are junk
__device__ __forceinline__ unsigned solid_nibble_top(unsigned w) {
    const unsigned nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;  // bit7 set where byte != 0
    return (nz >> 7) * 0x01020408u;   // bits 0, 8, 16, 24 -> 24..27: the 16 partial products land on distinct bits, none on 28..31
}

// bits 28..31 of the result = the solid mask of the four bytes of w (bit 28 + i = byte i non-zero), bits 24..27 zero, the rest junk:
// the flags sit at bits 7, 15, 23, 31 and the multiplier's shifts 21, 14, 7, 0 send them to 28..31; the other twelve partial products
// land on the distinct bits 7, 14, 15, 21, 22, 23 or overflow, so nothing carries into the top byte.  No shift: LOP3, add, LOP3, IMAD.
__device__ __forceinline__ unsigned solid_nibble_hi(unsigned w) {
    const unsigned nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
    return nz * 0x00204081u;
}

// solid ? 0 : maxd for the four bytes of w
__device__ __forceinline__ unsigned seed_word(unsigned w, unsigned maxd4) {
    unsigned nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;  // bit7 set where byte != 0
    unsigned mask = (nz >> 7) * 0xffu;                                   // 0xff where byte != 0
    return ~mask & maxd4;
}

// forward min-plus step through the four bytes of w (low byte first); c = running value
__device__ __forceinline__ unsigned sweep_word_up(unsigned w, unsigned& c) {
    const unsigned b0 = __viaddmin_u32(c, 1u, __byte_perm(w, 0u, 0x4440));
    const unsigned b1 = __viaddmin_u32(b0, 1u, __byte_perm(w, 0u, 0x4441));
    const unsigned b2 = __viaddmin_u32(b1, 1u, __byte_perm(w, 0u, 0x4442));
    const unsigned b3 = __viaddmin_u32(b2, 1u, __byte_perm(w, 0u, 0x4443));
    c = b3;
    return pack_bytes(b0, b1, b2, b3);
}
__device__ __forceinline__ unsigned sweep_word_down(unsigned w, unsigned& c) {
    const unsigned b3 = __viaddmin_u32(c, 1u, __byte_perm(w, 0u, 0x4443));
    const unsigned b2 = __viaddmin_u32(b3, 1u, __byte_perm(w, 0u, 0x4442));
    const unsigned b1 = __viaddmin_u32(b2, 1u, __byte_perm(w, 0u, 0x4441));
    const unsigned b0 = __viaddmin_u32(b1, 1u, __byte_perm(w, 0u, 0x4440));
    c = b0;
    return pack_bytes(b0, b1, b2, b3);
}

// ---- kernel 1: X and Y sweeps of one z-slice in shared memory ---------------------------------
// ManhattanDistanceX.comp:53-68 and ManhattanDistanceY.comp:35-49.
// All slices are resident at once (one wave), so the kernel lasts as long as one CTA's chain of phases: every phase
// has to keep all 12 warps busy.
// X needs no sweep over voxels: the 1-D distance to the nearest solid voxel of a row follows from the row's solid mask.  A quad
// (16 voxels) keeps its 16-bit mask; the carries that enter a quad from the left / right come from a scan over the row's quads
// (2 x 24 steps per row instead of 2 x 384); inside the quad the carries advance word by word through a 16-entry table indexed
// by the word's 4-bit mask (distance inside the word as u16 lanes, carry leaving the word on either side), and every voxel is
// min(inside the word, left carry + offset, right carry + offset): 2 + 4 VIADDMNMX per word against 8 + 8 PRMT for the byte chain,
// and no pass over the slice in shared memory.  Columns are cut into sy segments (4 x 32 rows for the 384 x 128 slice) so the Y sweep
// has 384 independent chains; the segments are joined exactly by carries, like the Z kernels: after the local two-sided sweeps a
// segment's boundary value + distance is what any other segment can see of it; the Y carries are folded into the write-out.
//   stage   16-byte loads (8 per thread, all in flight) -> solid mask per quad, carries leaving the quad
//   xcarry  per row: scan of the quad carries, both directions
//   X       per quad: distances from the mask, the two carries and the nibble table -> shared memory rows (odd stride in quads)
//   Y       local forward + backward sweep per (word column, y-segment)
//   ycarry  per (word column, y-segment): values arriving from above / below
//   store   Y carries applied, 16-byte stores
#ifndef VX_XY_THREADS
#define VX_XY_THREADS 384
#endif
#ifndef VX_XY_OCC
#define VX_XY_OCC 3
#endif
constexpr int XY_THREADS = VX_XY_THREADS;
constexpr int XY_BATCH = 8;   // 16-byte loads in flight per thread
constexpr int XY_MAX_SEG = 8;
constexpr unsigned NO_CARRY = 0x03ffu;  // above every distance, small enough to add offsets in a u16 lane

// EXACT: the slice is a whole number of XY_THREADS * XY_BATCH quads (no guards in the stage).  CNX / CNY / CSY != 0: slice size and
// number of Y segments as compile-time constants (the engine's 384 x 128 slice: every stride and quotient folds into the instructions).
template <bool EXACT, int CNX, int CNY, int CSY>
__global__ void __launch_bounds__(XY_THREADS, VX_XY_OCC) df_xy_slice_kernel(const uint8_t* __restrict__ blocks,
                                                                    uint8_t* __restrict__ df, int nx_arg, int ny_arg,
                                                                    int z_begin, unsigned maxd, int sx, int sy_arg) {
    extern __shared__ uint4 smem4[];
    unsigned* smem = reinterpret_cast<unsigned*>(smem4);
    const int nx = CNX ? CNX : nx_arg, ny = CNY ? CNY : ny_arg, sy = CSY ? CSY : sy_arg;
    const int qpr = nx >> 4;       // 16-byte quads per row
    const int sq = qpr | 1;        // row stride in quads (odd)
    const int sw = sq << 2;        // row stride in words
    const int nxw = nx >> 2;       // words per row
    const int nq = qpr * ny;       // quads in the slice (<= 4096)
    const int rps = ny / sy;       // rows per y-segment
    unsigned* qcar = smem + ny * sw;                                   // [ny][sq]: per quad, carry from the left | from the right << 16
    uint4* lut = reinterpret_cast<uint4*>(qcar + ((ny * sq + 3) & ~3));   // [16]
    uint4* ycar = lut + 16;                                            // [sy][4][qpr]: (fwd.e, fwd.o, bwd.e, bwd.o)
    const unsigned rdiv = ((1u << 20) + qpr - 1) / qpr;  // q / qpr == (q * rdiv) >> 20 for q < 4096, qpr <= 64
    const unsigned ydiv = ((1u << 20) + rps - 1) / rps;  // row / rps likewise (row < 4096)
    const int z = z_begin + blockIdx.x;
    const size_t slice_off = (size_t)z * nx * ny;
    const uint4* src = reinterpret_cast<const uint4*>(blocks + slice_off);
    const unsigned maxd4 = maxd * 0x01010101u;

    // ---- lut: per 4-voxel nibble of the solid mask (bit i = voxel i solid): x, y = distance of every voxel to the nearest solid of
    // the nibble as u16 lanes (v0 | v2 << 16, v1 | v3 << 16; maxd when the nibble is empty), z = carry leaving the word to the right
    // (4 - highest solid), w = carry leaving it to the left (lowest solid + 1), NO_CARRY when empty ----
    if (threadIdx.x < 16) {
        const unsigned n = threadIdx.x;
        unsigned d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned best = maxd;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n >> j & 1u) best = min(best, (unsigned)(i > j ? i - j : j - i));
            d[i] = best;
        }
        const unsigned f = n ? (unsigned)(4 - (31 - __clz(n))) : NO_CARRY, b = n ? (unsigned)__ffs(n) : NO_CARRY;
        lut[n] = make_uint4(d[0] | (d[2] << 16), d[1] | (d[3] << 16), f, b);
    }

    // ---- stage: 16-byte loads (all in flight); per quad the four 4-bit solid masks, kept as byte offsets into the table (mask * 16)
    // in the quad's slot of the slice, and the carries leaving the quad -> qcar ----
    for (int base = 0; base < nq; base += XY_THREADS * XY_BATCH) {
        uint4 v[XY_BATCH];
#pragma unroll
        for (int i = 0; i < XY_BATCH; ++i) {
            const int q = base + i * XY_THREADS + threadIdx.x;
            if (EXACT || q < nq) v[i] = __ldg(src + q);
        }
#pragma unroll
        for (int i = 0; i < XY_BATCH; ++i) {
            const int q = base + i * XY_THREADS + threadIdx.x;
            if (EXACT || q < nq) {
                const int row = (int)(((unsigned)q * rdiv) >> 20), col = q - row * qpr;
                const unsigned nb = __byte_perm(__byte_perm(solid_nibble_top(v[i].x), solid_nibble_top(v[i].y), 0x4473),
                                                __byte_perm(solid_nibble_top(v[i].z), solid_nibble_top(v[i].w), 0x4473), 0x5410);   // one mask per byte
                const unsigned t = nb | (nb >> 4);
                const unsigned m16 = __byte_perm(t, 0u, 0x4420);                      // bit j = voxel j of the quad solid
                smem[(row * sq + col) << 2] = nb << 4;
                // to the right: 16 - highest solid voxel; to the left: lowest solid voxel + 1
                qcar[row * sq + col] = m16 ? (unsigned)(__clz(m16) - 15) | ((unsigned)__ffs(m16) << 16) : NO_CARRY | (NO_CARRY << 16);
            }
        }
    }
    __syncthreads();

    // ---- X carries: per row, what reaches the first voxel of every quad from the left and its last voxel from the right
    // (ManhattanDistanceX.comp:53-68 at quad granularity: 2 * qpr steps per row instead of 2 * nx) ----
    {
        unsigned short* qc = reinterpret_cast<unsigned short*>(qcar);
        for (int item = threadIdx.x; item < 2 * ny; item += XY_THREADS) {
            const int back = item >= ny, row = item - back * ny;
            unsigned short* r = qc + ((row * sq) << 1) + back;
            unsigned c = NO_CARRY;
            if (!back) {
                for (int j = 0; j < qpr; ++j) { const unsigned f = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(f, c + 16u); }
            } else {
                for (int j = qpr - 1; j >= 0; --j) { const unsigned b = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(b, c + 16u); }
            }
        }
    }
    __syncthreads();

    // ---- X distances of every quad from its mask and the two carries: word by word the carries advance through the nibble table,
    // every voxel is min(inside the word, carry from the left + offset, carry from the right + offset) ----
    for (int q = threadIdx.x; q < nq; q += XY_THREADS) {
        const int row = (int)(((unsigned)q * rdiv) >> 20), col = q - row * qpr;
        const int slot = row * sq + col;
        const unsigned off4 = smem[slot << 2], cin = qcar[slot];
        uint4 l[4];
        unsigned cf[4], cb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) l[k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(lut) + ((off4 >> (8 * k)) & 0xffu));
        unsigned c = cin & 0xffffu;
#pragma unroll
        for (int k = 0; k < 4; ++k) { cf[k] = c; c = __viaddmin_u32(c, 4u, l[k].z); }
        c = cin >> 16;
#pragma unroll
        for (int k = 3; k >= 0; --k) { cb[k] = c; c = __viaddmin_u32(c, 4u, l[k].w); }
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned f2 = cf[k] * 0x00010001u, b2 = cb[k] * 0x00010001u;
            unsigned e = __viaddmin_u16x2(f2, 0x00020000u, l[k].x), o = __viaddmin_u16x2(f2, 0x00030001u, l[k].y);
            e = __viaddmin_u16x2(b2, 0x00010003u, e); o = __viaddmin_u16x2(b2, 0x00000002u, o);
            w[k] = pack_lanes(e, o);
        }
        smem4[slot] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();

    // ---- Y: local sweeps per (word column, y-segment), two u16x2 lane pairs per word ----
    for (int item = threadIdx.x; item < nxw * sy; item += XY_THREADS) {
        const int ys = item / nxw, col = item - ys * nxw;
        unsigned* cptr = smem + col;
        const int y0 = ys * rps, y1 = y0 + rps;
        unsigned e = 0x00ff00ffu, o = 0x00ff00ffu;  // min(v, 256) == v for the first row
#pragma unroll 4
        for (int y = y0; y < y1; ++y) {
            const unsigned w = cptr[y * sw];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * sw] = pack_lanes(e, o);
        }
#pragma unroll 4
        for (int y = y1 - 2; y >= y0; --y) {
            const unsigned w = cptr[y * sw];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * sw] = pack_lanes(e, o);
        }
    }
    __syncthreads();

    // ---- ycarry: what reaches the first / last row of (column, ys) from the other y-segments ----
    for (int item = threadIdx.x; item < nxw * sy; item += XY_THREADS) {
        const int ys = item / nxw, col = item - ys * nxw;
        unsigned fe = NO_CARRY * 0x00010001u, fo = fe, be = fe, bo = fe;
        for (int t = 0; t < sy; ++t) {
            if (t < ys) {
                const unsigned w = smem[((t + 1) * rps - 1) * sw + col], d = (unsigned)((ys - t - 1) * rps + 1), d2 = d | (d << 16);
                fe = __viaddmin_u16x2(even_lanes(w), d2, fe); fo = __viaddmin_u16x2(odd_lanes(w), d2, fo);
            } else if (t > ys) {
                const unsigned w = smem[(t * rps) * sw + col], d = (unsigned)((t - ys - 1) * rps + 1), d2 = d | (d << 16);
                be = __viaddmin_u16x2(even_lanes(w), d2, be); bo = __viaddmin_u16x2(odd_lanes(w), d2, bo);
            }
        }
        ycar[(ys * 4 + (col & 3)) * qpr + (col >> 2)] = make_uint4(fe, fo, be, bo);
    }
    __syncthreads();

    // ---- store: Y carries applied on the way out ----
    uint4* dst = reinterpret_cast<uint4*>(df + slice_off);
    for (int q = threadIdx.x; q < nq; q += XY_THREADS) {
        const int row = (int)(((unsigned)q * rdiv) >> 20), colq = q - row * qpr;
        const int ys = (int)(((unsigned)row * ydiv) >> 20);
        const unsigned df_ = (unsigned)(row - ys * rps), db_ = (unsigned)(rps - 1) - df_;
        const unsigned df2 = df_ | (df_ << 16), db2 = db_ | (db_ << 16);
        uint4 v = smem4[row * sq + colq];
        unsigned* vw = reinterpret_cast<unsigned*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 c = ycar[(ys * 4 + k) * qpr + colq];
            unsigned e = even_lanes(vw[k]), o = odd_lanes(vw[k]);
            e = __viaddmin_u16x2(c.x, df2, e); o = __viaddmin_u16x2(c.y, df2, o);
            e = __viaddmin_u16x2(c.z, db2, e); o = __viaddmin_u16x2(c.w, db2, o);
            vw[k] = pack_lanes(e, o);
        }
        dst[q] = v;
    }
}

// Y sweeps of one word column from both ends at once.  The two chains of the reference (down over all rows, then up over all rows) are
// one dependent VIADDMNMX after the other - a warp that does nothing else issues 7.5 instructions per ~20 clocks.  The L1 transform
// does not care in which order the sources reach a voxel, so: phase A sweeps rows 0 -> 63 downwards and rows 127 -> 64 upwards (four
// independent lane chains in flight per thread), phase B continues the downward chain through rows 64 -> 127, which now hold their
// distance to everything below them, and the upward chain through rows 63 -> 0, which hold their distance to everything above:
// every row ends as min over all rows of value + distance, the exact transform, with half the chain length per phase.
// `chunk_done(c)` is called when the 32-row chunk c holds final values (2 and 1 in the middle of phase B, 3 and 0 at its end).
template <class F>
__device__ __forceinline__ void y_sweep_bidir(unsigned* cptr, bool active, F&& chunk_done) {
    unsigned ed = 0x00ff00ffu, od = 0x00ff00ffu, eu = 0x00ff00ffu, ou = 0x00ff00ffu;   // min(v, 256) == v for the first row of a chain
    if (active) {
#pragma unroll 8
        for (int j = 0; j < 128 / 2; ++j) {
            const unsigned wd = cptr[j * 96], wu = cptr[(128 - 1 - j) * 96];
            ed = __viaddmin_u16x2(ed, 0x00010001u, even_lanes_lop(wd));
            od = __viaddmin_u16x2(od, 0x00010001u, odd_lanes(wd));
            eu = __viaddmin_u16x2(eu, 0x00010001u, even_lanes_lop(wu));
            ou = __viaddmin_u16x2(ou, 0x00010001u, odd_lanes(wu));
            cptr[j * 96] = pack_lanes(ed, od);
            cptr[(128 - 1 - j) * 96] = pack_lanes(eu, ou);
        }
    }
    for (int half = 0; half < 2; ++half) {
        if (active) {
#pragma unroll 8
            for (int j = half * 32; j < half * 32 + 32; ++j) {
                const unsigned wd = cptr[(128 / 2 + j) * 96], wu = cptr[(128 / 2 - 1 - j) * 96];
                ed = __viaddmin_u16x2(ed, 0x00010001u, even_lanes_lop(wd));
                od = __viaddmin_u16x2(od, 0x00010001u, odd_lanes(wd));
                eu = __viaddmin_u16x2(eu, 0x00010001u, even_lanes_lop(wu));
                ou = __viaddmin_u16x2(ou, 0x00010001u, odd_lanes(wu));
                cptr[(128 / 2 + j) * 96] = pack_lanes(ed, od);
                cptr[(128 / 2 - 1 - j) * 96] = pack_lanes(eu, ou);
            }
        }
        chunk_done(2 + half);
        chunk_done(1 - half);
    }
}

// the two chains one after the other: down over all rows, then up, a 32-row chunk at a time (chunk_done(3), (2), (1), (0))
template <class F>
__device__ __forceinline__ void y_sweep_seq(unsigned* cptr, bool active, F&& chunk_done) {
    unsigned e = 0x00ff00ffu, o = 0x00ff00ffu;  // min(v, 256) == v for the first row
    if (active) {
#pragma unroll 8
        for (int y = 0; y < 128; ++y) {
            const unsigned w = cptr[y * 96];
            e = __viaddmin_u16x2(e, 0x00010001u, even_lanes_lop(w));
            o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
            cptr[y * 96] = pack_lanes(e, o);
        }
    }
    for (int chunk = 3; chunk >= 0; --chunk) {
        if (active) {
            const int y_hi = chunk == 3 ? 128 - 2 : chunk * 32 + 31;
#pragma unroll 8
            for (int y = y_hi; y >= chunk * 32; --y) {
                const unsigned w = cptr[y * 96];
                e = __viaddmin_u16x2(e, 0x00010001u, even_lanes_lop(w));
                o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w));
                cptr[y * 96] = pack_lanes(e, o);
            }
        }
        chunk_done(chunk);
    }
}

// ---- kernel 1, second version (the engine's 384 x 128 slice) -------------------------------------------------------------
// Same arithmetic as df_xy_slice_kernel, reorganised around what the pipes of an SM can do per clock (tools/debug/pipe_bench.cu:
// VIADDMNMX / VIMNMX3 / PRMT / SHF 64 lanes / clk / SM on the ALU pipe, IMAD / VIADD 64 on the FMA pipe beside it, LDS.128 one warp
// instruction per 4 clocks):
//   * a thread owns 8 consecutive rows of one quad column, loads them with 8 x LDG.128 in flight and keeps the quads' nibble masks in
//     registers from the mask phase to the X phase (no staging pass through shared memory);
//   * X: the inside-the-word distances come from an 8-byte table entry (one conflict-free LDS.64), the carries leaving a word from a
//     4-byte one; "carry + offset" for the four lanes of a word are multiply-adds (FMA pipe) feeding one VIMNMX3 per lane pair;
//   * Y is one unsegmented chain per word column (no segment carries, no fold): 3 warps sweep the 96 columns down and back up through
//     the slice in shared memory (7 instructions per word and direction) while the SM's other CTAs are in their mask / X phases; which
//     3 of the first 4 warps do it rotates per CTA on an SM, so the Y warps of the co-resident CTAs spread over the 4 schedulers;
//   * the result leaves through the bulk-copy engine: as soon as the upward sweep has finished a 32-row chunk, one thread issues
//     cp.async.bulk.global.shared::cta for its 12 KB, so the store overlaps the rest of the sweep and costs no thread instructions.
constexpr int XY2_THREADS = 384;
constexpr int XY2_QPR = 24, XY2_ROWS = 128, XY2_WPR = 96;       // quads / rows / words per slice row
constexpr int XY2_QSTRIDE = 27;                                   // row stride of the quad-carry array (odd, and 8 rows apart never alias a row's 24 banks)
constexpr int XY2_SMEM = XY2_ROWS * XY2_QPR * 16 + XY2_ROWS * XY2_QSTRIDE * 4 + 16 * 8 + 16 * 4 + 16;
__device__ unsigned g_xy2_rotation[1024];   // per SM: CTAs that have started there (only its low two bits are used)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(XY2_THREADS, 3) df_xy2_kernel(const uint8_t* __restrict__ blocks, uint8_t* __restrict__ df, int z_begin, unsigned maxd, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* slice4 = reinterpret_cast<uint4*>(smem_raw);                                      // [128][24] quads, dense
    unsigned* slice = reinterpret_cast<unsigned*>(smem_raw);
    unsigned* qcar = reinterpret_cast<unsigned*>(smem_raw + XY2_ROWS * XY2_QPR * 16);         // [128][27]: carry from the left | from the right << 16
    uint2* lut_eo = reinterpret_cast<uint2*>(qcar + XY2_ROWS * XY2_QSTRIDE);                  // [16]: distances inside the word (v0 | v2 << 16, v1 | v3 << 16)
    unsigned* lut_zw = reinterpret_cast<unsigned*>(lut_eo + 16);                              // [16]: carry leaving to the right | to the left << 16
    int* rot_s = reinterpret_cast<int*>(lut_zw + 16);
    const int tid = threadIdx.x;
    if (dbg & 64) return;   // measurement aid: what a launch of this grid costs by itself 
    const int c0 = tid % XY2_QPR, rseg = tid / XY2_QPR;     // quad column, rows rseg * 8 .. + 7
    const size_t slice_off = (size_t)(z_begin + blockIdx.x) * (XY2_ROWS * XY2_QPR * 16);
    const uint4* src = reinterpret_cast<const uint4*>(blocks + slice_off) + (rseg * 8) * XY2_QPR + c0;

    // Ordering the load bursts of the co-resident CTAs (tickets per SM, each CTA issuing after the one before it) was measured and
    // dropped: 14.9 -> 15.9 us; the memory system does not serve the requests in issue order at this scale and the hand-over costs
    // two block barriers.
    if (tid == 32) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        *rot_s = (int)(atomicAdd(&g_xy2_rotation[smid & 1023u], 1u) & 3u);
    }
    // The CTAs of an SM start their load bursts ~1 us apart (class = position in launch order / SM count, which is how the block scheduler
    // fills the first wave; nothing but speed depends on that guess), so the second CTA's loads run under the first one's X phase instead of
    // beside its loads: 14.79 -> 14.52 us (0.5 / 1.0 / 1.5 / 2.0 us measured: 14.66 / 14.53 / 14.52 / 15.31; dbg >> 8 overrides, in units of
    // 64 ns; dbg & 128 switches it off).  Small, because the phases of co-resident CTAs are not what keeps memory and arithmetic apart.
    {
        const unsigned stagger = (dbg & 128) ? 0u : ((dbg >> 8) ? (unsigned)(dbg >> 8) * 64u : 1024u);
        if (stagger && blockIdx.x >= 148u) __nanosleep((blockIdx.x / 148u) * stagger);
    }
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i * XY2_QPR);

    if (tid < 16) {
        const unsigned n = tid;
        unsigned d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned best = maxd;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n >> j & 1u) best = min(best, (unsigned)(i > j ? i - j : j - i));
            d[i] = best;
        }
        const unsigned f = n ? (unsigned)(4 - (31 - __clz(n))) : NO_CARRY, b = n ? (unsigned)__ffs(n) : NO_CARRY;
        lut_eo[n] = make_uint2(d[0] | (d[2] << 16), d[1] | (d[3] << 16));
        lut_zw[n] = f | (b << 16);
    }

    // ---- masks: per quad the four 4-bit solid masks as byte offsets into the 8-byte table (mask * 8), and the carries leaving the quad ----
    unsigned nb8[8];   // per byte: mask << 4
    if (dbg & 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { nb8[i] = v[i].x ^ v[i].y ^ v[i].z ^ v[i].w; slice4[(rseg * 8 + i) * XY2_QPR + c0] = v[i]; }
    } else
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned nbh = __byte_perm(__byte_perm(solid_nibble_hi(v[i].x), solid_nibble_hi(v[i].y), 0x4473),
                                         __byte_perm(solid_nibble_hi(v[i].z), solid_nibble_hi(v[i].w), 0x4473), 0x5410);   // mask << 4 per byte
        const unsigned t = (nbh >> 4) | (nbh >> 8);
        const unsigned m16 = __byte_perm(t, 0u, 0x4420);                      // bit j = voxel j of the quad solid
        nb8[i] = nbh;
        // to the right: 16 - highest solid voxel; to the left: lowest solid voxel + 1
        qcar[(rseg * 8 + i) * XY2_QSTRIDE + c0] = m16 ? (unsigned)(__clz(m16) - 15) | ((unsigned)__ffs(m16) << 16) : NO_CARRY | (NO_CARRY << 16);
    }
    __syncthreads();

    // ---- X carries per row at quad granularity (ManhattanDistanceX.comp:53-68): what reaches the first voxel of every quad from the
    // left and its last voxel from the right ----
    if (tid < 2 * XY2_ROWS) {
        const int back = tid >= XY2_ROWS, row = tid - back * XY2_ROWS;
        unsigned short* r = reinterpret_cast<unsigned short*>(qcar) + ((row * XY2_QSTRIDE) << 1) + back;
        unsigned c = NO_CARRY;
        if (!back) {
#pragma unroll 8
            for (int j = 0; j < XY2_QPR; ++j) { const unsigned f = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(f, c + 16u); }
        } else {
#pragma unroll 8
            for (int j = XY2_QPR - 1; j >= 0; --j) { const unsigned b = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(b, c + 16u); }
        }
    }
    __syncthreads();

    // ---- X distances: word by word the carries advance through the table; every voxel is
    // min(inside the word, carry from the left + offset, carry from the right + offset) ----
    if (!(dbg & 4)) {
        const unsigned k1 = 0x00010001u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = rseg * 8 + i;
            const unsigned cin = qcar[row * XY2_QSTRIDE + c0];
            uint2 eo[4];
            unsigned zw[4], cf[4], cb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned off = (nb8[i] >> (8 * k + 1)) & 0x78u;
                eo[k] = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(lut_eo) + off);
                zw[k] = *reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(lut_zw) + (off >> 1));
            }
            unsigned c = cin & 0xffffu;
#pragma unroll
            for (int k = 0; k < 4; ++k) { cf[k] = c; c = __viaddmin_u32(c, 4u, zw[k] & 0xffffu); }
            c = cin >> 16;
#pragma unroll
            for (int k = 3; k >= 0; --k) { cb[k] = c; c = __viaddmin_u32(c, 4u, zw[k] >> 16); }
            unsigned w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // lanes (v0, v2) and (v1, v3): from the left + (0, 2) / (1, 3), from the right + (3, 1) / (2, 0)
                const unsigned e = __vimin3_u16x2(eo[k].x, cf[k] * k1 + 0x00020000u, cb[k] * k1 + 0x00010003u);
                const unsigned o = __vimin3_u16x2(eo[k].y, cf[k] * k1 + 0x00030001u, cb[k] * k1 + 0x00000002u);
                w[k] = pack_lanes(e, o);
            }
            slice4[row * XY2_QPR + c0] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncthreads();

    // ---- Y (ManhattanDistanceY.comp:35-49): one chain per word column, down and back up; bulk stores chunk by chunk on the way up ----
    const int warp = tid >> 5;
    const int yw = (warp - *rot_s) & 3;
    if (warp >= 4 || yw == 3) return;
    const int col = yw * 32 + (tid & 31);
    unsigned* cptr = slice + col;
    uint8_t* dst = df + slice_off;
    const bool issuer = yw == 0 && (tid & 31) == 0;
    auto store_chunk = [&](int chunk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the chunk's bytes become visible to the bulk-copy engine
        asm volatile("bar.sync 1, 96;" ::: "memory");
        if (issuer && !(dbg & 8)) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + chunk * 32 * XY2_WPR * 4),
                         "r"(smem_u32(slice + chunk * 32 * XY2_WPR)), "n"(32 * XY2_WPR * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    };
    // measured: the two chains one after the other 14.8 us, both ends at once (y_sweep_bidir, dbg & 256) 15.4 us - the column chains are
    // not bound by the latency of their dependent VIADDMNMX but by the LDS / STS round trips of a warp that walks shared memory alone
    if (dbg & 256) y_sweep_bidir(cptr, !(dbg & 1), store_chunk);
    else y_sweep_seq(cptr, !(dbg & 1), store_chunk);
    if (yw == 0 && (tid & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
}

// ---- kernel 1, third version: one persistent, warp-specialised CTA per SM ----------------------------------------------------------
// df_xy2_kernel is one wave of 384 CTAs: every slice is loaded in one burst at the start (5.9 us until the last slice is complete),
// then computed, then stored, and the measured 14.9 us are close to the SUM of the memory skeleton (8.5 us) and the sweeps (7.4 us).
// Here a CTA walks its slices (blockIdx, + gridDim, ...) through a 3-slot ring in shared memory and the stages of consecutive slices
// overlap:
//   bulk loads   cp.async.bulk.shared::cta.global, 48 KB per slice, complete_tx on full[slot].  The first three are issued one after
//                the other (each after the one before it has landed), so every SM has its first slice after a third of the burst.
//   front group  24 warps: masks (LDS.128 from the raw slot) -> quad-carry scan -> X distances written over the raw bytes of the slot,
//                then arrive on xdone[slot] and on to the next slice
//   Y group      4 warps (one per scheduler, 24 columns each): waits for xdone[slot], sweeps the columns down and up, hands every
//                finished 32-row chunk to a bulk store, and when the stores have read the slot issues the load of slice k + 3 into it
// so the loads of slice k + 1 / k + 2, the masks / X of slice k + 1, the Y sweeps of slice k and the stores of slice k - 1 are in
// flight together.  Same arithmetic as df_xy2_kernel, bit-exact.
// MEASURED (profiles/r2_q_df_xy3_phases.txt): 20.0 us against the 14.9 us of df_xy2_kernel, so it is NOT the default (set_option
// "df_xyver" 3).  Without the Y sweeps it takes 11.5 us, without masks / X 16.3 us, with neither 9.5 us: the Y stage is the bottleneck
// (2.85 us per slice: one warp per scheduler walking a 128-row chain twice, competing for issue slots with six front warps), and at
// 2.6 slices per SM the ring never reaches a steady state - the pipeline's fill and drain cost more than the overlap gains.
constexpr int XY3_FRONT_WARPS = 24, XY3_Y_WARPS = 4;
constexpr int XY3_FRONT = XY3_FRONT_WARPS * 32, XY3_THREADS = XY3_FRONT + XY3_Y_WARPS * 32;
constexpr int XY3_SLOTS = 3;
constexpr int XY3_SLICE = XY2_ROWS * XY2_QPR * 16;                        // 49152 bytes
constexpr int XY3_QCAR = XY2_ROWS * XY2_QSTRIDE * 4;                      // 13824 bytes
constexpr int XY3_SMEM = XY3_SLOTS * XY3_SLICE + 2 * XY3_QCAR + 16 * 8 + 16 * 4 + 2 * XY3_SLOTS * 8;

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(XY3_THREADS, 1) df_xy3_kernel(const uint8_t* __restrict__ blocks, uint8_t* __restrict__ df, int z_begin, int n_slices,
                                                                 unsigned maxd, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* slots = smem_raw;
    unsigned* qcar_base = reinterpret_cast<unsigned*>(smem_raw + XY3_SLOTS * XY3_SLICE);
    uint2* lut_eo = reinterpret_cast<uint2*>(smem_raw + XY3_SLOTS * XY3_SLICE + 2 * XY3_QCAR);
    unsigned* lut_zw = reinterpret_cast<unsigned*>(lut_eo + 16);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(lut_zw + 16);
    unsigned long long* xdone = full + XY3_SLOTS;
    const int tid = threadIdx.x;
    const int n_my = (n_slices - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // slices blockIdx, + gridDim, ...
    if (n_my <= 0) return;

    if (tid == 0) {
        for (int b = 0; b < XY3_SLOTS; ++b) { mbar_init(full + b, 1); mbar_init(xdone + b, XY3_FRONT); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 16) {
        const unsigned n = tid;
        unsigned d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned best = maxd;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n >> j & 1u) best = min(best, (unsigned)(i > j ? i - j : j - i));
            d[i] = best;
        }
        const unsigned f = n ? (unsigned)(4 - (31 - __clz(n))) : NO_CARRY, b = n ? (unsigned)__ffs(n) : NO_CARRY;
        lut_eo[n] = make_uint2(d[0] | (d[2] << 16), d[1] | (d[3] << 16));
        lut_zw[n] = f | (b << 16);
    }
    __syncthreads();
    const size_t slice0 = (size_t)(z_begin + (int)blockIdx.x) * XY3_SLICE, stride = (size_t)gridDim.x * XY3_SLICE;

    if (tid < XY3_FRONT) {
        // ================================ front group: masks, quad carries, X ================================
        const int c0 = tid % XY2_QPR, rseg = tid / XY2_QPR;     // quad column, rows rseg * 4 .. + 3
        for (int k = 0; k < n_my; ++k) {
            const int b = k % XY3_SLOTS;
            uint4* slice4 = reinterpret_cast<uint4*>(slots + b * XY3_SLICE);
            unsigned* qcar = qcar_base + (k & 1) * (XY3_QCAR / 4);
            mbar_wait(full + b, (unsigned)(k / XY3_SLOTS) & 1u);
            if (dbg & 2) { mbar_arrive(xdone + b); continue; }   // measurement aid: no masks / scan / X
            unsigned nb8[4];   // per byte: mask << 4
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = rseg * 4 + i;
                const uint4 v = slice4[row * XY2_QPR + c0];
                const unsigned nbh = __byte_perm(__byte_perm(solid_nibble_hi(v.x), solid_nibble_hi(v.y), 0x4473),
                                                 __byte_perm(solid_nibble_hi(v.z), solid_nibble_hi(v.w), 0x4473), 0x5410);
                const unsigned t = (nbh >> 4) | (nbh >> 8);
                const unsigned m16 = __byte_perm(t, 0u, 0x4420);
                nb8[i] = nbh;
                qcar[row * XY2_QSTRIDE + c0] = m16 ? (unsigned)(__clz(m16) - 15) | ((unsigned)__ffs(m16) << 16) : NO_CARRY | (NO_CARRY << 16);
            }
            asm volatile("bar.sync 2, %0;" ::"n"(XY3_FRONT) : "memory");
            if (tid < 2 * XY2_ROWS) {
                const int back = tid >= XY2_ROWS, row = tid - back * XY2_ROWS;
                unsigned short* r = reinterpret_cast<unsigned short*>(qcar) + ((row * XY2_QSTRIDE) << 1) + back;
                unsigned c = NO_CARRY;
                if (!back) {
#pragma unroll 8
                    for (int j = 0; j < XY2_QPR; ++j) { const unsigned f = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(f, c + 16u); }
                } else {
#pragma unroll 8
                    for (int j = XY2_QPR - 1; j >= 0; --j) { const unsigned bb = r[j << 1]; r[j << 1] = (unsigned short)c; c = min(bb, c + 16u); }
                }
            }
            asm volatile("bar.sync 2, %0;" ::"n"(XY3_FRONT) : "memory");
            const unsigned k1 = 0x00010001u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = rseg * 4 + i;
                const unsigned cin = qcar[row * XY2_QSTRIDE + c0];
                uint2 eo[4];
                unsigned zw[4], cf[4], cb[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const unsigned off = (nb8[i] >> (8 * kk + 1)) & 0x78u;
                    eo[kk] = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(lut_eo) + off);
                    zw[kk] = *reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(lut_zw) + (off >> 1));
                }
                unsigned c = cin & 0xffffu;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { cf[kk] = c; c = __viaddmin_u32(c, 4u, zw[kk] & 0xffffu); }
                c = cin >> 16;
#pragma unroll
                for (int kk = 3; kk >= 0; --kk) { cb[kk] = c; c = __viaddmin_u32(c, 4u, zw[kk] >> 16); }
                unsigned wv[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const unsigned e = __vimin3_u16x2(eo[kk].x, cf[kk] * k1 + 0x00020000u, cb[kk] * k1 + 0x00010003u);
                    const unsigned o = __vimin3_u16x2(eo[kk].y, cf[kk] * k1 + 0x00030001u, cb[kk] * k1 + 0x00000002u);
                    wv[kk] = pack_lanes(e, o);
                }
                slice4[row * XY2_QPR + c0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
            mbar_arrive(xdone + b);   // release: the Y group's wait acquires the X distances of this slot
        }
    } else {
        // ================================ Y group + loads / stores ================================
        const int yw = (tid - XY3_FRONT) >> 5, lane = tid & 31;
        const bool elected = yw == 0 && lane == 0;
        const bool active = lane < 24;
        const int col = yw * 24 + (active ? lane : 0);
        if (elected) {
            // the first loads one after the other: every SM's first slice is complete after a third of the burst
            for (int k = 0; k < n_my && k < XY3_SLOTS; ++k) {
                if (k > 0 && !(dbg & 4)) mbar_wait(full + (k - 1), 0u);
                mbar_expect_tx(full + k, XY3_SLICE);
                bulk_load(slots + k * XY3_SLICE, blocks + slice0 + (size_t)k * stride, XY3_SLICE, full + k);
            }
        }
        for (int k = 0; k < n_my; ++k) {
            const int b = k % XY3_SLOTS;
            unsigned* slice = reinterpret_cast<unsigned*>(slots + b * XY3_SLICE);
            mbar_wait(xdone + b, (unsigned)(k / XY3_SLOTS) & 1u);
            unsigned* cptr = slice + col;
            uint8_t* dst = df + slice0 + (size_t)k * stride;
            y_sweep_seq(cptr, active && !(dbg & 1), [&](int chunk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(XY3_Y_WARPS * 32) : "memory");
                if (elected) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + chunk * 32 * XY2_WPR * 4),
                                 "r"(smem_u32(slice + chunk * 32 * XY2_WPR)), "n"(32 * XY2_WPR * 4)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            });
            if (elected) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stores have read the slot
                if (k + XY3_SLOTS < n_my) {
                    mbar_expect_tx(full + b, XY3_SLICE);
                    bulk_load(slots + b * XY3_SLICE, blocks + slice0 + (size_t)(k + XY3_SLOTS) * stride, XY3_SLICE, full + b);
                }
            }
        }
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

// ---- kernel 2, register-resident variant (plane ranges of exactly ZR_WARPS * SEG planes) ------------------------
// Same decomposition as df_z_tile_kernel, but a lane keeps the SEG words of its column segment in registers: all of
// a lane's loads are in flight at once, the two local sweeps and the carry fold run on registers, and shared memory
// only carries the boundary planes of the 16 segments (8 KB).  Every warp-level load / store moves one 128-byte row
// of a plane.  SEG is a compile-time constant (24 for the 384-plane world, 12 / 6 / 3 for its 2 / 4 / 8 z-slabs), so
// the loops are straight-line code; at <= 42 registers three 512-thread CTAs fit an SM and the 384 tiles of the
// default grid are resident in one wave.
constexpr int ZR_WARPS = 16;
constexpr int ZR_THREADS = ZR_WARPS * 32;

template <int SEG, int CWPP>   // CWPP != 0: words per plane as a compile-time constant (plane offsets become load / store immediates)
__global__ void __launch_bounds__(ZR_THREADS, 3) df_z_reg_kernel(uint8_t* __restrict__ df, int words_per_plane_arg, int z0) {
    const int words_per_plane = CWPP ? CWPP : words_per_plane_arg;
    __shared__ uint4 bnd[ZR_WARPS][32];  // per segment and column: first-plane (e, o), last-plane (e, o) after the local sweeps
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    const bool col_ok = col < words_per_plane;
    unsigned* base = reinterpret_cast<unsigned*>(df) + (size_t)(z0 + s * SEG) * words_per_plane + (col_ok ? col : 0);
    const size_t ps = (size_t)words_per_plane;

    unsigned w[SEG];
    if (CWPP) {
#pragma unroll
        for (int i = 0; i < SEG; ++i) w[i] = base[(size_t)i * CWPP];   // one address register, SEG immediates
    } else {
        const unsigned* p = base;  // running pointer: one live address instead of SEG of them
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
            w[i] = *p;  // a column past the end re-reads column 0 of the tile and is never stored
            p += ps;
        }
    }
    // local forward sweep, then backward (the last plane is final after the forward sweep)
    unsigned e = 0x00ff00ffu, o = 0x00ff00ffu;  // min(v, 256) == v for the first plane
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w[i]));
        o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w[i]));
        w[i] = pack_lanes(e, o);
    }
    const unsigned last_e = e, last_o = o;
#pragma unroll
    for (int i = SEG - 2; i >= 0; --i) {
        e = __viaddmin_u16x2(e, 0x00010001u, even_lanes(w[i]));
        o = __viaddmin_u16x2(o, 0x00010001u, odd_lanes(w[i]));
        w[i] = pack_lanes(e, o);
    }
    bnd[s][lane] = make_uint4(e, o, last_e, last_o);
    __syncthreads();

    // carries of the other segments (see df_z_tile_kernel): fwd arrives at this segment's first plane from the last
    // plane of segment t < s, (s - t - 1) * SEG + 1 planes away; bwd arrives at its last plane likewise
    unsigned fe = 0x03ff03ffu, fo = 0x03ff03ffu, be = 0x03ff03ffu, bo = 0x03ff03ffu;
#pragma unroll
    for (int t = 0; t < ZR_WARPS; ++t) {
        const uint4 v = bnd[t][lane];
        if (t < s) {
            const unsigned d = (unsigned)((s - t - 1) * SEG + 1), d2 = d | (d << 16);
            fe = __viaddmin_u16x2(v.z, d2, fe);
            fo = __viaddmin_u16x2(v.w, d2, fo);
        } else if (t > s) {
            const unsigned d = (unsigned)((t - s - 1) * SEG + 1), d2 = d | (d << 16);
            be = __viaddmin_u16x2(v.x, d2, be);
            bo = __viaddmin_u16x2(v.y, d2, bo);
        }
    }
    // fold: plane i of the segment sees fwd + i and bwd + (SEG - 1 - i); lanes stay below 2^16.
    // The store addresses are re-derived from an opaque copy of the base pointer: left to itself the compiler keeps
    // the SEG load addresses alive across the whole kernel and spills them (ncu: STL/LDL on the critical path).
    asm volatile("" : "+l"(base));
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        unsigned ve = even_lanes(w[i]), vo = odd_lanes(w[i]);
        ve = __viaddmin_u16x2(fe, (unsigned)i * 0x00010001u, ve);
        vo = __viaddmin_u16x2(fo, (unsigned)i * 0x00010001u, vo);
        ve = __viaddmin_u16x2(be, (unsigned)(SEG - 1 - i) * 0x00010001u, ve);
        vo = __viaddmin_u16x2(bo, (unsigned)(SEG - 1 - i) * 0x00010001u, vo);
        if (CWPP) {
            if (col_ok) base[(size_t)i * CWPP] = pack_lanes(ve, vo);
        } else {
            if (col_ok) *base = pack_lanes(ve, vo);
            base += ps;
        }
    }
}

// ---- kernel 2, second register-resident variant: u16 lanes stay unpacked from the load to the store --------------------
// df_z_reg_kernel re-packs the column to bytes after each local sweep and unpacks it again (6 PRMT + 3 IMAD of its 34
// instructions per word).  Here a lane keeps the even / odd u16 lanes of its SEG words in 2 * SEG registers from the load to
// the store: per word 2 PRMT (unpack), 4 VIADDMNMX (local sweeps), 2 VIMNMX3 (fold: min of the value and both carries, whose
// "+ distance" terms are multiply-adds on the FMA pipe, which runs beside the ALU pipe that VIADDMNMX / PRMT occupy) and one
// multiply-add for the pack.
__device__ __forceinline__ unsigned opaque_u32(unsigned v) {
    unsigned r;
    asm("mov.u32 %0, %1;" : "=r"(r) : "r"(v));   // a value the compiler cannot fold: keeps k * i + c a multiply-add (FMA pipe)
    return r;
}

#ifndef VX_ZR2_OCC
#define VX_ZR2_OCC 2
#endif
// FOLD selects how "carry + distance" reaches the 3-input minimum (measured alternatives):
//   0: two VIADDMNMX per lane pair;  1: adds with immediates + VIMNMX3;  2: multiply-adds by the opaque constant k1 = 0x00010001 + VIMNMX3
// COLS = word columns per CTA (8, 16 or 32; a CTA has COLS * 16 threads, 64 registers each): 384 tiles of 32 columns leave a 148-SM
// GPU with 2 or 3 CTAs per SM (the slower SMs set the time); 1536 tiles of 8 columns (one 32-byte sector per plane row) balance to 6 %.
template <int SEG, int CWPP, int FOLD, int COLS>
__global__ void __launch_bounds__(COLS * ZR_WARPS, 1024 / (COLS * ZR_WARPS)) df_z_reg2_kernel(uint8_t* __restrict__ df, int words_per_plane_arg, int z0, unsigned k1) {
    const int words_per_plane = CWPP ? CWPP : words_per_plane_arg;
    if (k1 == 1) return;   // measurement aid: what a launch of this grid costs by itself
    __shared__ uint2 bnd_last[ZR_WARPS][COLS];   // per segment and column: last plane after the local sweeps, + 1 per lane (e, o)
    __shared__ uint2 bnd_first[ZR_WARPS][COLS];  // first plane likewise
    const int s = threadIdx.x / COLS, lane = threadIdx.x % COLS;
    const int col = blockIdx.x * COLS + lane;
    const bool col_ok = col < words_per_plane;
    unsigned* base = reinterpret_cast<unsigned*>(df) + (size_t)(z0 + s * SEG) * words_per_plane + (col_ok ? col : 0);
    const size_t ps = (size_t)words_per_plane;

    unsigned e[SEG], o[SEG];
    {
        unsigned w[SEG];
        if (CWPP) {
#pragma unroll
            for (int i = 0; i < SEG; ++i) w[i] = base[(size_t)i * CWPP];
        } else {
            const unsigned* p = base;
#pragma unroll
            for (int i = 0; i < SEG; ++i) { w[i] = *p; p += ps; }
        }
#pragma unroll
        for (int i = 0; i < SEG; ++i) { e[i] = even_lanes_lop(w[i]); o[i] = odd_lanes(w[i]); }
    }
    if (k1 == 0) {   // measurement aid: load + unpack + pack + store only
        asm volatile("" : "+l"(base));
#pragma unroll
        for (int i = 0; i < SEG; ++i)
            if (col_ok) base[(size_t)i * CWPP] = pack_lanes(e[i], o[i]);
        return;
    }
    // local forward sweep, then backward
#pragma unroll
    for (int i = 1; i < SEG; ++i) {
        e[i] = __viaddmin_u16x2(e[i - 1], 0x00010001u, e[i]);
        o[i] = __viaddmin_u16x2(o[i - 1], 0x00010001u, o[i]);
    }
    bnd_last[s][lane] = make_uint2(e[SEG - 1] + 0x00010001u, o[SEG - 1] + 0x00010001u);
#pragma unroll
    for (int i = SEG - 2; i >= 0; --i) {
        e[i] = __viaddmin_u16x2(e[i + 1], 0x00010001u, e[i]);
        o[i] = __viaddmin_u16x2(o[i + 1], 0x00010001u, o[i]);
    }
    bnd_first[s][lane] = make_uint2(e[0] + 0x00010001u, o[0] + 0x00010001u);
    __syncthreads();

    // carries of the other segments as min-plus chains over the segments (warp-uniform trip counts):
    //   fwd arriving at this segment's first plane = min over t < s of last_t + 1 + (s - 1 - t) * SEG,  bwd likewise from t > s
    constexpr unsigned SEG2 = (unsigned)SEG * 0x00010001u;
    unsigned fe = 0x03ff03ffu, fo = 0x03ff03ffu, be = 0x03ff03ffu, bo = 0x03ff03ffu;   // "no carry": above every distance
#pragma unroll 1
    for (int t = 0; t < s; ++t) {
        const uint2 v = bnd_last[t][lane];
        fe = __viaddmin_u16x2(fe, SEG2, v.x);
        fo = __viaddmin_u16x2(fo, SEG2, v.y);
    }
#pragma unroll 1
    for (int t = ZR_WARPS - 1; t > s; --t) {
        const uint2 v = bnd_first[t][lane];
        be = __viaddmin_u16x2(be, SEG2, v.x);
        bo = __viaddmin_u16x2(bo, SEG2, v.y);
    }
    // fold + pack + store: plane i sees fwd + i and bwd + (SEG - 1 - i); the sums stay below 2^16 per lane
    asm volatile("" : "+l"(base));
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        unsigned ve, vo;
        if (FOLD == 0) {
            ve = __viaddmin_u16x2(fe, (unsigned)i * 0x00010001u, e[i]);
            vo = __viaddmin_u16x2(fo, (unsigned)i * 0x00010001u, o[i]);
            ve = __viaddmin_u16x2(be, (unsigned)(SEG - 1 - i) * 0x00010001u, ve);
            vo = __viaddmin_u16x2(bo, (unsigned)(SEG - 1 - i) * 0x00010001u, vo);
        } else if (FOLD == 1) {
            ve = __vimin3_u16x2(e[i], fe + (unsigned)i * 0x00010001u, be + (unsigned)(SEG - 1 - i) * 0x00010001u);
            vo = __vimin3_u16x2(o[i], fo + (unsigned)i * 0x00010001u, bo + (unsigned)(SEG - 1 - i) * 0x00010001u);
        } else {
            ve = __vimin3_u16x2(e[i], k1 * (unsigned)i + fe, k1 * (unsigned)(SEG - 1 - i) + be);
            vo = __vimin3_u16x2(o[i], k1 * (unsigned)i + fo, k1 * (unsigned)(SEG - 1 - i) + bo);
        }
        if (CWPP) {
            if (col_ok) base[(size_t)i * CWPP] = pack_lanes(ve, vo);
        } else {
            if (col_ok) *base = pack_lanes(ve, vo);
            base += ps;
        }
    }
}

// ---- kernel 2, persistent variant: one CTA per SM loops over the column tiles and loads the words of its NEXT tile into registers
// before it works on the current one, so the L2 round trip of a tile overlaps the sweeps of the one before it (in df_z_reg2_kernel a
// CTA loads, computes and stores strictly one after the other, and with 2 or 3 tiles per SM nothing else fills the gaps).
template <int SEG, int CWPP>
__global__ void __launch_bounds__(ZR_THREADS, 1) df_z_persist_kernel(uint8_t* __restrict__ df, int z0, int ntiles) {
    __shared__ uint2 bnd_last[2][ZR_WARPS][32];   // double-buffered by tile parity: one barrier per tile is enough
    __shared__ uint2 bnd_first[2][ZR_WARPS][32];
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* plane0 = reinterpret_cast<unsigned*>(df) + (size_t)(z0 + s * SEG) * CWPP + lane;
    constexpr unsigned SEG2 = (unsigned)SEG * 0x00010001u;

    unsigned wn[SEG];
    int tile = blockIdx.x;
    if (tile < ntiles) {
#pragma unroll
        for (int i = 0; i < SEG; ++i) wn[i] = plane0[(size_t)i * CWPP + tile * 32];
    }
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        unsigned e[SEG], o[SEG];
#pragma unroll
        for (int i = 0; i < SEG; ++i) { e[i] = even_lanes_lop(wn[i]); o[i] = odd_lanes(wn[i]); }
        const int next = tile + gridDim.x;
        if (next < ntiles) {
#pragma unroll
            for (int i = 0; i < SEG; ++i) wn[i] = plane0[(size_t)i * CWPP + next * 32];
        }
#pragma unroll
        for (int i = 1; i < SEG; ++i) {
            e[i] = __viaddmin_u16x2(e[i - 1], 0x00010001u, e[i]);
            o[i] = __viaddmin_u16x2(o[i - 1], 0x00010001u, o[i]);
        }
        const int par = it & 1;
        bnd_last[par][s][lane] = make_uint2(e[SEG - 1] + 0x00010001u, o[SEG - 1] + 0x00010001u);
#pragma unroll
        for (int i = SEG - 2; i >= 0; --i) {
            e[i] = __viaddmin_u16x2(e[i + 1], 0x00010001u, e[i]);
            o[i] = __viaddmin_u16x2(o[i + 1], 0x00010001u, o[i]);
        }
        bnd_first[par][s][lane] = make_uint2(e[0] + 0x00010001u, o[0] + 0x00010001u);
        __syncthreads();   // the buffers of this parity are next written two tiles later, after the barrier of the tile in between
        unsigned fe = 0x03ff03ffu, fo = 0x03ff03ffu, be = 0x03ff03ffu, bo = 0x03ff03ffu;
#pragma unroll 1
        for (int t = 0; t < s; ++t) {
            const uint2 v = bnd_last[par][t][lane];
            fe = __viaddmin_u16x2(fe, SEG2, v.x);
            fo = __viaddmin_u16x2(fo, SEG2, v.y);
        }
#pragma unroll 1
        for (int t = ZR_WARPS - 1; t > s; --t) {
            const uint2 v = bnd_first[par][t][lane];
            be = __viaddmin_u16x2(be, SEG2, v.x);
            bo = __viaddmin_u16x2(bo, SEG2, v.y);
        }
        unsigned* out = plane0 + tile * 32;
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
            unsigned ve = __viaddmin_u16x2(fe, (unsigned)i * 0x00010001u, e[i]);
            unsigned vo = __viaddmin_u16x2(fo, (unsigned)i * 0x00010001u, o[i]);
            ve = __viaddmin_u16x2(be, (unsigned)(SEG - 1 - i) * 0x00010001u, ve);
            vo = __viaddmin_u16x2(bo, (unsigned)(SEG - 1 - i) * 0x00010001u, vo);
            out[(size_t)i * CWPP] = pack_lanes(ve, vo);
        }
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
        VX_CUDA(cudaFuncSetAttribute(df_xy_slice_kernel<true, 384, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VX_CUDA(cudaFuncSetAttribute(df_xy_slice_kernel<true, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VX_CUDA(cudaFuncSetAttribute(df_xy_slice_kernel<false, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VX_CUDA(cudaFuncSetAttribute(df_z_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VX_CUDA(cudaFuncSetAttribute(df_xy2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XY2_SMEM));
        VX_CUDA(cudaFuncSetAttribute(df_xy3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XY3_SMEM));
        attr_set = true;
    }
    return VXRT_OK;
}

// X + Y sweeps of slices [z0, z1), then the Z sweeps of the same plane range
static int launch_df_range(vxrt_ctx* c, int z0, int z1) {
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    const unsigned maxd = (unsigned)((nx + ny + nz) < 254 ? (nx + ny + nz) : 254);
    // segments per row / per column: as many as keep XY_THREADS chains busy, dividing the row / column evenly
    const int qpr = nx >> 4, nxw = nx >> 2;
    int sx = XY_THREADS / ny, sy = XY_THREADS / nxw;
    sx = sx < 1 ? 1 : (sx > XY_MAX_SEG ? XY_MAX_SEG : sx);
    sy = sy < 1 ? 1 : (sy > XY_MAX_SEG ? XY_MAX_SEG : sy);
    if (c->df_sx > 0) sx = c->df_sx;
    if (c->df_sy > 0) sy = c->df_sy;
    while (qpr % sx) --sx;
    while (ny % sy) --sy;
    const size_t smem_xy = (size_t)ny * (qpr | 1) * sizeof(uint4) + (((size_t)ny * (qpr | 1) + 3) & ~(size_t)3) * sizeof(unsigned) +
                           16 * sizeof(uint4) + (size_t)sy * 4 * qpr * sizeof(uint4);
    int rc = set_smem_attrs();
    if (rc) return rc;
    if (c->df_stage != 2) {
        if (c->df_xyver == 3 && nx == 384 && ny == 128)
            df_xy3_kernel<<<(z1 - z0) < c->sm_count ? (z1 - z0) : c->sm_count, XY3_THREADS, XY3_SMEM, c->stream>>>(c->d_blocks, c->d_df, z0, z1 - z0, maxd, c->df_dbg);
        else if (c->df_xyver == 2 && nx == 384 && ny == 128)
            df_xy2_kernel<<<z1 - z0, XY2_THREADS, XY2_SMEM, c->stream>>>(c->d_blocks, c->d_df, z0, maxd, c->df_dbg);
        else if (nx == 384 && ny == 128 && sy == 4 && XY_THREADS * XY_BATCH == 3072)   // the engine's slice (WORLD_SIZE_X x WORLD_SIZE_Y, Macros.h)
            df_xy_slice_kernel<true, 384, 128, 4><<<z1 - z0, XY_THREADS, smem_xy, c->stream>>>(c->d_blocks, c->d_df, nx, ny, z0, maxd, sx, sy);
        else if ((qpr * ny) % (XY_THREADS * XY_BATCH) == 0)
            df_xy_slice_kernel<true, 0, 0, 0><<<z1 - z0, XY_THREADS, smem_xy, c->stream>>>(c->d_blocks, c->d_df, nx, ny, z0, maxd, sx, sy);
        else
            df_xy_slice_kernel<false, 0, 0, 0><<<z1 - z0, XY_THREADS, smem_xy, c->stream>>>(c->d_blocks, c->d_df, nx, ny, z0, maxd, sx, sy);
    }
    VX_CUDA(cudaGetLastError());
    if (c->df_stage == 1) return VXRT_OK;
    const int wpp = (nx * ny) >> 2, nzr = z1 - z0;
    const int ztiles = (wpp + 31) / 32;
    const bool engine_grid = nzr == ZR_WARPS * 24 && wpp == 12288;
#define Z2(FOLD, COLS) df_z_reg2_kernel<24, 12288, FOLD, COLS><<<wpp / COLS, COLS * ZR_WARPS, 0, c->stream>>>(c->d_df, wpp, z0, (c->df_dbg & 64) ? 1u : (c->df_dbg & 32) ? 0u : 0x00010001u)
    if (engine_grid && c->df_zver == 2) Z2(0, 32);
    else if (engine_grid && c->df_zver == 3) Z2(1, 32);
    else if (engine_grid && c->df_zver == 4) Z2(2, 32);
    else if (engine_grid && c->df_zver == 5) Z2(0, 16);
    else if (engine_grid && c->df_zver == 6) Z2(2, 16);
    else if (engine_grid && c->df_zver == 7) Z2(0, 8);
    else if (engine_grid && c->df_zver == 8) Z2(2, 8);
    else if (engine_grid && c->df_zver == 9) df_z_persist_kernel<24, 12288><<<c->sm_count < 384 ? c->sm_count : 384, ZR_THREADS, 0, c->stream>>>(c->d_df, z0, 384);
#undef Z2
    else if (nzr == ZR_WARPS * 24 && wpp == 12288) df_z_reg_kernel<24, 12288><<<ztiles, ZR_THREADS, 0, c->stream>>>(c->d_df, wpp, z0);   // the engine's 384 x 128 x 384 grid
    else if (nzr == ZR_WARPS * 24) df_z_reg_kernel<24, 0><<<ztiles, ZR_THREADS, 0, c->stream>>>(c->d_df, wpp, z0);
    else if (nzr == ZR_WARPS * 12) df_z_reg_kernel<12, 0><<<ztiles, ZR_THREADS, 0, c->stream>>>(c->d_df, wpp, z0);
    else if (nzr == ZR_WARPS * 6) df_z_reg_kernel<6, 0><<<ztiles, ZR_THREADS, 0, c->stream>>>(c->d_df, wpp, z0);
    else if (nzr == ZR_WARPS * 3) df_z_reg_kernel<3, 0><<<ztiles, ZR_THREADS, 0, c->stream>>>(c->d_df, wpp, z0);
    else {
        const int seg = (nzr + Z_SEGS - 1) / Z_SEGS;
        const size_t smem_z = ((size_t)nzr * Z_TILE_QUADS + (size_t)Z_SEGS * Z_TILE_WORDS) * sizeof(uint4);
        df_z_tile_kernel<<<(wpp + Z_TILE_WORDS - 1) / Z_TILE_WORDS, Z_THREADS, smem_z, c->stream>>>(c->d_df, wpp, z0, z1, seg);
    }
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
