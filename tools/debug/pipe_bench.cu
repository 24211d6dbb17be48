// Pipe-throughput microbenchmark for the integer instructions of the distance-field kernels (sm_100a):
// which of VIADDMNMX.U16x2 / VIMNMX3.U16x2 / PRMT / LOP3 / VIADD / IMAD share an issue pipe.  Each kernel runs CH independent
// dependent chains per thread; "mix" kernels interleave two opcodes.  Prints thread-instructions per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu ; check the loop bodies with cuobjdump -sass
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 8, ITERS = 2048;
#define CHAINS(body)                                   \
    unsigned x[CH];                                    \
    for (int j = 0; j < CH; ++j) x[j] = in[threadIdx.x + j]; \
    unsigned a = in[100], b = in[101];                 \
    for (int it = 0; it < ITERS; ++it) {               \
        _Pragma("unroll") for (int j = 0; j < CH; ++j) { body; } \
    }                                                  \
    unsigned r = 0;                                    \
    for (int j = 0; j < CH; ++j) r ^= x[j];            \
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;

__global__ void k_viaddmnmx(const unsigned* in, unsigned* out) { CHAINS(x[j] = __viaddmin_u16x2(x[j], a, b)) }
__global__ void k_vimnmx3(const unsigned* in, unsigned* out) { CHAINS(x[j] = __vimin3_u16x2(x[j], a, b) ^ 0) }
__global__ void k_prmt(const unsigned* in, unsigned* out) { CHAINS(x[j] = __byte_perm(x[j], a, 0x4341)) }
__global__ void k_lop3(const unsigned* in, unsigned* out) { CHAINS(x[j] = (x[j] & a) ^ b) }
__global__ void k_imad(const unsigned* in, unsigned* out) { CHAINS(x[j] = x[j] * a + b) }
__global__ void k_add(const unsigned* in, unsigned* out) { CHAINS(x[j] = x[j] + x[j ^ 1]) }
__global__ void k_lop_addimm(const unsigned* in, unsigned* out) { CHAINS(x[j] = (x[j] + 0x00010001u) ^ a) }
__global__ void k_lop_imad(const unsigned* in, unsigned* out) { CHAINS(x[j] = (x[j] * a + b) ^ a) }
__global__ void k_lop_lop(const unsigned* in, unsigned* out) { CHAINS(x[j] = ((x[j] & b) | 0x00010001u) ^ a; x[j] = (x[j] | b) & (a + j)) }
__global__ void k_shf(const unsigned* in, unsigned* out) { CHAINS(x[j] = __funnelshift_r(x[j], a, 7)) }
__global__ void k_mix_vmn_imad(const unsigned* in, unsigned* out) { CHAINS(if (j & 1) x[j] = x[j] * a + b; else x[j] = __viaddmin_u16x2(x[j], a, b)) }
__global__ void k_mix_vmn_add(const unsigned* in, unsigned* out) { CHAINS(x[j] = __viaddmin_u16x2(x[j] + 0x00010001u, a, b)) }
__global__ void k_mix_vmn_prmt(const unsigned* in, unsigned* out) { CHAINS(if (j & 1) x[j] = __byte_perm(x[j], a, 0x4341); else x[j] = __viaddmin_u16x2(x[j], a, b)) }
__global__ void k_mix_vmn_lop(const unsigned* in, unsigned* out) { CHAINS(if (j & 1) x[j] = (x[j] & a) ^ b; else x[j] = __viaddmin_u16x2(x[j], a, b)) }
__global__ void k_mix_imad_prmt(const unsigned* in, unsigned* out) { CHAINS(if (j & 1) x[j] = __byte_perm(x[j], a, 0x4341); else x[j] = x[j] * a + b) }
// shared-memory load rates: LDS.32 / LDS.64 / LDS.128, conflict-free
__global__ void k_lds32(const unsigned* in, unsigned* out) {
    __shared__ unsigned s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = in[i & 255];
    __syncthreads();
    unsigned r = 0, idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) r += s[(idx + j * 256) & 4095];
        idx += r & 1 ? 0 : 32;   // keeps the addresses data dependent
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_lds128(const unsigned* in, unsigned* out) {
    __shared__ uint4 s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = make_uint4(in[i & 255], 1, 2, 3);
    __syncthreads();
    unsigned r = 0, idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { const uint4 v = s[(idx + j * 64) & 1023]; r += v.x ^ v.y ^ v.z ^ v.w; }
        idx += r & 1 ? 0 : 32;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
void run(const char* name, K k, const unsigned* in, unsigned* out, int sms, double per_iter) {
    const int threads = 512, blocks = sms * 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<blocks, threads>>>(in, out);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<<<blocks, threads>>>(in, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double instr = (double)blocks * threads * ITERS * per_iter;
    const double clocks = best * 1e-3 * clk_khz * 1e3;
    printf("%-18s %8.3f ms  %7.1f thread-instr/clk/SM (at the %d MHz attribute clock)\n", name, best, instr / clocks / sms, clk_khz / 1000);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *in, *out;
    cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 4 * 512 * 2 * 200);
    unsigned h[4096]; for (int i = 0; i < 4096; ++i) h[i] = i * 2654435761u;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run("VIADDMNMX.U16x2", k_viaddmnmx, in, out, sms, CH);
    run("VIMNMX3.U16x2", k_vimnmx3, in, out, sms, CH);
    run("PRMT", k_prmt, in, out, sms, CH);
    run("LOP3", k_lop3, in, out, sms, CH);
    run("IMAD", k_imad, in, out, sms, CH);
    run("add reg", k_add, in, out, sms, CH);
    run("LOP3+add imm", k_lop_addimm, in, out, sms, 2 * CH);
    run("LOP3+IMAD", k_lop_imad, in, out, sms, 2 * CH);
    run("LOP3+LOP3", k_lop_lop, in, out, sms, 2 * CH);
    run("SHF", k_shf, in, out, sms, CH);
    run("mix VMNMX+IMAD", k_mix_vmn_imad, in, out, sms, CH);
    run("VMNMX(add imm)", k_mix_vmn_add, in, out, sms, 2 * CH);
    run("mix VMNMX+PRMT", k_mix_vmn_prmt, in, out, sms, CH);
    run("mix VMNMX+LOP3", k_mix_vmn_lop, in, out, sms, CH);
    run("mix IMAD+PRMT", k_mix_imad_prmt, in, out, sms, CH);
    run("LDS.32 x8", k_lds32, in, out, sms, CH);
    run("LDS.128 x8", k_lds128, in, out, sms, CH);
    return 0;
}
