// trace_queue.cuh — the trace kernels of the wavefront pipelines walk a ray queue through a policy:
//     bool fetch(int idx, f3& o, f3& d)   false when queue entry idx carries no ray
//     void store(int idx, const TraceResult& r)
// One queue entry per thread.  A persistent-warp variant (every warp owning a chunk of the queue and
// refilling lanes whose ray had terminated, with the finish / fetch work batched over >= 8 parked lanes) was
// measured on config 4 and lost: GI 1.33 -> 1.52 ms, reflections 0.76 -> 0.84 ms at the best geometry
// (chunk 128, refill 8; larger chunks were worse).  The queues are latency bound, not lane bound: the
// refill rounds expose the queue loads once per round instead of once per warp and the resumable ray state
// costs 20 more registers (profiles/r1_e_wavefront_sweep.txt).  So did in-CTA compaction (live rays packed into the
// lowest threads through shared memory every 4..16 iterations, emptied warps exiting): GI 1.12 -> 1.32..1.56 ms
// although a replay of the oracle's iteration counts predicted 40 % fewer warp-iterations — 60 registers instead of
// 40 and one block barrier per round, at which seven warps wait for the slowest (profiles/r1_p_cta_compaction_sweep.txt).
#pragma once
#include "traverse.cuh"

constexpr int VX_TRACE_CTA = 128;
// Register allocation of the queue trace kernels.  With a bare __launch_bounds__(128) ptxas stops at 40 registers and pays for it inside
// the loop (a spill / fill pair and the grid dimensions reloaded from the constant bank every iteration: 70 instructions per DDA
// iteration).  Asking for ONE resident CTA lifts that target: 46 - 47 registers, no spill, 65 instructions per DDA iteration and 33
// instead of 38 per skip iteration, 10 instead of 12 CTAs per SM.  Measured (config 4, profiles/r2_x_tocc_sweep.txt): GI 1.024 -> 1.006 ms;
// asking for 10 / 11 / 12 CTAs (45 / 40 / 40 registers, the nudge constants then rebuilt per iteration) 1.022 / 1.030 / 1.031 ms.  Three
// more instructions cut from the DDA path (k == 0 test moved to the skip side) changed nothing: the loop is bound by the latency of
// its dependent chain at this occupancy as much as by issue slots.  0 = unspecified.
#ifndef VX_TRACE_OCC
#define VX_TRACE_OCC 1
#endif
#if VX_TRACE_OCC > 0
#define VX_TRACE_BOUNDS __launch_bounds__(VX_TRACE_CTA, VX_TRACE_OCC)
#else
#define VX_TRACE_BOUNDS __launch_bounds__(VX_TRACE_CTA)
#endif
inline int trace_queue_grid(size_t n) { return (int)((n + VX_TRACE_CTA - 1) / VX_TRACE_CTA); }

template <bool STATS, class Policy>
__device__ __forceinline__ void trace_queue(const GridView& g, Policy& pol, int count, int max_iter, LaneStats* st) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    f3 o, d;
    if (tid < count && pol.fetch(tid, o, d)) pol.store(tid, traverse_df<STATS>(g, o, d, max_iter, st));
}

// ---- iteration-capped passes with compaction BETWEEN launches -------------------------------------------------------------------
// A warp of trace_queue runs until its slowest ray ends; the bounce rays of a warp end after 5 ... 47 iterations, so 13 of 32 lanes
// execute the average instruction (profiles/r1_s_ncu_full_summary.txt).  Both in-kernel remedies above paid for the compaction inside
// the kernel.  Here the queue is walked in passes: pass j runs every ray it is given for the iterations [cap[j-1], cap[j]) of ITS loop;
// a ray that ends inside the pass is finished exactly as before, a survivor is appended (warp-aggregated, 16 bytes: current position +
// queue index + loop state) to a continuation queue, and the next launch starts from that dense queue.  The sequence of iterations of a
// ray, its arithmetic and therefore its result are unchanged bit for bit; only which thread of which launch executes an iteration differs.
// tools/analysis/lane_replay.py replays the oracle's per-ray iteration counts: caps (12, 24, max) leave 62 - 68 % of the warp-iterations
// of the uncapped kernel on the rooms world, (6, 12, 24, max) 60 - 66 %; sorting the queue by direction octant and coarse origin cell
// instead would leave 88 - 93 %.
// MEASURED (config 4, 1080p, profiles/r2_k_sweep_caps.txt): bit-identical for every schedule, and not faster - GI 1.047 ms uncapped,
// 1.044 (16, 32), 1.064 (12), 1.077 (12, 24), 1.244 (6, 12, 24); reflections 0.639 / 0.649 / 0.639 / 0.655 / 0.717.  What the replay
// does not see: a warp-iteration costs ~78 instructions, not ~36, because lanes in a DDA step and lanes in a skip step of the same
// iteration execute both blocks (that divergence survives any compaction of finished rays), every pass repeats the ray set-up and three
// dependent loads (continuation -> queue entry -> ray) in front of a loop that is now short, and each extra launch adds a tail.  Kept as
// an option (set_option "trace_caps"), off by default; tests/test_gpu_shade.py holds every schedule to bit identity.
template <bool STATS, class Policy>
__global__ void __launch_bounds__(VX_TRACE_CTA) trace_capped_kernel(GridView g, Policy pol, const int* __restrict__ count_ptr, int n_fixed, int it_begin,
                                                                    int it_end, int max_iter, const float4* __restrict__ cin, float4* __restrict__ cout,
                                                                    int* __restrict__ cout_count, TraceStatsDev* stats) {
    const int count = count_ptr ? *count_ptr : n_fixed;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    LaneStats ls = {0u, 0u, 0u, 0u};
    bool survivor = false;
    f3 cur = F3(0.0f);
    int idx = tid, state = 0;
    if (tid < count) {
        f3 o, d;
        bool have;
        if (cin) {
            const float4 e = cin[tid];
            const unsigned bits = __float_as_uint(e.w);
            idx = (int)(bits & 0x0fffffffu);
            state = (int)(bits >> 28);
            cur = F3(e.x, e.y, e.z);
            have = pol.fetch(idx, o, d);   // o = where the ray started (t is measured from there)
        } else {
            have = pol.fetch(idx, o, d);
            cur = o;
        }
        if (have) {
            const RaySetup rs = ray_setup(g, d);
            bool ended = false;
            for (int itr = it_begin; itr < it_end; ++itr) {
                const int c = df_iteration<STATS>(g, rs, cur, state, &ls);
                if (c == VX_ITER_CONTINUE) continue;
                if (c == VX_ITER_TAIL) {
                    bool Intersection = (state & 4) != 0;
                    int MinIdx = state & 3;
                    run_tail<STATS>(g, cur, d, itr, max_iter, Intersection, MinIdx, &ls);
                    state = MinIdx | (Intersection ? 4 : 0);
                }
                ended = true;
                break;
            }
            if (ended || it_end >= max_iter) pol.store(idx, trace_result<STATS>(g, rs, cur, o, state, &ls));
            else survivor = true;
        }
    }
    if (cout) {
        const unsigned lane = threadIdx.x & 31u;
        const unsigned m = __ballot_sync(0xffffffffu, survivor);
        if (m) {
            int base = 0;
            if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cout_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (survivor) cout[base + __popc(m & ((1u << lane) - 1u))] = make_float4(cur.x, cur.y, cur.z, __uint_as_float((unsigned)idx | ((unsigned)state << 28)));
        }
    }
    if (STATS) flush_stats(stats, ls);
}

// ---- adaptive hand-over: a warp gives up its stragglers ---------------------------------------------------------------------------
// The capped passes above move EVERY survivor at a fixed iteration (55 % of the rays at cap 16) and pay set-up + three dependent loads for
// each of them again.  Here the decision is the warp's own: it runs until at most T of its lanes still have a ray, then appends those
// (20 bytes: position, queue index | loop state, iterations done) to the continuation queue and exits; the next launch packs the
// stragglers of 32 / T warps into one.  tools/analysis/lane_replay.py (rooms world): T = 8 moves 23 % of the rays and leaves 70 - 75 % of
// the uncapped kernel's warp-iterations, (12, 8) moves 35 % + 8 % and leaves 60 - 67 %.  The iteration sequence of a ray is unchanged, so
// the results are bit-identical; the order of the continuation queue is not deterministic, which nothing downstream sees (every ray
// writes to its own queue index).  set_option "trace_spill" = thresholds as bytes, low byte first.
// MEASURED (config 4, 1080p, profiles/r2_u_spill_sweep.txt, ncu per launch in gpurun_out/r2_v_*): bit-identical for every schedule, and
// again not faster - GI 1.026 ms plain, 1.072 at the best setting (T = 12, votes every 4 iterations), reflections 0.640 -> 0.664.  At
// T = 8 the first pass runs with 20.2 lanes per instruction instead of 13.7 and takes 183 us instead of 210, but the vote + resumable
// state cost 60 registers instead of 40 (45 % instead of 64 % of the warps resident) and it retires only 20 % fewer warp instructions
// (160 M against 199 M) where the replay predicted 30 %; the continuation pass then runs 25 M more at 10 lanes and a 58 % L1 hit rate
// (its rays come from all over the frame), 46 us.  Off by default; tests/test_gpu_shade.py holds six schedules to bit identity.
#ifndef VX_SPILL_OCC
#define VX_SPILL_OCC 1
#endif
#ifndef VX_SPILL_CHUNK
#define VX_SPILL_CHUNK 2
#endif
template <bool STATS, bool SPILL, class Policy>
__global__ void __launch_bounds__(VX_TRACE_CTA, SPILL ? VX_SPILL_OCC : 1) trace_spill_kernel(GridView g, Policy pol, const int* __restrict__ count_ptr, int n_fixed, int threshold, int max_iter,
                                                                   const float4* __restrict__ cin, const unsigned* __restrict__ meta_in, float4* __restrict__ cout,
                                                                   unsigned* __restrict__ mout, int* __restrict__ cout_count, TraceStatsDev* stats) {
    const int count = count_ptr ? *count_ptr : n_fixed;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid - (int)(threadIdx.x & 31u) >= count) return;   // whole warp beyond the queue
    LaneStats ls = {0u, 0u, 0u, 0u};
    f3 cur = F3(0.0f), o = F3(0.0f), d = F3(1.0f);
    int idx = tid, state = 0, itr = 0;
    bool running = false;
    if (tid < count) {
        if (cin) {
            const float4 e = cin[tid];
            const unsigned meta = meta_in[tid];
            idx = __float_as_int(e.w);
            state = (int)(meta & 7u);
            itr = (int)(meta >> 3);
            cur = F3(e.x, e.y, e.z);
            running = pol.fetch(idx, o, d);   // o = where the ray started (t is measured from there)
        } else {
            running = pol.fetch(idx, o, d);
            cur = o;
        }
    }
    const bool have = running;
    const RaySetup rs = ray_setup(g, d);
    running = running && itr < max_iter;
    if (SPILL) {
        for (;;) {
            const unsigned m = __ballot_sync(0xffffffffu, running);
            if (__popc(m) <= threshold) break;
            if (running) {
                // VX_SPILL_CHUNK iterations between two votes
                const int stop = min(itr + VX_SPILL_CHUNK, max_iter);
#pragma unroll 1
                for (; itr < stop; ++itr) {
                    const int c = df_iteration<STATS>(g, rs, cur, state, &ls);
                    if (c == VX_ITER_CONTINUE) continue;
                    if (c == VX_ITER_TAIL) {
                        bool Intersection = (state & 4) != 0;
                        int MinIdx = state & 3;
                        run_tail<STATS>(g, cur, d, itr, max_iter, Intersection, MinIdx, &ls);
                        state = MinIdx | (Intersection ? 4 : 0);
                    }
                    running = false;
                    break;
                }
                if (itr >= max_iter) running = false;
            }
        }
    } else {
        if (running) {
            for (; itr < max_iter; ++itr) {
                const int c = df_iteration<STATS>(g, rs, cur, state, &ls);
                if (c == VX_ITER_CONTINUE) continue;
                if (c == VX_ITER_TAIL) {
                    bool Intersection = (state & 4) != 0;
                    int MinIdx = state & 3;
                    run_tail<STATS>(g, cur, d, itr, max_iter, Intersection, MinIdx, &ls);
                    state = MinIdx | (Intersection ? 4 : 0);
                }
                break;
            }
            running = false;
        }
    }
    if (have && !running) pol.store(idx, trace_result<STATS>(g, rs, cur, o, state, &ls));
    if (SPILL) {
        const unsigned lane = threadIdx.x & 31u;
        const unsigned m = __ballot_sync(0xffffffffu, running);
        if (m) {
            int base = 0;
            if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cout_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (running) {
                const int at = base + __popc(m & ((1u << lane) - 1u));
                cout[at] = make_float4(cur.x, cur.y, cur.z, __int_as_float(idx));
                mout[at] = (unsigned)state | ((unsigned)itr << 3);
            }
        }
    }
    if (STATS) flush_stats(stats, ls);
}

template <class Policy>
int launch_trace_spill(vxrt_ctx* c, const GridView& g, const Policy& pol, const int* count_ptr, size_t n, int max_iter, TraceStatsDev* stats) {
    int th[3], nth = 0;
    for (int j = 0; j < 3; ++j) {
        const int k = (c->trace_spill >> (8 * j)) & 0xff;
        if (k > 0 && k < 32) th[nth++] = k; else break;
    }
    TraceCont tc = {{nullptr, nullptr}, {nullptr, nullptr}, nullptr};
    if (nth) {
        const int rc = vxrt_ensure_trace_cont(c, n, &tc);
        if (rc != VXRT_OK) return rc;
        VX_CUDA(cudaMemsetAsync(tc.count, 0, 4 * sizeof(int), c->stream));
    }
    const bool st = c->stats_on;
    size_t bound = n;   // upper bound of the entries of pass j: every warp of pass j - 1 hands over at most th[j - 1] rays
    for (int j = 0; j <= nth; ++j) {
        const float4* cin = j ? tc.q[(j - 1) & 1] : nullptr;
        const unsigned* meta_in = j ? tc.meta[(j - 1) & 1] : nullptr;
        float4* cout = j < nth ? tc.q[j & 1] : nullptr;
        unsigned* mout = j < nth ? tc.meta[j & 1] : nullptr;
        int* cc = j < nth ? tc.count + j : nullptr;
        const int* cnt = j ? tc.count + (j - 1) : count_ptr;
        const int grid = trace_queue_grid(bound);
        if (j < nth) {
            if (st) trace_spill_kernel<true, true, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, th[j], max_iter, cin, meta_in, cout, mout, cc, stats);
            else trace_spill_kernel<false, true, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, th[j], max_iter, cin, meta_in, cout, mout, cc, stats);
            bound = ((bound + 31) / 32) * (size_t)th[j];
        } else {
            if (st) trace_spill_kernel<true, false, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, 0, max_iter, cin, meta_in, cout, mout, cc, stats);
            else trace_spill_kernel<false, false, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, 0, max_iter, cin, meta_in, cout, mout, cc, stats);
        }
    }
    c->launches += nth + 1;
    return VXRT_OK;
}


// Walks a queue of `n` entries (or *count_ptr of them) through the passes c->trace_caps describes.  Queue indices must fit 28 bits.
template <class Policy>
int launch_trace_capped(vxrt_ctx* c, const GridView& g, const Policy& pol, const int* count_ptr, size_t n, int max_iter, TraceStatsDev* stats) {
    if (c->trace_spill) return launch_trace_spill(c, g, pol, count_ptr, n, max_iter, stats);
    int caps[4], ncaps = 0;
    for (int j = 0; j < 3; ++j) {
        const int k = (c->trace_caps >> (8 * j)) & 0xff;
        if (k > 0 && k < max_iter && (ncaps == 0 || k > caps[ncaps - 1])) caps[ncaps++] = k;
    }
    caps[ncaps++] = max_iter;
    TraceCont tc = {{nullptr, nullptr}, {nullptr, nullptr}, nullptr};
    if (ncaps > 1) {
        if (n >= (1u << 28)) return vxrt_fail(VXRT_E_INVALID, "trace queue of %zu rays exceeds the 28-bit continuation index", n);
        const int rc = vxrt_ensure_trace_cont(c, n, &tc);
        if (rc != VXRT_OK) return rc;
        VX_CUDA(cudaMemsetAsync(tc.count, 0, 4 * sizeof(int), c->stream));
    }
    const int grid = trace_queue_grid(n);
    const bool st = c->stats_on;
    int begin = 0;
    for (int j = 0; j < ncaps; ++j) {
        const float4* cin = j ? tc.q[(j - 1) & 1] : nullptr;
        float4* cout = j + 1 < ncaps ? tc.q[j & 1] : nullptr;
        int* cc = j + 1 < ncaps ? tc.count + j : nullptr;
        const int* cnt = j ? tc.count + (j - 1) : count_ptr;
        if (st) trace_capped_kernel<true, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, begin, caps[j], max_iter, cin, cout, cc, stats);
        else trace_capped_kernel<false, Policy><<<grid, VX_TRACE_CTA, 0, c->stream>>>(g, pol, cnt, (int)n, begin, caps[j], max_iter, cin, cout, cc, stats);
        begin = caps[j];
    }
    c->launches += ncaps;
    return VXRT_OK;
}
