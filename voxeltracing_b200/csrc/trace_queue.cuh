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
inline int trace_queue_grid(size_t n) { return (int)((n + VX_TRACE_CTA - 1) / VX_TRACE_CTA); }

template <bool STATS, class Policy>
__device__ __forceinline__ void trace_queue(const GridView& g, Policy& pol, int count, int max_iter, LaneStats* st) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    f3 o, d;
    if (tid < count && pol.fetch(tid, o, d)) pol.store(tid, traverse_df<STATS>(g, o, d, max_iter, st));
}
