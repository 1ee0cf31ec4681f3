// ubcache.cuh — feature/ubcache on the device: per-user behaviour sequences in
// time-descending order (TimeSeq, cache.go:9-12) stored as one CSR in HBM, and TimeSeq.Filter
// (cache.go:71-94) evaluated for a whole batch: the first `count` items with ts <= maxTs.  Emits the
// [B,S] history-row matrix the train / predict entry points consume (most recent first, -1 padded:
// prepare.go:49-51, rcmd.go:517-522), removing the host from the per-sample path.
#pragma once
#include "common.cuh"

namespace ctr {

// First index in [beg, end) whose timestamp is <= mt (timestamps descend), or end.  Warp-cooperative: a few
// binary steps (uniform across the warp) narrow the range to 32 entries, then ONE coalesced load + ballot finishes —
// 3 dependent memory round trips for a 100-event history instead of 7.
__device__ __forceinline__ long long ub_first_leq(const long long* __restrict__ ts, long long beg, long long end, long long mt, int lane) {
    long long lo = beg, hi = end;                        // invariant: answer in [lo, hi]
    while (hi - lo > 32) { const long long mid = (lo + hi) >> 1; if (ts[mid] <= mt) hi = mid; else lo = mid + 1; }
    const long long i = lo + lane;
    const bool hit = i < hi && ts[i] <= mt;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    return m ? lo + (__ffs(m) - 1) : hi;
}

// one warp per sample: lower bound over the (descending) timestamps, then a coalesced
// write of up to S item rows
__global__ void __launch_bounds__(256)
k_ub_window(const long long* __restrict__ off, const long long* __restrict__ ts, const int* __restrict__ items,
            const int* __restrict__ user_row, const long long* __restrict__ max_ts, int B, int S, long n_users,
            int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += nwarps) {
        const int u = user_row[b];
        long long beg = 0, end = 0;
        if (u >= 0 && u < n_users) { beg = off[u]; end = off[u + 1]; }      // unknown user → empty (cache.go:61-64)
        long long first = end;
        if (end > beg) {
            long long mt = max_ts[b];
            if (mt == 0) mt = ts[beg];                                      // cache.go:72-74
            first = ub_first_leq(ts, beg, end, mt, lane);      // every lane of the warp takes this branch (u, beg, end are warp-uniform)
        }
        const long long avail = end - first;
        for (int s = lane; s < S; s += 32) hist[(long)b * S + s] = s < avail ? items[first + s] : -1;
    }
}

}  // namespace ctr
