// idmap.cuh — sparse external id → dense table row, on the device.
// The reference keys every cache by the decimal string of a Go int (`strconv.Itoa(sampleKey.UserId)`,
// rcmd.go:472,483,502; `itemEmbeddingMap.Get`, :502,519) — MovieLens ids are sparse, the tables here are
// dense.  An open-addressing table (linear probing, load factor <= 0.5) in HBM maps int64 ids to int32
// rows; a missing id yields -1, which the gather turns into a zero row exactly like the reference's
// "embedding not found → zeros" (rcmd.go:501-505,519-521).
#pragma once
#include "common.cuh"
#include "ubcache.cuh"

namespace ctr {

constexpr unsigned long long kIdEmpty = 0x8000000000000000ull;   // INT64_MIN is not a usable id

__device__ __forceinline__ unsigned long long idmap_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void __launch_bounds__(256)
k_idmap_clear(unsigned long long* __restrict__ keys, int* __restrict__ vals, unsigned long long cap) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cap; i += (unsigned long long)gridDim.x * blockDim.x) {
        keys[i] = kIdEmpty; vals[i] = -1;
    }
}

// row of ids[i] is i.  flags[0] |= 1 on a duplicate id, |= 2 on the reserved id.
__global__ void __launch_bounds__(256)
k_idmap_insert(unsigned long long* __restrict__ keys, int* __restrict__ vals, unsigned long long mask,
               const long long* __restrict__ ids, long n, int* __restrict__ flags) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const unsigned long long id = (unsigned long long)ids[i];
        if (id == kIdEmpty) { atomicOr(flags, 2); continue; }
        unsigned long long slot = idmap_hash(id) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(&keys[slot], kIdEmpty, id);
            if (prev == kIdEmpty) { vals[slot] = (int)i; break; }
            if (prev == id) { atomicOr(flags, 1); atomicMin(&vals[slot], (int)i); break; }
            slot = (slot + 1) & mask;
        }
    }
}

__global__ void __launch_bounds__(256)
k_idmap_lookup(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, unsigned long long mask,
               const long long* __restrict__ ids, long n, int* __restrict__ rows) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const unsigned long long id = (unsigned long long)ids[i];
        int row = -1;
        if (id != kIdEmpty) {
            unsigned long long slot = idmap_hash(id) & mask;
            for (;;) {
                const unsigned long long k = __ldg(&keys[slot]);
                if (k == id) { row = __ldg(&vals[slot]); break; }
                if (k == kIdEmpty) break;
                slot = (slot + 1) & mask;
            }
        }
        rows[i] = row;
    }
}

__device__ __forceinline__ int idmap_find(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                          unsigned long long mask, unsigned long long id) {
    if (id == kIdEmpty) return -1;
    unsigned long long slot = idmap_hash(id) & mask;
    for (;;) {
        const unsigned long long k = __ldg(&keys[slot]);
        if (k == id) return __ldg(&vals[slot]);
        if (k == kIdEmpty) return -1;
        slot = (slot + 1) & mask;
    }
}

// The whole key → index step of recommend.BatchPredict (rcmd.go:277-337) in one launch, one warp per key:
// lanes 0/1 resolve the user / item id, the warp then evaluates TimeSeq.Filter (cache.go:71-94) for the
// user's behaviour CSR (ub_off == nullptr: no UserBehavior provider → empty history) and writes hist [B,S].
// BatchPredict's per-sample failure rule (rcmd.go:299-307): a key whose user or item feature cannot be
// fetched becomes an all-zero X row — user, item and every history slot are dropped together.
// keys3 = [user ids | item ids | timestamps], each B long.
__global__ void __launch_bounds__(256)
k_keys_resolve(const unsigned long long* __restrict__ ukeys, const int* __restrict__ uvals, unsigned long long umask,
               const unsigned long long* __restrict__ ikeys, const int* __restrict__ ivals, unsigned long long imask,
               const long long* __restrict__ ub_off, const long long* __restrict__ ub_ts, const int* __restrict__ ub_items, long n_users,
               const long long* __restrict__ keys3, int B, int S,
               int* __restrict__ user_row, int* __restrict__ item_row, int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += nwarps) {
        int row = -1;
        if (lane == 0) row = idmap_find(ukeys, uvals, umask, (unsigned long long)keys3[b]);
        else if (lane == 1) row = idmap_find(ikeys, ivals, imask, (unsigned long long)keys3[B + b]);
        int u = __shfl_sync(0xffffffffu, row, 0), it = __shfl_sync(0xffffffffu, row, 1);
        if (u < 0 || it < 0) { u = -1; it = -1; }
        long long first = 0, end = 0;
        if (ub_off && u >= 0 && u < n_users) {
            const long long beg = ub_off[u]; end = ub_off[u + 1]; first = end;
            if (end > beg) {
                long long mt = keys3[2 * (long)B + b];
                if (mt == 0) mt = ub_ts[beg];
                first = ub_first_leq(ub_ts, beg, end, mt, lane);
            }
        }
        const long long avail = end - first;
        for (int s = lane; s < S; s += 32) hist[(long)b * S + s] = s < avail ? ub_items[first + s] : -1;
        if (lane == 0) { user_row[b] = u; item_row[b] = it; }
    }
}

// ---- recommend.GetSample's assembly for training (rcmd.go:339-460), on the device -----------------------------
// A sample whose user or item has no features is skipped by the reference's assembler goroutines
// (rcmd.go:378-382: `continue` on GetSampleVector's error); here: flag → exclusive scan → ordered compaction.
__global__ void __launch_bounds__(256)
k_keys_rows(const unsigned long long* __restrict__ ukeys, const int* __restrict__ uvals, unsigned long long umask,
            const unsigned long long* __restrict__ ikeys, const int* __restrict__ ivals, unsigned long long imask,
            const long long* __restrict__ uid, const long long* __restrict__ iid, long n, int* __restrict__ flag) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        flag[i] = (idmap_find(ukeys, uvals, umask, (unsigned long long)uid[i]) >= 0 &&
                   idmap_find(ikeys, ivals, imask, (unsigned long long)iid[i]) >= 0) ? 1 : 0;
}
// survivors are appended after the *count samples of the earlier chunks, input order kept
__global__ void __launch_bounds__(256)
k_keys_compact(const unsigned long long* __restrict__ ukeys, const int* __restrict__ uvals, unsigned long long umask,
               const unsigned long long* __restrict__ ikeys, const int* __restrict__ ivals, unsigned long long imask,
               const long long* __restrict__ uid, const long long* __restrict__ iid, const long long* __restrict__ ts,
               const float* __restrict__ label, long n, const int* __restrict__ flag, const int* __restrict__ pos,
               const unsigned long long* __restrict__ count, int* __restrict__ out_user, int* __restrict__ out_item,
               long long* __restrict__ out_ts, float* __restrict__ out_label) {
    const unsigned long long base = *count;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const unsigned long long o = base + (unsigned long long)pos[i];
        out_user[o] = idmap_find(ukeys, uvals, umask, (unsigned long long)uid[i]);
        out_item[o] = idmap_find(ikeys, ivals, imask, (unsigned long long)iid[i]);
        out_ts[o] = ts[i]; out_label[o] = label[i];
    }
}
__global__ void k_keys_advance(const int* __restrict__ chunk_total, unsigned long long* __restrict__ count) { *count += (unsigned long long)*chunk_total; }

}  // namespace ctr
