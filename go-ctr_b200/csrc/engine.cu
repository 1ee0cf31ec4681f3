// engine.cu — libctr_b200.so: handle, host-side mirror of model.Train / model.Predict
// (model/model.go:27-353 of auxten/go-ctr) and the C ABI of include/ctr_b200.h.
// sm_100a only; no CPU fallback: every compute entry point launches CUDA kernels or fails.
#include "../../include/ctr_b200.h"
#include "attn.cuh"
#include "mlp.cuh"
#include "umma_gemm.cuh"
#include "auc.cuh"
#include "ubcache.cuh"
#include "idmap.cuh"
#include "item2vec.cuh"
#include "mlp64.cuh"
#include "comm.cuh"

#include <nvtx3/nvToolsExt.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace ctr;

namespace {

std::string g_create_error;

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct Prof { double ms = 0; long n = 0; };

// Caller buffers are pageable (Go slices, numpy arrays): a cudaMemcpyAsync straight from them is staged
// synchronously by the driver and kills the copy/compute overlap.  The epoch entry points therefore copy each batch
// into a pinned ring first — with a few worker threads, a single memcpy stream tops out near 10 GB/s — and DMA from
// there.  The library never keeps a caller pointer past the call (cgo rule).
class CopyPool {
public:
    explicit CopyPool(int n) { for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); }); }
    ~CopyPool() { { std::lock_guard<std::mutex> l(m_); stop_ = true; } cv_.notify_all(); for (auto& t : th_) t.join(); }
    // copies n bytes, split over the workers and the calling thread; returns when done
    void copy(void* dst, const void* src, size_t n) {
        const size_t kMin = (size_t)512 << 10;
        const int parts = (int)std::min<size_t>(th_.size() + 1, std::max<size_t>(1, n / kMin));
        if (parts <= 1) { memcpy(dst, src, n); return; }
        const size_t chunk = (n / parts + 63) & ~(size_t)63;
        { std::lock_guard<std::mutex> l(m_);
          for (int i = 1; i < parts; i++) {
              const size_t off = (size_t)i * chunk; if (off >= n) break;
              jobs_.push_back({(char*)dst + off, (const char*)src + off, std::min(chunk, n - off)}); pending_++;
          } }
        cv_.notify_all();
        memcpy(dst, src, std::min(chunk, n));
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }
private:
    struct Job { char* d; const char* s; size_t n; };
    void run() {
        for (;;) {
            Job j;
            { std::unique_lock<std::mutex> l(m_);
              cv_.wait(l, [this] { return stop_ || !jobs_.empty(); });
              if (stop_ && jobs_.empty()) return;
              j = jobs_.back(); jobs_.pop_back(); }
            memcpy(j.d, j.s, j.n);
            { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    std::vector<std::thread> th_; std::vector<Job> jobs_;
    std::mutex m_; std::condition_variable cv_, done_; int pending_ = 0; bool stop_ = false;
};

}  // namespace

struct ctr_handle {
    ctr_config cfg{};
    int dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    mutable std::string err;
    std::mutex mu;                        // Predict may be hit concurrently (gin handlers, api.go:106-131)

    int in = 0, Kp = 0, H0p = 0, H1p = 0, Sp = 0, lddx = 0, Bmax = 0;
    int num_sms = 148;

    // tables (HBM resident)
    float* tab[3] = {nullptr, nullptr, nullptr};
    long tab_ld[3] = {0, 0, 0};
    int64_t tab_rows[3] = {0, 0, 0};      // logical (global) rows
    int64_t tab_local_rows[3] = {0, 0, 0};
    int tab_width[3] = {0, 0, 0};
    bool tab_sharded[3] = {false, false, false};   // world > 1: rows r % world == rank live here, the rest on the peers
    VmmBuf tab_vmm[3];                    // a sharded table lives in shareable (VMM) memory; tab[] points into it
    uint64_t tab_gen = 1;                 // bumped whenever an ITEM_* table is (re)allocated: peers must re-map it

    // learnables, padded storage: W0 [Kp,H0p] W1 [H0p,H1p] W2 [H1p] att [Sp]; grads / Adam moments alike
    float *W[4] = {}, *G[4] = {}, *Mo[4] = {}, *Vo[4] = {};
    float* Gflat = nullptr; size_t Gflat_n = 0;    // the four gradient tensors live in one buffer (one all-reduce, one memset)
    size_t wsize[4] = {};
    uint32_t step = 0;                    // optimiser steps taken (Adam t = step+1; dropout stream = step*4+layer)

    // activations
    float *X0 = nullptr, *H0d = nullptr, *H1d = nullptr, *P = nullptr, *Z = nullptr;
    float *dZ1 = nullptr, *dZ0 = nullptr, *dX = nullptr;
    float *dUb = nullptr, *dIt = nullptr;         // row-gradient buffers (deterministic / debug)
    size_t dUb_cap = 0;
    unsigned *keys = nullptr, *keys2 = nullptr, *pos = nullptr, *pos2 = nullptr;
    void* sort_tmp = nullptr; size_t sort_tmp_bytes = 0; size_t keys_cap = 0;
    double* d_cost = nullptr;
    float* hot_acc = nullptr; int hot_rows = 0, hot_reps = 0;
    float *emb_m = nullptr, *emb_v = nullptr;      // CTR_TABLE_ADAM: first / second moments of ITEM_EMB, same layout as the table
    // staging for host-pointer entry points
    int *s_user = nullptr, *s_item = nullptr, *s_hist = nullptr; float* s_label = nullptr;
    // epoch entry points (ctr_train_idx / ctr_train_keys): pinned host ring (filled by the copy pool from the caller's
    // pageable buffers) → device slots on a copy stream, overlapping the compute of the previous batch
    static constexpr int kPin = 3;
    unsigned char* feed_pin[kPin] = {}; unsigned char* feed_dev[2] = {}; size_t feed_bytes = 0;
    cudaEvent_t pin_free[kPin] = {};
    CopyPool* pool = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    double* d_costs = nullptr; size_t d_costs_cap = 0;
    // ctr_train_keys: resolved + compacted samples resident in HBM
    int *kt_user = nullptr, *kt_item = nullptr; long long* kt_ts = nullptr; float* kt_label = nullptr; size_t kt_cap = 0;
    int *kt_flag = nullptr, *kt_pos = nullptr; void* kt_scan_tmp = nullptr; size_t kt_scan_bytes = 0; size_t kt_chunk = 0;
    unsigned long long* kt_count = nullptr;
    static constexpr int kCnt = 8;                       // pinned ring of survivor counts read back chunk by chunk
    unsigned long long* kt_count_host = nullptr; cudaEvent_t kt_counted[kCnt] = {};
    // history windows of batch b+1 are cut on their own stream while batch b trains (two buffers)
    cudaStream_t win_stream = nullptr; int* kt_hist[2] = {}; cudaEvent_t kt_win[2] = {}, kt_used[2] = {};
    // dense-X residency
    float* dXd = nullptr; float* dYd = nullptr; size_t dXd_cap = 0, dYd_cap = 0;

    int64_t launches = 0;
    bool profiling = false;
    std::map<std::string, Prof> prof;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // tcgen05 GEMM engine (umma_gemm.cuh): pre-split K-major weight operands + TMA descriptors
    struct Umma {
        bool ready = false;
        bool dirty = true;                 // weights changed since the last split
        int stages_fwd0 = 0, stages_fwd1 = 0, stages_dz0 = 0, stages_dx = 0;
        int kbk_fwd0 = 32, kbk_fwd1 = 32, kbk_dz0 = 32, kbk_dx = 32;       // K elements per shared-memory stage of each GEMM
        int bn_fwd0 = 0, bn_fwd1 = 0, bn_dx = 0;
        float *Wt0[2] = {}, *Wt1[2] = {}, *W1s[2] = {}, *W0s[2] = {};     // [hi, lo]
        CUtensorMap mA_X0, mA_H0d, mA_dZ1, mA_dZ0;
        CUtensorMap mB_Wt0[2], mB_Wt1[2], mB_W1s[2], mB_W0s[2];
        // transposed accumulation (k_umma_gemm<.., TR>) for the GEMMs whose weight tile is <= 128 rows: 256-row activation boxes
        CUtensorMap mA_dZ0_tr, mB_W0s_tr[2];
        int stages_dx_tr = 0;
        CUtensorMap mK_X0, mK_H0d, mK_dZ0, mK_dZ1;      // {32 x ks} boxes over [batch, width] for the weight-gradient GEMMs
        CUtensorMap m3_X0, m3_H0d, m3_dZ0, m3_dZ1;      // the same operands as ONE {32, ks, width/32} box per k-block
        int dw_tma = 0;                                 // producer mode of k_umma_dw (DwArgs.tma)
        int dw_stages0 = 0, dw_stages1 = 0;
        int dw_ks = 16, dw_terms = 3;                   // samples per k-block; products per term (3 = error-compensated, fp32-grade)
    } um;

    // device-side ubcache (ubcache.cuh)
    long long *ub_off = nullptr, *ub_ts = nullptr; int* ub_items = nullptr; int64_t ub_users = 0, ub_n = 0;

    // device id maps (idmap.cuh): [CTR_IDMAP_USER, CTR_IDMAP_ITEM]
    unsigned long long* idm_keys[2] = {nullptr, nullptr}; int* idm_vals[2] = {nullptr, nullptr};
    unsigned long long idm_cap[2] = {0, 0}; int64_t idm_n[2] = {0, 0};
    long long *k_keys = nullptr, *k_host = nullptr; int* k_flags = nullptr;   // staging for ctr_batch_predict_keys (device, pinned host)

    // grow-only device scratch of the host-buffer entry points (gather / window / id lookup): no cudaMalloc on the serving path
    void* scratch[3] = {nullptr, nullptr, nullptr}; size_t scratch_cap[3] = {0, 0, 0};

    unsigned long long* umma_dbg = nullptr;     // CTR_UMMA_TIMELINE=1: timeline buffer of the last umma launch

    Comm comm;
};

namespace {

int set_err(const ctr_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return set_err(h, e_ == cudaErrorMemoryAllocation ? CTR_ENOMEM : CTR_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

#define RET(x) do { int r_ = (x); if (r_ != CTR_OK) return r_; } while (0)

// every kernel launch goes through here: counts it and, when profiling, brackets it with events
// CTR_NVTX=1: every launch sits in an NVTX range named like the profile key (nsys / ncu --nvtx timelines)
const bool g_nvtx = getenv("CTR_NVTX") != nullptr;

template <typename F>
int launch(ctr_handle* h, const char* name, F&& f) {
    if (h->profiling) cudaEventRecord(h->ev0, h->stream);
    if (g_nvtx) nvtxRangePushA(name);
    f();
    if (g_nvtx) nvtxRangePop();
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(h, CTR_ECUDA, "launch %s: %s", name, cudaGetErrorString(e));
    if (h->profiling) {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        Prof& p = h->prof[name]; p.ms += ms; p.n++;
    }
    return CTR_OK;
}

template <typename T>
int dalloc(ctr_handle* h, T** p, size_t n, bool zero = true) {
    CU(h, cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    if (zero) CU(h, cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    return CTR_OK;
}

// device scratch slot `i` with at least `bytes` (contents undefined); grows geometrically, freed with the handle
int scratch_get(ctr_handle* h, int i, size_t bytes, void** out) {
    if (h->scratch_cap[i] < bytes) {
        if (h->scratch[i]) { cudaStreamSynchronize(h->stream); cudaFree(h->scratch[i]); h->scratch[i] = nullptr; h->scratch_cap[i] = 0; }
        const size_t cap = std::max(bytes, h->scratch_cap[i] * 2);
        CU(h, cudaMalloc(&h->scratch[i], cap));
        h->scratch_cap[i] = cap;
    }
    *out = h->scratch[i];
    return CTR_OK;
}

Dims dims_of(const ctr_handle* h) { return Dims{h->cfg.uP, h->cfg.S, h->cfg.D, h->cfg.cF, h->in}; }

int grid_for_warps(const ctr_handle* h, int B) {
    int blocks = (B + 7) / 8;
    return std::max(1, std::min(blocks, h->num_sms * 8));
}

bool vec_ok(const ctr_handle* h, const RowSrc& r) {
    int D = h->cfg.D;
    if (D % 4) return false;
    int lpr = D / 4;
    if (lpr < 4 || lpr > 32 || (lpr & (lpr - 1))) return false;      // D in {16,32,64,128}; others use the generic kernels
    if (r.dense) return (r.ldx % 4 == 0) && (r.ub0 % 4 == 0) && (r.it0 % 4 == 0) && (((uintptr_t)r.X) % 16 == 0);
    return r.lde % 4 == 0;
}

// Lane mapping of the vector attention kernels for a row of D floats: LPR lanes x VPL float4 each, one
// load group in flight per warp, 128-thread blocks at >= 6-8 blocks/SM (tests/cuda/gather_probe.cu,
// bwd_probe.cu: occupancy beats deeper per-warp unrolling on B200).
int grid_attn(const ctr_handle* h, int B) {
    int blocks = (B + 3) / 4;
    return std::max(1, std::min(blocks, h->num_sms * 16));     // two waves of 8 resident blocks; 8..64 per SM measure within 1 %
}
template <int LPR, int VPL, int MINB>
void launch_fwd_vec(ctr_handle* h, const RowSrc& r, int B) {
    size_t smem = (size_t)4 * h->Kp * sizeof(float);
    const int g = grid_attn(h, B);
    switch (h->cfg.model) {
        case CTR_MODEL_YOUTUBE: k_attn_fwd_vec<LPR, VPL, 1, MODEL_YOUTUBE, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
        case CTR_MODEL_DIN_COS: k_attn_fwd_vec<LPR, VPL, 1, MODEL_DIN_COS, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
        default:                k_attn_fwd_vec<LPR, VPL, 1, MODEL_DIN_EUC, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
    }
}
template <int LPR, int VPL, int MINB>
void launch_bwd_vec(ctr_handle* h, const RowSrc& r, const BwdOut& o, int B) {
    size_t smem = (size_t)std::max(h->cfg.S, 1) * sizeof(float);
    const int g = grid_attn(h, B);
    switch (h->cfg.model) {
        case CTR_MODEL_YOUTUBE: k_attn_bwd_vec<LPR, VPL, 1, MODEL_YOUTUBE, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->dX, h->lddx, o, B); break;
        case CTR_MODEL_DIN_COS: k_attn_bwd_vec<LPR, VPL, 1, MODEL_DIN_COS, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->dX, h->lddx, o, B); break;
        default:                k_attn_bwd_vec<LPR, VPL, 1, MODEL_DIN_EUC, 128, MINB><<<g, 128, smem, h->stream>>>(r, dims_of(h), h->W[3], h->dX, h->lddx, o, B); break;
    }
}

// index-mode fast kernels (k_attn_fwd_idx / k_attn_bwd_idx): conditions beyond vec_ok
bool idx_fwd_ok(const ctr_handle* h, const RowSrc& r, int B) {
    static const bool off = getenv("CTR_ATTN_OLD") != nullptr;
    const ctr_config& c = h->cfg;
    return !off && !r.dense && c.S <= 64 && c.uP % 4 == 0 && r.ldu % 4 == 0 && r.ldi % 4 == 0 && (h->Kp - 2 * c.D) / 4 <= 32 &&
           r.ufeat && r.ifeat;
}
bool idx_bwd_ok(const ctr_handle* h, const RowSrc& r, int B) {
    static const bool off = getenv("CTR_ATTN_OLD") != nullptr;
    return !off && !r.dense && h->cfg.S <= 64;
}
template <int LPR, int VPL, int MINB, bool PEER>
void launch_fwd_idx(ctr_handle* h, const RowSrc& r, int B) {
    const int g = grid_attn(h, B);
    switch (h->cfg.model) {
        case CTR_MODEL_YOUTUBE: k_attn_fwd_idx<LPR, VPL, MODEL_YOUTUBE, PEER, 128, MINB><<<g, 128, 0, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
        case CTR_MODEL_DIN_COS: k_attn_fwd_idx<LPR, VPL, MODEL_DIN_COS, PEER, 128, MINB><<<g, 128, 0, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
        default:                k_attn_fwd_idx<LPR, VPL, MODEL_DIN_EUC, PEER, 128, MINB><<<g, 128, 0, h->stream>>>(r, dims_of(h), h->W[3], h->X0, h->Kp, h->Kp, B); break;
    }
}
template <int LPR, int VPL, int MINB, bool PEER>
void launch_bwd_idx(ctr_handle* h, const RowSrc& r, const BwdOut& o, int B) {
    const int g = grid_attn(h, B);
#define BWD_CASE(M) { if (o.sgd) k_attn_bwd_idx<LPR, VPL, M, true, PEER, 128, MINB><<<g, 128, 0, h->stream>>>(r, dims_of(h), h->W[3], h->dX, h->lddx, o, B); \
                      else       k_attn_bwd_idx<LPR, VPL, M, false, PEER, 128, MINB><<<g, 128, 0, h->stream>>>(r, dims_of(h), h->W[3], h->dX, h->lddx, o, B); }
    switch (h->cfg.model) {
        case CTR_MODEL_YOUTUBE: BWD_CASE(MODEL_YOUTUBE) break;
        case CTR_MODEL_DIN_COS: BWD_CASE(MODEL_DIN_COS) break;
        default:                BWD_CASE(MODEL_DIN_EUC) break;
    }
#undef BWD_CASE
}
template <bool PEER>
void dispatch_fwd_idx(ctr_handle* h, const RowSrc& r, int B) {
    if (PEER && h->cfg.D == 64) {
        // over NVLink the request size matters more than the shuffle count: 8 lanes x 2 float4 (128-byte contiguous
        // segments per load instruction) gathers at 582 GB/s, the local kernels' 4 x 4 mapping (64-byte segments) at
        // 465 GB/s, 16 x 1 (256-byte) at 494 GB/s (profiles/r02/shard_time_2gpu_maps.log).  CTR_PEER_FWD_MAP: 0 = 4x4, 2 = 16x1.
        static const int map = getenv("CTR_PEER_FWD_MAP") ? atoi(getenv("CTR_PEER_FWD_MAP")) : 1;
        if (map == 1) { launch_fwd_idx<8, 2, 8, PEER>(h, r, B); return; }
        if (map == 2) { launch_fwd_idx<16, 1, 8, PEER>(h, r, B); return; }
    }
    switch (h->cfg.D / 4) {
        case 4: launch_fwd_idx<4, 1, 8, PEER>(h, r, B); break;    case 8: launch_fwd_idx<4, 2, 8, PEER>(h, r, B); break;
        case 16: launch_fwd_idx<4, 4, 8, PEER>(h, r, B); break;
        default: launch_fwd_idx<8, 4, 8, PEER>(h, r, B); break;
    }
}
template <bool PEER>
void dispatch_bwd_idx(ctr_handle* h, const RowSrc& r, const BwdOut& o, int B) {
    if (PEER && h->cfg.D == 64) {
        static const int map = getenv("CTR_PEER_BWD_MAP") ? atoi(getenv("CTR_PEER_BWD_MAP")) : 0;
        if (map == 2) { launch_bwd_idx<16, 1, 8, PEER>(h, r, o, B); return; }
    }
    switch (h->cfg.D / 4) {
        case 4: launch_bwd_idx<4, 1, 8, PEER>(h, r, o, B); break;    case 8: launch_bwd_idx<4, 2, 8, PEER>(h, r, o, B); break;
        // D = 64: 8 lanes x 2 float4 (128-byte contiguous red.add / load segments per row) measured faster than
        // 4 x 4 (64-byte segments, 0.39 vs 0.34 ms) although the latter executes a third fewer instructions
        case 16: launch_bwd_idx<8, 2, 8, PEER>(h, r, o, B); break;
        default: launch_bwd_idx<16, 2, 8, PEER>(h, r, o, B); break;
    }
}

int attn_forward(ctr_handle* h, const RowSrc& r, int B) {
    if (r.world > 1)        // row-sharded tables: the owner's HBM is read over NVLink (comm_check vetted the dims)
        return launch(h, "attn_fwd_peer", [&] { dispatch_fwd_idx<true>(h, r, B); });
    if (vec_ok(h, r) && idx_fwd_ok(h, r, B))
        return launch(h, "attn_fwd_vec", [&] { dispatch_fwd_idx<false>(h, r, B); });
    if (vec_ok(h, r)) {
        return launch(h, "attn_fwd_vec", [&] {
            switch (h->cfg.D / 4) {           // float4 per row
                case 4: launch_fwd_vec<4, 1, 8>(h, r, B); break;    case 8: launch_fwd_vec<4, 2, 8>(h, r, B); break;
                case 16: launch_fwd_vec<4, 4, 8>(h, r, B); break;   default: launch_fwd_vec<8, 4, 8>(h, r, B); break;
            }
        });
    }
    return launch(h, "attn_fwd_gen", [&] {
        k_attn_fwd_gen<<<grid_for_warps(h, B), 256, 0, h->stream>>>(r, dims_of(h), h->cfg.model, h->W[3], h->X0, h->Kp, h->Kp, B);
    });
}

int attn_backward(ctr_handle* h, const RowSrc& r, const BwdOut& o, int B) {
    if (r.world > 1)
        return launch(h, "attn_bwd_peer", [&] { dispatch_bwd_idx<true>(h, r, o, B); });
    if (vec_ok(h, r) && idx_bwd_ok(h, r, B))
        return launch(h, "attn_bwd_vec", [&] { dispatch_bwd_idx<false>(h, r, o, B); });
    if (vec_ok(h, r)) {
        return launch(h, "attn_bwd_vec", [&] {
            switch (h->cfg.D / 4) {
                case 4: launch_bwd_vec<4, 1, 8>(h, r, o, B); break;    case 8: launch_bwd_vec<4, 2, 8>(h, r, o, B); break;
                case 16: launch_bwd_vec<8, 2, 8>(h, r, o, B); break;   default: launch_bwd_vec<16, 2, 8>(h, r, o, B); break;
            }
        });
    }
    return launch(h, "attn_bwd_gen", [&] {
        size_t smem = (size_t)std::max(h->cfg.S, 1) * sizeof(float);
        k_attn_bwd_gen<<<grid_for_warps(h, B), 256, smem, h->stream>>>(r, dims_of(h), h->cfg.model, h->W[3], h->dX, h->lddx, o, B);
    });
}

// ---- fp32 GEMM launchers ---------------------------------------------------------------------------
template <bool TA, bool TB, int EPI>
int gemm_big(ctr_handle* h, const char* name, GemmArgs g) {     // M = batch: 128x64 tiles
    constexpr int BM = 128, BN = 64;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, 1);
    return launch(h, name, [&] { k_sgemm<BM, BN, 16, 8, 4, TA, TB, EPI><<<grid, 256, 0, h->stream>>>(g); });
}
int gemm_dw(ctr_handle* h, const char* name, GemmArgs g) {      // K = batch: 64x64 tiles, split-K atomics
    constexpr int BM = 64, BN = 64;
    int tiles = ((g.N + BN - 1) / BN) * ((g.M + BM - 1) / BM);
    int want_z = std::max(1, (h->num_sms * 4) / std::max(tiles, 1));
    int kchunk = round_up(std::max(16, (g.K + want_z - 1) / want_z), 16);
    g.kchunk = kchunk;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (g.K + kchunk - 1) / kchunk);
    return launch(h, name, [&] { k_sgemm<BM, BN, 16, 4, 4, true, false, EPI_ATOMIC><<<grid, 256, 0, h->stream>>>(g); });
}


// ---- tcgen05 GEMM launchers --------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(ctr_handle* h, CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint32_t box_k = 32) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            return set_err(h, CTR_ECUDA, "cuTensorMapEncodeTiled entry point not available");
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_k, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(h, CTR_ECUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return CTR_OK;
}

// {32 fp32, rows, width/32} view of a row-major [rows, width] matrix: dim 2 walks the 32-column blocks (stride 128 bytes),
// so ONE box {32, box_rows, nblocks} lands as nblocks consecutive [box_rows x 128 B] column blocks in shared memory
int make_map_blocks(ctr_handle* h, CUtensorMap* m, const float* base, uint64_t rows, uint64_t width, uint64_t ld, uint32_t box_rows, uint32_t nblocks) {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q; void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return CTR_ECUDA;
        fn = (PFN)p;
    }
    cuuint64_t gdim[3] = {32, rows, width / 32};
    cuuint64_t gstr[2] = {ld * sizeof(float), 32 * sizeof(float)};
    cuuint32_t box[3] = {32, box_rows, nblocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CTR_OK : CTR_ECUDA;
}

constexpr size_t kUmmaEpiBytes = (size_t)umma::kEpiWarps * umma::kEpiStageBytes + 32;     // staged-epilogue tiles
// K elements per shared-memory stage: 32 (128-byte swizzle rows) or 16 (64-byte rows: half-size stages, twice as many).
// The main loop is bound by the L2→SM traffic of the weight tiles, not by the ring depth, so the choice
// matters little: measured, 16 wins by ~2 µs per GEMM for tiles >= 128 columns wide and loses 3.7 µs for the 96-wide one.
// CTR_UMMA_KBK = 16 | 32 forces one value.
int umma_kbk_for(int bn) {
    static const int force = getenv("CTR_UMMA_KBK") ? atoi(getenv("CTR_UMMA_KBK")) : 0;
    if (force == 16 || force == 32) return force;
    return bn >= 128 ? 16 : 32;
}
size_t umma_stage_bytes(int bn, int kbk) { return ((size_t)umma::kBlockM + (size_t)bn) * kbk * 4 * 2; }      // A + A_lo + B_hi + B_lo
int umma_stages(int bn, int kbk) {
    return (int)std::min<size_t>(kbk == 16 ? 6 : 4, ((size_t)227 * 1024 - 2048 - kUmmaEpiBytes) / umma_stage_bytes(bn, kbk));
}
size_t umma_smem(int bn, int stages, int kbk) {
    return (size_t)stages * umma_stage_bytes(bn, kbk) + 8 * (3 * stages + 4) + 16 + 1024 + kUmmaEpiBytes;
}
// transposed accumulation (dX): 256 activation rows per stage, 16 K-elements, no epilogue staging tiles; kTrSlack covers the 128-row
// read of a weight tile with fewer rows (the surplus accumulator rows are never stored)
constexpr int kTrKbk = 16;
constexpr size_t kTrSlack = 4096;
size_t umma_stage_bytes_tr(int bn) { return ((size_t)256 + (size_t)bn) * kTrKbk * 4 * 2; }
int umma_stages_tr(int bn) { return (int)std::min<size_t>(6, ((size_t)227 * 1024 - 2048 - kTrSlack) / umma_stage_bytes_tr(bn)); }
size_t umma_smem_tr(int bn, int stages) { return (size_t)stages * umma_stage_bytes_tr(bn) + 8 * (3 * stages + 4) + 16 + 1024 + kTrSlack; }
// worth it once the 256-row tiles still fill the machine; CTR_UMMA_NO_TR=1 keeps every GEMM on 128-row tiles
bool umma_use_tr(const ctr_handle* h, int M, int bn) {
    static const bool off = getenv("CTR_UMMA_NO_TR") != nullptr;
    // bn == 128 only: a narrower weight tile would still be read as 128 rows by the MMA, i.e. past its own buffer
    return !off && bn == 128 && (M + 255) / 256 >= h->num_sms;
}

bool umma_supported(const ctr_handle* h) {
    return h->H0p <= 256 && h->H1p <= 256 && round_up(2 * h->cfg.D, 16) <= 256 && h->H0p % 16 == 0 && h->H1p % 16 == 0;
}

// shared memory of one k-block of the weight-gradient GEMM: 8 A blocks + nb B blocks of (ks samples x 128 bytes), twice with lo copies
constexpr size_t kDwSlack = 4096;          // one 32-sample column block
template <class U> size_t dw_stage_bytes(const U& u, int nb) { return (size_t)(8 + nb) * u.dw_ks * 128 * (u.dw_terms == 3 ? 2 : 1); }

int umma_init(ctr_handle* h) {
    if (h->um.ready) return CTR_OK;
    const ctr_config& c = h->cfg;
    auto& u = h->um;
    u.bn_fwd0 = h->H0p; u.bn_fwd1 = h->H1p; u.bn_dx = round_up(2 * c.D, 16);
    u.kbk_fwd0 = u.kbk_dz0 = umma_kbk_for(u.bn_fwd0); u.kbk_fwd1 = umma_kbk_for(u.bn_fwd1); u.kbk_dx = umma_kbk_for(u.bn_dx);
    u.stages_fwd0 = u.stages_dz0 = umma_stages(u.bn_fwd0, u.kbk_fwd0); u.stages_fwd1 = umma_stages(u.bn_fwd1, u.kbk_fwd1);
    u.stages_dx = umma_stages(u.bn_dx, u.kbk_dx);
    if (u.stages_fwd0 < 2) return set_err(h, CTR_EINVAL, "tcgen05 GEMM: tile does not fit shared memory");
    auto ksw = [](int kbk) { return kbk == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B; };
    for (int i = 0; i < 2; i++) {
        RET(dalloc(h, &u.Wt0[i], (size_t)h->H0p * h->Kp)); RET(dalloc(h, &u.Wt1[i], (size_t)h->H1p * h->H0p));
        RET(dalloc(h, &u.W1s[i], (size_t)h->H0p * h->H1p)); RET(dalloc(h, &u.W0s[i], (size_t)u.bn_dx * h->H0p));
        RET(make_map(h, &u.mB_Wt0[i], u.Wt0[i], h->H0p, h->Kp, h->Kp, u.bn_fwd0, ksw(u.kbk_fwd0), u.kbk_fwd0));
        RET(make_map(h, &u.mB_Wt1[i], u.Wt1[i], h->H1p, h->H0p, h->H0p, u.bn_fwd1, ksw(u.kbk_fwd1), u.kbk_fwd1));
        RET(make_map(h, &u.mB_W1s[i], u.W1s[i], h->H0p, h->H1p, h->H1p, u.bn_fwd0, ksw(u.kbk_dz0), u.kbk_dz0));
        RET(make_map(h, &u.mB_W0s[i], u.W0s[i], u.bn_dx, h->H0p, h->H0p, u.bn_dx, ksw(u.kbk_dx), u.kbk_dx));
    }
    if (u.bn_dx == 128) {
        RET(make_map(h, &u.mA_dZ0_tr, h->dZ0, h->Bmax, h->H0p, h->H0p, 256, ksw(kTrKbk), kTrKbk));
        for (int i = 0; i < 2; i++) RET(make_map(h, &u.mB_W0s_tr[i], u.W0s[i], u.bn_dx, h->H0p, h->H0p, u.bn_dx, ksw(kTrKbk), kTrKbk));
        u.stages_dx_tr = umma_stages_tr(u.bn_dx);
    }
    RET(make_map(h, &u.mA_X0, h->X0, h->Bmax, h->Kp, h->Kp, umma::kBlockM, ksw(u.kbk_fwd0), u.kbk_fwd0));
    RET(make_map(h, &u.mA_H0d, h->H0d, h->Bmax, h->H0p, h->H0p, umma::kBlockM, ksw(u.kbk_fwd1), u.kbk_fwd1));
    RET(make_map(h, &u.mA_dZ1, h->dZ1, h->Bmax, h->H1p, h->H1p, umma::kBlockM, ksw(u.kbk_dz0), u.kbk_dz0));
    RET(make_map(h, &u.mA_dZ0, h->dZ0, h->Bmax, h->H0p, h->H0p, umma::kBlockM, ksw(u.kbk_dx), u.kbk_dx));
    // weight-gradient operands: rows limited to the training batch so that TMA zero-fills the K tail
    // weight gradients: error-compensated 3xTF32 like the other GEMMs (hi/lo copies double a stage, so 16-sample
    // k-blocks); CTR_DW_1XTF32=1 selects the round-1 arithmetic (one TF32-RN product per term, 32-sample k-blocks)
    static const bool dw1x = getenv("CTR_DW_1XTF32") != nullptr;
    u.dw_terms = dw1x ? 1 : 3; u.dw_ks = dw1x ? 32 : 16;
    RET(make_map(h, &u.mK_X0, h->X0, c.batch, h->Kp, h->Kp, u.dw_ks, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    RET(make_map(h, &u.mK_H0d, h->H0d, c.batch, h->H0p, h->H0p, u.dw_ks, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    RET(make_map(h, &u.mK_dZ0, h->dZ0, c.batch, h->H0p, h->H0p, u.dw_ks, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    RET(make_map(h, &u.mK_dZ1, h->dZ1, c.batch, h->H1p, h->H1p, u.dw_ks, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    // producer mode: one 3-D box per operand when the driver accepts the block view (CTR_DW_TMA = 0 | 1 | 2 overrides)
    u.dw_tma = getenv("CTR_DW_TMA") ? atoi(getenv("CTR_DW_TMA")) : 2;
    if (u.dw_tma == 2 && (make_map_blocks(h, &u.m3_X0, h->X0, c.batch, h->Kp, h->Kp, u.dw_ks, h->Kp / 32) != CTR_OK ||
                          make_map_blocks(h, &u.m3_H0d, h->H0d, c.batch, h->H0p, h->H0p, u.dw_ks, h->H0p / 32) != CTR_OK ||
                          make_map_blocks(h, &u.m3_dZ0, h->dZ0, c.batch, h->H0p, h->H0p, u.dw_ks, h->H0p / 32) != CTR_OK ||
                          make_map_blocks(h, &u.m3_dZ1, h->dZ1, c.batch, h->H1p, h->H1p, u.dw_ks, h->H1p / 32) != CTR_OK))
        u.dw_tma = 1;
    // kDwSlack: an operand with fewer than four 32-column blocks is still read as 128 rows by the MMA — the surplus rows
    // (never stored) come from whatever follows, which must lie inside the allocation
    u.dw_stages0 = (int)std::min<size_t>(6, ((size_t)227 * 1024 - 2048 - kDwSlack) / dw_stage_bytes(u, h->H0p / 32));
    u.dw_stages1 = (int)std::min<size_t>(6, ((size_t)227 * 1024 - 2048 - kDwSlack) / dw_stage_bytes(u, h->H1p / 32));
    CU(h, cudaFuncSetAttribute(umma::k_umma_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const size_t smax = (size_t)227 * 1024;
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_SIGMOID_DROP, true, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_SIGMOID_DROP, true, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_DSIGMOID, true, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_DSIGMOID, true, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_STORE, true, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_STORE, true, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    CU(h, cudaFuncSetAttribute(umma::k_umma_gemm<umma::UEPI_STORE, true, kTrKbk, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
    if (getenv("CTR_UMMA_TIMELINE")) RET(dalloc(h, &h->umma_dbg, 8192));
    u.ready = true; u.dirty = true;
    return CTR_OK;
}

int umma_split_weights(ctr_handle* h) {
    auto& u = h->um;
    if (!u.dirty) return CTR_OK;
    const ctr_config& c = h->cfg;
    umma::SplitJobs jobs;
    jobs.j[0] = {h->W[0], h->H0p, h->in, c.H0, 1, u.Wt0[0], u.Wt0[1], h->Kp};                                      // W0ᵀ [H0, in]
    jobs.j[1] = {h->W[1], h->H1p, c.H0, c.H1, 1, u.Wt1[0], u.Wt1[1], h->H0p};                                     // W1ᵀ [H1, H0]
    jobs.j[2] = {h->W[1], h->H1p, c.H0, c.H1, 0, u.W1s[0], u.W1s[1], h->H1p};                                     // W1  [H0, H1]
    jobs.j[3] = {h->W[0] + (long)c.uP * h->H0p, h->H0p, 2 * c.D, c.H0, 0, u.W0s[0], u.W0s[1], h->H0p};            // W0[uP:uP+2D, :]
    RET(launch(h, "split_weights_tf32", [&] { umma::k_split_weights<<<dim3(48, 4), 256, 0, h->stream>>>(jobs); }));
    u.dirty = false;
    return CTR_OK;
}

template <int EPI>
int umma_gemm(ctr_handle* h, const char* name, const CUtensorMap& mA, const CUtensorMap* mB, umma::Args a) {
    const int tile_rows = a.tr ? 256 : umma::kBlockM;
    const int tiles = (a.M + tile_rows - 1) / tile_rows;
    const int grid = std::min(tiles, h->num_sms);
    const size_t smem = a.tr ? umma_smem_tr(a.bn, a.stages) : umma_smem(a.bn, a.stages, a.kbk);
    a.dbg = h->umma_dbg;
    static const int pf = getenv("CTR_UMMA_PF") ? atoi(getenv("CTR_UMMA_PF")) : 0;     // measured: no gain (profiles/r02/EXPERIMENTS.md)
    a.pf = std::max(0, std::min(pf, 64));
    // kind::tf32 ignores the low 13 mantissa bits of an operand word, so the raw activation tile is A_hi as it lies: the
    // converters only write A_lo (bit-identical scores, test_raw_tile_as_hi_operand_is_bit_identical).  CTR_UMMA_RAWHI=0
    // writes the truncated copy in place as well.
    static const int rawhi = getenv("CTR_UMMA_RAWHI") ? atoi(getenv("CTR_UMMA_RAWHI")) : 1;
    a.rawhi = rawhi;
    static const bool epi_old = getenv("CTR_UMMA_EPI_OLD") != nullptr;
    a.staged_epi = epi_old ? 0 : 1;
    int rc = launch(h, name, [&] {
        if constexpr (EPI == umma::UEPI_STORE) {
            if (a.tr) { umma::k_umma_gemm<EPI, true, kTrKbk, true><<<grid, 448, smem, h->stream>>>(mA, mB[0], mB[1], a); return; }
        }
        if (a.kbk == 16) umma::k_umma_gemm<EPI, true, 16, false><<<grid, 448, smem, h->stream>>>(mA, mB[0], mB[1], a);
        else             umma::k_umma_gemm<EPI, true, 32, false><<<grid, 448, smem, h->stream>>>(mA, mB[0], mB[1], a);
    });
    if (rc == CTR_OK && h->umma_dbg) {
        std::vector<unsigned long long> t(8192);
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(t.data(), h->umma_dbg, 8192 * 8, cudaMemcpyDeviceToHost);
        cudaMemset(h->umma_dbg, 0, 8192 * 8);
        FILE* f = fopen("gpurun_out/umma_timeline.txt", "a");
        if (f) {
            unsigned long long t0 = t[0];
            fprintf(f, "# %s M=%d N=%d K=%d bn=%d kbk=%d stages=%d\n", name, a.M, a.N, a.K, a.bn, a.kbk, a.stages);
            for (int it = 0; it < 40 && t[it * 8]; it++)
                fprintf(f, "it %d tma_issue %llu conv_saw_full %llu conv_done %llu mma_saw %llu mma_issued %llu\n", it, t[it * 8] - t0, t[it * 8 + 1] - t0, t[it * 8 + 2] - t0, t[it * 8 + 3] - t0, t[it * 8 + 4] - t0);
            for (int tc = 0; tc < 5 && t[4096 + tc * 4]; tc++) fprintf(f, "tile %d epi_start %llu epi_done %llu\n", tc, t[4096 + tc * 4] - t0, t[4096 + tc * 4 + 1] - t0);
            fclose(f);
        }
    }
    return rc;
}

int umma_dw(ctr_handle* h, const char* name, const CUtensorMap& mA, const CUtensorMap& mB, umma::DwArgs a) {
    a.ks = h->um.dw_ks; a.terms = h->um.dw_terms; a.tma = h->um.dw_tma;
    // tcgen05 TF32 MMAs cover 8 samples each and take >= ~65 ns whatever their width: the 3-term main loop is bound by their
    // COUNT.  When C has <= 128 columns but > 128 rows (dW1: 200 x 96), accumulating Cᵀ needs one MMA per term, not two.
    static const bool noswap = getenv("CTR_DW_NO_SWAP") != nullptr;
    // nb >= 3: the 128-row read of the narrow operand then overruns its buffer by at most one block (kDwSlack)
    a.swap = (!noswap && a.M > 128 && a.N <= 128 && a.nb >= 3) ? 1 : 0;
    static const int rawhi = getenv("CTR_UMMA_RAWHI") ? atoi(getenv("CTR_UMMA_RAWHI")) : 1;
    a.rawhi = rawhi;
    const int total_kb = (a.K + a.ks - 1) / a.ks;
    const int grid = std::max(1, std::min(total_kb, h->num_sms));
    const size_t smem = (size_t)a.stages * dw_stage_bytes(h->um, a.nb) + 8 * (3 * a.stages + 2) + 16 + 1024 + kDwSlack;
    static const int pf = getenv("CTR_UMMA_DW_PF") ? atoi(getenv("CTR_UMMA_DW_PF")) : 0;
    a.pf = std::max(0, std::min(pf, 16));
    static const bool norot = getenv("CTR_DW_NO_ROTATE") != nullptr;
    a.rotate = norot ? 0 : 1;
    a.tl = h->umma_dbg;
    int rc = launch(h, name, [&] { umma::k_umma_dw<<<grid, 448, smem, h->stream>>>(mA, mB, a); });
    if (rc == CTR_OK && h->umma_dbg) {
        std::vector<unsigned long long> t(1024);
        cudaStreamSynchronize(h->stream);
        cudaMemcpy(t.data(), h->umma_dbg, 1024 * 8, cudaMemcpyDeviceToHost);
        cudaMemset(h->umma_dbg, 0, 1024 * 8);
        FILE* f = fopen("gpurun_out/umma_timeline.txt", "a");
        if (f) {
            fprintf(f, "# %s K=%d na=%d nb=%d ks=%d terms=%d stages=%d pf=%d grid=%d\n", name, a.K, a.na, a.nb, a.ks, a.terms, a.stages, a.pf, grid);
            fprintf(f, "first_box_landed %llu accumulators_done %llu kernel_end %llu  mma_batches:", t[1] - t[0], t[2] - t[0], t[3] - t[0]);
            for (int i = 8; i < 64 && t[i]; i++) fprintf(f, " %llu", t[i] - t[0]);
            fprintf(f, "\n");
            for (int it = 0; it < 24 && t[128 + it * 4]; it++)
                fprintf(f, "it %d tma_issue %llu conv_saw_full %llu conv_arrived %llu mma_saw_conv %llu\n", it, t[128 + it * 4] - t[0], t[128 + it * 4 + 1] - t[0],
                        t[128 + it * 4 + 2] - t[0], t[128 + it * 4 + 3] - t[0]);
            fclose(f);
        }
    }
    return rc;
}

// the weight-gradient GEMMs run on tcgen05 when the padded widths fit 8 column blocks and the batch
// is the configured training batch (the TMA maps zero-fill rows beyond it)
bool use_umma_dw(const ctr_handle* h, int B) {
    static const bool off = getenv("CTR_DW_FP32") != nullptr;
    return B == h->cfg.batch && h->Kp <= 256 && h->H0p <= 256 && h->H1p <= 256 && !off;
}

bool use_umma(const ctr_handle* h) {
    return h->cfg.gemm != CTR_GEMM_FP32 && umma_supported(h);     // AUTO → tcgen05 when the shape qualifies
}

// One pass of the hot path over one batch (model.go:107-196 inner loop body).
//   training: dropout on, backward + optimiser step (update) or gradients only (!update)
struct StepOpts {
    bool training = false, update = false;
    bool want_rows = false;              // fill h->dUb / h->dIt
    const float* d_label = nullptr;
    // sharded-table step (comm_impl.cuh): rows live in a per-lookup buffer, gradients are scaled to the
    // global mean and the dense gradients are all-reduced before Adam
    bool comm = false; float grad_scale = 1.0f; int adam_batch = 0;
    bool peer = false;                   // rows come from / go to the owners' shards over NVLink (RowSrc.world > 1)
    // replicated-table step: row gradients are summed into table_grad (layout of the table), all-reduced with the
    // dense gradients and applied on every rank
    float* table_grad = nullptr; size_t table_grad_n = 0;
};

int ensure_rowgrad_buffers(ctr_handle* h, int B) {
    size_t need = (size_t)B * h->cfg.S * h->cfg.D;
    if (h->dUb_cap >= need && h->dUb) return CTR_OK;
    if (h->dUb) cudaFree(h->dUb);
    if (h->dIt) cudaFree(h->dIt);
    h->dUb = h->dIt = nullptr;
    RET(dalloc(h, &h->dUb, need));
    RET(dalloc(h, &h->dIt, (size_t)B * h->cfg.D));
    h->dUb_cap = need;
    return CTR_OK;
}

// Replica accumulators for the popular rows of ITEM_EMB (rows [0, hot_rows), see k_attn_bwd_vec).
// cfg.reserved[0]: 0 = auto (all rows of a small table / the first 32768 of a large one — keep rows
// ordered by popularity, as word2vec-style vocabularies are), >0 = that many rows, <0 = off.
int ensure_hot(ctr_handle* h) {
    const int64_t rows = h->tab_local_rows[CTR_TABLE_ITEM_EMB];
    int want = h->cfg.reserved[0];
    if (want == 0) want = (int)std::min<int64_t>(rows, 32768);
    if (want < 0) want = 0;
    want = (int)std::min<int64_t>(want, rows);
    if (h->hot_acc && h->hot_rows == want) return CTR_OK;
    if (h->hot_acc) { cudaFree(h->hot_acc); h->hot_acc = nullptr; }
    h->hot_rows = want; h->hot_reps = 0;
    if (want == 0) return CTR_OK;
    const size_t row_bytes = (size_t)h->cfg.D * sizeof(float);
    int reps = (int)std::min<size_t>(32, std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)want * row_bytes)));
    h->hot_reps = reps;
    return dalloc(h, &h->hot_acc, (size_t)reps * want * h->cfg.D);
}

// CTR_TABLE_ADAM: moments of the embedding rows (zero-initialised, like the solver's, model.go:88)
int ensure_moments(ctr_handle* h) {
    if (h->emb_m) return CTR_OK;
    const size_t n = (size_t)std::max<int64_t>(h->tab_local_rows[CTR_TABLE_ITEM_EMB], 1) * h->tab_ld[CTR_TABLE_ITEM_EMB];
    RET(dalloc(h, &h->emb_m, n)); RET(dalloc(h, &h->emb_v, n));
    return CTR_OK;
}

int deterministic_table_update(ctr_handle* h, const RowSrc& r, int B, int adam_batch) {
    const int S = h->cfg.S, D = h->cfg.D;
    size_t n = (size_t)B * (S + 1);
    if (h->keys_cap < n) {
        for (void* p : {(void*)h->keys, (void*)h->keys2, (void*)h->pos, (void*)h->pos2, h->sort_tmp}) if (p) cudaFree(p);
        h->keys = h->keys2 = h->pos = h->pos2 = nullptr; h->sort_tmp = nullptr;
        RET(dalloc(h, &h->keys, n, false)); RET(dalloc(h, &h->keys2, n, false));
        RET(dalloc(h, &h->pos, n, false)); RET(dalloc(h, &h->pos2, n, false));
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->keys, h->keys2, h->pos, h->pos2, (int)n, 0, 32, h->stream);
        CU(h, cudaMalloc(&h->sort_tmp, bytes));
        h->sort_tmp_bytes = bytes; h->keys_cap = n;
    }
    RET(launch(h, "scatter_keys", [&] {
        k_scatter_keys<<<std::min<int>((int)((n + 255) / 256), h->num_sms * 8), 256, 0, h->stream>>>(r.hist, r.item_row, S, B, r.nvalid, r.n_emb, h->keys, h->pos);
    }));
    RET(launch(h, "cub_radix_sort_pairs", [&] {
        cub::DeviceRadixSort::SortPairs(h->sort_tmp, h->sort_tmp_bytes, h->keys, h->keys2, h->pos, h->pos2, (int)n, 0, 32, h->stream);
    }));
    if (h->cfg.table_opt == CTR_TABLE_ADAM) {
        RET(ensure_moments(h));
        const ctr_config& c = h->cfg;
        const int t = (int)h->step + 1;                    // the dense solver's step counter (k_adam runs after this)
        RowAdam ra{c.table_lr, c.beta1, c.beta2, c.eps, (float)(1.0 - std::pow((double)c.beta1, (double)t)), (float)(1.0 - std::pow((double)c.beta2, (double)t)),
                   adam_batch > 1 ? 1.0f / (float)adam_batch : 1.0f};
        return launch(h, "segment_adam", [&] {
            k_segment_adam<<<std::min<int>((int)((n * 32 + 255) / 256), h->num_sms * 16), 256, 0, h->stream>>>(
                h->keys2, h->pos2, (long)n, h->dUb, h->dIt, S, D, h->tab[CTR_TABLE_ITEM_EMB], h->emb_m, h->emb_v, h->tab_ld[CTR_TABLE_ITEM_EMB], ra);
        });
    }
    RET(launch(h, "segment_sgd", [&] {
        k_segment_sgd<<<std::min<int>((int)((n * 32 + 255) / 256), h->num_sms * 16), 256, 0, h->stream>>>(
            h->keys2, h->pos2, (long)n, h->dUb, h->dIt, S, D, h->tab[CTR_TABLE_ITEM_EMB], h->tab_ld[CTR_TABLE_ITEM_EMB], h->cfg.table_lr);
    }));
    return CTR_OK;
}

int step_core(ctr_handle* h, const RowSrc& r, int B, const StepOpts& o) {
    const ctr_config& c = h->cfg;
    if (B <= 0 || B > h->Bmax) return set_err(h, CTR_EINVAL, "batch %d outside (0, %d]", B, h->Bmax);
    const float d0 = o.training ? c.dropout0 : 0.0f, d1 = o.training ? c.dropout1 : 0.0f;

    RET(attn_forward(h, r, B));
    const bool um = use_umma(h);
    if (um) {
        RET(umma_init(h));
        RET(umma_split_weights(h));
        umma::Args a{}; a.M = B; a.N = c.H0; a.Nz = h->H0p; a.K = h->Kp; a.bn = h->um.bn_fwd0; a.C = h->H0d; a.ldc = h->H0p;
        a.drop_p = d0; a.seed = c.seed; a.stream = h->step * 4u + 0u; a.stages = h->um.stages_fwd0; a.kbk = h->um.kbk_fwd0;
        RET(umma_gemm<umma::UEPI_SIGMOID_DROP>(h, "umma_fwd0_sigmoid", h->um.mA_X0, h->um.mB_Wt0, a));
        umma::Args b{}; b.M = B; b.N = c.H1; b.Nz = h->H1p; b.K = h->H0p; b.bn = h->um.bn_fwd1; b.C = h->H1d; b.ldc = h->H1p;
        b.drop_p = d1; b.seed = c.seed; b.stream = h->step * 4u + 1u; b.stages = h->um.stages_fwd1; b.kbk = h->um.kbk_fwd1;
        RET(umma_gemm<umma::UEPI_SIGMOID_DROP>(h, "umma_fwd1_sigmoid", h->um.mA_H0d, h->um.mB_Wt1, b));
    } else {
    {   // h0 = dropout(sigmoid(x·W0))   din.go:307-308
        GemmArgs g{}; g.A = h->X0; g.lda = h->Kp; g.B = h->W[0]; g.ldb = h->H0p; g.C = h->H0d; g.ldc = h->H0p;
        g.M = B; g.N = c.H0; g.K = h->in; g.Nz = h->H0p; g.drop_p = d0; g.seed = c.seed; g.stream = h->step * 4u + 0u;
        RET((gemm_big<false, false, EPI_SIGMOID_DROP>(h, "sgemm_fwd0_sigmoid", g)));
    }
    {   // h1 = dropout(sigmoid(h0·W1))  din.go:311-312
        GemmArgs g{}; g.A = h->H0d; g.lda = h->H0p; g.B = h->W[1]; g.ldb = h->H1p; g.C = h->H1d; g.ldc = h->H1p;
        g.M = B; g.N = c.H1; g.K = c.H0; g.Nz = h->H1p; g.drop_p = d1; g.seed = c.seed; g.stream = h->step * 4u + 1u;
        RET((gemm_big<false, false, EPI_SIGMOID_DROP>(h, "sgemm_fwd1_sigmoid", g)));
    }
    }
    const bool bwd = o.training;
    if (bwd) CU(h, cudaMemsetAsync(h->d_cost, 0, sizeof(double), h->stream));
    {
        HeadArgs a{}; a.H1d = h->H1d; a.ldh = h->H1p; a.H1 = c.H1; a.H1p = h->H1p; a.w2 = h->W[2];
        a.y = bwd ? o.d_label : nullptr; a.B = B; a.nvalid = r.nvalid; a.drop_p = d1;
        a.p = h->P; a.logit = h->Z; a.dZ1 = bwd ? h->dZ1 : nullptr; a.lddz = h->H1p; a.dW2 = h->G[2]; a.cost_sum = h->d_cost;
        // few, long-lived blocks: every block ends with H1 same-address atomics (dW2) and one on the cost — with one block per
        // 64 rows those serialised in L2 (~27 cycles each per address) and cost more than the 50 MB the kernel streams
        static const int hb = getenv("CTR_HEAD_BLOCKS") ? atoi(getenv("CTR_HEAD_BLOCKS")) : 2;
        const int hgrid = std::max(1, std::min((B + 63) / 64, h->num_sms * std::max(1, hb)));
        RET(launch(h, bwd ? "head_fwd_bwd" : "head_fwd", [&] { k_head<<<hgrid, 256, 0, h->stream>>>(a); }));
    }
    if (!bwd) return CTR_OK;

    const bool umdw = um && use_umma_dw(h, B);
    if (umdw) {   // dW1 += h0dᵀ · dZ1   (tcgen05, MN-major operands, split-K over the grid)
        umma::DwArgs a{}; a.K = B; a.na = h->H0p / 32; a.nb = h->H1p / 32; a.M = c.H0; a.N = h->H1p; a.C = h->G[1]; a.ldc = h->H1p;
        a.stages = h->um.dw_stages1;
        const bool t3 = h->um.dw_tma == 2;
        RET(umma_dw(h, "umma_dW1_splitk", t3 ? h->um.m3_H0d : h->um.mK_H0d, t3 ? h->um.m3_dZ1 : h->um.mK_dZ1, a));
    } else {   // dW1 += h0dᵀ · dZ1
        GemmArgs g{}; g.A = h->H0d; g.lda = h->H0p; g.B = h->dZ1; g.ldb = h->H1p; g.C = h->G[1]; g.ldc = h->H1p;
        g.M = c.H0; g.N = c.H1; g.K = B;
        RET(gemm_dw(h, "sgemm_dW1_splitk", g));
    }
    if (um) {
        umma::Args a{}; a.M = B; a.N = c.H0; a.Nz = h->H0p; a.K = h->H1p; a.bn = h->um.bn_fwd0; a.C = h->dZ0; a.ldc = h->H0p;
        a.H = h->H0d; a.ldh = h->H0p; a.drop_p = d0; a.stages = h->um.stages_dz0; a.kbk = h->um.kbk_dz0;
        RET(umma_gemm<umma::UEPI_DSIGMOID>(h, "umma_dZ0_dsigmoid", h->um.mA_dZ1, h->um.mB_W1s, a));
    } else {   // dZ0 = (dZ1 · W1ᵀ) ⊙ dsigmoid(h0d)
        GemmArgs g{}; g.A = h->dZ1; g.lda = h->H1p; g.B = h->W[1]; g.ldb = h->H1p; g.C = h->dZ0; g.ldc = h->H0p;
        g.M = B; g.N = c.H0; g.K = c.H1; g.Nz = h->H0p; g.H = h->H0d; g.ldh = h->H0p; g.drop_p = d0;
        RET((gemm_big<false, true, EPI_DSIGMOID>(h, "sgemm_dZ0_dsigmoid", g)));
    }
    if (umdw) {   // dW0 += x0ᵀ · dZ0
        umma::DwArgs a{}; a.K = B; a.na = h->Kp / 32; a.nb = h->H0p / 32; a.M = h->in; a.N = h->H0p; a.C = h->G[0]; a.ldc = h->H0p;
        a.stages = h->um.dw_stages0;
        const bool t3 = h->um.dw_tma == 2;
        RET(umma_dw(h, "umma_dW0_splitk", t3 ? h->um.m3_X0 : h->um.mK_X0, t3 ? h->um.m3_dZ0 : h->um.mK_dZ0, a));
    } else {   // dW0 += x0ᵀ · dZ0
        GemmArgs g{}; g.A = h->X0; g.lda = h->Kp; g.B = h->dZ0; g.ldb = h->H0p; g.C = h->G[0]; g.ldc = h->H0p;
        g.M = h->in; g.N = c.H0; g.K = B;
        RET(gemm_dw(h, "sgemm_dW0_splitk", g));
    }
    const bool learn_rows = o.update && c.table_opt != CTR_TABLE_FROZEN && !r.dense;
    const bool need_attn_bwd = c.model != CTR_MODEL_YOUTUBE || learn_rows || o.want_rows;
    if (need_attn_bwd) {
        if (um) {
            umma::Args a{}; a.M = B; a.N = 2 * c.D; a.Nz = h->lddx; a.K = h->H0p; a.bn = h->um.bn_dx; a.C = h->dX; a.ldc = h->lddx;
            a.stages = h->um.stages_dx; a.kbk = h->um.kbk_dx;
            if (umma_use_tr(h, B, a.bn) && h->um.stages_dx_tr >= 2) {
                a.tr = 1; a.kbk = kTrKbk; a.stages = h->um.stages_dx_tr;
                RET(umma_gemm<umma::UEPI_STORE>(h, "umma_dX", h->um.mA_dZ0_tr, h->um.mB_W0s_tr, a));
            } else
            RET(umma_gemm<umma::UEPI_STORE>(h, "umma_dX", h->um.mA_dZ0, h->um.mB_W0s, a));
        } else {   // d concat[:, uP:uP+2D] = dZ0 · W0[uP:uP+2D, :]ᵀ
            GemmArgs g{}; g.A = h->dZ0; g.lda = h->H0p; g.B = h->W[0] + (long)c.uP * h->H0p; g.ldb = h->H0p;
            g.C = h->dX; g.ldc = h->lddx; g.M = B; g.N = 2 * c.D; g.K = c.H0; g.Nz = h->lddx;
            RET((gemm_big<false, true, EPI_STORE>(h, "sgemm_dX", g)));
        }
        const bool fused_only = o.comm || o.table_grad;        // multi-GPU steps always scatter with red.add
        const bool sorted_rows = c.table_opt == CTR_TABLE_SGD_DETERMINISTIC || c.table_opt == CTR_TABLE_ADAM;   // per-unique-row update from sorted keys
        const bool adam_rows = learn_rows && c.table_opt == CTR_TABLE_ADAM;
        if (adam_rows && o.peer) return set_err(h, CTR_ESTATE, "CTR_TABLE_ADAM is not available with row-sharded tables (use CTR_TABLE_SGD, or a replicated table)");
        const bool buffers = o.want_rows || (learn_rows && sorted_rows && !fused_only);
        if (buffers) RET(ensure_rowgrad_buffers(h, B));
        BwdOut bo{}; bo.datt = h->G[3]; bo.dUb = buffers ? h->dUb : nullptr; bo.dIt = buffers ? h->dIt : nullptr;
        bo.sgd = (learn_rows && (c.table_opt == CTR_TABLE_SGD || fused_only)) ? 1 : 0;
        // what leaves the kernel: -lr/world * gradient (SGD: added straight into rows), or 1/world * gradient (Adam on a
        // replicated table: the all-reduced sum is the global-batch gradient the row solver consumes)
        bo.neg_lr = (adam_rows ? 1.0f : -c.table_lr) * o.grad_scale;
        bo.scatter_base = o.table_grad ? o.table_grad : h->tab[CTR_TABLE_ITEM_EMB];
        const bool hot = bo.sgd && vec_ok(h, r) && !o.comm;
        if (hot) { RET(ensure_hot(h)); bo.hot_acc = h->hot_acc; bo.hot_rows = h->hot_rows; bo.hot_reps = h->hot_reps; }
        const bool peer_hot = o.peer && bo.sgd && h->comm.hot_k > 0;
        if (peer_hot) { bo.hot_acc = h->comm.hot_acc; bo.hot_rows = h->comm.hot_k; bo.hot_reps = h->comm.hot_reps; }
        // sharded tables: nobody's red.add may land in a row before every rank's forward has read it
        if (o.peer && bo.sgd) RET(comm_barrier(h));
        RET(attn_backward(h, r, bo, B));
        if (peer_hot)      // replica accumulators of the replicated hot rows → this rank's gradient sum (all-reduced below)
            RET(launch(h, "hot_rows_fold", [&] {
                k_hot_apply<<<h->num_sms * 4, 256, 0, h->stream>>>(h->comm.hot_sum, c.D, h->comm.hot_acc, h->comm.hot_k, h->comm.hot_reps, c.D, 1.0f);
            }));
        if (hot && h->hot_rows > 0)
            RET(launch(h, "hot_rows_apply", [&] {
                k_hot_apply<<<h->num_sms * 4, 256, 0, h->stream>>>(bo.scatter_base, h->tab_ld[CTR_TABLE_ITEM_EMB], h->hot_acc,
                                                                h->hot_rows, h->hot_reps, c.D, 1.0f);
            }));
        if (learn_rows && sorted_rows && !fused_only) RET(deterministic_table_update(h, r, B, o.adam_batch > 0 ? o.adam_batch : B));
    }
    if (o.update) {
        const bool tg = o.table_grad && c.table_opt != CTR_TABLE_FROZEN;
        const bool ph = o.peer && c.table_opt != CTR_TABLE_FROZEN && h->comm.hot_k > 0;
        if (o.comm || o.table_grad) RET(comm_allreduce_grads(h, tg ? o.table_grad : ph ? h->comm.hot_sum : nullptr, tg ? o.table_grad_n : ph ? (size_t)h->comm.hot_k * c.D : 0));
        if (ph)
            RET(launch(h, "hot_rows_apply", [&] {
                k_apply_table_grad<<<h->num_sms * 8, 256, 0, h->stream>>>(h->comm.hot_tab, h->comm.hot_sum, (long)h->comm.hot_k * c.D / 4);
            }));
        if (tg && c.table_opt == CTR_TABLE_ADAM) {
            RET(ensure_moments(h));
            const int t = (int)h->step + 1; const int ab = o.adam_batch > 0 ? o.adam_batch : B;
            RowAdam ra{c.table_lr, c.beta1, c.beta2, c.eps, (float)(1.0 - std::pow((double)c.beta1, (double)t)), (float)(1.0 - std::pow((double)c.beta2, (double)t)),
                       ab > 1 ? 1.0f / (float)ab : 1.0f};
            RET(launch(h, "apply_table_adam", [&] {
                k_apply_table_adam<<<h->num_sms * 8, 256, 0, h->stream>>>(h->tab[CTR_TABLE_ITEM_EMB], o.table_grad, h->emb_m, h->emb_v,
                                                                         (long)h->tab_rows[CTR_TABLE_ITEM_EMB], (int)h->tab_ld[CTR_TABLE_ITEM_EMB], ra);
            }));
        } else if (tg)
            RET(launch(h, "apply_table_grad", [&] {
                k_apply_table_grad<<<h->num_sms * 8, 256, 0, h->stream>>>(h->tab[CTR_TABLE_ITEM_EMB], o.table_grad, (long)(o.table_grad_n / 4));
            }));
        AdamArgs a{};
        const int in = h->in;
        a.t[0] = AdamTensor{h->W[0], h->G[0], h->Mo[0], h->Vo[0], in, c.H0, h->H0p};
        a.t[1] = AdamTensor{h->W[1], h->G[1], h->Mo[1], h->Vo[1], c.H0, c.H1, h->H1p};
        a.t[2] = AdamTensor{h->W[2], h->G[2], h->Mo[2], h->Vo[2], 1, c.H1, h->H1p};
        a.t[3] = AdamTensor{h->W[3], h->G[3], h->Mo[3], h->Vo[3], 1, c.S, h->Sp};
        a.nt = c.model == CTR_MODEL_YOUTUBE ? 3 : 4;       // Learnable(): dnn.go:153 vs din.go:161-169
        const int t = (int)h->step + 1;
        const int ab = o.adam_batch > 0 ? o.adam_batch : B;
        a.gscale = o.grad_scale;
        a.lr = c.lr; a.l2 = c.l2; a.inv_batch = ab > 1 ? 1.0f / (float)ab : 1.0f; a.b1 = c.beta1; a.b2 = c.beta2; a.eps = c.eps;
        a.c1 = (float)(1.0 - std::pow((double)c.beta1, (double)t));
        a.c2 = (float)(1.0 - std::pow((double)c.beta2, (double)t));
        const bool split_in_adam = um && h->um.ready;
        if (split_in_adam) {     // the next step's K-major TF32 operand copies come straight out of the optimiser (same layouts as umma_split_weights)
            auto& u = h->um;
            a.sp[0][0] = SplitOut{u.Wt0[0], u.Wt0[1], (long)h->Kp, 1, 0, in};                     // W0ᵀ [H0, in]
            a.sp[0][1] = SplitOut{u.W0s[0], u.W0s[1], (long)h->H0p, 0, c.uP, 2 * c.D};            // W0[uP:uP+2D, :]
            a.sp[1][0] = SplitOut{u.Wt1[0], u.Wt1[1], (long)h->H0p, 1, 0, c.H0};                  // W1ᵀ [H1, H0]
            a.sp[1][1] = SplitOut{u.W1s[0], u.W1s[1], (long)h->H1p, 0, 0, c.H0};                  // W1  [H0, H1]
            a.nsp[0] = a.nsp[1] = 2;
        }
        RET(launch(h, "adam_dense", [&] { k_adam<<<h->num_sms * 2, 256, 0, h->stream>>>(a); }));
        h->um.dirty = !split_in_adam;
        h->step++;
    }
    return CTR_OK;
}

int zero_grads(ctr_handle* h) {
    CU(h, cudaMemsetAsync(h->Gflat, 0, h->Gflat_n * sizeof(float), h->stream));
    return CTR_OK;
}

// Placement of the ITEM_* tables under world > 1: sharded by row % world (large tables), or replicated on every
// rank.  cfg.reserved[1]: 0 = by size (> 32 MB shards), 1 = always shard, 2 = always replicate.  USER_FEAT is
// always replicated.
bool table_sharded(const ctr_handle* h, int which, int64_t nrows, int32_t width) {
    if (h->comm.world <= 1 || which == CTR_TABLE_USER_FEAT) return false;
    const int pol = h->cfg.reserved[1];
    if (pol == 1) return true;
    if (pol == 2) return false;
    return (size_t)nrows * round_up(width, 4) * sizeof(float) > kReplicateBytes;
}

// (re)allocates table `which` for nrows logical rows; decides the placement; leaves the rows zeroed
void table_free(ctr_handle* h, int which) {
    if (h->tab_vmm[which].live) { cudaStreamSynchronize(h->stream); vmm_free(&h->tab_vmm[which]); }
    else if (h->tab[which]) cudaFree(h->tab[which]);
    h->tab[which] = nullptr;
}
// device memory of a table: shareable VMM memory when the table is sharded (peers map it), cudaMalloc otherwise
int table_mem(ctr_handle* h, int which, bool shard, size_t bytes) {
    table_free(h, which);
    if (shard) {
        std::string err;
        if (!vmm_alloc(bytes, h->dev, &h->tab_vmm[which], &err)) return set_err(h, CTR_ENOMEM, "table %d (%zu bytes): %s", which, bytes, err.c_str());
        h->tab[which] = (float*)h->tab_vmm[which].ptr;
    } else CU(h, cudaMalloc(&h->tab[which], bytes));
    CU(h, cudaMemsetAsync(h->tab[which], 0, bytes, h->stream));
    return CTR_OK;
}
int table_alloc(ctr_handle* h, int which, int64_t nrows, int32_t width) {
    const long ld = round_up(width, 4);                       // 16-byte aligned rows for 128-bit loads
    const bool shard = table_sharded(h, which, nrows, width);
    const int64_t local = shard ? (nrows - h->comm.rank + h->comm.world - 1) / h->comm.world : nrows;
    const size_t bytes = (size_t)std::max<int64_t>(local, 1) * ld * sizeof(float);
    RET(table_mem(h, which, shard, bytes));
    h->tab_ld[which] = ld; h->tab_rows[which] = nrows; h->tab_local_rows[which] = local; h->tab_width[which] = width;
    h->tab_sharded[which] = shard;
    if (which == CTR_TABLE_ITEM_EMB) {
        h->comm.replicate = h->comm.world > 1 && !shard;
        if (h->hot_acc) { cudaFree(h->hot_acc); h->hot_acc = nullptr; h->hot_rows = 0; }
        for (float** p : {&h->emb_m, &h->emb_v}) if (*p) { cudaFree(*p); *p = nullptr; }       // a new table starts a new solver state
    }
    if (which != CTR_TABLE_USER_FEAT) h->tab_gen++;           // peers hold mappings of the old allocation
    return CTR_OK;
}

RowSrc idx_src(const ctr_handle* h, const int* d_user, const int* d_item, const int* d_hist, int B) {
    RowSrc r{};
    r.emb = h->tab[CTR_TABLE_ITEM_EMB]; r.lde = h->tab_ld[CTR_TABLE_ITEM_EMB];
    r.ufeat = h->tab[CTR_TABLE_USER_FEAT]; r.ldu = h->tab_ld[CTR_TABLE_USER_FEAT];
    r.ifeat = h->tab[CTR_TABLE_ITEM_FEAT]; r.ldi = h->tab_ld[CTR_TABLE_ITEM_FEAT];
    r.user_row = d_user; r.item_row = d_item; r.hist = d_hist;
    r.dense = 0; r.nvalid = B;
    r.n_emb = (int)std::min<int64_t>(h->tab_rows[CTR_TABLE_ITEM_EMB], 0x7fffffff);
    r.n_user = (int)std::min<int64_t>(h->tab_rows[CTR_TABLE_USER_FEAT], 0x7fffffff);
    r.n_ifeat = (int)std::min<int64_t>(h->tab_rows[CTR_TABLE_ITEM_FEAT], 0x7fffffff);
    r.world = 1;
    return r;
}

int check_tables(const ctr_handle* h) {
    const ctr_config& c = h->cfg;
    if (!h->tab[CTR_TABLE_ITEM_EMB] || h->tab_width[CTR_TABLE_ITEM_EMB] != c.D)
        return set_err(h, CTR_ESTATE, "ITEM_EMB table not uploaded with width D=%d", c.D);
    if (c.uP > 0 && (!h->tab[CTR_TABLE_USER_FEAT] || h->tab_width[CTR_TABLE_USER_FEAT] != c.uP))
        return set_err(h, CTR_ESTATE, "USER_FEAT table not uploaded with width uP=%d", c.uP);
    if (c.cF > 0 && (!h->tab[CTR_TABLE_ITEM_FEAT] || h->tab_width[CTR_TABLE_ITEM_FEAT] != c.cF))
        return set_err(h, CTR_ESTATE, "ITEM_FEAT table not uploaded with width cF=%d", c.cF);
    if (h->comm.world > 1 && (h->tab_sharded[CTR_TABLE_ITEM_EMB] || h->tab_sharded[CTR_TABLE_ITEM_FEAT]))
        return set_err(h, CTR_ESTATE, "this entry point gathers locally: ITEM_EMB / ITEM_FEAT must both be replicated (or both sharded for the train / predict entry points)");
    return CTR_OK;
}

int stage_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, const float* label, int B) {
    const int S = h->cfg.S;
    CU(h, cudaMemcpyAsync(h->s_user, user_row, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->s_item, item_row, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
    if (S > 0) CU(h, cudaMemcpyAsync(h->s_hist, hist, sizeof(int) * (size_t)B * S, cudaMemcpyHostToDevice, h->stream));
    if (label) CU(h, cudaMemcpyAsync(h->s_label, label, sizeof(float) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
    return CTR_OK;
}

int read_cost(ctr_handle* h, int B, float* cost) {
    double s = 0;
    CU(h, cudaMemcpyAsync(&s, h->d_cost, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    *cost = -(float)(s / (double)B);         // -mean, cost.go:15
    if (h->comm.h_err && *h->comm.h_err) return set_err(h, CTR_ECOMM, "peer barrier timed out waiting for rank %d (a rank left the step loop?)", *h->comm.h_err - 1);
    return CTR_OK;
}

int dense_src(ctr_handle* h, const int32_t ranges[8], int32_t xcols, RowSrc* out) {
    const ctr_config& c = h->cfg;
    const int up0 = ranges[0], up1 = ranges[1], ub0 = ranges[2], ub1 = ranges[3], it0 = ranges[4], it1 = ranges[5], cx0 = ranges[6], cx1 = ranges[7];
    if (up1 - up0 != c.uP || ub1 - ub0 != c.S * c.D || it1 - it0 != c.D || cx1 - cx0 != c.cF)
        return set_err(h, CTR_EINVAL, "SampleInfo ranges do not match the model dims (uP=%d S*D=%d D=%d cF=%d)", c.uP, c.S * c.D, c.D, c.cF);
    for (int i = 0; i < 8; i++) if (ranges[i] < 0 || ranges[i] > xcols) return set_err(h, CTR_EINVAL, "SampleInfo range outside [0, xcols=%d]", xcols);
    RowSrc r{}; r.dense = 1; r.world = 1; r.ldx = xcols; r.up0 = up0; r.ub0 = ub0; r.it0 = it0; r.cx0 = cx0;
    *out = r;
    return CTR_OK;
}

int upload_dense(ctr_handle* h, const float* X, const float* Y, int64_t n, int32_t xcols) {
    size_t need = (size_t)n * xcols;
    if (h->dXd_cap < need) { if (h->dXd) cudaFree(h->dXd); h->dXd = nullptr; CU(h, cudaMalloc(&h->dXd, need * sizeof(float))); h->dXd_cap = need; }
    CU(h, cudaMemcpyAsync(h->dXd, X, need * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (Y) {
        if (h->dYd_cap < (size_t)n) { if (h->dYd) cudaFree(h->dYd); h->dYd = nullptr; CU(h, cudaMalloc(&h->dYd, (size_t)n * sizeof(float))); h->dYd_cap = (size_t)n; }
        CU(h, cudaMemcpyAsync(h->dYd, Y, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    return CTR_OK;
}

void gaussian_fill(std::vector<float>& w, uint32_t seed, uint32_t stream) {
    for (size_t i = 0; i < w.size(); i++) {       // Box–Muller on the counter RNG
        uint64_t z = mix64(seed, stream, (uint64_t)i);
        double u1 = ((double)(z >> 40) + 1.0) * (1.0 / 16777217.0);
        double u2 = (double)((z >> 8) & 0xFFFFFFull) * (1.0 / 16777216.0);
        w[i] = (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
    }
}

}  // namespace

#include "comm_impl.cuh"

// item2vec host plan: Huffman tree as dictionary.HuffnamTree builds it (huffman.go:23-57) and every word's
// root→leaf path as Node.GetPath(maxDepth) returns it (node.go:26-43), as a CSR of (inner node, child code).
// dictionary.HuffnamTree (huffman.go:23-57) over leaves already in stable count order (`leaves` = word ids sorted by
// count, ties in id order): merged nodes are created in non-decreasing value order and inserted before every queued
// node of equal value, so they win ties and equal-valued merged nodes pop newest-first.  O(V).
// parent / code are indexed by tree node (leaves [0,V), inner nodes V + creation index); node_val [V-1].
// Ids that never occur (count 0) are not dictionary words in the reference (the dictionary is built from the stream,
// dictionary.go:70-81).  Here they go through the literal procedure like every other leaf (they chain into a comb of
// zero-valued nodes below the lightest real word, which leaves every real word's path what the procedure gives it), but
// no PATH is built for them — they are never trained and walking a comb as deep as their number is quadratic.
// Returns the number of inner nodes created.
static int i2v_huffman_sorted(const int64_t* cnt_by_word, const int* leaves, int V, std::vector<int>& parent, std::vector<unsigned char>& code,
                              std::vector<long long>& node_val) {
    parent.assign((size_t)2 * V - 1, -1); code.assign((size_t)2 * V - 1, 0); node_val.assign((size_t)std::max(V - 1, 1), 0);
    const size_t first = 0;
    const int Vp = V;
    // merged queue as runs of equal value; inside a run the newest node pops first, runs are in non-decreasing value order
    std::vector<long long> run_val; std::vector<int> run_begin, run_top;      // run r holds ids mq[run_begin[r] .. run_top[r])
    std::vector<int> mq; mq.reserve((size_t)V);
    size_t rh = 0, lh = first; int next_id = V;
    for (int made = 0; made < Vp - 1; made++) {
        long long pv[2]; int pid[2];
        for (int k = 0; k < 2; k++) {
            while (rh + 1 < run_val.size() && run_top[rh] == run_begin[rh]) rh++;          // never step past the last run: it may refill
            const bool have_m = rh < run_val.size() && run_top[rh] > run_begin[rh];
            const bool use_m = have_m && (lh >= (size_t)V || run_val[rh] <= cnt_by_word[leaves[lh]]);
            if (use_m) { pv[k] = run_val[rh]; pid[k] = mq[(size_t)--run_top[rh]]; if (rh + 1 == run_val.size()) mq.resize((size_t)run_top[rh]); }
            else { pv[k] = cnt_by_word[leaves[lh]]; pid[k] = leaves[lh]; lh++; }
        }
        const long long val = pv[0] + pv[1]; const int id = next_id++;
        node_val[(size_t)(id - V)] = val;
        code[(size_t)pid[0]] = 0; code[(size_t)pid[1]] = 1; parent[(size_t)pid[0]] = id; parent[(size_t)pid[1]] = id;
        if (!run_val.empty() && run_val.back() == val && run_top.back() == (int)mq.size()) { mq.push_back(id); run_top.back() = (int)mq.size(); }
        else if (!run_val.empty() && run_top.back() == run_begin.back() && run_top.back() == (int)mq.size()) {      // reuse the emptied last run
            run_val.back() = val; mq.push_back(id); run_top.back() = (int)mq.size();
        } else { run_val.push_back(val); run_begin.push_back((int)mq.size()); mq.push_back(id); run_top.push_back((int)mq.size()); }
    }
    return std::max(Vp - 1, 0);
}

// item2vec host plan (CPU-only entry ctr_i2v_paths): Huffman tree + every word's root→leaf path as Node.GetPath(maxDepth)
// returns it (node.go:26-43), as a CSR of (inner node, child code).
static void i2v_build_paths(const std::vector<int64_t>& cnt, int V, int max_depth, std::vector<int64_t>& node_val,
                            std::vector<long long>& poff, std::vector<int>& pnode, std::vector<unsigned char>& pcode) {
    std::vector<int> leaves((size_t)V);
    for (int i = 0; i < V; i++) leaves[(size_t)i] = i;
    std::stable_sort(leaves.begin(), leaves.end(), [&](int x, int y) { return cnt[(size_t)x] < cnt[(size_t)y]; });
    std::vector<int> parent; std::vector<unsigned char> code; std::vector<long long> nv;
    i2v_huffman_sorted(cnt.data(), leaves.data(), V, parent, code, nv);
    node_val.assign(nv.begin(), nv.begin() + std::max(V - 1, 0));
    poff.assign((size_t)V + 1, 0); pnode.clear(); pcode.clear();
    std::vector<int> chain;
    for (int w = 0; w < V; w++) {                          // Node.GetPath(maxDepth), node.go:26-43
        if (cnt[(size_t)w] == 0) { poff[(size_t)w + 1] = (long long)pnode.size(); continue; }     // never occurs: no path
        chain.clear();
        for (int p = w; p != -1; p = parent[(size_t)p]) chain.push_back(p);
        const int len = (int)chain.size(), depth = std::min(max_depth, len);
        for (int i = 0; i < depth - 1; i++) { pnode.push_back(chain[(size_t)(len - 1 - i)] - V); pcode.push_back(code[(size_t)chain[(size_t)(len - 2 - i)]]); }
        poff[(size_t)w + 1] = (long long)pnode.size();
    }
}

// =====================================================================================================
extern "C" {

int ctr_abi_version(void) { return CTR_B200_ABI_VERSION; }

void ctr_config_default(ctr_config* c, int model) {
    memset(c, 0, sizeof *c);
    c->model = model;
    c->uP = 52; c->S = 10; c->D = 16; c->cF = 53;          // example/movielens: feature.go:87-196, rcmd.go:22-24
    c->H0 = 200; c->H1 = 80;                                // din.go:17-18
    c->batch = 200; c->pred_batch = 100;                    // model_test.go:22, dinimpl_test.go
    c->lr = 0.01f; c->l2 = 1e-4f; c->beta1 = 0.9f; c->beta2 = 0.999f; c->eps = 1e-8f;   // model.go:88
    c->dropout0 = c->dropout1 = (model == CTR_MODEL_YOUTUBE) ? 0.003f : 0.005f;           // dnn.go:136-137, din.go:204-205
    c->seed = 0; c->table_opt = CTR_TABLE_FROZEN; c->table_lr = 0.0f; c->gemm = CTR_GEMM_AUTO;
    c->device = 0; c->rank = 0; c->world = 1;
}

const char* ctr_last_error(const ctr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ctr_create(const ctr_config* cfg, ctr_handle** out) {
    if (!cfg || !out) return set_err(nullptr, CTR_EINVAL, "null argument");
    *out = nullptr;
    const ctr_config& c = *cfg;
    if (c.model < 0 || c.model > 2) return set_err(nullptr, CTR_EINVAL, "unknown model %d", c.model);
    if (c.table_opt < CTR_TABLE_FROZEN || c.table_opt > CTR_TABLE_ADAM) return set_err(nullptr, CTR_EINVAL, "unknown table optimiser %d", c.table_opt);
    if (c.uP < 0 || c.cF < 0 || c.S < 1 || c.D < 1 || c.H0 < 1 || c.H1 < 1 || c.batch < 1 || c.pred_batch < 1)
        return set_err(nullptr, CTR_EINVAL, "bad dims");
    if (c.D > 32 * kGenAcc) return set_err(nullptr, CTR_EINVAL, "D=%d > %d unsupported", c.D, 32 * kGenAcc);
    if (c.H1 > 128) return set_err(nullptr, CTR_EINVAL, "H1=%d > 128 unsupported", c.H1);
    if (c.world < 1 || c.rank < 0 || c.rank >= c.world) return set_err(nullptr, CTR_EINVAL, "bad rank/world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_err(nullptr, CTR_ENODEV, "no CUDA device: this engine has no CPU fallback"); }
    if (c.device < 0 || c.device >= ndev) return set_err(nullptr, CTR_ENODEV, "device %d of %d", c.device, ndev);
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, c.device);
    if (prop.major != 10) return set_err(nullptr, CTR_ENODEV, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);

    ctr_handle* h = new ctr_handle();
    h->cfg = c; h->dev = c.device; h->num_sms = prop.multiProcessorCount;
    h->comm.rank = c.rank; h->comm.world = c.world;
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(nullptr, CTR_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); ctr_destroy(h); return CTR_ECUDA; } } while (0)
#define RC(x) do { int r_ = (x); if (r_ != CTR_OK) { g_create_error = h->err; ctr_destroy(h); return r_; } } while (0)
    CUC(cudaSetDevice(h->dev));
    CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true;
    CUC(cudaEventCreate(&h->ev0)); CUC(cudaEventCreate(&h->ev1));
    h->in = c.uP + 2 * c.D + c.cF;                      // din.go:187
    h->Kp = round_up(h->in, 32); h->H0p = round_up(c.H0, 32); h->H1p = round_up(c.H1, 32); h->Sp = round_up(c.S, 4);
    h->lddx = round_up(2 * c.D, 4);
    h->Bmax = std::max(c.batch, c.pred_batch);
    h->wsize[0] = (size_t)h->Kp * h->H0p; h->wsize[1] = (size_t)h->H0p * h->H1p; h->wsize[2] = h->H1p; h->wsize[3] = h->Sp;
    h->Gflat_n = h->wsize[0] + h->wsize[1] + h->wsize[2] + h->wsize[3];
    RC(dalloc(h, &h->Gflat, h->Gflat_n));
    for (size_t i = 0, off = 0; i < 4; off += h->wsize[i], i++) h->G[i] = h->Gflat + off;
    for (int i = 0; i < 4; i++) { RC(dalloc(h, &h->W[i], h->wsize[i])); RC(dalloc(h, &h->Mo[i], h->wsize[i])); RC(dalloc(h, &h->Vo[i], h->wsize[i])); }
    const size_t B = h->Bmax;
    RC(dalloc(h, &h->X0, B * h->Kp)); RC(dalloc(h, &h->H0d, B * h->H0p)); RC(dalloc(h, &h->H1d, B * h->H1p));
    RC(dalloc(h, &h->P, B)); RC(dalloc(h, &h->Z, B));
    RC(dalloc(h, &h->dZ1, B * h->H1p)); RC(dalloc(h, &h->dZ0, B * h->H0p)); RC(dalloc(h, &h->dX, B * h->lddx));
    RC(dalloc(h, &h->d_cost, 1));
    RC(dalloc(h, &h->s_user, B)); RC(dalloc(h, &h->s_item, B)); RC(dalloc(h, &h->s_hist, B * c.S)); RC(dalloc(h, &h->s_label, B));
    RC(ctr_init_weights(h, c.seed));
    CUC(cudaStreamSynchronize(h->stream));
#undef CUC
#undef RC
    *out = h;
    return CTR_OK;
}

void ctr_destroy(ctr_handle* h) {
    if (!h) return;
    cudaSetDevice(h->dev);
    if (h->stream) cudaStreamSynchronize(h->stream);
    comm_destroy(h);
    for (int i = 0; i < 3; i++) table_free(h, i);
    for (void* p : {(void*)h->ub_off, (void*)h->ub_ts, (void*)h->ub_items}) if (p) cudaFree(p);
    for (void* p : {(void*)h->idm_keys[0], (void*)h->idm_keys[1], (void*)h->idm_vals[0], (void*)h->idm_vals[1], (void*)h->k_keys, (void*)h->k_flags}) if (p) cudaFree(p);
    if (h->k_host) cudaFreeHost(h->k_host);
    for (int i = 0; i < 2; i++) for (float* p : {h->um.Wt0[i], h->um.Wt1[i], h->um.W1s[i], h->um.W0s[i]}) if (p) cudaFree(p);
    for (int i = 0; i < 4; i++) for (float* p : {h->W[i], h->Mo[i], h->Vo[i]}) if (p) cudaFree(p);
    if (h->Gflat) cudaFree(h->Gflat);
    for (void* p : {(void*)h->X0, (void*)h->H0d, (void*)h->H1d, (void*)h->P, (void*)h->Z, (void*)h->dZ1, (void*)h->dZ0, (void*)h->dX,
                    (void*)h->dUb, (void*)h->dIt, (void*)h->keys, (void*)h->keys2, (void*)h->pos, (void*)h->pos2, h->sort_tmp,
                    (void*)h->d_cost, (void*)h->s_user, (void*)h->s_item, (void*)h->s_hist, (void*)h->s_label, (void*)h->dXd, (void*)h->dYd, (void*)h->hot_acc, (void*)h->emb_m, (void*)h->emb_v})
        if (p) cudaFree(p);
    for (int i = 0; i < 3; i++) if (h->scratch[i]) cudaFree(h->scratch[i]);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < 2; i++) { if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]); }
    for (void* p : {(void*)h->feed_dev[0], (void*)h->feed_dev[1], (void*)h->d_costs, (void*)h->kt_user, (void*)h->kt_item, (void*)h->kt_ts, (void*)h->kt_label,
                    (void*)h->kt_flag, (void*)h->kt_pos, h->kt_scan_tmp, (void*)h->kt_count}) if (p) cudaFree(p);
    for (int i = 0; i < ctr_handle::kPin; i++) { if (h->feed_pin[i]) cudaFreeHost(h->feed_pin[i]); if (h->pin_free[i]) cudaEventDestroy(h->pin_free[i]); }
    if (h->kt_count_host) cudaFreeHost(h->kt_count_host);
    for (int i = 0; i < 2; i++) { if (h->kt_hist[i]) cudaFree(h->kt_hist[i]); if (h->kt_win[i]) cudaEventDestroy(h->kt_win[i]); if (h->kt_used[i]) cudaEventDestroy(h->kt_used[i]); }
    if (h->win_stream) cudaStreamDestroy(h->win_stream);
    for (int i = 0; i < ctr_handle::kCnt; i++) if (h->kt_counted[i]) cudaEventDestroy(h->kt_counted[i]);
    delete h->pool; h->pool = nullptr;
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int ctr_set_weights(ctr_handle* h, const float* mlp0, const float* mlp1, const float* mlp2, const float* att0) {
    if (!h || !mlp0 || !mlp1 || !mlp2) return set_err(h, CTR_EINVAL, "null weights");
    std::lock_guard<std::mutex> lk(h->mu);
    const ctr_config& c = h->cfg;
    CU(h, cudaSetDevice(h->dev));
    for (int i = 0; i < 4; i++) {
        CU(h, cudaMemsetAsync(h->W[i], 0, h->wsize[i] * sizeof(float), h->stream));
        CU(h, cudaMemsetAsync(h->G[i], 0, h->wsize[i] * sizeof(float), h->stream));
        CU(h, cudaMemsetAsync(h->Mo[i], 0, h->wsize[i] * sizeof(float), h->stream));
        CU(h, cudaMemsetAsync(h->Vo[i], 0, h->wsize[i] * sizeof(float), h->stream));
    }
    CU(h, cudaMemcpy2DAsync(h->W[0], h->H0p * sizeof(float), mlp0, c.H0 * sizeof(float), c.H0 * sizeof(float), h->in, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpy2DAsync(h->W[1], h->H1p * sizeof(float), mlp1, c.H1 * sizeof(float), c.H1 * sizeof(float), c.H0, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->W[2], mlp2, c.H1 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (att0) CU(h, cudaMemcpyAsync(h->W[3], att0, c.S * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    else { std::vector<float> ones(c.S, 1.0f); CU(h, cudaMemcpyAsync(h->W[3], ones.data(), c.S * sizeof(float), cudaMemcpyHostToDevice, h->stream)); CU(h, cudaStreamSynchronize(h->stream)); }
    CU(h, cudaStreamSynchronize(h->stream));
    h->step = 0;
    h->um.dirty = true;
    return CTR_OK;
}

int ctr_init_weights(ctr_handle* h, uint32_t seed) {
    if (!h) return CTR_EINVAL;
    const ctr_config& c = h->cfg;
    std::vector<float> w0((size_t)h->in * c.H0), w1((size_t)c.H0 * c.H1), w2(c.H1), a(c.S, 1.0f);   // att0 = 1: din.go:181
    gaussian_fill(w0, seed, 0); gaussian_fill(w1, seed, 1); gaussian_fill(w2, seed, 2);
    return ctr_set_weights(h, w0.data(), w1.data(), w2.data(), a.data());
}

int ctr_get_weights(ctr_handle* h, float* mlp0, float* mlp1, float* mlp2, float* att0) {
    if (!h) return CTR_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    const ctr_config& c = h->cfg;
    CU(h, cudaSetDevice(h->dev));
    if (mlp0) CU(h, cudaMemcpy2DAsync(mlp0, c.H0 * sizeof(float), h->W[0], h->H0p * sizeof(float), c.H0 * sizeof(float), h->in, cudaMemcpyDeviceToHost, h->stream));
    if (mlp1) CU(h, cudaMemcpy2DAsync(mlp1, c.H1 * sizeof(float), h->W[1], h->H1p * sizeof(float), c.H1 * sizeof(float), c.H0, cudaMemcpyDeviceToHost, h->stream));
    if (mlp2) CU(h, cudaMemcpyAsync(mlp2, h->W[2], c.H1 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (att0) CU(h, cudaMemcpyAsync(att0, h->W[3], c.S * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

int ctr_table_upload(ctr_handle* h, int which, const float* rows, int64_t nrows, int32_t width) {
    if (!h || !rows || which < 0 || which > 2 || nrows < 1 || width < 1) return set_err(h, CTR_EINVAL, "bad table upload arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    const ctr_config& c = h->cfg;
    const int want = which == CTR_TABLE_USER_FEAT ? c.uP : which == CTR_TABLE_ITEM_FEAT ? c.cF : c.D;
    if (width != want) return set_err(h, CTR_EINVAL, "table %d width %d != model dim %d", which, width, want);
    CU(h, cudaSetDevice(h->dev));
    RET(table_alloc(h, which, nrows, width));
    const long ld = h->tab_ld[which]; const int64_t local = h->tab_local_rows[which];
    if (!h->tab_sharded[which]) {
        CU(h, cudaMemcpy2DAsync(h->tab[which], ld * sizeof(float), rows, (size_t)width * sizeof(float), (size_t)width * sizeof(float), (size_t)nrows, cudaMemcpyHostToDevice, h->stream));
    } else if (local > 0) {   // owner(row) = row % world; local row = row / world
        CU(h, cudaMemcpy2DAsync(h->tab[which], ld * sizeof(float), rows + (size_t)h->comm.rank * width, (size_t)width * h->comm.world * sizeof(float),
                                (size_t)width * sizeof(float), (size_t)local, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

int ctr_table_fill(ctr_handle* h, int which, int64_t nrows, int32_t width, uint32_t seed, int32_t dist, float scale) {
    if (!h || which < 0 || which > 2 || nrows < 1 || width < 1) return set_err(h, CTR_EINVAL, "bad table fill arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    const ctr_config& c = h->cfg;
    const int want = which == CTR_TABLE_USER_FEAT ? c.uP : which == CTR_TABLE_ITEM_FEAT ? c.cF : c.D;
    if (width != want) return set_err(h, CTR_EINVAL, "table %d width %d != model dim %d", which, width, want);
    CU(h, cudaSetDevice(h->dev));
    RET(table_alloc(h, which, nrows, width));
    const bool shard = h->tab_sharded[which];
    RET(launch(h, "table_fill", [&] {
        k_table_fill<<<h->num_sms * 8, 256, 0, h->stream>>>(h->tab[which], h->tab_ld[which], (long)h->tab_local_rows[which], width, shard ? h->comm.world : 1,
                                                           shard ? h->comm.rank : 0, seed, (uint32_t)which, dist, scale);
    }));
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

int ctr_table_download(ctr_handle* h, int which, float* rows, int64_t nrows, int32_t width) {
    if (!h || !rows || which < 0 || which > 2) return set_err(h, CTR_EINVAL, "bad table download arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->tab[which] || nrows != h->tab_rows[which] || width != h->tab_width[which]) return set_err(h, CTR_EINVAL, "table %d shape mismatch", which);
    CU(h, cudaSetDevice(h->dev));
    const bool shard = h->tab_sharded[which];
    if (shard && which == CTR_TABLE_ITEM_EMB) RET(comm_hot_writeback(h));
    if (!shard) {
        CU(h, cudaMemcpy2DAsync(rows, (size_t)width * sizeof(float), h->tab[which], h->tab_ld[which] * sizeof(float), (size_t)width * sizeof(float), (size_t)nrows, cudaMemcpyDeviceToHost, h->stream));
    } else if (h->tab_local_rows[which] > 0) {   // fills only this rank's rows (row % world == rank)
        CU(h, cudaMemcpy2DAsync(rows + (size_t)h->comm.rank * width, (size_t)width * h->comm.world * sizeof(float), h->tab[which], h->tab_ld[which] * sizeof(float),
                                (size_t)width * sizeof(float), (size_t)h->tab_local_rows[which], cudaMemcpyDeviceToHost, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

int ctr_gather_rows(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, int64_t B, float* X) {
    if (!h || !user_row || !item_row || !hist || !X || B < 1) return set_err(h, CTR_EINVAL, "bad gather arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    RET(check_tables(h));
    CU(h, cudaSetDevice(h->dev));
    const ctr_config& c = h->cfg;
    const long xc = (long)c.uP + (long)c.S * c.D + c.D + c.cF;
    float* dout = nullptr;
    const int chunk = h->Bmax;
    RET(scratch_get(h, 0, (size_t)chunk * xc * sizeof(float), (void**)&dout));
    int rc = CTR_OK;
    for (int64_t s = 0; s < B && rc == CTR_OK; s += chunk) {
        int nb = (int)std::min<int64_t>(chunk, B - s);
        rc = stage_idx(h, user_row + s, item_row + s, hist + s * c.S, nullptr, nb);
        if (rc != CTR_OK) break;
        RowSrc r = idx_src(h, h->s_user, h->s_item, h->s_hist, nb);
        rc = launch(h, "gather_rows", [&] { k_gather_rows<<<grid_for_warps(h, nb), 256, 0, h->stream>>>(r, dims_of(h), dout, xc, nb); });
        if (rc != CTR_OK) break;
        if (cudaMemcpyAsync(X + s * xc, dout, (size_t)nb * xc * sizeof(float), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) rc = set_err(h, CTR_ECUDA, "gather D2H: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

int ctr_train_dense(ctr_handle* h, const float* X, const float* Y, int64_t n, int32_t xcols, const int32_t ranges[8],
                    int32_t epochs, int32_t early_stop, float* last_cost, int32_t* epochs_run) {
    if (!h || !X || !Y || !ranges || n < 1 || xcols < 1 || epochs < 0) return set_err(h, CTR_EINVAL, "bad train arguments");
    // the dense-X compatibility route has no gradient exchange: refusing beats training world replicas that silently diverge
    if (h->comm.world > 1) return set_err(h, CTR_ESTATE, "ctr_train_dense is single-GPU; with world=%d train through the index entry points", h->comm.world);
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    RowSrc base{};
    RET(dense_src(h, ranges, xcols, &base));
    RET(upload_dense(h, X, Y, n, xcols));
    const int B = h->cfg.batch;
    const int64_t batches = n / B + (n % B != 0);           // model.go:96-99
    float best = INFINITY, cost = 0.0f; int no_improve = 0, ep = 0;
    for (ep = 0; ep < epochs; ep++) {
        for (int64_t b = 0; b < batches; b++) {
            const int64_t start = b * B; int64_t end = start + B;
            if (start >= n) break;
            if (end > n) end = n;
            RowSrc r = base; r.X = h->dXd + start * xcols; r.nvalid = (int)(end - start);   // tail rows → zeros, label 0 (model.go:357-371)
            StepOpts o; o.training = true; o.update = true; o.d_label = h->dYd + start;
            RET(step_core(h, r, B, o));
        }
        RET(read_cost(h, B, &cost));                        // cost of the epoch's last batch, model.go:198
        if (cost < best) { best = cost; no_improve = 0; } else no_improve++;
        if (early_stop != 0 && no_improve >= early_stop) { ep++; break; }   // model.go:206-209
    }
    if (last_cost) *last_cost = cost;
    if (epochs_run) *epochs_run = ep;
    return CTR_OK;
}

int ctr_predict_dense(ctr_handle* h, const float* X, int64_t n, int32_t xcols, const int32_t ranges[8], float* out) {
    if (!h || !X || !ranges || !out || n < 1 || xcols < 1) return set_err(h, CTR_EINVAL, "bad predict arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    RowSrc base{};
    RET(dense_src(h, ranges, xcols, &base));
    RET(upload_dense(h, X, nullptr, n, xcols));
    const int B = h->cfg.pred_batch;
    const int64_t batches = n / B + (n % B != 0);
    for (int64_t b = 0; b < batches; b++) {
        const int64_t start = b * B; int64_t end = start + B;
        if (start >= n) break;
        if (end > n) end = n;
        RowSrc r = base; r.X = h->dXd + start * xcols; r.nvalid = (int)(end - start);
        StepOpts o;                                          // dropout off: din.go:133-145
        RET(step_core(h, r, B, o));
        CU(h, cudaMemcpyAsync(out + start, h->P, (size_t)(end - start) * sizeof(float), cudaMemcpyDeviceToHost, h->stream));   // model.go:344-347
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

namespace {
// options of one local-gather train step; with world > 1 (replicated ITEM_EMB) the row gradients go through
// comm.table_grad and everything is all-reduced before the update
int train_opts(ctr_handle* h, const float* d_label, int B, StepOpts* out) {
    StepOpts o; o.training = true; o.update = true; o.d_label = d_label;
    if (h->comm.world > 1) {
        Comm& cm = h->comm;
        const size_t n = (size_t)h->tab_rows[CTR_TABLE_ITEM_EMB] * h->tab_ld[CTR_TABLE_ITEM_EMB];
        if (cm.table_grad_n != n) {
            if (cm.table_grad) cudaFree(cm.table_grad);
            cm.table_grad = nullptr; cm.table_grad_n = 0;
            RET(dalloc(h, &cm.table_grad, n)); cm.table_grad_n = n;
        }
        o.table_grad = cm.table_grad; o.table_grad_n = n; o.grad_scale = 1.0f / (float)cm.world; o.adam_batch = B * cm.world;
    }
    *out = o;
    return CTR_OK;
}

bool emb_sharded(const ctr_handle* h) { return h->comm.world > 1 && h->tab_sharded[CTR_TABLE_ITEM_EMB]; }

// one train step on device-resident ids; rows >= nvalid are the zero-padded tail trained as label 0 (model.go:357-371)
int train_step_dev(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, const float* d_label, int B, int nvalid) {
    if (emb_sharded(h)) return comm_train_step(h, d_user, d_item, d_hist, d_label, B, nvalid);
    RET(check_tables(h));
    RowSrc r = idx_src(h, d_user, d_item, d_hist, B);
    r.nvalid = nvalid;
    StepOpts o;
    RET(train_opts(h, d_label, B, &o));
    return step_core(h, r, B, o);
}
}  // namespace

int ctr_train_step_idx_dev(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, const float* d_label, int32_t B) {
    if (!h || !d_user || !d_item || !d_hist || !d_label) return set_err(h, CTR_EINVAL, "null argument");
    if (B != h->cfg.batch) return set_err(h, CTR_EINVAL, "B=%d != configured batch %d", B, h->cfg.batch);
    return train_step_dev(h, d_user, d_item, d_hist, d_label, B, B);
}

int ctr_train_step_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, const float* label, int32_t B, ctr_step_stats* stats) {
    if (!h || !user_row || !item_row || !hist || !label) return set_err(h, CTR_EINVAL, "null argument");
    if (B != h->cfg.batch) return set_err(h, CTR_EINVAL, "B=%d != configured batch %d", B, h->cfg.batch);
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    const int64_t l0 = h->launches;
    RET(stage_idx(h, user_row, item_row, hist, label, B));
    RET(ctr_train_step_idx_dev(h, h->s_user, h->s_item, h->s_hist, h->s_label, B));
    float cost = 0;
    RET(read_cost(h, B * h->comm.world, &cost));      // d_cost was summed over the ranks
    if (stats) { stats->cost = cost; stats->ms_device = 0; stats->launches = (int32_t)(h->launches - l0); stats->reserved = 0; }
    return CTR_OK;
}

namespace {
// lazily creates the pinned ring, the two device slots and the copy stream; slot size fits one idx batch
int feed_init(ctr_handle* h) {
    if (h->copy_stream) return CTR_OK;
    const size_t B = (size_t)h->Bmax;
    h->feed_bytes = (B * (size_t)(h->cfg.S + 3) * 4 + 255) & ~(size_t)255;      // [user | item | hist | label] — also fits [uid | iid | ts | label] chunks
    h->feed_bytes = std::max(h->feed_bytes, (size_t)B * 28 + 256);
    CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU(h, cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming)); CU(h, cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
        CU(h, cudaMalloc(&h->feed_dev[i], h->feed_bytes));
    }
    for (int i = 0; i < ctr_handle::kPin; i++) {
        CU(h, cudaHostAlloc(&h->feed_pin[i], h->feed_bytes, cudaHostAllocDefault));
        CU(h, cudaEventCreateWithFlags(&h->pin_free[i], cudaEventDisableTiming));
    }
    h->pool = new CopyPool(7);          // a host thread copies ~5 GB/s on the GPU boxes; 8 streams keep a 14 MB batch under the step time
    return CTR_OK;
}
int costs_reserve(ctr_handle* h, size_t nb) {
    if (h->d_costs_cap >= nb) return CTR_OK;
    if (h->d_costs) cudaFree(h->d_costs);
    h->d_costs = nullptr; RET(dalloc(h, &h->d_costs, nb)); h->d_costs_cap = nb;
    return CTR_OK;
}
}  // namespace

int ctr_train_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, const float* label,
                  int64_t n, float* costs) {
    if (!h || !user_row || !item_row || !hist || !label || n < 1) return set_err(h, CTR_EINVAL, "bad train arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    if (!emb_sharded(h)) RET(check_tables(h));
    const int B = h->cfg.batch, S = h->cfg.S;
    const int64_t nb = (n + B - 1) / B;                      // model.go:96-99
    RET(feed_init(h));
    RET(costs_reserve(h, (size_t)nb));
    const size_t o_item = (size_t)B * 4, o_hist = (size_t)B * 8, o_label = (size_t)B * (S + 2) * 4;
    CU(h, cudaStreamSynchronize(h->stream));                 // a device slot may still be in use by an earlier call
    for (int64_t b = 0; b < nb; b++) {
        const int slot = (int)(b & 1), ps = (int)(b % ctr_handle::kPin);
        const int64_t start = b * B; const size_t nv = (size_t)std::min<int64_t>(B, n - start);
        // caller memory → pinned slot (host threads; overlaps the GPU work already queued), then one DMA
        if (b >= ctr_handle::kPin) CU(h, cudaEventSynchronize(h->pin_free[ps]));
        unsigned char* pin = h->feed_pin[ps];
        memcpy(pin, user_row + start, nv * 4); memcpy(pin + o_item, item_row + start, nv * 4); memcpy(pin + o_label, label + start, nv * 4);
        h->pool->copy(pin + o_hist, hist + start * S, nv * S * 4);
        // H2D of batch b on the copy stream, after the compute that last used this device slot (batch b-2)
        if (b >= 2) CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0));
        CU(h, cudaMemcpyAsync(h->feed_dev[slot], pin, o_label + (size_t)B * 4, cudaMemcpyHostToDevice, h->copy_stream));
        CU(h, cudaEventRecord(h->pin_free[ps], h->copy_stream));
        CU(h, cudaEventRecord(h->ev_copied[slot], h->copy_stream));
        // compute of batch b on the engine stream (overlaps the staging + H2D of batch b+1)
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
        unsigned char* d = h->feed_dev[slot];
        RET(train_step_dev(h, (const int*)d, (const int*)(d + o_item), (const int*)(d + o_hist), (const float*)(d + o_label), B, (int)nv));   // ragged tail → zero rows with label 0 (model.go:357-371)
        CU(h, cudaMemcpyAsync(h->d_costs + b, h->d_cost, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CU(h, cudaEventRecord(h->ev_consumed[slot], h->stream));
    }
    std::vector<double> hc((size_t)nb);
    CU(h, cudaMemcpyAsync(hc.data(), h->d_costs, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaStreamSynchronize(h->copy_stream));
    if (h->comm.h_err && *h->comm.h_err) return set_err(h, CTR_ECOMM, "peer barrier timed out waiting for rank %d", *h->comm.h_err - 1);
    if (costs) for (int64_t b = 0; b < nb; b++) costs[b] = -(float)(hc[(size_t)b] / ((double)B * h->comm.world));   // d_cost is summed over the ranks
    return CTR_OK;
}

// recommend.Train (rcmd.go:197-246) fed by sample keys: GetSample's assembly (rcmd.go:339-460) happens on the device —
// id maps resolve {UserId, ItemId}, samples whose user or item has no features are DROPPED as the reference's
// assembler skips them (rcmd.go:378-382), the survivors stay resident in HBM in input order, and every batch's
// history rows come from the device ubcache at the sample's timestamp (GetUserBehavior(uid, S, -1, ts),
// rcmd.go:509; prepare.go:13-38) right before its step — then model.Train's epoch loop (model.go:96-209).
int ctr_train_keys(ctr_handle* h, const int64_t* user_ids, const int64_t* item_ids, const int64_t* ts, const float* label, int64_t n,
                   int32_t epochs, int32_t early_stop, float* last_cost, int32_t* epochs_run, int64_t* rows_used) {
    if (!h || !user_ids || !item_ids || !ts || !label || n < 1 || epochs < 0) return set_err(h, CTR_EINVAL, "bad train arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    if (!h->idm_keys[0] || !h->idm_keys[1]) return set_err(h, CTR_ESTATE, "ctr_train_keys needs both id maps (ctr_idmap_build)");
    if (!emb_sharded(h)) RET(check_tables(h));
    const int B = h->cfg.batch, S = h->cfg.S;
    RET(feed_init(h));
    // ---- resident sample arrays (+B slack so a padded tail never indexes past the end)
    const size_t cap = (size_t)n + (size_t)B;
    if (h->kt_cap < cap) {
        for (void* p : {(void*)h->kt_user, (void*)h->kt_item, (void*)h->kt_ts, (void*)h->kt_label}) if (p) cudaFree(p);
        h->kt_user = h->kt_item = nullptr; h->kt_ts = nullptr; h->kt_label = nullptr; h->kt_cap = 0;
        RET(dalloc(h, &h->kt_user, cap)); RET(dalloc(h, &h->kt_item, cap)); RET(dalloc(h, &h->kt_ts, cap)); RET(dalloc(h, &h->kt_label, cap));
        h->kt_cap = cap;
    }
    // One GPU: the first epoch STREAMS — small chunks, and the steps over the samples already resident are queued while
    // the host stages the next chunk (the survivor count of chunk c-1 comes back through a pinned ring).  Several ranks:
    // every step is collective, so the ranks first agree on the batch count (resolve everything, then train).
    const bool streaming = epochs > 0 && h->comm.world == 1;
    const size_t chunk = std::min<size_t>((size_t)n, streaming ? std::min<size_t>(h->feed_bytes / 28, (size_t)2 * B) : h->feed_bytes / 28);   // keys per staging slot
    if (h->kt_chunk < chunk) {
        for (void* p : {(void*)h->kt_flag, (void*)h->kt_pos, h->kt_scan_tmp}) if (p) cudaFree(p);
        h->kt_flag = h->kt_pos = nullptr; h->kt_scan_tmp = nullptr;
        RET(dalloc(h, &h->kt_flag, chunk + 1)); RET(dalloc(h, &h->kt_pos, chunk + 1));
        cub::DeviceScan::ExclusiveSum(nullptr, h->kt_scan_bytes, h->kt_flag, h->kt_pos, (int)(chunk + 1), h->stream);
        CU(h, cudaMalloc(&h->kt_scan_tmp, h->kt_scan_bytes));
        h->kt_chunk = chunk;
    }
    if (!h->kt_count) RET(dalloc(h, &h->kt_count, 1));
    if (!h->kt_count_host) {
        CU(h, cudaHostAlloc(&h->kt_count_host, sizeof(unsigned long long) * ctr_handle::kCnt, cudaHostAllocDefault));
        for (int i = 0; i < ctr_handle::kCnt; i++) CU(h, cudaEventCreateWithFlags(&h->kt_counted[i], cudaEventDisableTiming));
    }
    CU(h, cudaMemsetAsync(h->kt_count, 0, sizeof(unsigned long long), h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (!h->win_stream) {
        CU(h, cudaStreamCreateWithFlags(&h->win_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            RET(dalloc(h, &h->kt_hist[i], (size_t)h->Bmax * S, false));
            CU(h, cudaEventCreateWithFlags(&h->kt_win[i], cudaEventDisableTiming)); CU(h, cudaEventCreateWithFlags(&h->kt_used[i], cudaEventDisableTiming));
        }
    }
    // one step of model.Train's loop over the resident samples [b*B, b*B + B) of which `have` exist so far.  The batch's
    // history windows (GetUserBehavior at the sample's timestamp) are cut on the window stream into one of two buffers,
    // so the window kernel of batch b+1 runs under the train step of batch b.
    int64_t nrun = 0;
    auto run_batch = [&](int64_t b, int64_t have) -> int {
        const int64_t start = std::min<int64_t>(b * B, have);
        const int nv = (int)std::max<int64_t>(0, std::min<int64_t>(B, have - start));
        const int k = (int)(nrun & 1);
        int* hist = h->kt_hist[k];
        if (nv > 0) {
            if (nrun >= 2) CU(h, cudaStreamWaitEvent(h->win_stream, h->kt_used[k], 0));         // the step two batches back has read this buffer
            if (h->ub_off) {
                RET(launch(h, "ubcache_window", [&] {
                    k_ub_window<<<grid_for_warps(h, nv), 256, 0, h->win_stream>>>(h->ub_off, h->ub_ts, h->ub_items, h->kt_user + start, h->kt_ts + start, nv, S,
                                                                                 (long)h->ub_users, hist);
                }));
            } else CU(h, cudaMemsetAsync(hist, 0xff, sizeof(int) * (size_t)nv * S, h->win_stream));   // no UserBehavior provider → empty history (rcmd.go:498,509)
            CU(h, cudaEventRecord(h->kt_win[k], h->win_stream));
            CU(h, cudaStreamWaitEvent(h->stream, h->kt_win[k], 0));
        }
        RET(train_step_dev(h, h->kt_user + start, h->kt_item + start, hist, h->kt_label + start, B, nv));
        CU(h, cudaEventRecord(h->kt_used[k], h->stream));
        nrun++;
        return CTR_OK;
    };
    // ---- keys → (user row, item row, ts, label), unresolvable samples dropped, order kept.  The resolve kernels of a
    // chunk follow its H2D on the COPY stream, so they (and the staging of the next chunk) overlap the train steps
    // running on the engine stream; the engine stream waits for a chunk's event before it reads that chunk's samples.
    cudaStream_t rs = h->copy_stream;
    int64_t nchunks = ((int64_t)n + (int64_t)chunk - 1) / (int64_t)chunk;
    int64_t queued = 0;                                       // first-epoch batches already on the engine stream (streaming)
    for (int64_t c = 0; c < nchunks; c++) {
        const int slot = (int)(c & 1), ps = (int)(c % ctr_handle::kPin);
        const int64_t start = c * (int64_t)chunk; const size_t m = (size_t)std::min<int64_t>((int64_t)chunk, n - start);
        if (c >= ctr_handle::kPin) CU(h, cudaEventSynchronize(h->pin_free[ps]));
        unsigned char* pin = h->feed_pin[ps];
        h->pool->copy(pin, user_ids + start, m * 8); h->pool->copy(pin + m * 8, item_ids + start, m * 8);
        h->pool->copy(pin + m * 16, ts + start, m * 8); h->pool->copy(pin + m * 24, label + start, m * 4);
        CU(h, cudaMemcpyAsync(h->feed_dev[slot], pin, m * 28, cudaMemcpyHostToDevice, rs));     // after resolve(c-2), which read this slot (stream order)
        CU(h, cudaEventRecord(h->pin_free[ps], rs));
        const long long* dk = (const long long*)h->feed_dev[slot];
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((m + 255) / 256, (size_t)h->num_sms * 8));
        RET(launch(h, "keys_rows", [&] {
            k_keys_rows<<<grid, 256, 0, rs>>>(h->idm_keys[0], h->idm_vals[0], h->idm_cap[0] - 1, h->idm_keys[1], h->idm_vals[1], h->idm_cap[1] - 1,
                                             dk, dk + m, (long)m, h->kt_flag);
        }));
        RET(launch(h, "cub_exclusive_scan", [&] { cub::DeviceScan::ExclusiveSum(h->kt_scan_tmp, h->kt_scan_bytes, h->kt_flag, h->kt_pos, (int)(m + 1), rs); }));
        RET(launch(h, "keys_compact", [&] {
            k_keys_compact<<<grid, 256, 0, rs>>>(h->idm_keys[0], h->idm_vals[0], h->idm_cap[0] - 1, h->idm_keys[1], h->idm_vals[1], h->idm_cap[1] - 1,
                                                dk, dk + m, dk + 2 * m, (const float*)(dk + 3 * m), (long)m, h->kt_flag, h->kt_pos, h->kt_count,
                                                h->kt_user, h->kt_item, h->kt_ts, h->kt_label);
        }));
        RET(launch(h, "keys_advance", [&] { k_keys_advance<<<1, 1, 0, rs>>>(h->kt_pos + m, h->kt_count); }));
        if (streaming) {
            const int cs = (int)(c % ctr_handle::kCnt);
            CU(h, cudaMemcpyAsync(h->kt_count_host + cs, h->kt_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, rs));
            CU(h, cudaEventRecord(h->kt_counted[cs], rs));
            if (c >= 1) {       // survivors up to chunk c-1: queue their full batches while chunk c resolves
                const int pc = (int)((c - 1) % ctr_handle::kCnt);
                CU(h, cudaEventSynchronize(h->kt_counted[pc]));
                const int64_t have = (int64_t)h->kt_count_host[pc];
                if ((queued + 1) * (int64_t)B <= have) CU(h, cudaStreamWaitEvent(h->stream, h->kt_counted[pc], 0));
                while ((queued + 1) * (int64_t)B <= have) { RET(run_batch(queued, have)); queued++; }
            }
        }
    }
    CU(h, cudaStreamSynchronize(rs));                          // every sample resolved; the engine stream may still be training
    unsigned long long used = 0;
    CU(h, cudaMemcpyAsync(&used, h->kt_count, sizeof used, cudaMemcpyDeviceToHost, rs));
    CU(h, cudaStreamSynchronize(rs));
    if (rows_used) *rows_used = (int64_t)used;
    // ---- model.Train's loop over the resident samples.  world > 1: every step is collective, so all ranks run the
    // batch count of the rank with the most samples (ranks that ran out train on all-padding batches)
    int64_t nmax = (int64_t)used;
    if (h->comm.world > 1) RET(comm_max_i64(h, &nmax));
    if (nmax == 0) return set_err(h, CTR_ENOTFOUND, "no sample resolved: every key has an unknown user or item");
    const int64_t batches = nmax / B + (nmax % B != 0);     // model.go:96-99
    float best = INFINITY, cost = 0.0f; int no_improve = 0, ep = 0;
    for (ep = 0; ep < epochs; ep++) {
        for (int64_t b = (ep == 0 ? queued : 0); b < batches; b++) RET(run_batch(b, (int64_t)used));
        RET(read_cost(h, B * h->comm.world, &cost));       // cost of the epoch's last batch, model.go:198
        if (cost < best) { best = cost; no_improve = 0; } else no_improve++;
        if (early_stop != 0 && no_improve >= early_stop) { ep++; break; }   // model.go:206-209
    }
    if (last_cost) *last_cost = cost;
    if (epochs_run) *epochs_run = ep;
    return CTR_OK;
}

int ctr_predict_idx_dev(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, int32_t B, float* d_out) {
    if (!h || !d_user || !d_item || !d_hist) return set_err(h, CTR_EINVAL, "null argument");
    if (B < 1 || B > h->Bmax) return set_err(h, CTR_EINVAL, "batch %d outside (0, %d]", B, h->Bmax);
    if (emb_sharded(h)) return comm_predict(h, d_user, d_item, d_hist, B, d_out);
    RET(check_tables(h));
    RowSrc r = idx_src(h, d_user, d_item, d_hist, B);
    StepOpts o;
    RET(step_core(h, r, B, o));
    if (d_out) CU(h, cudaMemcpyAsync(d_out, h->P, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return CTR_OK;
}

int ctr_predict_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, int64_t n, float* out) {
    if (!h || !user_row || !item_row || !hist || !out || n < 1) return set_err(h, CTR_EINVAL, "bad predict arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    const int B = h->cfg.pred_batch, S = h->cfg.S;
    for (int64_t s = 0; s < n; s += B) {
        const int nb = (int)std::min<int64_t>(B, n - s);
        RET(stage_idx(h, user_row + s, item_row + s, hist + s * S, nullptr, nb));
        RET(ctr_predict_idx_dev(h, h->s_user, h->s_item, h->s_hist, nb, nullptr));
        CU(h, cudaMemcpyAsync(out + s, h->P, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));    // staging buffers are reused by the next chunk
    }
    return CTR_OK;
}

int ctr_last_cost(ctr_handle* h, float* cost) {
    if (!h || !cost) return CTR_EINVAL;
    CU(h, cudaSetDevice(h->dev));
    return read_cost(h, h->cfg.batch * h->comm.world, cost);
}

int ctr_sync(ctr_handle* h) {
    if (!h) return CTR_EINVAL;
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->comm.h_err && *h->comm.h_err) return set_err(h, CTR_ECOMM, "peer barrier timed out waiting for rank %d (a rank left the step loop?)", *h->comm.h_err - 1);
    return CTR_OK;
}

void* ctr_get_stream(ctr_handle* h) { return h ? (void*)h->stream : nullptr; }

int ctr_set_stream(ctr_handle* h, void* s) {
    if (!h) return CTR_EINVAL;
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)s;
    return CTR_OK;
}

int64_t ctr_launch_count(const ctr_handle* h) { return h ? h->launches : 0; }

int ctr_profile_enable(ctr_handle* h, int on) { if (!h) return CTR_EINVAL; h->profiling = on != 0; return CTR_OK; }
int ctr_profile_reset(ctr_handle* h) { if (!h) return CTR_EINVAL; h->prof.clear(); return CTR_OK; }
int ctr_profile_get(ctr_handle* h, const char* kernel, double* ms, int64_t* n) {
    if (!h || !kernel) return CTR_EINVAL;
    auto it = h->prof.find(kernel);
    if (it == h->prof.end()) { if (ms) *ms = 0; if (n) *n = 0; return CTR_OK; }
    if (ms) *ms = it->second.ms; if (n) *n = it->second.n;
    return CTR_OK;
}
int ctr_profile_dump(ctr_handle* h, char* buf, int64_t len) {
    if (!h || !buf || len < 1) return CTR_EINVAL;
    std::string s;
    for (auto& kv : h->prof) { char line[256]; snprintf(line, sizeof line, "%s %.6f %ld\n", kv.first.c_str(), kv.second.ms, kv.second.n); s += line; }
    snprintf(buf, (size_t)len, "%s", s.c_str());
    return CTR_OK;
}

int ctr_debug_grads_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row, const int32_t* hist, const float* label,
                        int32_t B, int32_t training, float* dmlp0, float* dmlp1, float* dmlp2, float* datt0,
                        float* dUb, float* dIt, float* p, float* logit, float* cost) {
    if (!h || !user_row || !item_row || !hist || !label) return set_err(h, CTR_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    RET(check_tables(h));
    const ctr_config& c = h->cfg;
    if (B < 1 || B > h->Bmax) return set_err(h, CTR_EINVAL, "batch %d outside (0, %d]", B, h->Bmax);
    RET(stage_idx(h, user_row, item_row, hist, label, B));
    RET(zero_grads(h));
    RowSrc r = idx_src(h, h->s_user, h->s_item, h->s_hist, B);
    StepOpts o; o.training = true; o.update = false; o.want_rows = true; o.d_label = h->s_label;
    ctr_config saved = h->cfg;
    if (!training) { h->cfg.dropout0 = 0; h->cfg.dropout1 = 0; }
    int rc = step_core(h, r, B, o);
    h->cfg = saved;
    RET(rc);
    if (dmlp0) CU(h, cudaMemcpy2DAsync(dmlp0, c.H0 * sizeof(float), h->G[0], h->H0p * sizeof(float), c.H0 * sizeof(float), h->in, cudaMemcpyDeviceToHost, h->stream));
    if (dmlp1) CU(h, cudaMemcpy2DAsync(dmlp1, c.H1 * sizeof(float), h->G[1], h->H1p * sizeof(float), c.H1 * sizeof(float), c.H0, cudaMemcpyDeviceToHost, h->stream));
    if (dmlp2) CU(h, cudaMemcpyAsync(dmlp2, h->G[2], c.H1 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (datt0) CU(h, cudaMemcpyAsync(datt0, h->G[3], c.S * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (dUb) CU(h, cudaMemcpyAsync(dUb, h->dUb, (size_t)B * c.S * c.D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (dIt) CU(h, cudaMemcpyAsync(dIt, h->dIt, (size_t)B * c.D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (p) CU(h, cudaMemcpyAsync(p, h->P, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (logit) CU(h, cudaMemcpyAsync(logit, h->Z, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    float cst = 0;
    RET(read_cost(h, B, &cst));
    if (cost) *cost = cst;
    RET(zero_grads(h));
    CU(h, cudaStreamSynchronize(h->stream));
    return CTR_OK;
}

int ctr_ubcache_upload(ctr_handle* h, const int64_t* offsets, const int64_t* ts, const int32_t* item_rows, int64_t n_users, int64_t n) {
    if (!h || !offsets || n_users < 1 || n < 0 || (n > 0 && (!ts || !item_rows))) return set_err(h, CTR_EINVAL, "bad ubcache arguments");
    if (offsets[0] != 0 || offsets[n_users] != n) return set_err(h, CTR_EINVAL, "ubcache offsets must run from 0 to n");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    for (void* p : {(void*)h->ub_off, (void*)h->ub_ts, (void*)h->ub_items}) if (p) cudaFree(p);
    h->ub_off = h->ub_ts = nullptr; h->ub_items = nullptr;
    RET(dalloc(h, &h->ub_off, (size_t)n_users + 1, false)); RET(dalloc(h, &h->ub_ts, (size_t)n, false)); RET(dalloc(h, &h->ub_items, (size_t)n, false));
    CU(h, cudaMemcpyAsync(h->ub_off, offsets, sizeof(int64_t) * (size_t)(n_users + 1), cudaMemcpyHostToDevice, h->stream));
    if (n > 0) {
        CU(h, cudaMemcpyAsync(h->ub_ts, ts, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->ub_items, item_rows, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    h->ub_users = n_users; h->ub_n = n;
    return CTR_OK;
}

int ctr_ubcache_window_dev(ctr_handle* h, const int32_t* d_user_row, const int64_t* d_max_ts, int32_t B, int32_t* d_hist_rows) {
    if (!h || !d_user_row || !d_max_ts || !d_hist_rows || B < 1) return set_err(h, CTR_EINVAL, "bad ubcache window arguments");
    if (!h->ub_off) return set_err(h, CTR_ESTATE, "ubcache not uploaded");
    CU(h, cudaSetDevice(h->dev));
    return launch(h, "ubcache_window", [&] {
        k_ub_window<<<grid_for_warps(h, B), 256, 0, h->stream>>>(h->ub_off, h->ub_ts, h->ub_items, d_user_row, (const long long*)d_max_ts, B, h->cfg.S,
                                                               (long)h->ub_users, d_hist_rows);
    });
}

int ctr_ubcache_window(ctr_handle* h, const int32_t* user_row, const int64_t* max_ts, int32_t B, int32_t* hist_rows) {
    if (!h || !user_row || !max_ts || !hist_rows || B < 1) return set_err(h, CTR_EINVAL, "bad ubcache window arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    int* du = nullptr; long long* dt = nullptr; int* dh = nullptr;
    int rc = CTR_OK;
    RET(scratch_get(h, 0, sizeof(int) * (size_t)B, (void**)&du)); RET(scratch_get(h, 1, sizeof(long long) * (size_t)B, (void**)&dt));
    RET(scratch_get(h, 2, sizeof(int) * (size_t)B * h->cfg.S, (void**)&dh));
    {
        cudaMemcpyAsync(du, user_row, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(dt, max_ts, sizeof(long long) * (size_t)B, cudaMemcpyHostToDevice, h->stream);
        rc = ctr_ubcache_window_dev(h, du, (const int64_t*)dt, B, dh);
        if (rc == CTR_OK && (cudaMemcpyAsync(hist_rows, dh, sizeof(int) * (size_t)B * h->cfg.S, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                             cudaStreamSynchronize(h->stream) != cudaSuccess)) rc = set_err(h, CTR_ECUDA, "ubcache window copy: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

// ---- sparse ids, serving keys, checkpoint (rows f3 / f4) -------------------------------------------------
int ctr_idmap_build(ctr_handle* h, int which, const int64_t* ids, int64_t n) {
    if (!h || !ids || which < 0 || which > 1 || n < 1 || n > 0x7fffffff) return set_err(h, CTR_EINVAL, "bad idmap arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    for (void* p : {(void*)h->idm_keys[which], (void*)h->idm_vals[which]}) if (p) cudaFree(p);
    h->idm_keys[which] = nullptr; h->idm_vals[which] = nullptr; h->idm_n[which] = 0;
    unsigned long long cap = 64;
    while (cap < 2ull * (unsigned long long)n) cap <<= 1;                 // load factor <= 0.5
    RET(dalloc(h, &h->idm_keys[which], (size_t)cap, false)); RET(dalloc(h, &h->idm_vals[which], (size_t)cap, false));
    if (!h->k_flags) RET(dalloc(h, &h->k_flags, 1));
    long long* d_ids = nullptr;
    CU(h, cudaMalloc(&d_ids, sizeof(long long) * (size_t)n));
    int flags = 0, rc = CTR_OK;
    cudaMemcpyAsync(d_ids, ids, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
    cudaMemsetAsync(h->k_flags, 0, sizeof(int), h->stream);
    rc = launch(h, "idmap_clear", [&] { k_idmap_clear<<<h->num_sms * 8, 256, 0, h->stream>>>(h->idm_keys[which], h->idm_vals[which], cap); });
    if (rc == CTR_OK) rc = launch(h, "idmap_insert", [&] {
        k_idmap_insert<<<h->num_sms * 8, 256, 0, h->stream>>>(h->idm_keys[which], h->idm_vals[which], cap - 1, d_ids, (long)n, h->k_flags);
    });
    if (rc == CTR_OK && (cudaMemcpyAsync(&flags, h->k_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                         cudaStreamSynchronize(h->stream) != cudaSuccess)) rc = set_err(h, CTR_ECUDA, "idmap build: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_ids);
    RET(rc);
    if (flags) {
        cudaFree(h->idm_keys[which]); cudaFree(h->idm_vals[which]); h->idm_keys[which] = nullptr; h->idm_vals[which] = nullptr;
        return set_err(h, CTR_EINVAL, flags & 2 ? "idmap: id INT64_MIN is reserved" : "idmap: duplicate ids");
    }
    h->idm_cap[which] = cap; h->idm_n[which] = n;
    return CTR_OK;
}

int ctr_idmap_lookup_dev(ctr_handle* h, int which, const int64_t* d_ids, int64_t n, int32_t* d_rows) {
    if (!h || !d_ids || !d_rows || which < 0 || which > 1 || n < 1) return set_err(h, CTR_EINVAL, "bad idmap lookup arguments");
    if (!h->idm_keys[which]) return set_err(h, CTR_ESTATE, "idmap %d not built", which);
    CU(h, cudaSetDevice(h->dev));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, h->num_sms * 8));
    return launch(h, "idmap_lookup", [&] {
        k_idmap_lookup<<<grid, 256, 0, h->stream>>>(h->idm_keys[which], h->idm_vals[which], h->idm_cap[which] - 1, (const long long*)d_ids, (long)n, d_rows);
    });
}

int ctr_idmap_lookup(ctr_handle* h, int which, const int64_t* ids, int64_t n, int32_t* rows) {
    if (!h || !ids || !rows || which < 0 || which > 1 || n < 1) return set_err(h, CTR_EINVAL, "bad idmap lookup arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    long long* d_ids = nullptr; int* d_rows = nullptr;
    int rc = CTR_OK;
    RET(scratch_get(h, 0, sizeof(long long) * (size_t)n, (void**)&d_ids)); RET(scratch_get(h, 1, sizeof(int) * (size_t)n, (void**)&d_rows));
    {
        cudaMemcpyAsync(d_ids, ids, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
        rc = ctr_idmap_lookup_dev(h, which, (const int64_t*)d_ids, n, d_rows);
        if (rc == CTR_OK && (cudaMemcpyAsync(rows, d_rows, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                             cudaStreamSynchronize(h->stream) != cudaSuccess)) rc = set_err(h, CTR_ECUDA, "idmap lookup copy: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

int ctr_batch_predict_keys(ctr_handle* h, const int64_t* user_ids, const int64_t* item_ids, const int64_t* ts, int64_t n, float* scores) {
    if (!h || !user_ids || !item_ids || !ts || !scores || n < 1) return set_err(h, CTR_EINVAL, "bad batch predict arguments");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    if (!h->idm_keys[0] || !h->idm_keys[1]) return set_err(h, CTR_ESTATE, "batch predict needs both id maps (ctr_idmap_build)");
    const int B = h->cfg.pred_batch, S = h->cfg.S;
    if (!h->k_keys) {       // device + pinned staging: [user ids | item ids | ts] and {first user row}
        RET(dalloc(h, &h->k_keys, (size_t)3 * B, false));
        CU(h, cudaMallocHost(&h->k_host, sizeof(long long) * (size_t)3 * B + sizeof(float) * (size_t)B + 16));
    }
    long long* hk = h->k_host;
    float* hs = (float*)(hk + (size_t)3 * B);
    int* hfirst = (int*)(hs + B);
    for (int64_t s = 0; s < n; s += B) {
        const int nb = (int)std::min<int64_t>(B, n - s);
        memcpy(hk, user_ids + s, sizeof(long long) * (size_t)nb);
        memcpy(hk + nb, item_ids + s, sizeof(long long) * (size_t)nb);
        memcpy(hk + 2 * (size_t)nb, ts + s, sizeof(long long) * (size_t)nb);
        CU(h, cudaMemcpyAsync(h->k_keys, hk, sizeof(long long) * (size_t)3 * nb, cudaMemcpyHostToDevice, h->stream));
        RET(launch(h, "keys_resolve", [&] {
            k_keys_resolve<<<grid_for_warps(h, nb), 256, 0, h->stream>>>(h->idm_keys[0], h->idm_vals[0], h->idm_cap[0] - 1, h->idm_keys[1], h->idm_vals[1], h->idm_cap[1] - 1,
                                                                    h->ub_off, h->ub_ts, h->ub_items, (long)h->ub_users, h->k_keys, nb, S, h->s_user, h->s_item, h->s_hist);
        }));
        if (s == 0) CU(h, cudaMemcpyAsync(hfirst, h->s_user, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        RET(ctr_predict_idx_dev(h, h->s_user, h->s_item, h->s_hist, nb, nullptr));
        CU(h, cudaMemcpyAsync(hs, h->P, (size_t)nb * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        // rcmd.go:300-303: an unresolvable FIRST key is an error, later ones become zero rows
        if (s == 0 && *hfirst < 0) return set_err(h, CTR_ENOTFOUND, "get sample vector error: user %lld or item %lld has no features", (long long)user_ids[0], (long long)item_ids[0]);
        memcpy(scores + s, hs, sizeof(float) * (size_t)nb);
    }
    return CTR_OK;
}

namespace {
struct CkptHeader {
    char magic[8];                 // "CTRB200\0"
    int32_t version, model, uP, S, D, cF, H0, H1, rank, world;
    uint32_t step;
    int32_t has_table[3];
    int64_t tab_rows[3], tab_local_rows[3];
    int32_t tab_width[3];
    int32_t replicated;            // ITEM_EMB placement under world > 1 (1 = every rank holds the whole table)
    int32_t has_moments;           // ITEM_EMB Adam moments follow the tables (CTR_TABLE_ADAM)
};
constexpr size_t kCkptChunk = 32u << 20;

// device [rows, ld] → file as compact [rows, width], through a pinned bounce buffer
int ckpt_write_2d(ctr_handle* h, FILE* f, const float* d, long ld, int64_t rows, int width, float* bounce) {
    const int64_t per = std::max<int64_t>(1, (int64_t)(kCkptChunk / sizeof(float)) / width);
    for (int64_t r = 0; r < rows; r += per) {
        const int64_t nr = std::min(per, rows - r);
        CU(h, cudaMemcpy2DAsync(bounce, (size_t)width * sizeof(float), d + (size_t)r * ld, (size_t)ld * sizeof(float), (size_t)width * sizeof(float), (size_t)nr, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        if (fwrite(bounce, sizeof(float), (size_t)nr * width, f) != (size_t)nr * width) return set_err(h, CTR_EIO, "checkpoint: short write");
    }
    return CTR_OK;
}
int ckpt_read_2d(ctr_handle* h, FILE* f, float* d, long ld, int64_t rows, int width, float* bounce) {
    const int64_t per = std::max<int64_t>(1, (int64_t)(kCkptChunk / sizeof(float)) / width);
    for (int64_t r = 0; r < rows; r += per) {
        const int64_t nr = std::min(per, rows - r);
        if (fread(bounce, sizeof(float), (size_t)nr * width, f) != (size_t)nr * width) return set_err(h, CTR_EIO, "checkpoint: truncated file");
        CU(h, cudaMemcpy2DAsync(d + (size_t)r * ld, (size_t)ld * sizeof(float), bounce, (size_t)width * sizeof(float), (size_t)width * sizeof(float), (size_t)nr, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return CTR_OK;
}
// the four learnable tensors in their logical (reference JSON) shapes
void dense_shape(const ctr_handle* h, int i, int64_t* rows, int* width, long* ld) {
    const ctr_config& c = h->cfg;
    if (i == 0) { *rows = h->in; *width = c.H0; *ld = h->H0p; }
    else if (i == 1) { *rows = c.H0; *width = c.H1; *ld = h->H1p; }
    else if (i == 2) { *rows = 1; *width = c.H1; *ld = h->H1p; }
    else { *rows = 1; *width = c.S; *ld = h->Sp; }
}
}  // namespace

int ctr_checkpoint_save(ctr_handle* h, const char* path) {
    if (!h || !path) return set_err(h, CTR_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    if (h->tab_sharded[CTR_TABLE_ITEM_EMB]) RET(comm_hot_writeback(h));
    CU(h, cudaStreamSynchronize(h->stream));
    FILE* f = fopen(path, "wb");
    if (!f) return set_err(h, CTR_EIO, "checkpoint: cannot open %s for writing", path);
    float* bounce = nullptr;
    if (cudaMallocHost(&bounce, kCkptChunk) != cudaSuccess) { fclose(f); return set_err(h, CTR_ENOMEM, "checkpoint bounce buffer"); }
    const ctr_config& c = h->cfg;
    CkptHeader hd{};
    memcpy(hd.magic, "CTRB200", 8);
    hd.version = 2; hd.model = c.model; hd.uP = c.uP; hd.S = c.S; hd.D = c.D; hd.cF = c.cF; hd.H0 = c.H0; hd.H1 = c.H1;
    hd.rank = h->comm.rank; hd.world = h->comm.world; hd.step = h->step; hd.replicated = h->comm.replicate ? 1 : 0;
    hd.has_moments = (h->emb_m && h->tab[CTR_TABLE_ITEM_EMB]) ? 1 : 0;
    for (int t = 0; t < 3; t++) { hd.has_table[t] = h->tab[t] != nullptr; hd.tab_rows[t] = h->tab_rows[t]; hd.tab_local_rows[t] = h->tab_local_rows[t]; hd.tab_width[t] = h->tab_width[t]; }
    int rc = fwrite(&hd, sizeof hd, 1, f) == 1 ? CTR_OK : set_err(h, CTR_EIO, "checkpoint: short write");
    for (int i = 0; i < 4 && rc == CTR_OK; i++) {
        int64_t rows; int width; long ld; dense_shape(h, i, &rows, &width, &ld);
        for (float* src : {h->W[i], h->Mo[i], h->Vo[i]}) if (rc == CTR_OK) rc = ckpt_write_2d(h, f, src, ld, rows, width, bounce);
    }
    for (int t = 0; t < 3 && rc == CTR_OK; t++)
        if (h->tab[t]) rc = ckpt_write_2d(h, f, h->tab[t], h->tab_ld[t], h->tab_local_rows[t], h->tab_width[t], bounce);
    if (hd.has_moments)
        for (float* src : {h->emb_m, h->emb_v})
            if (rc == CTR_OK) rc = ckpt_write_2d(h, f, src, h->tab_ld[CTR_TABLE_ITEM_EMB], h->tab_local_rows[CTR_TABLE_ITEM_EMB], h->tab_width[CTR_TABLE_ITEM_EMB], bounce);
    cudaFreeHost(bounce);
    if (fclose(f) != 0 && rc == CTR_OK) rc = set_err(h, CTR_EIO, "checkpoint: close failed");
    return rc;
}

int ctr_checkpoint_load(ctr_handle* h, const char* path) {
    if (!h || !path) return set_err(h, CTR_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(h, CTR_EIO, "checkpoint: cannot open %s", path);
    CkptHeader hd{};
    const ctr_config& c = h->cfg;
    int rc = CTR_OK;
    if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "CTRB200", 8) != 0 || hd.version != 2) rc = set_err(h, CTR_EIO, "checkpoint: %s is not a version-2 ctr-b200 snapshot", path);
    else if (hd.model != c.model || hd.uP != c.uP || hd.S != c.S || hd.D != c.D || hd.cF != c.cF || hd.H0 != c.H0 || hd.H1 != c.H1)
        rc = set_err(h, CTR_EINVAL, "checkpoint: model dims differ from this handle");
    else if (hd.rank != h->comm.rank || hd.world != h->comm.world) rc = set_err(h, CTR_EINVAL, "checkpoint: written by rank %d/%d, this handle is %d/%d", hd.rank, hd.world, h->comm.rank, h->comm.world);
    float* bounce = nullptr;
    if (rc == CTR_OK && cudaMallocHost(&bounce, kCkptChunk) != cudaSuccess) rc = set_err(h, CTR_ENOMEM, "checkpoint bounce buffer");
    for (int i = 0; i < 4 && rc == CTR_OK; i++) {
        int64_t rows; int width; long ld; dense_shape(h, i, &rows, &width, &ld);
        for (float* dst : {h->W[i], h->Mo[i], h->Vo[i]}) if (rc == CTR_OK) rc = ckpt_read_2d(h, f, dst, ld, rows, width, bounce);
    }
    for (int t = 0; t < 3 && rc == CTR_OK; t++) {
        if (!hd.has_table[t]) continue;
        const int want = t == CTR_TABLE_USER_FEAT ? c.uP : t == CTR_TABLE_ITEM_FEAT ? c.cF : c.D;
        if (hd.tab_width[t] != want || hd.tab_rows[t] < 1 || hd.tab_local_rows[t] < 0) { rc = set_err(h, CTR_EIO, "checkpoint: bad table %d header", t); break; }
        const long ld = round_up(hd.tab_width[t], 4);
        const size_t bytes = (size_t)std::max<int64_t>(hd.tab_local_rows[t], 1) * ld * sizeof(float);
        // placement travels with the snapshot: a shard holds fewer rows than the logical table
        h->tab_sharded[t] = h->comm.world > 1 && t != CTR_TABLE_USER_FEAT && (t == CTR_TABLE_ITEM_EMB ? hd.replicated != 1 : hd.tab_local_rows[t] != hd.tab_rows[t]);
        rc = table_mem(h, t, h->tab_sharded[t], bytes);
        if (rc != CTR_OK) break;
        h->tab_ld[t] = ld; h->tab_rows[t] = hd.tab_rows[t]; h->tab_local_rows[t] = hd.tab_local_rows[t]; h->tab_width[t] = hd.tab_width[t];
        if (t == CTR_TABLE_ITEM_EMB) h->comm.replicate = h->comm.world > 1 && hd.replicated == 1;
        if (t != CTR_TABLE_USER_FEAT) h->tab_gen++;
        rc = ckpt_read_2d(h, f, h->tab[t], ld, hd.tab_local_rows[t], hd.tab_width[t], bounce);
        if (t == CTR_TABLE_ITEM_EMB && h->hot_acc) { cudaFree(h->hot_acc); h->hot_acc = nullptr; h->hot_rows = 0; }
    }
    for (float** p : {&h->emb_m, &h->emb_v}) if (*p) { cudaFree(*p); *p = nullptr; }
    if (rc == CTR_OK && hd.has_moments) {
        rc = ensure_moments(h);
        for (float* dst : {h->emb_m, h->emb_v})
            if (rc == CTR_OK) rc = ckpt_read_2d(h, f, dst, h->tab_ld[CTR_TABLE_ITEM_EMB], h->tab_local_rows[CTR_TABLE_ITEM_EMB], h->tab_width[CTR_TABLE_ITEM_EMB], bounce);
    }
    if (bounce) cudaFreeHost(bounce);
    fclose(f);
    if (rc == CTR_OK) { h->step = hd.step; h->um.dirty = true; cudaStreamSynchronize(h->stream); }
    return rc;
}

int ctr_roc_auc(ctr_handle* h, const float* pred, const float* y, int64_t n, double* auc) {
    if (!h || !pred || !y || !auc || n < 1 || n > 0x7fffffff) return set_err(h, CTR_EINVAL, "bad auc arguments (n must be in [1, 2^31))");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    { cudaError_t e = auc_run(h->stream, pred, y, (long)n, auc); if (e != cudaSuccess) return set_err(h, CTR_ECUDA, "auc: %s", cudaGetErrorString(e)); return CTR_OK; }
}

int ctr_i2v_paths(const int64_t* count, int32_t V, int32_t max_depth, int64_t* path_off, int32_t* path_node, uint8_t* path_code, int64_t cap) {
    if (!count || !path_off || V < 2 || max_depth < 2) return set_err(nullptr, CTR_EINVAL, "bad item2vec path arguments");
    std::vector<int64_t> cnt(count, count + V), node_val; std::vector<long long> poff; std::vector<int> pnode; std::vector<unsigned char> pcode;
    i2v_build_paths(cnt, V, max_depth, node_val, poff, pnode, pcode);
    for (int i = 0; i <= V; i++) path_off[i] = poff[(size_t)i];
    if ((int64_t)pnode.size() > cap) return set_err(nullptr, CTR_EINVAL, "path buffers too small: need %lld", (long long)pnode.size());
    if (path_node) memcpy(path_node, pnode.data(), sizeof(int) * pnode.size());
    if (path_code) memcpy(path_code, pcode.data(), pcode.size());
    return CTR_OK;
}

void ctr_i2v_config_default(ctr_i2v_config* c) {
    memset(c, 0, sizeof *c);
    c->dim = 16; c->window = 5; c->iter = 1;                  // rcmd.go:22-26, 543
    c->min_count = 5; c->max_depth = 100;                      // options.go:47-48
    c->init_lr = 0.025f; c->min_lr = 0.025f * 1.0e-4f; c->subsample = 1.0e-3f; c->update_lr_batch = 100000;
    c->seed = 0; c->device = 0;
}

namespace {
// Large transfers between PAGEABLE caller memory and the device (item2vec: the token stream in, the vector table out):
// two pinned bounce buffers, host-side memcpy by the copy pool overlapping the DMA of the other buffer.
cudaError_t big_copy(void* dst, const void* src, size_t bytes, bool to_device, cudaStream_t st) {
    const size_t kChunk = (size_t)32 << 20;
    if (bytes < ((size_t)8 << 20)) { cudaError_t e = cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st); return e != cudaSuccess ? e : cudaStreamSynchronize(st); }
    unsigned char* pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; i++) { e = cudaHostAlloc(&pin[i], kChunk, cudaHostAllocDefault); if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); }
    if (e == cudaSuccess) {
        CopyPool pool(7);
        const size_t nchunks = (bytes + kChunk - 1) / kChunk;
        if (to_device) {
            for (size_t c = 0; c < nchunks && e == cudaSuccess; c++) {
                const int b = (int)(c & 1); const size_t off = c * kChunk, m = std::min(kChunk, bytes - off);
                if (c >= 2) e = cudaEventSynchronize(ev[b]);
                pool.copy(pin[b], (const char*)src + off, m);
                if (e == cudaSuccess) e = cudaMemcpyAsync((char*)dst + off, pin[b], m, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        } else {
            // DMA chunk c+1 while the pool copies chunk c out of its bounce buffer
            if (nchunks > 0) { e = cudaMemcpyAsync(pin[0], src, std::min(kChunk, bytes), cudaMemcpyDeviceToHost, st); if (e == cudaSuccess) e = cudaEventRecord(ev[0], st); }
            for (size_t c = 0; c < nchunks && e == cudaSuccess; c++) {
                const int b = (int)(c & 1); const size_t off = c * kChunk, m = std::min(kChunk, bytes - off);
                if (c + 1 < nchunks) {
                    const size_t off2 = (c + 1) * kChunk, m2 = std::min(kChunk, bytes - off2);
                    e = cudaMemcpyAsync(pin[b ^ 1], (const char*)src + off2, m2, cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess) e = cudaEventRecord(ev[b ^ 1], st);
                }
                if (e == cudaSuccess) e = cudaEventSynchronize(ev[b]);
                if (e == cudaSuccess) pool.copy((char*)dst + off, pin[b], m);
            }
        }
    }
    for (int i = 0; i < 2; i++) { if (pin[i]) cudaFreeHost(pin[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    return e;
}

struct I2vDist { int rank = 0, world = 1; void* nccl = nullptr; long sync_every = 0; };

// embedding.TrainEmbedding on the device.  Stream-proportional work (dictionary counts, MinCount filter) and the
// per-word Huffman paths are built on the GPU; the host only runs the O(V) two-queue Huffman merge over the
// count-sorted leaves.  dist (world > 1): every rank trains on its own shard of the stream with a full replica of
// both vector tables; dictionary counts are all-reduced (identical tree everywhere) and the replicas are averaged
// every sync_every positions (local SGD / model averaging — the reference's own trainer is Hogwild over goroutines,
// word2vec.go:165-169; across GPUs the racy shared memory becomes periodic averaging).
int i2v_train_impl(const ctr_i2v_config* cfg, const int32_t* tokens, int64_t n, int32_t V, float* emb_out, ctr_i2v_stats* stats, I2vDist* dist) {
    if (!cfg || !tokens || (!emb_out && !dist) || n < 1 || V < 2 || n > 0x7fffffffLL) return set_err(nullptr, CTR_EINVAL, "bad item2vec arguments");
    const ctr_i2v_config& c = *cfg;
    const int D = c.dim, W = c.window;
    if (D < 4 || D > 128 || (D & (D - 1)) || W < 1 || c.iter < 1 || c.max_depth < 2 || c.update_lr_batch < 1)
        return set_err(nullptr, CTR_EINVAL, "item2vec: dim must be a power of two in [4,128], window/iter >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_err(nullptr, CTR_ENODEV, "no CUDA device: this engine has no CPU fallback"); }
    const int world = dist ? dist->world : 1;
    const bool seq = c.reserved[0] == 1;
    if (seq && world > 1) return set_err(nullptr, CTR_EINVAL, "the sequential float64 mode is single-GPU");
#define CI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = set_err(nullptr, CTR_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); goto done; } } while (0)
#define CN(call) do { int r_ = (call); if (r_ != 0) { rc = set_err(nullptr, CTR_ECOMM, "%s: %s", #call, g_nccl.GetErrorString(r_)); goto done; } } while (0)
    int rc = CTR_OK;
    int *d_tok = nullptr, *d_doc = nullptr, *d_pnode = nullptr, *d_sid = nullptr, *d_sid2 = nullptr, *d_parent = nullptr, *d_bad = nullptr;
    unsigned long long *d_cnt64 = nullptr, *d_skey = nullptr, *d_skey2 = nullptr, *d_ctr = nullptr, *d_misc = nullptr;
    unsigned char *d_keep = nullptr, *d_pcode = nullptr, *d_code = nullptr;
    long long *d_poff = nullptr, *d_nval = nullptr; long* d_nd = nullptr;
    double* d_z = nullptr; float *d_syn0 = nullptr, *d_syn1 = nullptr, *d_lr = nullptr, *d_nsc = nullptr, *d_wsc = nullptr;
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms_total = 0; int launches = 0; unsigned long long hc[3] = {0, 0, 0};
    {
        CI(cudaSetDevice(c.device));
        cudaDeviceProp prop{}; CI(cudaGetDeviceProperties(&prop, c.device));
        if (prop.major != 10) { rc = set_err(nullptr, CTR_ENODEV, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor); goto done; }
        const int G = prop.multiProcessorCount * 8;
        CI(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); CI(cudaEventCreate(&e0)); CI(cudaEventCreate(&e1));
        // ---- dictionary counts over the whole stream (dictionary.go:70-81), on the device
        CI(cudaMalloc(&d_tok, sizeof(int) * (size_t)n)); CI(cudaMalloc(&d_doc, sizeof(int) * (size_t)n)); CI(cudaMalloc(&d_keep, (size_t)n));
        CI(cudaMalloc(&d_cnt64, sizeof(unsigned long long) * (size_t)V)); CI(cudaMalloc(&d_bad, sizeof(int))); CI(cudaMalloc(&d_nd, sizeof(long)));
        CI(cudaMalloc(&d_misc, 4 * sizeof(unsigned long long)));
        CI(big_copy(d_tok, tokens, sizeof(int) * (size_t)n, true, st));
        CI(cudaMemsetAsync(d_cnt64, 0, sizeof(unsigned long long) * (size_t)V, st)); CI(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
        k_i2v_count<<<G, 256, 0, st>>>(d_tok, (long)n, V, d_cnt64, d_bad); launches++;
        unsigned long long n_global = (unsigned long long)n;
        long nd_max_global = 0;
        if (world > 1) {
            CN(g_nccl.AllReduce(d_cnt64, d_cnt64, (size_t)V, kNcclUint64, kNcclSum, dist->nccl, st));
            CI(cudaMemcpyAsync(d_misc, &n_global, sizeof n_global, cudaMemcpyHostToDevice, st));
            CN(g_nccl.AllReduce(d_misc, d_misc, 1, kNcclUint64, kNcclSum, dist->nccl, st));
            CI(cudaMemcpyAsync(&n_global, d_misc, sizeof n_global, cudaMemcpyDeviceToHost, st));
        }
        // ---- the training document: words rarer than MinCount are dropped (memory.go:53-62), order kept
        k_i2v_keep<<<G, 256, 0, st>>>(d_tok, (long)n, d_cnt64, c.min_count, d_keep); launches++;
        cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_tok, d_keep, d_doc, d_nd, (int)n, st);
        {
            size_t b2 = 0, b3 = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, b2, d_skey, d_skey2, d_sid, d_sid2, V, 0, 64, st);
            cub::DeviceScan::ExclusiveSum(nullptr, b3, d_poff, d_poff, V + 1, st);
            tmp_bytes = std::max(tmp_bytes, std::max(b2, b3));
        }
        CI(cudaMalloc(&d_tmp, tmp_bytes));
        { size_t tb = tmp_bytes; cub::DeviceSelect::Flagged(d_tmp, tb, d_tok, d_keep, d_doc, d_nd, (int)n, st); launches++; }
        long nd = 0; int bad = 0;
        CI(cudaMemcpyAsync(&nd, d_nd, sizeof nd, cudaMemcpyDeviceToHost, st)); CI(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, st));
        CI(cudaStreamSynchronize(st));
        if (bad) { rc = set_err(nullptr, CTR_EINVAL, "a token lies outside [0, vocab)"); goto done; }
        nd_max_global = nd;
        if (world > 1) {
            unsigned long long v = (unsigned long long)nd;
            CI(cudaMemcpyAsync(d_misc, &v, sizeof v, cudaMemcpyHostToDevice, st));
            CN(g_nccl.AllReduce(d_misc, d_misc, 1, kNcclUint64, kNcclMax, dist->nccl, st));
            CI(cudaMemcpyAsync(&v, d_misc, sizeof v, cudaMemcpyDeviceToHost, st)); CI(cudaStreamSynchronize(st));
            nd_max_global = (long)v;
        }
        // concurrency: (centre, context) pairs in flight ~ vocabulary / 4, at most the whole machine
        const int rpw = 32 / (D / 4);
        const long max_warps = (long)prop.multiProcessorCount * 8 * 8;
        const long warps = std::max<long>(1, std::min<long>(std::min<long>(max_warps, (std::max<long>(nd, 1) + 63) / 64), std::max<long>(1, (long)V / 4 / rpw)));
        const int grid = (int)((warps + 7) / 8);
        const double Ceff = (double)grid * 8 * rpw;
        // ---- per-word tables + count-sorted leaves (stable: ties stay in id order, like the reference's sort)
        CI(cudaMalloc(&d_z, sizeof(double) * (size_t)V)); CI(cudaMalloc(&d_wsc, sizeof(float) * (size_t)V));
        CI(cudaMalloc(&d_skey, sizeof(unsigned long long) * (size_t)V)); CI(cudaMalloc(&d_skey2, sizeof(unsigned long long) * (size_t)V));
        CI(cudaMalloc(&d_sid, sizeof(int) * (size_t)V)); CI(cudaMalloc(&d_sid2, sizeof(int) * (size_t)V));
        k_i2v_word_tables<<<G, 256, 0, st>>>(d_cnt64, V, (double)c.subsample, Ceff, (double)n_global, d_z, d_wsc, d_skey, d_sid); launches++;
        { size_t tb = tmp_bytes; cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_skey, d_skey2, d_sid, d_sid2, V, 0, 64, st); launches++; }
        std::vector<int64_t> cnt((size_t)V); std::vector<int> leaves((size_t)V);
        CI(cudaMemcpyAsync(cnt.data(), d_cnt64, sizeof(int64_t) * (size_t)V, cudaMemcpyDeviceToHost, st));
        CI(cudaMemcpyAsync(leaves.data(), d_sid2, sizeof(int) * (size_t)V, cudaMemcpyDeviceToHost, st));
        CI(cudaStreamSynchronize(st));
        // ---- host: the Huffman merge (huffman.go:23-57), O(V)
        std::vector<int> parent; std::vector<unsigned char> code; std::vector<long long> node_val;
        const int n_inner = i2v_huffman_sorted(cnt.data(), leaves.data(), V, parent, code, node_val);
        if (n_inner < 1) { rc = set_err(nullptr, CTR_EINVAL, "item2vec needs at least two distinct words in the stream"); goto done; }
        // ---- paths (node.go:26-43) on the device
        CI(cudaMalloc(&d_parent, sizeof(int) * parent.size())); CI(cudaMalloc(&d_code, code.size())); CI(cudaMalloc(&d_nval, sizeof(long long) * node_val.size()));
        CI(cudaMalloc(&d_poff, sizeof(long long) * ((size_t)V + 1))); CI(cudaMalloc(&d_nsc, sizeof(float) * node_val.size()));
        CI(cudaMemcpyAsync(d_parent, parent.data(), sizeof(int) * parent.size(), cudaMemcpyHostToDevice, st));
        CI(cudaMemcpyAsync(d_code, code.data(), code.size(), cudaMemcpyHostToDevice, st));
        CI(cudaMemcpyAsync(d_nval, node_val.data(), sizeof(long long) * node_val.size(), cudaMemcpyHostToDevice, st));
        CI(cudaMemsetAsync(d_poff, 0, sizeof(long long) * ((size_t)V + 1), st));
        k_i2v_paths<<<G, 256, 0, st>>>(d_parent, d_code, d_cnt64, V, c.max_depth, 0, d_poff, nullptr, nullptr); launches++;
        { size_t tb = tmp_bytes; cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_poff, d_poff, V + 1, st); launches++; }
        long long nsteps = 0;
        CI(cudaMemcpyAsync(&nsteps, d_poff + V, sizeof nsteps, cudaMemcpyDeviceToHost, st)); CI(cudaStreamSynchronize(st));
        CI(cudaMalloc(&d_pnode, sizeof(int) * (size_t)std::max<long long>(nsteps, 1))); CI(cudaMalloc(&d_pcode, (size_t)std::max<long long>(nsteps, 1)));
        k_i2v_paths<<<G, 256, 0, st>>>(d_parent, d_code, d_cnt64, V, c.max_depth, 1, d_poff, d_pnode, d_pcode); launches++;
        k_i2v_node_scale<<<G, 256, 0, st>>>(d_nval, V - 1, Ceff, (double)n_global, d_nsc); launches++;
        CI(cudaMalloc(&d_ctr, 3 * sizeof(unsigned long long))); CI(cudaMemsetAsync(d_ctr, 0, 3 * sizeof(unsigned long long), st));
        if (seq) {
            // sequential float64 parity mode (item2vec.cuh): one warp, the reference's single-goroutine order
            double *d64_0 = nullptr, *d64_1 = nullptr, *d_lut64 = nullptr; float* d_out32 = nullptr;
            std::vector<double> lut64(1000);
            for (int i = 0; i < 1000; i++) { double e = std::exp(((double)i / 1000.0 * 2.0 - 1.0) * 6.0); lut64[(size_t)i] = e / (e + 1.0); }   // sigmoid_table.go:28-45
            cudaError_t e1_ = cudaMalloc(&d64_0, sizeof(double) * (size_t)V * D), e2_ = cudaMalloc(&d64_1, sizeof(double) * (size_t)std::max(V - 1, 1) * D),
                        e3_ = cudaMalloc(&d_lut64, sizeof(double) * 1000), e4_ = cudaMalloc(&d_out32, sizeof(float) * (size_t)V * D);
            if (e1_ != cudaSuccess || e2_ != cudaSuccess || e3_ != cudaSuccess || e4_ != cudaSuccess) rc = set_err(nullptr, CTR_ENOMEM, "item2vec float64 tables");
            if (rc == CTR_OK) {
                cudaMemcpyAsync(d_lut64, lut64.data(), sizeof(double) * 1000, cudaMemcpyHostToDevice, st);
                cudaMemsetAsync(d64_1, 0, sizeof(double) * (size_t)std::max(V - 1, 1) * D, st);
                k_i2v_init64<<<G, 256, 0, st>>>(d64_0, (long)V * D, D, c.seed); launches++;
                I2vSeqArgs sa{}; sa.doc = d_doc; sa.nd = nd; sa.n_stream = (long)n; sa.z = d_z; sa.poff = d_poff; sa.pnode = d_pnode; sa.pcode = d_pcode;
                sa.syn0 = d64_0; sa.syn1 = d64_1; sa.lut = d_lut64; sa.D = D; sa.W = W; sa.upd = c.update_lr_batch; sa.iters = c.iter;
                sa.init_lr = (double)c.init_lr; sa.min_lr = (double)c.min_lr; sa.seed = c.seed; sa.counters = d_ctr;
                cudaEventRecord(e0, st);
                k_i2v_seq_f64<<<1, 32, 0, st>>>(sa); launches++;
                cudaEventRecord(e1, st);
                k_f64_to_f32_i2v<<<G, 256, 0, st>>>(d64_0, d_out32, (long)V * D); launches++;
                cudaMemcpyAsync(emb_out, d_out32, sizeof(float) * (size_t)V * D, cudaMemcpyDeviceToHost, st);
                cudaMemcpyAsync(hc, d_ctr, sizeof hc, cudaMemcpyDeviceToHost, st);
                if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = set_err(nullptr, CTR_ECUDA, "item2vec sequential mode: %s", cudaGetErrorString(cudaGetLastError()));
                else {
                    cudaEventElapsedTime(&ms_total, e0, e1);
                    if (stats) {
                        stats->doc_len = nd; stats->trained_positions = (int64_t)hc[0]; stats->pairs = (int64_t)hc[1]; stats->node_visits = (int64_t)hc[2];
                        stats->algorithmic_bytes = 2.0 * D * 8.0 * ((double)hc[1] + (double)hc[2]); stats->ms_device = ms_total; stats->launches = launches;
                    }
                }
            }
            for (void* p : {(void*)d64_0, (void*)d64_1, (void*)d_lut64, (void*)d_out32}) if (p) cudaFree(p);
            goto done;
        }
        // ---- vector tables: syn0 = (U-0.5)/dim (word2vec.go:103-111), inner nodes zero (huffman.go:40)
        std::vector<float> lut(1000);
        for (int i = 0; i < 1000; i++) { double e = std::exp(((double)i / 1000.0 * 2.0 - 1.0) * 6.0); lut[(size_t)i] = (float)(e / (e + 1.0)); }
        CI(cudaMemcpyToSymbolAsync(c_i2v_lut, lut.data(), sizeof(float) * 1000, 0, cudaMemcpyHostToDevice, st));
        CI(cudaMalloc(&d_syn0, sizeof(float) * (size_t)V * D)); CI(cudaMalloc(&d_syn1, sizeof(float) * (size_t)(V - 1) * D));
        CI(cudaMemsetAsync(d_syn1, 0, sizeof(float) * (size_t)(V - 1) * D, st));
        k_i2v_init<<<G, 256, 0, st>>>(d_syn0, (long)V * D, D, c.seed); launches++;
        const long nchunks = nd_max_global / c.update_lr_batch + 2;
        CI(cudaMalloc(&d_lr, sizeof(float) * (size_t)nchunks));
        std::vector<float> lr_tab((size_t)nchunks);
        double lr = c.init_lr;                                                           // w.currentlr persists across iterations
        // segments: one per iteration on a single GPU; every sync_every positions when replicas have to be averaged
        const long seg = (world > 1 && dist->sync_every > 0) ? dist->sync_every : std::max<long>(nd_max_global, 1);
        for (int it = 0; it < c.iter; it++) {
            for (long k = 0; k < nchunks; k++) {                                         // observe(), word2vec.go:223-233 (positions of all ranks advance together)
                lr_tab[(size_t)k] = (float)lr;
                const double seen = (double)(k + 1) * c.update_lr_batch * world;
                if ((double)(k + 1) * c.update_lr_batch <= (double)nd_max_global) lr = lr < (double)c.min_lr ? (double)c.min_lr : (double)c.init_lr * (1.0 - seen / (double)n_global);
            }
            CI(cudaMemcpyAsync(d_lr, lr_tab.data(), sizeof(float) * (size_t)nchunks, cudaMemcpyHostToDevice, st));
            I2vArgs a{}; a.doc = d_doc; a.nd = nd; a.z = d_z; a.poff = d_poff; a.pnode = d_pnode; a.pcode = d_pcode; a.syn0 = d_syn0; a.syn1 = d_syn1;
            a.D = D; a.W = W; a.lr_tab = d_lr; a.upd = c.update_lr_batch; a.seed = c.seed; a.iter = it; a.counters = d_ctr;
            a.node_scale = d_nsc; a.word_scale = d_wsc;
            static const bool no_hot = getenv("CTR_I2V_NO_HOT") != nullptr;
            a.hot_n = no_hot ? 0 : std::min(kI2vHot, n_inner); a.hot_base = n_inner - a.hot_n;      // the last nodes created = the top of the tree
            const size_t sm = (size_t)std::max(a.hot_n, 1) * D * sizeof(float);
            CI(cudaEventRecord(e0, st));
            for (long p0 = 0; p0 < std::max<long>(nd_max_global, 1); p0 += seg) {
                a.pos_begin = p0; a.pos_end = std::min<long>(nd, p0 + seg);
                if (a.pos_end > a.pos_begin) {
                    switch (D / 4) {
                        case 1: k_i2v_skipgram_hs<1><<<grid, 256, sm, st>>>(a); break;   case 2: k_i2v_skipgram_hs<2><<<grid, 256, sm, st>>>(a); break;
                        case 4: k_i2v_skipgram_hs<4><<<grid, 256, sm, st>>>(a); break;   case 8: k_i2v_skipgram_hs<8><<<grid, 256, sm, st>>>(a); break;
                        case 16: k_i2v_skipgram_hs<16><<<grid, 256, sm, st>>>(a); break; default: k_i2v_skipgram_hs<32><<<grid, 256, sm, st>>>(a); break;
                    }
                    launches++;
                }
                if (world > 1) {      // model averaging: every replica continues from the mean of all replicas
                    CN(g_nccl.GroupStart());
                    CN(g_nccl.AllReduce(d_syn0, d_syn0, (size_t)V * D, kNcclFloat32, kNcclSum, dist->nccl, st));
                    CN(g_nccl.AllReduce(d_syn1, d_syn1, (size_t)(V - 1) * D, kNcclFloat32, kNcclSum, dist->nccl, st));
                    CN(g_nccl.GroupEnd());
                    k_i2v_scale<<<G, 256, 0, st>>>(d_syn0, (long)V * D, 1.0f / (float)world);
                    k_i2v_scale<<<G, 256, 0, st>>>(d_syn1, (long)(V - 1) * D, 1.0f / (float)world);
                    launches += 3;
                }
            }
            CI(cudaGetLastError());
            CI(cudaEventRecord(e1, st)); CI(cudaStreamSynchronize(st));                   // lr_tab is reused by the next iteration
            float ms = 0; CI(cudaEventElapsedTime(&ms, e0, e1)); ms_total += ms;
        }
        CI(cudaStreamSynchronize(st));
        if (emb_out) CI(big_copy(emb_out, d_syn0, sizeof(float) * (size_t)V * D, false, st));
        CI(cudaMemcpyAsync(hc, d_ctr, sizeof hc, cudaMemcpyDeviceToHost, st));
        CI(cudaStreamSynchronize(st));
        if (stats) {
            stats->doc_len = nd; stats->trained_positions = (int64_t)hc[0]; stats->pairs = (int64_t)hc[1]; stats->node_visits = (int64_t)hc[2];
            stats->algorithmic_bytes = 2.0 * D * 4.0 * ((double)hc[1] + (double)hc[2]); stats->ms_device = ms_total; stats->launches = launches;
        }
    }
done:
#undef CI
#undef CN
    for (void* p : {(void*)d_tok, (void*)d_doc, (void*)d_keep, (void*)d_cnt64, (void*)d_bad, (void*)d_nd, (void*)d_misc, (void*)d_z, (void*)d_wsc, (void*)d_skey, (void*)d_skey2,
                    (void*)d_sid, (void*)d_sid2, (void*)d_parent, (void*)d_code, (void*)d_nval, (void*)d_poff, (void*)d_nsc, (void*)d_pnode, (void*)d_pcode, (void*)d_ctr,
                    (void*)d_syn0, (void*)d_syn1, (void*)d_lr, d_tmp}) if (p) cudaFree(p);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return rc;
}
}  // namespace

int ctr_i2v_train(const ctr_i2v_config* cfg, const int32_t* tokens, int64_t n, int32_t V, float* emb_out, ctr_i2v_stats* stats) {
    return i2v_train_impl(cfg, tokens, n, V, emb_out, stats, nullptr);
}

// BASELINE configs[4] (item2vec over 8 GPUs): one process per GPU, each with its own shard of the item stream.
int ctr_i2v_train_dist(const ctr_i2v_config* cfg, const int32_t* tokens_shard, int64_t n_shard, int32_t V, float* emb_out, ctr_i2v_stats* stats,
                       int32_t rank, int32_t world, const void* nccl_id, int32_t id_bytes, int64_t sync_every) {
    if (world < 1 || rank < 0 || rank >= world) return set_err(nullptr, CTR_EINVAL, "bad rank/world");
    if (world == 1) return i2v_train_impl(cfg, tokens_shard, n_shard, V, emb_out, stats, nullptr);
    if (!nccl_id || id_bytes != 128 || !cfg) return set_err(nullptr, CTR_EINVAL, "unique id must be 128 bytes");
    std::string err;
    if (!nccl_load(&err)) return set_err(nullptr, CTR_ECOMM, "%s", err.c_str());
    if (cudaSetDevice(cfg->device) != cudaSuccess) return set_err(nullptr, CTR_ECUDA, "cudaSetDevice(%d)", cfg->device);
    I2vDist d; d.rank = rank; d.world = world; d.sync_every = sync_every > 0 ? (long)sync_every : 4L << 20;
    Uid u; memcpy(&u, nccl_id, 128);
    int r = g_nccl.CommInitRank(&d.nccl, world, u, rank);
    if (r != 0) return set_err(nullptr, CTR_ECOMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    const int rc = i2v_train_impl(cfg, tokens_shard, n_shard, V, emb_out, stats, &d);
    g_nccl.CommDestroy(d.nccl);
    return rc;
}

// ---- model/mlp: the float64 MLP classifier (row a11) ---------------------------------------------------------
struct ctr_mlp {
    ctr_mlp_config cfg{};
    Mlp64Dims d{};
    long np = 0;
    int dev = 0, num_sms = 148;
    cudaStream_t stream = nullptr;
    mutable std::string err;
    std::mutex mu;
    double *params = nullptr, *grads = nullptr, *ms = nullptr, *vs = nullptr;
    double *act[kMlpMaxLayers] = {}, *del[kMlpMaxLayers] = {};       // [rows_cap, units[l]]
    long rows_cap = 0;
    double *X64 = nullptr; float* Y32 = nullptr; long* idx = nullptr; size_t x_cap = 0, y_cap = 0, idx_cap = 0;
    double* scal = nullptr;                                           // [0] batch loss, [1] epoch accumulator
    double adam_calls = 0;                                            // AdamOptimizer64.t
    bool fitted = false;
};
namespace {
int mlp_err(const ctr_mlp* m, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (m) m->err = buf; else g_create_error = buf;
    return code;
}
#define MCU(m, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return mlp_err(m, e_ == cudaErrorMemoryAllocation ? CTR_ENOMEM : CTR_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
int mlp_grid(const ctr_mlp* m, long work) { return (int)std::max<long>(1, std::min<long>((work + 255) / 256, (long)m->num_sms * 8)); }
int mlp_rows(ctr_mlp* m, long rows) {
    if (m->rows_cap >= rows) return CTR_OK;
    for (int l = 1; l < m->d.n_layers; l++) {
        if (m->act[l]) cudaFree(m->act[l]);
        if (m->del[l]) cudaFree(m->del[l]);
        m->act[l] = m->del[l] = nullptr;
        MCU(m, cudaMalloc(&m->act[l], sizeof(double) * (size_t)rows * m->d.units[l]));
        MCU(m, cudaMalloc(&m->del[l], sizeof(double) * (size_t)rows * m->d.units[l]));
    }
    m->rows_cap = rows;
    return CTR_OK;
}
// f32 host matrix → f64 device matrix (mlp.go:47-52 / :17-29), through a bounded f32 staging buffer
int mlp_upload_x(ctr_mlp* m, const float* X, int64_t n, int nin) {
    const size_t need = (size_t)n * nin;
    if (m->x_cap < need) { if (m->X64) cudaFree(m->X64); m->X64 = nullptr; MCU(m, cudaMalloc(&m->X64, need * sizeof(double))); m->x_cap = need; }
    const size_t chunk = std::min<size_t>(need, (size_t)16 << 20);
    float* st = nullptr;
    MCU(m, cudaMalloc(&st, chunk * sizeof(float)));
    int rc = CTR_OK;
    for (size_t o = 0; o < need && rc == CTR_OK; o += chunk) {
        const size_t c = std::min(chunk, need - o);
        if (cudaMemcpyAsync(st, X + o, c * sizeof(float), cudaMemcpyHostToDevice, m->stream) != cudaSuccess) { rc = mlp_err(m, CTR_ECUDA, "upload X: %s", cudaGetErrorString(cudaGetLastError())); break; }
        k_f32_to_f64<<<mlp_grid(m, (long)c), 256, 0, m->stream>>>(st, m->X64 + o, (long)c);
        if (cudaStreamSynchronize(m->stream) != cudaSuccess) rc = mlp_err(m, CTR_ECUDA, "upload X: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(st);
    return rc;
}
// forwardPass (basemlp64.go:259-275) over m rows; layer 0 reads X64 through idx (may be null = identity) from row0
void mlp_forward(ctr_mlp* m, const long* idx, long row0, int rows) {
    const int L = m->d.n_layers, nin = m->d.units[0];
    for (int l = 0; l + 1 < L; l++)
        k_mlp64_layer<<<mlp_grid(m, (long)rows * m->d.units[l + 1]), 256, 0, m->stream>>>(m->params, m->d, l, m->X64 + (idx ? 0 : row0 * nin), nin, idx,
                                                                                      l == 0 ? nullptr : m->act[l], m->act[l + 1], rows);
}
}  // namespace

void ctr_mlp_config_default(ctr_mlp_config* c, int32_t n_features) {
    memset(c, 0, sizeof *c);
    c->n_layers = 3; c->units[0] = n_features; c->units[1] = 100; c->units[2] = 1;      // HiddenLayerSizes {100}, basemlp64.go:239
    c->hidden_act = CTR_MLP_RELU; c->batch = 200; c->max_iter = 200; c->n_iter_no_change = 10; c->shuffle = 1;   // :229-254
    c->adaptive = 0; c->warm_start = 0; c->seed = 0; c->device = 0;
    c->alpha = 1e-4; c->lr_init = 1e-3; c->beta1 = 0.9; c->beta2 = 0.999; c->eps = 1e-8; c->tol = 1e-4;
}
const char* ctr_mlp_last_error(const ctr_mlp* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

int ctr_mlp_create(const ctr_mlp_config* cfg, ctr_mlp** out) {
    if (!cfg || !out) return mlp_err(nullptr, CTR_EINVAL, "null argument");
    *out = nullptr;
    const ctr_mlp_config& c = *cfg;
    if (c.n_layers < 2 || c.n_layers > kMlpMaxLayers) return mlp_err(nullptr, CTR_EINVAL, "n_layers must be in [2, %d]", kMlpMaxLayers);
    for (int i = 0; i < c.n_layers; i++) if (c.units[i] < 1) return mlp_err(nullptr, CTR_EINVAL, "layer %d has %d units", i, c.units[i]);
    if (c.units[c.n_layers - 1] != 1) return mlp_err(nullptr, CTR_EINVAL, "binary classifier: one output unit (basemlp64.go:423-425)");
    if (c.hidden_act < CTR_MLP_RELU || c.hidden_act > CTR_MLP_IDENTITY || c.max_iter < 1) return mlp_err(nullptr, CTR_EINVAL, "bad activation / max_iter");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return mlp_err(nullptr, CTR_ENODEV, "no CUDA device: this engine has no CPU fallback"); }
    if (c.device < 0 || c.device >= ndev) return mlp_err(nullptr, CTR_ENODEV, "device %d of %d", c.device, ndev);
    cudaDeviceProp prop{}; cudaGetDeviceProperties(&prop, c.device);
    if (prop.major != 10) return mlp_err(nullptr, CTR_ENODEV, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
    ctr_mlp* m = new ctr_mlp();
    m->cfg = c; m->dev = c.device; m->num_sms = prop.multiProcessorCount;
    m->d.n_layers = c.n_layers; m->d.hidden_act = c.hidden_act;
    long off = 0;
    for (int l = 0; l < c.n_layers; l++) { m->d.units[l] = c.units[l]; if (l + 1 < c.n_layers) { m->d.off[l] = off; off += (long)(1 + c.units[l]) * c.units[l + 1]; } }
    m->np = off;
    bool ok = cudaSetDevice(m->dev) == cudaSuccess && cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (double** p : {&m->params, &m->grads, &m->ms, &m->vs}) ok = ok && cudaMalloc(p, sizeof(double) * (size_t)m->np) == cudaSuccess && cudaMemset(*p, 0, sizeof(double) * (size_t)m->np) == cudaSuccess;
    ok = ok && cudaMalloc(&m->scal, 2 * sizeof(double)) == cudaSuccess;
    if (!ok) { mlp_err(nullptr, CTR_ECUDA, "ctr_mlp_create: %s", cudaGetErrorString(cudaGetLastError())); ctr_mlp_destroy(m); return CTR_ECUDA; }
    *out = m;
    return CTR_OK;
}

void ctr_mlp_destroy(ctr_mlp* m) {
    if (!m) return;
    cudaSetDevice(m->dev);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (void* p : {(void*)m->params, (void*)m->grads, (void*)m->ms, (void*)m->vs, (void*)m->X64, (void*)m->Y32, (void*)m->idx, (void*)m->scal}) if (p) cudaFree(p);
    for (int l = 0; l < kMlpMaxLayers; l++) { if (m->act[l]) cudaFree(m->act[l]); if (m->del[l]) cudaFree(m->del[l]); }
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int ctr_mlp_get_params(ctr_mlp* m, double* params, int64_t cap, int64_t* np) {
    if (!m) return CTR_EINVAL;
    std::lock_guard<std::mutex> lk(m->mu);
    if (np) *np = m->np;
    if (!params) return CTR_OK;
    if (cap < m->np) return mlp_err(m, CTR_EINVAL, "parameter buffer too small: need %ld", m->np);
    MCU(m, cudaSetDevice(m->dev));
    MCU(m, cudaMemcpyAsync(params, m->params, sizeof(double) * (size_t)m->np, cudaMemcpyDeviceToHost, m->stream));
    MCU(m, cudaStreamSynchronize(m->stream));
    return CTR_OK;
}
int ctr_mlp_set_params(ctr_mlp* m, const double* params, int64_t np) {
    if (!m || !params) return CTR_EINVAL;
    std::lock_guard<std::mutex> lk(m->mu);
    if (np != m->np) return mlp_err(m, CTR_EINVAL, "expected %ld packed parameters, got %lld", m->np, (long long)np);
    MCU(m, cudaSetDevice(m->dev));
    MCU(m, cudaMemcpyAsync(m->params, params, sizeof(double) * (size_t)m->np, cudaMemcpyHostToDevice, m->stream));
    MCU(m, cudaStreamSynchronize(m->stream));
    m->fitted = true;
    return CTR_OK;
}

int ctr_mlp_fit(ctr_mlp* m, const float* X, const float* Y, int64_t n, int32_t xcols, int32_t* n_iter, double* loss_curve) {
    if (!m || !X || !Y || n < 1) return mlp_err(m, CTR_EINVAL, "bad fit arguments");
    std::lock_guard<std::mutex> lk(m->mu);
    const ctr_mlp_config& c = m->cfg;
    const int nin = m->d.units[0];
    if (xcols != nin) return mlp_err(m, CTR_EINVAL, "X has %d columns, the network %d inputs", xcols, nin);
    MCU(m, cudaSetDevice(m->dev));
    long bs = c.batch;
    if (bs <= 0) bs = n < 200 ? n : 200; else if (bs > n) bs = n;                       // basemlp64.go:517-527
    RET(mlp_upload_x(m, X, n, nin));
    if (m->y_cap < (size_t)n) { if (m->Y32) cudaFree(m->Y32); m->Y32 = nullptr; MCU(m, cudaMalloc(&m->Y32, sizeof(float) * (size_t)n)); m->y_cap = (size_t)n; }
    if (m->idx_cap < (size_t)n) { if (m->idx) cudaFree(m->idx); m->idx = nullptr; MCU(m, cudaMalloc(&m->idx, sizeof(long) * (size_t)n)); m->idx_cap = (size_t)n; }
    MCU(m, cudaMemcpyAsync(m->Y32, Y, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    RET(mlp_rows(m, bs));
    if (!c.warm_start || !m->fitted) {
        // initialize, basemlp64.go:459-476: every packed element (intercepts too) = U[0,1)·sqrt(f/(fi+fo)), f = 6, or 2 for a
        // logistic hidden activation — non-negative; counter RNG instead of the reference's time-seeded source
        std::vector<double> p((size_t)m->np);
        long pos = 0;
        for (int l = 0; l + 1 < m->d.n_layers; l++) {
            const int fi = m->d.units[l], fo = m->d.units[l + 1];
            const double bound = std::sqrt((c.hidden_act == CTR_MLP_LOGISTIC ? 2.0 : 6.0) / (double)(fi + fo));
            const long end = pos + (long)(1 + fi) * fo;
            for (; pos < end; pos++) p[(size_t)pos] = (double)(mix64(c.seed, 7, (uint64_t)pos) >> 11) * (1.0 / 9007199254740992.0) * bound;
        }
        MCU(m, cudaMemcpyAsync(m->params, p.data(), sizeof(double) * (size_t)m->np, cudaMemcpyHostToDevice, m->stream));
        MCU(m, cudaStreamSynchronize(m->stream));
    }
    MCU(m, cudaMemsetAsync(m->ms, 0, sizeof(double) * (size_t)m->np, m->stream));        // a fresh optimizer per fit (:541-556)
    MCU(m, cudaMemsetAsync(m->vs, 0, sizeof(double) * (size_t)m->np, m->stream));
    m->adam_calls = 0;
    std::vector<long> idx((size_t)n);
    for (int64_t i = 0; i < n; i++) idx[(size_t)i] = i;
    double lr_init = c.lr_init, lr_last = c.lr_init, best = INFINITY;
    int no_improve = 0, it = 0;
    const int L = m->d.n_layers;
    for (it = 0; it < c.max_iter;) {
        if (c.shuffle) {                                                 // rand.Shuffle (:788): Fisher–Yates from the top
            for (long i = n - 1; i > 0; i--) { const long j = (long)(mix64(c.seed, 100u + (uint32_t)it, (uint64_t)i) % (uint64_t)(i + 1)); std::swap(idx[(size_t)i], idx[(size_t)j]); }
        }
        MCU(m, cudaMemcpyAsync(m->idx, idx.data(), sizeof(long) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
        MCU(m, cudaMemsetAsync(m->scal, 0, 2 * sizeof(double), m->stream));
        for (long b0 = 0; b0 < n; b0 += bs) {                            // :791-808
            const int rows = (int)std::min<long>(bs, n - b0);
            const long* bi = m->idx + b0;
            mlp_forward(m, bi, 0, rows);
            k_mlp64_loss<<<1, 256, 0, m->stream>>>(m->params, m->d, m->act[L - 1], m->Y32, bi, rows, c.alpha, m->del[L - 1], m->scal, m->scal + 1);
            for (int l = L - 2; l >= 0; l--) {                           // backprop :382-402
                k_mlp64_grads<<<mlp_grid(m, (long)(m->d.units[l] + 1) * m->d.units[l + 1]), 256, 0, m->stream>>>(
                    m->params, m->grads, m->d, l, m->X64, nin, bi, l == 0 ? nullptr : m->act[l], m->del[l + 1], rows, c.alpha);
                if (l >= 1) k_mlp64_delta<<<mlp_grid(m, (long)rows * m->d.units[l]), 256, 0, m->stream>>>(m->params, m->d, l, m->act[l], m->del[l + 1], m->del[l], rows);
            }
            k_mlp64_adam<<<mlp_grid(m, m->np), 256, 0, m->stream>>>(m->params, m->grads, m->ms, m->vs, m->np, m->adam_calls, lr_init, c.beta1, c.beta2, c.eps);
            m->adam_calls += 1;
        }
        it++;
        double sc[2] = {0, 0};
        MCU(m, cudaMemcpyAsync(sc, m->scal, sizeof sc, cudaMemcpyDeviceToHost, m->stream));
        MCU(m, cudaStreamSynchronize(m->stream));
        MCU(m, cudaGetLastError());
        const double loss = sc[1] / (double)n;                           // :810-811
        if (loss_curve) loss_curve[it - 1] = loss;
        {   // the learning rate the optimizer last used (:1088): decides "adaptive" stopping (:1059-1064)
            const double e = m->adam_calls * (double)m->np;
            lr_last = lr_init * std::sqrt(1.0 - std::pow(c.beta2, e)) / (1.0 - std::pow(c.beta1, e));
        }
        if (loss > best - c.tol) no_improve++; else no_improve = 0;      // :886-892
        if (loss < best) best = loss;
        if (no_improve > c.n_iter_no_change) {                           // :826-840
            if (!c.adaptive) break;                                      // triggerStopping :1053-1058
            if (lr_last <= 1e-6) break;
            lr_init *= .8;
            no_improve = 0;
        }
    }
    if (n_iter) *n_iter = it;
    m->fitted = true;
    return CTR_OK;
}

int ctr_mlp_predict(ctr_mlp* m, const float* X, int64_t n, int32_t xcols, float* out) {
    if (!m || !X || !out || n < 1) return mlp_err(m, CTR_EINVAL, "bad predict arguments");
    std::lock_guard<std::mutex> lk(m->mu);
    if (!m->fitted) return mlp_err(m, CTR_ESTATE, "predict before fit");
    const int nin = m->d.units[0], L = m->d.n_layers;
    if (xcols != nin) return mlp_err(m, CTR_EINVAL, "X has %d columns, the network %d inputs", xcols, nin);
    MCU(m, cudaSetDevice(m->dev));
    RET(mlp_upload_x(m, X, n, nin));
    const long chunk = 8192;
    RET(mlp_rows(m, std::min<long>(chunk, n)));
    float* d_out = nullptr;
    MCU(m, cudaMalloc(&d_out, sizeof(float) * (size_t)std::min<long>(chunk, n)));
    int rc = CTR_OK;
    for (long r0 = 0; r0 < n && rc == CTR_OK; r0 += chunk) {            // predictProbas :897-931: raw probabilities for the f64 class
        const int rows = (int)std::min<long>(chunk, n - r0);
        mlp_forward(m, nullptr, r0, rows);
        k_f64_to_f32<<<mlp_grid(m, rows), 256, 0, m->stream>>>(m->act[L - 1], d_out, rows);
        if (cudaMemcpyAsync(out + r0, d_out, sizeof(float) * (size_t)rows, cudaMemcpyDeviceToHost, m->stream) != cudaSuccess ||
            cudaStreamSynchronize(m->stream) != cudaSuccess) rc = mlp_err(m, CTR_ECUDA, "predict: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d_out);
    return rc;
}

int ctr_comm_unique_id(void* id_out, int32_t* id_bytes) { return comm_unique_id(id_out, id_bytes); }
int ctr_comm_init(ctr_handle* h, const void* id, int32_t id_bytes) {
    if (!h || !id) return set_err(h, CTR_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(h, cudaSetDevice(h->dev));
    return comm_init(h, id, id_bytes);
}

}  // extern "C"
