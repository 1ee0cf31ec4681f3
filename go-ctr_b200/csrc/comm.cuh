// comm.cuh — multi-GPU state of one rank.  ITEM_EMB rows are sharded
// owner(row) = row % world, local row = row / world; the dense weights and the small feature tables
// are replicated.  Per step: the (S+1)*B lookups of the local batch are bucketed by owner on the
// device, their ids travel to the owners (all-to-all), the owners gather the rows and send them back,
// the local attention/MLP kernels run on the received rows, row gradients return the same way and
// the owners apply them; dense gradients are all-reduced so every rank takes the identical Adam step.
// NCCL is loaded with dlopen so that a process that already carries torch's bundled libnccl shares it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ctr_handle;

namespace ctr {

struct Comm {
    int rank = 0, world = 1;
    void* nccl = nullptr;         // ncclComm_t
    bool ready = false;
    // exchange plan of the current batch
    int* d_cnt = nullptr;         // [world] lookups per owner            (device)
    int* d_cursor = nullptr;      // [world] fill cursors                  (device)
    int* d_rcnt = nullptr;        // [world] lookups requested from us     (device)
    int* send_rows = nullptr;     // [L] owner-local row ids, bucketed by owner
    int* slot_hist = nullptr;     // [B,S]  position of each history lookup in send order (-1 = missing)
    int* slot_item = nullptr;     // [B]
    int* recv_rows = nullptr;     // [n_recv] rows requested from this rank
    float* rows_out = nullptr;    // [n_recv, D] gathered rows / received gradients (owner side)
    float* rows_local = nullptr;  // [L, D] rows of the local batch in send order
    float* grad_local = nullptr;  // [L, D] -lr/world * gradient per lookup
    size_t cap_L = 0, cap_recv = 0;
    int h_scnt[64] = {0}, h_rcnt[64] = {0}, h_soff[65] = {0}, h_roff[65] = {0};
    double bytes_sent = 0;        // payload bytes this rank has put on NVLink (rows + gradients + ids)
    // de-duplicated exchange (tables up to 32M rows): each distinct row of the batch crosses NVLink once
    bool dedup = false;
    long Imax = 0, Q = 0;         // q(row) = (row % world) * Imax + row / world, Q = world * Imax
    int* flags = nullptr;         // [Q+1] 1 where the batch touches q
    int* pos = nullptr;           // [Q+1] exclusive scan of flags = slot of q in owner-bucketed send order
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    float* rep_acc = nullptr;     // [reps, cap_U, D] replica accumulators of the per-row gradients
    int reps = 0; size_t cap_U = 0;
    // small ITEM_EMB tables (<= kReplicateBytes) are not sharded at all: every rank keeps the whole table, the
    // batch's row gradients are summed into table_grad [rows, ld], all-reduced together with the dense
    // gradients and applied identically everywhere — no per-step exchange plan, no host sync
    bool replicate = false;
    float* table_grad = nullptr; size_t table_grad_n = 0;
};
constexpr size_t kReplicateBytes = (size_t)32 << 20;

}  // namespace ctr

static int comm_allreduce_grads(ctr_handle* h, float* extra, size_t extra_n);
static int comm_train_step(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, const float* d_label, int32_t B);
static int comm_predict(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, int32_t B, float* d_out);
static int comm_unique_id(void* id_out, int32_t* id_bytes);
static int comm_init(ctr_handle* h, const void* id, int32_t id_bytes);
static void comm_destroy(ctr_handle* h);
