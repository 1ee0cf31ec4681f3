// comm.cuh — multi-GPU state of one rank: ITEM_EMB rows are sharded
// owner(row) = row % world; indices travel to the owners with an NCCL all-to-all, rows come back,
// row gradients return the same way, dense gradients are all-reduced.  NCCL is loaded with dlopen
// so that a process that already carries torch's bundled libnccl shares it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ctr_handle;

namespace ctr {

struct Comm {
    int rank = 0, world = 1;
    void* lib = nullptr;          // dlopen handle of libnccl
    void* nccl = nullptr;         // ncclComm_t
    bool ready = false;
    // exchange buffers (device)
    int* send_rows = nullptr;  int* recv_rows = nullptr;      // requested row ids, bucketed by owner
    int* send_cnt = nullptr;   int* recv_cnt = nullptr;       // [world] counts (device)
    float* send_buf = nullptr; float* recv_buf = nullptr;     // row payloads
    int* slot_of = nullptr;                                   // [B*(S+1)] position of each lookup in recv order
    size_t cap_rows = 0;
    float* rows_local = nullptr;                              // [B, (S+1)*D] gathered rows in sample order
    float* grads_local = nullptr;
};

}  // namespace ctr

static int comm_allreduce_grads(ctr_handle* h);
static int comm_train_step(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, const float* d_label, int32_t B);
static int comm_predict(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, int32_t B, float* d_out);
static int comm_unique_id(void* id_out, int32_t* id_bytes);
static int comm_init(ctr_handle* h, const void* id, int32_t id_bytes);
static void comm_destroy(ctr_handle* h);
