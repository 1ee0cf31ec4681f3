// comm_impl.cuh — included by engine.cu after ctr_handle and step_core are defined.
#pragma once
#include <dlfcn.h>

namespace {

// ---- the NCCL entry points we use, resolved at run time -------------------------------------------
struct Uid { char internal[128]; };      // ncclUniqueId
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Uid /* by value */, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
enum { kNcclInt32 = 2, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

NcclApi g_nccl;

bool nccl_load(std::string* err) {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* lib = nullptr;
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { *err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { *err = std::string("missing NCCL symbol ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return true;
}

#define NC(h, call) do { int r_ = (call); if (r_ != 0) return set_err(h, CTR_ECOMM, "%s: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

// ---- exchange-plan kernels -----------------------------------------------------------------------------
// lookup p = b*(S+1)+slot (slot S = target item); owner = row % world
__global__ void k_owner_count(const int* __restrict__ hist, const int* __restrict__ item_row, int S, int B, int world, int* __restrict__ cnt) {
    __shared__ int sc[64];
    if (threadIdx.x < 64) sc[threadIdx.x] = 0;
    __syncthreads();
    const long n = (long)B * (S + 1);
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        const int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
        const int row = sl < S ? hist[(long)b * S + sl] : item_row[b];
        if (row >= 0) atomicAdd(&sc[row % world], 1);
    }
    __syncthreads();
    if (threadIdx.x < world && sc[threadIdx.x]) atomicAdd(cnt + threadIdx.x, sc[threadIdx.x]);
}
// assigns every lookup its position in the owner-bucketed send order
__global__ void k_owner_fill(const int* __restrict__ hist, const int* __restrict__ item_row, int S, int B, int world,
                             int* __restrict__ cursor, int* __restrict__ send_rows, int* __restrict__ slot_hist, int* __restrict__ slot_item) {
    const long n = (long)B * (S + 1);
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        const int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
        const int row = sl < S ? hist[(long)b * S + sl] : item_row[b];
        int pos = -1;
        // one atomic per (warp, owner): lanes that target the same owner claim a contiguous run together
        const int own = row >= 0 ? row % world : -1;
        const unsigned peers = __match_any_sync(__activemask(), own);
        if (row >= 0) {
            const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
            int base = 0;
            if (lane == leader) base = atomicAdd(cursor + own, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            pos = base + __popc(peers & ((1u << lane) - 1u));
            send_rows[pos] = row / world;
        }
        if (sl < S) slot_hist[(long)b * S + sl] = pos; else slot_item[b] = pos;
    }
}
// owner side: out[i] = table[rows[i]]   (one float4 per thread)
__global__ void k_gather_local(const int* __restrict__ rows, long n, const float* __restrict__ table, long lde, int D, float* __restrict__ out) {
    const int q4 = D / 4;
    const long total = n * q4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long rix = i / q4; const int c = (int)(i % q4) * 4;
        *reinterpret_cast<float4*>(out + rix * D + c) = ldg4_stream(table + (long)rows[rix] * lde + c);
    }
}
// owner side: table[rows[i]] += g[i]   (already scaled by -lr/world)
__global__ void k_scatter_local(const int* __restrict__ rows, long n, const float* __restrict__ g, float* __restrict__ table, long lde, int D) {
    const int q4 = D / 4;
    const long total = n * q4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long rix = i / q4; const int c = (int)(i % q4) * 4;
        red_add4(table + (long)rows[rix] * lde + c, *reinterpret_cast<const float4*>(g + rix * D + c));
    }
}

// ---- de-duplicated plan: mark → exclusive scan → emit --------------------------------------------------
__global__ void k_dd_mark(const int* __restrict__ hist, const int* __restrict__ item_row, int S, int B, int world, long Imax, int* __restrict__ flags) {
    const long n = (long)B * (S + 1);
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        const int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
        const int row = sl < S ? hist[(long)b * S + sl] : item_row[b];
        if (row >= 0) {
            int* f = flags + (long)(row % world) * Imax + row / world;
            if (*reinterpret_cast<volatile int*>(f) == 0) *f = 1;      // popular rows: read-mostly instead of a store storm on one line
        }
    }
}
__global__ void k_dd_counts(const int* __restrict__ pos, long Imax, int world, int* __restrict__ cnt) {
    const int j = threadIdx.x;
    if (j < world) cnt[j] = pos[(long)(j + 1) * Imax] - pos[(long)j * Imax];
}
__global__ void k_dd_emit(const int* __restrict__ flags, const int* __restrict__ pos, long Q, long Imax, int* __restrict__ send_rows) {
    for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < Q; q += (long)gridDim.x * blockDim.x)
        if (flags[q]) send_rows[pos[q]] = (int)(q % Imax);
}
__global__ void k_dd_slots(const int* __restrict__ hist, const int* __restrict__ item_row, int S, int B, int world, long Imax,
                           const int* __restrict__ pos, int* __restrict__ slot_hist, int* __restrict__ slot_item) {
    const long n = (long)B * (S + 1);
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        const int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
        const int row = sl < S ? hist[(long)b * S + sl] : item_row[b];
        const int slot = row >= 0 ? pos[(long)(row % world) * Imax + row / world] : -1;
        if (sl < S) slot_hist[(long)b * S + sl] = slot; else slot_item[b] = slot;
    }
}

int comm_ensure(ctr_handle* h, size_t L, size_t nrecv) {
    Comm& cm = h->comm;
    const int D = h->cfg.D, W = cm.world;
    if (!cm.d_cnt) {
        RET(dalloc(h, &cm.d_cnt, (size_t)W)); RET(dalloc(h, &cm.d_cursor, (size_t)W)); RET(dalloc(h, &cm.d_rcnt, (size_t)W));
        const long I = (long)h->tab_rows[CTR_TABLE_ITEM_EMB];
        cm.Imax = (I + W - 1) / W; cm.Q = cm.Imax * W;
        cm.dedup = cm.Q <= ((long)32 << 20) && !getenv("CTR_NO_DEDUP");
        if (cm.dedup) {
            RET(dalloc(h, &cm.flags, (size_t)cm.Q + 1)); RET(dalloc(h, &cm.pos, (size_t)cm.Q + 1));
            cub::DeviceScan::ExclusiveSum(nullptr, cm.scan_tmp_bytes, cm.flags, cm.pos, (int)(cm.Q + 1), h->stream);
            CU(h, cudaMalloc(&cm.scan_tmp, cm.scan_tmp_bytes));
            // replica accumulators for the per-row gradients (popular rows would otherwise serialise in L2)
            cm.cap_U = (size_t)std::min<long>((long)L, cm.Q);
            cm.reps = (int)std::min<size_t>(16, std::max<size_t>(1, ((size_t)64 << 20) / (cm.cap_U * D * sizeof(float))));
            if (cm.reps > 1) RET(dalloc(h, &cm.rep_acc, (size_t)cm.reps * cm.cap_U * D));
        }
    }
    if (cm.cap_L < L) {
        for (void* p : {(void*)cm.send_rows, (void*)cm.slot_hist, (void*)cm.slot_item, (void*)cm.rows_local, (void*)cm.grad_local}) if (p) cudaFree(p);
        RET(dalloc(h, &cm.send_rows, L)); RET(dalloc(h, &cm.slot_hist, L)); RET(dalloc(h, &cm.slot_item, L));
        RET(dalloc(h, &cm.rows_local, L * D)); RET(dalloc(h, &cm.grad_local, L * D));
        cm.cap_L = L;
    }
    if (cm.cap_recv < nrecv) {
        if (cm.recv_rows) cudaFree(cm.recv_rows);
        if (cm.rows_out) cudaFree(cm.rows_out);
        const size_t cap = nrecv + nrecv / 4 + 1024;
        RET(dalloc(h, &cm.recv_rows, cap)); RET(dalloc(h, &cm.rows_out, cap * D));
        cm.cap_recv = cap;
    }
    return CTR_OK;
}

// Buckets the batch's lookups by owner, exchanges ids, gathers at the owners and brings the rows back:
// afterwards cm.rows_local[cm.slot_*] holds every row of the local batch.
int comm_fetch_rows(ctr_handle* h, const int* d_item, const int* d_hist, int B) {
    Comm& cm = h->comm;
    const int S = h->cfg.S, D = h->cfg.D, W = cm.world;
    const size_t L = (size_t)B * (S + 1);
    RET(comm_ensure(h, L, cm.cap_recv));
    const int grid = std::min<int>((int)((L + 255) / 256), h->num_sms * 8);
    if (cm.dedup) {
        CU(h, cudaMemsetAsync(cm.flags, 0, sizeof(int) * ((size_t)cm.Q + 1), h->stream));
        RET(launch(h, "shard_dedup_mark", [&] { k_dd_mark<<<grid, 256, 0, h->stream>>>(d_hist, d_item, S, B, W, cm.Imax, cm.flags); }));
        RET(launch(h, "cub_exclusive_scan", [&] { cub::DeviceScan::ExclusiveSum(cm.scan_tmp, cm.scan_tmp_bytes, cm.flags, cm.pos, (int)(cm.Q + 1), h->stream); }));
        RET(launch(h, "shard_dedup_counts", [&] { k_dd_counts<<<1, 64, 0, h->stream>>>(cm.pos, cm.Imax, W, cm.d_cnt); }));
    } else {
        CU(h, cudaMemsetAsync(cm.d_cnt, 0, sizeof(int) * W, h->stream));
        RET(launch(h, "shard_owner_count", [&] { k_owner_count<<<grid, 256, 0, h->stream>>>(d_hist, d_item, S, B, W, cm.d_cnt); }));
    }
    // counts to every owner (one int each way), then both count vectors to the host to size the exchange
    NC(h, g_nccl.GroupStart());
    for (int j = 0; j < W; j++) {
        NC(h, g_nccl.Send(cm.d_cnt + j, 1, kNcclInt32, j, cm.nccl, h->stream));
        NC(h, g_nccl.Recv(cm.d_rcnt + j, 1, kNcclInt32, j, cm.nccl, h->stream));
    }
    NC(h, g_nccl.GroupEnd());
    CU(h, cudaMemcpyAsync(cm.h_scnt, cm.d_cnt, sizeof(int) * W, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(cm.h_rcnt, cm.d_rcnt, sizeof(int) * W, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cm.h_soff[0] = cm.h_roff[0] = 0;
    for (int j = 0; j < W; j++) { cm.h_soff[j + 1] = cm.h_soff[j] + cm.h_scnt[j]; cm.h_roff[j + 1] = cm.h_roff[j] + cm.h_rcnt[j]; }
    RET(comm_ensure(h, L, (size_t)cm.h_roff[W]));
    if (cm.dedup) {
        RET(launch(h, "shard_dedup_emit", [&] {
            k_dd_emit<<<std::min<long>((cm.Q + 255) / 256, (long)h->num_sms * 8), 256, 0, h->stream>>>(cm.flags, cm.pos, cm.Q, cm.Imax, cm.send_rows);
        }));
        RET(launch(h, "shard_dedup_slots", [&] { k_dd_slots<<<grid, 256, 0, h->stream>>>(d_hist, d_item, S, B, W, cm.Imax, cm.pos, cm.slot_hist, cm.slot_item); }));
    } else {
        CU(h, cudaMemcpyAsync(cm.d_cursor, cm.h_soff, sizeof(int) * W, cudaMemcpyHostToDevice, h->stream));
        RET(launch(h, "shard_owner_fill", [&] {
            k_owner_fill<<<grid, 256, 0, h->stream>>>(d_hist, d_item, S, B, W, cm.d_cursor, cm.send_rows, cm.slot_hist, cm.slot_item);
        }));
    }
    // ids to the owners
    NC(h, g_nccl.GroupStart());
    for (int j = 0; j < W; j++) {
        if (cm.h_scnt[j]) NC(h, g_nccl.Send(cm.send_rows + cm.h_soff[j], (size_t)cm.h_scnt[j], kNcclInt32, j, cm.nccl, h->stream));
        if (cm.h_rcnt[j]) NC(h, g_nccl.Recv(cm.recv_rows + cm.h_roff[j], (size_t)cm.h_rcnt[j], kNcclInt32, j, cm.nccl, h->stream));
    }
    NC(h, g_nccl.GroupEnd());
    // owners gather, rows come back in the requester's send order
    const long nrecv = cm.h_roff[W];
    if (nrecv > 0)
        RET(launch(h, "shard_gather_local", [&] {
            k_gather_local<<<std::min<long>((nrecv * (D / 4) + 255) / 256, (long)h->num_sms * 16), 256, 0, h->stream>>>(
                cm.recv_rows, nrecv, h->tab[CTR_TABLE_ITEM_EMB], h->tab_ld[CTR_TABLE_ITEM_EMB], D, cm.rows_out);
        }));
    NC(h, g_nccl.GroupStart());
    for (int j = 0; j < W; j++) {
        if (cm.h_rcnt[j]) NC(h, g_nccl.Send(cm.rows_out + (size_t)cm.h_roff[j] * D, (size_t)cm.h_rcnt[j] * D, kNcclFloat32, j, cm.nccl, h->stream));
        if (cm.h_scnt[j]) NC(h, g_nccl.Recv(cm.rows_local + (size_t)cm.h_soff[j] * D, (size_t)cm.h_scnt[j] * D, kNcclFloat32, j, cm.nccl, h->stream));
    }
    NC(h, g_nccl.GroupEnd());
    cm.bytes_sent += (double)(cm.h_soff[W] - cm.h_scnt[cm.rank]) * 4 + (double)(cm.h_roff[W] - cm.h_rcnt[cm.rank]) * D * 4;
    return CTR_OK;
}

RowSrc comm_src(ctr_handle* h, const int* d_user, const int* d_item, int B) {
    Comm& cm = h->comm;
    RowSrc r{};
    r.emb = cm.rows_local; r.lde = h->cfg.D;               // the received rows act as the table, slots as row ids
    r.ufeat = h->tab[CTR_TABLE_USER_FEAT]; r.ldu = h->tab_ld[CTR_TABLE_USER_FEAT];
    r.ifeat = h->tab[CTR_TABLE_ITEM_FEAT]; r.ldi = h->tab_ld[CTR_TABLE_ITEM_FEAT];
    r.user_row = d_user; r.item_row = cm.slot_item; r.hist = cm.slot_hist; r.item_feat_row = d_item;
    r.dense = 0; r.nvalid = B;
    return r;
}

int comm_check(ctr_handle* h) {
    const ctr_config& c = h->cfg;
    if (!h->comm.ready) return set_err(h, CTR_ESTATE, "world=%d but ctr_comm_init was not called", h->comm.world);
    if (c.D % 4) return set_err(h, CTR_EINVAL, "sharded tables need D %% 4 == 0");
    if (!h->tab[CTR_TABLE_ITEM_EMB] || h->tab_width[CTR_TABLE_ITEM_EMB] != c.D) return set_err(h, CTR_ESTATE, "ITEM_EMB shard not uploaded");
    if (c.uP > 0 && !h->tab[CTR_TABLE_USER_FEAT]) return set_err(h, CTR_ESTATE, "USER_FEAT not uploaded");
    if (c.cF > 0 && !h->tab[CTR_TABLE_ITEM_FEAT]) return set_err(h, CTR_ESTATE, "ITEM_FEAT not uploaded");
    return CTR_OK;
}

}  // namespace

// dense gradients + the batch cost: sum over ranks (every rank then takes the identical Adam step)
static int comm_allreduce_grads(ctr_handle* h, float* extra, size_t extra_n) {
    Comm& cm = h->comm;
    if (!cm.ready) return set_err(h, CTR_ESTATE, "world=%d but ctr_comm_init was not called", cm.world);
    NC(h, g_nccl.GroupStart());
    if (extra && extra_n) NC(h, g_nccl.AllReduce(extra, extra, extra_n, kNcclFloat32, kNcclSum, cm.nccl, h->stream));
    const int nt = h->cfg.model == CTR_MODEL_YOUTUBE ? 3 : 4;
    for (int i = 0; i < nt; i++) NC(h, g_nccl.AllReduce(h->G[i], h->G[i], h->wsize[i], kNcclFloat32, kNcclSum, cm.nccl, h->stream));
    NC(h, g_nccl.AllReduce(h->d_cost, h->d_cost, 1, kNcclFloat64, kNcclSum, cm.nccl, h->stream));
    NC(h, g_nccl.GroupEnd());
    return CTR_OK;
}

static int comm_train_step(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, const float* d_label, int32_t B) {
    RET(comm_check(h));
    Comm& cm = h->comm;
    const ctr_config& c = h->cfg;
    const int D = c.D, W = cm.world;
    RET(comm_fetch_rows(h, d_item, d_hist, B));
    RowSrc r = comm_src(h, d_user, d_item, B);
    const bool learn = c.table_opt != CTR_TABLE_FROZEN;
    if (learn) CU(h, cudaMemsetAsync(cm.grad_local, 0, (size_t)cm.h_soff[W] * D * sizeof(float), h->stream));
    StepOpts o; o.training = true; o.update = true; o.d_label = d_label;
    o.comm = true; o.scatter_base = cm.grad_local; o.grad_scale = 1.0f / (float)W; o.adam_batch = B * W;
    if (cm.dedup && cm.reps > 1) { o.rep_acc = cm.rep_acc; o.rep_rows = cm.h_soff[W]; o.rep_n = cm.reps; }
    RET(step_core(h, r, B, o));
    if (learn) {
        // row gradients go home: the reverse of the row exchange, then the owners apply them
        NC(h, g_nccl.GroupStart());
        for (int j = 0; j < W; j++) {
            if (cm.h_scnt[j]) NC(h, g_nccl.Send(cm.grad_local + (size_t)cm.h_soff[j] * D, (size_t)cm.h_scnt[j] * D, kNcclFloat32, j, cm.nccl, h->stream));
            if (cm.h_rcnt[j]) NC(h, g_nccl.Recv(cm.rows_out + (size_t)cm.h_roff[j] * D, (size_t)cm.h_rcnt[j] * D, kNcclFloat32, j, cm.nccl, h->stream));
        }
        NC(h, g_nccl.GroupEnd());
        const long nrecv = cm.h_roff[W];
        if (nrecv > 0)
            RET(launch(h, "shard_scatter_local", [&] {
                k_scatter_local<<<std::min<long>((nrecv * (D / 4) + 255) / 256, (long)h->num_sms * 16), 256, 0, h->stream>>>(
                    cm.recv_rows, nrecv, cm.rows_out, h->tab[CTR_TABLE_ITEM_EMB], h->tab_ld[CTR_TABLE_ITEM_EMB], D);
            }));
        cm.bytes_sent += (double)(cm.h_soff[W] - cm.h_scnt[cm.rank]) * D * 4;
    }
    return CTR_OK;
}

static int comm_predict(ctr_handle* h, const int32_t* d_user, const int32_t* d_item, const int32_t* d_hist, int32_t B, float* d_out) {
    RET(comm_check(h));
    RET(comm_fetch_rows(h, d_item, d_hist, B));
    RowSrc r = comm_src(h, d_user, d_item, B);
    StepOpts o;
    RET(step_core(h, r, B, o));
    if (d_out) CU(h, cudaMemcpyAsync(d_out, h->P, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return CTR_OK;
}

static int comm_unique_id(void* id_out, int32_t* id_bytes) {
    std::string err;
    if (!id_out || !id_bytes || *id_bytes < 128) return CTR_EINVAL;
    if (!nccl_load(&err)) { g_create_error = err; return CTR_ECOMM; }
    Uid u;
    if (g_nccl.GetUniqueId(&u) != 0) return CTR_ECOMM;
    memcpy(id_out, &u, 128); *id_bytes = 128;
    return CTR_OK;
}

static int comm_init(ctr_handle* h, const void* id, int32_t id_bytes) {
    std::string err;
    if (id_bytes != 128) return set_err(h, CTR_EINVAL, "unique id must be 128 bytes");
    if (h->comm.world < 2) return set_err(h, CTR_EINVAL, "ctr_comm_init with world=%d", h->comm.world);
    if (h->comm.world > 64) return set_err(h, CTR_EINVAL, "world > 64 unsupported");
    if (!nccl_load(&err)) return set_err(h, CTR_ECOMM, "%s", err.c_str());
    Uid u; memcpy(&u, id, 128);
    NC(h, g_nccl.CommInitRank(&h->comm.nccl, h->comm.world, u, h->comm.rank));
    h->comm.ready = true;
    return CTR_OK;
}

static void comm_destroy(ctr_handle* h) {
    Comm& cm = h->comm;
    for (void* p : {(void*)cm.flags, (void*)cm.pos, cm.scan_tmp, (void*)cm.rep_acc}) if (p) cudaFree(p);
    for (void* p : {(void*)cm.d_cnt, (void*)cm.d_cursor, (void*)cm.d_rcnt, (void*)cm.send_rows, (void*)cm.slot_hist, (void*)cm.slot_item,
                    (void*)cm.recv_rows, (void*)cm.rows_out, (void*)cm.rows_local, (void*)cm.grad_local, (void*)cm.table_grad}) if (p) cudaFree(p);
    if (cm.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(cm.nccl);
    cm.nccl = nullptr; cm.ready = false;
}
