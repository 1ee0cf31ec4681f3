// Micro-benchmark (needs 2 GPUs with peer access; `gpurun --gpus 2`): what does NVLink 5 / NVSwitch give the
// row-sharded embedding path when kernels touch the owner's table directly (no NCCL, no staging)?
//   gather   256-byte rows, uniform random ids, LDG.128 4 lanes x 4 float4 per row (k_attn_fwd_idx's mapping),
//            remote fraction 0 / 0.5 / 0.875 / 1
//   red4     red.global.add.v4.f32 of 256-byte rows into the peer's table (k_attn_bwd_idx's fused scatter)
//   red1     the same bytes as scalar red.global.add.f32
//   st4      st.global.v4.f32 of 256-byte rows into a sequential staging buffer on the peer
//   bidir    both GPUs gather from / red into each other at the same time
//   barrier  flag round trip of a two-GPU device-side barrier (st.release.sys / ld.acquire.sys)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/_build/peer_probe tests/cuda/peer_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <random>
#include <unistd.h>
#include <sys/wait.h>
#include <sys/socket.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int D = 64, S = 51;

__device__ __forceinline__ float4 ldg4s(const float* p) {
    float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ void red4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory"); }

struct Tabs { const float* t[2]; };
struct TabsW { float* t[2]; };

// ids carry the owner in bit 0 (row % 2) like the engine; local row = id >> 1
__global__ void __launch_bounds__(128, 8) k_gather(Tabs tb, const int* __restrict__ idx, float* __restrict__ out, int B) {
    constexpr int LPR = 4, VPL = 4, RPW = 8;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * 4;
    for (int b = blockIdx.x * 4 + (threadIdx.x >> 5); b < B; b += nw) {
        int i0 = lane < S ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
        float4 acc[VPL];
#pragma unroll
        for (int q = 0; q < VPL; q++) acc[q] = make_float4(0, 0, 0, 0);
        for (int s0 = 0; s0 < S; s0 += RPW) {
            int s = s0 + sub;
            int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
            int id = s < S ? (s < 32 ? a0 : a1) : -1;
            const float* base = tb.t[id & 1] + (long)(id >> 1) * D + lir * 4;
#pragma unroll
            for (int q = 0; q < VPL; q++) {
                float4 u = id >= 0 ? ldg4s(base + q * LPR * 4) : make_float4(0, 0, 0, 0);
                acc[q].x += u.x; acc[q].y += u.y; acc[q].z += u.z; acc[q].w += u.w;
            }
        }
        if (sub == 0)
#pragma unroll
            for (int q = 0; q < VPL; q++) *reinterpret_cast<float4*>(out + (long)b * D + (q * LPR + lir) * 4) = acc[q];
    }
}

// MODE 0: red.v4, 1: scalar red, 2: st.v4 into a sequential buffer (slot = b*S+s)
template <int MODE>
__global__ void __launch_bounds__(128, 8) k_scatter(TabsW tb, const int* __restrict__ idx, int B) {
    constexpr int LPR = 8, VPL = 2, RPW = 4;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * 4;
    for (int b = blockIdx.x * 4 + (threadIdx.x >> 5); b < B; b += nw) {
        int i0 = lane < S ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
        for (int s0 = 0; s0 < S; s0 += RPW) {
            int s = s0 + sub;
            int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
            int id = s < S ? (s < 32 ? a0 : a1) : -1;
            if (id < 0) continue;
            float4 v = make_float4(1e-9f * lane, 1e-9f, 2e-9f, 3e-9f);
            if (MODE == 2) {
                float* dst = tb.t[id & 1] + ((long)b * S + s) * D + lir * 4;
#pragma unroll
                for (int q = 0; q < VPL; q++) *reinterpret_cast<float4*>(dst + q * LPR * 4) = v;
            } else {
                float* dst = tb.t[id & 1] + (long)(id >> 1) * D + lir * 4;
#pragma unroll
                for (int q = 0; q < VPL; q++) {
                    if (MODE == 0) red4(dst + q * LPR * 4, v);
                    else { red1(dst + q * LPR * 4, v.x); red1(dst + q * LPR * 4 + 1, v.y); red1(dst + q * LPR * 4 + 2, v.z); red1(dst + q * LPR * 4 + 3, v.w); }
                }
            }
        }
    }
}

// two-GPU barrier: write epoch to the peer's flag, wait for mine
__global__ void k_barrier(volatile unsigned long long* mine, unsigned long long* peer, unsigned long long epoch) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(peer), "l"(epoch) : "memory");
        unsigned long long v = 0; long spins = 0;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v < epoch && ++spins < (1l << 28));
    }
}

#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* m_ = nullptr; cuGetErrorString(r_, &m_); printf("driver error %s at %s:%d\n", m_ ? m_ : "?", __FILE__, __LINE__); exit(1); } } while (0)

static void send_fd(int sock, int fd) {
    char dummy = 'x'; struct iovec iov = {&dummy, 1};
    char ctl[CMSG_SPACE(sizeof(int))]; memset(ctl, 0, sizeof ctl);
    struct msghdr msg; memset(&msg, 0, sizeof msg);
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
    struct cmsghdr* c = CMSG_FIRSTHDR(&msg); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    if (sendmsg(sock, &msg, 0) != 1) { perror("sendmsg"); exit(1); }
}
static int recv_fd(int sock) {
    char dummy; struct iovec iov = {&dummy, 1};
    char ctl[CMSG_SPACE(sizeof(int))]; memset(ctl, 0, sizeof ctl);
    struct msghdr msg; memset(&msg, 0, sizeof msg);
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
    if (recvmsg(sock, &msg, 0) != 1) { perror("recvmsg"); exit(1); }
    struct cmsghdr* c = CMSG_FIRSTHDR(&msg); int fd = -1; memcpy(&fd, CMSG_DATA(c), sizeof(int)); return fd;
}
// physical allocation through the VMM API (2 MB granules), exportable as a POSIX file descriptor
static CUmemGenericAllocationHandle vmm_create(int dev, size_t* bytes) {
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0; CKD(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    *bytes = (*bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h; CKD(cuMemCreate(&h, *bytes, &prop, 0));
    return h;
}
static void* vmm_map(CUmemGenericAllocationHandle h, size_t bytes, int dev) {
    CUdeviceptr va = 0; CKD(cuMemAddressReserve(&va, bytes, 0, 0, 0)); CKD(cuMemMap(va, bytes, 0, h, 0));
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = dev; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CKD(cuMemSetAccess(va, bytes, &acc, 1));
    return (void*)va;
}

// argv: [rows per GPU] [ipc|vmm]   — GPU 1's table lives in a forked child process and is mapped here with
// "ipc": cudaIpcOpenMemHandle (legacy CUDA IPC), "vmm": cuMemCreate + POSIX-fd export / cuMemImportFromShareableHandle +
// cuMemMap (2 MB pages) — instead of in-process peer access
int main(int argc, char** argv) {
    const long R_arg = argc > 1 ? atol(argv[1]) : 12'500'000 / 2;
    const bool vmm = argc > 2 && !strcmp(argv[2], "vmm");
    const bool ipc = (argc > 2 && !strcmp(argv[2], "ipc")) || vmm;
    int pfd[2] = {-1, -1}, cfd[2] = {-1, -1}, sp[2] = {-1, -1}; pid_t child = 0;
    cudaIpcMemHandle_t ipc_handle;
    int vmm_fd = -1; size_t vmm_bytes = (size_t)R_arg * 64 * 4;
    if (ipc) {
        if (pipe(pfd) || pipe(cfd) || socketpair(AF_UNIX, SOCK_STREAM, 0, sp)) return 1;
        child = fork();
        if (child == 0) {      // child: owns GPU 1's table, exports it, waits for the parent to finish
            CK(cudaSetDevice(1)); CK(cudaFree(0));
            if (vmm) {
                size_t bytes = vmm_bytes;
                CUmemGenericAllocationHandle h = vmm_create(1, &bytes);
                void* t = vmm_map(h, bytes, 1);
                CK(cudaMemset(t, 0, bytes)); CK(cudaDeviceSynchronize());
                int fd = -1; CKD(cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
                if (write(pfd[1], &bytes, sizeof bytes) != (ssize_t)sizeof bytes) return 1;
                send_fd(sp[1], fd);
            } else {
                float* t = nullptr; CK(cudaMalloc(&t, R_arg * 64 * 4)); CK(cudaMemset(t, 0, R_arg * 64 * 4)); CK(cudaDeviceSynchronize());
                cudaIpcMemHandle_t hd; CK(cudaIpcGetMemHandle(&hd, t));
                if (write(pfd[1], &hd, sizeof hd) != (ssize_t)sizeof hd) return 1;
            }
            char c; if (read(cfd[0], &c, 1) != 1) return 1;
            return 0;
        }
        if (vmm) { if (read(pfd[0], &vmm_bytes, sizeof vmm_bytes) != (ssize_t)sizeof vmm_bytes) return 1; vmm_fd = recv_fd(sp[0]); }
        else if (read(pfd[0], &ipc_handle, sizeof ipc_handle) != (ssize_t)sizeof ipc_handle) return 1;
    }
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs, have %d\n", nd); return 0; }
    int can01 = 0, can10 = 0; CK(cudaDeviceCanAccessPeer(&can01, 0, 1)); CK(cudaDeviceCanAccessPeer(&can10, 1, 0));
    printf("peer access 0->1 %d, 1->0 %d\n", can01, can10);
    if (!can01 || !can10) return 0;
    const long R = R_arg;                     // local rows per GPU (default 1.6 GB each; >> L2)
    printf("rows per GPU %ld (%.1f GB), %s\n", R, R * 256.0 / 1e9, vmm ? "GPU 1's table: VMM allocation of another process, imported through a POSIX fd and cuMemMap'ed" : ipc ? "GPU 1's table mapped through CUDA IPC from another process" : "in-process peer access");
    const int B = 65536;
    float* tab[2]; float* out[2]; int* idx[2][4]; float* stage[2]; unsigned long long* flag[2];
    cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        if (vmm && d == 1) {
            CK(cudaSetDevice(0));
            CUmemGenericAllocationHandle h; CKD(cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)vmm_fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
            tab[1] = (float*)vmm_map(h, vmm_bytes, 0);
            CK(cudaSetDevice(1));
        } else if (ipc && d == 1) { CK(cudaSetDevice(0)); CK(cudaIpcOpenMemHandle((void**)&tab[1], ipc_handle, cudaIpcMemLazyEnablePeerAccess)); CK(cudaSetDevice(1)); }
        else { CK(cudaMalloc(&tab[d], R * D * 4)); CK(cudaMemset(tab[d], 0, R * D * 4)); }
        CK(cudaMalloc(&out[d], (long)B * D * 4));
        CK(cudaMalloc(&stage[d], (long)B * S * D * 4));
        CK(cudaMalloc(&flag[d], 8)); CK(cudaMemset(flag[d], 0, 8));
        CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
    }
    // id sets: remote fraction f from GPU d's point of view: owner bit = (1-d) with prob f
    const double fr[4] = {0.0, 0.5, 0.875, 1.0};
    std::mt19937_64 rng(1);
    for (int d = 0; d < 2; d++)
        for (int k = 0; k < 4; k++) {
            std::vector<int> h((long)B * S);
            for (auto& v : h) {
                long lr = (long)(rng() % R); int own = ((rng() >> 11) * (1.0 / 9007199254740992.0) < fr[k]) ? 1 - d : d;
                v = (int)(lr * 2 + own);
            }
            CK(cudaSetDevice(d)); CK(cudaMalloc(&idx[d][k], h.size() * 4)); CK(cudaMemcpy(idx[d][k], h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        }
    const double bytes = (double)B * S * D * 4;
    auto time_one = [&](int d, auto&& fn, int reps) {
        CK(cudaSetDevice(d));
        fn(d); CK(cudaStreamSynchronize(st[d]));
        CK(cudaEventRecord(e0[d], st[d]));
        for (int i = 0; i < reps; i++) fn(d);
        CK(cudaEventRecord(e1[d], st[d])); CK(cudaStreamSynchronize(st[d]));
        float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); return ms / reps;
    };
    const int grid = 148 * 16;
    for (int k = 0; k < 4; k++) {
        float ms = time_one(0, [&](int d) { Tabs t{{tab[0], tab[1]}}; k_gather<<<grid, 128, 0, st[d]>>>(t, idx[d][k], out[d], B); }, 10);
        printf("gather  remote %.3f : %.3f ms  %.0f GB/s total, %.0f GB/s over NVLink\n", fr[k], ms, bytes / ms / 1e6, bytes * fr[k] / ms / 1e6);
    }
    for (int k = 0; k < 4; k++) {
        float ms = time_one(0, [&](int d) { TabsW t{{tab[0], tab[1]}}; k_scatter<0><<<grid, 128, 0, st[d]>>>(t, idx[d][k], B); }, 10);
        printf("red4    remote %.3f : %.3f ms  %.0f GB/s payload, %.0f GB/s over NVLink\n", fr[k], ms, bytes / ms / 1e6, bytes * fr[k] / ms / 1e6);
    }
    for (int k = 2; k < 4; k++) {
        float ms = time_one(0, [&](int d) { TabsW t{{tab[0], tab[1]}}; k_scatter<1><<<grid, 128, 0, st[d]>>>(t, idx[d][k], B); }, 5);
        printf("red1    remote %.3f : %.3f ms  %.0f GB/s payload\n", fr[k], ms, bytes / ms / 1e6);
    }
    for (int k = 2; k < 4; k++) {
        float ms = time_one(0, [&](int d) { TabsW t{{stage[0], stage[1]}}; k_scatter<2><<<grid, 128, 0, st[d]>>>(t, idx[d][k], B); }, 10);
        printf("st4     remote %.3f : %.3f ms  %.0f GB/s payload\n", fr[k], ms, bytes / ms / 1e6);
    }
    // bidirectional: both GPUs at once (remote 0.875), wall time over both streams
    for (int mode = 0; mode < 2 && !ipc; mode++) {
        const int reps = 10;
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); }
        for (int i = 0; i < reps; i++)
            for (int d = 0; d < 2; d++) {
                CK(cudaSetDevice(d));
                if (mode == 0) { Tabs t{{tab[0], tab[1]}}; k_gather<<<grid, 128, 0, st[d]>>>(t, idx[d][2], out[d], B); }
                else { TabsW t{{tab[0], tab[1]}}; k_scatter<0><<<grid, 128, 0, st[d]>>>(t, idx[d][2], B); }
            }
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e1[d], st[d])); }
        for (int d = 0; d < 2; d++) {
            CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d]));
            float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); ms /= reps;
            printf("bidir %s gpu%d remote 0.875: %.3f ms  %.0f GB/s payload, %.0f GB/s over NVLink each way\n", mode == 0 ? "gather" : "red4  ", d, ms, bytes / ms / 1e6, bytes * 0.875 / ms / 1e6);
        }
    }
    // gather on gpu0 + red4 on gpu1 hitting gpu0... (fwd of one rank overlapping bwd of another) skipped: covered by bidir.
    // barrier round trip
    {
        const int reps = 200;
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        CK(cudaSetDevice(0)); CK(cudaEventRecord(e0[0], st[0]));
        for (int i = 1; i <= reps; i++)
            for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); k_barrier<<<1, 32, 0, st[d]>>>(flag[d], flag[1 - d], (unsigned long long)i); }
        CK(cudaSetDevice(0)); CK(cudaEventRecord(e1[0], st[0]));
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
        float ms; CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
        printf("barrier kernel (launch + flag exchange): %.2f us each\n", ms * 1e3 / reps);
    }
    if (ipc) { char c = 1; if (write(cfd[1], &c, 1) != 1) return 1; int st = 0; waitpid(child, &st, 0); }
    printf("done\n");
    return 0;
}
