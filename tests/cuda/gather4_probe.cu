// Micro-benchmark: does a TMA producer (cp.async.bulk.tensor.2d ... tile::gather4, 4 table rows per instruction, mbarrier
// complete_tx) feeding consumer warps through a shared-memory ring beat the one-warp-per-sample LDG.128 design for the
// embedding gather + cosine gate?  (VERDICT r1 #6.)  256-byte rows, uniform random ids, table >> L2.
//   ldg      baseline: k_attn_fwd_idx's mapping (4 lanes x 4 float4 per row, 8 rows per step), gate math per row
//   tma<S,C> producer warp + C consumer warps per block, S ring slots of one sample (52 rows = 13 gather4) each
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/_build/gather4_probe tests/cuda/gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int D = 64, S = 51, SP = 52;          // rows per sample, padded to a multiple of 4 (the pad id repeats row 0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 ldg4s(const float* p) {
    float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---- baseline: one warp per sample, LDG.128, cosine gate against the sample's last row ("item")
__global__ void __launch_bounds__(128, 8) k_ldg(const float* __restrict__ emb, const int* __restrict__ idx, float* __restrict__ out, int B) {
    constexpr int LPR = 4, VPL = 4, RPW = 8;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * 4;
    for (int b = blockIdx.x * 4 + (threadIdx.x >> 5); b < B; b += nw) {
        const int* ids = idx + (long)b * SP;
        int i0 = lane < S ? ids[lane] : -1, i1 = lane + 32 < S ? ids[lane + 32] : -1;
        const int item = __shfl_sync(0xffffffffu, i1, (S - 1) & 31);
        float4 v[VPL], acc[VPL]; float ny2 = 0.f;
#pragma unroll
        for (int q = 0; q < VPL; q++) { v[q] = ldg4s(emb + (long)item * D + (q * LPR + lir) * 4); ny2 += dot4(v[q], v[q]); acc[q] = make_float4(0, 0, 0, 0); }
        for (int o = 1; o < LPR; o <<= 1) ny2 += __shfl_xor_sync(0xffffffffu, ny2, o);
        const float ny = sqrtf(ny2);
        for (int s0 = 0; s0 < S - 1; s0 += RPW) {
            const int s = s0 + sub;
            const int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
            const int id = s < S - 1 ? (s < 32 ? a0 : a1) : -1;
            float4 u[VPL]; float dot = 0.f, nx2 = 0.f;
#pragma unroll
            for (int q = 0; q < VPL; q++) { u[q] = id >= 0 ? ldg4s(emb + (long)id * D + (q * LPR + lir) * 4) : make_float4(0, 0, 0, 0); }
#pragma unroll
            for (int q = 0; q < VPL; q++) { dot += dot4(u[q], v[q]); nx2 += dot4(u[q], u[q]); }
            for (int o = 1; o < LPR; o <<= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); nx2 += __shfl_xor_sync(0xffffffffu, nx2, o); }
            const float a = sigm((dot / (sqrtf(nx2) * ny + 1e-8f) + 1.f) * 0.5f);
#pragma unroll
            for (int q = 0; q < VPL; q++) { acc[q].x += a * u[q].x; acc[q].y += a * u[q].y; acc[q].z += a * u[q].z; acc[q].w += a * u[q].w; }
        }
#pragma unroll
        for (int q = 0; q < VPL; q++) {
            for (int o = LPR; o < 32; o <<= 1) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
            }
            if (sub == 0) *reinterpret_cast<float4*>(out + (long)b * D + (q * LPR + lir) * 4) = acc[q];
        }
    }
}

// ---- TMA gather4 producer / consumer ring
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(cnt) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 :: "r"(dst), "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

template <int STAGES, int CONS>
__global__ void __launch_bounds__(32 * (CONS + 1)) k_tma(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, float* __restrict__ out, int B) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int SLOT = SP * D * 4;                                   // 13 312 bytes per sample
    const uint32_t base = (s32(smem) + 127u) & ~127u;
    const uint32_t bars = base + STAGES * SLOT;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int per = (B + gridDim.x - 1) / gridDim.x;
    const int b0 = blockIdx.x * per, b1 = min(B, b0 + per);
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0, b = b0; b < b1; b++, i++) {
                const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1u;
                mbar_wait(empty(s), ph ^ 1u);
                mbar_expect_tx(full(s), SLOT);
                const int4* ids = reinterpret_cast<const int4*>(idx + (long)b * SP);
#pragma unroll
                for (int g = 0; g < SP / 4; g++) { const int4 r = __ldg(ids + g); gather4(base + s * SLOT + g * 1024, &map, full(s), r.x, r.y, r.z, r.w); }
            }
        }
    } else {
        // consumer warp c takes samples i with i % CONS == c; 8 lanes x 2 float4 per row, 4 rows per step from shared memory
        const int c = warp - 1;
        constexpr int LPR = 4, VPL = 4, RPW = 8;
        const int lir = lane % LPR, sub = lane / LPR;
        for (int i = 0, b = b0; b < b1; b++, i++) {
            const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1u;
            if (i % CONS != c) continue;
            mbar_wait(full(s), ph);
            const float* rows = reinterpret_cast<const float*>(smem + (base - s32(smem)) + s * SLOT);
            float4 v[VPL], acc[VPL]; float ny2 = 0.f;
#pragma unroll
            for (int q = 0; q < VPL; q++) { v[q] = *reinterpret_cast<const float4*>(rows + (S - 1) * D + (q * LPR + lir) * 4); ny2 += dot4(v[q], v[q]); acc[q] = make_float4(0, 0, 0, 0); }
            for (int o = 1; o < LPR; o <<= 1) ny2 += __shfl_xor_sync(0xffffffffu, ny2, o);
            const float ny = sqrtf(ny2);
            for (int s0 = 0; s0 < S - 1; s0 += RPW) {
                const int r = s0 + sub;
                float4 u[VPL]; float dot = 0.f, nx2 = 0.f;
#pragma unroll
                for (int q = 0; q < VPL; q++) u[q] = r < S - 1 ? *reinterpret_cast<const float4*>(rows + r * D + (q * LPR + lir) * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < VPL; q++) { dot += dot4(u[q], v[q]); nx2 += dot4(u[q], u[q]); }
                for (int o = 1; o < LPR; o <<= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); nx2 += __shfl_xor_sync(0xffffffffu, nx2, o); }
                const float a = sigm((dot / (sqrtf(nx2) * ny + 1e-8f) + 1.f) * 0.5f);
#pragma unroll
                for (int q = 0; q < VPL; q++) { acc[q].x += a * u[q].x; acc[q].y += a * u[q].y; acc[q].z += a * u[q].z; acc[q].w += a * u[q].w; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(s));                           // slot free
#pragma unroll
            for (int q = 0; q < VPL; q++) {
                for (int o = LPR; o < 32; o <<= 1) {
                    acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                    acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
                }
                if (sub == 0) *reinterpret_cast<float4*>(out + (long)b * D + (q * LPR + lir) * 4) = acc[q];
            }
        }
    }
}

template <int STAGES, int CONS>
float run_tma(const CUtensorMap& map, const int* d_idx, float* d_out, int B, int blocks_per_sm, const char* tag, double bytes, const float* ref_out) {
    const size_t smem = (size_t)STAGES * SP * D * 4 + 16 * STAGES + 256;
    CK(cudaFuncSetAttribute(k_tma<STAGES, CONS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = 148 * blocks_per_sm;
    k_tma<STAGES, CONS><<<grid, 32 * (CONS + 1), smem>>>(map, d_idx, d_out, B);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: launch failed: %s\n", tag, cudaGetErrorString(e)); exit(1); }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; i++) k_tma<STAGES, CONS><<<grid, 32 * (CONS + 1), smem>>>(map, d_idx, d_out, B);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    std::vector<float> h(1024), r(1024);
    CK(cudaMemcpy(h.data(), d_out, 4096, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r.data(), ref_out, 4096, cudaMemcpyDeviceToHost));
    double maxd = 0; for (int i = 0; i < 1024; i++) maxd = fmax(maxd, fabs((double)h[i] - r[i]));
    printf("%-28s smem %6zu B, %d blocks/SM: %.3f ms  %.0f GB/s   (max |diff| vs ldg %.2e)\n", tag, smem, blocks_per_sm, ms, bytes / ms / 1e6, maxd);
    return ms;
}

int main() {
    const long R = 12'500'000; const int B = 65536;
    float* emb; CK(cudaMalloc(&emb, R * D * 4));
    { std::vector<float> h((size_t)1 << 22); std::mt19937 g(3); for (auto& v : h) v = (float)((g() & 0xffff) / 65536.0 - 0.5);
      for (long o = 0; o < R * D; o += (long)h.size()) CK(cudaMemcpy(emb + o, h.data(), std::min<long>((long)h.size(), R * D - o) * 4, cudaMemcpyHostToDevice)); }
    std::vector<int> hidx((size_t)B * SP); std::mt19937_64 g(1);
    for (int b = 0; b < B; b++) { for (int s = 0; s < S; s++) hidx[(size_t)b * SP + s] = (int)(g() % R); hidx[(size_t)b * SP + S] = hidx[(size_t)b * SP]; }
    int* d_idx; CK(cudaMalloc(&d_idx, hidx.size() * 4)); CK(cudaMemcpy(d_idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    float *o_ref, *o_tma; CK(cudaMalloc(&o_ref, (long)B * D * 4)); CK(cudaMalloc(&o_tma, (long)B * D * 4));
    const double bytes = (double)B * S * D * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_ldg<<<148 * 16, 128>>>(emb, d_idx, o_ref, B); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 10; i++) k_ldg<<<148 * 16, 128>>>(emb, d_idx, o_ref, B); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    printf("%-28s %.3f ms  %.0f GB/s\n", "ldg 4x4, 8 blocks/SM", ms, bytes / ms / 1e6);
    // tensor map over the table: [R rows, 64 cols] fp32, box {64, 1} — gather4 picks 4 independent rows per instruction
    CUtensorMap map; cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)R}; cuuint64_t gstr[1] = {(cuuint64_t)D * 4}; cuuint32_t estr[2] = {1, 1};
    CUresult r = CUDA_ERROR_UNKNOWN;
    for (cuuint32_t brows : {1u, 4u}) {
        cuuint32_t box[2] = {(cuuint32_t)D, brows};
        r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, emb, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled box {64,%u}: %d\n", brows, (int)r);
        if (r == CUDA_SUCCESS) break;
    }
    if (r != CUDA_SUCCESS) return 1;
    run_tma<4, 4>(map, d_idx, o_tma, B, 2, "tma gather4 4 slots, 4 cons", bytes, o_ref);
    run_tma<4, 4>(map, d_idx, o_tma, B, 3, "tma gather4 4 slots, 4 cons", bytes, o_ref);
    run_tma<8, 4>(map, d_idx, o_tma, B, 2, "tma gather4 8 slots, 4 cons", bytes, o_ref);
    run_tma<6, 6>(map, d_idx, o_tma, B, 2, "tma gather4 6 slots, 6 cons", bytes, o_ref);
    run_tma<16, 8>(map, d_idx, o_tma, B, 1, "tma gather4 16 slots, 8 cons", bytes, o_ref);
    run_tma<8, 8>(map, d_idx, o_tma, B, 2, "tma gather4 8 slots, 8 cons", bytes, o_ref);
    printf("done\n");
    return 0;
}
