// Micro-benchmark: how fast can one B200 gather 256-byte embedding rows (uniform random ids, table >> L2)?
// Variants of the load path of k_attn_fwd_vec; the "compute" is a plain row sum so memory dominates.
//   v0  LDG.128 to registers, 4 lanes x 4 float4 per row, 2 groups (16 rows) in flight per warp
//   v1  LDG.128 to registers, 8 lanes x 2 float4 per row, 2 groups (8 rows) in flight, more warps
//   v2  cp.async.bulk (TMA 1-D bulk copy, UBLKCP) one per row into a per-warp double buffer, mbarrier
//   v3  cp.async (LDGSTS) 16 B per lane into the same double buffer
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/_build/gather_probe tests/cuda/gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

constexpr int D = 64, S = 51;

__device__ __forceinline__ float4 ldg4s(const float* p) {
    float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int LPR, int VPL, int UNR>
__global__ void __launch_bounds__(256) k_ldg(const float* __restrict__ emb, const int* __restrict__ idx, float* __restrict__ out, int B) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * 8;
    for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < B; b += nw) {
        int i0 = lane < S ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
        float4 acc[VPL];
#pragma unroll
        for (int q = 0; q < VPL; q++) acc[q] = make_float4(0, 0, 0, 0);
        for (int s0 = 0; s0 < S; s0 += UNR * RPW) {
            float4 u[UNR][VPL];
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                int s = s0 + j * RPW + sub;
                int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
                int id = s < S ? (s < 32 ? a0 : a1) : -1;
#pragma unroll
                for (int q = 0; q < VPL; q++) u[j][q] = id >= 0 ? ldg4s(emb + (long)id * D + (q * LPR + lir) * 4) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < UNR; j++)
#pragma unroll
                for (int q = 0; q < VPL; q++) { acc[q].x += u[j][q].x; acc[q].y += u[j][q].y; acc[q].z += u[j][q].z; acc[q].w += u[j][q].w; }
        }
#pragma unroll
        for (int q = 0; q < VPL; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
            }
            if (sub == 0) *reinterpret_cast<float4*>(out + (long)b * D + (q * LPR + lir) * 4) = acc[q];
        }
    }
}

// per-warp double buffer [2][S][D] in smem; MODE 0 = cp.async.bulk + mbarrier, 1 = cp.async 16B (LDGSTS)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_stage(const float* __restrict__ emb, const int* __restrict__ idx, float* __restrict__ out, int B) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* buf = reinterpret_cast<float*>(smem) + (long)w * 2 * S * D;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * S * D * 4) + w * 2;
    if (MODE == 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nw = gridDim.x * WARPS;
    auto issue = [&](int b, int slot) {
        float* dst = buf + (long)slot * S * D;
        if (MODE == 0) {
            // lanes 0..S-1 each issue bulk copies for rows lane, lane+32 (256 B each); lane 0 arms the barrier
            int n = 0;
            int r0 = lane < S ? idx[(long)b * S + lane] : -1, r1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
            n = (r0 >= 0) + (r1 >= 0);
            int tot = n;
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bars[slot])), "r"(tot * D * 4) : "memory");
            __syncwarp();
            if (r0 >= 0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                      :: "r"(s32(dst + lane * D)), "l"(emb + (long)r0 * D), "r"(D * 4), "r"(s32(&bars[slot])) : "memory");
            if (r1 >= 0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                      :: "r"(s32(dst + (lane + 32) * D)), "l"(emb + (long)r1 * D), "r"(D * 4), "r"(s32(&bars[slot])) : "memory");
        } else {
            int i0 = lane < S ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
            // 16 lanes per row, 2 rows per instruction
            for (int s0 = 0; s0 < S; s0 += 2) {
                int s = s0 + (lane >> 4);
                int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
                int id = s < S ? (s < 32 ? a0 : a1) : -1;
                if (id >= 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s32(dst + s * D + (lane & 15) * 4)), "l"(emb + (long)id * D + (lane & 15) * 4) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    int b = blockIdx.x * WARPS + w;
    uint32_t phase[2] = {0, 0};
    if (b < B) issue(b, 0);
    int slot = 0;
    for (; b < B; b += nw, slot ^= 1) {
        int nb = b + nw;
        if (nb < B) issue(nb, slot ^ 1);
        if (MODE == 0) {
            asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" :: "r"(s32(&bars[slot])), "r"(phase[slot]) : "memory");
            phase[slot] ^= 1;
        } else {
            if (nb < B) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        const float* src = buf + (long)slot * S * D;
        float4 acc0 = make_float4(0, 0, 0, 0);
        // lanes 0..15 cover a row; two rows per step
        for (int s0 = 0; s0 < S; s0 += 2) {
            int s = s0 + (lane >> 4);
            if (s < S) { float4 u = *reinterpret_cast<const float4*>(src + s * D + (lane & 15) * 4); acc0.x += u.x; acc0.y += u.y; acc0.z += u.z; acc0.w += u.w; }
        }
        acc0.x += __shfl_xor_sync(0xffffffffu, acc0.x, 16); acc0.y += __shfl_xor_sync(0xffffffffu, acc0.y, 16);
        acc0.z += __shfl_xor_sync(0xffffffffu, acc0.z, 16); acc0.w += __shfl_xor_sync(0xffffffffu, acc0.w, 16);
        if (lane < 16) *reinterpret_cast<float4*>(out + (long)b * D + lane * 4) = acc0;
        __syncwarp();
    }
}


// ---- forward attention math on top of the gather: cosine gate + weighted mean (k_attn_fwd_vec's inner loop)
__device__ __forceinline__ float frcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float fsq(float x) { return x > 0.0f ? x * rsqrtf(x) : 0.0f; }
__device__ __forceinline__ float sigf(float x) { float r = frcp(1.0f + __expf(-x)); r = x > 15.0f ? 1.0f : r; return x < -88.0f ? 0.0f : r; }
__device__ __forceinline__ float d4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
template <int W> __device__ __forceinline__ float gsum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// PIPE = 0: load UNR groups, then compute (current engine structure). PIPE = 1: ping-pong prefetch of one group.
template <int LPR, int VPL, int UNR, int PIPE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_gate(const float* __restrict__ emb, const int* __restrict__ idx, const int* __restrict__ item,
                                                 const float* __restrict__ att, float* __restrict__ out, int B) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * (NT / 32);
    const float invS = 1.0f / 50.0f;
    for (int b = blockIdx.x * (NT / 32) + (threadIdx.x >> 5); b < B; b += nw) {
        const int i0 = lane < 50 ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < 50 ? idx[(long)b * S + lane + 32] : -1;
        const float* ip = emb + (long)item[b] * D;
        float4 v[VPL], acc[VPL]; float ny2 = 0.f;
#pragma unroll
        for (int q = 0; q < VPL; q++) { v[q] = ldg4s(ip + (q * LPR + lir) * 4); ny2 += d4(v[q], v[q]); acc[q] = make_float4(0, 0, 0, 0); }
        const float ny = fsq(gsum<LPR>(ny2));
        auto load = [&](float4 (&u)[VPL], int s) {
            int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
            int id = s < 50 ? (s < 32 ? a0 : a1) : -1;
#pragma unroll
            for (int q = 0; q < VPL; q++) u[q] = id >= 0 ? ldg4s(emb + (long)id * D + (q * LPR + lir) * 4) : make_float4(0, 0, 0, 0);
        };
        auto comp = [&](const float4 (&u)[VPL], int s) {
            float dot = 0.f, nx2 = 0.f;
#pragma unroll
            for (int q = 0; q < VPL; q++) { dot += d4(u[q], v[q]); nx2 += d4(u[q], u[q]); }
            dot = gsum<LPR>(dot); nx2 = gsum<LPR>(nx2);
            const float cs = dot * frcp(fsq(nx2) * ny + 1e-8f);
            const float a = sigf((cs + 1.0f) * 0.5f * (s < 50 ? __ldg(att + s) : 0.0f));
#pragma unroll
            for (int q = 0; q < VPL; q++) { acc[q].x = fmaf(a, u[q].x, acc[q].x); acc[q].y = fmaf(a, u[q].y, acc[q].y); acc[q].z = fmaf(a, u[q].z, acc[q].z); acc[q].w = fmaf(a, u[q].w, acc[q].w); }
        };
        if (PIPE == 0) {
            for (int s0 = 0; s0 < 50; s0 += UNR * RPW) {
                float4 u[UNR][VPL];
#pragma unroll
                for (int j = 0; j < UNR; j++) load(u[j], s0 + j * RPW + sub);
#pragma unroll
                for (int j = 0; j < UNR; j++) comp(u[j], s0 + j * RPW + sub);
            }
        } else {
            float4 ua[VPL], ub[VPL];
            load(ua, sub);
            for (int s0 = 0; s0 < 50; s0 += 2 * RPW) {
                load(ub, s0 + RPW + sub);
                comp(ua, s0 + sub);
                load(ua, s0 + 2 * RPW + sub);
                comp(ub, s0 + RPW + sub);
            }
        }
#pragma unroll
        for (int q = 0; q < VPL; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
            }
            if (sub == 0) *reinterpret_cast<float4*>(out + (long)b * D + (q * LPR + lir) * 4) = make_float4(acc[q].x * invS, acc[q].y * invS, acc[q].z * invS, acc[q].w * invS);
        }
    }
}

// ---- same math, rows staged through a per-warp shared-memory ring with cp.async (LDGSTS): NST-1 groups of
// 8 rows stay in flight per warp at no register cost, and the pipeline runs across sample boundaries.
// Sequence per sample: position 0 = item row, 1..50 = history rows; group g = positions 8g..8g+7 (7 groups).
__device__ __forceinline__ void cp16(void* dst, const void* src, int bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
template <int NST, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_gate_ring(const float* __restrict__ emb, const int* __restrict__ idx, const int* __restrict__ item,
                                                      const float* __restrict__ att, float* __restrict__ out, int B) {
    extern __shared__ float4 ring[];
    constexpr int NG = 7;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lir = lane & 3, sub = lane >> 2;
    const int nw = gridDim.x * (NT / 32), w = blockIdx.x * (NT / 32) + wib;
    float4* wb = ring + wib * NST * 128;
    const float invS = 1.0f / 50.0f;
    auto load_ids = [&](int b, int& p0, int& p1) {
        p0 = p1 = -1;
        if (b < B) { p0 = lane == 0 ? item[b] : idx[(long)b * S + lane - 1]; if (lane + 32 <= 50) p1 = idx[(long)b * S + lane + 31]; }
    };
    const int nsamp = w < B ? (B - w + nw - 1) / nw : 0;
    const int T = nsamp * NG;
    int i0, i1, n0, n1;                       // ids of the sample being issued, and of the one after it
    load_ids(w, i0, i1); load_ids(w + nw, n0, n1);
    int ib = w, ig = 0, kiss = 0;             // issue cursor
    auto issue = [&]() {
        if (kiss < T) {
            const int slot = kiss % NST;
            const int src = ig < 4 ? i0 : i1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rg = 2 * j + (lane >> 4), p = 8 * ig + rg;
                const int id = __shfl_sync(0xffffffffu, src, p & 31);
                const int c = lane & 15;
                cp16(wb + slot * 128 + rg * 16 + (c ^ ((rg & 1) << 2)), emb + (long)(id >= 0 ? id : 0) * D + c * 4, id >= 0 ? 16 : 0);
            }
            if (++ig == NG) { ig = 0; ib += nw; i0 = n0; i1 = n1; load_ids(ib + nw, n0, n1); }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        kiss++;
    };
#pragma unroll
    for (int j = 0; j < NST - 1; j++) issue();
    float4 v[4], acc[4]; float ny = 0.f;
    int cb = w, cg = 0;
    for (int k = 0; k < T; k++) {
        issue();
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
        __syncwarp();
        const float4* sl = wb + (k % NST) * 128;
        if (cg == 0) {
            float ny2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) { v[q] = sl[q * 4 + lir]; ny2 += d4(v[q], v[q]); acc[q] = make_float4(0, 0, 0, 0); }
            ny = fsq(gsum<4>(ny2));
        }
        {
            const int p = 8 * cg + sub, s = p - 1;
            float4 u[4]; float dot = 0.f, nx2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) { u[q] = sl[sub * 16 + ((q * 4 + lir) ^ ((sub & 1) << 2))]; dot += d4(u[q], v[q]); nx2 += d4(u[q], u[q]); }
            dot = gsum<4>(dot); nx2 = gsum<4>(nx2);
            const float cs = dot * frcp(fsq(nx2) * ny + 1e-8f);
            float a = sigf((cs + 1.0f) * 0.5f * ((s >= 0 && s < 50) ? __ldg(att + s) : 0.0f));
            if (s < 0 || s >= 50) a = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++) { acc[q].x = fmaf(a, u[q].x, acc[q].x); acc[q].y = fmaf(a, u[q].y, acc[q].y); acc[q].z = fmaf(a, u[q].z, acc[q].z); acc[q].w = fmaf(a, u[q].w, acc[q].w); }
        }
        if (++cg == NG) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o); acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                    acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, o); acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, o);
                }
                if (sub == 0) *reinterpret_cast<float4*>(out + (long)cb * D + (q * 4 + lir) * 4) = make_float4(acc[q].x * invS, acc[q].y * invS, acc[q].z * invS, acc[q].w * invS);
            }
            cg = 0; cb += nw;
        }
        __syncwarp();
    }
}

int main(int argc, char** argv) {
    long I = argc > 1 ? atol(argv[1]) : 12500000; int B = 65536;
    float* emb; int* idx; float* out;
    cudaMalloc(&emb, (size_t)I * D * 4); cudaMalloc(&idx, (size_t)B * S * 4); cudaMalloc(&out, (size_t)B * D * 4);
    cudaMemset(emb, 0, (size_t)I * D * 4);
    std::vector<int> h((size_t)B * S); uint64_t x = 88172645463325252ull;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (int)(x % (uint64_t)I); }
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)B * S * D * 4;
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(e0); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-28s %.3f ms  %.0f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int g : {148 * 4, 148 * 8, 148 * 16}) {
        printf("grid %d\n", g);
        run("v0 ldg 4x4 unr2", [&] { k_ldg<4, 4, 2><<<g, 256>>>(emb, idx, out, B); });
        run("v0b ldg 4x4 unr1", [&] { k_ldg<4, 4, 1><<<g, 256>>>(emb, idx, out, B); });
        run("v1 ldg 8x2 unr2", [&] { k_ldg<8, 2, 2><<<g, 256>>>(emb, idx, out, B); });
        run("v1b ldg 8x2 unr4", [&] { k_ldg<8, 2, 4><<<g, 256>>>(emb, idx, out, B); });
        run("v1c ldg 16x1 unr4", [&] { k_ldg<16, 1, 4><<<g, 256>>>(emb, idx, out, B); });
    }
    {
        constexpr int W = 8; size_t sm = (size_t)W * 2 * S * D * 4 + W * 16;
        cudaFuncSetAttribute(k_stage<0, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(k_stage<1, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        run("v2 bulk W8 grid148", [&] { k_stage<0, W><<<148, W * 32, sm>>>(emb, idx, out, B); });
        run("v3 ldgsts W8 grid148", [&] { k_stage<1, W><<<148, W * 32, sm>>>(emb, idx, out, B); });
    }
    {
        constexpr int W = 4; size_t sm = (size_t)W * 2 * S * D * 4 + W * 16;
        cudaFuncSetAttribute(k_stage<0, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(k_stage<1, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        run("v2 bulk W4 grid296", [&] { k_stage<0, W><<<296, W * 32, sm>>>(emb, idx, out, B); });
        run("v3 ldgsts W4 grid296", [&] { k_stage<1, W><<<296, W * 32, sm>>>(emb, idx, out, B); });
    }

    {
        int* item; float* att; cudaMalloc(&item, B * 4); cudaMalloc(&att, 64 * 4);
        cudaMemcpy(item, h.data(), B * 4, cudaMemcpyHostToDevice);
        std::vector<float> ones(64, 1.0f); cudaMemcpy(att, ones.data(), 64 * 4, cudaMemcpyHostToDevice);
        // real rows so the math is not degenerate
        { std::vector<float> r((size_t)1 << 20); for (size_t i = 0; i < r.size(); i++) r[i] = (float)((i * 2654435761u) % 1000) / 1000.0f - 0.5f;
          for (size_t off = 0; off < (size_t)I * D; off += r.size()) cudaMemcpy(emb + off, r.data(), std::min(r.size(), (size_t)I * D - off) * 4, cudaMemcpyHostToDevice); }
        printf("--- gather + cosine gate (forward math)\n");
        run("g 4x4 unr2 nopipe 256/1", [&] { k_gate<4, 4, 2, 0, 256, 1><<<148 * 8, 256>>>(emb, idx, item, att, out, B); });
        run("g 4x4 pipe 256/2", [&] { k_gate<4, 4, 1, 1, 256, 2><<<148 * 8, 256>>>(emb, idx, item, att, out, B); });
        run("g 4x4 pipe 128/5", [&] { k_gate<4, 4, 1, 1, 128, 5><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 4x4 pipe 128/6", [&] { k_gate<4, 4, 1, 1, 128, 6><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 8x2 pipe 128/8", [&] { k_gate<8, 2, 1, 1, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 8x2 unr2 nopipe 128/8", [&] { k_gate<8, 2, 2, 0, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 8x2 unr4 nopipe 128/6", [&] { k_gate<8, 2, 4, 0, 128, 6><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 4x4 unr1 nopipe 128/8", [&] { k_gate<4, 4, 1, 0, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        run("g 16x1 unr4 nopipe 128/8", [&] { k_gate<16, 1, 4, 0, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
        {
            auto ringrun = [&](const char* name, auto kern, int nst, int nt, int blocks_per_sm) {
                size_t sm = (size_t)(nt / 32) * nst * 2048;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                run(name, [&] { kern<<<148 * blocks_per_sm, nt, sm>>>(emb, idx, item, att, out, B); });
            };
            ringrun("ring nst3 128/8", k_gate_ring<3, 128, 8>, 3, 128, 8);
            ringrun("ring nst4 128/6", k_gate_ring<4, 128, 6>, 4, 128, 6);
            ringrun("ring nst2 128/8", k_gate_ring<2, 128, 8>, 2, 128, 8);
            ringrun("ring nst3 128/6", k_gate_ring<3, 128, 6>, 3, 128, 6);
            ringrun("ring nst6 128/4", k_gate_ring<6, 128, 4>, 6, 128, 4);
            ringrun("ring nst3 256/4", k_gate_ring<3, 256, 4>, 3, 256, 4);
            ringrun("ring nst3 128/7", k_gate_ring<3, 128, 7>, 3, 128, 7);
            ringrun("ring nst4 128/7", k_gate_ring<4, 128, 7>, 4, 128, 7);
        }
        run("g 2x8 pipe 128/4", [&] { k_gate<2, 8, 1, 1, 128, 4><<<148 * 16, 128>>>(emb, idx, item, att, out, B); });
    }
    return 0;
}
