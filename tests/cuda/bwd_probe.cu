// Micro-benchmark for the backward attention kernel (re-gather + cosine-gate backward + red.add scatter):
// which lane mapping / occupancy gets closest to the HBM roofline?  Table 12.5M x 64 fp32, uniform ids.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/_build/bwd_probe tests/cuda/bwd_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

constexpr int D = 64, S = 50;

__device__ __forceinline__ float frcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float fsq(float x) { return x > 0.0f ? x * rsqrtf(x) : 0.0f; }
__device__ __forceinline__ float sigf(float x) { float r = frcp(1.0f + __expf(-x)); r = x > 15.0f ? 1.0f : r; return x < -88.0f ? 0.0f : r; }
__device__ __forceinline__ float d4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
template <int W> __device__ __forceinline__ float gsum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void red4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldc4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int LPR, int VPL, int UNR, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_bwd(float* __restrict__ emb, const int* __restrict__ idx, const int* __restrict__ item, const float* __restrict__ att,
      const float* __restrict__ dX, float* __restrict__ datt, float neg_lr, int B) {
    __shared__ float sdatt[64];
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lir = lane % LPR, sub = lane / LPR;
    const int nw = gridDim.x * (NT / 32);
    const float invS = 1.0f / (float)S;
    if (threadIdx.x < 64) sdatt[threadIdx.x] = 0.f;
    __syncthreads();
    for (int b = blockIdx.x * (NT / 32) + (threadIdx.x >> 5); b < B; b += nw) {
        const int i0 = lane < S ? idx[(long)b * S + lane] : -1, i1 = lane + 32 < S ? idx[(long)b * S + lane + 32] : -1;
        const int irow = item[b];
        float4 g[VPL], v[VPL], dvu[VPL]; float ny2 = 0.f;
#pragma unroll
        for (int q = 0; q < VPL; q++) {
            g[q] = __ldg(reinterpret_cast<const float4*>(dX + (long)b * 2 * D + (q * LPR + lir) * 4));
            v[q] = ldc4(emb + (long)irow * D + (q * LPR + lir) * 4);
            ny2 += d4(v[q], v[q]); dvu[q] = make_float4(0, 0, 0, 0);
        }
        ny2 = gsum<LPR>(ny2);
        const float ny = fsq(ny2), rny = ny2 > 0.f ? rsqrtf(ny2) : 0.f;
        float kvsum = 0.f;
        for (int s0 = 0; s0 < S; s0 += UNR * RPW) {
            float4 u[UNR][VPL]; int id[UNR];
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const int s = s0 + j * RPW + sub;
                const int a0 = __shfl_sync(0xffffffffu, i0, s & 31), a1 = __shfl_sync(0xffffffffu, i1, s & 31);
                id[j] = s < S ? (s < 32 ? a0 : a1) : -1;
#pragma unroll
                for (int q = 0; q < VPL; q++) u[j][q] = id[j] >= 0 ? ldc4(emb + (long)id[j] * D + (q * LPR + lir) * 4) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const int s = s0 + j * RPW + sub;
                const float att_s = s < S ? __ldg(att + s) : 0.f;
                float gu = 0.f, dot = 0.f, nx2 = 0.f;
#pragma unroll
                for (int q = 0; q < VPL; q++) { gu += d4(g[q], u[j][q]); dot += d4(u[j][q], v[q]); nx2 += d4(u[j][q], u[j][q]); }
                gu = gsum<LPR>(gu); dot = gsum<LPR>(dot); nx2 = gsum<LPR>(nx2);
                const float nx = fsq(nx2), iden = frcp(nx * ny + 1e-8f), cs = dot * iden, w = (cs + 1.f) * .5f;
                const float a = sigf(w * att_s), dz = gu * invS * a * (1.f - a);
                if (lir == 0 && s < S) atomicAdd(&sdatt[s], dz * w);
                const float cc = .5f * dz * att_s;
                const float c1 = a * invS * neg_lr, c2 = cc * iden, c3 = nx2 > 0.f ? -cc * cs * ny * iden * rsqrtf(nx2) * neg_lr : 0.f;
                kvsum += cc * cs * nx * iden * rny;
                if (id[j] >= 0) {
#pragma unroll
                    for (int q = 0; q < VPL; q++) {
                        const float4 uu = u[j][q];
                        float4 du;
                        du.x = fmaf(c3, uu.x, fmaf(c2 * neg_lr, v[q].x, c1 * g[q].x)); du.y = fmaf(c3, uu.y, fmaf(c2 * neg_lr, v[q].y, c1 * g[q].y));
                        du.z = fmaf(c3, uu.z, fmaf(c2 * neg_lr, v[q].z, c1 * g[q].z)); du.w = fmaf(c3, uu.w, fmaf(c2 * neg_lr, v[q].w, c1 * g[q].w));
                        dvu[q].x = fmaf(c2, uu.x, dvu[q].x); dvu[q].y = fmaf(c2, uu.y, dvu[q].y); dvu[q].z = fmaf(c2, uu.z, dvu[q].z); dvu[q].w = fmaf(c2, uu.w, dvu[q].w);
                        red4(emb + (long)id[j] * D + (q * LPR + lir) * 4, du);
                    }
                }
            }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) kvsum += __shfl_xor_sync(0xffffffffu, kvsum, o);
#pragma unroll
        for (int q = 0; q < VPL; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                dvu[q].x += __shfl_xor_sync(0xffffffffu, dvu[q].x, o); dvu[q].y += __shfl_xor_sync(0xffffffffu, dvu[q].y, o);
                dvu[q].z += __shfl_xor_sync(0xffffffffu, dvu[q].z, o); dvu[q].w += __shfl_xor_sync(0xffffffffu, dvu[q].w, o);
            }
            if (sub == 0) {
                const float4 gi = __ldg(reinterpret_cast<const float4*>(dX + (long)b * 2 * D + D + (q * LPR + lir) * 4));
                float4 dv = make_float4((gi.x + dvu[q].x - kvsum * v[q].x) * neg_lr, (gi.y + dvu[q].y - kvsum * v[q].y) * neg_lr,
                                        (gi.z + dvu[q].z - kvsum * v[q].z) * neg_lr, (gi.w + dvu[q].w - kvsum * v[q].w) * neg_lr);
                red4(emb + (long)irow * D + (q * LPR + lir) * 4, dv);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < S && sdatt[threadIdx.x] != 0.f) atomicAdd(datt + threadIdx.x, sdatt[threadIdx.x]);
}

int main(int argc, char** argv) {
    long I = argc > 1 ? atol(argv[1]) : 12500000; int B = 65536;
    float *emb, *dX, *att, *datt; int *idx, *item;
    cudaMalloc(&emb, (size_t)I * D * 4); cudaMalloc(&idx, (size_t)B * S * 4); cudaMalloc(&item, B * 4);
    cudaMalloc(&dX, (size_t)B * 2 * D * 4); cudaMalloc(&att, 64 * 4); cudaMalloc(&datt, 64 * 4);
    { std::vector<float> r((size_t)1 << 20); for (size_t i = 0; i < r.size(); i++) r[i] = (float)((i * 2654435761u) % 1000) / 1000.0f - 0.5f;
      for (size_t off = 0; off < (size_t)I * D; off += r.size()) cudaMemcpy(emb + off, r.data(), std::min(r.size(), (size_t)I * D - off) * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dX, r.data(), std::min(r.size(), (size_t)B * 2 * D) * 4, cudaMemcpyHostToDevice); for (size_t off = r.size(); off < (size_t)B * 2 * D; off += r.size()) cudaMemcpy(dX + off, r.data(), std::min(r.size(), (size_t)B * 2 * D - off) * 4, cudaMemcpyHostToDevice); }
    std::vector<int> h((size_t)B * S); uint64_t x = 88172645463325252ull;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (int)(x % (uint64_t)I); }
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(item, h.data(), B * 4, cudaMemcpyHostToDevice);
    std::vector<float> ones(64, 1.0f); cudaMemcpy(att, ones.data(), 64 * 4, cudaMemcpyHostToDevice); cudaMemset(datt, 0, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 2.0 * (double)B * (S + 1) * D * 4;
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(e0); for (int i = 0; i < 10; i++) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-28s %.3f ms  %.0f GB/s (RMW bytes)  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    const float nlr = -1e-6f;
    run("b 8x2 unr2 256/2 (engine)", [&] { k_bwd<8, 2, 2, 256, 2><<<148 * 8, 256>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 8x2 unr1 128/8", [&] { k_bwd<8, 2, 1, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 8x2 unr2 128/6", [&] { k_bwd<8, 2, 2, 128, 6><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 4x4 unr1 128/4", [&] { k_bwd<4, 4, 1, 128, 4><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 4x4 unr1 128/5", [&] { k_bwd<4, 4, 1, 128, 5><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 4x4 unr1 128/6", [&] { k_bwd<4, 4, 1, 128, 6><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 16x1 unr2 128/8", [&] { k_bwd<16, 1, 2, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 16x1 unr1 128/8", [&] { k_bwd<16, 1, 1, 128, 8><<<148 * 16, 128>>>(emb, idx, item, att, dX, datt, nlr, B); });
    run("b 8x2 unr1 256/4", [&] { k_bwd<8, 2, 1, 256, 4><<<148 * 8, 256>>>(emb, idx, item, att, dX, datt, nlr, B); });
    return 0;
}
