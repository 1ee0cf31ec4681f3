// mlp.cuh — the dense half of the hot path: the no-bias sigmoid MLP (din.go:307-315,
// dnn.go:172-177), BCE (cost.go:9-17), their backward (what G.Grad builds, model.go:56) and
// gorgonia's AdamSolver.Step (model.go:88,192).  This file is the exact-fp32 engine
// (CTR_GEMM_FP32): a register-tiled FFMA SGEMM with fused epilogues.  The tcgen05 engine
// (umma_gemm.cuh) replaces the two forward GEMMs and the dX/dH GEMMs when enabled and is validated
// against this one.
#pragma once
#include "common.cuh"

namespace ctr {

enum { EPI_STORE = 0, EPI_SIGMOID_DROP = 1, EPI_DSIGMOID = 2, EPI_ATOMIC = 3 };

struct GemmArgs {
    const float* A; long lda;     // TA=false: A[M,K] row-major ; TA=true: stored [K,M]
    const float* B; long ldb;     // TB=false: B[K,N] row-major ; TB=true: stored [N,K]
    float* C; long ldc;           // C[M,N]
    int M, N, K;
    int Nz;                       // columns [N, Nz) of C are written as zeros (keeps K-padding of the next layer clean)
    int kchunk;                   // split-K: each blockIdx.z handles kchunk of K (EPI_ATOMIC)
    // epilogue parameters
    const float* H; long ldh;     // EPI_DSIGMOID: stored post-dropout activation
    float drop_p; uint32_t seed, stream;
};

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB, int EPI>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_sgemm(GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    static_assert(BK % 4 == 0 && BM % 4 == 0 && BN % 4 == 0 && TM % 4 == 0 && TN % 4 == 0, "tile");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int kbeg = 0, kend = g.K;
    if (EPI == EPI_ATOMIC) { kbeg = blockIdx.z * g.kchunk; kend = min(g.K, kbeg + g.kchunk); }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- A tile → As[k][m]
        if (!TA) {
            constexpr int KQ = BK / 4;
            for (int e = tid; e < BM * KQ; e += NT) {
                int m = e / KQ, kq = e % KQ;
                int gm = m0 + m, gk = k0 + kq * 4;
                float4 v = zero4();
                if (gm < g.M) {
                    const float* p = g.A + (long)gm * g.lda + gk;
                    if (gk + 3 < kend) v = ldg4(p);
                    else { if (gk < kend) v.x = __ldg(p); if (gk + 1 < kend) v.y = __ldg(p + 1); if (gk + 2 < kend) v.z = __ldg(p + 2); }
                }
                As[kq * 4 + 0][m] = v.x; As[kq * 4 + 1][m] = v.y; As[kq * 4 + 2][m] = v.z; As[kq * 4 + 3][m] = v.w;
            }
        } else {
            constexpr int MQ = BM / 4;
            for (int e = tid; e < BK * MQ; e += NT) {
                int k = e / MQ, mq = e % MQ;
                int gk = k0 + k, gm = m0 + mq * 4;
                float4 v = zero4();
                if (gk < kend) {
                    const float* p = g.A + (long)gk * g.lda + gm;
                    if (gm + 3 < g.M) v = ldg4(p);
                    else { if (gm < g.M) v.x = __ldg(p); if (gm + 1 < g.M) v.y = __ldg(p + 1); if (gm + 2 < g.M) v.z = __ldg(p + 2); }
                }
                *reinterpret_cast<float4*>(&As[k][mq * 4]) = v;
            }
        }
        // ---- B tile → Bs[k][n]
        if (!TB) {
            constexpr int NQ = BN / 4;
            for (int e = tid; e < BK * NQ; e += NT) {
                int k = e / NQ, nq = e % NQ;
                int gk = k0 + k, gn = n0 + nq * 4;
                float4 v = zero4();
                if (gk < kend) {
                    const float* p = g.B + (long)gk * g.ldb + gn;
                    if (gn + 3 < g.N) v = ldg4(p);
                    else { if (gn < g.N) v.x = __ldg(p); if (gn + 1 < g.N) v.y = __ldg(p + 1); if (gn + 2 < g.N) v.z = __ldg(p + 2); }
                }
                *reinterpret_cast<float4*>(&Bs[k][nq * 4]) = v;
            }
        } else {
            constexpr int KQ = BK / 4;
            for (int e = tid; e < BN * KQ; e += NT) {
                int n = e / KQ, kq = e % KQ;
                int gn = n0 + n, gk = k0 + kq * 4;
                float4 v = zero4();
                if (gn < g.N) {
                    const float* p = g.B + (long)gn * g.ldb + gk;
                    if (gk + 3 < kend) v = ldg4(p);
                    else { if (gk < kend) v.x = __ldg(p); if (gk + 1 < kend) v.y = __ldg(p + 1); if (gk + 2 < kend) v.z = __ldg(p + 2); }
                }
                Bs[kq * 4 + 0][n] = v.x; Bs[kq * 4 + 1][n] = v.y; Bs[kq * 4 + 2][n] = v.z; Bs[kq * 4 + 3][n] = v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
                b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
#pragma unroll
    for (int i = 0; i < TM; i++) {
        int gm = m0 + ty * TM + i;
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int gn = n0 + tx * TN + j;
            float* c = g.C + (long)gm * g.ldc + gn;
            if (gn < g.N) {
                float v = acc[i][j];
                if (EPI == EPI_SIGMOID_DROP) {
                    v = sigmoid32(v) * drop_keep(g.drop_p, g.seed, g.stream, (uint64_t)gm * (uint64_t)g.N + gn);
                    *c = v;
                } else if (EPI == EPI_DSIGMOID) {
                    *c = v * dsigmoid_drop(__ldg(g.H + (long)gm * g.ldh + gn), g.drop_p);
                } else if (EPI == EPI_ATOMIC) {
                    atomicAdd(c, v);
                } else {
                    *c = v;
                }
            } else if (gn < g.Nz && EPI != EPI_ATOMIC) {
                *c = 0.0f;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Output layer + loss + first backward step, one warp per sample (din.go:315, cost.go:9-17):
//   z2 = h1d·w2 ; p = sigmoid(z2) ; cost += y ln p + (1-y) ln(1-p)
//   dz2 = (p - y)/B ; dZ1 = dz2 w2ᵀ ⊙ dsigmoid(h1d) ; dW2 += h1dᵀ dz2
// -------------------------------------------------------------------------------------------------
struct HeadArgs {
    const float* H1d; long ldh; int H1; int H1p;     // [B, ldh], true width H1, zero-fill to H1p
    const float* w2;                                  // [H1]
    const float* y;                                   // [B] labels (null: predict only); rows >= nvalid use label 0
    int B, nvalid;
    float drop_p;
    float* p; float* logit;                           // [B] (logit may be null)
    float* dZ1; long lddz;                            // [B, lddz] (null: predict only)
    float* dW2;                                       // [H1] atomics
    double* cost_sum;                                 // scalar atomic: sum of y ln p + (1-y) ln(1-p)
};

__global__ void __launch_bounds__(256) k_head(HeadArgs a) {
    __shared__ float s_dw2[8][128];     // per-warp partials, H1p <= 128
    __shared__ double s_cost[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const bool train = a.dZ1 != nullptr;
    // 8 lanes per row (float4 columns l8, l8+8, l8+16, l8+24), 4 rows per warp instruction, R such groups in flight:
    // every load / store instruction moves full 128-byte segments of 4 rows and a row sum needs 3 shuffles, not 5
    const int l8 = lane & 7, sub = lane >> 3;
    float4 w2v[4], dw2v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = (l8 + 8 * j) * 4;
        w2v[j].x = c + 0 < a.H1 ? __ldg(a.w2 + c + 0) : 0.0f; w2v[j].y = c + 1 < a.H1 ? __ldg(a.w2 + c + 1) : 0.0f;
        w2v[j].z = c + 2 < a.H1 ? __ldg(a.w2 + c + 2) : 0.0f; w2v[j].w = c + 3 < a.H1 ? __ldg(a.w2 + c + 3) : 0.0f;
        dw2v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double cost = 0.0;
    const float invB = 1.0f / (float)a.B;
    constexpr int R = 2;
    for (long b0 = (long)(blockIdx.x * (blockDim.x >> 5) + wib) * (4 * R); b0 < a.B; b0 += (long)nwarps * (4 * R)) {
        float4 h[R][4]; float z[R];
#pragma unroll
        for (int i = 0; i < R; i++) {
            const long b = b0 + 4 * i + sub;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = (l8 + 8 * j) * 4;
                h[i][j] = (c < a.H1p && b < a.B) ? ldg4(a.H1d + b * a.ldh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < R; i++) {
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                t = fmaf(h[i][j].x, w2v[j].x, t); t = fmaf(h[i][j].y, w2v[j].y, t);
                t = fmaf(h[i][j].z, w2v[j].z, t); t = fmaf(h[i][j].w, w2v[j].w, t);
            }
            z[i] = t;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
            for (int i = 0; i < R; i++) z[i] += __shfl_xor_sync(0xffffffffu, z[i], o);
        }
#pragma unroll
        for (int i = 0; i < R; i++) {
            const long b = b0 + 4 * i + sub;
            if (b >= a.B) continue;
            const float p = sigmoid32(z[i]);
            if (l8 == 0) { a.p[b] = p; if (a.logit) a.logit[b] = z[i]; }
            if (!train) continue;
            const float y = (b < a.nvalid) ? __ldg(a.y + b) : 0.0f;
            if (l8 == 0) cost += (double)(logf(p) * y + logf(1.0f - p) * (1.0f - y));
            const float dz2 = (p - y) * invB;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = (l8 + 8 * j) * 4;
                if (c < a.H1p) {
                    float4 d;       // columns >= H1 have w2 = 0: the padding of dZ1 stays zero
                    d.x = dz2 * w2v[j].x * dsigmoid_drop(h[i][j].x, a.drop_p); d.y = dz2 * w2v[j].y * dsigmoid_drop(h[i][j].y, a.drop_p);
                    d.z = dz2 * w2v[j].z * dsigmoid_drop(h[i][j].z, a.drop_p); d.w = dz2 * w2v[j].w * dsigmoid_drop(h[i][j].w, a.drop_p);
                    *reinterpret_cast<float4*>(a.dZ1 + b * a.lddz + c) = d;
                }
                dw2v[j].x = fmaf(h[i][j].x, dz2, dw2v[j].x); dw2v[j].y = fmaf(h[i][j].y, dz2, dw2v[j].y);
                dw2v[j].z = fmaf(h[i][j].z, dz2, dw2v[j].z); dw2v[j].w = fmaf(h[i][j].w, dz2, dw2v[j].w);
            }
        }
    }
    if (!train) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {       // the warp's four row groups → one partial per column
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            dw2v[j].x += __shfl_xor_sync(0xffffffffu, dw2v[j].x, o); dw2v[j].y += __shfl_xor_sync(0xffffffffu, dw2v[j].y, o);
            dw2v[j].z += __shfl_xor_sync(0xffffffffu, dw2v[j].z, o); dw2v[j].w += __shfl_xor_sync(0xffffffffu, dw2v[j].w, o);
        }
        if (sub == 0) *reinterpret_cast<float4*>(&s_dw2[wib][(l8 + 8 * j) * 4]) = dw2v[j];
    }
    cost += __shfl_xor_sync(0xffffffffu, cost, 8); cost += __shfl_xor_sync(0xffffffffu, cost, 16);     // the l8 == 0 lanes of the four groups
    if (lane == 0) s_cost[wib] = cost;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < a.H1; k += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < nw; w++) s += s_dw2[w][k];
        atomicAdd(a.dW2 + k, s);
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += s_cost[w];
        atomicAdd(a.cost_sum, s);
    }
}

// -------------------------------------------------------------------------------------------------
// gorgonia AdamSolver.Step over the learnables (model.go:88,192): g += l2 w ; g *= 1/batch ;
// m,v update ; w -= lr (m/c1)/(sqrt(v/c2)+eps) ; g = 0.  c1,c2 = 1-β^t computed on the host.
// One launch covers all four tensors (logical [rows, cols] inside padded storage with stride ld).
// -------------------------------------------------------------------------------------------------
struct AdamTensor { float* w; float* g; float* m; float* v; int rows, cols; long ld; };
// TF32 hi/lo operand copy of a weight tensor, written by the optimiser itself (the tcgen05 GEMMs of the next step read
// them; no separate split launch): element (r, c) goes to [c*old + r] (transpose) or [r*old + c], rows [r0, r0+nr) only
struct SplitOut { float* hi; float* lo; long old; int transpose, r0, nr; };
struct AdamArgs {
    AdamTensor t[4]; int nt;
    SplitOut sp[4][2]; int nsp[4];
    float lr, l2, inv_batch, b1, b2, eps, c1, c2;
    float gscale;       // 1/world after the all-reduce(sum) of the ranks' local-mean gradients, else 1
};

// element j of tensor TI (a compile-time index: the kernel parameters are read from the constant bank, not copied)
template <int TI>
__device__ __forceinline__ void adam_elem(const AdamArgs& a, int j) {
    const AdamTensor& t = a.t[TI];
    const int r = j / t.cols, c = j - r * t.cols;
    const long off = (long)r * t.ld + c;
    const float w = t.w[off];
    float g = t.g[off] * a.gscale;
    if (a.l2 != 0.0f) g = g + a.l2 * w;
    g = g * a.inv_batch;
    const float m = a.b1 * t.m[off] + (1.0f - a.b1) * g;
    const float v = a.b2 * t.v[off] + (1.0f - a.b2) * (g * g);
    t.m[off] = m; t.v[off] = v;
    const float wn = w - a.lr * (m / a.c1) / (sqrtf(v / a.c2) + a.eps);
    t.w[off] = wn;
    t.g[off] = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k >= a.nsp[TI]) break;
        const SplitOut& so = a.sp[TI][k];
        const int rr = r - so.r0;
        if (rr < 0 || rr >= so.nr) continue;
        const float hi = __uint_as_float(__float_as_uint(wn) & 0xFFFFE000u);
        const long o = so.transpose ? (long)c * so.old + rr : (long)rr * so.old + c;
        so.hi[o] = hi; so.lo[o] = wn - hi;
    }
}

__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ AdamArgs a) {
    // one flat index space over the (up to four) tensors: a thread's elements are independent, so their loads are all in
    // flight together — the tensors one after the other cost one dependent DRAM round trip each for 63 k parameters
    const int p1 = a.t[0].rows * a.t[0].cols, p2 = p1 + a.t[1].rows * a.t[1].cols, p3 = p2 + a.t[2].rows * a.t[2].cols;
    const int p4 = a.nt > 3 ? p3 + a.t[3].rows * a.t[3].cols : p3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p4; i += gridDim.x * blockDim.x) {
        if (i < p1) adam_elem<0>(a, i);
        else if (i < p2) adam_elem<1>(a, i - p1);
        else if (i < p3) adam_elem<2>(a, i - p2);
        else adam_elem<3>(a, i - p3);
    }
}

}  // namespace ctr
