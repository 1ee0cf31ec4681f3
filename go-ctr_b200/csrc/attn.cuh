// attn.cuh — embedding-row gather fused with DIN's attention ActivationUnit / YouTube mean-pool
// (forward), and its backward fused with the gradient scatter-add + SGD row update.
//
// Replaces, per batch: recommend.GetSampleVector's lookups + concat (rcmd.go:462-536),
// model.CosineSimilarity (activation.go:57-83) / EucDistance (:23-50), the attention weight and
// sigmoid gate (din.go:231-276), G.Mean pooling (din.go:298, dnn.go:164-167) and G.Concat
// (din.go:301, dnn.go:170).  HBM-bound: (S+1) rows of D floats per sample, one pass forward, one
// pass backward (+ the L2-side red.add that writes each touched row back once).
//
// Mapping: one warp per sample.  Vector kernels: a row of D=4*LPR floats is covered by LPR lanes
// with one 128-bit load each, so a warp issues 32/LPR rows per load instruction; the dot products
// reduce inside the LPR-lane group with __shfl_xor.  Generic kernels: any D<=256, scalar loads.
#pragma once
#include "common.cuh"

namespace ctr {

enum { MODEL_YOUTUBE = 0, MODEL_DIN_COS = 1, MODEL_DIN_EUC = 2 };

// prefetched history indices of one sample, two per lane (S <= 64), broadcast by shuffle
struct HistIdx {
    int i0, i1;
    __device__ __forceinline__ void load(const RowSrc& r, const Dims& d, int b, int lane) {
        i0 = i1 = -1;
        if (!r.dense && b < r.nvalid) {
            const int* h = r.hist + (long)b * d.S;
            if (lane < d.S) i0 = __ldg(h + lane);
            if (lane + 32 < d.S) i1 = __ldg(h + lane + 32);
        }
    }
    // all 32 lanes must call; s may differ per lane
    __device__ __forceinline__ int get(int s) const {
        int a0 = __shfl_sync(0xffffffffu, i0, s & 31);
        int a1 = __shfl_sync(0xffffffffu, i1, s & 31);
        return s < 32 ? a0 : a1;
    }
};

__device__ __forceinline__ const float* ub_ptr(const RowSrc& r, const Dims& d, const HistIdx& hi,
                                               bool use_hi, int b, int s) {
    // every lane of the warp reaches the shuffle inside hi.get()
    int sidx = use_hi ? hi.get(s < d.S ? s : 0) : -1;
    if (s >= d.S || b >= r.nvalid) return nullptr;
    if (r.dense) return r.X + (long)b * r.ldx + r.ub0 + (long)s * d.D;
    int idx = use_hi ? sidx : __ldg(r.hist + (long)b * d.S + s);
    return idx >= 0 ? r.emb + (long)idx * r.lde : nullptr;
}

// -------------------------------------------------------------------------------------------------
// forward, vector path.  Writes the MLP input row X0[b] = [uProfile | pooled | item | ctx | 0-pad].
// -------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256)
k_attn_fwd_vec(RowSrc r, Dims d, int model, const float* __restrict__ att,
               float* __restrict__ X0, long ldx0, int Kp, int B) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int lir = lane % LPR, sub = lane / LPR;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    float* row = smem + (long)wib * Kp;
    const float invS = 1.0f / (float)d.S;
    const bool use_hi = (!r.dense) && d.S <= 64;

    for (int b = blockIdx.x * (blockDim.x >> 5) + wib; b < B; b += nwarps) {
        HistIdx hi; hi.load(r, d, b, lane);
        const float* ip = src_it(r, d, b);
        const float4 v = ip ? ldg4(ip + lir * 4) : zero4();
        const float ny = sqrtf(group_sum<LPR>(dot4(v, v)));
        float4 acc = zero4();
        for (int s0 = 0; s0 < d.S; s0 += 2 * RPW) {
            // two independent row loads in flight per lane
            const int sA = s0 + sub, sB = s0 + RPW + sub;
            const float* pA = ub_ptr(r, d, hi, use_hi, b, sA);
            const float* pB = ub_ptr(r, d, hi, use_hi, b, sB);
            const float4 uA = pA ? ldg4_stream(pA + lir * 4) : zero4();
            const float4 uB = pB ? ldg4_stream(pB + lir * 4) : zero4();
            float aA = 1.0f, aB = 1.0f;
            if (model == MODEL_DIN_COS) {
                float dA = group_sum<LPR>(dot4(uA, v)), nA = group_sum<LPR>(dot4(uA, uA));
                float dB = group_sum<LPR>(dot4(uB, v)), nB = group_sum<LPR>(dot4(uB, uB));
                float wA = (dA / (sqrtf(nA) * ny + 1e-8f) + 1.0f) * 0.5f;
                float wB = (dB / (sqrtf(nB) * ny + 1e-8f) + 1.0f) * 0.5f;
                aA = sigmoid32(wA * (sA < d.S ? __ldg(att + sA) : 0.0f));
                aB = sigmoid32(wB * (sB < d.S ? __ldg(att + sB) : 0.0f));
            } else if (model == MODEL_DIN_EUC) {
                float4 eA = make_float4(uA.x - v.x, uA.y - v.y, uA.z - v.z, uA.w - v.w);
                float4 eB = make_float4(uB.x - v.x, uB.y - v.y, uB.z - v.z, uB.w - v.w);
                float wA = 1.0f - sqrtf(group_sum<LPR>(dot4(eA, eA)));
                float wB = 1.0f - sqrtf(group_sum<LPR>(dot4(eB, eB)));
                aA = sigmoid32(wA * (sA < d.S ? __ldg(att + sA) : 0.0f));
                aB = sigmoid32(wB * (sB < d.S ? __ldg(att + sB) : 0.0f));
            }
            acc = fma4(aA, uA, acc);     // slots beyond S and missing rows carry u == 0
            acc = fma4(aB, uB, acc);
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        // assemble the concat row in shared memory, then one coalesced 128-bit store pass
        const float* pu = src_up(r, b);
        for (int j = lane; j < d.uP; j += 32) row[j] = pu ? __ldg(pu + j) : 0.0f;
        if (sub == 0) {
            float* q = row + d.uP + lir * 4;
            q[0] = acc.x * invS; q[1] = acc.y * invS; q[2] = acc.z * invS; q[3] = acc.w * invS;
            float* q2 = row + d.uP + d.D + lir * 4;
            q2[0] = v.x; q2[1] = v.y; q2[2] = v.z; q2[3] = v.w;
        }
        const float* pc = src_cx(r, b);
        for (int j = lane; j < d.cF; j += 32) row[d.uP + 2 * d.D + j] = pc ? __ldg(pc + j) : 0.0f;
        for (int j = d.in + lane; j < Kp; j += 32) row[j] = 0.0f;
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(X0 + (long)b * ldx0);
        const float4* src4 = reinterpret_cast<const float4*>(row);
        for (int j = lane; j < Kp / 4; j += 32) dst[j] = src4[j];
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// forward, generic path (any D <= 256, unaligned sources — the dense-X compatibility route and odd
// dims such as the reference test's D=7, model_test.go:24-28).
// -------------------------------------------------------------------------------------------------
constexpr int kGenAcc = 8;   // D <= 32*kGenAcc

__global__ void __launch_bounds__(256)
k_attn_fwd_gen(RowSrc r, Dims d, int model, const float* __restrict__ att,
               float* __restrict__ X0, long ldx0, int Kp, int B) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float invS = 1.0f / (float)d.S;
    HistIdx hi; hi.i0 = hi.i1 = -1;
    for (int b = blockIdx.x * (blockDim.x >> 5) + wib; b < B; b += nwarps) {
        const float* ip = src_it(r, d, b);
        float vv[kGenAcc], acc[kGenAcc];
        float ny2 = 0.0f;
#pragma unroll
        for (int j = 0; j < kGenAcc; j++) {
            int k = lane + 32 * j;
            vv[j] = (ip && k < d.D) ? __ldg(ip + k) : 0.0f;
            ny2 += vv[j] * vv[j]; acc[j] = 0.0f;
        }
        const float ny = sqrtf(warp_sum(ny2));
        for (int s = 0; s < d.S; s++) {
            const float* up = ub_ptr(r, d, hi, false, b, s);
            float uu[kGenAcc]; float dot = 0.0f, nx2 = 0.0f, d2 = 0.0f;
#pragma unroll
            for (int j = 0; j < kGenAcc; j++) {
                int k = lane + 32 * j;
                uu[j] = (up && k < d.D) ? __ldg(up + k) : 0.0f;
                dot += uu[j] * vv[j]; nx2 += uu[j] * uu[j];
                float e = uu[j] - vv[j]; d2 += e * e;
            }
            float a = 1.0f;
            if (model == MODEL_DIN_COS) {
                dot = warp_sum(dot); nx2 = warp_sum(nx2);
                float w = (dot / (sqrtf(nx2) * ny + 1e-8f) + 1.0f) * 0.5f;
                a = sigmoid32(w * __ldg(att + s));
            } else if (model == MODEL_DIN_EUC) {
                d2 = warp_sum(d2);
                a = sigmoid32((1.0f - sqrtf(d2)) * __ldg(att + s));
            }
#pragma unroll
            for (int j = 0; j < kGenAcc; j++) acc[j] = fmaf(a, uu[j], acc[j]);
        }
        float* dst = X0 + (long)b * ldx0;
        const float* pu = src_up(r, b);
        for (int j = lane; j < d.uP; j += 32) dst[j] = pu ? __ldg(pu + j) : 0.0f;
#pragma unroll
        for (int j = 0; j < kGenAcc; j++) {
            int k = lane + 32 * j;
            if (k < d.D) { dst[d.uP + k] = acc[j] * invS; dst[d.uP + d.D + k] = vv[j]; }
        }
        const float* pc = src_cx(r, b);
        for (int j = lane; j < d.cF; j += 32) dst[d.uP + 2 * d.D + j] = pc ? __ldg(pc + j) : 0.0f;
        for (int j = d.in + lane; j < Kp; j += 32) dst[j] = 0.0f;
    }
}

// -------------------------------------------------------------------------------------------------
// backward.  dX[b] = [g (d cost/d pooled) | gi (d cost/d item through the MLP input)].
// Emits d cost/d att0 (sum over batch), and for every gathered row either
//   * sgd != 0: table_row += -lr * grad with red.global.add.v4.f32 (fused scatter-add + SGD), or
//   * dUb/dIt buffers (deterministic update path, debug hook, multi-GPU return leg).
// Analytic reverse of din.go:231-298 (see DESIGN.md §kernels for the derivation):
//   da_s = g·u_s / S ; dz_s = da_s a_s (1-a_s) ; datt_s += dz_s w_s ; dw_s = dz_s att_s
//   cosine:  c = dw_s/2 ; du_s = a_s g/S + c (v/den - cos·|v| u_s/(|u_s| den))
//                         dv  += c (u_s/den - cos·|u_s| v/(|v| den))
//   euclid:  du_s = a_s g/S - dw_s (u_s-v)/dist ; dv += dw_s (u_s-v)/dist
//   mean  :  du_s = g/S
// -------------------------------------------------------------------------------------------------
struct BwdOut {
    float* datt;        // [S] accumulated with atomics (zeroed by the optimiser step)
    float* dUb;         // [B,S,D] or null
    float* dIt;         // [B,D]   or null
    int    sgd;         // fused scatter-add + SGD into r.emb
    float  neg_lr;      // -table_lr
};

template <int LPR>
__global__ void __launch_bounds__(256)
k_attn_bwd_vec(RowSrc r, Dims d, int model, const float* __restrict__ att,
               const float* __restrict__ dX, long lddx, BwdOut o, int B) {
    extern __shared__ __align__(16) float smem[];     // datt partials [S]
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int lir = lane % LPR, sub = lane / LPR;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float invS = 1.0f / (float)d.S;
    const bool use_hi = (!r.dense) && d.S <= 64;
    for (int j = threadIdx.x; j < d.S; j += blockDim.x) smem[j] = 0.0f;
    __syncthreads();

    for (int b = blockIdx.x * (blockDim.x >> 5) + wib; b < B; b += nwarps) {
        HistIdx hi; hi.load(r, d, b, lane);
        const float4 g = ldg4(dX + (long)b * lddx + lir * 4);
        const float4 gi = ldg4(dX + (long)b * lddx + d.D + lir * 4);
        const float* ip = src_it(r, d, b);
        // the table is written by this kernel (sgd mode): coherent loads, no .nc
        const float4 v = ip ? *reinterpret_cast<const float4*>(ip + lir * 4) : zero4();
        const float ny = sqrtf(group_sum<LPR>(dot4(v, v)));
        float4 dv = sub == 0 ? gi : zero4();
        for (int s0 = 0; s0 < d.S; s0 += RPW) {
            const int s = s0 + sub;
            const float* up = ub_ptr(r, d, hi, use_hi, b, s);
            const float4 u = up ? *reinterpret_cast<const float4*>(up + lir * 4) : zero4();
            float4 du;
            if (model == MODEL_YOUTUBE) {
                du = make_float4(g.x * invS, g.y * invS, g.z * invS, g.w * invS);
            } else {
                const float att_s = s < d.S ? __ldg(att + s) : 0.0f;
                const float gu = group_sum<LPR>(dot4(g, u));
                if (model == MODEL_DIN_COS) {
                    const float dot = group_sum<LPR>(dot4(u, v));
                    const float nx = sqrtf(group_sum<LPR>(dot4(u, u)));
                    const float den = nx * ny + 1e-8f;
                    const float cs = dot / den;
                    const float w = (cs + 1.0f) * 0.5f;
                    const float a = sigmoid32(w * att_s);
                    const float dz = gu * invS * a * (1.0f - a);
                    if (lir == 0 && s < d.S) atomicAdd(&smem[s], dz * w);
                    const float cc = 0.5f * dz * att_s;
                    const float iden = 1.0f / den;
                    const float ku = nx > 0.0f ? cs * ny / (nx * den) : 0.0f;
                    const float kv = ny > 0.0f ? cs * nx / (ny * den) : 0.0f;
                    const float ag = a * invS;
                    du = make_float4(ag * g.x + cc * (v.x * iden - ku * u.x), ag * g.y + cc * (v.y * iden - ku * u.y),
                                     ag * g.z + cc * (v.z * iden - ku * u.z), ag * g.w + cc * (v.w * iden - ku * u.w));
                    dv.x += cc * (u.x * iden - kv * v.x); dv.y += cc * (u.y * iden - kv * v.y);
                    dv.z += cc * (u.z * iden - kv * v.z); dv.w += cc * (u.w * iden - kv * v.w);
                } else {
                    const float4 e = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
                    const float dist = sqrtf(group_sum<LPR>(dot4(e, e)));
                    const float w = 1.0f - dist;
                    const float a = sigmoid32(w * att_s);
                    const float dz = gu * invS * a * (1.0f - a);
                    if (lir == 0 && s < d.S) atomicAdd(&smem[s], dz * w);
                    const float dw = dz * att_s;
                    const float k = dist > 0.0f ? dw / dist : 0.0f;
                    const float ag = a * invS;
                    du = make_float4(ag * g.x - k * e.x, ag * g.y - k * e.y, ag * g.z - k * e.z, ag * g.w - k * e.w);
                    if (s < d.S) { dv.x += k * e.x; dv.y += k * e.y; dv.z += k * e.z; dv.w += k * e.w; }
                }
            }
            if (s < d.S) {
                if (o.dUb) *reinterpret_cast<float4*>(o.dUb + ((long)b * d.S + s) * d.D + lir * 4) = du;
                if (o.sgd && up)
                    red_add4(const_cast<float*>(up) + lir * 4,
                             make_float4(o.neg_lr * du.x, o.neg_lr * du.y, o.neg_lr * du.z, o.neg_lr * du.w));
            }
        }
#pragma unroll
        for (int of = LPR; of < 32; of <<= 1) {
            dv.x += __shfl_xor_sync(0xffffffffu, dv.x, of);
            dv.y += __shfl_xor_sync(0xffffffffu, dv.y, of);
            dv.z += __shfl_xor_sync(0xffffffffu, dv.z, of);
            dv.w += __shfl_xor_sync(0xffffffffu, dv.w, of);
        }
        if (sub == 0) {
            if (o.dIt) *reinterpret_cast<float4*>(o.dIt + (long)b * d.D + lir * 4) = dv;
            if (o.sgd && ip)
                red_add4(const_cast<float*>(ip) + lir * 4,
                         make_float4(o.neg_lr * dv.x, o.neg_lr * dv.y, o.neg_lr * dv.z, o.neg_lr * dv.w));
        }
    }
    __syncthreads();
    if (model != MODEL_YOUTUBE && o.datt)
        for (int j = threadIdx.x; j < d.S; j += blockDim.x)
            if (smem[j] != 0.0f) atomicAdd(o.datt + j, smem[j]);
}

__global__ void __launch_bounds__(256)
k_attn_bwd_gen(RowSrc r, Dims d, int model, const float* __restrict__ att,
               const float* __restrict__ dX, long lddx, BwdOut o, int B) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float invS = 1.0f / (float)d.S;
    HistIdx hi; hi.i0 = hi.i1 = -1;
    for (int j = threadIdx.x; j < d.S; j += blockDim.x) smem[j] = 0.0f;
    __syncthreads();
    for (int b = blockIdx.x * (blockDim.x >> 5) + wib; b < B; b += nwarps) {
        const float* ip = src_it(r, d, b);
        float vv[kGenAcc], gg[kGenAcc], dv[kGenAcc];
        float ny2 = 0.0f;
#pragma unroll
        for (int j = 0; j < kGenAcc; j++) {
            int k = lane + 32 * j;
            vv[j] = (ip && k < d.D) ? ip[k] : 0.0f;
            gg[j] = k < d.D ? __ldg(dX + (long)b * lddx + k) : 0.0f;
            dv[j] = k < d.D ? __ldg(dX + (long)b * lddx + d.D + k) : 0.0f;
            ny2 += vv[j] * vv[j];
        }
        const float ny = sqrtf(warp_sum(ny2));
        for (int s = 0; s < d.S; s++) {
            const float* up = ub_ptr(r, d, hi, false, b, s);
            float uu[kGenAcc], du[kGenAcc];
            float dot = 0.0f, nx2 = 0.0f, d2 = 0.0f, gu = 0.0f;
#pragma unroll
            for (int j = 0; j < kGenAcc; j++) {
                int k = lane + 32 * j;
                uu[j] = (up && k < d.D) ? up[k] : 0.0f;
                dot += uu[j] * vv[j]; nx2 += uu[j] * uu[j]; gu += gg[j] * uu[j];
                float e = uu[j] - vv[j]; d2 += e * e;
            }
            if (model == MODEL_YOUTUBE) {
#pragma unroll
                for (int j = 0; j < kGenAcc; j++) du[j] = gg[j] * invS;
            } else {
                const float att_s = __ldg(att + s);
                gu = warp_sum(gu);
                if (model == MODEL_DIN_COS) {
                    dot = warp_sum(dot);
                    const float nx = sqrtf(warp_sum(nx2));
                    const float den = nx * ny + 1e-8f, cs = dot / den, w = (cs + 1.0f) * 0.5f;
                    const float a = sigmoid32(w * att_s);
                    const float dz = gu * invS * a * (1.0f - a);
                    if (lane == 0) atomicAdd(&smem[s], dz * w);
                    const float cc = 0.5f * dz * att_s, iden = 1.0f / den;
                    const float ku = nx > 0.0f ? cs * ny / (nx * den) : 0.0f;
                    const float kv = ny > 0.0f ? cs * nx / (ny * den) : 0.0f;
#pragma unroll
                    for (int j = 0; j < kGenAcc; j++) {
                        du[j] = a * invS * gg[j] + cc * (vv[j] * iden - ku * uu[j]);
                        dv[j] += cc * (uu[j] * iden - kv * vv[j]);
                    }
                } else {
                    const float dist = sqrtf(warp_sum(d2));
                    const float w = 1.0f - dist, a = sigmoid32(w * att_s);
                    const float dz = gu * invS * a * (1.0f - a);
                    if (lane == 0) atomicAdd(&smem[s], dz * w);
                    const float dw = dz * att_s, k2 = dist > 0.0f ? dw / dist : 0.0f;
#pragma unroll
                    for (int j = 0; j < kGenAcc; j++) {
                        float e = uu[j] - vv[j];
                        du[j] = a * invS * gg[j] - k2 * e;
                        dv[j] += k2 * e;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kGenAcc; j++) {
                int k = lane + 32 * j;
                if (k < d.D) {
                    if (o.dUb) o.dUb[((long)b * d.S + s) * d.D + k] = du[j];
                    if (o.sgd && up) atomicAdd(const_cast<float*>(up) + k, o.neg_lr * du[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kGenAcc; j++) {
            int k = lane + 32 * j;
            if (k < d.D) {
                if (o.dIt) o.dIt[(long)b * d.D + k] = dv[j];
                if (o.sgd && ip) atomicAdd(const_cast<float*>(ip) + k, o.neg_lr * dv[j]);
            }
        }
    }
    __syncthreads();
    if (model != MODEL_YOUTUBE && o.datt)
        for (int j = threadIdx.x; j < d.S; j += blockDim.x)
            if (smem[j] != 0.0f) atomicAdd(o.datt + j, smem[j]);
}

// -------------------------------------------------------------------------------------------------
// recommend.GetSampleVector materialised (rcmd.go:462-536): X[b] = [user | S history rows | item
// emb | item feat].  Bit-exact copies; one warp per sample.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gather_rows(RowSrc r, Dims d, float* __restrict__ X, long ldx, int B) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    HistIdx hi; hi.i0 = hi.i1 = -1;
    for (int b = blockIdx.x * (blockDim.x >> 5) + wib; b < B; b += nwarps) {
        float* x = X + (long)b * ldx;
        const float* pu = src_up(r, b);
        for (int j = lane; j < d.uP; j += 32) x[j] = pu ? __ldg(pu + j) : 0.0f;
        for (int s = 0; s < d.S; s++) {
            const float* up = ub_ptr(r, d, hi, false, b, s);
            for (int j = lane; j < d.D; j += 32) x[d.uP + (long)s * d.D + j] = up ? __ldg(up + j) : 0.0f;
        }
        const float* ip = src_it(r, d, b);
        float* xi = x + d.uP + (long)d.S * d.D;
        for (int j = lane; j < d.D; j += 32) xi[j] = ip ? __ldg(ip + j) : 0.0f;
        const float* pc = src_cx(r, b);
        for (int j = lane; j < d.cF; j += 32) xi[d.D + j] = pc ? __ldg(pc + j) : 0.0f;
    }
}

// -------------------------------------------------------------------------------------------------
// Deterministic row update (CTR_TABLE_SGD_DETERMINISTIC): keys[p] = table row of gradient slot
// p = b*(S+1)+slot (slot S = target item), sorted stably; one warp walks each equal-key segment in
// ascending p (== (b, slot) order) accumulating in double, then applies row -= lr * sum once.
// -------------------------------------------------------------------------------------------------
__global__ void k_scatter_keys(const int* __restrict__ hist, const int* __restrict__ item_row,
                               int S, int B, unsigned* __restrict__ keys, unsigned* __restrict__ pos) {
    long n = (long)B * (S + 1);
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
        int row = sl < S ? hist[(long)b * S + sl] : item_row[b];
        keys[p] = row >= 0 ? (unsigned)row : 0xFFFFFFFFu;
        pos[p] = (unsigned)p;
    }
}

__global__ void __launch_bounds__(256)
k_segment_sgd(const unsigned* __restrict__ keys, const unsigned* __restrict__ pos, long n,
              const float* __restrict__ dUb, const float* __restrict__ dIt, int S, int D,
              float* __restrict__ emb, long lde, float lr) {
    const int lane = threadIdx.x & 31;
    long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long i = w; i < n; i += nw) {
        unsigned key = keys[i];
        if (key == 0xFFFFFFFFu) continue;
        if (i > 0 && keys[i - 1] == key) continue;        // not a segment head
        for (int k0 = 0; k0 < D; k0 += 32) {
            int k = k0 + lane;
            double acc = 0.0;
            for (long j = i; j < n && keys[j] == key; j++) {
                unsigned p = pos[j];
                int b = (int)(p / (unsigned)(S + 1)), sl = (int)(p % (unsigned)(S + 1));
                if (k < D) acc += (double)(sl < S ? dUb[((long)b * S + sl) * D + k] : dIt[(long)b * D + k]);
            }
            if (k < D) {
                float* e = emb + (long)key * lde + k;
                *e = (float)((double)*e - (double)lr * acc);
            }
        }
    }
}

// synthetic table rows generated in place (ctr_table_fill): global row = local*world + rank
__global__ void k_table_fill(float* __restrict__ t, long ld, long local_rows, int width, int world, int rank,
                             uint32_t seed, uint32_t stream, int dist, float scale) {
    long n = local_rows * (long)width;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long lr = i / width; int c = (int)(i % width);
        uint64_t ctr = (uint64_t)(lr * world + rank) * (uint64_t)width + c;
        uint64_t z = mix64(seed, stream, ctr);
        float v;
        if (dist == 0) v = (float)(z >> 40) * (1.0f / 16777216.0f);
        else {
            float u1 = ((float)(z >> 40) + 1.0f) * (1.0f / 16777217.0f);
            float u2 = (float)((z >> 8) & 0xFFFFFFull) * (1.0f / 16777216.0f);
            v = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
        }
        t[lr * ld + c] = v * scale;
    }
}

}  // namespace ctr
