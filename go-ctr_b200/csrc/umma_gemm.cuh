// umma_gemm.cuh — the dense MLP GEMMs on Blackwell's 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// hand-written for sm_100a.  C[M,N] = epilogue(A[M,K] · B[N,K]ᵀ), fp32 in / fp32 out.
//
// Precision: the reference's MLP is float32 (gonum Sgemm under gorgonia, din.go:307-315) and the
// parity bar is 1e-4 relative on scores, which single-pass TF32 (10-bit mantissa) cannot hold.  So
// each product is error-compensated "3xTF32": x = hi + lo with hi = x & 0xFFFFE000 (exactly
// representable in TF32) and lo = x - hi (exact in fp32), and
//     A·B ≈ A_hi·B_hi + A_hi·B_lo + A_lo·B_hi            (fp32 accumulation in TMEM).
// B (weights, tiny) is pre-split once per step by k_split_weights; A (activations, streamed from
// HBM exactly once) is split inside the kernel by two converter warps, tile by tile in shared
// memory — the operation is elementwise, so it is oblivious to the 128-byte swizzle TMA applied.
//
// One persistent CTA per SM, 14 warps, warp-specialised:
//   warp 0   TMA producer: cp.async.bulk.tensor 2D, SWIZZLE_128B, {32 fp32 x 128 rows} A boxes and
//            {32 x BLOCK_N} B_hi/B_lo boxes into a STAGES-deep smem ring (mbarrier complete_tx)
//   warp 1   MMA issuer: one elected thread, tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BLOCK_N,
//            K=8 per instruction; 4 k-steps x 3 products per 32-wide k-block; tcgen05.commit frees
//            the smem stage and, after the last k-block, publishes the TMEM accumulator
//   warps 2-5 converters (warp 2 also owns tcgen05.alloc/dealloc)
//   warps 6-13 epilogue: tcgen05.ld 32x32b.x16 (TMEM lane = tile row; warp w reads lane quadrant
//            w%4, the two warps of a quadrant take alternate 16-column chunks), next chunk's TMEM
//            load in flight while the current one is transformed; fused sigmoid+dropout / dsigmoid /
//            plain store, 64-byte row segments to global
// Two TMEM accumulators (double buffering) overlap tile i's epilogue with tile i+1's MMAs.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ctr {
namespace umma {

enum { UEPI_STORE = 0, UEPI_SIGMOID_DROP = 1, UEPI_DSIGMOID = 2 };

constexpr int kBlockM = 128;
// K per shared-memory stage is a template parameter of k_umma_gemm (KBK = 32 or 16 fp32: one 128- or 64-byte swizzle row)

struct Args {
    int M, N, Nz, K;            // K multiple of 32; columns [N, Nz) of C written as zeros
    int bn;                     // BLOCK_N (multiple of 16, <= 256) >= Nz
    float* C; long ldc;
    const float* H; long ldh;   // UEPI_DSIGMOID: stored post-dropout activation
    float drop_p; uint32_t seed, stream;
    int stages;
    int staged_epi;             // 1: the epilogue goes through per-warp shared-memory staging (coalesced 128-byte row segments)
    int tr;                     // 1: transposed accumulation (template TR): 256-row tiles, weight tile as the MMA's M operand; needs bn <= 128
    int rawhi;                  // 1: the MMA reads the raw fp32 tile as A_hi (the tensor core ignores the low 13 mantissa bits)
    int kbk;                    // K elements per shared-memory stage (32 or 16): selects the kernel instantiation and the maps' box
    int pf;                     // activation k-blocks requested into L2 ahead of the shared-memory ring (0 = off)
    unsigned long long* dbg;    // optional timeline of CTA 0 (globaltimer ns): [it*8 + event], tiles at [4096 + t*4 + e]
};

constexpr int kEpiWarps = 8;
constexpr int kEpiRowBytes = 144;                                  // 32 floats + 16 bytes of padding: conflict-free float4 access both ways
constexpr int kEpiStageBytes = 32 * kEpiRowBytes;                  // one warp's 32 x 32 staging tile

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define UDBG(idx) do { if (a.dbg && blockIdx.x == 0 && (idx) < 8192) a.dbg[(idx)] = gtimer(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no barrier): the later tma_load_2d of the same box is an L2 hit
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" :: "l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// asynchronous TMEM → register load of 16 consecutive columns of this thread's lane; the registers
// are valid only after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), LBO unused (=1); version 1 (sm_100); layout type 2.
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Same for SWIZZLE_64B: rows of 64 bytes (16 fp32 of K), 8-row groups 512 bytes apart; layout type 4.
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
template <int KBK> __device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return KBK == 32 ? desc_k_sw128(saddr) : desc_k_sw64(saddr); }
// instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// KBK: fp32 elements of K per shared-memory stage — 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B:
// half-size stages, twice as many of them in the same shared memory)
// TR (plain-store epilogue only: dX): the accumulator holds Cᵀ — the WEIGHT tile (bn <= 128 rows) is the MMA's M operand
// and 256 batch rows are its N operand.  A TF32 MMA covers K = 8 and occupies the tensor pipe >= ~65 ns however narrow it
// is, so one 128x256x8 MMA per term replaces two 128xbnx8 ones, and the epilogue stores coalesced rows without staging.
// Measured: dX 35.2 → 31.5 µs; the same variant with the sigmoid epilogue made fwd1 slower (33.4 → 43.1 µs) and was dropped.
template <int EPI, bool SPLIT3, int KBK, bool TR>
__global__ void __launch_bounds__(448, 1)
k_umma_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
            const __grid_constant__ CUtensorMap tmBlo, Args a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int bn = a.bn, stages = a.stages;
    static_assert(!TR || EPI == UEPI_STORE, "transposed accumulation is wired for the plain-store epilogue only");
    constexpr int kTileRows = TR ? 256 : kBlockM;                     // batch rows per tile
    constexpr uint32_t kABytes = (uint32_t)kTileRows * KBK * 4u;
    const uint32_t bBytes = (uint32_t)bn * (uint32_t)KBK * 4u;
    const uint32_t stBytes = kABytes * (SPLIT3 ? 2u : 1u) + bBytes * (SPLIT3 ? 2u : 1u);
    // stage s: [A | (Alo) | Bhi | (Blo)]
    auto sA = [&](int s) { return base + (uint32_t)s * stBytes; };
    auto sAlo = [&](int s) { return sA(s) + kABytes; };
    auto sBhi = [&](int s) { return sA(s) + kABytes * (SPLIT3 ? 2u : 1u); };
    auto sBlo = [&](int s) { return sBhi(s) + bBytes; };
    const uint32_t bars = base + (uint32_t)stages * stBytes;          // 8-byte barriers
    auto full = [&](int s) { return bars + 8u * s; };
    auto conv = [&](int s) { return bars + 8u * (stages + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * stages + s); };
    auto tfull = [&](int i) { return bars + 8u * (3 * stages + i); };
    auto tempty = [&](int i) { return bars + 8u * (3 * stages + 2 + i); };
    const uint32_t tmem_slot = bars + 8u * (3 * stages + 4);
    const uint32_t epi_base = (tmem_slot + 16u + 15u) & ~15u;          // kEpiWarps staging tiles (staged epilogue)
    const int acc_stride = TR ? 256 : (bn <= 128 ? 128 : 256);
    const uint32_t ncols = 2u * acc_stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(full(s), 1); mbar_init(conv(s), 128); mbar_init(empty(s), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 256); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int num_k = a.K / KBK;
    const int num_tiles = (a.M + kTileRows - 1) / kTileRows;

    if (warp == 0) {
        if (lane == 0) {                                            // ---------------- TMA producer
            // The ring holds only `stages` (2-3) 16 KB activation boxes per SM — far too few bytes in flight to cover the
            // loaded HBM latency.  The boxes of the next a.pf k-blocks are therefore requested into L2 ahead of the ring.
            const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            const uint32_t total = (uint32_t)my_tiles * (uint32_t)num_k;
            auto prefetch_a = [&](uint32_t j) {
                if (j < total) tma_prefetch_2d(&tmA, (int)(j % num_k) * KBK, ((int)blockIdx.x + (int)(j / num_k) * (int)gridDim.x) * kTileRows);
            };
            for (uint32_t j = stages; j < (uint32_t)(stages + a.pf); j++) prefetch_a(j);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_k; kb++, it++) {
                    const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                    if (a.pf) prefetch_a(it + stages + a.pf);
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full(s), kABytes + bBytes * (SPLIT3 ? 2u : 1u));
                    UDBG(it * 8 + 0);
                    tma_load_2d(sA(s), &tmA, full(s), kb * KBK, tile * kTileRows);
                    tma_load_2d(sBhi(s), &tmBhi, full(s), kb * KBK, 0);
                    if (SPLIT3) tma_load_2d(sBlo(s), &tmBlo, full(s), kb * KBK, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                            // ---------------- MMA issuer
            const uint32_t idesc = TR ? idesc_tf32(kBlockM, 256) : idesc_tf32(kBlockM, bn);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
                const int acc = tcount & 1; const uint32_t aph = (tcount >> 1) & 1u;
                mbar_wait(tempty(acc), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                for (int kb = 0; kb < num_k; kb++, it++) {
                    const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(SPLIT3 ? conv(s) : full(s), ph);
                    UDBG(it * 8 + 3);
                    tc_fence_after();
                    const uint64_t dAhi = desc_k<KBK>(sA(s)), dBhi = desc_k<KBK>(sBhi(s));
                    const uint64_t dAlo = SPLIT3 ? desc_k<KBK>(sAlo(s)) : 0, dBlo = SPLIT3 ? desc_k<KBK>(sBlo(s)) : 0;
#pragma unroll
                    for (int k = 0; k < KBK / 8; k++) {         // UMMA_K = 8 tf32 = 32 bytes = +2 in the address field
                        const uint64_t ko = (uint64_t)(k * 2);
                        if (TR) {           // rows of D = weight rows, columns of D = the tile's 256 batch rows
                            umma_tf32(d_tmem, dBhi + ko, dAhi + ko, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            if (SPLIT3) {
                                umma_tf32(d_tmem, dBlo + ko, dAhi + ko, idesc, 1u);
                                umma_tf32(d_tmem, dBhi + ko, dAlo + ko, idesc, 1u);
                            }
                        } else {
                            umma_tf32(d_tmem, dAhi + ko, dBhi + ko, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            if (SPLIT3) {
                                umma_tf32(d_tmem, dAhi + ko, dBlo + ko, idesc, 1u);
                                umma_tf32(d_tmem, dAlo + ko, dBhi + ko, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(empty(s));                          // smem stage reusable once these MMAs retire
                    UDBG(it * 8 + 4);
                }
                umma_commit(tfull(acc));                            // accumulator complete
            }
        }
    } else if (warp < 6) {
        if (SPLIT3) {                                               // ---------------- converters (warps 2..5, 128 threads)
            const int c = threadIdx.x - 64;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_k; kb++, it++) {
                    const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(full(s), ph);
                    if (c == 0) UDBG(it * 8 + 1);
                    const uint32_t pa = sA(s), pl = sAlo(s);
                    constexpr int NV = (int)(kABytes / 16u / 128u);     // float4 per thread: 128 threads x NV x 16 B = the activation tile
                    constexpr int NP = NV < 8 ? NV : 8;                  // handled NP at a time: all loads of a part first
#pragma unroll
                    for (int p0 = 0; p0 < NV; p0 += NP) {
                        float4 x[NP];
#pragma unroll
                        for (int i = 0; i < NP; i++)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                                         : "r"(pa + 16u * (c + 128 * (p0 + i))));
#pragma unroll
                        for (int i = 0; i < NP; i++) {
                            float4 h, l;
                            h.x = __uint_as_float(__float_as_uint(x[i].x) & 0xFFFFE000u); l.x = x[i].x - h.x;
                            h.y = __uint_as_float(__float_as_uint(x[i].y) & 0xFFFFE000u); l.y = x[i].y - h.y;
                            h.z = __uint_as_float(__float_as_uint(x[i].z) & 0xFFFFE000u); l.z = x[i].z - h.z;
                            h.w = __uint_as_float(__float_as_uint(x[i].w) & 0xFFFFE000u); l.w = x[i].w - h.w;
                            // kind::tf32 reads the upper 19 bits of each 32-bit operand word, i.e. the raw tile already IS A_hi
                            // (bit-identical results, tests/test_gpu_umma.py); a.rawhi = 0 writes the truncated copy anyway
                            if (!a.rawhi) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(pa + 16u * (c + 128 * (p0 + i))), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(pl + 16u * (c + 128 * (p0 + i))), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                        }
                    }
                    fence_proxy_async();                            // generic-proxy writes → visible to tcgen05.mma
                    mbar_arrive(conv(s));
                    if (c == 0) UDBG(it * 8 + 2);
                }
            }
        }
    } else {                                                        // ---------------- epilogue (warps 6..13)
        const int q = warp & 3;                                     // TMEM lane quadrant this warp may read
        const int half = (warp - 6) >> 2;                           // which alternate 16-column chunks
        const int row_in_tile = q * 32 + lane;
        const float inv_keep = a.drop_p > 0.0f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
        const float keep_thr = 1.0f - a.drop_p;
        const int nchunks = bn / 16;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
            const int acc = tcount & 1; const uint32_t aph = (tcount >> 1) & 1u;
            mbar_wait(tfull(acc), aph);
            if (threadIdx.x == 192) UDBG(4096 + tcount * 4 + 0);
            tc_fence_after();
            const long gm = (long)tile * kBlockM + row_in_tile;
            const uint32_t trow = tmem_base + (uint32_t)(acc * acc_stride) + ((uint32_t)(q * 32) << 16);
            const uint32_t ctr0 = (uint32_t)((uint64_t)gm * (uint64_t)a.N);
            if (TR) {
                // lane = output feature f (accumulator row), accumulator columns = the tile's 256 batch rows: the warp's store
                // of one column is 32 consecutive floats of one row of C — coalesced without staging
                const int f = q * 32 + lane;
                const long tile_row0 = (long)tile * kTileRows;
                if (q * 32 < a.Nz) {
                    for (int ci = half; ci < kTileRows / 16; ci += 2) {
                        uint32_t r[16];
                        tmem_ld16_async(trow + (uint32_t)(ci * 16), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const long grow = tile_row0 + ci * 16 + j;
                            float v = __uint_as_float(r[j]);
                            v = f < a.N ? v : 0.0f;
                            if (grow < a.M && f < a.Nz) a.C[grow * a.ldc + f] = v;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty(acc));
                continue;
            }
            if (a.staged_epi) {
                // Staged epilogue.  tcgen05.ld hands every thread one ROW of the tile, so storing straight from registers
                // writes 32 different rows per instruction (16 useful bytes of each 128-byte line).  Here a warp transposes
                // 32 rows x 32 columns through its private shared-memory tile: thread r writes its row as float4s (row pitch
                // 144 B: the 8 lanes of a quarter-warp cover all 32 banks), then every instruction moves 4 rows x 128
                // contiguous bytes between shared and global memory — full lines, for the C store and for the H load.
                const uint32_t stg = epi_base + (uint32_t)(warp - 6) * kEpiStageBytes;
                const uint32_t my_row = stg + (uint32_t)lane * kEpiRowBytes;
                const int sub = lane >> 3, c4 = (lane & 7) * 4;                     // coalesced phase: row 4*j + sub, columns c4..c4+3
                const long tile_row0 = (long)tile * kBlockM + q * 32;
                const int nch32 = (a.Nz + 31) / 32;
                const float kp = a.drop_p > 0.0f ? 1.0f - a.drop_p : 1.0f;
                // dsigmoid: the H tile of a chunk is fetched (coalesced) one chunk AHEAD — issued as soon as the current chunk's H has
                // moved from the registers to the staging tile — the epilogue of a K = 96 GEMM is otherwise a chain of dependent round trips
                float4 hn[8];
                auto fetch_h = [&](int c0n) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const long gr = tile_row0 + 4 * j + sub;
                        hn[j] = (gr < a.M && c0n + c4 < a.Nz) ? ldg4(a.H + gr * a.ldh + c0n + c4) : zero4();
                    }
                };
                if (EPI == UEPI_DSIGMOID && half < nch32) fetch_h(half * 32);
                for (int cj = half; cj < nch32; cj += 2) {
                    const int c0 = cj * 32;
                    uint32_t ra[16], rb[16];
                    tmem_ld16_async(trow + (uint32_t)c0, ra);
                    if (c0 + 16 < bn) tmem_ld16_async(trow + (uint32_t)(c0 + 16), rb);
                    if (EPI == UEPI_DSIGMOID) {                                      // H tile → staging
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(stg + (uint32_t)(4 * j + sub) * kEpiRowBytes + (uint32_t)c4 * 4u),
                                         "f"(hn[j].x), "f"(hn[j].y), "f"(hn[j].z), "f"(hn[j].w) : "memory");
                        __syncwarp();
                        if (cj + 2 < nch32) fetch_h((cj + 2) * 32);      // hn is free again: the next chunk's H is in flight under this chunk's math and stores
                    }
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 16; j++) { v[j] = __uint_as_float(ra[j]); v[16 + j] = (c0 + 16 < bn) ? __uint_as_float(rb[j]) : 0.0f; }
                    if (EPI == UEPI_DSIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 hq;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(hq.x), "=f"(hq.y), "=f"(hq.z), "=f"(hq.w) : "r"(my_row + (uint32_t)j * 4u));
                            const float hs[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) { const float hh = hs[e] * kp; v[j + e] *= inv_keep * hh * (1.0f - hh); }   // keep*h*(1-h), h = hd*(1-p)
                        }
                        __syncwarp();
                    } else if (EPI == UEPI_SIGMOID_DROP) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float x = v[j];
                            float sg = __fdividef(1.0f, 1.0f + __expf(-x));
                            sg = x > 15.0f ? 1.0f : sg;
                            sg = x < -88.0f ? 0.0f : sg;
                            v[j] = sg;
                        }
                        if (a.drop_p > 0.0f) {
#pragma unroll
                            for (int j = 0; j < 32; j++) v[j] *= uniform24(a.seed, a.stream, ctr0 + (uint32_t)(c0 + j)) < keep_thr ? inv_keep : 0.0f;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = (c0 + j < a.N) ? v[j] : 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(my_row + (uint32_t)j * 4u), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const long gr = tile_row0 + 4 * j + sub;
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                     : "r"(stg + (uint32_t)(4 * j + sub) * kEpiRowBytes + (uint32_t)c4 * 4u));
                        if (gr < a.M && c0 + c4 < a.Nz) *reinterpret_cast<float4*>(a.C + gr * a.ldc + c0 + c4) = t;
                    }
                    __syncwarp();
                }
                tc_fence_before();
                mbar_arrive(tempty(acc));
                if (threadIdx.x == 192) UDBG(4096 + tcount * 4 + 1);
                continue;
            }
            // one 16-column chunk: registers → fused epilogue → 64 contiguous bytes of this thread's row.
            // Straight-line and branch-free so the 16 elements overlap in the pipes.
            auto process = [&](const uint32_t (&r)[16], int ci) {
                const int c0 = ci * 16;
                if (gm >= a.M || c0 >= a.Nz) return;
                float v[16];
                float* crow = a.C + gm * a.ldc + c0;
                if (EPI == UEPI_DSIGMOID) {
                    float hs[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 t = (c0 + j < a.Nz) ? ldg4(a.H + gm * a.ldh + c0 + j) : zero4();
                        hs[j] = t.x; hs[j + 1] = t.y; hs[j + 2] = t.z; hs[j + 3] = t.w;
                    }
                    const float kp = a.drop_p > 0.0f ? 1.0f - a.drop_p : 1.0f;
#pragma unroll
                    for (int j = 0; j < 16; j++) {                  // keep*h*(1-h) with h = hd*(1-p); hd == 0 → 0
                        const float hh = hs[j] * kp;
                        v[j] = __uint_as_float(r[j]) * (inv_keep * hh * (1.0f - hh));
                    }
                } else if (EPI == UEPI_SIGMOID_DROP) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float x = __uint_as_float(r[j]);
                        float sg = __fdividef(1.0f, 1.0f + __expf(-x));   // x < -88 → exp = inf → 0
                        sg = x > 15.0f ? 1.0f : sg;
                        sg = x < -88.0f ? 0.0f : sg;
                        v[j] = sg;
                    }
                    if (a.drop_p > 0.0f) {
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            v[j] *= uniform24(a.seed, a.stream, ctr0 + (uint32_t)(c0 + j)) < keep_thr ? inv_keep : 0.0f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = (c0 + j < a.N) ? v[j] : 0.0f;
                const int nvalid = min(16, a.Nz - c0);              // Nz and ldc are multiples of 4
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    if (j < nvalid) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            };
            uint32_t ra[16], rb[16];                                // ping-pong, statically indexed (registers)
            int ci = half;
            if (ci < nchunks) tmem_ld16_async(trow + (uint32_t)(ci * 16), ra);
            while (ci < nchunks) {
                tmem_ld_wait();
                if (ci + 2 < nchunks) tmem_ld16_async(trow + (uint32_t)((ci + 2) * 16), rb);    // in flight during the math below
                process(ra, ci);
                ci += 2;
                if (ci >= nchunks) break;
                tmem_ld_wait();
                if (ci + 2 < nchunks) tmem_ld16_async(trow + (uint32_t)((ci + 2) * 16), ra);
                process(rb, ci);
                ci += 2;
            }
            tc_fence_before();
            mbar_arrive(tempty(acc));
            if (threadIdx.x == 192) UDBG(4096 + tcount * 4 + 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, ncols);
}

// W (logical [rows, cols], stride ld) → hi / lo TF32 split, optionally transposed, into padded
// K-major operand buffers [orows_pad, ocols_pad] (stride old); padding stays zero.  All operand copies of a
// step are produced by ONE launch (blockIdx.y = job): the work is tiny, launch gaps would dominate.
struct SplitJob { const float* W; long ld; int rows, cols, transpose; float* hi; float* lo; long old; };
struct SplitJobs { SplitJob j[4]; };
__global__ void __launch_bounds__(256)
k_split_weights(SplitJobs jobs) {
    const SplitJob& jb = jobs.j[blockIdx.y];
    const long n = (long)jb.rows * jb.cols;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / jb.cols), c = (int)(i % jb.cols);
        const float x = jb.W[(long)r * jb.ld + c];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const long o = jb.transpose ? (long)c * jb.old + r : (long)r * jb.old + c;
        jb.hi[o] = h; jb.lo[o] = x - h;
    }
}

// -------------------------------------------------------------------------------------------------
// Weight-gradient GEMM: C[M,N] += Aᵀ·B with A [K, M] and B [K, N] row-major — the reduction runs over
// the batch (K = samples), so both operands are MN-major for the tensor core (model.go:56: d cost /
// d mlp0 = concatᵀ·dZ0, d cost/d mlp1 = h0ᵀ·dZ1).  Split-K over the persistent grid: every CTA
// reduces its own slice of the batch into two TMEM accumulators (rows 0..127 and 128..255 of C, all N
// columns) and adds them into C with red.global.add.v4.f32.
//   MN-major TF32 operands exist only in the SWIZZLE_128B_BASE32B shared-memory layout (32-byte swizzle
//   atoms inside 128-byte rows, pattern period 4 rows): TMA boxes are {32 fp32 along M/N, 32 samples}
//   written with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B — one 4 KB column block each — i.e. the canonical
//   layout ((8,n),(4,k)) with LBO = 4096 (next 32-wide column block) and SBO = 512 (next 4 samples);
//   a k-step of 8 samples starts 1024 bytes further on.
//   Converter warps round both operands to TF32 (round-to-nearest-even, in place) so that the single
//   tf32 product per k-step is unbiased; the 2^-12 relative rounding noise averages out over the batch.
// -------------------------------------------------------------------------------------------------
struct DwArgs {
    int K;                      // batch rows to reduce (rows >= K are zero-filled by TMA)
    int na, nb;                 // 32-wide column blocks of A (<= 8) and B (<= 8)
    int M, N;                   // valid rows / cols of C
    float* C; long ldc;
    int stages;
    int rawhi;                  // 3 terms: the raw tiles serve as the hi operands (no in-place truncation)
    int swap;                   // 1: accumulate Cᵀ — needs N (cols of C) <= 128; halves the MMA count when M > 128 >= N
    int tma;                    // producer: 0 = 2-D boxes from one thread, 1 = one lane per 2-D box, 2 = one 3-D box per operand (3-D maps)
    int ks;                     // samples per k-block / TMA box: 32 or 16
    int terms;                  // 3 = error-compensated 3xTF32 (hi/lo split of both operands, fp32-grade), 1 = one TF32-RN product per term
    int pf;                     // k-blocks requested into L2 ahead of the ring (0 = off)
    int rotate;                 // epilogue: CTA-dependent starting chunk (spreads the same-address red traffic)
    unsigned long long* tl;     // optional timeline of CTA 0 (globaltimer ns): start, first box landed, every 4th MMA batch, accumulators done, epilogue done
    int dbg;                    // probe only: 1 = epilogue adds 1.0, 2 = K-major descriptors, 3 = swap LBO/SBO, 4 = no swizzle-atom K offset
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t blk_bytes = 4096u) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(blk_bytes >> 4) << 16;     // LBO: next 32-wide column block along M/N (4096 for 32-sample boxes, 2048 for 16)
    d |= (uint64_t)(512 >> 4) << 32;           // SBO: next group of 4 samples along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ uint32_t idesc_tf32_mn(int M, int N) {
    return idesc_tf32(M, N) | (1u << 15) | (1u << 16);      // A and B MN-major
}
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u = __float_as_uint(x);
    u += 0xFFFu + ((u >> 13) & 1u);
    return __uint_as_float(u & 0xFFFFE000u);
}

__global__ void __launch_bounds__(448, 1)
k_umma_dw(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, DwArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = a.stages;
    const uint32_t blk = (uint32_t)a.ks * 128u;                      // one 32-column block of a k-block: ks samples x 128 bytes
    const uint32_t aBytes = (uint32_t)a.na * blk, bBytes = (uint32_t)a.nb * blk;
    const bool split = a.terms == 3;
    const uint32_t hiBytes = 8u * blk + bBytes;                      // A always reserves 8 blocks (two M halves)
    const uint32_t stBytes = hiBytes * (split ? 2u : 1u);            // stage: [A | B] and, for the 3-term product, [A_lo | B_lo] behind it
    auto sA = [&](int s) { return base + (uint32_t)s * stBytes; };
    auto sB = [&](int s) { return sA(s) + 8u * blk; };
    const uint32_t bars = base + (uint32_t)stages * stBytes;
    auto full = [&](int s) { return bars + 8u * s; };
    auto conv = [&](int s) { return bars + 8u * (stages + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * stages + s); };
    const uint32_t tdone = bars + 8u * (3 * stages);
    const uint32_t tmem_slot = tdone + 8u;
    const int bn = a.nb * 32;
    const uint32_t ncols = 512;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(full(s), 1); mbar_init(conv(s), 384); mbar_init(empty(s), 1); }
        mbar_init(tdone, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

#define DWDBG(idx) do { if (a.tl && blockIdx.x == 0 && (idx) < 1024) a.tl[(idx)] = gtimer(); } while (0)
    if (threadIdx.x == 0) DWDBG(0);
    // this CTA's slice of the batch, in ks-sample k-blocks
    const int total_kb = (a.K + a.ks - 1) / a.ks;
    const int per = (total_kb + gridDim.x - 1) / gridDim.x;
    const int kb0 = blockIdx.x * per, kb1 = min(total_kb, kb0 + per);
    const int nkb = max(0, kb1 - kb0);
    const bool two_halves = a.M > 128;

    if (warp == 0) {
        // TMA producer.  A k-block is na + nb column blocks of {32 fp32, ks samples}.  Issued as 2-D boxes by ONE thread the
        // main loop ran at ~12 boxes/µs whatever their size — the issue rate of the producer thread — so (a.tma == 2)
        // each operand comes as ONE 3-D box {32, ks, blocks} over a [32 | batch | width/32] view of the matrix, or
        // (a.tma == 1) every lane of the producer warp issues one of the 2-D boxes.
        if (a.tma == 1) {
            for (int it = 0; it < nkb; it++) {
                const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                mbar_wait(empty(s), ph ^ 1u);
                if (lane == 0) mbar_expect_tx(full(s), aBytes + bBytes);
                __syncwarp();
                const int row = (kb0 + it) * a.ks;
                if (lane < a.na) tma_load_2d(sA(s) + blk * lane, &tmA, full(s), 32 * lane, row);
                else if (lane < a.na + a.nb) tma_load_2d(sB(s) + blk * (lane - a.na), &tmB, full(s), 32 * (lane - a.na), row);
            }
        } else if (lane == 0) {
            auto prefetch_ab = [&](int j) {
                if (j >= nkb) return;
                const int prow = (kb0 + j) * a.ks;
                if (a.tma == 2) { tma_prefetch_3d(&tmA, 0, prow, 0); tma_prefetch_3d(&tmB, 0, prow, 0); return; }
                for (int q = 0; q < a.na; q++) tma_prefetch_2d(&tmA, 32 * q, prow);
                for (int q = 0; q < a.nb; q++) tma_prefetch_2d(&tmB, 32 * q, prow);
            };
            for (int j = stages; j < stages + a.pf; j++) prefetch_ab(j);
            for (int it = 0; it < nkb; it++) {
                const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                if (a.pf) prefetch_ab(it + stages + a.pf);
                mbar_wait(empty(s), ph ^ 1u);
                DWDBG(128 + it * 4 + 0);
                mbar_expect_tx(full(s), aBytes + bBytes);
                const int row = (kb0 + it) * a.ks;
                if (a.tma == 2) {
                    tma_load_3d(sA(s), &tmA, full(s), 0, row, 0);
                    tma_load_3d(sB(s), &tmB, full(s), 0, row, 0);
                } else {
                    for (int j = 0; j < a.na; j++) tma_load_2d(sA(s) + blk * j, &tmA, full(s), 32 * j, row);
                    for (int j = 0; j < a.nb; j++) tma_load_2d(sB(s) + blk * j, &tmB, full(s), 32 * j, row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = a.dbg == 2 ? idesc_tf32(kBlockM, bn) : idesc_tf32_mn(kBlockM, bn);
            const int ksteps = a.ks / 8;
            const uint32_t idesc_sw = idesc_tf32_mn(kBlockM, a.na * 32);
            const uint64_t lo_off = (uint64_t)(hiBytes >> 4);            // hi → lo copy of the same operand, in descriptor address units
            for (int it = 0; it < nkb; it++) {
                const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
                mbar_wait(conv(s), ph);
                DWDBG(128 + it * 4 + 3);
                tc_fence_after();
                uint64_t dA0 = desc_mn_sw128(sA(s), blk), dA1 = desc_mn_sw128(sA(s) + 4u * blk, blk), dB = desc_mn_sw128(sB(s), blk);
                if (a.dbg == 2) { dA0 = desc_k_sw128(sA(s)); dA1 = desc_k_sw128(sA(s) + 4u * 4096u); dB = desc_k_sw128(sB(s)); }
                if (a.dbg == 3) {   // LBO <-> SBO
                    auto sw = [](uint64_t d) { return (d & ~((uint64_t)0x3FFF << 16) & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(4096 >> 4) << 32); };
                    dA0 = sw(dA0); dA1 = sw(dA1); dB = sw(dB);
                }
                for (int k = 0; k < ksteps; k++) {                  // 8 samples per k-step = +1024 bytes = +64 in the address field
                    const uint64_t ko = (uint64_t)(k * 64);
                    const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                    if (a.swap) {       // Cᵀ: the narrow operand supplies the 128 accumulator rows, the wide one all N columns — one MMA per term
                        umma_tf32(tmem_base, dB + ko, dA0 + ko, idesc_sw, acc);
                        if (split) {
                            umma_tf32(tmem_base, dB + lo_off + ko, dA0 + ko, idesc_sw, 1u);
                            umma_tf32(tmem_base, dB + ko, dA0 + lo_off + ko, idesc_sw, 1u);
                        }
                        continue;
                    }
                    umma_tf32(tmem_base, dA0 + ko, dB + ko, idesc, acc);
                    if (split) {                                    // + A_lo·B_hi + A_hi·B_lo (A_lo·B_lo is below fp32 resolution)
                        umma_tf32(tmem_base, dA0 + lo_off + ko, dB + ko, idesc, 1u);
                        umma_tf32(tmem_base, dA0 + ko, dB + lo_off + ko, idesc, 1u);
                    }
                    if (two_halves) {
                        umma_tf32(tmem_base + 256u, dA1 + ko, dB + ko, idesc, acc);
                        if (split) {
                            umma_tf32(tmem_base + 256u, dA1 + lo_off + ko, dB + ko, idesc, 1u);
                            umma_tf32(tmem_base + 256u, dA1 + ko, dB + lo_off + ko, idesc, 1u);
                        }
                    }
                }
                umma_commit(empty(s));
                DWDBG(8 + it);
            }
            umma_commit(tdone);
        }
    } else {
        // converters.  1 term: TF32 round-to-nearest in place (unbiased single product).  3 terms: hi = the 19 bits the
        // tensor core reads, written in place, lo = x - hi (exact) into the stage's second half.  All twelve remaining warps
        // convert (the eight epilogue warps have nothing else to do until the accumulators are complete).
        const int c = threadIdx.x - 64;
        const int n16 = (int)((aBytes + bBytes) / 16u);
        for (int it = 0; it < nkb; it++) {
            const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
            mbar_wait(full(s), ph);
            if (c == 0 && it == 0) DWDBG(1);
            if (c == 0) DWDBG(128 + it * 4 + 1);
            const uint32_t pa = sA(s), pb = sB(s);
            const int na16 = (int)(aBytes / 16u);
            for (int i0 = c; i0 < n16; i0 += 384 * 4) {
                float4 x[4]; uint32_t ad[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + 384 * u;
                    ad[u] = i < na16 ? pa + 16u * i : pb + 16u * (i - na16);
                    if (i < n16) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "r"(ad[u]));
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + 384 * u >= n16) continue;
                    if (split) {
                        float4 hv, lv;
                        hv.x = __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u); lv.x = x[u].x - hv.x;
                        hv.y = __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u); lv.y = x[u].y - hv.y;
                        hv.z = __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u); lv.z = x[u].z - hv.z;
                        hv.w = __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u); lv.w = x[u].w - hv.w;
                        if (!a.rawhi) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(ad[u]), "f"(hv.x), "f"(hv.y), "f"(hv.z), "f"(hv.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(ad[u] + hiBytes), "f"(lv.x), "f"(lv.y), "f"(lv.z), "f"(lv.w) : "memory");
                    } else {
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(ad[u]), "f"(tf32_rn(x[u].x)), "f"(tf32_rn(x[u].y)),
                                     "f"(tf32_rn(x[u].z)), "f"(tf32_rn(x[u].w)) : "memory");
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(conv(s));
            if (c == 0) DWDBG(128 + it * 4 + 2);
        }
        if (warp >= 6 && nkb > 0) {                                 // epilogue warps 6..13: C += accumulators
            const int q = warp & 3, half = (warp - 6) >> 2;
            mbar_wait(tdone, 0);
            if (threadIdx.x == 192) DWDBG(2);
            tc_fence_after();
            if (a.swap) {
                // accumulator rows = columns of C (lane m = q*32 + lane), accumulator columns = rows of C: a warp's red.add of
                // one accumulator column covers 32 consecutive floats of one row of C
                const int m = q * 32 + lane;
                const int nck = a.na * 2;                                    // 16-column chunks of the accumulator
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
                const int mine = (nck - half + 1) / 2;
                const int rot = a.rotate ? (int)blockIdx.x : 0;
                if (q * 32 < a.N)
                    for (int kk = 0; kk < mine; kk++) {
                        const int ci = half + 2 * ((kk + rot) % mine);
                        uint32_t r[16];
                        tmem_ld16_async(trow + (uint32_t)(ci * 16), r);
                        tmem_ld_wait();
                        if (m < a.N) {
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                if (ci * 16 + j < a.M) atomicAdd(a.C + (long)(ci * 16 + j) * a.ldc + m, __uint_as_float(r[j]));
                        }
                    }
            } else {
            const int nchunks = bn / 16;
            // every CTA adds the same [M, N] block: walking it in the same order would put all 148 CTAs on the same
            // addresses at the same moment (same-address red ops serialise in L2), so each CTA starts at its own chunk
            const int nh = two_halves ? 2 : 1;
            const int per_half = (nchunks - half + 1) / 2;                   // chunks half, half + 2, ... of this warp
            const int rot = a.rotate ? (int)blockIdx.x : 0;
            for (int hh = 0; hh < nh; hh++) {
                const int h = (hh + rot) % nh;
                const int m = h * 128 + q * 32 + lane;
                const uint32_t trow = tmem_base + (uint32_t)(h * 256) + ((uint32_t)(q * 32) << 16);
                for (int kk = 0; kk < per_half; kk++) {
                    const int ci = half + 2 * ((kk + rot / nh) % per_half);
                    uint32_t r[16];
                    tmem_ld16_async(trow + (uint32_t)(ci * 16), r);
                    tmem_ld_wait();
                    if (m < a.M) {
                        float* crow = a.C + (long)m * a.ldc + ci * 16;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            if (ci * 16 + j < a.N)                  // N is a multiple of 4 on the padded grad storage
                                red_add4(crow + j, a.dbg == 1 ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                    }
                }
            }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) DWDBG(3);
    if (warp == 2) tmem_dealloc(tmem_base, ncols);
}

}  // namespace umma
}  // namespace ctr
