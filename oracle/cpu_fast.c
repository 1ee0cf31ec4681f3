/*
 * cpu_fast.c — the TIMED CPU arm of bench.py (cpu_baseline / --impl reference): one index-fed train step of
 * go-ctr's DIN / YouTube graph the way the reference's CPU path spends its time — float32 throughout, the three
 * dense layers as blocked, thread-parallel SGEMMs (the reference: gonum v0.11.0 Sgemm under gorgonia's Mul nodes,
 * din.go:307-315, model.go:189), everything else parallel over samples (the reference: rcmd.go:375 assembler
 * goroutines).  Same algorithm as ctr_oracle.c (the CHECKER, which accumulates in double and stays the parity
 * reference); tests/test_oracle_fast.py holds the two against each other.
 *
 * TEST INFRASTRUCTURE ONLY — see ctr_oracle.h.  Built with -O3 -mavx2 -mfma -fopenmp (no -march=native: the .so is
 * built in the CPU container and travels to the GPU box's host).
 *
 * The embedding-row update (engine extension — the reference never learns embeddings, din.go:161-169) is Hogwild
 * (plain float adds, racy across threads) like the reference's own item2vec trainer (optimizer.go:107-129).
 */
#include "ctr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Scratch that lives across steps: a step of 65 536 samples uses ~350 MB of activations; fresh mmap'd pages every step
 * would make 128 threads queue on page faults.  Test infrastructure: one caller at a time. */
#define NSCRATCH 16
static struct { void* p; size_t cap; } g_scr[NSCRATCH];
static void* scratch(int slot, size_t bytes, int zero) {
    if (g_scr[slot].cap < bytes) { free(g_scr[slot].p); g_scr[slot].p = malloc(bytes); g_scr[slot].cap = g_scr[slot].p ? bytes : 0; }
    if (zero && g_scr[slot].p) memset(g_scr[slot].p, 0, bytes);
    return g_scr[slot].p;
}

static inline float sig32(float x) { return x < -88.0f ? 0.0f : x > 15.0f ? 1.0f : 1.0f / (1.0f + expf(-x)); }

/* C[M,N] = A[M,K] · B[K,N] (all row-major).  Row blocks of 4 x column panels of 32: the 4x32 accumulator tile
 * lives in 16 AVX2 registers, B's panel row is loaded once per k and reused by the 4 rows. */
static void sgemm_nn(int M, int N, int K, const float* A, long lda, const float* B, long ldb, float* C, long ldc) {
#pragma omp parallel for schedule(static)
    for (int i0 = 0; i0 < M; i0 += 4) {
        const int mi = M - i0 < 4 ? M - i0 : 4;
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int nj = N - j0 < 32 ? N - j0 : 32;
            float acc[4][32];
            memset(acc, 0, sizeof acc);
            const float* a0 = A + (long)i0 * lda;
            for (int k = 0; k < K; k++) {
                const float* b = B + (long)k * ldb + j0;
                const float x0 = a0[k], x1 = mi > 1 ? a0[lda + k] : 0.0f, x2 = mi > 2 ? a0[2 * lda + k] : 0.0f, x3 = mi > 3 ? a0[3 * lda + k] : 0.0f;
#pragma omp simd
                for (int j = 0; j < 32; j++) {
                    const float bv = j < nj ? b[j] : 0.0f;
                    acc[0][j] += x0 * bv; acc[1][j] += x1 * bv; acc[2][j] += x2 * bv; acc[3][j] += x3 * bv;
                }
            }
            for (int i = 0; i < mi; i++) memcpy(C + (long)(i0 + i) * ldc + j0, acc[i], sizeof(float) * (size_t)nj);
        }
    }
}
/* C[M,N] = A[M,K] · Bᵀ, B is [N,K] row-major (dZ · Wᵀ): dot products along contiguous k */
static void sgemm_nt(int M, int N, int K, const float* A, long lda, const float* B, long ldb, float* C, long ldc) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; i++) {
        const float* a = A + (long)i * lda;
        for (int j = 0; j < N; j++) {
            const float* b = B + (long)j * ldb;
            float s = 0.0f;
#pragma omp simd reduction(+ : s)
            for (int k = 0; k < K; k++) s += a[k] * b[k];
            C[(long)i * ldc + j] = s;
        }
    }
}
/* C[M,N] = Aᵀ · B with A [K,M], B [K,N] (weight gradients: the reduction runs over the batch K).  Every thread
 * accumulates its slice of the batch into a private [M,N] tile, then the tiles are summed. */
static void sgemm_tn(int M, int N, int K, const float* A, long lda, const float* B, long ldb, float* C, long ldc) {
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    if (T > 64) T = 64;
    float* part = (float*)scratch(0, sizeof(float) * (size_t)T * M * N, 0);
#pragma omp parallel num_threads(T)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* P = part + (size_t)tid * M * N;
        memset(P, 0, sizeof(float) * (size_t)M * N);
        const long k0 = (long)K * tid / T, k1 = (long)K * (tid + 1) / T;
        for (long k = k0; k < k1; k++) {
            const float* a = A + k * lda; const float* b = B + k * ldb;
            for (int i = 0; i < M; i++) {
                const float x = a[i];
                if (x == 0.0f) continue;
                float* p = P + (long)i * N;
#pragma omp simd
                for (int j = 0; j < N; j++) p[j] += x * b[j];
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            float s = 0.0f;
            for (int t = 0; t < T; t++) s += part[((size_t)t * M + i) * N + j];
            C[(long)i * ldc + j] = s;
        }
}

/* One train step (forward, BCE, backward, Adam on the dense weights, Hogwild SGD on the touched rows).
 * Returns the batch cost (cost.go:9-17).  Same signature family as orc_train_step_idx. */
float orc_fast_train_step_idx(const orc_cfg* c, const orc_solver* s, orc_adam_state* st,
                              float* W0, float* W1, float* W2, float* att,
                              const float* user_feat, long ldu, const float* item_feat, long ldi,
                              float* item_emb, long lde, long n_items,
                              const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                              const float* y, int B, float table_lr, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)n_items;
    const int uP = c->uP, S = c->S, D = c->D, cF = c->cF, H0 = c->H0, H1 = c->H1, in = uP + 2 * D + cF;
    const float invS = 1.0f / (float)S;
    float* X0 = (float*)scratch(1, sizeof(float) * (size_t)B * in, 0);
    float* A = (float*)scratch(2, sizeof(float) * (size_t)B * S, 0);           /* gate values a_s (din.go:273) */
    float* H0d = (float*)scratch(3, sizeof(float) * (size_t)B * H0, 0);        /* post-dropout activations */
    float* H1d = (float*)scratch(4, sizeof(float) * (size_t)B * H1, 0);
    float* dZ1 = (float*)scratch(5, sizeof(float) * (size_t)B * H1, 0);
    float* dZ0 = (float*)scratch(6, sizeof(float) * (size_t)B * H0, 0);
    float* dX = (float*)scratch(7, sizeof(float) * (size_t)B * 2 * D, 0);
    float* P = (float*)scratch(8, sizeof(float) * (size_t)B, 0);
    const uint32_t step = (uint32_t)st->t;
    /* ---- gather + attention forward (rcmd.go:462-536, din.go:224-301 / dnn.go:164-170) */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        float* x = X0 + (long)b * in;
        const int ur = user_row[b], ir = item_row[b];
        if (ur >= 0) memcpy(x, user_feat + (long)ur * ldu, sizeof(float) * (size_t)uP); else memset(x, 0, sizeof(float) * (size_t)uP);
        float* pooled = x + uP; float* v = x + uP + D;
        if (ir >= 0) { memcpy(v, item_emb + (long)ir * lde, sizeof(float) * (size_t)D); memcpy(v + D, item_feat + (long)ir * ldi, sizeof(float) * (size_t)cF); }
        else memset(v, 0, sizeof(float) * (size_t)(D + cF));
        float ny2 = 0.0f;
        for (int k = 0; k < D; k++) { ny2 += v[k] * v[k]; pooled[k] = 0.0f; }
        const float ny = sqrtf(ny2);
        for (int t = 0; t < S; t++) {
            const int hr = hist[(long)b * S + t];
            float a = c->model == ORC_YOUTUBE ? 1.0f : sig32((c->model == ORC_DIN_COS ? 0.5f : 1.0f) * att[t]);   /* missing row: cos = 0 / dist = 0 */
            if (hr >= 0) {
                const float* u = item_emb + (long)hr * lde;
                if (c->model == ORC_DIN_COS) {
                    float dot = 0.0f, nx2 = 0.0f;
                    for (int k = 0; k < D; k++) { dot += u[k] * v[k]; nx2 += u[k] * u[k]; }
                    a = sig32((dot / (sqrtf(nx2) * ny + 1e-8f) + 1.0f) * 0.5f * att[t]);
                } else if (c->model == ORC_DIN_EUC) {
                    float d2 = 0.0f;
                    for (int k = 0; k < D; k++) { const float e = u[k] - v[k]; d2 += e * e; }
                    a = sig32((1.0f - sqrtf(d2)) * att[t]);
                }
                for (int k = 0; k < D; k++) pooled[k] += a * u[k];
            } else if (c->model == ORC_DIN_EUC) a = sig32((1.0f - ny) * att[t]);      /* zero row: dist = |v| */
            A[(long)b * S + t] = a;
        }
        for (int k = 0; k < D; k++) pooled[k] *= invS;                       /* G.Mean, din.go:298 */
    }
    /* ---- MLP forward (din.go:307-315) */
    sgemm_nn(B, H0, in, X0, in, W0, H0, H0d, H0);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)B * H0; i++) {
        float h = sig32(H0d[i]);
        if (c->d0 > 0.0f) h = orc_uniform24(s->seed, step * 4u + 0u, (uint64_t)i) < (1.0f - c->d0) ? h / (1.0f - c->d0) : 0.0f;
        H0d[i] = h;
    }
    sgemm_nn(B, H1, H0, H0d, H0, W1, H1, H1d, H1);
    double cost = 0.0;
    float* dW2p = NULL; int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    dW2p = (float*)calloc((size_t)T * H1, sizeof(float));
#pragma omp parallel for schedule(static) reduction(+ : cost)
    for (int b = 0; b < B; b++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* h1 = H1d + (long)b * H1; float z2 = 0.0f;
        for (int j = 0; j < H1; j++) {
            float h = sig32(h1[j]);
            if (c->d1 > 0.0f) h = orc_uniform24(s->seed, step * 4u + 1u, (uint64_t)b * H1 + j) < (1.0f - c->d1) ? h / (1.0f - c->d1) : 0.0f;
            h1[j] = h; z2 += h * W2[j];
        }
        const float p = sig32(z2);
        P[b] = p;
        cost += (double)(y[b] * logf(p) + (1.0f - y[b]) * logf(1.0f - p));          /* cost.go:9-17 */
        const float g2 = (p - y[b]) / (float)B;
        for (int j = 0; j < H1; j++) {
            const float hd = h1[j];
            dW2p[(long)tid * H1 + j] += g2 * hd;
            float ds = 0.0f;                                                          /* keep * h (1-h) from hd = h*keep */
            if (hd != 0.0f) { const float h = hd * (1.0f - c->d1); ds = (1.0f / (1.0f - c->d1)) * h * (1.0f - h); }
            dZ1[(long)b * H1 + j] = g2 * W2[j] * ds;
        }
    }
    float* g0 = (float*)malloc(sizeof(float) * (size_t)in * H0);
    float* g1 = (float*)malloc(sizeof(float) * (size_t)H0 * H1);
    float* g2v = (float*)calloc((size_t)H1, sizeof(float));
    float* ga = (float*)calloc((size_t)S, sizeof(float));
    for (int t = 0; t < T; t++) for (int j = 0; j < H1; j++) g2v[j] += dW2p[(long)t * H1 + j];
    free(dW2p);
    /* ---- backward through the dense layers */
    sgemm_tn(H0, H1, B, H0d, H0, dZ1, H1, g1, H1);                                    /* dW1 = h0ᵀ dZ1 */
    sgemm_nt(B, H0, H1, dZ1, H1, W1, H1, dZ0, H0);                                    /* dZ0 = dZ1 W1ᵀ ... */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)B * H0; i++) {
        const float hd = H0d[i]; float ds = 0.0f;
        if (hd != 0.0f) { const float h = hd * (1.0f - c->d0); ds = (1.0f / (1.0f - c->d0)) * h * (1.0f - h); }
        dZ0[i] *= ds;                                                                 /* ... ⊙ σ' */
    }
    sgemm_tn(in, H0, B, X0, in, dZ0, H0, g0, H0);                                     /* dW0 = xᵀ dZ0 */
    sgemm_nt(B, 2 * D, H0, dZ0, H0, W0 + (long)uP * H0, H0, dX, 2 * D);               /* d[pooled | item] */
    /* ---- attention backward + row updates (analytic reverse of din.go:231-298) */
    float* gap = (float*)calloc((size_t)T * S, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        const float* g = dX + (long)b * 2 * D; const float* gi = g + D;
        const float* v = X0 + (long)b * in + uP + D;
        float dv[512];
        float ny2 = 0.0f;
        for (int k = 0; k < D; k++) { dv[k] = gi[k]; ny2 += v[k] * v[k]; }
        const float ny = sqrtf(ny2);
        for (int t = 0; t < S; t++) {
            const int hr = hist[(long)b * S + t];
            if (hr < 0) continue;                       /* zero row: u = 0 → no gradient into att (g·u = 0) nor into a table row */
            float* u = item_emb + (long)hr * lde;
            const float a = A[(long)b * S + t];
            float du[512];
            if (c->model == ORC_YOUTUBE) { for (int k = 0; k < D; k++) du[k] = g[k] * invS; }
            else {
                float gu = 0.0f, dot = 0.0f, nx2 = 0.0f, d2 = 0.0f;
                for (int k = 0; k < D; k++) { gu += g[k] * u[k]; dot += u[k] * v[k]; nx2 += u[k] * u[k]; const float e = u[k] - v[k]; d2 += e * e; }
                const float dz = gu * invS * a * (1.0f - a);
                if (c->model == ORC_DIN_COS) {
                    const float nx = sqrtf(nx2), den = nx * ny + 1e-8f, cs = dot / den, w = (cs + 1.0f) * 0.5f;
                    gap[(long)tid * S + t] += dz * w;
                    const float cc = 0.5f * dz * att[t], iden = 1.0f / den;
                    const float ku = nx > 0.0f ? cs * ny / (nx * den) : 0.0f, kv = ny > 0.0f ? cs * nx / (ny * den) : 0.0f;
                    for (int k = 0; k < D; k++) { du[k] = a * invS * g[k] + cc * (v[k] * iden - ku * u[k]); dv[k] += cc * (u[k] * iden - kv * v[k]); }
                } else {
                    const float dist = sqrtf(d2), w = 1.0f - dist;
                    gap[(long)tid * S + t] += dz * w;
                    const float k2 = dist > 0.0f ? dz * att[t] / dist : 0.0f;
                    for (int k = 0; k < D; k++) { const float e = u[k] - v[k]; du[k] = a * invS * g[k] - k2 * e; dv[k] += k2 * e; }
                }
            }
            if (table_lr != 0.0f) for (int k = 0; k < D; k++) u[k] -= table_lr * du[k];      /* Hogwild */
        }
        const int ir = item_row[b];
        if (table_lr != 0.0f && ir >= 0) { float* e = item_emb + (long)ir * lde; for (int k = 0; k < D; k++) e[k] -= table_lr * dv[k]; }
    }
    for (int t = 0; t < T; t++) for (int j = 0; j < S; j++) ga[j] += gap[(long)t * S + j];
    free(gap);
    /* ---- solver.Step (model.go:88,192) */
    st->t++;
    orc_adam_step(W0, g0, st->m0, st->v0, (long)in * H0, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    orc_adam_step(W1, g1, st->m1, st->v1, (long)H0 * H1, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    orc_adam_step(W2, g2v, st->m2, st->v2, H1, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    if (c->model != ORC_YOUTUBE) orc_adam_step(att, ga, st->ma, st->va, S, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    free(g0); free(g1); free(g2v); free(ga);
    return -(float)(cost / (double)B);
}
