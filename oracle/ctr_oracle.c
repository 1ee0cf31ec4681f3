/*
 * ctr_oracle.c — CPU restatement of go-ctr's CTR hot path.  TEST INFRASTRUCTURE ONLY (see
 * ctr_oracle.h).  Plain C99, double accumulation with results rounded to float32 at the same op
 * boundaries where the reference's float32 graph (gorgonia, model.DT = Float32, model.go:14)
 * materialises a tensor.  Cites are relative to /root/reference.
 */
#include "ctr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Known-answer-test subjects
 * ---------------------------------------------------------------------------------------------- */

/* gorgonia v0.9.17 float32 sigmoid (third-party, restated from its published source:
 * operatorPointwise_unary.go `_sigmoidf32`): saturates outside [-88, 15].  UNPINNED. */
float orc_sigmoid32(float x) {
    if (x < -88.0f) return 0.0f;
    if (x > 15.0f) return 1.0f;
    return (float)(1.0 / (1.0 + exp(-(double)x)));
}

/* model/cost.go:9-17.  The constant float32(1.0+1e-8) is exactly 1.0f (cost.go:12), so p==1
 * yields log(0) = -Inf exactly as in the reference. */
float orc_bce32(const float* pred, const float* y, int n) {
    const float one_eps = (float)(1.0 + 1e-8);
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        float pos = (float)log((double)pred[i]) * y[i];
        float neg = (float)log((double)(one_eps - pred[i])) * (1.0f - y[i]);
        acc += (double)(float)(pos + neg);
    }
    return -(float)(acc / (double)n);
}

float orc_mse32(const float* pred, const float* y, int n) {  /* cost.go:20-23 */
    double acc = 0.0;
    for (int i = 0; i < n; i++) { float d = pred[i] - y[i]; acc += (double)(float)(d * d); }
    return (float)(acc / (double)n);
}

float orc_rms32(const float* pred, const float* y, int n) {  /* cost.go:26-29 */
    return (float)sqrt((double)orc_mse32(pred, y, n));
}

/* activation.go:11-16: ((x-|x|)*slope + (x+|x|)) * 0.5 */
void orc_prelu32(const float* x, float slope, int n, float* out) {
    for (int i = 0; i < n; i++) {
        float ax = fabsf(x[i]);
        float negative = (x[i] - ax) * slope;
        float positive = x[i] + ax;
        out[i] = (negative + positive) * 0.5f;
    }
}

/* activation.go:23-50: sqrt(sum((x-y)^2, last axis)), broadcasting the size-1 middle axis. */
int orc_euc_distance(const float* x, int sx, const float* y, int sy, int n, int d, float* out) {
    if (sx != sy && sx != 1 && sy != 1) return -1;
    int s = sx > sy ? sx : sy;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < s; j++) {
            const float* xr = x + ((long)i * sx + (sx == 1 ? 0 : j)) * d;
            const float* yr = y + ((long)i * sy + (sy == 1 ? 0 : j)) * d;
            double acc = 0.0;
            for (int k = 0; k < d; k++) { float df = xr[k] - yr[k]; acc += (double)(float)(df * df); }
            out[(long)i * s + j] = (float)sqrt((double)(float)acc);
        }
    return 0;
}

/* activation.go:57-83: sum(x*y) / (|x|*|y| + 1e-8) */
int orc_cosine(const float* x, int sx, const float* y, int sy, int n, int d, float* out) {
    if (sx != sy && sx != 1 && sy != 1) return -1;
    int s = sx > sy ? sx : sy;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < s; j++) {
            const float* xr = x + ((long)i * sx + (sx == 1 ? 0 : j)) * d;
            const float* yr = y + ((long)i * sy + (sy == 1 ? 0 : j)) * d;
            double dot = 0.0, nx = 0.0, ny = 0.0;
            for (int k = 0; k < d; k++) {
                dot += (double)(float)(xr[k] * yr[k]);
                nx += (double)(float)(xr[k] * xr[k]);
                ny += (double)(float)(yr[k] * yr[k]);
            }
            float xn = (float)sqrt((double)(float)nx), yn = (float)sqrt((double)(float)ny);
            float den = (float)(xn * yn) + 1e-8f;
            out[(long)i * s + j] = (float)dot / den;
        }
    return 0;
}

/* utils/util.go:131-148 → nn/metrics/ranking.go:144 → binaryClfCurve :13-57, ROCCurve :71-103,
 * AUC :106-118.  Labels binarised at 0.5; equal scores form one threshold group (:27-35). */
typedef struct { double s; int pos; } orc_sl;
static int orc_cmp_desc(const void* a, const void* b) {
    double x = ((const orc_sl*)a)->s, y = ((const orc_sl*)b)->s;
    return (x < y) - (x > y);
}
double orc_roc_auc(const float* pred, const float* y, int n) {
    if (n <= 0) return NAN;
    orc_sl* v = (orc_sl*)malloc(sizeof(orc_sl) * (size_t)n);
    for (int i = 0; i < n; i++) { v[i].s = (double)pred[i]; v[i].pos = y[i] > 0.5f; }
    qsort(v, (size_t)n, sizeof(orc_sl), orc_cmp_desc);
    double* fps = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    double* tps = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    int m = 0; double tp = 0, fp = 0;
    for (int i = 0; i < n;) {
        int j = i;
        while (j < n && !(v[j].s < v[i].s)) { if (v[j].pos) tp += 1; else fp += 1; j++; }
        fps[m] = fp; tps[m] = tp; m++;
        i = j;
    }
    /* ROCCurve :74-79: prepend (0,0) when the first threshold already has false positives */
    int off = 0;
    if (m == 0 || fps[0] != 0.0) off = 1;
    double fpmax = fps[m - 1], tpmax = tps[m - 1], auc = 0.0, xp = 0.0, yp = 0.0;
    if (fpmax <= 0.0 || tpmax <= 0.0) { free(v); free(fps); free(tps); return NAN; }
    for (int i = -off; i < m; i++) {
        double xx = i < 0 ? 0.0 : fps[i] * (1.0 / fpmax);
        double yy = i < 0 ? 0.0 : tps[i] * (1.0 / tpmax);
        auc += (xx - xp) * (yy + yp) / 2.0;
        xp = xx; yp = yy;
    }
    free(v); free(fps); free(tps);
    return auc;
}

/* feature/ubcache/cache.go:71-94 */
int orc_ub_filter(const int64_t* ts, int n, int64_t max_ts, int64_t max_len, int* start) {
    if (n == 0) { *start = 0; return 0; }
    if (max_ts == 0) max_ts = ts[0];
    int count = (int)max_len;
    if (count == 0) count = n;
    int i;
    for (i = 0; i < n; i++) if (ts[i] <= max_ts) break;
    if (i + count > n) count = n - i;
    *start = i;
    return count;
}

/* feature/multihot.go:26-35: FNV-1 32-bit (multiply then xor), int(sum32) % size */
int orc_hash_onehot32(const char* s, int size) {
    uint32_t h = 2166136261u;
    for (const unsigned char* p = (const unsigned char*)s; *p; p++) { h *= 16777619u; h ^= *p; }
    return (int)((int64_t)h % size);
}

/* ------------------------------------------------------------------------------------------------
 * Counter RNG.  The reference draws dropout masks and N(0,1) weights from time-seeded generators
 * (gorgonia UniformRandomNode / go_rng) that can never be reproduced; the engine and
 * the oracle instead share this *specification* (each implements it independently) so masks and
 * initial weights are bit-identical on both sides.
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_mix64(uint32_t seed, uint32_t stream, uint64_t ctr) {
    uint64_t z = (((uint64_t)seed << 32) | stream) * 0x9E3779B97F4A7C15ull + ctr * 0xD1B54A32D192ED03ull
                 + 0x632BE59BD9B4E019ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
/* dropout draws: 32-bit multiply-xorshift hash of (seed, stream, low 32 bits of the counter) */
static uint32_t orc_hash32(uint32_t seed, uint32_t stream, uint32_t ctr) {
    uint32_t x = ctr ^ (seed * 0x9E3779B1u) ^ (stream * 0x85EBCA77u + 0xC2B2AE3Du);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
float orc_uniform24(uint32_t seed, uint32_t stream, uint64_t ctr) {
    return (float)(orc_hash32(seed, stream, (uint32_t)ctr) >> 8) * (1.0f / 16777216.0f);
}
void orc_gaussian_init(float* w, long n, uint32_t seed, uint32_t stream) {
    for (long i = 0; i < n; i++) {   /* Box–Muller, one draw per element */
        uint64_t z = orc_mix64(seed, stream, (uint64_t)i);
        double u1 = ((double)(z >> 40) + 1.0) * (1.0 / 16777217.0);           /* (0,1) */
        double u2 = (double)((z >> 8) & 0xFFFFFFull) * (1.0 / 16777216.0);     /* [0,1) */
        w[i] = (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
    }
}

/* ------------------------------------------------------------------------------------------------
 * a1: recommend.GetSampleVector (rcmd.go:462-536), index form.  user feature (:473-481), history
 * embeddings most-recent-first with zero tail / zeros for a missing embedding (:509-530), target
 * item embedding or zeros (:501-505), item feature (:484-492); concat order (:533).
 * ---------------------------------------------------------------------------------------------- */
void orc_gather_rows(const float* user_feat, long ldu, const float* item_feat, long ldi,
                     const float* item_emb, long lde, int uP, int cF, int S, int D,
                     const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                     long B, float* X) {
    long xc = (long)uP + (long)S * D + D + cF;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; b++) {
        float* x = X + b * xc;
        if (user_row[b] >= 0) memcpy(x, user_feat + (long)user_row[b] * ldu, sizeof(float) * (size_t)uP);
        else memset(x, 0, sizeof(float) * (size_t)uP);   /* BatchPredict's zero row, rcmd.go:296-306 */
        float* ub = x + uP;
        for (int s = 0; s < S; s++) {
            int32_t r = hist[b * S + s];
            if (r >= 0) memcpy(ub + (long)s * D, item_emb + (long)r * lde, sizeof(float) * (size_t)D);
            else memset(ub + (long)s * D, 0, sizeof(float) * (size_t)D);
        }
        float* it = ub + (long)S * D;
        if (item_row[b] >= 0) {
            memcpy(it, item_emb + (long)item_row[b] * lde, sizeof(float) * (size_t)D);
            memcpy(it + D, item_feat + (long)item_row[b] * ldi, sizeof(float) * (size_t)cF);
        } else {
            memset(it, 0, sizeof(float) * (size_t)(D + cF));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Forward (a4 din.go:219-323, a5 dnn.go:162-184, a6 din.go:307-315, a7 cost.go)
 * ---------------------------------------------------------------------------------------------- */
struct orc_ws {
    int B, in, S, D, H0, H1;
    float *concat;  /* [B,in]  */
    float *ub;      /* [B,S*D] copy of the behaviour block */
    float *it;      /* [B,D]   */
    float *a;       /* [B,S]   sigmoid(w*att) */
    float *w;       /* [B,S]   attention weight before att0 */
    float *cs;      /* [B,S]   cosine (DIN_COS) or distance (DIN_EUC) */
    float *nx;      /* [B,S]   |ub_s| */
    float *ny;      /* [B]     |item| */
    float *den;     /* [B,S]   |ub_s||item| + 1e-8 */
    float *dot;     /* [B,S] */
    float *h0, *k0; /* [B,H0]  sigmoid output, dropout factor (0 or 1/(1-p)) */
    float *h1, *k1; /* [B,H1] */
    float *p;       /* [B] */
};

static float* orc_alloc(long n) { return (float*)calloc((size_t)(n > 0 ? n : 1), sizeof(float)); }

orc_ws* orc_ws_new(const orc_cfg* c, int B) {
    orc_ws* ws = (orc_ws*)calloc(1, sizeof(orc_ws));
    ws->B = B; ws->S = c->S; ws->D = c->D; ws->H0 = c->H0; ws->H1 = c->H1;
    ws->in = c->uP + 2 * c->D + c->cF;
    ws->concat = orc_alloc((long)B * ws->in);
    ws->ub = orc_alloc((long)B * c->S * c->D);
    ws->it = orc_alloc((long)B * c->D);
    ws->a = orc_alloc((long)B * c->S); ws->w = orc_alloc((long)B * c->S);
    ws->cs = orc_alloc((long)B * c->S); ws->nx = orc_alloc((long)B * c->S);
    ws->den = orc_alloc((long)B * c->S); ws->dot = orc_alloc((long)B * c->S);
    ws->ny = orc_alloc(B);
    ws->h0 = orc_alloc((long)B * c->H0); ws->k0 = orc_alloc((long)B * c->H0);
    ws->h1 = orc_alloc((long)B * c->H1); ws->k1 = orc_alloc((long)B * c->H1);
    ws->p = orc_alloc(B);
    return ws;
}
void orc_ws_free(orc_ws* ws) {
    if (!ws) return;
    free(ws->concat); free(ws->ub); free(ws->it); free(ws->a); free(ws->w); free(ws->cs);
    free(ws->nx); free(ws->ny); free(ws->den); free(ws->dot);
    free(ws->h0); free(ws->k0); free(ws->h1); free(ws->k1); free(ws->p); free(ws);
}

/* attention / pooling for one sample; writes pooled[D] */
static void orc_pool_one(const orc_cfg* c, const float* att, const float* ub, const float* it,
                         float* pooled, float* a_out, float* w_out, float* cs_out, float* nx_out,
                         float* ny_out, float* den_out, float* dot_out) {
    const int S = c->S, D = c->D;
    double acc[512];
    double* accp = D <= 512 ? acc : (double*)malloc(sizeof(double) * (size_t)D);
    for (int k = 0; k < D; k++) accp[k] = 0.0;
    if (c->model == ORC_YOUTUBE) {                      /* dnn.go:164-167: Mean over axis 1 */
        for (int s = 0; s < S; s++) for (int k = 0; k < D; k++) accp[k] += (double)ub[s * D + k];
        for (int s = 0; s < S; s++) { a_out[s] = 1.0f; w_out[s] = 0; cs_out[s] = 0; nx_out[s] = 0; den_out[s] = 0; dot_out[s] = 0; }
        *ny_out = 0;
    } else {
        double ny2 = 0.0;
        for (int k = 0; k < D; k++) ny2 += (double)(float)(it[k] * it[k]);
        float ny = (float)sqrt((double)(float)ny2);
        *ny_out = ny;
        for (int s = 0; s < S; s++) {
            const float* u = ub + (long)s * D;
            float w;
            if (c->model == ORC_DIN_COS) {              /* activation.go:75-82, din.go:231-237 */
                double dot = 0.0, nx2 = 0.0;
                for (int k = 0; k < D; k++) { dot += (double)(float)(u[k] * it[k]); nx2 += (double)(float)(u[k] * u[k]); }
                float nx = (float)sqrt((double)(float)nx2);
                float den = (float)(nx * ny) + 1e-8f;
                float cs = (float)dot / den;
                w = (cs + 1.0f) / 2.0f;
                cs_out[s] = cs; nx_out[s] = nx; den_out[s] = den; dot_out[s] = (float)dot;
            } else {                                    /* din.go:230 (euclidean variant) */
                double d2 = 0.0;
                for (int k = 0; k < D; k++) { float df = u[k] - it[k]; d2 += (double)(float)(df * df); }
                float dist = (float)sqrt((double)(float)d2);
                w = 1.0f - dist;
                cs_out[s] = dist; nx_out[s] = 0; den_out[s] = 0; dot_out[s] = 0;
            }
            float a = orc_sigmoid32(w * att[s]);       /* din.go:266-274 */
            a_out[s] = a; w_out[s] = w;
            for (int k = 0; k < D; k++) accp[k] += (double)(float)(u[k] * a);   /* :264-276 */
        }
    }
    for (int k = 0; k < D; k++) pooled[k] = (float)(accp[k] / (double)S);  /* G.Mean, din.go:298 */
    if (accp != acc) free(accp);
}

void orc_forward(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                 const float* X, long ldx, const orc_ranges* r, int B, int nvalid,
                 int training, uint32_t seed, uint32_t step, orc_ws* ws, float* p, float* logit) {
    const int uP = c->uP, S = c->S, D = c->D, cF = c->cF, H0 = c->H0, H1 = c->H1;
    const int in = uP + 2 * D + cF;
    orc_ws* lws = ws ? ws : orc_ws_new(c, B);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        float* cc = lws->concat + (long)b * in;
        float* ubc = lws->ub + (long)b * S * D;
        float* itc = lws->it + (long)b * D;
        if (b < nvalid) {
            const float* x = X + (long)b * ldx;
            memcpy(cc, x + r->up[0], sizeof(float) * (size_t)uP);
            memcpy(ubc, x + r->ub[0], sizeof(float) * (size_t)S * D);
            memcpy(itc, x + r->it[0], sizeof(float) * (size_t)D);
            memcpy(cc + uP + 2 * D, x + r->cx[0], sizeof(float) * (size_t)cF);
        } else {                                        /* FillTensorRows model.go:357-371 */
            memset(cc, 0, sizeof(float) * (size_t)in);
            memset(ubc, 0, sizeof(float) * (size_t)S * D);
            memset(itc, 0, sizeof(float) * (size_t)D);
        }
        orc_pool_one(c, att, ubc, itc, cc + uP, lws->a + (long)b * S, lws->w + (long)b * S,
                     lws->cs + (long)b * S, lws->nx + (long)b * S, lws->ny + b,
                     lws->den + (long)b * S, lws->dot + (long)b * S);
        memcpy(cc + uP + D, itc, sizeof(float) * (size_t)D);            /* concat, din.go:301 */
        /* layer 0: sigmoid(concat · W0) then dropout (din.go:307-308) */
        float* h0 = lws->h0 + (long)b * H0; float* k0 = lws->k0 + (long)b * H0;
        double z[1024];
        for (int j = 0; j < H0; j++) z[j] = 0.0;
        for (int k = 0; k < in; k++) { double xv = cc[k]; const float* wr = W0 + (long)k * H0; for (int j = 0; j < H0; j++) z[j] += xv * (double)wr[j]; }
        for (int j = 0; j < H0; j++) {
            h0[j] = orc_sigmoid32((float)z[j]);
            float keep = 1.0f;
            if (training && c->d0 > 0.0f)
                keep = orc_uniform24(seed, step * 4u + 0u, (uint64_t)b * H0 + j) < (1.0f - c->d0) ? 1.0f / (1.0f - c->d0) : 0.0f;
            k0[j] = keep;
        }
        /* layer 1 (din.go:311-312) */
        float* h1 = lws->h1 + (long)b * H1; float* k1 = lws->k1 + (long)b * H1;
        for (int j = 0; j < H1; j++) z[j] = 0.0;
        for (int k = 0; k < H0; k++) { double xv = (double)(float)(h0[k] * k0[k]); const float* wr = W1 + (long)k * H1; for (int j = 0; j < H1; j++) z[j] += xv * (double)wr[j]; }
        for (int j = 0; j < H1; j++) {
            h1[j] = orc_sigmoid32((float)z[j]);
            float keep = 1.0f;
            if (training && c->d1 > 0.0f)
                keep = orc_uniform24(seed, step * 4u + 1u, (uint64_t)b * H1 + j) < (1.0f - c->d1) ? 1.0f / (1.0f - c->d1) : 0.0f;
            k1[j] = keep;
        }
        /* layer 2 (din.go:315) */
        double z2 = 0.0;
        for (int k = 0; k < H1; k++) z2 += (double)(float)(h1[k] * k1[k]) * (double)W2[k];
        float pp = orc_sigmoid32((float)z2);
        lws->p[b] = pp;
        if (p) p[b] = pp;
        if (logit) logit[b] = (float)z2;
    }
    if (!ws) orc_ws_free(lws);
}

/* ------------------------------------------------------------------------------------------------
 * Backward (a8: what G.Grad(cost, Learnable...) model.go:56 differentiates; analytic).
 * dz2 uses the fused stable form (p - y)/B, equal to the chain through cost.go:9-17 and the
 * sigmoid wherever that chain is finite (the reference's op-by-op chain gives NaN when p==1.0f).
 * ---------------------------------------------------------------------------------------------- */
float orc_backward(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                   const orc_ws* ws, const float* y, int B,
                   float* dW0, float* dW1, float* dW2, float* datt, float* dUb, float* dIt) {
    const int uP = c->uP, S = c->S, D = c->D, H0 = c->H0, H1 = c->H1;
    const int in = ws->in;
    float cost = orc_bce32(ws->p, y, B);
    float* dz0 = orc_alloc((long)B * H0);
    float* dz1 = orc_alloc((long)B * H1);
    float* dz2 = orc_alloc(B);
    double* dattB = (double*)calloc((size_t)B * (size_t)S, sizeof(double));   /* per-sample datt terms */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        const float* h0 = ws->h0 + (long)b * H0; const float* k0 = ws->k0 + (long)b * H0;
        const float* h1 = ws->h1 + (long)b * H1; const float* k1 = ws->k1 + (long)b * H1;
        double g2 = ((double)ws->p[b] - (double)y[b]) / (double)B;
        dz2[b] = (float)g2;
        for (int j = 0; j < H1; j++)
            dz1[(long)b * H1 + j] = (float)(g2 * (double)W2[j] * (double)k1[j] * (double)h1[j] * (1.0 - (double)h1[j]));
        for (int i = 0; i < H0; i++) {
            double acc = 0.0; const float* wr = W1 + (long)i * H1;
            for (int j = 0; j < H1; j++) acc += (double)dz1[(long)b * H1 + j] * (double)wr[j];
            dz0[(long)b * H0 + i] = (float)(acc * (double)k0[i] * (double)h0[i] * (1.0 - (double)h0[i]));
        }
        /* d concat for the pooled and item columns */
        double g[512], gi[512];
        for (int k = 0; k < D; k++) {
            double a0 = 0.0, a1 = 0.0;
            const float* w0 = W0 + (long)(uP + k) * H0; const float* w1 = W0 + (long)(uP + D + k) * H0;
            for (int i = 0; i < H0; i++) { double d = dz0[(long)b * H0 + i]; a0 += d * (double)w0[i]; a1 += d * (double)w1[i]; }
            g[k] = (double)(float)a0; gi[k] = (double)(float)a1;
        }
        const float* ub = ws->ub + (long)b * S * D; const float* it = ws->it + (long)b * D;
        double dv[512];
        for (int k = 0; k < D; k++) dv[k] = gi[k];
        for (int s = 0; s < S; s++) {
            const float* u = ub + (long)s * D;
            float* du = dUb ? dUb + ((long)b * S + s) * D : NULL;
            if (c->model == ORC_YOUTUBE) {
                if (du) for (int k = 0; k < D; k++) du[k] = (float)(g[k] / (double)S);
                continue;
            }
            double a = ws->a[(long)b * S + s], w = ws->w[(long)b * S + s];
            double gu = 0.0;
            for (int k = 0; k < D; k++) gu += g[k] * (double)u[k];
            double da = gu / (double)S;
            double dz = da * a * (1.0 - a);
            dattB[(long)b * S + s] = dz * w;
            double dw = dz * (double)att[s];
            if (c->model == ORC_DIN_COS) {
                double cc = 0.5 * dw;
                double nx = ws->nx[(long)b * S + s], ny = ws->ny[b], den = ws->den[(long)b * S + s];
                double cs = ws->cs[(long)b * S + s];
                double ku = nx > 0.0 ? cs * ny / (nx * den) : 0.0;
                double kv = ny > 0.0 ? cs * nx / (ny * den) : 0.0;
                for (int k = 0; k < D; k++) {
                    if (du) du[k] = (float)(a * g[k] / (double)S + cc * ((double)it[k] / den - ku * (double)u[k]));
                    dv[k] += cc * ((double)u[k] / den - kv * (double)it[k]);
                }
            } else {
                double dist = ws->cs[(long)b * S + s];
                double inv = dist > 0.0 ? 1.0 / dist : 0.0;
                for (int k = 0; k < D; k++) {
                    double df = ((double)u[k] - (double)it[k]) * inv;
                    if (du) du[k] = (float)(a * g[k] / (double)S - dw * df);
                    dv[k] += dw * df;
                }
            }
        }
        if (dIt) for (int k = 0; k < D; k++) dIt[(long)b * D + k] = (float)dv[k];
    }
    /* weight grads: deterministic (each output element owned by one thread, b ascending) */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < in; k++) {
        double acc[1024];
        for (int i = 0; i < H0; i++) acc[i] = 0.0;
        for (int b = 0; b < B; b++) { double xv = ws->concat[(long)b * in + k]; const float* d = dz0 + (long)b * H0; for (int i = 0; i < H0; i++) acc[i] += xv * (double)d[i]; }
        for (int i = 0; i < H0; i++) dW0[(long)k * H0 + i] = (float)acc[i];
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < H0; i++) {
        double acc[1024];
        for (int j = 0; j < H1; j++) acc[j] = 0.0;
        for (int b = 0; b < B; b++) { double xv = (double)(float)(ws->h0[(long)b * H0 + i] * ws->k0[(long)b * H0 + i]); const float* d = dz1 + (long)b * H1; for (int j = 0; j < H1; j++) acc[j] += xv * (double)d[j]; }
        for (int j = 0; j < H1; j++) dW1[(long)i * H1 + j] = (float)acc[j];
    }
    for (int j = 0; j < H1; j++) {
        double acc = 0.0;
        for (int b = 0; b < B; b++) acc += (double)(float)(ws->h1[(long)b * H1 + j] * ws->k1[(long)b * H1 + j]) * (double)dz2[b];
        dW2[j] = (float)acc;
    }
    for (int s = 0; s < S; s++) {
        double acc = 0.0;
        for (int b = 0; b < B; b++) acc += dattB[(long)b * S + s];
        if (datt) datt[s] = (float)acc;
    }
    free(dz0); free(dz1); free(dz2); free(dattB);
    return cost;
}

/* ------------------------------------------------------------------------------------------------
 * a9: gorgonia v0.9.17 AdamSolver.Step (solvers.go; third-party, restated; UNPINNED), float32
 * branch, as configured at model.go:88: WithLearnRate(0.01), WithBatchSize(B), WithL2Reg(1e-4).
 *   iter++ ; c1 = 1-β1^iter ; c2 = 1-β2^iter
 *   g += l2·w ; g *= 1/batch
 *   m = β1 m + (1-β1) g ; v = β2 v + (1-β2) g²
 *   w -= η · (m/c1) / (sqrt(v/c2) + ε) ; g = 0
 * ---------------------------------------------------------------------------------------------- */
void orc_adam_step(float* w, float* g, float* m, float* v, long n, int t,
                   float lr, float l2, float batch, float b1, float b2, float eps) {
    float c1 = (float)(1.0 - pow((double)b1, (double)t));
    float c2 = (float)(1.0 - pow((double)b2, (double)t));
    float inv_b = batch > 1.0f ? 1.0f / batch : 1.0f;
    for (long i = 0; i < n; i++) {
        float gi = g[i];
        if (l2 != 0.0f) gi = gi + l2 * w[i];
        gi = gi * inv_b;
        float mi = b1 * m[i] + (1.0f - b1) * gi;
        float vi = b2 * v[i] + (1.0f - b2) * (gi * gi);
        m[i] = mi; v[i] = vi;
        float mh = mi / c1;
        float vh = (float)sqrt((double)(vi / c2)) + eps;
        w[i] = w[i] - lr * mh / vh;
        g[i] = 0.0f;
    }
}

/* ------------------------------------------------------------------------------------------------
 * a3: model.Train (model.go:27-213) and a10: model.Predict (model.go:242-353) on a dense X.
 * ---------------------------------------------------------------------------------------------- */
int orc_train_dense(const orc_cfg* c, const orc_solver* s, float* W0, float* W1, float* W2, float* att,
                    const float* X, const float* Y, long n, int xcols, const orc_ranges* r,
                    int batch, int epochs, int early_stop, float* last_cost, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const int in = c->uP + 2 * c->D + c->cF, H0 = c->H0, H1 = c->H1, S = c->S;
    long n0 = (long)in * H0, n1 = (long)H0 * H1;
    float *g0 = orc_alloc(n0), *g1 = orc_alloc(n1), *g2 = orc_alloc(H1), *ga = orc_alloc(S);
    float *m0 = orc_alloc(n0), *v0 = orc_alloc(n0), *m1 = orc_alloc(n1), *v1 = orc_alloc(n1);
    float *m2 = orc_alloc(H1), *v2 = orc_alloc(H1), *ma = orc_alloc(S), *va = orc_alloc(S);
    float* yb = orc_alloc(batch);
    orc_ws* ws = orc_ws_new(c, batch);
    long batches = n / batch + (n % batch != 0);                 /* model.go:96-99 */
    float best = INFINITY, cost = 0.0f; int no_improve = 0, t = 0, ep = 0;
    uint32_t step = 0;
    for (ep = 0; ep < epochs; ep++) {
        for (long b = 0; b < batches; b++) {
            long start = b * batch, end = start + batch;
            if (start >= n) break;
            if (end > n) end = n;
            int nv = (int)(end - start);
            for (int i = 0; i < batch; i++) yb[i] = i < nv ? Y[start + i] : 0.0f;  /* :174-184 */
            orc_forward(c, W0, W1, W2, att, X + start * xcols, xcols, r, batch, nv, 1, s->seed, step, ws, NULL, NULL);
            cost = orc_backward(c, W0, W1, W2, att, ws, yb, batch, g0, g1, g2, ga, NULL, NULL);
            t++;                                                  /* solver.Step model.go:192 */
            orc_adam_step(W0, g0, m0, v0, n0, t, s->lr, s->l2, (float)batch, s->b1, s->b2, s->eps);
            orc_adam_step(W1, g1, m1, v1, n1, t, s->lr, s->l2, (float)batch, s->b1, s->b2, s->eps);
            orc_adam_step(W2, g2, m2, v2, H1, t, s->lr, s->l2, (float)batch, s->b1, s->b2, s->eps);
            if (c->model != ORC_YOUTUBE)                          /* Learnable(): din.go:161-169 vs dnn.go:153 */
                orc_adam_step(att, ga, ma, va, S, t, s->lr, s->l2, (float)batch, s->b1, s->b2, s->eps);
            step++;
        }
        if (cost < best) { best = cost; no_improve = 0; } else no_improve++;   /* :198-204 */
        if (early_stop != 0 && no_improve >= early_stop) { ep++; break; }       /* :206-209 */
    }
    if (last_cost) *last_cost = cost;
    orc_ws_free(ws);
    free(g0); free(g1); free(g2); free(ga); free(m0); free(v0); free(m1); free(v1);
    free(m2); free(v2); free(ma); free(va); free(yb);
    return ep;
}

void orc_predict_dense(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                       const float* X, long n, int xcols, const orc_ranges* r, int batch, float* out) {
    long batches = n / batch + (n % batch != 0);
    float* pb = orc_alloc(batch);
    orc_cfg cp = *c; cp.d0 = 0.0f; cp.d1 = 0.0f;    /* prediction graph has no dropout: din.go:133-145 */
    for (long b = 0; b < batches; b++) {
        long start = b * batch, end = start + batch;
        if (start >= n) break;
        if (end > n) end = n;
        orc_forward(&cp, W0, W1, W2, att, X + start * xcols, xcols, r, batch, (int)(end - start), 0, 0, 0, NULL, pb, NULL);
        memcpy(out + start, pb, sizeof(float) * (size_t)(end - start));        /* model.go:344-347 */
    }
    free(pb);
}

/* ------------------------------------------------------------------------------------------------
 * Index-form train step (engine fast path; the reference materialises X first, rcmd.go:339-460,
 * then trains on it — this does the same per batch).  Row update is the engine's extension:
 * the reference never learns embeddings (din.go:161-169).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t row; long pos; } orc_rp;
static int orc_cmp_rp(const void* a, const void* b) {
    const orc_rp *x = (const orc_rp*)a, *y = (const orc_rp*)b;
    if (x->row != y->row) return x->row < y->row ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

float orc_train_step_idx(const orc_cfg* c, const orc_solver* s, orc_adam_state* st,
                         float* W0, float* W1, float* W2, float* att,
                         const float* user_feat, long ldu, const float* item_feat, long ldi,
                         float* item_emb, long lde, long n_items,
                         const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                         const float* y, int B, float table_lr, float* p_out, int nthreads,
                         float* emb_m, float* emb_v) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const int uP = c->uP, S = c->S, D = c->D, cF = c->cF, H0 = c->H0, H1 = c->H1;
    const int in = uP + 2 * D + cF;
    long xc = (long)uP + (long)S * D + D + cF;
    float* X = orc_alloc((long)B * xc);
    orc_gather_rows(user_feat, ldu, item_feat, ldi, item_emb, lde, uP, cF, S, D, user_row, item_row, hist, B, X);
    orc_ranges r = { {0, uP}, {uP, uP + S * D}, {uP + S * D, uP + S * D + D}, {uP + S * D + D, (int)xc} };
    orc_ws* ws = orc_ws_new(c, B);
    long n0 = (long)in * H0, n1 = (long)H0 * H1;
    float *g0 = orc_alloc(n0), *g1 = orc_alloc(n1), *g2 = orc_alloc(H1), *ga = orc_alloc(S);
    float* dUb = table_lr != 0.0f ? orc_alloc((long)B * S * D) : NULL;
    float* dIt = table_lr != 0.0f ? orc_alloc((long)B * D) : NULL;
    orc_forward(c, W0, W1, W2, att, X, xc, &r, B, B, 1, s->seed, (uint32_t)st->t, ws, p_out, NULL);
    float cost = orc_backward(c, W0, W1, W2, att, ws, y, B, g0, g1, g2, ga, dUb, dIt);
    st->t++;
    orc_adam_step(W0, g0, st->m0, st->v0, n0, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    orc_adam_step(W1, g1, st->m1, st->v1, n1, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    orc_adam_step(W2, g2, st->m2, st->v2, H1, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    if (c->model != ORC_YOUTUBE)
        orc_adam_step(att, ga, st->ma, st->va, S, st->t, s->lr, s->l2, (float)B, s->b1, s->b2, s->eps);
    if (table_lr != 0.0f) {
        /* accumulate duplicates in (b, slot) order in double, then one SGD update per row.
         * Sparse: sort the (row, position) pairs instead of clearing an n_items x D accumulator. */
        long npos = (long)B * (S + 1);
        int T = 1;
#ifdef _OPENMP
        T = omp_get_max_threads();
#endif
        /* every thread owns the rows with row % T == tid: it collects their (row, position) pairs in
         * position order, sorts by row (stable through the pos key) and applies its rows' updates */
#pragma omp parallel num_threads(T)
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            long m = 0, cap = npos / T + 1024;
            orc_rp* rp = (orc_rp*)malloc(sizeof(orc_rp) * (size_t)cap);
            for (long p = 0; p < npos; p++) {
                int b = (int)(p / (S + 1)), sl = (int)(p % (S + 1));
                int32_t row = sl < S ? hist[(long)b * S + sl] : item_row[b];
                if (row < 0 || row % T != tid) continue;
                if (m == cap) { cap *= 2; rp = (orc_rp*)realloc(rp, sizeof(orc_rp) * (size_t)cap); }
                rp[m].row = row; rp[m].pos = p; m++;
            }
            qsort(rp, (size_t)m, sizeof(orc_rp), orc_cmp_rp);
            for (long i = 0; i < m;) {
                long j = i; double acc[512];
                for (int k = 0; k < D; k++) acc[k] = 0.0;
                while (j < m && rp[j].row == rp[i].row) {
                    int b = (int)(rp[j].pos / (S + 1)), sl = (int)(rp[j].pos % (S + 1));
                    const float* gsrc = sl < S ? dUb + ((long)b * S + sl) * D : dIt + (long)b * D;
                    for (int k = 0; k < D; k++) acc[k] += (double)gsrc[k];
                    j++;
                }
                float* e = item_emb + (long)rp[i].row * lde;
                if (emb_m) {
                    /* engine extension CTR_TABLE_ADAM: the dense solver's update (model.go:88; orc_adam_step) on the
                     * touched row with its summed gradient — g *= 1/B first, no L2, shared step counter t */
                    float* mm = emb_m + (long)rp[i].row * lde; float* vv = emb_v + (long)rp[i].row * lde;
                    const float c1 = (float)(1.0 - pow((double)s->b1, (double)st->t)), c2 = (float)(1.0 - pow((double)s->b2, (double)st->t));
                    for (int k = 0; k < D; k++) {
                        float g = (float)acc[k] * (B > 1 ? 1.0f / (float)B : 1.0f);
                        mm[k] = s->b1 * mm[k] + (1.0f - s->b1) * g;
                        vv[k] = s->b2 * vv[k] + (1.0f - s->b2) * g * g;
                        e[k] -= table_lr * (mm[k] / c1) / (sqrtf(vv[k] / c2) + s->eps);
                    }
                } else
                for (int k = 0; k < D; k++) e[k] = (float)((double)e[k] - (double)table_lr * acc[k]);
                i = j;
            }
            free(rp);
        }
    }
    orc_ws_free(ws);
    free(X); free(g0); free(g1); free(g2); free(ga); free(dUb); free(dIt);
    return cost;
}
