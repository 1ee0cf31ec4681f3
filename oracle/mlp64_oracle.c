/*
 * mlp64_oracle.c — CPU restatement of go-ctr's float64 MLP path (BASELINE
 * configs[0] "plumbing" case): model/mlp/mlp.go:15-65 (f32 → f64 adapters) over
 * nn/neural_network/basemlp64.go (`BaseMultilayerPerceptron64`, a scikit-learn MLP clone).
 *
 * TEST INFRASTRUCTURE ONLY (see ctr_oracle.h).  SURVEY §8d scopes this row as "oracle only, no GPU".
 *
 * PARITY: the reference's numeric tests for this class (multilayer_perceptron_test.go:82-101,164)
 * read the MicroChip dataset from a third-party module (github.com/pa-m/sklearn/datasets, go.mod:13)
 * that is absent from /root/reference ⇒ only the data-independent known answers are pinned:
 *   loss(θ=0) = ln 2 = 0.693 (multilayer_perceptron_test.go:86-91, chkLoss ±1e-3),
 *   model gradient == finite-difference gradient within 1e-4 (:118-130, restated on synthetic data).
 * Weight init and the epoch shuffle are time-seeded in the reference (basemlp64.go:448,500); here
 * they use the counter RNG of ctr_oracle.c (orc_mix64) so that runs are reproducible.
 */
#include "ctr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double u53(uint32_t seed, uint32_t stream, uint64_t ctr) {
    return (double)(orc_mix64(seed, stream, ctr) >> 11) * (1.0 / 9007199254740992.0);
}

/* packed parameter layout = basemlp64.go:459-463: per layer [intercepts(fo) | coefs(fi×fo row-major)] */
long orc_mlp64_nparams(const orc_mlp64_cfg* c) {
    long n = 0;
    for (int i = 0; i + 1 < c->n_layers; i++) n += (long)(1 + c->units[i]) * c->units[i + 1];
    return n;
}
static long layer_off(const orc_mlp64_cfg* c, int layer) {
    long n = 0;
    for (int i = 0; i < layer; i++) n += (long)(1 + c->units[i]) * c->units[i + 1];
    return n;
}

/* initialize, basemlp64.go:459-476: every packed element (intercepts too) = U[0,1)·sqrt(f/(fi+fo)),
 * f = 6, or 2 for logistic hidden activation — non-negative, unlike sklearn's U(−b, b). */
void orc_mlp64_init(const orc_mlp64_cfg* c, double* params) {
    long pos = 0;
    for (int i = 0; i + 1 < c->n_layers; i++) {
        int fi = c->units[i], fo = c->units[i + 1];
        double factor = c->hidden_act == ORC_MLP_LOGISTIC ? 2.0 : 6.0;
        double bound = sqrt(factor / (double)(fi + fo));
        long end = pos + (long)(1 + fi) * fo;
        for (; pos < end; pos++) params[pos] = u53(c->seed, 7, (uint64_t)pos) * bound;
    }
}

/* forwardPass, basemlp64.go:259-275.  act[l] is [n, units[l]]; act[0] = X. */
static void forward(const orc_mlp64_cfg* c, const double* params, double** act, long n) {
    int L = c->n_layers;
    for (int l = 0; l + 1 < L; l++) {
        int fi = c->units[l], fo = c->units[l + 1];
        const double* b = params + layer_off(c, l);
        const double* W = b + fo;
        const double* A = act[l];
        double* Z = act[l + 1];
        int last = (l + 1 == L - 1);
#pragma omp parallel for schedule(static)
        for (long r = 0; r < n; r++) {
            double* z = Z + r * fo;
            for (int j = 0; j < fo; j++) z[j] = 0.0;
            const double* a = A + r * fi;
            for (int k = 0; k < fi; k++) {
                double ak = a[k];
                const double* w = W + (long)k * fo;
                for (int j = 0; j < fo; j++) z[j] += ak * w[j];
            }
            for (int j = 0; j < fo; j++) z[j] += b[j];                 /* addIntercepts64 :205 */
            if (!last) {
                if (c->hidden_act == ORC_MLP_RELU) { for (int j = 0; j < fo; j++) if (z[j] < 0) z[j] = 0; }       /* :96-104 */
                else if (c->hidden_act == ORC_MLP_LOGISTIC) { for (int j = 0; j < fo; j++) z[j] = 1 / (1 + exp(-z[j])); }  /* :82-88 */
            } else {
                for (int j = 0; j < fo; j++) z[j] = 1 / (1 + exp(-z[j])); /* out_activation logistic, :423-425 */
            }
        }
    }
}

/* binary_log_loss, basemlp64.go:180-195 */
static double binary_log_loss(const double* y, const double* h, long n, int cols) {
    double sum = 0, hmin = nextafter(0.0, 1.0), hmax = nextafter(1.0, 0.0);
    for (long i = 0; i < n * cols; i++) {
        double hv = h[i];
        if (hv < hmin) hv = hmin; else if (hv > hmax) hv = hmax;
        sum += -y[i] * log(hv) - (1 - y[i]) * log1p(-hv);
    }
    return sum / (double)n;
}

typedef struct { double** act; double** del; long cap; } mlp_ws;
static mlp_ws* ws_new(const orc_mlp64_cfg* c, long n) {
    mlp_ws* w = (mlp_ws*)calloc(1, sizeof(*w));
    w->act = (double**)calloc(c->n_layers, sizeof(double*));
    w->del = (double**)calloc(c->n_layers, sizeof(double*));
    for (int l = 1; l < c->n_layers; l++) {
        w->act[l] = (double*)malloc(sizeof(double) * n * c->units[l]);
        w->del[l] = (double*)malloc(sizeof(double) * n * c->units[l]);
    }
    w->cap = n;
    return w;
}
static void ws_free(const orc_mlp64_cfg* c, mlp_ws* w) {
    for (int l = 1; l < c->n_layers; l++) { free(w->act[l]); free(w->del[l]); }
    free(w->act); free(w->del); free(w);
}

/* backprop, basemlp64.go:340-406 (+computeLossGrad :322-331).  Returns the batch loss incl. the L2
 * term; grads in the packed layout. */
static double backprop(const orc_mlp64_cfg* c, const double* params, const double* X, const double* y,
                       long n, mlp_ws* ws, double* grads) {
    int L = c->n_layers;
    ws->act[0] = (double*)X;
    forward(c, params, ws->act, n);
    int no = c->units[L - 1];
    double loss = binary_log_loss(y, ws->act[L - 1], n, no);
    double sq = 0;                                                      /* sumCoefSquares :311-319 */
    for (int l = 0; l + 1 < L; l++) {
        const double* W = params + layer_off(c, l) + c->units[l + 1];
        long cnt = (long)c->units[l] * c->units[l + 1];
        for (long i = 0; i < cnt; i++) sq += W[i] * W[i];
    }
    loss += (0.5 * c->alpha) * sq / (double)n;                           /* :361 */
    for (long i = 0; i < n * no; i++) ws->del[L - 1][i] = ws->act[L - 1][i] - y[i];  /* :373-381 */
    for (int l = L - 2; l >= 0; l--) {
        int fi = c->units[l], fo = c->units[l + 1];
        const double* b = params + layer_off(c, l);
        const double* W = b + fo;
        double* gb = grads + layer_off(c, l);
        double* gW = gb + fo;
        const double* A = ws->act[l];
        const double* Dl = ws->del[l + 1];
        /* coefGrads = actᵀ·delta / n + alpha/n · coefs  (:326-327) */
#pragma omp parallel for schedule(static)
        for (int k = 0; k < fi; k++) {
            double* g = gW + (long)k * fo;
            for (int j = 0; j < fo; j++) g[j] = 0.0;
            for (long r = 0; r < n; r++) {
                double a = A[r * fi + k];
                if (a == 0.0) continue;
                const double* d = Dl + r * fo;
                for (int j = 0; j < fo; j++) g[j] += a * d[j];
            }
            for (int j = 0; j < fo; j++) g[j] = g[j] / (double)n + c->alpha / (double)n * W[(long)k * fo + j];
        }
        for (int j = 0; j < fo; j++) {                                   /* matRowMean64 :213-226 */
            double s = 0;
            for (long r = 0; r < n; r++) s += Dl[r * fo + j];
            gb[j] = s / (double)n;
        }
        if (l >= 1) {                                                    /* :388-396 */
            double* Dp = ws->del[l];
#pragma omp parallel for schedule(static)
            for (long r = 0; r < n; r++) {
                const double* d = Dl + r * fo;
                const double* z = A + r * fi;
                double* dp = Dp + r * fi;
                for (int k = 0; k < fi; k++) {
                    const double* w = W + (long)k * fo;
                    double s = 0;
                    for (int j = 0; j < fo; j++) s += d[j] * w[j];
                    if (c->hidden_act == ORC_MLP_RELU) { if (z[k] == 0.0) s = 0.0; }        /* :140-148 */
                    else if (c->hidden_act == ORC_MLP_LOGISTIC) s *= z[k] * (1 - z[k]);      /* :124-131 */
                    dp[k] = s;
                }
            }
        }
    }
    return loss;
}

double orc_mlp64_loss_grad(const orc_mlp64_cfg* c, const double* params, const double* X, const double* y,
                           long n, double* grads) {
    mlp_ws* ws = ws_new(c, n);
    double* g = grads ? grads : (double*)malloc(sizeof(double) * orc_mlp64_nparams(c));
    double loss = backprop(c, params, X, y, n, ws, g);
    if (!grads) free(g);
    ws_free(c, ws);
    return loss;
}

/* AdamOptimizer64.updateParams, basemlp64.go:1075-1091 — note the β-powers advance once per
 * parameter ELEMENT, not once per step, so the bias correction vanishes within one call. */
void orc_mlp64_adam(const orc_mlp64_cfg* c, orc_mlp64_adam_state* st, double* params, const double* grads, long np) {
    if (st->t == 0) {
        memset(st->ms, 0, sizeof(double) * np);
        memset(st->vs, 0, sizeof(double) * np);
        st->beta1t = 1; st->beta2t = 1;
    }
    st->t += 1;
    for (long i = 0; i < np; i++) {
        double g = grads[i];
        st->ms[i] = c->beta1 * st->ms[i] + (1 - c->beta1) * g;
        st->vs[i] = c->beta2 * st->vs[i] + (1 - c->beta2) * g * g;
        st->beta1t *= c->beta1;
        st->beta2t *= c->beta2;
        st->lr = st->lr_init * sqrt(1 - st->beta2t) / (1. - st->beta1t);
        params[i] += -st->lr * st->ms[i] / (sqrt(st->vs[i]) + c->eps);
    }
}

/* fit → fitStochastic, basemlp64.go:484-567,729-857 (solver adam, no validation split).  X [n, units[0]]
 * and y [n, units[L-1]] are read through a permutation instead of being swapped in place (:788) and
 * restored (:855) — same visiting order.  loss_curve (may be NULL) receives one value per epoch.
 * Returns the number of epochs run (NIter). */
int orc_mlp64_fit(const orc_mlp64_cfg* c, double* params, const double* X, const double* y, long n,
                  double* loss_curve, double* final_lr_init) {
    int L = c->n_layers, nin = c->units[0], no = c->units[L - 1];
    long np = orc_mlp64_nparams(c);
    long bs = c->batch;
    if (bs <= 0) { bs = n < 200 ? n : 200; } else if (bs > n) bs = n;   /* :517-527 */
    mlp_ws* ws = ws_new(c, bs);
    double* grads = (double*)malloc(sizeof(double) * np);
    double* Xb = (double*)malloc(sizeof(double) * bs * nin);
    double* yb = (double*)malloc(sizeof(double) * bs * no);
    long* idx = (long*)malloc(sizeof(long) * n);
    for (long i = 0; i < n; i++) idx[i] = i;
    orc_mlp64_adam_state st;
    memset(&st, 0, sizeof(st));
    st.ms = (double*)malloc(sizeof(double) * np);
    st.vs = (double*)malloc(sizeof(double) * np);
    st.lr_init = c->lr_init; st.lr = c->lr_init;
    double best = INFINITY;
    int no_improve = 0, it = 0;
    for (it = 0; it < c->max_iter;) {
        if (c->shuffle) {                                               /* rand.Shuffle: Fisher–Yates from the top */
            for (long i = n - 1; i > 0; i--) {
                long j = (long)(orc_mix64(c->seed, 100 + (uint32_t)it, (uint64_t)i) % (uint64_t)(i + 1));
                long t = idx[i]; idx[i] = idx[j]; idx[j] = t;
            }
        }
        double acc = 0;
        for (long b0 = 0; b0 < n; b0 += bs) {                           /* :791-808 */
            long b1 = b0 + bs > n ? n : b0 + bs, m = b1 - b0;
            for (long r = 0; r < m; r++) {
                memcpy(Xb + r * nin, X + idx[b0 + r] * nin, sizeof(double) * nin);
                memcpy(yb + r * no, y + idx[b0 + r] * no, sizeof(double) * no);
            }
            double bl = backprop(c, params, Xb, yb, m, ws, grads);
            acc += bl * (double)m;
            orc_mlp64_adam(c, &st, params, grads, np);
        }
        it++;
        double loss = acc / (double)n;                                  /* :810-811 */
        if (loss_curve) loss_curve[it - 1] = loss;
        if (loss > best - c->tol) no_improve++; else no_improve = 0;    /* :886-892 */
        if (loss < best) best = loss;
        if (no_improve > c->n_iter_no_change) {                         /* :826-840 */
            if (!c->adaptive) break;                                    /* triggerStopping :1053-1058 */
            if (st.lr <= 1e-6) break;                                   /* :1059-1064 */
            st.lr_init *= .8;                                           /* :1065 */
            no_improve = 0;
        }
    }
    if (final_lr_init) *final_lr_init = st.lr_init;
    ws->act[0] = NULL;
    free(st.ms); free(st.vs); free(idx); free(Xb); free(yb); free(grads);
    ws_free(c, ws);
    return it;
}

/* predict → predictProbas, basemlp64.go:897-931: with 0/1 labels no LabelBinarizer is attached
 * (:590-594) and the float64 class returns the raw probabilities (no toLogits). */
void orc_mlp64_predict(const orc_mlp64_cfg* c, const double* params, const double* X, long n, double* out) {
    mlp_ws* ws = ws_new(c, n);
    ws->act[0] = (double*)X;
    forward(c, params, ws->act, n);
    memcpy(out, ws->act[c->n_layers - 1], sizeof(double) * n * c->units[c->n_layers - 1]);
    ws_free(c, ws);
}

/* SimpleMlpFitWrap.Fit / SimpleMlpPredWrap.Predict, model/mlp/mlp.go:45-65,15-39: float32 samples →
 * float64, fit, and probabilities → float32. */
int orc_mlp_fit_wrap(const orc_mlp64_cfg* c, double* params, const float* X32, const float* Y32, long n,
                     double* loss_curve) {
    int nin = c->units[0];
    double* X = (double*)malloc(sizeof(double) * n * nin);
    double* y = (double*)malloc(sizeof(double) * n);
    for (long i = 0; i < n * nin; i++) X[i] = (double)X32[i];
    for (long i = 0; i < n; i++) y[i] = (double)Y32[i];
    orc_mlp64_init(c, params);
    int it = orc_mlp64_fit(c, params, X, y, n, loss_curve, NULL);
    free(X); free(y);
    return it;
}
void orc_mlp_predict_wrap(const orc_mlp64_cfg* c, const double* params, const float* X32, long n, float* out) {
    int nin = c->units[0];
    double* X = (double*)malloc(sizeof(double) * n * nin);
    double* p = (double*)malloc(sizeof(double) * n);
    for (long i = 0; i < n * nin; i++) X[i] = (double)X32[i];
    orc_mlp64_predict(c, params, X, n, p);
    for (long i = 0; i < n; i++) out[i] = (float)p[i];
    free(X); free(p);
}
