/*
 * ctr_oracle.h — CPU restatement of go-ctr's CTR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link,
 * import or execute it — and there only as the checker or the reported CPU baseline, never as the
 * thing measured or shipped.  The product (go-ctr_b200/libctr_b200.so) contains none of this code
 * and has no CPU fallback.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * PARITY PINNING:
 *   pinned by the reference's own known-answer tests (tests/golden/kat.json):
 *     orc_bce32            model/cost_test.go:12-55   (0.8746, 0.3335 ±1e-4)
 *     orc_mse32/orc_rms32  model/cost_test.go:58-90
 *     orc_prelu32          model/activation_test.go:11-24
 *     orc_euc_distance     model/activation_test.go:26-148
 *     orc_cosine           model/activation_test.go:150-234
 *     orc_roc_auc          nn/metrics/ranking_test.go:9-42, utils/util_test.go:25-32 (0.75)
 *     orc_ub_filter        feature/ubcache/cache_test.go:9-38
 *   PARITY UNPINNED (no reference test, and gorgonia v0.9.17 / gonum v0.11.0 are third-party Go
 *   modules absent from /root/reference; no Go toolchain here): full DIN/YouTube forward scores,
 *   every gradient, the Adam trajectory, dropout, weight init.  For those the oracle restates the
 *   published algorithm (formulas cited per function) and is cross-checked by finite differences.
 */
#ifndef CTR_ORACLE_H
#define CTR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_YOUTUBE = 0, ORC_DIN_COS = 1, ORC_DIN_EUC = 2 };

typedef struct {
    int model;          /* ORC_* */
    int uP, S, D, cF;   /* uProfileDim, uBehaviorSize, uBehaviorDim(=iFeatureDim), cFeatureDim */
    int H0, H1;         /* 200, 80 in the reference (din.go:14-19) */
    float d0, d1;       /* dropout probabilities (din.go:204-205: .005; dnn.go:136-137: .003) */
} orc_cfg;

/* model.SampleInfo (rcmd.go:132-137): [start,end) column ranges inside a dense X row */
typedef struct { int up[2], ub[2], it[2], cx[2]; } orc_ranges;

/* ---- small known-answer-test subjects ------------------------------------------------------- */
float orc_bce32(const float* pred, const float* y, int n);                 /* cost.go:9-17 */
float orc_mse32(const float* pred, const float* y, int n);                 /* cost.go:20-23 */
float orc_rms32(const float* pred, const float* y, int n);                 /* cost.go:26-29 */
void  orc_prelu32(const float* x, float slope, int n, float* out);         /* activation.go:11-16 */
/* x:[n,sx,d], y:[n,sy,d] with sx==sy or one of them 1 (broadcast); out:[n,max(sx,sy)]; returns 0,
 * or -1 for unsupported shapes.  activation.go:23-50 / :57-83 */
int   orc_euc_distance(const float* x, int sx, const float* y, int sy, int n, int d, float* out);
int   orc_cosine(const float* x, int sx, const float* y, int sy, int n, int d, float* out);
float orc_sigmoid32(float x);                                              /* gorgonia _sigmoidf32 */
double orc_roc_auc(const float* pred, const float* y, int n);              /* util.go:131-148, ranking.go:13-150 */
/* ubcache TimeSeq.Filter (cache.go:71-94): ts desc-sorted; returns count, *start = first index */
int   orc_ub_filter(const int64_t* ts, int n, int64_t max_ts, int64_t max_len, int* start);
int   orc_hash_onehot32(const char* s, int size);                          /* feature/multihot.go:26-35 (FNV-1) */

/* ---- counter RNG shared by spec with the engine (dropout masks, weight init) ------------------ */
uint64_t orc_mix64(uint32_t seed, uint32_t stream, uint64_t ctr);
float    orc_uniform24(uint32_t seed, uint32_t stream, uint64_t ctr);      /* [0,1), 24 bits */
void     orc_gaussian_init(float* w, long n, uint32_t seed, uint32_t stream); /* N(0,1), din.go:187-191 */

/* ---- a1: GetSampleVector (rcmd.go:462-536), index form -------------------------------------- */
/* vec = [user_feat[user_row] | emb[hist[0..S)] (row<0 → zeros) | emb[item_row] | item_feat[item_row]] */
void orc_gather_rows(const float* user_feat, long ldu, const float* item_feat, long ldi,
                     const float* item_emb, long lde, int uP, int cF, int S, int D,
                     const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                     long B, float* X /* [B, uP+S*D+D+cF] */);

/* ---- a4–a7: forward ------------------------------------------------------------------------- */
typedef struct orc_ws orc_ws;   /* saved activations for backward */
orc_ws* orc_ws_new(const orc_cfg* c, int B);
void    orc_ws_free(orc_ws* ws);

/* One forward over a batch of B rows of dense X (rows >= nvalid are the zero-padded tail,
 * model.go:132-184,357-371).  training!=0 applies dropout with masks = orc_uniform24(seed,
 * step*4+layer, b*H+j) < 1-p, scaled 1/(1-p).  Writes p[B]; optional logit[B] (pre-sigmoid z2). */
void orc_forward(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                 const float* X, long ldx, const orc_ranges* r, int B, int nvalid,
                 int training, uint32_t seed, uint32_t step, orc_ws* ws, float* p, float* logit);

/* ---- a8: backward (analytic reverse of a4-a7; + d/d(ub), d/d(item) = engine extension) ------ */
/* y[B] (tail rows label 0).  Grads are of cost = BCE mean.  dUb [B,S*D], dIt [B,D] may be NULL.
 * dIt includes both the MLP-input path and the attention path.  Returns cost (cost.go:9-17). */
float orc_backward(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                   const orc_ws* ws, const float* y, int B,
                   float* dW0, float* dW1, float* dW2, float* datt, float* dUb, float* dIt);

/* ---- a9: gorgonia AdamSolver.Step (model.go:88,192) ----------------------------------------- */
void orc_adam_step(float* w, float* g, float* m, float* v, long n, int t,
                   float lr, float l2, float batch, float b1, float b2, float eps);

/* ---- a3/a10: model.Train / model.Predict on dense X ----------------------------------------- */
typedef struct {
    float lr, l2, b1, b2, eps;      /* 0.01, 1e-4, .9, .999, 1e-8 (model.go:88) */
    uint32_t seed;                  /* dropout seed */
} orc_solver;
/* weights are updated in place; returns epochs run; *last_cost = cost of last batch of last epoch
 * (model.go:198). */
int orc_train_dense(const orc_cfg* c, const orc_solver* s, float* W0, float* W1, float* W2, float* att,
                    const float* X, const float* Y, long n, int xcols, const orc_ranges* r,
                    int batch, int epochs, int early_stop, float* last_cost, int nthreads);
void orc_predict_dense(const orc_cfg* c, const float* W0, const float* W1, const float* W2, const float* att,
                       const float* X, long n, int xcols, const orc_ranges* r, int batch, float* out);

/* ---- index-form train step (engine fast path): gather → fwd → bwd → Adam(dense) → SGD(rows) --- */
typedef struct {
    float *m0, *v0, *m1, *v1, *m2, *v2, *ma, *va; int t;
} orc_adam_state;
/* table_lr == 0 → frozen table (reference behaviour, din.go:161-169). Duplicated rows accumulate
 * in (b, slot) order in double, then row -= table_lr * grad. Returns cost. */
/* the TIMED CPU arm (cpu_fast.c): same step, float32, blocked thread-parallel SGEMMs, Hogwild row update */
float orc_fast_train_step_idx(const orc_cfg* c, const orc_solver* s, orc_adam_state* st,
                              float* W0, float* W1, float* W2, float* att,
                              const float* user_feat, long ldu, const float* item_feat, long ldi,
                              float* item_emb, long lde, long n_items,
                              const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                              const float* y, int B, float table_lr, int nthreads);
float orc_train_step_idx(const orc_cfg* c, const orc_solver* s, orc_adam_state* st,
                         float* W0, float* W1, float* W2, float* att,
                         const float* user_feat, long ldu, const float* item_feat, long ldi,
                         float* item_emb, long lde, long n_items,
                         const int32_t* user_row, const int32_t* item_row, const int32_t* hist,
                         const float* y, int B, float table_lr, float* p_out, int nthreads,
                         float* emb_m, float* emb_v /* non-NULL: lazy Adam on the touched rows instead of SGD */);

#ifdef __cplusplus
}
#endif
#endif

/* ============================================================================================== *
 * item2vec (BASELINE config 5): embedding.TrainEmbedding (wordemb.go:9-32) =
 * wego word2vec SkipGram + HierarchicalSoftmax, float64, restated single-threaded.
 * PARITY UNPINNED: the reference's tests assert only dimensions / non-zero vectors
 * (wordemb_test.go:23) and the trainer is Hogwild + time-seeded (racy global LCG modelutil.go:22-29,
 * math/rand init word2vec.go:103-111) — not reproducible even against itself.
 * ============================================================================================== */
#ifndef CTR_ORACLE_I2V_H
#define CTR_ORACLE_I2V_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    int dim, window, iter;       /* wordemb.go:9: (window, dim, iter) ; rcmd.go:22-26: 16, 5, 1 */
    int min_count, max_depth;    /* options.go: 5, 100 */
    double init_lr, min_lr;      /* 0.025, 0.025e-4 */
    double subsample;            /* 1e-3 */
    int update_lr_batch;         /* 100000 */
    uint32_t seed;
    int rng_mode;                /* 0: del from the global LCG in token order (modelutil.go:26-29), 1: counter RNG per position (the parallelisable spec the engine uses) */
} orc_i2v_cfg;

/* Huffman tree exactly as dictionary.HuffnamTree builds it (huffman.go:23-57): stable sort by count,
 * merge the two smallest, insert the merged node before the first node with Val >= merged.Val.
 * parent[2V-1]: nodes 0..V-1 are the words, V..2V-2 the merged nodes in creation order (root = 2V-2,
 * parent[root] = -1); code[2V-1]: 0 = left, 1 = right.  literal != 0 uses the reference's O(V^2) array
 * procedure, 0 an O(V log V) equivalent.  V >= 2. */
void orc_i2v_huffman(const int64_t* count, int V, int literal, int32_t* parent, uint8_t* code);
/* path of word w: inner nodes from the root down, truncated like Node.GetPath(max_depth)
 * (node.go:26-43): returns n = number of (node, child code) steps, node ids are merged-node indices
 * 0..V-2 */
int orc_i2v_path(const int32_t* parent, const uint8_t* code, int V, int w, int max_depth, int32_t* nodes, uint8_t* codes);
/* full trainer.  tokens: word ids in [0,V) in corpus order (dictionary.Add assigns ids by first
 * appearance, dictionary.go:70-81).  emb_out [V, dim] float32 (GenEmbeddingMap32, word2vec.go:298).
 * syn1_out (may be NULL) [V-1, dim] float64 inner-node vectors.  Returns trained positions. */
long orc_i2v_train(const orc_i2v_cfg* c, const int32_t* tokens, long n, int V, float* emb_out, double* syn1_out);
/* the 1000-entry sigmoid table (sigmoid_table.go:28-45) */
double orc_i2v_sigmoid_lut(double x);
/* initial embedding value of element i: (u - 0.5)/dim with the counter RNG (word2vec.go:103-111 uses math/rand) */
double orc_i2v_init(uint32_t seed, long i, int dim);
#ifdef __cplusplus
}
#endif
#endif

/* ============================================================================================== *
 * float64 MLP (BASELINE configs[0]): model/mlp/mlp.go over
 * nn/neural_network/basemlp64.go.  Oracle only (no GPU path is in scope for this row).
 * Pinned: loss(θ=0)=ln 2 and gradient == finite differences (multilayer_perceptron_test.go:86-91,
 * :118-130); everything else PARITY UNPINNED (dataset of the reference's tests is third-party).
 * ============================================================================================== */
#ifndef CTR_ORACLE_MLP64_H
#define CTR_ORACLE_MLP64_H
#ifdef __cplusplus
extern "C" {
#endif
enum { ORC_MLP_RELU = 0, ORC_MLP_LOGISTIC = 1, ORC_MLP_IDENTITY = 2 };
typedef struct {
    int n_layers;            /* len(layerUnits) = hidden + 2 */
    int units[8];            /* [nFeatures, hidden..., nOutputs]  (basemlp64.go:495-497) */
    int hidden_act;          /* ORC_MLP_*; output is always logistic + binary_log_loss (:423-425) */
    int batch;               /* 200 (:233) */
    int max_iter;            /* 200 (:238) */
    int n_iter_no_change;    /* 10 (:254) */
    int shuffle;             /* true (:241) */
    int adaptive;            /* LearningRate == "adaptive" (feature_test.go:36) */
    uint32_t seed;
    double alpha;            /* 1e-4 (:232) */
    double lr_init;          /* 1e-3 (:235) */
    double beta1, beta2, eps;/* .9 .999 1e-8 (:250-252) */
    double tol;              /* 1e-4 (:243) */
} orc_mlp64_cfg;
typedef struct { double *ms, *vs; double beta1t, beta2t, t, lr_init, lr; } orc_mlp64_adam_state;

long   orc_mlp64_nparams(const orc_mlp64_cfg* c);
void   orc_mlp64_init(const orc_mlp64_cfg* c, double* params);
double orc_mlp64_loss_grad(const orc_mlp64_cfg* c, const double* params, const double* X, const double* y,
                           long n, double* grads);
void   orc_mlp64_adam(const orc_mlp64_cfg* c, orc_mlp64_adam_state* st, double* params, const double* grads, long np);
int    orc_mlp64_fit(const orc_mlp64_cfg* c, double* params, const double* X, const double* y, long n,
                     double* loss_curve, double* final_lr_init);
void   orc_mlp64_predict(const orc_mlp64_cfg* c, const double* params, const double* X, long n, double* out);
int    orc_mlp_fit_wrap(const orc_mlp64_cfg* c, double* params, const float* X32, const float* Y32, long n,
                        double* loss_curve);
void   orc_mlp_predict_wrap(const orc_mlp64_cfg* c, const double* params, const float* X32, long n, float* out);
#ifdef __cplusplus
}
#endif
#endif
