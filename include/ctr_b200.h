/*
 * ctr_b200.h — C ABI of libctr_b200.so: the B200 (sm_100a) engine behind go-ctr's CTR hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one Go-side interface of the reference
 * (auxten/go-ctr @ c181363c; file:line relative to the repository root) and is exactly what a cgo
 * shim binds (see INTEGRATION.md and go/ctrb200/ctrb200.go).  Plain pointers and sizes only — no
 * torch / C++ types.  The caller owns every host buffer; the library copies and never retains a
 * caller pointer after return (cgo pointer-passing rule).  Every function returns 0 on success or a
 * CTR_E* code; the message is available from ctr_last_error().  Nothing aborts or exits.
 * There is no CPU fallback: without a CUDA device ctr_create fails with CTR_ENODEV.
 *
 * Layouts: all matrices row-major float32; indices int32 dense row ids (the id→row map stays on
 * the Go side where the reference keeps it, rcmd.go:472-505); -1 = missing row → zeros
 * (rcmd.go:502-505,519).
 */
#ifndef CTR_B200_H
#define CTR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_B200_ABI_VERSION 1

enum { CTR_OK = 0, CTR_EINVAL = 1, CTR_ENODEV = 2, CTR_ECUDA = 3, CTR_ENOMEM = 4, CTR_ESTATE = 5, CTR_ECOMM = 6, CTR_ENOTFOUND = 7, CTR_EIO = 8 };

/* which graph: model/youtube/dnn.go:162-184 | model/din/din.go:219-323 (cosine, live) |
 * din.go:230 + model/activation.go:23-50 (euclidean ActivationUnit variant) */
enum { CTR_MODEL_YOUTUBE = 0, CTR_MODEL_DIN_COS = 1, CTR_MODEL_DIN_EUC = 2 };

enum { CTR_TABLE_USER_FEAT = 0, CTR_TABLE_ITEM_FEAT = 1, CTR_TABLE_ITEM_EMB = 2 };

/* embedding-table optimiser.  FROZEN is the reference's behaviour (embeddings are inputs, not
 * learnables: din.go:161-169).  SGD fuses gradient scatter-add + update into the backward kernel
 * with red.global.add.v4.f32 (order-nondeterministic, Hogwild within a batch like the reference's
 * own item2vec trainer).  SGD_DETERMINISTIC sorts (row, sample, slot) keys and reduces each row's
 * segment in a fixed order — the parity-test mode.  ADAM applies the dense solver's update (model.go:88: Adam, step
 * table_lr, cfg betas / eps, gradient divided by the batch size first, no L2) once per DISTINCT row of the batch with
 * that row's summed gradient (sorted keys → segment reduction, deterministic); rows the batch did not touch keep their
 * moments ("lazy" Adam).  Single GPU or replicated table; the moments are part of the checkpoint. */
enum { CTR_TABLE_FROZEN = 0, CTR_TABLE_SGD = 1, CTR_TABLE_SGD_DETERMINISTIC = 2, CTR_TABLE_ADAM = 3 };

/* dense-layer GEMM engine: exact fp32 FFMA, or tcgen05 3xTF32 (error-compensated, fp32 accumulate
 * in TMEM).  AUTO = tcgen05 when the shape qualifies. */
enum { CTR_GEMM_AUTO = 0, CTR_GEMM_FP32 = 1, CTR_GEMM_TCGEN05_3XTF32 = 2 };

typedef struct ctr_handle ctr_handle;

/* Hyper-parameters the reference hard-codes, as runtime fields (SURVEY.md §5 "Config / flags").
 * ctr_config_default() fills the reference's values. */
typedef struct {
    int32_t model;            /* CTR_MODEL_* */
    int32_t uP;               /* uProfileDim            model.go:27 */
    int32_t S;                /* uBehaviorSize          rcmd.go:24  (10) */
    int32_t D;                /* uBehaviorDim == iFeatureDim (din.go:176-178)  rcmd.go:22 (16) */
    int32_t cF;               /* cFeatureDim */
    int32_t H0, H1;           /* din.go:17-18 (200, 80) */
    int32_t batch;            /* model.Train batchSize */
    int32_t pred_batch;       /* model.InitForwardOnlyVm batchSize (model.go:215) */
    float   lr, l2, beta1, beta2, eps;   /* model.go:88: 0.01, 1e-4; gorgonia Adam defaults .9 .999 1e-8 */
    float   dropout0, dropout1;          /* din.go:204-205 (.005) / dnn.go:136-137 (.003) */
    uint32_t seed;            /* dropout-mask / weight-init counter RNG seed */
    int32_t table_opt;        /* CTR_TABLE_* */
    float   table_lr;         /* SGD step for embedding rows (engine extension) */
    int32_t gemm;             /* CTR_GEMM_* */
    int32_t device;           /* CUDA device ordinal */
    int32_t rank, world;      /* one handle (process) per GPU; large ITEM_* tables are row-sharded (owner(row) = row % world, world in
                                 {2,4,8}: peers read / update the owner's HBM over NVLink), small ones replicated */
    int32_t reserved[8];      /* tuning knobs, 0 = automatic: [0] hot rows (ids [0, n): keep rows ordered by popularity) with replica
                                 accumulators — and, on row-sharded tables, a replica on every rank (< 0: none; default 32768);
                                 [1] ITEM_EMB / ITEM_FEAT placement under world > 1: 1 = always shard, 2 = always replicate
                                 (default: replicate tables <= 32 MB, shard larger ones) */
} ctr_config;

typedef struct {
    float   cost;             /* BCE mean of this batch (model/cost.go:9-17) */
    float   ms_device;        /* device time of the step (CUDA events on the engine stream); 0 if not measured */
    int32_t launches;         /* kernels launched by the step */
    int32_t reserved;
} ctr_step_stats;

int  ctr_abi_version(void);
void ctr_config_default(ctr_config* cfg, int model);

/* din.NewDinNet / youtube.NewYoutubeDnn (din.go:171, dnn.go:119) — allocates device state. */
int  ctr_create(const ctr_config* cfg, ctr_handle** out);
void ctr_destroy(ctr_handle* h);
/* last error message of this handle (or of the last failed ctr_create when h == NULL) */
const char* ctr_last_error(const ctr_handle* h);

/* G.WithInit(G.Gaussian(0,1)) for mlp0/1/2 and ValuesOf(1) for att0 (din.go:181-191) with the
 * engine's counter RNG (streams 0,1,2) — the reference's draws are time-seeded, never reproducible. */
int ctr_init_weights(ctr_handle* h, uint32_t seed);
/* DinNet.Marshal / NewDinNetFromJson field layout (din.go:41-52, dnn.go:38-47): mlp0 [in,H0],
 * mlp1 [H0,H1], mlp2 [H1,1], att0 [1,S] row-major; att0 may be NULL for YouTube.  set also resets
 * the Adam moments and step counter (a fresh solver is built per model.Train call, model.go:88). */
int ctr_set_weights(ctr_handle* h, const float* mlp0, const float* mlp1, const float* mlp2, const float* att0);
int ctr_get_weights(ctr_handle* h, float* mlp0, float* mlp1, float* mlp2, float* att0);

/* Feature / embedding tables resident in HBM: replaces UserFeatureCache / ItemFeatureCache /
 * itemEmbeddingMap (rcmd.go:30-36, 473-505).  rows is [nrows, width] row-major.  With world > 1,
 * ITEM_EMB upload takes the FULL table on every rank and keeps rows r % world == rank — or all of them when the
 * table is small enough to be replicated (<= 32 MB, see ctr_config.reserved[1]); download mirrors that (a sharded
 * table's download fills this rank's rows only).  ITEM_FEAT follows the same rule; USER_FEAT is always replicated.
 * Row ids the entry points receive are range-checked on the device: an id outside [0, rows) reads as a missing row. */
int ctr_table_upload(ctr_handle* h, int which, const float* rows, int64_t nrows, int32_t width);
int ctr_table_download(ctr_handle* h, int which, float* rows, int64_t nrows, int32_t width);
/* Synthetic table generated on the device with the counter RNG (benchmarks with 10M-100M rows,
 * BASELINE.json configs[2..3]): element (r, c) of the FULL table = scale * draw(seed, stream=which,
 * ctr=r*width+c), dist 0 = U[0,1), 1 = N(0,1) (Box-Muller) — identical on every rank / shard. */
int ctr_table_fill(ctr_handle* h, int which, int64_t nrows, int32_t width, uint32_t seed, int32_t dist, float scale);

/* recommend.GetSampleVector for a batch (rcmd.go:462-536): X[b] = [user_feat | emb[hist[b,0..S)] |
 * emb[item] | item_feat], bit-exact copies.  X is [B, uP + S*D + D + cF]. */
int ctr_gather_rows(ctr_handle* h, const int32_t* user_row, const int32_t* item_row,
                    const int32_t* hist_rows, int64_t B, float* X);

/* model.Train (model.go:27-213) on the dense X / Y that recommend.GetSample built
 * (rcmd.go:339-460).  ranges = SampleInfo {UserProfile, UserBehavior, ItemFeature, CtxFeature} as
 * 4 [start,end) pairs (rcmd.go:132-137).  Zero-pads the ragged last batch and trains on it with
 * label 0 (model.go:132-184,357-371); *last_cost = cost of the last batch (model.go:198);
 * early_stop = no-improvement epochs (0 = off, model.go:199-209).  Single-GPU route (the reference's own shape of
 * use); with world > 1 it returns CTR_ESTATE — multi-GPU training goes through the index entry points. */
int ctr_train_dense(ctr_handle* h, const float* X, const float* Y, int64_t n, int32_t xcols,
                    const int32_t ranges[8], int32_t epochs, int32_t early_stop,
                    float* last_cost, int32_t* epochs_run);
/* model.Predict (model.go:242-353): batches of pred_batch, zero-padded tail, dropout off
 * (din.go:133-145).  out is [n]. */
int ctr_predict_dense(ctr_handle* h, const float* X, int64_t n, int32_t xcols,
                      const int32_t ranges[8], float* out);

/* The B200-native fast path: one model.Train inner-loop iteration (model.go:107-196) fed by row
 * indices instead of a materialised X; gathers from the HBM tables inside the kernels.
 * hist_rows is [B,S], most-recent-first, -1 padded (prepare.go:49-51, rcmd.go:517-522).
 * B must equal cfg.batch.  Host buffers: copied H2D inside the call. */
int ctr_train_step_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row,
                       const int32_t* hist_rows, const float* label, int32_t B, ctr_step_stats* stats);
/* model.Train's batch loop for one pass over n samples (model.go:107-196) fed by row ids: batches of
 * cfg.batch are consumed in order, the ragged tail is zero-padded with label 0 (model.go:357-371), and
 * the host→device copy of batch i+1 overlaps the compute of batch i: the caller's (pageable) buffers are copied into
 * an internal pinned ring by a few host threads and DMA'd from there on a second stream.  One coarse call
 * per epoch is the shape a cgo caller wants.  costs (may be NULL) receives ceil(n/batch) batch costs.
 * With world > 1 every step is collective (both placements): all ranks must call with the same n. */
int ctr_train_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row,
                  const int32_t* hist_rows, const float* label, int64_t n, float* costs);
/* recommend.Train (rcmd.go:197-246) fed by sample keys {UserId, ItemId, Timestamp, Label} (rcmd.go:189-194) — the
 * whole of GetSample (rcmd.go:339-460) runs on the device: the id maps (ctr_idmap_build) resolve the keys, samples
 * whose user or item has no features are dropped as the reference's assembler skips them (rcmd.go:378-382), the
 * survivors stay in HBM in input order, and each batch's history rows are windowed from the device ubcache at the
 * sample's timestamp (GetUserBehavior(uid, S, -1, ts): rcmd.go:509, prepare.go:13-38; no ubcache uploaded → empty
 * history) right before its step.  Then model.Train's loop: batches of cfg.batch, zero-padded tail with label 0,
 * *last_cost = cost of the epoch's last batch, early_stop = no-improvement epochs (0 = off) (model.go:96-209).
 * Host traffic: 28 bytes per sample, staged through an internal pinned ring (the caller's buffers may be pageable).
 * *rows_used = samples that survived.  world > 1: collective — every rank passes its own keys; all ranks run the
 * batch count of the rank with the most surviving samples. */
int ctr_train_keys(ctr_handle* h, const int64_t* user_ids, const int64_t* item_ids, const int64_t* ts, const float* label,
                   int64_t n, int32_t epochs, int32_t early_stop, float* last_cost, int32_t* epochs_run, int64_t* rows_used);
/* recommend.BatchPredict → model.Predict (rcmd.go:277-337) fed by indices. out is [n]. */
int ctr_predict_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row,
                    const int32_t* hist_rows, int64_t n, float* out);

/* Same step with DEVICE-resident index / label buffers; asynchronous on the engine stream, no host
 * sync (stats->cost is not filled; read it with ctr_last_cost after ctr_sync).  Used when the caller
 * keeps the sample stream in HBM (bench.py's `value` leg; the device-side ubcache row f2).
 * The *_dev entry points, ctr_sync and ctr_last_cost do not take the handle's mutex: drive a handle's device-side
 * entry points from one thread (the host-buffer entry points serialise themselves). */
int ctr_train_step_idx_dev(ctr_handle* h, const int32_t* d_user_row, const int32_t* d_item_row,
                           const int32_t* d_hist_rows, const float* d_label, int32_t B);
int ctr_predict_idx_dev(ctr_handle* h, const int32_t* d_user_row, const int32_t* d_item_row,
                        const int32_t* d_hist_rows, int32_t B, float* d_out);
int ctr_last_cost(ctr_handle* h, float* cost);
int ctr_sync(ctr_handle* h);
/* the cudaStream_t the engine launches on (so a harness can record its own events on it), and a
 * way to adopt the caller's stream instead */
void* ctr_get_stream(ctr_handle* h);
int   ctr_set_stream(ctr_handle* h, void* cuda_stream);
/* number of kernels the engine has launched since creation */
int64_t ctr_launch_count(const ctr_handle* h);
/* per-kernel device-time profile: enable → every launch is bracketed by events (slow; for bench's
 * roofline leg).  ctr_profile_get returns accumulated ms and launch count for a kernel name. */
int ctr_profile_enable(ctr_handle* h, int on);
int ctr_profile_get(ctr_handle* h, const char* kernel, double* ms_total, int64_t* launches);
int ctr_profile_reset(ctr_handle* h);
int ctr_profile_dump(ctr_handle* h, char* buf, int64_t buflen);   /* "name ms launches\n" lines */

/* Test hook: forward+backward of one index batch WITHOUT any update; returns what G.Grad would
 * (model.go:56) plus the engine's row gradients.  Any output pointer may be NULL.
 * dUb [B,S,D], dIt [B,D], p [B], logit [B]. */
int ctr_debug_grads_idx(ctr_handle* h, const int32_t* user_row, const int32_t* item_row,
                        const int32_t* hist_rows, const float* label, int32_t B, int32_t training,
                        float* dmlp0, float* dmlp1, float* dmlp2, float* datt0,
                        float* dUb, float* dIt, float* p, float* logit, float* cost);

/* feature/ubcache on the device.  upload: every user's behaviour sequence in
 * time-DESCENDING order (ubcache.TimeSeq, cache.go:9-12) as one CSR: offsets [n_users+1], ts [n],
 * item_rows [n] (dense ITEM_EMB rows).  window: TimeSeq.Filter(maxTs, S) for a batch (cache.go:71-94;
 * maxTs == 0 means "from the newest"), i.e. what GetUserBehavior(uid, S, -1, sample.Timestamp) returns
 * (prepare.go:13-38), written as hist_rows [B,S], -1 padded.  The _dev variant leaves the result in
 * device memory for ctr_train_step_idx_dev / ctr_predict_idx_dev. */
int ctr_ubcache_upload(ctr_handle* h, const int64_t* offsets, const int64_t* ts, const int32_t* item_rows, int64_t n_users, int64_t n);
int ctr_ubcache_window(ctr_handle* h, const int32_t* user_row, const int64_t* max_ts, int32_t B, int32_t* hist_rows);
int ctr_ubcache_window_dev(ctr_handle* h, const int32_t* d_user_row, const int64_t* d_max_ts, int32_t B, int32_t* d_hist_rows);

/* ---- sparse ids, serving keys, checkpoint ------------------------------------
 * The reference keys its caches by the decimal string of a Go int (rcmd.go:472,483,502,519); the engine's
 * tables are dense.  ctr_idmap_build puts `ids[i] → row i` into a device hash table (which: CTR_IDMAP_*);
 * duplicate ids or INT64_MIN are rejected (CTR_EINVAL).  lookup writes the row, or -1 for an unknown id —
 * which the gather reads as a zero row, the reference's "not found → zeros" (rcmd.go:501-505,519-521). */
enum { CTR_IDMAP_USER = 0, CTR_IDMAP_ITEM = 1 };
int ctr_idmap_build(ctr_handle* h, int which, const int64_t* ids, int64_t n);
int ctr_idmap_lookup(ctr_handle* h, int which, const int64_t* ids, int64_t n, int32_t* rows);
int ctr_idmap_lookup_dev(ctr_handle* h, int which, const int64_t* d_ids, int64_t n, int32_t* d_rows);

/* recommend.BatchPredict (rcmd.go:277-337) over sample keys {UserId, ItemId, Timestamp} (rcmd.go:189-194),
 * entirely on the device: id maps → rows, ubcache window at the sample's timestamp → history rows,
 * forward → scores [n].  Needs both id maps; without an uploaded ubcache the history is empty (the
 * reference's "UserBehavior not implemented → zeros", rcmd.go:498,509).  A key whose user or item is
 * unknown scores as an all-zero X row (rcmd.go:299-307) — unless it is key 0, which fails the call with
 * CTR_ENOTFOUND as the reference returns the error (rcmd.go:300-303).  recommend.Rank (rcmd.go:248-275) =
 * this with one user id, the candidate item ids and ts = now. */
int ctr_batch_predict_keys(ctr_handle* h, const int64_t* user_ids, const int64_t* item_ids, const int64_t* ts,
                           int64_t n, float* scores);

/* Binary snapshot of everything a resumed run needs: dims, optimiser step, dense weights with their Adam
 * moments, and the three tables (this rank's shard of ITEM_EMB when world > 1 — use one path per rank).
 * The reference has no checkpoint at all (JSON weights only, din.go:62); that JSON stays
 * available through ctr_get_weights / ctr_set_weights.  load requires a handle created with the same
 * dims, rank and world; tables are (re)allocated from the file. */
int ctr_checkpoint_save(ctr_handle* h, const char* path);
int ctr_checkpoint_load(ctr_handle* h, const char* path);

/* utils.RocAuc32 (util.go:131-148 → nn/metrics/ranking.go:144): labels binarised at 0.5, tied
 * scores grouped, trapezoid. Sorted on the device. */
int ctr_roc_auc(ctr_handle* h, const float* pred, const float* y, int64_t n, double* auc);

/* ---- item2vec (BASELINE config 5) ------------------------------------------------
 * embedding.TrainEmbedding(ch, window, dim, iter) (feature/embedding/wordemb.go:9-32): SkipGram +
 * HierarchicalSoftmax word2vec over the users' item sequences; its output is the ITEM_EMB table the hot
 * path gathers from (recommend.Train, rcmd.go:207-213).  tokens are word ids in [0, vocab) in corpus
 * order — ids are assigned by first appearance, as dictionary.Add does (dictionary.go:70-81); the Go side
 * keeps the string↔id map.  emb_out [vocab, dim] float32 == GenEmbeddingMap32 (word2vec.go:298-324);
 * words rarer than min_count keep their initial vector (they are filtered from the training document only,
 * memory.go:53-62).  Stand-alone call: no handle; errors via ctr_last_error(NULL). */
typedef struct {
    int32_t dim, window, iter;            /* rcmd.go:22-26: 16, 5 ; iter 1 (rcmd.go:543) */
    int32_t min_count, max_depth;         /* options.go: 5, 100 */
    float   init_lr, min_lr, subsample;   /* 0.025, 0.025e-4, 1e-3 */
    int32_t update_lr_batch;              /* 100000 */
    uint32_t seed;
    int32_t device;
    int32_t reserved[8];
} ctr_i2v_config;
typedef struct {
    int64_t doc_len;                      /* tokens left after the MinCount filter */
    int64_t trained_positions;            /* positions that passed the subsampling trial, all iterations */
    int64_t pairs, node_visits;           /* (centre, context) pairs trained; Huffman nodes updated */
    double  algorithmic_bytes;            /* 2*dim*4 * (node_visits + pairs): row read-modify-writes */
    float   ms_device;                    /* device time of the training kernels */
    int32_t launches;
} ctr_i2v_stats;
void ctr_i2v_config_default(ctr_i2v_config* cfg);
/* host-only: the Huffman paths the trainer walks — dictionary.HuffnamTree (huffman.go:23-57) +
 * Node.GetPath(max_depth) (node.go:26-43) for word counts `count[vocab]`; path_off [vocab+1], then
 * (inner node id in creation order, child code) per step.  Runs without a GPU (used by the CPU tests). */
int  ctr_i2v_paths(const int64_t* count, int32_t vocab, int32_t max_depth, int64_t* path_off, int32_t* path_node, uint8_t* path_code, int64_t cap);
int  ctr_i2v_train(const ctr_i2v_config* cfg, const int32_t* tokens, int64_t n, int32_t vocab,
                   float* emb_out, ctr_i2v_stats* stats);
/* BASELINE configs[4] shape: the item stream sharded over `world` processes (one per GPU), every rank passes ITS shard.
 * Dictionary counts are all-reduced (NCCL) so every rank builds the identical Huffman tree; every rank trains a full
 * replica of both vector tables on its shard and the replicas are averaged (all-reduce) every sync_every centre
 * positions (0 = 4 Mi) — the reference's trainer is Hogwild over goroutines in one address space (word2vec.go:165-169);
 * across GPUs the shared memory becomes periodic model averaging.  The learning rate follows the GLOBAL position count.
 * emb_out receives the averaged table (identical on every rank; ranks that do not need it may pass NULL and save the
 * 4·vocab·dim-byte copy to the host); id = ncclUniqueId bytes from ctr_comm_unique_id on rank 0.
 * ctr_i2v_config.reserved[0] = 1 (single GPU only): sequential float64 parity mode — one warp walks the document in
 * order, reproducing the CPU restatement of the reference bit for bit (tests/test_gpu_i2v.py). */
int  ctr_i2v_train_dist(const ctr_i2v_config* cfg, const int32_t* tokens_shard, int64_t n_shard, int32_t vocab,
                        float* emb_out, ctr_i2v_stats* stats, int32_t rank, int32_t world,
                        const void* nccl_id, int32_t id_bytes, int64_t sync_every);

/* ---- model/mlp: the float64 MLP classifier of the reference's default path (BASELINE configs[0]; SURVEY.md §8a row
 * a11) — main.go:42-52 → model/mlp/mlp.go:45-65 (SimpleMlpFitWrap.Fit / SimpleMlpPredWrap.Predict: float32 samples →
 * float64 → nn.MLPClassifier) → nn/neural_network/basemlp64.go.  Float64 arithmetic on the device; packed parameter
 * vector in the reference's layout (basemlp64.go:459-463: per layer [intercepts | coefs row-major]). -------------- */
enum { CTR_MLP_RELU = 0, CTR_MLP_LOGISTIC = 1, CTR_MLP_IDENTITY = 2 };
typedef struct ctr_mlp ctr_mlp;
typedef struct {
    int32_t n_layers;                 /* len(layerUnits): hidden layers + 2 (basemlp64.go:495-497) */
    int32_t units[8];                 /* [nFeatures, HiddenLayerSizes..., 1] */
    int32_t hidden_act;               /* Activation "relu" (:229) | "logistic" | "identity"; the output is logistic (:423-425) */
    int32_t batch, max_iter, n_iter_no_change;   /* BatchSize 200, MaxIter 200, NIterNoChange 10 (:233,238,254) */
    int32_t shuffle, adaptive;        /* Shuffle true (:241); LearningRate == "adaptive" */
    int32_t warm_start;               /* WarmStart (:245): keep the current parameters instead of re-initialising */
    uint32_t seed;                    /* init + shuffle counter RNG (the reference's source is time-seeded, :448,500) */
    int32_t device;
    double alpha, lr_init, beta1, beta2, eps, tol;   /* 1e-4, 1e-3, .9, .999, 1e-8, 1e-4 (:232-252) */
} ctr_mlp_config;
void ctr_mlp_config_default(ctr_mlp_config* cfg, int32_t n_features);      /* NewBaseMultilayerPerceptron64, :227-256 */
int  ctr_mlp_create(const ctr_mlp_config* cfg, ctr_mlp** out);             /* nn.NewMLPClassifier */
void ctr_mlp_destroy(ctr_mlp* m);
const char* ctr_mlp_last_error(const ctr_mlp* m);
/* SimpleMlpFitWrap.Fit (mlp.go:45-65): X [n, xcols] float32, Y [n] float32 in {0,1} → fit (basemlp64.go:484-567 →
 * fitStochastic :729-857, solver adam).  *n_iter = epochs run (NIter); loss_curve (may be NULL) receives max_iter
 * per-epoch losses at most. */
int  ctr_mlp_fit(ctr_mlp* m, const float* X, const float* Y, int64_t n, int32_t xcols, int32_t* n_iter, double* loss_curve);
/* SimpleMlpPredWrap.Predict (mlp.go:15-39): probabilities as float32 [n] (the 64-bit class returns them raw,
 * basemlp64.go:897-931). */
int  ctr_mlp_predict(ctr_mlp* m, const float* X, int64_t n, int32_t xcols, float* out);
/* packed parameter vector; ctr_mlp_get_params(m, NULL, 0, &np) returns the length */
int  ctr_mlp_get_params(ctr_mlp* m, double* params, int64_t cap, int64_t* np);
int  ctr_mlp_set_params(ctr_mlp* m, const double* params, int64_t np);

/* Multi-GPU (world > 1): one handle per rank/process.  The id is ncclUniqueId bytes produced on
 * rank 0 by ctr_comm_unique_id and distributed by the host (torch.distributed / Go). */
int ctr_comm_unique_id(void* id_out, int32_t* id_bytes /* in: capacity, out: used */);
int ctr_comm_init(ctr_handle* h, const void* id, int32_t id_bytes);

#ifdef __cplusplus
}
#endif
#endif
