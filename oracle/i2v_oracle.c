/*
 * i2v_oracle.c — CPU restatement of go-ctr's item2vec trainer (feature/embedding).
 * TEST INFRASTRUCTURE ONLY (see ctr_oracle.h).  Cites relative to the reference's feature/embedding directory.
 */
#include "ctr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* sigmoid_table.go:28-45 */
static double g_lut[1000]; static int g_lut_ok = 0;
static void lut_init(void) {
    for (int i = 0; i < 1000; i++) { double e = exp(((double)i / 1000.0 * 2.0 - 1.0) * 6.0); g_lut[i] = e / (e + 1.0); }
    g_lut_ok = 1;
}
double orc_i2v_sigmoid_lut(double x) {
    if (!g_lut_ok) lut_init();
    return g_lut[(int)((x + 6.0) * (1000.0 / 6.0 / 2.0))];
}

double orc_i2v_init(uint32_t seed, long i, int dim) {
    double u = (double)(orc_mix64(seed, 100u, (uint64_t)i) >> 11) * (1.0 / 9007199254740992.0);
    return (u - 0.5) / (double)dim;
}

typedef struct { int64_t val; int32_t id; } hnode;
static int cmp_hnode(const void* a, const void* b) {
    const hnode *x = (const hnode*)a, *y = (const hnode*)b;
    if (x->val != y->val) return x->val < y->val ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);      /* sort.SliceStable keeps id order among equals */
}

void orc_i2v_huffman(const int64_t* count, int V, int literal, int32_t* parent, uint8_t* code) {
    hnode* nodes = (hnode*)malloc(sizeof(hnode) * (size_t)(2 * V));
    for (int i = 0; i < V; i++) { nodes[i].val = count[i]; nodes[i].id = i; }
    qsort(nodes, (size_t)V, sizeof(hnode), cmp_hnode);
    for (int i = 0; i < 2 * V - 1; i++) { parent[i] = -1; code[i] = 0; }
    int next_id = V;
    if (literal) {                                      /* huffman.go:35-54, array insertion */
        int len = V; hnode* arr = nodes;
        while (len > 1) {
            hnode left = arr[0], right = arr[1];
            hnode merged; merged.val = left.val + right.val; merged.id = next_id++;
            code[left.id] = 0; code[right.id] = 1; parent[left.id] = merged.id; parent[right.id] = merged.id;
            arr += 2; len -= 2;
            int lo = 0, hi = len;                       /* sort.Search: first idx with Val >= merged.Val */
            while (lo < hi) { int mid = (lo + hi) / 2; if (arr[mid].val >= merged.val) hi = mid; else lo = mid + 1; }
            memmove(arr + lo + 1, arr + lo, sizeof(hnode) * (size_t)(len - lo));   /* room exists: we consumed two slots */
            arr[lo] = merged; len++;
        }
    } else {
        /* equivalent without the O(V) shifts: merged values are produced in non-decreasing order, a new
         * merged node goes before every queued node of equal value (older merged nodes and leaves), so
         * merged nodes of equal value form a LIFO run and win ties against leaves */
        /* merged nodes are produced in non-decreasing value order; equal values form a run that behaves as
         * a stack (the newest was inserted before the older ones).  Only the last run receives pushes and
         * runs are popped from their top, so one array in creation order plus (start, top) per run suffices. */
        hnode* mq = (hnode*)malloc(sizeof(hnode) * (size_t)V);
        long* rstart = (long*)malloc(sizeof(long) * (size_t)V); long* rtop = (long*)malloc(sizeof(long) * (size_t)V);
        long nruns = 0, hr = 0, mt = 0; int lh = 0;
        for (int made = 0; made < V - 1; made++) {
            hnode pick[2];
            for (int k = 0; k < 2; k++) {
                while (hr < nruns - 1 && rtop[hr] == rstart[hr]) hr++;
                int have_m = hr < nruns && rtop[hr] > rstart[hr];
                int use_m = have_m && (lh >= V || mq[rstart[hr]].val <= nodes[lh].val);
                if (use_m) { pick[k] = mq[--rtop[hr]]; if (hr == nruns - 1) mt = rtop[hr]; }
                else pick[k] = nodes[lh++];
            }
            hnode merged; merged.val = pick[0].val + pick[1].val; merged.id = next_id++;
            code[pick[0].id] = 0; code[pick[1].id] = 1; parent[pick[0].id] = merged.id; parent[pick[1].id] = merged.id;
            if (nruns > 0 && rtop[nruns - 1] == mt && ((rtop[nruns - 1] > rstart[nruns - 1] && mq[rstart[nruns - 1]].val == merged.val) ||
                                                      (rtop[nruns - 1] == rstart[nruns - 1]))) {
                if (rtop[nruns - 1] == rstart[nruns - 1]) { rstart[nruns - 1] = mt; }   /* reuse the emptied last run */
                mq[mt++] = merged; rtop[nruns - 1] = mt;
            } else { rstart[nruns] = mt; mq[mt++] = merged; rtop[nruns] = mt; nruns++; }
        }
        free(rstart); free(rtop);
        free(mq);
    }
    free(nodes);
}

int orc_i2v_path(const int32_t* parent, const uint8_t* code, int V, int w, int max_depth, int32_t* nodes, uint8_t* codes) {
    /* node.go:26-43: cache = root..leaf (leaf included); GetPath(depth) = first min(depth, len) entries;
     * optimizer.go:113-115 walks i in [0, len(path)-1): node path[i] (inner), child code path[i+1].Code */
    int32_t chain[4096]; int len = 0;
    for (int p = w; p != -1; p = parent[p]) chain[len++] = p;      /* leaf..root */
    int depth = max_depth < len ? max_depth : len;
    int n = 0;
    for (int i = 0; i < depth - 1; i++) {
        int node = chain[len - 1 - i], child = chain[len - 2 - i];
        nodes[n] = node - V; codes[n] = code[child]; n++;
    }
    return n;
}

long orc_i2v_train(const orc_i2v_cfg* c, const int32_t* tokens, long n, int V, float* emb_out, double* syn1_out) {
    const int D = c->dim, W = c->window;
    if (!g_lut_ok) lut_init();
    /* dictionary counts over the whole stream (dictionary.go:70-81); the training doc drops words with
     * count < MinCount (memory.go:53-62, cpsutil.go:74-78) but they stay in the dictionary */
    int64_t* cnt = (int64_t*)calloc((size_t)V, sizeof(int64_t));
    for (long i = 0; i < n; i++) cnt[tokens[i]]++;
    int32_t* doc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1)); long nd = 0;
    for (long i = 0; i < n; i++) if (!(0 <= c->min_count && cnt[tokens[i]] < c->min_count)) doc[nd++] = tokens[i];
    /* param = (rand-0.5)/dim (word2vec.go:103-111); inner-node vectors zero (huffman.go:40) */
    double* syn0 = (double*)malloc(sizeof(double) * (size_t)V * D);
    for (long i = 0; i < (long)V * D; i++) syn0[i] = orc_i2v_init(c->seed, i, D);
    double* syn1 = (double*)calloc((size_t)(V > 1 ? V - 1 : 1) * D, sizeof(double));
    /* subsample.go:34-38 */
    double* z = (double*)malloc(sizeof(double) * (size_t)V);
    for (int i = 0; i < V; i++) { double v = cnt[i] > 0 ? 1.0 - sqrt(c->subsample / (double)cnt[i]) : 0.0; z[i] = v < 0 ? 0 : v; }
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * V)); uint8_t* code = (uint8_t*)malloc((size_t)(2 * V));
    orc_i2v_huffman(cnt, V, V <= 4096, parent, code);
    int32_t* pn = (int32_t*)malloc(sizeof(int32_t) * 4096); uint8_t* pc = (uint8_t*)malloc(4096);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)D);
    uint64_t lcg = 1;                                   /* modelutil.go:22 */
    long trained = 0;
    double lr = c->init_lr;                             /* w.currentlr persists across iterations */
    for (int it = 0; it < c->iter; it++) {
        long seen = 0;
        for (long pos = 0; pos < nd; pos++) {
            const int id = doc[pos];
            /* Subsampler.Trial (subsample.go:45-52): train when z > U[0,1) */
            double u = (double)(orc_mix64(c->seed, 200u + (uint32_t)it, (uint64_t)pos) >> 11) * (1.0 / 9007199254740992.0);
            if (z[id] > u) {
                int del;
                if (c->rng_mode == 0) { lcg = lcg * 25214903917ull + 11ull; del = (int)(lcg % (uint64_t)W); }
                else del = (int)(orc_mix64(c->seed, 300u + (uint32_t)it, (uint64_t)pos) % (uint64_t)W);
                int np = orc_i2v_path(parent, code, V, id, c->max_depth, pn, pc);
                for (int a = del; a < W * 2 + 1 - del; a++) {              /* model.go:59-77 */
                    if (a == W) continue;
                    long cpos = pos - W + a;
                    if (cpos < 0 || cpos >= nd) continue;
                    double* ctx = syn0 + (long)doc[cpos] * D;
                    for (int k = 0; k < D; k++) tmp[k] = 0.0;
                    for (int i = 0; i < np; i++) {                          /* optimizer.go:113-128 */
                        double* nv = syn1 + (long)pn[i] * D;
                        double inner = 0.0;
                        for (int k = 0; k < D; k++) inner += ctx[k] * nv[k];
                        if (inner <= -6.0 || inner >= 6.0) break;            /* `return`: abandons the rest of the path */
                        double g = (1.0 - (double)pc[i] - orc_i2v_sigmoid_lut(inner)) * lr;
                        for (int k = 0; k < D; k++) { tmp[k] += g * nv[k]; nv[k] += g * ctx[k]; }
                    }
                    for (int k = 0; k < D; k++) ctx[k] += tmp[k];
                }
                trained++;
            }
            /* observe (word2vec.go:223-233): every UpdateLRBatch positions */
            seen++;
            if (seen % c->update_lr_batch == 0) lr = lr < c->min_lr ? c->min_lr : c->init_lr * (1.0 - (double)seen / (double)n);
        }
    }
    for (long i = 0; i < (long)V * D; i++) emb_out[i] = (float)syn0[i];
    if (syn1_out && V > 1) memcpy(syn1_out, syn1, sizeof(double) * (size_t)(V - 1) * D);
    free(cnt); free(doc); free(syn0); free(syn1); free(z); free(parent); free(code); free(pn); free(pc); free(tmp);
    return trained;
}
