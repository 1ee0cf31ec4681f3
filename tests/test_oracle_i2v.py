"""item2vec oracle (BASELINE config 5): Huffman tree = dictionary.HuffnamTree's procedure, paths =
Node.GetPath, trainer = skipGram.trainOne + hierarchicalSoftmax.optim.  The reference pins nothing numeric
here (wordemb_test.go:23 checks only shapes), so: structural invariants + a planted-cluster quality check."""
import numpy as np

from oracle import oracle as orc


def brute_huffman(count):
    """The reference's array procedure written naively in Python (huffman.go:23-57)."""
    V = len(count)
    nodes = sorted(range(V), key=lambda i: (count[i], i))          # SliceStable by Val
    val = list(count) + [0] * (V - 1)
    parent = [-1] * (2 * V - 1); code = [0] * (2 * V - 1)
    nxt = V
    while len(nodes) > 1:
        l, r = nodes[0], nodes[1]
        val[nxt] = val[l] + val[r]
        code[l], code[r] = 0, 1; parent[l] = parent[r] = nxt
        nodes = nodes[2:]
        idx = next((i for i, x in enumerate(nodes) if val[x] >= val[nxt]), len(nodes))
        nodes.insert(idx, nxt); nxt += 1
    return np.array(parent, np.int32), np.array(code, np.uint8)


def test_huffman_literal_and_fast_agree_with_the_reference_procedure():
    rng = np.random.default_rng(0)
    for V in (2, 3, 7, 64, 501):
        for hi in (3, 50, 100000):                                  # many ties … few ties
            cnt = rng.integers(1, hi, V)
            want_p, want_c = brute_huffman(cnt.tolist())
            for literal in (True, False):
                p, c = orc.i2v_huffman(cnt, literal)
                assert p.tolist() == want_p.tolist() and c.tolist() == want_c.tolist(), (V, hi, literal)


def test_paths_are_root_to_parent_with_child_codes_and_depth_cap():
    rng = np.random.default_rng(1)
    V = 200
    cnt = rng.integers(1, 1000, V)
    parent, code = orc.i2v_huffman(cnt)
    root = 2 * V - 2
    for w in (0, 17, 199):
        nodes, codes = orc.i2v_path(parent, code, V, w)
        chain = []; p = w
        while p != -1:
            chain.append(p); p = parent[p]
        chain = chain[::-1]                                         # root … leaf
        assert chain[0] == root and len(nodes) == len(chain) - 1
        assert (nodes + V).tolist() == chain[:-1] and codes.tolist() == [int(code[x]) for x in chain[1:]]
        n3, _ = orc.i2v_path(parent, code, V, w, max_depth=3)       # GetPath(3) → 2 steps
        assert len(n3) == min(2, len(nodes))
    # frequent words sit nearer the root
    depth = [len(orc.i2v_path(parent, code, V, w)[0]) for w in range(V)]
    assert depth[int(np.argmax(cnt))] <= depth[int(np.argmin(cnt))]


def test_sigmoid_table_and_init_range():
    assert abs(orc.lib().orc_i2v_sigmoid_lut(0.0) - 0.5) < 4e-3     # 1000 buckets over [-6,6)
    assert 0.0 < orc.lib().orc_i2v_sigmoid_lut(-5.99) < 0.01 and 0.99 < orc.lib().orc_i2v_sigmoid_lut(5.99) < 1.0
    v = [orc.lib().orc_i2v_init(3, i, 16) for i in range(1000)]
    assert -0.5 / 16 <= min(v) and max(v) < 0.5 / 16 and abs(np.mean(v)) < 0.003


def planted_corpus(rng, n_clusters=8, per=12, n_users=600, seq=40):
    """Users stay inside one cluster of items → items of a cluster co-occur."""
    toks = []
    for u in range(n_users):
        c = rng.integers(0, n_clusters)
        toks.extend((c * per + rng.integers(0, per, seq)).tolist())
    toks = np.array(toks, np.int32)
    # dictionary ids by first appearance (dictionary.go:70-81)
    _, first = np.unique(toks, return_index=True)
    order = toks[np.sort(first)]
    remap = np.empty(n_clusters * per, np.int32); remap[order] = np.arange(order.size)
    return remap[toks], remap, n_clusters, per


def neighbour_purity(emb, remap, n_clusters, per):
    e = emb / (np.linalg.norm(emb, axis=1, keepdims=True) + 1e-12)
    cluster_of = np.empty(n_clusters * per, np.int64); cluster_of[remap] = np.arange(n_clusters * per) // per
    sim = e @ e.T; np.fill_diagonal(sim, -2)
    nn = np.argsort(-sim, axis=1)[:, :5]
    return float(np.mean(cluster_of[nn] == cluster_of[:, None]))


def test_trainer_learns_planted_clusters_and_modes_agree_in_quality():
    rng = np.random.default_rng(2)
    toks, remap, nc, per = planted_corpus(rng)
    V = nc * per
    res = {}
    for mode in (0, 1):
        cfg = orc.i2v_cfg(dim=16, window=5, iters=3, seed=5, rng_mode=mode)
        emb, syn1, trained = orc.i2v_train(cfg, toks, V, want_syn1=True)
        assert emb.shape == (V, 16) and np.isfinite(emb).all() and np.abs(syn1).max() > 0
        assert 0.8 * toks.size * 3 < trained <= toks.size * 3       # z ≈ 1 - sqrt(1e-3/count): almost every position trains
        res[mode] = neighbour_purity(emb, remap, nc, per)
    assert res[0] > 0.9 and res[1] > 0.9, res                        # chance = 1/8
