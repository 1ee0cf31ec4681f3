"""One ctr_i2v_train call on the bench's synthetic stream (for an ncu capture of k_i2v_skipgram_hs).
PYTHONPATH=. python tests/tools/i2v_once.py [tokens] [vocab]"""
import sys

import numpy as np

import go_ctr_b200 as g

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 1_500_000
rng = np.random.default_rng(42)
raw = (np.floor(np.exp(rng.random(n) * np.log(V))).astype(np.int64) - 1).clip(0, V - 1)
uniq, toks = np.unique(raw, return_inverse=True)
emb, st = g.i2v_train_ids(toks.astype(np.int32), int(uniq.size), dim=64, window=5, iter=1, seed=1)
print("doc_len", st.doc_len, "ms", st.ms_device, "pairs", st.pairs, "node_visits", st.node_visits)
