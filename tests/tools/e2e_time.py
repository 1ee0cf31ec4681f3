"""Where the host-fed epoch entry points spend their time (1 GPU): PYTHONPATH=. python tests/tools/e2e_time.py [rows] [steps]"""
import json
import sys
import time

import numpy as np

import go_ctr_b200 as g
from tests.util import make_batch


def main():
    I = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    U, uP, S, D, cF, B = 138_493, 52, 50, 64, 53, 65_536
    cfg = g.engine.default_config(g.MODEL_DIN_COS, uP=uP, S=S, D=D, cF=cF, batch=B, pred_batch=B, table_opt=g.TABLE_SGD, table_lr=0.01)
    eng = g.Engine(cfg)
    eng.table_fill(g.TABLE_USER_FEAT, U, uP, 1, 0, 1.0); eng.table_fill(g.TABLE_ITEM_FEAT, I, cF, 2, 0, 1.0)
    eng.table_fill(g.TABLE_ITEM_EMB, I, D, 3, 1, 0.125)
    rng = np.random.default_rng(0)
    n = B * steps
    ur, ir, hist, y = make_batch(rng, U, I, n, S, pad_frac=0.2)
    out = {}
    eng.train_idx(ur[:2 * B], ir[:2 * B], hist[:2 * B], y[:2 * B])
    for rep in range(2):
        t0 = time.perf_counter(); eng.train_idx(ur, ir, hist, y); out["train_idx_ms_per_step_%d" % rep] = 1e3 * (time.perf_counter() - t0) / steps
    t0 = time.perf_counter(); tmp = np.empty_like(hist); np.copyto(tmp, hist); out["numpy_copy_hist_GBs"] = hist.nbytes / (time.perf_counter() - t0) / 1e9
    uid = np.arange(U, dtype=np.int64) * 7 + 3; iid = np.arange(I, dtype=np.int64) * 5 + 11
    eng.idmap_build(g.IDMAP_USER, uid); eng.idmap_build(g.IDMAP_ITEM, iid)
    L = 2 * S
    off = np.arange(U + 1, dtype=np.int64) * L
    ts = np.tile(np.arange(L, 0, -1, dtype=np.int64) * 1000, U)
    items = rng.integers(0, I, U * L).astype(np.int32)
    eng.ubcache_upload(off, ts, items)
    su = rng.integers(0, U, n); si = rng.integers(0, I, n)
    ku = uid[su]; ki = iid[si]; kt = rng.integers(1000, (L + 1) * 1000, n).astype(np.int64)
    eng.train_keys(ku[:2 * B], ki[:2 * B], kt[:2 * B], y[:2 * B])
    for rep in range(2):
        t0 = time.perf_counter(); eng.train_keys(ku, ki, kt, y, epochs=0); out["keys_resolve_only_ms_%d" % rep] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter(); eng.train_keys(ku, ki, kt, y, epochs=1); out["keys_epoch_ms_per_step_%d" % rep] = 1e3 * (time.perf_counter() - t0) / steps
    eng.profile(True); eng.profile_reset()
    eng.train_keys(ku, ki, kt, y, epochs=1)
    out["keys_kernels_ms"] = {k: round(v / nn, 4) for k, (v, nn) in eng.profile_dump().items()}
    eng.profile(False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
