"""Latency of the serving call (ctr_batch_predict_keys = recommend.BatchPredict on the device) for small
candidate lists.  Run on the GPU box: PYTHONPATH=. python tests/tools/serve_latency.py"""
import json
import time

import numpy as np

import go_ctr_b200 as g
from go_ctr_b200 import serving


def main():
    U, I, uP, S, D, cF = 138_493, 27_278, 52, 50, 64, 53
    rng = np.random.default_rng(0)
    eng = g.Engine(g.engine.default_config(g.MODEL_DIN_COS, uP=uP, S=S, D=D, cF=cF, batch=1024, pred_batch=1024))
    eng.table_fill(g.TABLE_USER_FEAT, U, uP, 1, 0, 1.0); eng.table_fill(g.TABLE_ITEM_FEAT, I, cF, 2, 0, 1.0)
    eng.table_fill(g.TABLE_ITEM_EMB, I, D, 3, 1, 0.125)
    uid = rng.permutation(10 * U)[:U].astype(np.int64); iid = rng.permutation(10 * I)[:I].astype(np.int64)
    serving.load_id_maps(eng, uid, iid)
    lens = rng.integers(0, 200, U); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ts = np.concatenate([np.sort(rng.integers(1, 10**6, n))[::-1] for n in lens]).astype(np.int64)
    eng.ubcache_upload(off, ts, rng.integers(0, I, ts.size).astype(np.int32))
    out = {}
    for n in (1, 16, 100, 1000):
        ku = np.full(n, uid[5]); ki = iid[rng.integers(0, I, n)]; kt = np.full(n, 10**6, np.int64)
        for _ in range(20):
            eng.batch_predict_keys(ku, ki, kt)
        lat = []
        for _ in range(300):
            t0 = time.perf_counter(); eng.batch_predict_keys(ku, ki, kt); lat.append((time.perf_counter() - t0) * 1e6)
        lat = np.sort(lat)
        out[str(n)] = {"p50_us": round(float(lat[150]), 1), "p99_us": round(float(lat[297]), 1), "launches": None}
    l0 = eng.launch_count(); eng.batch_predict_keys(ku, ki, kt); out["launches_per_call"] = eng.launch_count() - l0
    print(json.dumps({"serve_latency": out}))


if __name__ == "__main__":
    main()
