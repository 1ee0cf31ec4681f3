"""Round-2 parity tests through the C ABI: config-shaped AUC parity, key-fed training, pageable epoch entry,
table Adam, ragged tails on the deterministic path, out-of-range ids."""
import numpy as np
import pytest

import go_ctr_b200 as g
from oracle import oracle as orc
from tests.util import assert_mostly_close, make_batch, make_tables, scaled_init

pytestmark = pytest.mark.gpu

NS = dict(uP=52, S=50, D=64, cF=53)
OMODEL = {g.MODEL_YOUTUBE: orc.YOUTUBE, g.MODEL_DIN_COS: orc.DIN_COS, g.MODEL_DIN_EUC: orc.DIN_EUC}


def _engine(model, B, U, I, rng, **kw):
    cfg = g.engine.default_config(model, batch=B, pred_batch=kw.pop("pred_batch", B), **NS, **kw)
    eng = g.Engine(cfg)
    uf, itf, emb = make_tables(rng, U, I, NS["uP"], NS["cF"], NS["D"])
    eng.table_upload(g.TABLE_USER_FEAT, uf); eng.table_upload(g.TABLE_ITEM_FEAT, itf); eng.table_upload(g.TABLE_ITEM_EMB, emb)
    return eng, cfg, (uf, itf, emb)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("model", [g.MODEL_DIN_COS, g.MODEL_DIN_EUC, g.MODEL_YOUTUBE], ids=["din_cos", "din_euc", "youtube"])
def test_auc_parity_at_the_baseline_config_shape(model):
    """BASELINE configs[1] shape (SURVEY §8d cfg 2): D=64, S=50, uP=52, cF=53, I=27 278 items with Zipf(1.05)
    popularity, 20 % padded histories, labels ~ Bernoulli(teacher DIN score); 50 000 train / 10 000 test samples,
    batch 200, reference init N(0,1) and solver (model.go:88), dropout on.  Engine (tcgen05 3xTF32 forward/dgrad,
    1xTF32 wgrad) and oracle (f32 with double accumulation) train from the same init on the same split:
    test AUC within +-0.002 (north_star), and well above chance."""
    rng = np.random.default_rng(2024)
    U, I, B, NTR, NTE = 20000, 27278, 200, 50000, 10000
    eng, cfg, (uf, itf, emb) = _engine(model, B, U, I, rng, pred_batch=2000, seed=7)
    ur, ir, hist, _ = make_batch(rng, U, I, NTR + NTE, NS["S"], pad_frac=0.2, zipf=True)
    S, D, uP, cF = NS["S"], NS["D"], NS["uP"], NS["cF"]
    ranges = orc.make_ranges(uP, S, D, cF)
    # teacher: a DIN with moderate weights scores every sample; the label is a Bernoulli draw of its standardised logit
    # (Bayes AUC ~0.86, so a trained model lands far from chance and the +-0.002 bar is informative)
    tcfg = orc.make_cfg(orc.DIN_COS, uP, S, D, cF, 200, 80)
    Wt = scaled_init(orc, tcfg, 99, s0=0.25, s1=0.25, s2=1.0)
    zt = np.concatenate([orc.forward(tcfg, Wt, orc.gather_rows(uf, itf, emb, ur[i:i + 5000], ir[i:i + 5000], hist[i:i + 5000]), ranges)[1]
                         for i in range(0, NTR + NTE, 5000)]).astype(np.float64)
    pt = 1.0 / (1.0 + np.exp(-2.0 * (zt - np.median(zt)) / zt.std()))
    y = (rng.random(NTR + NTE) < pt).astype(np.float32)
    ocfg = orc.make_cfg(OMODEL[model], uP, S, D, cF, 200, 80, cfg.dropout0, cfg.dropout1)
    W = [w.copy() for w in orc.init_weights(ocfg, 7)]
    eng.set_weights(*W)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(seed=7), W, uf, itf, emb)
    EPOCHS = 2
    for _ in range(EPOCHS):
        costs = eng.train_idx(ur[:NTR], ir[:NTR], hist[:NTR], y[:NTR])
        for s in range(0, NTR, B):
            oc, _ = tr.step(ur[s:s + B], ir[s:s + B], hist[s:s + B], y[s:s + B])
    assert abs(costs[-1] - oc) < 0.05, (costs[-1], oc)
    sl = slice(NTR, NTR + NTE)
    p_gpu = eng.predict_idx(ur[sl], ir[sl], hist[sl])
    o0 = orc.make_cfg(OMODEL[model], uP, S, D, cF, 200, 80)
    p_cpu = np.concatenate([orc.forward(o0, tr.W, orc.gather_rows(uf, itf, emb, ur[i:i + 5000], ir[i:i + 5000], hist[i:i + 5000]), ranges)[0]
                            for i in range(NTR, NTR + NTE, 5000)])
    auc_gpu = eng.roc_auc(p_gpu, y[sl]); auc_cpu = orc.roc_auc(p_cpu, y[sl])
    assert abs(auc_gpu - orc.roc_auc(p_gpu, y[sl])) < 1e-9            # device AUC == reference-semantics AUC
    assert auc_cpu > 0.75 and abs(auc_gpu - auc_cpu) <= 0.002, (auc_gpu, auc_cpu)


def _keys_world(rng, U, I, S, n):
    """sparse external ids, per-user behaviour sequences (time-descending CSR), sample keys"""
    user_ids = rng.choice(10 * U, U, replace=False).astype(np.int64) + 1
    item_ids = rng.choice(10 * I, I, replace=False).astype(np.int64) + 1
    lens = rng.integers(0, 2 * S, U)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ts = np.concatenate([np.sort(rng.integers(1, 10_000, l))[::-1] for l in lens]).astype(np.int64) if off[-1] else np.zeros(0, np.int64)
    items = rng.integers(0, I, off[-1]).astype(np.int32)
    su = rng.integers(0, U, n); si = rng.integers(0, I, n)
    k_user = user_ids[su].copy(); k_item = item_ids[si].copy()
    k_ts = rng.integers(0, 11_000, n).astype(np.int64)
    y = (rng.random(n) > 0.5).astype(np.float32)
    return user_ids, item_ids, off, ts, items, su.astype(np.int32), si.astype(np.int32), k_user, k_item, k_ts, y


def test_train_keys_equals_train_idx_on_the_same_samples():
    """ctr_train_keys (GetSample on the device: id maps, drop-unknown, ubcache window per batch) == ctr_train_idx fed
    with the rows / histories the host computes for the surviving samples (rcmd.go:339-460, prepare.go:13-38)."""
    rng = np.random.default_rng(8)
    U, I, B, n = 400, 3000, 256, 256 * 5 + 77
    S = NS["S"]
    engs = []
    for _ in range(2):
        eng, cfg, tabs = _engine(g.MODEL_DIN_COS, B, U, I, np.random.default_rng(1), table_opt=g.TABLE_SGD_DETERMINISTIC, table_lr=0.3, seed=4)
        engs.append(eng)
    user_ids, item_ids, off, ts, items, su, si, k_user, k_item, k_ts, y = _keys_world(rng, U, I, S, n)
    k_user[::13] = 9_999_999_999          # unknown users / items: the reference's assembler skips those samples
    k_item[5::29] = -7
    keep = np.ones(n, bool); keep[::13] = False; keep[5::29] = False
    for eng in engs:
        eng.idmap_build(g.IDMAP_USER, user_ids); eng.idmap_build(g.IDMAP_ITEM, item_ids)
        eng.ubcache_upload(off, ts, items)
    ep, cost, used = engs[0].train_keys(k_user, k_item, k_ts, y, epochs=2)
    assert ep == 2 and used == int(keep.sum())
    hist = engs[1].ubcache_window(su[keep], k_ts[keep])
    for _ in range(2):
        costs = engs[1].train_idx(su[keep], si[keep], hist, y[keep])
    assert abs(cost - costs[-1]) <= 1e-6 * max(1.0, abs(cost))
    for a, b in zip(engs[0].get_weights(), engs[1].get_weights()):       # split-K weight gradients add in arbitrary order
        assert_mostly_close(a, b, 1e-3, 1e-5, 0.999, "dense weights")
    np.testing.assert_allclose(engs[0].table_download(g.TABLE_ITEM_EMB, I, NS["D"]), engs[1].table_download(g.TABLE_ITEM_EMB, I, NS["D"]), rtol=1e-4, atol=1e-6)


def test_epoch_entry_with_pageable_buffers_and_ragged_tail_on_the_deterministic_path():
    """ctr_train_idx from plain (pageable) numpy arrays through the internal pinned ring, n not a multiple of the
    batch: equals the oracle stepping the same batches with the tail zero-padded as label-0 rows (model.go:357-371);
    padded rows must not touch the table (their staging ids are stale)."""
    rng = np.random.default_rng(21)
    U, I, B = 300, 800, 128
    n = 3 * B + 50
    eng, cfg, (uf, itf, emb) = _engine(g.MODEL_DIN_COS, B, U, I, rng, table_opt=g.TABLE_SGD_DETERMINISTIC, table_lr=0.4, seed=6, dropout0=0.0, dropout1=0.0)
    ocfg = orc.make_cfg(orc.DIN_COS, NS["uP"], NS["S"], NS["D"], NS["cF"], 200, 80, 0.0, 0.0)
    W = scaled_init(orc, ocfg, 2)
    eng.set_weights(*W)
    ur, ir, hist, y = make_batch(rng, U, I, n, NS["S"], zipf=True)
    costs = eng.train_idx(ur, ir, hist, y)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(seed=6), W, uf, itf, emb)
    want = []
    for s in range(0, n, B):
        e = min(n, s + B); pad = B - (e - s)
        u2 = np.concatenate([ur[s:e], np.full(pad, -1, np.int32)]); i2 = np.concatenate([ir[s:e], np.full(pad, -1, np.int32)])
        h2 = np.concatenate([hist[s:e], np.full((pad, NS["S"]), -1, np.int32)]); y2 = np.concatenate([y[s:e], np.zeros(pad, np.float32)])
        want.append(tr.step(u2, i2, h2, y2, table_lr=0.4)[0])
    np.testing.assert_allclose(costs, np.array(want, np.float32), rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(eng.table_download(g.TABLE_ITEM_EMB, I, NS["D"]), tr.emb, rtol=2e-3, atol=2e-5)


def test_table_adam_follows_the_oracle_row_solver():
    """CTR_TABLE_ADAM: the dense solver's update applied once per distinct row with the row's summed gradient."""
    rng = np.random.default_rng(31)
    U, I, B = 200, 500, 256
    eng, cfg, (uf, itf, emb) = _engine(g.MODEL_DIN_COS, B, U, I, rng, table_opt=g.TABLE_ADAM, table_lr=0.01, seed=3, dropout0=0.0, dropout1=0.0)
    ocfg = orc.make_cfg(orc.DIN_COS, NS["uP"], NS["S"], NS["D"], NS["cF"], 200, 80, 0.0, 0.0)
    W = scaled_init(orc, ocfg, 5)
    eng.set_weights(*W)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(seed=3), W, uf, itf, emb)
    for _ in range(4):
        b = make_batch(rng, U, I, B, NS["S"], zipf=True)
        st = eng.train_step_idx(*b)
        oc, _ = tr.step(*b, table_lr=0.01, table_adam=True)
        assert abs(st.cost - oc) <= 2e-4 * max(1.0, abs(oc))
    got = eng.table_download(g.TABLE_ITEM_EMB, I, NS["D"])
    assert np.abs(got - emb).max() > 5e-3                           # Adam moved the touched rows by ~lr per step
    # an Adam step is ~lr * sign(g) at first: elements whose tiny gradient flips sign inside fp32 noise may differ
    assert_mostly_close(got, tr.emb, 2e-3, 2e-4, 0.995, "ITEM_EMB after 4 Adam steps")
    # checkpoint carries the moments: a resumed handle continues identically
    import os, tempfile
    path = os.path.join(tempfile.mkdtemp(), "adam.ckpt")
    eng.checkpoint_save(path)
    eng2, _, _ = _engine(g.MODEL_DIN_COS, B, U, I, np.random.default_rng(0), table_opt=g.TABLE_ADAM, table_lr=0.01, seed=3, dropout0=0.0, dropout1=0.0)
    eng2.checkpoint_load(path)
    b = make_batch(rng, U, I, B, NS["S"], zipf=True)
    c1 = eng.train_step_idx(*b).cost; c2 = eng2.train_step_idx(*b).cost
    assert abs(c1 - c2) <= 1e-6 * max(1.0, abs(c1))
    np.testing.assert_allclose(eng.table_download(g.TABLE_ITEM_EMB, I, NS["D"]), eng2.table_download(g.TABLE_ITEM_EMB, I, NS["D"]), rtol=1e-4, atol=1e-6)


def test_out_of_range_ids_read_as_missing_rows():
    """A stale id from the host (>= table rows, or negative) must behave exactly like -1: zero row, no update — never
    an out-of-bounds access (the reference's "not found -> zeros", rcmd.go:501-505,519-521)."""
    rng = np.random.default_rng(41)
    U, I, B = 100, 300, 128
    eng, cfg, tabs = _engine(g.MODEL_DIN_COS, B, U, I, rng, table_opt=g.TABLE_SGD, table_lr=0.2, seed=1, dropout0=0.0, dropout1=0.0)
    eng2, _, _ = _engine(g.MODEL_DIN_COS, B, U, I, np.random.default_rng(41), table_opt=g.TABLE_SGD, table_lr=0.2, seed=1, dropout0=0.0, dropout1=0.0)
    ur, ir, hist, y = make_batch(rng, U, I, B, NS["S"])
    bad_u, bad_i, bad_h = ur.copy(), ir.copy(), hist.copy()
    clean_u, clean_i, clean_h = ur.copy(), ir.copy(), hist.copy()
    bad_u[::7] = U + 5; clean_u[::7] = -1
    bad_i[::5] = 2_000_000_000; clean_i[::5] = -1
    bad_h[:, ::3] = I; clean_h[:, ::3] = -1
    bad_h[3, 1] = -12345; clean_h[3, 1] = -1
    assert eng.gather_rows(bad_u, bad_i, bad_h).tobytes() == eng2.gather_rows(clean_u, clean_i, clean_h).tobytes()
    assert eng.predict_idx(bad_u, bad_i, bad_h).tobytes() == eng2.predict_idx(clean_u, clean_i, clean_h).tobytes()
    c1 = eng.train_step_idx(bad_u, bad_i, bad_h, y).cost; c2 = eng2.train_step_idx(clean_u, clean_i, clean_h, y).cost
    assert abs(c1 - c2) <= 1e-6 * max(1.0, abs(c1))
    np.testing.assert_allclose(eng.table_download(g.TABLE_ITEM_EMB, I, NS["D"]), eng2.table_download(g.TABLE_ITEM_EMB, I, NS["D"]), rtol=1e-5, atol=1e-7)
