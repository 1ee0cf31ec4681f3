"""Rows f3/f4: sparse-id maps, BatchPredict/Rank over sample keys, and the binary checkpoint,
through the C ABI, against the oracle / a plain dict restatement of rcmd.go:277-337,462-536."""
import numpy as np
import pytest

import go_ctr_b200 as g
from go_ctr_b200 import serving
from oracle import oracle as orc
from tests.util import make_batch, make_tables, scaled_init

pytestmark = pytest.mark.gpu


def test_idmap_maps_sparse_ids_to_rows_bit_exact():
    rng = np.random.default_rng(0)
    n = 200_003
    ids = rng.choice(np.arange(-5_000_000, 5_000_000, dtype=np.int64), n, replace=False)
    ids[:3] = [0, -1, np.iinfo(np.int64).max]
    ids = np.unique(ids); rng.shuffle(ids); n = ids.size
    eng = g.Engine(g.engine.default_config(g.MODEL_DIN_COS, batch=8, pred_batch=8))
    eng.idmap_build(g.IDMAP_ITEM, ids)
    q = np.concatenate([ids[rng.integers(0, n, 50_000)], rng.integers(6_000_000, 7_000_000, 1000), [np.iinfo(np.int64).min]])
    want = {int(v): r for r, v in enumerate(ids)}
    got = eng.idmap_lookup(g.IDMAP_ITEM, q)
    assert got.tolist() == [want.get(int(v), -1) for v in q]
    with pytest.raises(g.CtrError):
        eng.idmap_build(g.IDMAP_USER, np.array([5, 7, 5], np.int64))          # duplicate ids
    with pytest.raises(g.CtrError):
        eng.idmap_lookup(g.IDMAP_USER, q)                                      # not built (the failed build left nothing)
    eng.idmap_build(g.IDMAP_USER, np.array([42], np.int64))
    assert eng.idmap_lookup(g.IDMAP_USER, [42, 43]).tolist() == [0, -1]


def _serving_engine(model, seed=0, U=60, I=150, uP=9, S=6, D=16, cF=7, pred_batch=64):
    rng = np.random.default_rng(seed)
    cfg = g.engine.default_config(model, uP=uP, S=S, D=D, cF=cF, batch=pred_batch, pred_batch=pred_batch)
    eng = g.Engine(cfg)
    uf, itf, emb = make_tables(rng, U, I, uP, cF, D)
    for w, t in ((g.TABLE_USER_FEAT, uf), (g.TABLE_ITEM_FEAT, itf), (g.TABLE_ITEM_EMB, emb)):
        eng.table_upload(w, t)
    ocfg = orc.make_cfg(model, uP, S, D, cF, 200, 80)
    W = scaled_init(orc, ocfg, seed + 1); eng.set_weights(*W)
    uid = rng.choice(10_000_000, U, replace=False).astype(np.int64) + 1
    iid = rng.choice(10_000_000, I, replace=False).astype(np.int64) + 1
    serving.load_id_maps(eng, uid, iid)
    lens = rng.integers(0, 15, U); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ts = np.concatenate([np.sort(rng.integers(1, 1000, n))[::-1] for n in lens] + [np.zeros(0, np.int64)]).astype(np.int64)
    items = rng.integers(0, I, ts.size).astype(np.int32)
    return eng, ocfg, W, (uf, itf, emb), (uid, iid), (off, ts, items, lens), rng


def _expected_scores(ocfg, W, tabs, ids, ub, keys, with_ub=True):
    """rcmd.go:277-337 + 462-536 restated with dicts: rows (or a zero row), history window, oracle forward."""
    uf, itf, emb = tabs; uid, iid = ids; off, ts, items, lens = ub
    umap = {int(v): r for r, v in enumerate(uid)}; imap = {int(v): r for r, v in enumerate(iid)}
    S = ocfg.S
    n = len(keys)
    ur = np.full(n, -1, np.int32); ir = np.full(n, -1, np.int32); hist = np.full((n, S), -1, np.int32)
    for k, (u, i, t) in enumerate(keys):
        r, c = umap.get(int(u), -1), imap.get(int(i), -1)
        if r < 0 or c < 0:
            continue                                         # GetSampleVector error → zero X row (rcmd.go:299-307)
        ur[k], ir[k] = r, c
        if with_ub and lens[r] > 0:
            start, cnt = orc.ub_filter(ts[off[r]:off[r + 1]], int(t), S)
            hist[k, :cnt] = items[off[r] + start: off[r] + start + cnt]
    X = orc.gather_rows(uf, itf, emb, ur, ir, hist)
    p, _ = orc.forward(ocfg, W, X, orc.make_ranges(ocfg.uP, ocfg.S, ocfg.D, ocfg.cF))
    return p, X


@pytest.mark.parametrize("model", [g.MODEL_DIN_COS, g.MODEL_YOUTUBE])
def test_batch_predict_over_sample_keys(model):
    eng, ocfg, W, tabs, ids, ub, rng = _serving_engine(model)
    uid, iid = ids
    n = 333                                                  # > pred_batch: several chunks, ragged tail
    ku = uid[rng.integers(0, uid.size, n)]; ki = iid[rng.integers(0, iid.size, n)]; kt = rng.integers(0, 1100, n).astype(np.int64)
    ku[5] = 999_999_999_999; ki[9] = -4; ku[200] = 0         # unknown user / item → zero rows (not key 0)
    keys = list(zip(ku, ki, kt))
    # without an uploaded ubcache the history is empty (UserBehavior not implemented → zeros, rcmd.go:498,509)
    want, X = _expected_scores(ocfg, W, tabs, ids, ub, keys, with_ub=False)
    got = serving.BatchPredict(eng, [serving.Sample(int(u), int(i), int(t)) for u, i, t in keys])
    assert got.shape == (n, 1) and got.dtype == np.float32
    np.testing.assert_allclose(got[:, 0], want, rtol=1e-4, atol=1e-7)
    eng.ubcache_upload(*ub[:3])
    want, X = _expected_scores(ocfg, W, tabs, ids, ub, keys)
    assert not X[5].any() and not X[9].any() and not X[200].any()
    got = eng.batch_predict_keys(ku, ki, kt)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-7)
    assert got[5] == got[9] == got[200]                      # the score of an all-zero row
    # key 0 unresolvable → the call fails like the reference (rcmd.go:300-303)
    with pytest.raises(g.CtrError) as e:
        eng.batch_predict_keys(np.array([123456789012, uid[0]]), np.array([iid[0], iid[1]]), np.array([5, 5]))
    assert e.value.code == g.ENOTFOUND and "get sample vector error" in str(e.value)


def test_rank_scores_candidates_in_caller_order():
    eng, ocfg, W, tabs, ids, ub, rng = _serving_engine(g.MODEL_DIN_COS, seed=3)
    eng.ubcache_upload(*ub[:3])
    uid, iid = ids
    cands = [int(v) for v in iid[rng.integers(0, iid.size, 20)]]
    scores = serving.Rank(eng, int(uid[7]), cands, now=500)
    assert [s.ItemId for s in scores] == cands
    want, _ = _expected_scores(ocfg, W, tabs, ids, ub, [(uid[7], c, 500) for c in cands])
    np.testing.assert_allclose([s.Score for s in scores], want, rtol=1e-4, atol=1e-7)


def test_checkpoint_resume_continues_the_same_run(tmp_path):
    """save → keep training == load into a fresh handle → same training: weights, Adam moments, step counter
    (dropout stream) and the learnt table all come back (the reference has no checkpoint).
    The state itself round-trips bit for bit; the continued steps agree to float-reduction-order noise
    (datt0 / dW accumulate with atomics)."""
    uP, S, D, cF, B, U, I = 12, 5, 16, 6, 128, 50, 90
    rng = np.random.default_rng(1)
    kw = dict(uP=uP, S=S, D=D, cF=cF, batch=B, pred_batch=B, seed=9, table_opt=g.TABLE_SGD_DETERMINISTIC, table_lr=0.05)
    tabs = make_tables(rng, U, I, uP, cF, D)
    batches = [make_batch(rng, U, I, B, S) for _ in range(5)]

    def fresh():
        e = g.Engine(g.engine.default_config(g.MODEL_DIN_COS, **kw))
        for w, t in zip((g.TABLE_USER_FEAT, g.TABLE_ITEM_FEAT, g.TABLE_ITEM_EMB), tabs):
            e.table_upload(w, t)
        e.set_weights(*scaled_init(orc, orc.make_cfg(g.MODEL_DIN_COS, uP, S, D, cF, 200, 80), 4))
        return e

    a = fresh()
    for b in batches[:3]:
        a.train_step_idx(*b)
    path = tmp_path / "snap.ctr"
    a.checkpoint_save(path)
    w_saved = a.get_weights(); t_saved = a.table_download(g.TABLE_ITEM_EMB, I, D)
    costs_a = [a.train_step_idx(*b).cost for b in batches[3:]]
    b_eng = g.Engine(g.engine.default_config(g.MODEL_DIN_COS, **kw))          # no tables, random weights
    b_eng.checkpoint_load(path)
    for x, y in zip(w_saved, b_eng.get_weights()):
        assert x.tobytes() == y.tobytes()
    assert t_saved.tobytes() == b_eng.table_download(g.TABLE_ITEM_EMB, I, D).tobytes()
    costs_b = [b_eng.train_step_idx(*b).cost for b in batches[3:]]
    np.testing.assert_allclose(costs_a, costs_b, rtol=1e-5)
    for x, y in zip(a.get_weights(), b_eng.get_weights()):
        # a step of Adam is ~lr = 0.01 per element: lost moments or a reset step counter would show as >= 1e-3
        np.testing.assert_allclose(x, y, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(a.table_download(g.TABLE_ITEM_EMB, I, D), b_eng.table_download(g.TABLE_ITEM_EMB, I, D), rtol=1e-4, atol=1e-6)
    ur, ir, hist, _ = batches[0]
    np.testing.assert_allclose(a.predict_idx(ur, ir, hist), b_eng.predict_idx(ur, ir, hist), rtol=1e-4, atol=1e-6)
    # and a handle that does NOT load the moments drifts visibly (the test can tell the difference)
    c = fresh(); c.set_weights(*w_saved); c.table_upload(g.TABLE_ITEM_EMB, t_saved)
    for b in batches[3:]:
        c.train_step_idx(*b)
    assert max(np.abs(x - y).max() for x, y in zip(a.get_weights(), c.get_weights())) > 1e-3
    # a handle with other dims refuses the file; a truncated file is an I/O error, not a crash
    other = g.Engine(g.engine.default_config(g.MODEL_DIN_COS, **{**kw, "S": S + 1}))
    with pytest.raises(g.CtrError):
        other.checkpoint_load(path)
    raw = path.read_bytes(); (tmp_path / "cut.ctr").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(g.CtrError):
        g.Engine(g.engine.default_config(g.MODEL_DIN_COS, **kw)).checkpoint_load(tmp_path / "cut.ctr")
