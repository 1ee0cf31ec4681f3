"""Host logic of the serving mirror (go-ctr_b200/serving.py), no GPU: recommend.Rank / BatchPredict
(rcmd.go:248-337) build the key arrays the C ABI expects and keep the caller's candidate order."""
import numpy as np
import pytest

from go_ctr_b200 import serving


class FakeEngine:
    def __init__(self):
        self.calls = []

    def batch_predict_keys(self, u, i, t):
        self.calls.append((u.copy(), i.copy(), t.copy()))
        assert u.dtype == i.dtype == t.dtype == np.int64
        return (i % 7).astype(np.float32) / 7.0

    def idmap_build(self, which, ids):
        self.calls.append(("idmap", which, np.asarray(ids).tolist()))


def test_batch_predict_packs_keys_and_returns_column():
    e = FakeEngine()
    keys = [serving.Sample(5, 70, 111), serving.Sample(6, 71, 0), serving.Sample(5, 9, 333, Label=1.0)]
    y = serving.BatchPredict(e, keys)
    u, i, t = e.calls[0]
    assert u.tolist() == [5, 6, 5] and i.tolist() == [70, 71, 9] and t.tolist() == [111, 0, 333]
    assert y.shape == (3, 1) and y.dtype == np.float32                    # tensor.Shape{len(sampleKeys), 1}
    with pytest.raises(ValueError):
        serving.BatchPredict(e, [])


def test_rank_keeps_candidate_order_and_stamps_now(monkeypatch):
    e = FakeEngine()
    monkeypatch.setattr(serving.time, "time", lambda: 1700000000.9)
    out = serving.Rank(e, 42, [30, 10, 20])
    u, i, t = e.calls[0]
    assert u.tolist() == [42, 42, 42] and i.tolist() == [30, 10, 20] and t.tolist() == [1700000000] * 3   # time.Now().Unix()
    assert [s.ItemId for s in out] == [30, 10, 20]
    assert [s.Score for s in out] == pytest.approx([(30 % 7) / 7, (10 % 7) / 7, (20 % 7) / 7])
    out = serving.Rank(e, 42, [1], now=5)
    assert e.calls[1][2].tolist() == [5]


def test_load_id_maps_builds_both_maps():
    e = FakeEngine()
    serving.load_id_maps(e, [11, 12], [7, 8, 9])
    from go_ctr_b200 import engine as en
    assert e.calls == [("idmap", en.IDMAP_USER, [11, 12]), ("idmap", en.IDMAP_ITEM, [7, 8, 9])]
