"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
hot path.  Vectors: tests/golden/kat.json (copied from the cited reference tests)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


@pytest.mark.parametrize("case", KAT["bce"])
def test_bce(case):  # model/cost_test.go:12-55
    assert abs(orc.bce32(case["pred"], case["y"]) - case["want"]) <= case["tol"]


def test_bce_eps_is_inert():  # cost.go:12: float32(1.0+1e-8) == 1 → p==1,y==0 gives +Inf cost
    assert np.float32(1.0 + 1e-8) == np.float32(1.0)
    assert orc.bce32([1.0], [0.0]) == float("inf")


@pytest.mark.parametrize("case", KAT["mse"])
def test_mse(case):
    assert abs(orc.mse32(case["pred"], case["y"]) - case["want"]) <= case["tol"]


@pytest.mark.parametrize("case", KAT["rms"])
def test_rms(case):
    assert abs(orc.rms32(case["pred"], case["y"]) - case["want"]) <= case["tol"]


@pytest.mark.parametrize("case", KAT["prelu"])
def test_prelu(case):
    np.testing.assert_array_equal(orc.prelu32(case["x"], case["slope"]), np.array(case["want"], np.float32))


@pytest.mark.parametrize("case", KAT["euc"])
def test_euc(case):  # ShouldResemble == exact
    x = np.array(case["x"], np.float32).reshape(case["x_shape"])
    y = np.array(case["y"], np.float32).reshape(case["y_shape"])
    out = orc.euc_distance(x, y)
    assert list(out.shape) == case["want_shape"]
    np.testing.assert_array_equal(out.ravel(), np.array(case["want"], np.float32))


@pytest.mark.parametrize("case", KAT["cosine"])
def test_cosine(case):
    x = np.array(case["x"], np.float32).reshape(case["x_shape"])
    y = np.array(case["y"], np.float32).reshape(case["y_shape"])
    out = orc.cosine(x, y)
    assert list(out.shape) == case["want_shape"]
    np.testing.assert_array_equal(out.ravel(), np.array(case["want"], np.float32))


@pytest.mark.parametrize("key,fn", [("euc_error", orc.euc_distance), ("cosine_error", orc.cosine)])
def test_dim_mismatch_is_an_error(key, fn):
    for case in KAT[key]:
        with pytest.raises(ValueError):
            fn(np.zeros(case["x_shape"], np.float32), np.zeros(case["y_shape"], np.float32))


@pytest.mark.parametrize("case", KAT["auc"])
def test_auc(case):
    assert orc.roc_auc(case["pred"], case["y"]) == case["want"]


def test_auc_matches_sklearn_with_ties():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    y = (rng.random(5000) > 0.6).astype(np.float32)
    s = np.round(rng.random(5000) * 0.5 + y * 0.2, 2).astype(np.float32)   # many ties
    assert abs(orc.roc_auc(s, y) - roc_auc_score(y, s)) < 1e-12


def test_ub_filter():
    k = KAT["ub_filter"]
    ts = np.array(k["ts"], np.int64)
    for c in k["cases"]:
        start, cnt = orc.ub_filter(ts, c["max_ts"], c["max_len"])
        assert ts[start:start + cnt].tolist() == c["want"]


def test_hash_onehot32():
    k = KAT["hash_onehot32"]
    for s, want in k["cases"].items():
        assert orc.hash_onehot32(s, k["size"]) == want


def test_sigmoid32_saturation():
    assert orc.sigmoid32(-89.0) == 0.0
    assert orc.sigmoid32(15.5) == 1.0
    assert abs(orc.sigmoid32(0.3) - 1 / (1 + np.exp(-0.3))) < 1e-7
