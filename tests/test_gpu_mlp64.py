"""model/mlp on the device (row a11): the float64 MLP classifier through the C ABI against oracle/mlp64_oracle.c."""
import numpy as np
import pytest

import go_ctr_b200 as g
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _data(n, nf, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((n, nf), np.float32)
    w = rng.standard_normal(nf)
    y = ((X - 0.5) @ w + 0.3 * rng.standard_normal(n) > 0).astype(np.float32)
    return X, y


def test_loss_at_zero_parameters_is_ln2():
    """multilayer_perceptron_test.go:86-91: loss(theta = 0) = ln 2 (chkLoss +-1e-3)."""
    X, y = _data(400, 30, 0)
    m = g.MLPClassifier(30, hidden=(10,), batch=400, max_iter=1, shuffle=False, lr_init=1e-12, warm_start=True)
    m.set_params(np.zeros(m.get_params().size))
    m.fit(X, y)
    assert abs(m.loss_curve[0] - np.log(2.0)) < 1e-9


@pytest.mark.parametrize("act", ["relu", "logistic"])
def test_one_adam_step_matches_the_oracle_gradient_and_update(act):
    """One full-batch step from known parameters: loss, and parameters after AdamOptimizer64.updateParams with the
    per-element beta powers (basemlp64.go:1082-1090)."""
    nf, n = 281, 200
    X, y = _data(n, nf, 1)
    ocfg = orc.mlp64_cfg(nf, hidden=(100,), activation=act, batch=n, max_iter=1, shuffle=False, seed=5)
    p0 = orc.mlp64_init(ocfg)
    po = p0.copy()
    it, curve, _ = orc.mlp64_fit(ocfg, po, X.astype(np.float64), y.astype(np.float64).reshape(-1, 1))
    m = g.MLPClassifier(nf, hidden=(100,), activation=act, batch=n, max_iter=1, shuffle=False, seed=5, warm_start=True)
    m.set_params(p0)
    m.fit(X, y)
    assert abs(m.loss_curve[0] - curve[0]) <= 1e-12 * max(1.0, abs(curve[0]))
    np.testing.assert_allclose(m.get_params(), po, rtol=1e-9, atol=1e-12)


def test_fit_follows_the_oracle_over_epochs_and_predicts_alike():
    """configs[0] shape: 281 -> 100 -> 1, relu, batch 200, shuffled minibatches, same init (counter RNG): loss curve
    and probabilities stay with the CPU restatement (float64 both sides; FMA contraction / libm ulps only)."""
    nf, n = 281, 3000
    X, y = _data(n, nf, 2)
    ocfg = orc.mlp64_cfg(nf, hidden=(100,), max_iter=5, seed=9)
    po = orc.mlp64_init(ocfg)
    it, curve, _ = orc.mlp64_fit(ocfg, po, X.astype(np.float64), y.astype(np.float64).reshape(-1, 1))
    m = g.MLPClassifier(nf, hidden=(100,), max_iter=5, seed=9)
    m.fit(X, y)
    assert m.n_iter == it
    np.testing.assert_allclose(m.loss_curve, curve, rtol=1e-7)
    pg = m.predict(X[:500]); pc = orc.mlp64_predict(ocfg, po, X[:500].astype(np.float64)).ravel().astype(np.float32)
    np.testing.assert_allclose(pg, pc, rtol=1e-5, atol=1e-7)
    assert curve[-1] < curve[0]


def test_stopping_rule_and_errors():
    """tol / n_iter_no_change stop (basemlp64.go:826-840,886-890); argument errors are codes, not aborts."""
    X, y = _data(600, 20, 3)
    ocfg = orc.mlp64_cfg(20, hidden=(8,), max_iter=60, seed=1, tol=1e-2, n_iter_no_change=2)
    po = orc.mlp64_init(ocfg)
    it, curve, _ = orc.mlp64_fit(ocfg, po, X.astype(np.float64), y.astype(np.float64).reshape(-1, 1))
    m = g.MLPClassifier(20, hidden=(8,), max_iter=60, seed=1, tol=1e-2, n_iter_no_change=2)
    m.fit(X, y)
    assert m.n_iter == it < 60
    with pytest.raises(g.CtrError, match="columns"):
        m.predict(np.zeros((3, 19), np.float32))
    with pytest.raises(g.CtrError, match="before fit"):
        g.MLPClassifier(20).predict(np.zeros((3, 20), np.float32))
