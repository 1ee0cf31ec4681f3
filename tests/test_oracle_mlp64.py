"""Row a11: the float64 sklearn-clone MLP behind mlp.SimpleMlpFitWrap — oracle only.

Known answers that do not need the reference's (third-party) MicroChip data:
  * loss at θ = 0 is ln 2 (multilayer_perceptron_test.go:86-91, ±1e-3);
  * packed gradient == finite differences within 1e-4 (multilayer_perceptron_test.go:118-130);
plus an independent numpy restatement of backprop / the per-element Adam quirk (basemlp64.go:1075-1091),
and the cfg1 plumbing run on the model_test.go:64-77 label rule (scaled down for CPU time)."""
import numpy as np
import pytest

from oracle import oracle as orc


def _np_loss_grad(cfg, params, X, y):
    units = [cfg.units[i] for i in range(cfg.n_layers)]
    off, Ws, bs = 0, [], []
    for fi, fo in zip(units[:-1], units[1:]):
        bs.append(params[off:off + fo]); off += fo
        Ws.append(params[off:off + fi * fo].reshape(fi, fo)); off += fi * fo
    acts = [X]
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = acts[-1] @ W + b
        if l < len(Ws) - 1:
            z = np.maximum(z, 0) if cfg.hidden_act == 0 else 1 / (1 + np.exp(-z))
        else:
            z = 1 / (1 + np.exp(-z))
        acts.append(z)
    n = X.shape[0]
    h = np.clip(acts[-1], np.nextafter(0, 1), np.nextafter(1, 0))
    loss = (-(y * np.log(h)) - (1 - y) * np.log1p(-h)).sum() / n + 0.5 * cfg.alpha * sum((W * W).sum() for W in Ws) / n
    delta = acts[-1] - y
    g = np.empty_like(params)
    offs = np.cumsum([0] + [(1 + fi) * fo for fi, fo in zip(units[:-1], units[1:])])
    for l in range(len(Ws) - 1, -1, -1):
        fi, fo = units[l], units[l + 1]
        g[offs[l]:offs[l] + fo] = delta.mean(0)
        g[offs[l] + fo:offs[l + 1]] = ((acts[l].T @ delta + cfg.alpha * Ws[l]) / n).ravel()
        if l:
            delta = delta @ Ws[l].T
            delta = delta * (acts[l] != 0) if cfg.hidden_act == 0 else delta * acts[l] * (1 - acts[l])
    return loss, g


def _data(rng, n, f):
    X = rng.random((n, f))
    y = (X[:, :3].sum(1) + 0.3 * rng.standard_normal(n) > 1.5).astype(np.float64)[:, None]
    return X, y


def test_zero_theta_loss_is_ln2():
    rng = np.random.default_rng(0)
    X, y = _data(rng, 118, 27)
    cfg = orc.mlp64_cfg(27, hidden=(), activation="logistic", alpha=1.0)
    loss, g = orc.mlp64_loss_grad(cfg, np.zeros(orc.mlp64_nparams(cfg)), X, y)
    assert abs(loss - 0.693) < 1e-3 and abs(loss - np.log(2)) < 1e-15
    # at θ=0 the gradient is mean((0.5 - y)·[1, x]) — independent of alpha
    np.testing.assert_allclose(g[0], (0.5 - y).mean(), rtol=1e-13)
    np.testing.assert_allclose(g[1:], ((0.5 - y) * X).mean(0), rtol=1e-12)


@pytest.mark.parametrize("act,hidden", [("relu", (16,)), ("logistic", (12, 7)), ("relu", ())])
def test_gradient_matches_numpy_and_finite_differences(act, hidden):
    rng = np.random.default_rng(1)
    X, y = _data(rng, 64, 20)
    cfg = orc.mlp64_cfg(20, hidden=hidden, activation=act, alpha=0.7, seed=3)
    p = orc.mlp64_init(cfg)
    assert (p >= 0).all()                       # basemlp64.go:472-475: U[0,1)·bound, never negative
    p = p - 0.4 * p.mean() * (rng.random(p.size) < 0.5)   # mix signs so some relu units are off
    loss, g = orc.mlp64_loss_grad(cfg, p, X, y)
    l2, g2 = _np_loss_grad(cfg, p, X, y)
    assert abs(loss - l2) < 1e-12
    np.testing.assert_allclose(g, g2, rtol=1e-10, atol=1e-13)
    fd = np.empty_like(p); h = 1e-6
    for i in range(p.size):
        q = p.copy(); q[i] += h; lp, _ = orc.mlp64_loss_grad(cfg, q, X, y)
        q[i] -= 2 * h; lm, _ = orc.mlp64_loss_grad(cfg, q, X, y)
        fd[i] = (lp - lm) / (2 * h)
    assert np.abs(fd - g).max() < 1e-4          # the reference's own tolerance


def test_init_bounds():
    cfg = orc.mlp64_cfg(281, hidden=(100,), seed=11)
    p = orc.mlp64_init(cfg)
    assert p.size == 101 * 281 + 100 + 1 * 100 + 1 + 0 or p.size == (1 + 281) * 100 + (1 + 100) * 1
    b0 = np.sqrt(6 / 381); b1 = np.sqrt(6 / 101)
    first = p[:(1 + 281) * 100]; second = p[(1 + 281) * 100:]
    assert 0 <= first.min() and first.max() < b0 and first.max() > 0.99 * b0
    assert 0 <= second.min() and second.max() < b1
    assert abs(first.mean() - b0 / 2) < 0.01 * b0


def test_adam_advances_beta_powers_per_element():
    """basemlp64.go:1082-1087: after k elements the correction uses β^k, so element 0 of step 1 moves by
    lr·sqrt(1-β2)/(1-β1)·m/(sqrt(v)+eps) and far elements by ≈ lr·m/(sqrt(v)+eps)."""
    import ctypes as C
    cfg = orc.mlp64_cfg(30, hidden=(40,))
    n = orc.mlp64_nparams(cfg)
    rng = np.random.default_rng(5)
    p0 = rng.standard_normal(n); g = rng.standard_normal(n)
    p = p0.copy(); ms = np.zeros(n); vs = np.zeros(n)

    class St(C.Structure):
        _fields_ = [("ms", C.POINTER(C.c_double)), ("vs", C.POINTER(C.c_double)), ("beta1t", C.c_double),
                    ("beta2t", C.c_double), ("t", C.c_double), ("lr_init", C.c_double), ("lr", C.c_double)]
    dp = C.POINTER(C.c_double)
    st = St(ms.ctypes.data_as(dp), vs.ctypes.data_as(dp), 0, 0, 0, 1e-3, 1e-3)
    L = orc.lib()
    for step in range(2):
        L.orc_mlp64_adam(C.byref(cfg), C.byref(st), p.ctypes.data_as(dp), g.ctypes.data_as(dp), C.c_long(n))
    # numpy restatement
    q = p0.copy(); m = np.zeros(n); v = np.zeros(n); k = 0
    for step in range(2):
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        e = np.arange(k + 1, k + n + 1, dtype=np.float64); k += n
        lr = 1e-3 * np.sqrt(1 - 0.999 ** e) / (1 - 0.9 ** e)
        q += -lr * m / (np.sqrt(v) + 1e-8)
    np.testing.assert_allclose(p, q, rtol=1e-9, atol=1e-12)
    assert st.t == 2 and abs(st.beta2t - 0.999 ** (2 * n)) < 1e-9


def _cfg1_samples(rng, n, uP=52, S=10, D=16, cF=53):
    """model_test.go:44-77: random profile/ctx/2nd-behaviour/item columns, label = round(0.6·(d1+d2))."""
    W = uP + S * D + D + cF
    X = np.zeros((n, W), np.float32)
    ub0, it0, cx0 = uP, uP + S * D, uP + S * D + D
    X[:, :uP] = rng.random((n, uP), dtype=np.float32)
    X[:, cx0:cx0 + cF] = rng.random((n, cF), dtype=np.float32)
    X[:, ub0 + D:ub0 + 2 * D] = rng.random((n, D), dtype=np.float32)
    X[:, it0:it0 + D] = rng.random((n, D), dtype=np.float32)
    # distances over the reference test's widths (uProfileDim=5, uBehaviorDim=7, model_test.go:23-27): over all
    # 52/16 columns the rule degenerates to label 0 for every sample
    ku, kd = 5, 7
    d1 = np.abs(X[:, :ku] - X[:, cx0:cx0 + ku]).sum(1) / ku
    d2 = np.abs(X[:, ub0 + D:ub0 + D + kd] - X[:, it0:it0 + kd]).sum(1) / kd
    y = np.round((d1 + d2) * 0.6).astype(np.float32)
    return X, y


def test_cfg1_plumbing_fit_predict_auc():
    """feature_test.go:30-43 shape of use: NewMLPClassifier([100], relu, adam, 1e-5), adaptive lr .0025 →
    SimpleMlpFitWrap.Fit → Predict (float32 probabilities) → AUC > 0.5 (model_test.go:112)."""
    rng = np.random.default_rng(42)
    X, y = _cfg1_samples(rng, 4000)
    assert X.shape[1] == 281 and 0.1 < y.mean() < 0.9
    cfg = orc.mlp64_cfg(281, hidden=(100,), activation="relu", alpha=1e-5, lr_init=.0025, adaptive=True, max_iter=12, seed=1)
    pred = orc.SimpleMlpFitWrap(cfg).Fit(X, y)
    assert len(pred.loss_curve) == 12 and pred.loss_curve[-1] < pred.loss_curve[0]
    Xt, yt = _cfg1_samples(rng, 1000)
    p = pred.Predict(Xt)
    assert p.dtype == np.float32 and p.shape == (1000, 1) and (p >= 0).all() and (p <= 1).all()
    assert orc.roc_auc(pred.Predict(X)[:, 0], y) > 0.5         # the reference's own bar (model_test.go:112)
    # dense U[0,1) inputs + the non-negative init (basemlp64.go:472-475) saturate the net; one-hot-like inputs
    # (what rcmd.GetSampleVector really produces for MovieLens) are learnt quickly
    def sparse(n):
        Xs = (rng.random((n, 281)) < 0.08).astype(np.float32)
        return Xs, ((Xs @ np.linspace(-1, 1, 281)) + 0.3 * rng.standard_normal(n) > 0).astype(np.float32)
    Xs, ys = sparse(4000); Xst, yst = sparse(1000)
    ps = orc.SimpleMlpFitWrap(cfg).Fit(Xs, ys)
    assert (np.diff(ps.loss_curve) < 0).all() and orc.roc_auc(ps.Predict(Xst)[:, 0], yst) > 0.95
    # shuffle is a permutation of visits, not of the caller's data (restored at :855)
    X2 = X.copy(); orc.SimpleMlpFitWrap(cfg).Fit(X2, y); assert (X2 == X).all()


def test_no_improvement_stop_and_adaptive_lr():
    rng = np.random.default_rng(2)
    X = rng.random((300, 8)); y = (rng.random((300, 1)) < 0.5).astype(np.float64)   # unlearnable labels
    cfg = orc.mlp64_cfg(8, hidden=(), activation="logistic", batch=300, max_iter=400, lr_init=1e-9, tol=1e-4,
                        n_iter_no_change=10, shuffle=False)
    p = orc.mlp64_init(cfg)
    it, curve, lr = orc.mlp64_fit(cfg, p, X, y)
    assert it == 12            # epoch 1 sets best; 11 more non-improving epochs make the count exceed 10 (:826)
    cfg.adaptive = 1
    p = orc.mlp64_init(cfg)
    it, curve, lr = orc.mlp64_fit(cfg, p, X, y)
    assert it == 12 and lr == 1e-9   # effective lr 1e-9 <= 1e-6 → "Learning rate too small. Stopping." (:1059)
    cfg.lr_init = 1e-5
    p = orc.mlp64_init(cfg)
    it, curve, lr = orc.mlp64_fit(cfg, p, X, y)
    assert it > 12 and lr < 1e-5 and abs(np.log(lr / 1e-5) / np.log(0.8) - round(np.log(lr / 1e-5) / np.log(0.8))) < 1e-9
