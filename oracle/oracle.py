"""ctypes binding of the CPU oracle (oracle/ctr_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/ctr_oracle.h.  Imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc.so")

YOUTUBE, DIN_COS, DIN_EUC = 0, 1, 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ctr_oracle.c", "i2v_oracle.c", "mlp64_oracle.c", "cpu_fast.c", "ctr_oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class Cfg(C.Structure):
    _fields_ = [("model", C.c_int), ("uP", C.c_int), ("S", C.c_int), ("D", C.c_int), ("cF", C.c_int),
                ("H0", C.c_int), ("H1", C.c_int), ("d0", C.c_float), ("d1", C.c_float)]


class Ranges(C.Structure):
    _fields_ = [("up", C.c_int * 2), ("ub", C.c_int * 2), ("it", C.c_int * 2), ("cx", C.c_int * 2)]


class Solver(C.Structure):
    _fields_ = [("lr", C.c_float), ("l2", C.c_float), ("b1", C.c_float), ("b2", C.c_float),
                ("eps", C.c_float), ("seed", C.c_uint32)]


class AdamState(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_float)) for n in ("m0", "v0", "m1", "v1", "m2", "v2", "ma", "va")] + [("t", C.c_int)]


_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_bce32.restype = C.c_float
        L.orc_mse32.restype = C.c_float
        L.orc_rms32.restype = C.c_float
        L.orc_sigmoid32.restype = C.c_float
        L.orc_sigmoid32.argtypes = [C.c_float]
        L.orc_roc_auc.restype = C.c_double
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_uniform24.restype = C.c_float
        L.orc_uniform24.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_gaussian_init.argtypes = [_fp, C.c_long, C.c_uint32, C.c_uint32]
        L.orc_prelu32.argtypes = [_fp, C.c_float, C.c_int, _fp]
        L.orc_ws_new.restype = C.c_void_p
        L.orc_ws_free.argtypes = [C.c_void_p]
        L.orc_backward.restype = C.c_float
        L.orc_train_step_idx.restype = C.c_float
        L.orc_fast_train_step_idx.restype = C.c_float
        L.orc_hash_onehot32.argtypes = [C.c_char_p, C.c_int]
        L.orc_i2v_train.restype = C.c_long
        L.orc_i2v_sigmoid_lut.restype = C.c_double
        L.orc_i2v_sigmoid_lut.argtypes = [C.c_double]
        L.orc_i2v_init.restype = C.c_double
        L.orc_i2v_init.argtypes = [C.c_uint32, C.c_long, C.c_int]
        L.orc_mlp64_nparams.restype = C.c_long
        L.orc_mlp64_loss_grad.restype = C.c_double
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def make_cfg(model, uP, S, D, cF, H0=200, H1=80, d0=0.0, d1=0.0):
    return Cfg(model, uP, S, D, cF, H0, H1, d0, d1)


def make_ranges(uP, S, D, cF):
    """SampleInfo as GetSample lays it out (rcmd.go:403-421)."""
    r = Ranges()
    r.up[:] = [0, uP]
    r.ub[:] = [uP, uP + S * D]
    r.it[:] = [uP + S * D, uP + S * D + D]
    r.cx[:] = [uP + S * D + D, uP + S * D + D + cF]
    return r


def default_solver(seed=0):
    return Solver(0.01, 1e-4, 0.9, 0.999, 1e-8, seed)      # model.go:88


# ---- KAT subjects -------------------------------------------------------------------------------
def bce32(pred, y):
    p, pp = _f(pred); t, tp = _f(y)
    return float(lib().orc_bce32(pp, tp, C.c_int(p.size)))


def mse32(pred, y):
    p, pp = _f(pred); t, tp = _f(y)
    return float(lib().orc_mse32(pp, tp, C.c_int(p.size)))


def rms32(pred, y):
    p, pp = _f(pred); t, tp = _f(y)
    return float(lib().orc_rms32(pp, tp, C.c_int(p.size)))


def prelu32(x, slope):
    a, ap = _f(x); out = np.empty_like(a)
    lib().orc_prelu32(ap, C.c_float(slope), C.c_int(a.size), out.ctypes.data_as(_fp))
    return out


def _pair(fn, x, y):
    x = np.asarray(x, np.float32); y = np.asarray(y, np.float32)
    if x.ndim != y.ndim:
        raise ValueError("x, y shapes not supported: %s, %s" % (x.shape, y.shape))   # activation.go:45-47,58-61
    if x.ndim == 2:
        x3, y3 = x[:, None, :], y[:, None, :]
    else:
        x3, y3 = x, y
    n, sx, d = x3.shape; sy = y3.shape[1]
    xa, xp = _f(x3); ya, yp = _f(y3)
    out = np.empty((n, max(sx, sy)), np.float32)
    rc = fn(xp, C.c_int(sx), yp, C.c_int(sy), C.c_int(n), C.c_int(d), out.ctypes.data_as(_fp))
    if rc != 0:
        raise ValueError("x, y shapes not supported")
    return out[:, 0] if x.ndim == 2 else out


def euc_distance(x, y):
    return _pair(lib().orc_euc_distance, x, y)


def cosine(x, y):
    return _pair(lib().orc_cosine, x, y)


def sigmoid32(x):
    return float(lib().orc_sigmoid32(C.c_float(x)))


def roc_auc(pred, y):
    p, pp = _f(pred); t, tp = _f(y)
    return float(lib().orc_roc_auc(pp, tp, C.c_int(p.size)))


def ub_filter(ts, max_ts, max_len):
    ts = np.ascontiguousarray(ts, np.int64); start = C.c_int(0)
    cnt = lib().orc_ub_filter(ts.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(ts.size),
                              C.c_int64(max_ts), C.c_int64(max_len), C.byref(start))
    return start.value, cnt


def hash_onehot32(s, size):
    return int(lib().orc_hash_onehot32(s.encode(), size))


def uniform24(seed, stream, ctr):
    return float(lib().orc_uniform24(seed, stream, ctr))


def gaussian_init(n, seed, stream):
    w = np.empty(n, np.float32)
    lib().orc_gaussian_init(w.ctypes.data_as(_fp), C.c_long(n), C.c_uint32(seed), C.c_uint32(stream))
    return w


def init_weights(cfg, seed):
    """N(0,1) for mlp0/1/2 (din.go:187-191), ones for att0 (din.go:181); streams 0,1,2."""
    inn = cfg.uP + 2 * cfg.D + cfg.cF
    return (gaussian_init(inn * cfg.H0, seed, 0).reshape(inn, cfg.H0),
            gaussian_init(cfg.H0 * cfg.H1, seed, 1).reshape(cfg.H0, cfg.H1),
            gaussian_init(cfg.H1, seed, 2).reshape(cfg.H1, 1),
            np.ones(cfg.S, np.float32))


# ---- hot path -----------------------------------------------------------------------------------
def gather_rows(user_feat, item_feat, item_emb, user_row, item_row, hist):
    uf, ufp = _f(user_feat); itf, itfp = _f(item_feat); em, emp = _f(item_emb)
    ur, urp = _i(user_row); ir, irp = _i(item_row); hs, hsp = _i(hist)
    B, S = hs.shape; uP = uf.shape[1]; cF = itf.shape[1]; D = em.shape[1]
    X = np.empty((B, uP + S * D + D + cF), np.float32)
    lib().orc_gather_rows(ufp, C.c_long(uP), itfp, C.c_long(cF), emp, C.c_long(D), uP, cF, S, D,
                          urp, irp, hsp, C.c_long(B), X.ctypes.data_as(_fp))
    return X


class Workspace:
    def __init__(self, cfg, B):
        self.cfg, self.B = cfg, B
        self.h = lib().orc_ws_new(C.byref(cfg), C.c_int(B))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ws_free(C.c_void_p(self.h)); self.h = None


def forward(cfg, W, X, ranges, nvalid=None, training=False, seed=0, step=0, ws=None):
    """W = (W0, W1, W2, att).  Returns (p[B], logit[B])."""
    Xa, Xp = _f(X); B = Xa.shape[0]
    nvalid = B if nvalid is None else nvalid
    w = [_f(a) for a in W]
    p = np.empty(B, np.float32); z = np.empty(B, np.float32)
    lib().orc_forward(C.byref(cfg), w[0][1], w[1][1], w[2][1], w[3][1], Xp, C.c_long(Xa.shape[1]),
                      C.byref(ranges), C.c_int(B), C.c_int(nvalid), C.c_int(int(training)),
                      C.c_uint32(seed), C.c_uint32(step), C.c_void_p(ws.h) if ws else None,
                      p.ctypes.data_as(_fp), z.ctypes.data_as(_fp))
    return p, z


def backward(cfg, W, ws, y, want_rows=True):
    """Returns dict(cost, dW0, dW1, dW2, datt, dUb, dIt)."""
    B = ws.B; inn = cfg.uP + 2 * cfg.D + cfg.cF
    w = [_f(a) for a in W]; ya, yp = _f(y)
    g0 = np.zeros((inn, cfg.H0), np.float32); g1 = np.zeros((cfg.H0, cfg.H1), np.float32)
    g2 = np.zeros((cfg.H1, 1), np.float32); ga = np.zeros(cfg.S, np.float32)
    dUb = np.zeros((B, cfg.S, cfg.D), np.float32) if want_rows else None
    dIt = np.zeros((B, cfg.D), np.float32) if want_rows else None
    cost = lib().orc_backward(C.byref(cfg), w[0][1], w[1][1], w[2][1], w[3][1], C.c_void_p(ws.h), yp,
                              C.c_int(B), g0.ctypes.data_as(_fp), g1.ctypes.data_as(_fp),
                              g2.ctypes.data_as(_fp), ga.ctypes.data_as(_fp),
                              dUb.ctypes.data_as(_fp) if want_rows else None,
                              dIt.ctypes.data_as(_fp) if want_rows else None)
    return dict(cost=float(cost), dW0=g0, dW1=g1, dW2=g2, datt=ga, dUb=dUb, dIt=dIt)


def adam_step(w, g, m, v, t, lr=0.01, l2=1e-4, batch=1.0, b1=0.9, b2=0.999, eps=1e-8):
    """In place on float32 contiguous arrays."""
    for a in (w, g, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().orc_adam_step(w.ctypes.data_as(_fp), g.ctypes.data_as(_fp), m.ctypes.data_as(_fp),
                        v.ctypes.data_as(_fp), C.c_long(w.size), C.c_int(t), C.c_float(lr), C.c_float(l2),
                        C.c_float(batch), C.c_float(b1), C.c_float(b2), C.c_float(eps))


def train_dense(cfg, solver, W, X, Y, ranges, batch, epochs, early_stop=0, nthreads=0):
    """model.Train. W arrays are updated in place (must be float32 contiguous). Returns (epochs_run, last_cost)."""
    Xa, Xp = _f(X); Ya, Yp = _f(Y)
    for a in W:
        assert a.dtype == np.float32 and a.flags.c_contiguous
    cost = C.c_float(0)
    ep = lib().orc_train_dense(C.byref(cfg), C.byref(solver), *[a.ctypes.data_as(_fp) for a in W],
                               Xp, Yp, C.c_long(Xa.shape[0]), C.c_int(Xa.shape[1]), C.byref(ranges),
                               C.c_int(batch), C.c_int(epochs), C.c_int(early_stop), C.byref(cost),
                               C.c_int(nthreads))
    return ep, cost.value


def predict_dense(cfg, W, X, ranges, batch):
    Xa, Xp = _f(X); w = [_f(a) for a in W]
    out = np.empty(Xa.shape[0], np.float32)
    lib().orc_predict_dense(C.byref(cfg), w[0][1], w[1][1], w[2][1], w[3][1], Xp, C.c_long(Xa.shape[0]),
                            C.c_int(Xa.shape[1]), C.byref(ranges), C.c_int(batch), out.ctypes.data_as(_fp))
    return out


class IdxTrainer:
    """Holds Adam state for orc_train_step_idx."""

    def __init__(self, cfg, solver, W, user_feat, item_feat, item_emb):
        self.cfg, self.solver = cfg, solver
        self.W = [np.ascontiguousarray(a, np.float32).copy() for a in W]
        self.uf = np.ascontiguousarray(user_feat, np.float32)
        self.itf = np.ascontiguousarray(item_feat, np.float32)
        self.emb = np.ascontiguousarray(item_emb, np.float32).copy()
        self._mv = [np.zeros_like(a) for a in self.W for _ in (0, 1)]
        self.st = AdamState(*[a.ctypes.data_as(_fp) for a in self._mv], 0)

    def step(self, user_row, item_row, hist, y, table_lr=0.0, nthreads=0, table_adam=False):
        ur, urp = _i(user_row); ir, irp = _i(item_row); hs, hsp = _i(hist); ya, yp = _f(y)
        B = hs.shape[0]; p = np.empty(B, np.float32)
        if table_adam and not hasattr(self, "emb_m"):
            self.emb_m = np.zeros_like(self.emb); self.emb_v = np.zeros_like(self.emb)
        mp = self.emb_m.ctypes.data_as(_fp) if table_adam else None
        vp = self.emb_v.ctypes.data_as(_fp) if table_adam else None
        cost = lib().orc_train_step_idx(
            C.byref(self.cfg), C.byref(self.solver), C.byref(self.st),
            *[a.ctypes.data_as(_fp) for a in self.W],
            self.uf.ctypes.data_as(_fp), C.c_long(self.uf.shape[1]),
            self.itf.ctypes.data_as(_fp), C.c_long(self.itf.shape[1]),
            self.emb.ctypes.data_as(_fp), C.c_long(self.emb.shape[1]), C.c_long(self.emb.shape[0]),
            urp, irp, hsp, yp, C.c_int(B), C.c_float(table_lr), p.ctypes.data_as(_fp), C.c_int(nthreads), mp, vp)
        return float(cost), p


    def step_fast(self, user_row, item_row, hist, y, table_lr=0.0, nthreads=0):
        """The timed CPU arm (oracle/cpu_fast.c): same step in float32 with blocked, thread-parallel SGEMMs."""
        ur, urp = _i(user_row); ir, irp = _i(item_row); hs, hsp = _i(hist); ya, yp = _f(y)
        cost = lib().orc_fast_train_step_idx(
            C.byref(self.cfg), C.byref(self.solver), C.byref(self.st),
            *[a.ctypes.data_as(_fp) for a in self.W],
            self.uf.ctypes.data_as(_fp), C.c_long(self.uf.shape[1]),
            self.itf.ctypes.data_as(_fp), C.c_long(self.itf.shape[1]),
            self.emb.ctypes.data_as(_fp), C.c_long(self.emb.shape[1]), C.c_long(self.emb.shape[0]),
            urp, irp, hsp, yp, C.c_int(hs.shape[0]), C.c_float(table_lr), C.c_int(nthreads))
        return float(cost)


# ---- item2vec (BASELINE config 5) -----------------------------------------------------------------
class I2vCfg(C.Structure):
    _fields_ = [("dim", C.c_int), ("window", C.c_int), ("iter", C.c_int), ("min_count", C.c_int), ("max_depth", C.c_int),
                ("init_lr", C.c_double), ("min_lr", C.c_double), ("subsample", C.c_double), ("update_lr_batch", C.c_int),
                ("seed", C.c_uint32), ("rng_mode", C.c_int)]


def i2v_cfg(dim=16, window=5, iters=1, min_count=5, max_depth=100, init_lr=0.025, subsample=1e-3, seed=0, rng_mode=1):
    """embedding.TrainEmbedding's fixed options (wordemb.go:9-32, options.go:41-60)."""
    return I2vCfg(dim, window, iters, min_count, max_depth, init_lr, init_lr * 1e-4, subsample, 100000, seed, rng_mode)


def i2v_huffman(count, literal=True):
    cnt = np.ascontiguousarray(count, np.int64); V = cnt.size
    parent = np.empty(2 * V - 1, np.int32); code = np.empty(2 * V - 1, np.uint8)
    lib().orc_i2v_huffman(cnt.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(V), C.c_int(int(literal)),
                          parent.ctypes.data_as(_ip), code.ctypes.data_as(C.POINTER(C.c_uint8)))
    return parent, code


def i2v_path(parent, code, V, w, max_depth=100):
    nodes = np.empty(4096, np.int32); codes = np.empty(4096, np.uint8)
    n = lib().orc_i2v_path(parent.ctypes.data_as(_ip), code.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(V), C.c_int(w),
                           C.c_int(max_depth), nodes.ctypes.data_as(_ip), codes.ctypes.data_as(C.POINTER(C.c_uint8)))
    return nodes[:n].copy(), codes[:n].copy()


def i2v_train(cfg, tokens, V, want_syn1=False):
    tok, tp = _i(tokens)
    emb = np.empty((V, cfg.dim), np.float32)
    syn1 = np.zeros((max(V - 1, 1), cfg.dim), np.float64) if want_syn1 else None
    trained = lib().orc_i2v_train(C.byref(cfg), tp, C.c_long(tok.size), C.c_int(V), emb.ctypes.data_as(_fp),
                                  syn1.ctypes.data_as(C.POINTER(C.c_double)) if want_syn1 else None)
    return (emb, syn1, trained) if want_syn1 else (emb, trained)


# ---- float64 MLP (row a11, BASELINE configs[0]) -----------------------------------------------------
class Mlp64Cfg(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("units", C.c_int * 8), ("hidden_act", C.c_int), ("batch", C.c_int),
                ("max_iter", C.c_int), ("n_iter_no_change", C.c_int), ("shuffle", C.c_int), ("adaptive", C.c_int),
                ("seed", C.c_uint32), ("alpha", C.c_double), ("lr_init", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("tol", C.c_double)]


_dp = C.POINTER(C.c_double)
MLP_ACT = {"relu": 0, "logistic": 1, "identity": 2}


def mlp64_cfg(n_features, hidden=(100,), activation="relu", alpha=1e-4, batch=200, max_iter=200, lr_init=1e-3,
              adaptive=False, shuffle=True, seed=0, tol=1e-4, n_iter_no_change=10, n_outputs=1):
    """NewMLPClassifier(hidden, activation, "adam", alpha) with NewBaseMultilayerPerceptron64's defaults
    (multilayer_perceptron.go:81-90, basemlp64.go:228-256)."""
    units = [n_features, *hidden, n_outputs]
    c = Mlp64Cfg()
    c.n_layers = len(units)
    for i, u in enumerate(units):
        c.units[i] = u
    c.hidden_act = MLP_ACT[activation]; c.batch = batch; c.max_iter = max_iter; c.n_iter_no_change = n_iter_no_change
    c.shuffle = int(shuffle); c.adaptive = int(adaptive); c.seed = seed; c.alpha = alpha; c.lr_init = lr_init
    c.beta1, c.beta2, c.eps, c.tol = 0.9, 0.999, 1e-8, tol
    return c


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def mlp64_nparams(cfg):
    return int(lib().orc_mlp64_nparams(C.byref(cfg)))


def mlp64_init(cfg):
    p = np.empty(mlp64_nparams(cfg), np.float64)
    lib().orc_mlp64_init(C.byref(cfg), p.ctypes.data_as(_dp))
    return p


def mlp64_loss_grad(cfg, params, X, y):
    X, xp = _d(X); y, yp = _d(y); params, pp = _d(params)
    g = np.empty_like(params)
    loss = lib().orc_mlp64_loss_grad(C.byref(cfg), pp, xp, yp, C.c_long(X.shape[0]), g.ctypes.data_as(_dp))
    return float(loss), g


def mlp64_fit(cfg, params, X, y):
    """in-place on params; returns (n_iter, loss_curve, final lr_init)."""
    X, xp = _d(X); y, yp = _d(y)
    assert params.dtype == np.float64 and params.flags.c_contiguous
    curve = np.zeros(cfg.max_iter, np.float64); lr = C.c_double(0)
    it = lib().orc_mlp64_fit(C.byref(cfg), params.ctypes.data_as(_dp), xp, yp, C.c_long(X.shape[0]),
                             curve.ctypes.data_as(_dp), C.byref(lr))
    return it, curve[:it], lr.value


def mlp64_predict(cfg, params, X):
    X, xp = _d(X); params, pp = _d(params)
    out = np.empty((X.shape[0], cfg.units[cfg.n_layers - 1]), np.float64)
    lib().orc_mlp64_predict(C.byref(cfg), pp, xp, C.c_long(X.shape[0]), out.ctypes.data_as(_dp))
    return out


class SimpleMlpFitWrap:
    """model/mlp/mlp.go:41-65 — float32 TrainSample in, predictor returning float32 [n,1] out."""

    def __init__(self, cfg):
        self.cfg = cfg

    def Fit(self, X32, Y32):
        X, xp = _f(X32); Y, yp = _f(Y32)
        params = np.empty(mlp64_nparams(self.cfg), np.float64)
        curve = np.zeros(self.cfg.max_iter, np.float64)
        it = lib().orc_mlp_fit_wrap(C.byref(self.cfg), params.ctypes.data_as(_dp), xp, yp, C.c_long(X.shape[0]),
                                    curve.ctypes.data_as(_dp))
        return SimpleMlpPredWrap(self.cfg, params, curve[:it])


class SimpleMlpPredWrap:
    def __init__(self, cfg, params, loss_curve):
        self.cfg, self.params, self.loss_curve = cfg, params, loss_curve

    def Predict(self, X32):
        X, xp = _f(X32)
        out = np.empty((X.shape[0], 1), np.float32)
        lib().orc_mlp_predict_wrap(C.byref(self.cfg), self.params.ctypes.data_as(_dp), xp, C.c_long(X.shape[0]),
                                   out.ctypes.data_as(_fp))
        return out
