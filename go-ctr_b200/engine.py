"""ctypes binding of libctr_b200.so (include/ctr_b200.h).  This is the same C ABI the cgo shim
binds (go/ctrb200/ctrb200.go); Python is only the test / bench harness language here because the
image has no Go toolchain.  There is no CPU fallback: if the shared library is missing or no B200
is present, construction raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = ["Engine", "Config", "MLPClassifier", "MlpConfig", "StepStats", "CtrError", "load_library", "MODEL_YOUTUBE", "MODEL_DIN_COS",
           "MODEL_DIN_EUC", "TABLE_USER_FEAT", "TABLE_ITEM_FEAT", "TABLE_ITEM_EMB", "TABLE_FROZEN", "TABLE_SGD",
           "TABLE_SGD_DETERMINISTIC", "TABLE_ADAM", "GEMM_AUTO", "GEMM_FP32", "GEMM_TCGEN05_3XTF32", "EXPORTS", "IDMAP_USER", "IDMAP_ITEM", "ENOTFOUND"]

MODEL_YOUTUBE, MODEL_DIN_COS, MODEL_DIN_EUC = 0, 1, 2
TABLE_USER_FEAT, TABLE_ITEM_FEAT, TABLE_ITEM_EMB = 0, 1, 2
TABLE_FROZEN, TABLE_SGD, TABLE_SGD_DETERMINISTIC, TABLE_ADAM = 0, 1, 2, 3
GEMM_AUTO, GEMM_FP32, GEMM_TCGEN05_3XTF32 = 0, 1, 2
IDMAP_USER, IDMAP_ITEM = 0, 1
ENOTFOUND = 7

# every symbol include/ctr_b200.h declares (tests check the .so exports all of them)
EXPORTS = ["ctr_abi_version", "ctr_config_default", "ctr_create", "ctr_destroy", "ctr_last_error",
           "ctr_init_weights", "ctr_set_weights", "ctr_get_weights", "ctr_table_upload", "ctr_table_download", "ctr_table_fill",
           "ctr_gather_rows", "ctr_train_dense", "ctr_predict_dense", "ctr_train_step_idx", "ctr_train_idx", "ctr_train_keys", "ctr_predict_idx",
           "ctr_train_step_idx_dev", "ctr_predict_idx_dev", "ctr_last_cost", "ctr_sync", "ctr_get_stream",
           "ctr_set_stream", "ctr_launch_count", "ctr_profile_enable", "ctr_profile_get", "ctr_profile_reset",
           "ctr_profile_dump", "ctr_debug_grads_idx", "ctr_ubcache_upload", "ctr_ubcache_window", "ctr_ubcache_window_dev", "ctr_idmap_build", "ctr_idmap_lookup", "ctr_idmap_lookup_dev",
           "ctr_batch_predict_keys", "ctr_checkpoint_save", "ctr_checkpoint_load", "ctr_roc_auc", "ctr_i2v_config_default", "ctr_i2v_paths", "ctr_i2v_train", "ctr_i2v_train_dist", "ctr_comm_unique_id", "ctr_comm_init",
           "ctr_mlp_config_default", "ctr_mlp_create", "ctr_mlp_destroy", "ctr_mlp_last_error", "ctr_mlp_fit", "ctr_mlp_predict",
           "ctr_mlp_get_params", "ctr_mlp_set_params"]


class CtrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ctr_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("uP", C.c_int32), ("S", C.c_int32), ("D", C.c_int32), ("cF", C.c_int32),
                ("H0", C.c_int32), ("H1", C.c_int32), ("batch", C.c_int32), ("pred_batch", C.c_int32),
                ("lr", C.c_float), ("l2", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("dropout0", C.c_float), ("dropout1", C.c_float), ("seed", C.c_uint32),
                ("table_opt", C.c_int32), ("table_lr", C.c_float), ("gemm", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32), ("reserved", C.c_int32 * 8)]


class MlpConfig(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("units", C.c_int32 * 8), ("hidden_act", C.c_int32), ("batch", C.c_int32), ("max_iter", C.c_int32),
                ("n_iter_no_change", C.c_int32), ("shuffle", C.c_int32), ("adaptive", C.c_int32), ("warm_start", C.c_int32), ("seed", C.c_uint32),
                ("device", C.c_int32), ("alpha", C.c_double), ("lr_init", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("tol", C.c_double)]


class StepStats(C.Structure):
    _fields_ = [("cost", C.c_float), ("ms_device", C.c_float), ("launches", C.c_int32), ("reserved", C.c_int32)]


_lib = None
_fp = C.POINTER(C.c_float)
_lp = C.POINTER(C.c_int64)
_ip = C.POINTER(C.c_int32)


def load_library():
    """Loads the in-tree libctr_b200.so.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.SO):
        raise ImportError("libctr_b200.so is not built: run `python __graft_entry__.py build` "
                          "(this package has no CPU fallback)")
    L = C.CDLL(_build.SO)
    L.ctr_last_error.restype = C.c_char_p
    L.ctr_last_error.argtypes = [C.c_void_p]
    L.ctr_get_stream.restype = C.c_void_p
    L.ctr_get_stream.argtypes = [C.c_void_p]
    L.ctr_launch_count.restype = C.c_int64
    L.ctr_launch_count.argtypes = [C.c_void_p]
    L.ctr_destroy.argtypes = [C.c_void_p]
    L.ctr_destroy.restype = None
    L.ctr_config_default.restype = None
    L.ctr_mlp_last_error.restype = C.c_char_p
    L.ctr_mlp_last_error.argtypes = [C.c_void_p]
    L.ctr_mlp_destroy.argtypes = [C.c_void_p]
    L.ctr_mlp_destroy.restype = None
    L.ctr_mlp_config_default.restype = None
    _lib = L
    return L


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def default_config(model, **kw):
    cfg = Config()
    load_library().ctr_config_default(C.byref(cfg), C.c_int(model))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class Engine:
    """One engine handle == one DinNet / YoutubeDnn plus its HBM tables (din.go:171, dnn.go:119)."""

    def __init__(self, cfg):
        self.L = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.ctr_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise CtrError(rc, self.L.ctr_last_error(None).decode())
        self.inn = cfg.uP + 2 * cfg.D + cfg.cF

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ctr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CtrError(rc, self.L.ctr_last_error(self.h).decode())

    # ---- weights (Marshal layout, din.go:41-52)
    def init_weights(self, seed):
        self._ck(self.L.ctr_init_weights(self.h, C.c_uint32(seed)))

    def set_weights(self, mlp0, mlp1, mlp2, att0=None):
        c = self.cfg
        a0, p0 = _f(mlp0); a1, p1 = _f(mlp1); a2, p2 = _f(mlp2)
        assert a0.size == self.inn * c.H0 and a1.size == c.H0 * c.H1 and a2.size == c.H1
        if att0 is not None:
            a3, p3 = _f(att0); assert a3.size == c.S
        else:
            p3 = None
        self._ck(self.L.ctr_set_weights(self.h, p0, p1, p2, p3))

    def get_weights(self):
        c = self.cfg
        w0 = np.empty((self.inn, c.H0), np.float32); w1 = np.empty((c.H0, c.H1), np.float32)
        w2 = np.empty((c.H1, 1), np.float32); at = np.empty(c.S, np.float32)
        self._ck(self.L.ctr_get_weights(self.h, w0.ctypes.data_as(_fp), w1.ctypes.data_as(_fp),
                                        w2.ctypes.data_as(_fp), at.ctypes.data_as(_fp)))
        return w0, w1, w2, at

    # ---- tables
    def table_upload(self, which, rows):
        a, p = _f(rows)
        self._ck(self.L.ctr_table_upload(self.h, C.c_int(which), p, C.c_int64(a.shape[0]), C.c_int32(a.shape[1])))

    def table_fill(self, which, nrows, width, seed=0, dist=0, scale=1.0):
        self._ck(self.L.ctr_table_fill(self.h, C.c_int(which), C.c_int64(nrows), C.c_int32(width), C.c_uint32(seed),
                                       C.c_int32(dist), C.c_float(scale)))

    def table_download(self, which, nrows, width):
        out = np.zeros((nrows, width), np.float32)
        self._ck(self.L.ctr_table_download(self.h, C.c_int(which), out.ctypes.data_as(_fp), C.c_int64(nrows), C.c_int32(width)))
        return out

    def gather_rows(self, user_row, item_row, hist):
        c = self.cfg
        u, up = _i(user_row); it, ip = _i(item_row); hs, hp = _i(hist)
        B = u.size
        X = np.empty((B, c.uP + c.S * c.D + c.D + c.cF), np.float32)
        self._ck(self.L.ctr_gather_rows(self.h, up, ip, hp, C.c_int64(B), X.ctypes.data_as(_fp)))
        return X

    # ---- dense X (model.Train / model.Predict)
    def train_dense(self, X, Y, ranges, epochs, early_stop=0):
        Xa, Xp = _f(X); Ya, Yp = _f(Y)
        r = (C.c_int32 * 8)(*[int(v) for v in ranges])
        cost = C.c_float(0); ep = C.c_int32(0)
        self._ck(self.L.ctr_train_dense(self.h, Xp, Yp, C.c_int64(Xa.shape[0]), C.c_int32(Xa.shape[1]), r,
                                        C.c_int32(epochs), C.c_int32(early_stop), C.byref(cost), C.byref(ep)))
        return ep.value, cost.value

    def predict_dense(self, X, ranges):
        Xa, Xp = _f(X)
        r = (C.c_int32 * 8)(*[int(v) for v in ranges])
        out = np.empty(Xa.shape[0], np.float32)
        self._ck(self.L.ctr_predict_dense(self.h, Xp, C.c_int64(Xa.shape[0]), C.c_int32(Xa.shape[1]), r, out.ctypes.data_as(_fp)))
        return out

    # ---- index fast path
    def train_step_idx(self, user_row, item_row, hist, label):
        u, up = _i(user_row); it, ip = _i(item_row); hs, hp = _i(hist); y, yp = _f(label)
        st = StepStats()
        self._ck(self.L.ctr_train_step_idx(self.h, up, ip, hp, yp, C.c_int32(u.size), C.byref(st)))
        return st

    def train_step_idx_ptr(self, up, ip, hp, yp, B, stats=None):
        """Raw host pointers (ints) — pinned buffers owned by the caller (bench e2e leg)."""
        self._ck(self.L.ctr_train_step_idx(self.h, C.c_void_p(up), C.c_void_p(ip), C.c_void_p(hp), C.c_void_p(yp),
                                           C.c_int32(B), C.byref(stats) if stats is not None else None))

    def train_step_idx_dev(self, d_user, d_item, d_hist, d_label, B):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous."""
        self._ck(self.L.ctr_train_step_idx_dev(self.h, C.c_void_p(d_user), C.c_void_p(d_item), C.c_void_p(d_hist),
                                               C.c_void_p(d_label), C.c_int32(B)))

    def train_idx(self, user_row, item_row, hist, label):
        """One pass over n samples in batches of cfg.batch (pipelined H2D); returns the batch costs."""
        u, up = _i(user_row); it, ip = _i(item_row); hs, hp = _i(hist); y, yp = _f(label)
        nb = (u.size + self.cfg.batch - 1) // self.cfg.batch
        costs = np.empty(nb, np.float32)
        self._ck(self.L.ctr_train_idx(self.h, up, ip, hp, yp, C.c_int64(u.size), costs.ctypes.data_as(_fp)))
        return costs

    def train_idx_ptr(self, up, ip, hp, yp, n, costs_ptr=None):
        """Raw host pointers (pinned buffers owned by the caller)."""
        self._ck(self.L.ctr_train_idx(self.h, C.c_void_p(up), C.c_void_p(ip), C.c_void_p(hp), C.c_void_p(yp), C.c_int64(n),
                                      C.c_void_p(costs_ptr) if costs_ptr else None))

    def train_keys(self, user_ids, item_ids, ts, label, epochs=1, early_stop=0):
        """recommend.Train over sample keys (rcmd.go:197-246); returns (epochs_run, last_cost, rows_used)."""
        u = np.ascontiguousarray(user_ids, np.int64); i = np.ascontiguousarray(item_ids, np.int64); t = np.ascontiguousarray(ts, np.int64)
        y, yp = _f(label)
        assert u.size == i.size == t.size == y.size
        cost = C.c_float(0); ep = C.c_int32(0); used = C.c_int64(0)
        self._ck(self.L.ctr_train_keys(self.h, u.ctypes.data_as(_lp), i.ctypes.data_as(_lp), t.ctypes.data_as(_lp), yp, C.c_int64(u.size),
                                       C.c_int32(epochs), C.c_int32(early_stop), C.byref(cost), C.byref(ep), C.byref(used)))
        return ep.value, cost.value, used.value

    def predict_idx(self, user_row, item_row, hist):
        u, up = _i(user_row); it, ip = _i(item_row); hs, hp = _i(hist)
        out = np.empty(u.size, np.float32)
        self._ck(self.L.ctr_predict_idx(self.h, up, ip, hp, C.c_int64(u.size), out.ctypes.data_as(_fp)))
        return out

    def predict_idx_dev(self, d_user, d_item, d_hist, B, d_out=None):
        self._ck(self.L.ctr_predict_idx_dev(self.h, C.c_void_p(d_user), C.c_void_p(d_item), C.c_void_p(d_hist),
                                            C.c_int32(B), C.c_void_p(d_out) if d_out else None))

    def last_cost(self):
        c = C.c_float(0)
        self._ck(self.L.ctr_last_cost(self.h, C.byref(c)))
        return c.value

    def sync(self):
        self._ck(self.L.ctr_sync(self.h))

    @property
    def stream(self):
        return self.L.ctr_get_stream(self.h)

    def set_stream(self, ptr):
        self._ck(self.L.ctr_set_stream(self.h, C.c_void_p(ptr)))

    def launch_count(self):
        return int(self.L.ctr_launch_count(self.h))

    def profile(self, on):
        self._ck(self.L.ctr_profile_enable(self.h, C.c_int(int(on))))

    def profile_reset(self):
        self._ck(self.L.ctr_profile_reset(self.h))

    def profile_dump(self):
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.L.ctr_profile_dump(self.h, buf, C.c_int64(len(buf))))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def debug_grads_idx(self, user_row, item_row, hist, label, training=False):
        c = self.cfg
        u, up = _i(user_row); it, ip = _i(item_row); hs, hp = _i(hist); y, yp = _f(label)
        B = u.size
        g0 = np.empty((self.inn, c.H0), np.float32); g1 = np.empty((c.H0, c.H1), np.float32)
        g2 = np.empty((c.H1, 1), np.float32); ga = np.empty(c.S, np.float32)
        dUb = np.empty((B, c.S, c.D), np.float32); dIt = np.empty((B, c.D), np.float32)
        p = np.empty(B, np.float32); z = np.empty(B, np.float32); cost = C.c_float(0)
        self._ck(self.L.ctr_debug_grads_idx(self.h, up, ip, hp, yp, C.c_int32(B), C.c_int32(int(training)),
                                            g0.ctypes.data_as(_fp), g1.ctypes.data_as(_fp), g2.ctypes.data_as(_fp),
                                            ga.ctypes.data_as(_fp), dUb.ctypes.data_as(_fp), dIt.ctypes.data_as(_fp),
                                            p.ctypes.data_as(_fp), z.ctypes.data_as(_fp), C.byref(cost)))
        return dict(cost=cost.value, dW0=g0, dW1=g1, dW2=g2, datt=ga, dUb=dUb, dIt=dIt, p=p, logit=z)

    def comm_unique_id(self):
        buf = C.create_string_buffer(256); n = C.c_int32(256)
        rc = self.L.ctr_comm_unique_id(buf, C.byref(n))
        if rc != 0:
            raise CtrError(rc, "ctr_comm_unique_id failed (NCCL not loadable?)")
        return bytes(buf.raw[:n.value])

    def comm_init(self, uid):
        self._ck(self.L.ctr_comm_init(self.h, C.c_char_p(uid), C.c_int32(len(uid))))

    def ubcache_upload(self, offsets, ts, item_rows):
        off = np.ascontiguousarray(offsets, np.int64); t = np.ascontiguousarray(ts, np.int64); it = np.ascontiguousarray(item_rows, np.int32)
        self._ck(self.L.ctr_ubcache_upload(self.h, off.ctypes.data_as(C.POINTER(C.c_int64)), t.ctypes.data_as(C.POINTER(C.c_int64)),
                                           it.ctypes.data_as(_ip), C.c_int64(off.size - 1), C.c_int64(t.size)))

    def ubcache_window(self, user_row, max_ts):
        u, up = _i(user_row); t = np.ascontiguousarray(max_ts, np.int64)
        out = np.empty((u.size, self.cfg.S), np.int32)
        self._ck(self.L.ctr_ubcache_window(self.h, up, t.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int32(u.size), out.ctypes.data_as(_ip)))
        return out

    def idmap_build(self, which, ids):
        a = np.ascontiguousarray(ids, np.int64)
        self._ck(self.L.ctr_idmap_build(self.h, C.c_int(which), a.ctypes.data_as(_lp), C.c_int64(a.size)))

    def idmap_lookup(self, which, ids):
        a = np.ascontiguousarray(ids, np.int64)
        out = np.empty(a.size, np.int32)
        self._ck(self.L.ctr_idmap_lookup(self.h, C.c_int(which), a.ctypes.data_as(_lp), C.c_int64(a.size), out.ctypes.data_as(_ip)))
        return out

    def batch_predict_keys(self, user_ids, item_ids, ts):
        """recommend.BatchPredict (rcmd.go:277-337) over Sample keys."""
        u = np.ascontiguousarray(user_ids, np.int64); i = np.ascontiguousarray(item_ids, np.int64); t = np.ascontiguousarray(ts, np.int64)
        assert u.size == i.size == t.size
        out = np.empty(u.size, np.float32)
        self._ck(self.L.ctr_batch_predict_keys(self.h, u.ctypes.data_as(_lp), i.ctypes.data_as(_lp), t.ctypes.data_as(_lp), C.c_int64(u.size),
                                                out.ctypes.data_as(_fp)))
        return out

    def checkpoint_save(self, path):
        self._ck(self.L.ctr_checkpoint_save(self.h, C.c_char_p(str(path).encode())))

    def checkpoint_load(self, path):
        self._ck(self.L.ctr_checkpoint_load(self.h, C.c_char_p(str(path).encode())))

    def roc_auc(self, pred, y):
        p, pp = _f(pred); t, tp = _f(y)
        auc = C.c_double(0)
        self._ck(self.L.ctr_roc_auc(self.h, pp, tp, C.c_int64(p.size), C.byref(auc)))
        return auc.value


MLP_ACT = {"relu": 0, "logistic": 1, "identity": 2}


class MLPClassifier:
    """nn.MLPClassifier on the device (float64): NewMLPClassifier(hidden, activation, "adam", alpha) with
    NewBaseMultilayerPerceptron64's defaults (multilayer_perceptron.go:81-90, basemlp64.go:227-256)."""

    def __init__(self, n_features, hidden=(100,), activation="relu", alpha=1e-4, batch=200, max_iter=200, lr_init=1e-3,
                 adaptive=False, shuffle=True, seed=0, tol=1e-4, n_iter_no_change=10, warm_start=False, device=0):
        self.L = load_library()
        cfg = MlpConfig()
        self.L.ctr_mlp_config_default(C.byref(cfg), C.c_int32(n_features))
        units = [n_features, *hidden, 1]
        cfg.n_layers = len(units)
        for i, u in enumerate(units):
            cfg.units[i] = u
        cfg.hidden_act = MLP_ACT[activation]; cfg.alpha = alpha; cfg.batch = batch; cfg.max_iter = max_iter; cfg.lr_init = lr_init
        cfg.adaptive = int(adaptive); cfg.shuffle = int(shuffle); cfg.seed = seed; cfg.tol = tol; cfg.n_iter_no_change = n_iter_no_change
        cfg.warm_start = int(warm_start); cfg.device = device
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.ctr_mlp_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise CtrError(rc, self.L.ctr_mlp_last_error(None).decode())
        self.n_iter = 0; self.loss_curve = np.zeros(0)

    def _ck(self, rc):
        if rc != 0:
            raise CtrError(rc, self.L.ctr_mlp_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ctr_mlp_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fit(self, X, Y):
        Xa, Xp = _f(X); Ya, Yp = _f(Y)
        it = C.c_int32(0); curve = np.zeros(self.cfg.max_iter, np.float64)
        self._ck(self.L.ctr_mlp_fit(self.h, Xp, Yp, C.c_int64(Xa.shape[0]), C.c_int32(Xa.shape[1]), C.byref(it), curve.ctypes.data_as(C.POINTER(C.c_double))))
        self.n_iter = it.value; self.loss_curve = curve[:it.value]
        return self

    def predict(self, X):
        Xa, Xp = _f(X)
        out = np.empty(Xa.shape[0], np.float32)
        self._ck(self.L.ctr_mlp_predict(self.h, Xp, C.c_int64(Xa.shape[0]), C.c_int32(Xa.shape[1]), out.ctypes.data_as(_fp)))
        return out

    def get_params(self):
        n = C.c_int64(0)
        self._ck(self.L.ctr_mlp_get_params(self.h, None, C.c_int64(0), C.byref(n)))
        p = np.empty(n.value, np.float64)
        self._ck(self.L.ctr_mlp_get_params(self.h, p.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n.value), C.byref(n)))
        return p

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float64)
        self._ck(self.L.ctr_mlp_set_params(self.h, p.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(p.size)))
