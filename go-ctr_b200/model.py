"""Host-side mirror of the reference's plugin / operator interface for the hot path, over the C ABI.

Names, argument order and error behaviour follow the Go reference so the parity tests read like
the reference's own tests (model/model_test.go, example/movielens/dinimpl_test.go):

    recommend.SampleInfo / TrainSample           rcmd.go:56-63, 132-137
    din.NewDinNet / NewDinNetFromJson / Marshal   din.go:62-147, 171
    youtube.NewYoutubeDnn / ...FromJson / Marshal dnn.go:49-160
    model.Train / InitForwardOnlyVm / Predict     model.go:27, 215, 242
    movielens dinImpl / YoutubeDnnImpl (Fitter + PredictAbstract)   dinimpl.go:13-92, youtube.go:13-91
    mlp.SimpleMlpFitWrap / SimpleMlpPredWrap (Fitter + PredictAbstract) model/mlp/mlp.go:11-65

The gorgonia-typed half of model.Model (Fwd/Graph/Vm, model.go:16-25) has no meaning on a GPU
engine; the seam is Fitter / PredictAbstract + Train / Predict + the Marshal JSON schema
(SURVEY.md §8b).  The Go binding of the same C ABI is go/ctrb200/ctrb200.go.
"""
import json
from dataclasses import dataclass, field

import numpy as np

from . import engine as _e

__all__ = ["SampleInfo", "TrainSample", "DinNet", "YoutubeDnn", "NewDinNet", "NewDinNetFromJson",
           "NewYoutubeDnn", "NewYoutubeDnnFromJson", "Train", "InitForwardOnlyVm", "Predict",
           "DinImpl", "YoutubeDnnImpl", "RocAuc32", "SimpleMlpFitWrap", "SimpleMlpPredWrap"]

mlp0_1, mlp1_2 = 200, 80          # din.go:17-18


@dataclass
class SampleInfo:                 # rcmd.go:132-137
    UserProfileRange: tuple = (0, 0)
    UserBehaviorRange: tuple = (0, 0)
    ItemFeatureRange: tuple = (0, 0)
    CtxFeatureRange: tuple = (0, 0)

    def flat(self):
        return [*self.UserProfileRange, *self.UserBehaviorRange, *self.ItemFeatureRange, *self.CtxFeatureRange]


@dataclass
class TrainSample:                # rcmd.go:56-63
    X: np.ndarray = None
    Y: np.ndarray = None
    Rows: int = 0
    XCols: int = 0
    Info: SampleInfo = field(default_factory=SampleInfo)


class _Net:
    """Common state of DinNet / YoutubeDnn: dims + weights in the Marshal layout."""
    kind = None
    d0 = d1 = 0.0

    def __init__(self, uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, seed=0, gemm=_e.GEMM_AUTO):
        if uBehaviorDim != iFeatureDim:           # din.go:176-178
            raise ValueError("uBehaviorDim %d != iFeatureDim %d" % (uBehaviorDim, iFeatureDim))
        self.uProfileDim, self.uBehaviorSize, self.uBehaviorDim = uProfileDim, uBehaviorSize, uBehaviorDim
        self.iFeatureDim, self.cFeatureDim = iFeatureDim, cFeatureDim
        self.seed, self.gemm = seed, gemm
        self.weights = None           # (mlp0, mlp1, mlp2, att0) host copies, set after Train / FromJson
        self.engine = None            # the "VM" (model.go:81, 237)
        self.batchSize = 0

    def _config(self, batch, pred_batch, training):
        return _e.default_config(self.kind, uP=self.uProfileDim, S=self.uBehaviorSize, D=self.uBehaviorDim,
                                 cF=self.cFeatureDim, H0=mlp0_1, H1=mlp1_2, batch=batch, pred_batch=pred_batch,
                                 dropout0=self.d0 if training else 0.0, dropout1=self.d1 if training else 0.0,
                                 seed=self.seed, gemm=self.gemm)

    def Marshal(self):
        if self.weights is None:
            if self.engine is None:
                raise RuntimeError("model has no weights yet")
            self.weights = self.engine.get_weights()
        w0, w1, w2, att = self.weights
        m = {"uProfileDim": self.uProfileDim, "uBehaviorSize": self.uBehaviorSize, "uBehaviorDim": self.uBehaviorDim,
             "iFeatureDim": self.iFeatureDim, "cFeatureDim": self.cFeatureDim,
             "mlp0": [float(v) for v in np.asarray(w0, np.float32).ravel()],
             "mlp1": [float(v) for v in np.asarray(w1, np.float32).ravel()],
             "mlp2": [float(v) for v in np.asarray(w2, np.float32).ravel()]}
        if self.kind != _e.MODEL_YOUTUBE:
            m["att0"] = [float(v) for v in np.asarray(att, np.float32).ravel()]
        return json.dumps(m).encode()

    @classmethod
    def _from_json(cls, data):
        m = json.loads(data)
        net = cls(m["uProfileDim"], m["uBehaviorSize"], m["uBehaviorDim"], m["iFeatureDim"], m["cFeatureDim"])
        net.d0 = net.d1 = 0.0     # FromJson leaves d0/d1 unset ⇒ dropout is the identity (din.go:133-145)
        inn = m["uProfileDim"] + m["uBehaviorDim"] + m["iFeatureDim"] + m["cFeatureDim"]
        att = np.asarray(m.get("att0", [1.0] * m["uBehaviorSize"]), np.float32)
        net.weights = (np.asarray(m["mlp0"], np.float32).reshape(inn, mlp0_1),
                       np.asarray(m["mlp1"], np.float32).reshape(mlp0_1, mlp1_2),
                       np.asarray(m["mlp2"], np.float32).reshape(mlp1_2, 1), att)
        return net


class DinNet(_Net):
    kind = _e.MODEL_DIN_COS
    d0 = d1 = 0.005               # din.go:204-205


class YoutubeDnn(_Net):
    kind = _e.MODEL_YOUTUBE
    d0 = d1 = 0.003               # dnn.go:136-137


def NewDinNet(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, **kw):
    return DinNet(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, **kw)


def NewYoutubeDnn(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, **kw):
    return YoutubeDnn(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, **kw)


def NewDinNetFromJson(data):
    return DinNet._from_json(data)


def NewYoutubeDnnFromJson(data):
    return YoutubeDnn._from_json(data)


def _check_dims(m, uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim):
    if (m.uProfileDim, m.uBehaviorSize, m.uBehaviorDim, m.iFeatureDim, m.cFeatureDim) != \
            (uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim):
        raise ValueError("model dims do not match the arguments")
    if uBehaviorDim != iFeatureDim:               # Fwd: din.go:221-223
        raise ValueError("uBehaviorDim %d != iFeatureDim %d" % (uBehaviorDim, iFeatureDim))


def Train(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim,
          numExamples, batchSize, epochs, earlyStop, si, inputs, targets, m):
    """model.Train (model.go:27-213).  inputs [numExamples, XCols] float32, targets [numExamples(,1)].
    Returns (epochs_run, last_cost); the Go function returns only err — errors raise here."""
    _check_dims(m, uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim)
    inputs = np.ascontiguousarray(inputs, np.float32)
    targets = np.ascontiguousarray(targets, np.float32).reshape(-1)
    if inputs.shape[0] < numExamples or targets.size < numExamples:
        raise ValueError("numExamples exceeds the tensors")
    eng = _e.Engine(m._config(batchSize, batchSize, training=True))
    if m.weights is not None:
        eng.set_weights(*m.weights)
    else:
        eng.init_weights(m.seed)                  # G.Gaussian(0,1) / ValuesOf(1): din.go:181-191
    ep, cost = eng.train_dense(inputs[:numExamples], targets[:numExamples], si.flat(), epochs, earlyStop)
    m.weights = eng.get_weights()
    m.engine = eng
    m.batchSize = batchSize
    return ep, cost


def InitForwardOnlyVm(uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, batchSize, m):
    """model.InitForwardOnlyVm (model.go:215-240): builds the forward-only engine for Predict."""
    _check_dims(m, uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim)
    if m.weights is None:
        raise RuntimeError("model has no weights")
    eng = _e.Engine(m._config(batchSize, batchSize, training=False))
    eng.set_weights(*m.weights)
    m.engine = eng
    m.batchSize = batchSize


def Predict(m, numExamples, batchSize, si, inputs):
    """model.Predict (model.go:242-353) → []float32 of length numExamples."""
    if m.engine is None:
        raise RuntimeError("InitForwardOnlyVm was not called")
    if batchSize != m.batchSize:
        raise ValueError("batchSize differs from InitForwardOnlyVm's")
    inputs = np.ascontiguousarray(inputs, np.float32)
    return m.engine.predict_dense(inputs[:numExamples], si.flat())


class _Impl:
    """example/movielens dinImpl / YoutubeDnnImpl: recommend.Fitter + recommend.PredictAbstract
    (dinimpl.go:13-92, youtube.go:13-91)."""
    _new = None
    _from_json = None

    def __init__(self, uBehaviorSize, uBehaviorDim, PredBatchSize=100, BatchSize=200, epochs=100, earlyStop=0, seed=0):
        self.uBehaviorSize, self.uBehaviorDim = uBehaviorSize, uBehaviorDim    # rcmd.UserBehaviorLen / ItemEmbDim
        self.PredBatchSize, self.BatchSize, self.epochs, self.earlyStop = PredBatchSize, BatchSize, epochs, earlyStop
        self.seed = seed
        self.learner = self.pred = self.sampleInfo = None

    def Fit(self, trainSample):
        si = trainSample.Info
        self.uProfileDim = si.UserProfileRange[1] - si.UserProfileRange[0]
        self.cFeatureDim = si.CtxFeatureRange[1] - si.CtxFeatureRange[0]
        self.iFeatureDim = self.uBehaviorDim
        self.sampleInfo = si
        if trainSample.Rows != len(trainSample.Y):       # dinimpl.go:52-56
            raise ValueError("number of examples %d and labels %d do not match" % (trainSample.Rows, len(trainSample.Y)))
        inputs = np.asarray(trainSample.X, np.float32).reshape(trainSample.Rows, trainSample.XCols)
        dims = (self.uProfileDim, self.uBehaviorSize, self.uBehaviorDim, self.iFeatureDim, self.cFeatureDim)
        self.learner = type(self)._new(*dims, seed=self.seed)
        Train(*dims, trainSample.Rows, self.BatchSize, self.epochs, self.earlyStop, si, inputs, trainSample.Y, self.learner)
        self.pred = type(self)._from_json(self.learner.Marshal())       # dinimpl.go:73-78
        InitForwardOnlyVm(*dims, self.PredBatchSize, self.pred)
        return self

    def Predict(self, X):
        X = np.asarray(X, np.float32)
        y = Predict(self.pred, X.shape[0], self.PredBatchSize, self.sampleInfo, X)
        return y.reshape(-1, 1)                                          # dinimpl.go:39


class DinImpl(_Impl):
    _new = staticmethod(NewDinNet)
    _from_json = staticmethod(NewDinNetFromJson)


class YoutubeDnnImpl(_Impl):
    _new = staticmethod(NewYoutubeDnn)
    _from_json = staticmethod(NewYoutubeDnnFromJson)


class SimpleMlpPredWrap:
    """model/mlp/mlp.go:11-39: PredictAbstract over the fitted classifier — float32 X in, float32 [n, 1] out."""

    def __init__(self, pred):
        self.pred = pred

    def Predict(self, X):
        X = np.asarray(X, np.float32)
        return self.pred.predict(X).reshape(-1, 1)


class SimpleMlpFitWrap:
    """model/mlp/mlp.go:41-65: Fitter over nn.MLPClassifier (main.go:42-52 builds it with NewMLPClassifier([]int{100},
    "relu", "adam", 1e-4) — pass the same through **mlp_kw).  Model is created at Fit time from the sample width."""

    def __init__(self, **mlp_kw):
        self.mlp_kw = mlp_kw
        self.Model = None

    def Fit(self, trainSample):
        X = np.asarray(trainSample.X, np.float32).reshape(trainSample.Rows, trainSample.XCols)
        Y = np.asarray(trainSample.Y, np.float32)
        self.Model = _e.MLPClassifier(trainSample.XCols, **self.mlp_kw)
        self.Model.fit(X, Y)
        return SimpleMlpPredWrap(self.Model)


def RocAuc32(pred, y, engine=None):
    """utils.RocAuc32 (util.go:131-148) on the device."""
    eng = engine or _e.Engine(_e.default_config(_e.MODEL_YOUTUBE, batch=1, pred_batch=1))
    return eng.roc_auc(pred, y)
