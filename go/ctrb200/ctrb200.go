// Package ctrb200 is the cgo shim that maps go-ctr's unchanged plugin surface —
// recommend.Fitter / recommend.PredictAbstract (recommend/rcmd.go:87-97) and the
// model.Train / model.Predict call shape (model/model.go:27, 242) — onto libctr_b200.so.
//
// NOT BUILT IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md §8b "Compile
// reality").  It is the binding a go-ctr maintainer adds; every C entry point it calls is covered
// by the Python ctypes tests in tests/, which drive the identical C ABI.
//
// Build (on a machine with Go, CUDA 12.9 and a B200):
//
//	CGO_CFLAGS="-I${CTR_B200}/include" CGO_LDFLAGS="-L${CTR_B200}/go-ctr_b200 -lctr_b200" go build ./...
package ctrb200

/*
#cgo CFLAGS:  -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../go-ctr_b200 -lctr_b200 -Wl,-rpath,${SRCDIR}/../../go-ctr_b200
#include <stdlib.h>
#include "ctr_b200.h"
*/
import "C"

import (
	"encoding/json"
	"fmt"
	"runtime"
	"unsafe"

	rcmd "github.com/auxten/go-ctr/recommend"
	"gorgonia.org/tensor"
)

// Kind selects the graph the engine evaluates.
type Kind int

const (
	YoutubeDnn Kind = C.CTR_MODEL_YOUTUBE // model/youtube/dnn.go
	DinCosine  Kind = C.CTR_MODEL_DIN_COS // model/din/din.go (live variant)
	DinEuclid  Kind = C.CTR_MODEL_DIN_EUC // din.go:230
)

// Engine owns one device handle (weights + HBM tables).
type Engine struct {
	h   *C.ctr_handle
	cfg C.ctr_config
}

func (e *Engine) err(rc C.int) error {
	if rc == C.CTR_OK {
		return nil
	}
	return fmt.Errorf("ctr_b200 error %d: %s", int(rc), C.GoString(C.ctr_last_error(e.h)))
}

// NewEngine mirrors din.NewDinNet / youtube.NewYoutubeDnn (din.go:171, dnn.go:119).
func NewEngine(kind Kind, uProfileDim, uBehaviorSize, uBehaviorDim, iFeatureDim, cFeatureDim, batchSize, predBatchSize int) (*Engine, error) {
	if uBehaviorDim != iFeatureDim { // din.go:176-178
		return nil, fmt.Errorf("uBehaviorDim %d != iFeatureDim %d", uBehaviorDim, iFeatureDim)
	}
	e := &Engine{}
	C.ctr_config_default(&e.cfg, C.int(kind))
	e.cfg.uP, e.cfg.S, e.cfg.D, e.cfg.cF = C.int32_t(uProfileDim), C.int32_t(uBehaviorSize), C.int32_t(uBehaviorDim), C.int32_t(cFeatureDim)
	e.cfg.batch, e.cfg.pred_batch = C.int32_t(batchSize), C.int32_t(predBatchSize)
	if rc := C.ctr_create(&e.cfg, &e.h); rc != C.CTR_OK {
		return nil, fmt.Errorf("ctr_create: %d: %s", int(rc), C.GoString(C.ctr_last_error(nil)))
	}
	runtime.SetFinalizer(e, func(e *Engine) { e.Close() })
	return e, nil
}

func (e *Engine) Close() {
	if e.h != nil {
		C.ctr_destroy(e.h)
		e.h = nil
	}
}

func f32(p []float32) *C.float { return (*C.float)(unsafe.Pointer(&p[0])) }
func i32(p []int32) *C.int32_t { return (*C.int32_t)(unsafe.Pointer(&p[0])) }

func ranges(si *rcmd.SampleInfo) [8]C.int32_t {
	return [8]C.int32_t{
		C.int32_t(si.UserProfileRange[0]), C.int32_t(si.UserProfileRange[1]),
		C.int32_t(si.UserBehaviorRange[0]), C.int32_t(si.UserBehaviorRange[1]),
		C.int32_t(si.ItemFeatureRange[0]), C.int32_t(si.ItemFeatureRange[1]),
		C.int32_t(si.CtxFeatureRange[0]), C.int32_t(si.CtxFeatureRange[1]),
	}
}

// Train has model.Train's meaning (model.go:27-213) for the engine's model.
func (e *Engine) Train(numExamples, epochs, earlyStop int, si *rcmd.SampleInfo, inputs, targets []float32, xcols int) (lastCost float32, err error) {
	r := ranges(si)
	var cost C.float
	var ran C.int32_t
	rc := C.ctr_train_dense(e.h, f32(inputs), f32(targets), C.int64_t(numExamples), C.int32_t(xcols), &r[0],
		C.int32_t(epochs), C.int32_t(earlyStop), &cost, &ran)
	return float32(cost), e.err(rc)
}

// Predict has model.Predict's meaning (model.go:242-353).
func (e *Engine) Predict(numExamples int, si *rcmd.SampleInfo, inputs []float32, xcols int) ([]float32, error) {
	r := ranges(si)
	y := make([]float32, numExamples)
	rc := C.ctr_predict_dense(e.h, f32(inputs), C.int64_t(numExamples), C.int32_t(xcols), &r[0], f32(y))
	return y, e.err(rc)
}

// UploadTable puts a feature / embedding table in HBM (replaces UserFeatureCache /
// ItemFeatureCache / itemEmbeddingMap, rcmd.go:30-36).
func (e *Engine) UploadTable(which int, rows []float32, nrows, width int) error {
	return e.err(C.ctr_table_upload(e.h, C.int(which), f32(rows), C.int64_t(nrows), C.int32_t(width)))
}

// TrainStepIdx is one inner-loop iteration of model.Train (model.go:107-196) fed by row ids.
func (e *Engine) TrainStepIdx(userRow, itemRow, histRows []int32, label []float32) (cost float32, err error) {
	var st C.ctr_step_stats
	rc := C.ctr_train_step_idx(e.h, i32(userRow), i32(itemRow), i32(histRows), f32(label), C.int32_t(len(userRow)), &st)
	return float32(st.cost), e.err(rc)
}

// PredictIdx is recommend.BatchPredict → model.Predict (rcmd.go:277-337) fed by row ids.
func (e *Engine) PredictIdx(userRow, itemRow, histRows []int32) ([]float32, error) {
	y := make([]float32, len(userRow))
	rc := C.ctr_predict_idx(e.h, i32(userRow), i32(itemRow), i32(histRows), C.int64_t(len(userRow)), f32(y))
	return y, e.err(rc)
}

// TrainIdx is one model.Train pass (one epoch) over n samples given as row ids in host memory: batches are
// copied H2D on a second stream while the previous batch computes; costs receives one BCE per batch.
func (e *Engine) TrainIdx(userRow, itemRow, histRows []int32, label []float32, costs []float32) error {
	var cp *C.float
	if len(costs) > 0 {
		cp = f32(costs)
	}
	return e.err(C.ctr_train_idx(e.h, i32(userRow), i32(itemRow), i32(histRows), f32(label), C.int64_t(len(userRow)), cp))
}

func i64(p []int64) *C.int64_t { return (*C.int64_t)(unsafe.Pointer(&p[0])) }

// LoadIDMaps replaces the strconv.Itoa-keyed caches (rcmd.go:472,483,502): userIds[r] / itemIds[r] is the
// external id of table row r.
func (e *Engine) LoadIDMaps(userIds, itemIds []int64) error {
	if err := e.err(C.ctr_idmap_build(e.h, C.CTR_IDMAP_USER, i64(userIds), C.int64_t(len(userIds)))); err != nil {
		return err
	}
	return e.err(C.ctr_idmap_build(e.h, C.CTR_IDMAP_ITEM, i64(itemIds), C.int64_t(len(itemIds))))
}

// UploadUserBehavior puts every user's (time-descending) behaviour sequence in HBM (feature/ubcache,
// prepare.go:13-38): offsets [nUsers+1], ts and itemRows [n].
func (e *Engine) UploadUserBehavior(offsets, ts []int64, itemRows []int32) error {
	return e.err(C.ctr_ubcache_upload(e.h, i64(offsets), i64(ts), i32(itemRows), C.int64_t(len(offsets)-1), C.int64_t(len(ts))))
}

// BatchPredict is recommend.BatchPredict (rcmd.go:277-337) over sample keys, entirely on the device.
func (e *Engine) BatchPredict(sampleKeys []rcmd.Sample) ([]float32, error) {
	n := len(sampleKeys)
	u, it, ts := make([]int64, n), make([]int64, n), make([]int64, n)
	for i, k := range sampleKeys {
		u[i], it[i], ts[i] = int64(k.UserId), int64(k.ItemId), k.Timestamp
	}
	y := make([]float32, n)
	rc := C.ctr_batch_predict_keys(e.h, i64(u), i64(it), i64(ts), C.int64_t(n), f32(y))
	return y, e.err(rc)
}

// SaveCheckpoint / LoadCheckpoint: binary snapshot of weights, Adam moments, step counter and tables.
func (e *Engine) SaveCheckpoint(path string) error {
	cs := C.CString(path)
	defer C.free(unsafe.Pointer(cs))
	return e.err(C.ctr_checkpoint_save(e.h, cs))
}
func (e *Engine) LoadCheckpoint(path string) error {
	cs := C.CString(path)
	defer C.free(unsafe.Pointer(cs))
	return e.err(C.ctr_checkpoint_load(e.h, cs))
}

// dinModel is the reference's JSON schema (din.go:41-52; dnn.go:38-47 without att0).
type dinModel struct {
	UProfileDim   int       `json:"uProfileDim"`
	UBehaviorSize int       `json:"uBehaviorSize"`
	UBehaviorDim  int       `json:"uBehaviorDim"`
	IFeatureDim   int       `json:"iFeatureDim"`
	CFeatureDim   int       `json:"cFeatureDim"`
	Mlp0          []float32 `json:"mlp0"`
	Mlp1          []float32 `json:"mlp1"`
	Mlp2          []float32 `json:"mlp2"`
	Att0          []float32 `json:"att0,omitempty"`
}

// Marshal emits exactly what DinNet.Marshal / YoutubeDnn.Marshal emit (din.go:62, dnn.go:49), so
// din.NewDinNetFromJson can load an engine-trained model and vice versa.
func (e *Engine) Marshal() ([]byte, error) {
	in := int(e.cfg.uP + 2*e.cfg.D + e.cfg.cF)
	m := dinModel{UProfileDim: int(e.cfg.uP), UBehaviorSize: int(e.cfg.S), UBehaviorDim: int(e.cfg.D),
		IFeatureDim: int(e.cfg.D), CFeatureDim: int(e.cfg.cF),
		Mlp0: make([]float32, in*int(e.cfg.H0)), Mlp1: make([]float32, int(e.cfg.H0*e.cfg.H1)),
		Mlp2: make([]float32, int(e.cfg.H1)), Att0: make([]float32, int(e.cfg.S))}
	if err := e.err(C.ctr_get_weights(e.h, f32(m.Mlp0), f32(m.Mlp1), f32(m.Mlp2), f32(m.Att0))); err != nil {
		return nil, err
	}
	if Kind(e.cfg.model) == YoutubeDnn {
		m.Att0 = nil
	}
	return json.Marshal(m)
}

// LoadJSON is NewDinNetFromJson / NewYoutubeDnnFromJson for the engine (din.go:82, dnn.go:63).
func (e *Engine) LoadJSON(data []byte) error {
	var m dinModel
	if err := json.Unmarshal(data, &m); err != nil {
		return err
	}
	var att *C.float
	if len(m.Att0) > 0 {
		att = f32(m.Att0)
	}
	return e.err(C.ctr_set_weights(e.h, f32(m.Mlp0), f32(m.Mlp1), f32(m.Mlp2), att))
}

// Impl is the drop-in for example/movielens dinImpl / YoutubeDnnImpl (dinimpl.go:13-92,
// youtube.go:13-91): it satisfies recommend.Fitter and recommend.PredictAbstract, so
// recommend.Train(ctx, recSys, &ctrb200.Impl{...}) (rcmd.go:196,229) works unchanged.
type Impl struct {
	Kind                               Kind
	PredBatchSize, BatchSize, Epochs   int
	EarlyStop                          int
	UBehaviorSize, UBehaviorDim        int // rcmd.UserBehaviorLen, rcmd.ItemEmbDim
	sampleInfo                         *rcmd.SampleInfo
	xcols                              int
	pred                               *Engine
}

// Fit implements recommend.Fitter (rcmd.go:95-97); body follows dinImpl.Fit (dinimpl.go:44-92).
func (d *Impl) Fit(trainSample *rcmd.TrainSample) (rcmd.PredictAbstract, error) {
	uP := trainSample.Info.UserProfileRange[1] - trainSample.Info.UserProfileRange[0]
	cF := trainSample.Info.CtxFeatureRange[1] - trainSample.Info.CtxFeatureRange[0]
	d.sampleInfo, d.xcols = &trainSample.Info, trainSample.XCols
	if trainSample.Rows != len(trainSample.Y) {
		return nil, fmt.Errorf("number of examples %d and labels %d do not match", trainSample.Rows, len(trainSample.Y))
	}
	learner, err := NewEngine(d.Kind, uP, d.UBehaviorSize, d.UBehaviorDim, d.UBehaviorDim, cF, d.BatchSize, d.BatchSize)
	if err != nil {
		return nil, err
	}
	defer learner.Close()
	if _, err = learner.Train(trainSample.Rows, d.Epochs, d.EarlyStop, d.sampleInfo, trainSample.X, trainSample.Y, trainSample.XCols); err != nil {
		return nil, err
	}
	js, err := learner.Marshal() // learner → JSON → predictor, dinimpl.go:73-89
	if err != nil {
		return nil, err
	}
	pred, err := NewEngine(d.Kind, uP, d.UBehaviorSize, d.UBehaviorDim, d.UBehaviorDim, cF, d.PredBatchSize, d.PredBatchSize)
	if err != nil {
		return nil, err
	}
	if err = pred.LoadJSON(js); err != nil {
		return nil, err
	}
	d.pred = pred
	return d, nil
}

// Predict implements recommend.PredictAbstract (rcmd.go:87-89); body follows dinImpl.Predict
// (dinimpl.go:32-42), including returning nil on failure.
func (d *Impl) Predict(X tensor.Tensor) tensor.Tensor {
	numPred := X.Shape()[0]
	y, err := d.pred.Predict(numPred, d.sampleInfo, X.Data().([]float32), d.xcols)
	if err != nil {
		return nil
	}
	return tensor.NewDense(tensor.Float32, tensor.Shape{numPred, 1}, tensor.WithBacking(y))
}

// TrainKeys is recommend.Train's sample path (rcmd.go:197-246) on the device: the samples go down as keys (28 bytes
// each), GetSample's assembly (id maps, drop-unknown, history window at the sample's timestamp; rcmd.go:339-460) and
// model.Train's epoch loop run in HBM.  Needs LoadIDMaps (and UploadUserBehavior for DIN's history).
func (e *Engine) TrainKeys(samples []rcmd.Sample, epochs, earlyStop int) (lastCost float32, rowsUsed int64, err error) {
	n := len(samples)
	u, it, ts, y := make([]int64, n), make([]int64, n), make([]int64, n), make([]float32, n)
	for i, s := range samples {
		u[i], it[i], ts[i], y[i] = int64(s.UserId), int64(s.ItemId), s.Timestamp, s.Label
	}
	var cost C.float
	var ran C.int32_t
	var used C.int64_t
	rc := C.ctr_train_keys(e.h, i64(u), i64(it), i64(ts), f32(y), C.int64_t(n), C.int32_t(epochs), C.int32_t(earlyStop), &cost, &ran, &used)
	return float32(cost), int64(used), e.err(rc)
}

// SetTableOptimizer selects how embedding rows learn (engine extension; the reference keeps them frozen,
// din.go:161-169): C.CTR_TABLE_FROZEN | SGD | SGD_DETERMINISTIC | ADAM.  Call before NewEngine's ctr_create in a
// custom constructor, or use NewEngineWith.
func NewEngineWith(kind Kind, uP, S, D, cF, batchSize, predBatchSize, tableOpt int, tableLr float32) (*Engine, error) {
	e := &Engine{}
	C.ctr_config_default(&e.cfg, C.int(kind))
	e.cfg.uP, e.cfg.S, e.cfg.D, e.cfg.cF = C.int32_t(uP), C.int32_t(S), C.int32_t(D), C.int32_t(cF)
	e.cfg.batch, e.cfg.pred_batch = C.int32_t(batchSize), C.int32_t(predBatchSize)
	e.cfg.table_opt, e.cfg.table_lr = C.int32_t(tableOpt), C.float(tableLr)
	if rc := C.ctr_create(&e.cfg, &e.h); rc != C.CTR_OK {
		return nil, fmt.Errorf("ctr_create: %d: %s", int(rc), C.GoString(C.ctr_last_error(nil)))
	}
	runtime.SetFinalizer(e, func(e *Engine) { e.Close() })
	return e, nil
}

// ---- model/mlp: drop-in for mlp.SimpleMlpFitWrap / SimpleMlpPredWrap (model/mlp/mlp.go:11-65) -----------------

// MlpFitWrap satisfies recommend.Fitter like mlp.SimpleMlpFitWrap{Model: nn.NewMLPClassifier([]int{100}, "relu",
// "adam", 1e-4)} (main.go:42-52): float32 TrainSample in, float64 training on the device.
type MlpFitWrap struct {
	Hidden      []int   // HiddenLayerSizes, default {100}
	Activation  int     // C.CTR_MLP_RELU (default) | CTR_MLP_LOGISTIC | CTR_MLP_IDENTITY
	Alpha       float64 // 1e-4
	MaxIter     int     // 200
	Adaptive    bool    // LearningRate == "adaptive"
	Seed        uint32
}

// MlpPredWrap satisfies recommend.PredictAbstract like mlp.SimpleMlpPredWrap.
type MlpPredWrap struct {
	m     *C.ctr_mlp
	xcols int
}

func (f *MlpFitWrap) Fit(trainSample *rcmd.TrainSample) (rcmd.PredictAbstract, error) {
	var cfg C.ctr_mlp_config
	C.ctr_mlp_config_default(&cfg, C.int32_t(trainSample.XCols))
	if len(f.Hidden) > 0 {
		cfg.n_layers = C.int32_t(len(f.Hidden) + 2)
		for i, u := range f.Hidden {
			cfg.units[i+1] = C.int32_t(u)
		}
		cfg.units[len(f.Hidden)+1] = 1
	}
	cfg.hidden_act = C.int32_t(f.Activation)
	if f.Alpha != 0 {
		cfg.alpha = C.double(f.Alpha)
	}
	if f.MaxIter != 0 {
		cfg.max_iter = C.int32_t(f.MaxIter)
	}
	if f.Adaptive {
		cfg.adaptive = 1
	}
	cfg.seed = C.uint32_t(f.Seed)
	var m *C.ctr_mlp
	if rc := C.ctr_mlp_create(&cfg, &m); rc != C.CTR_OK {
		return nil, fmt.Errorf("ctr_mlp_create: %d: %s", int(rc), C.GoString(C.ctr_mlp_last_error(nil)))
	}
	var iters C.int32_t
	if rc := C.ctr_mlp_fit(m, f32(trainSample.X), f32(trainSample.Y), C.int64_t(trainSample.Rows), C.int32_t(trainSample.XCols), &iters, nil); rc != C.CTR_OK {
		err := fmt.Errorf("ctr_mlp_fit: %d: %s", int(rc), C.GoString(C.ctr_mlp_last_error(m)))
		C.ctr_mlp_destroy(m)
		return nil, err
	}
	p := &MlpPredWrap{m: m, xcols: trainSample.XCols}
	runtime.SetFinalizer(p, func(p *MlpPredWrap) { C.ctr_mlp_destroy(p.m) })
	return p, nil
}

// Predict follows SimpleMlpPredWrap.Predict (mlp.go:15-39), including nil on failure.
func (p *MlpPredWrap) Predict(X tensor.Tensor) tensor.Tensor {
	n := X.Shape()[0]
	y := make([]float32, n)
	if rc := C.ctr_mlp_predict(p.m, f32(X.Data().([]float32)), C.int64_t(n), C.int32_t(p.xcols), f32(y)); rc != C.CTR_OK {
		return nil
	}
	return tensor.NewDense(tensor.Float32, tensor.Shape{n, 1}, tensor.WithBacking(y))
}
