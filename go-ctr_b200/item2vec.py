"""Host-side mirror of feature/embedding for the device trainer (include/ctr_b200.h ctr_i2v_train).

    embedding.TrainEmbedding(inputCh, window, dim, iter)        wordemb.go:9-32
    model.GenEmbeddingMap32() map[string][]float32              word2vec.go:298-324

The channel carries one WORD per message (cpsutil.ReadWord, cpsutil.go:46-53); ids are assigned by first
appearance (dictionary.Add, dictionary.go:70-81)."""
import ctypes as C

import numpy as np

from . import engine as _e

__all__ = ["i2v_paths", "I2vConfig", "I2vStats", "i2v_default_config", "i2v_train_ids", "TrainEmbedding", "EmbeddingModel"]


class I2vConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("window", C.c_int32), ("iter", C.c_int32), ("min_count", C.c_int32), ("max_depth", C.c_int32),
                ("init_lr", C.c_float), ("min_lr", C.c_float), ("subsample", C.c_float), ("update_lr_batch", C.c_int32),
                ("seed", C.c_uint32), ("device", C.c_int32), ("reserved", C.c_int32 * 8)]


class I2vStats(C.Structure):
    _fields_ = [("doc_len", C.c_int64), ("trained_positions", C.c_int64), ("pairs", C.c_int64), ("node_visits", C.c_int64),
                ("algorithmic_bytes", C.c_double), ("ms_device", C.c_float), ("launches", C.c_int32)]


def i2v_default_config(**kw):
    cfg = I2vConfig()
    _e.load_library().ctr_i2v_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def i2v_paths(count, max_depth=100):
    """Host-side Huffman paths of the trainer (no GPU needed): returns (path_off [V+1], nodes, codes)."""
    L = _e.load_library()
    cnt = np.ascontiguousarray(count, np.int64); V = cnt.size
    cap = V * min(max_depth, 64) + 64
    off = np.empty(V + 1, np.int64); nodes = np.empty(cap, np.int32); codes = np.empty(cap, np.uint8)
    rc = L.ctr_i2v_paths(cnt.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int32(V), C.c_int32(max_depth), off.ctypes.data_as(C.POINTER(C.c_int64)),
                         nodes.ctypes.data_as(C.POINTER(C.c_int32)), codes.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(cap))
    if rc != 0:
        raise _e.CtrError(rc, L.ctr_last_error(None).decode())
    return off, nodes[:off[-1]], codes[:off[-1]]


def i2v_train_ids(tokens, vocab, cfg=None, **kw):
    """tokens: int32 word ids in corpus order.  Returns (emb [vocab, dim] float32, I2vStats)."""
    L = _e.load_library()
    cfg = cfg or i2v_default_config(**kw)
    tok = np.ascontiguousarray(tokens, np.int32)
    emb = np.empty((vocab, cfg.dim), np.float32)
    st = I2vStats()
    rc = L.ctr_i2v_train(C.byref(cfg), tok.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(tok.size), C.c_int32(vocab),
                         emb.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st))
    if rc != 0:
        raise _e.CtrError(rc, L.ctr_last_error(None).decode())
    return emb, st


class EmbeddingModel:
    def __init__(self, id2word, emb, stats):
        self.id2word, self.emb, self.stats = id2word, emb, stats

    def GenEmbeddingMap32(self):
        if len(self.id2word) == 0:
            raise ValueError("dictionary is empty")                          # word2vec.go:305-308
        return {w: self.emb[i] for i, w in enumerate(self.id2word)}


def TrainEmbedding(inputCh, window, dim, iter, **kw):
    """inputCh: iterable of words (item ids as strings)."""
    word2id, id2word, toks = {}, [], []
    for w in inputCh:
        i = word2id.get(w)
        if i is None:
            i = word2id[w] = len(id2word); id2word.append(w)
        toks.append(i)
    emb, st = i2v_train_ids(np.asarray(toks, np.int32), len(id2word), window=window, dim=dim, iter=iter, **kw)
    return EmbeddingModel(id2word, emb, st)
