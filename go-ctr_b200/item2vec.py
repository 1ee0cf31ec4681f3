"""Host-side mirror of feature/embedding for the device trainer (include/ctr_b200.h ctr_i2v_train).

    embedding.TrainEmbedding(inputCh, window, dim, iter)        wordemb.go:9-32
    model.GenEmbeddingMap32() map[string][]float32              word2vec.go:298-324

    vector.Save(f, dic, mat, ...)  word2vec text format           model/modelutil/vector/vector.go:40-67
    emb.Load(r) / parse / parseLine → []Embedding{Word,Dim,Vector,Norm}   emb/embedding.go:73-131

The channel carries one WORD per message (cpsutil.ReadWord, cpsutil.go:46-53); ids are assigned by first
appearance (dictionary.Add, dictionary.go:70-81)."""
import ctypes as C

import numpy as np

from . import engine as _e

__all__ = ["i2v_paths", "I2vConfig", "I2vStats", "i2v_default_config", "i2v_train_ids", "i2v_train_dist", "TrainEmbedding", "EmbeddingModel",
           "Embedding", "SaveVectors", "LoadVectors", "ParseLine"]


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


def i2v_train_dist(tokens_shard, vocab, rank, world, nccl_id, sync_every=0, cfg=None, want_table=True, **kw):
    """One rank of the multi-GPU trainer (ctr_i2v_train_dist): this rank's shard of the stream in, the averaged table out
    (want_table=False: this rank skips the copy of the table to the host and returns None for it)."""
    L = _e.load_library()
    cfg = cfg or i2v_default_config(**kw)
    tok = np.ascontiguousarray(tokens_shard, np.int32)
    emb = np.empty((vocab, cfg.dim), np.float32) if want_table else None
    st = I2vStats()
    rc = L.ctr_i2v_train_dist(C.byref(cfg), tok.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(tok.size), C.c_int32(vocab),
                              emb.ctypes.data_as(C.POINTER(C.c_float)) if want_table else None, C.byref(st), C.c_int32(rank), C.c_int32(world),
                              C.c_char_p(nccl_id), C.c_int32(len(nccl_id)), C.c_int64(sync_every))
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


# ---- the word2vec text format the reference saves / loads its vectors in -------------------------------------
class Embedding:
    """emb.Embedding (emb/embedding.go:27-32): Word, Dim, Vector (float64), Norm = sqrt(sum v^2) (embutil.Norm)."""

    def __init__(self, Word, Vector):
        self.Word = Word
        self.Vector = np.asarray(Vector, np.float64)
        self.Dim = int(self.Vector.size)
        n = 0.0
        for v in self.Vector:                      # embutil/embutil.go: sequential float64 sum
            n += float(v) * float(v)
        self.Norm = float(np.sqrt(n))

    def Validate(self):                            # embedding.go:34-43
        if self.Word == "":
            raise ValueError("word must not be empty")
        if self.Dim == 0:
            raise ValueError("Dim of %s is zero" % self.Word)

    def __eq__(self, o):
        return (self.Word, self.Dim, self.Norm) == (o.Word, o.Dim, o.Norm) and np.array_equal(self.Vector, o.Vector)


def ParseLine(line):
    """emb.parseLine (embedding.go:107-131): whitespace-separated fields, the first is the word."""
    fields = line.split()
    if len(fields) < 2:
        raise ValueError("Must be over 2 lenghth for word and vector elems")
    return Embedding(fields[0], [float(x) for x in fields[1:]])


def LoadVectors(r):
    """emb.Load (embedding.go:73-105): one embedding per line; lines that START WITH A SPACE are skipped; every vector
    must be non-empty.  r: iterable of lines, a file object, or one string."""
    lines = r.splitlines() if isinstance(r, str) else r
    out = []
    for line in lines:
        line = line.rstrip("\r\n")
        if line.startswith(" "):
            continue
        e = ParseLine(line)
        e.Validate()
        out.append(e)
    if out and any(e.Dim != out[0].Dim for e in out):      # Embeddings.Validate, :61-71
        raise ValueError("dimension for all vectors must be the same")
    return out


def SaveVectors(f, words, mat):
    """vector.Save (vector.go:40-67): per word `word ` then every element as `%f ` (six decimals, trailing space), newline."""
    mat = np.asarray(mat)
    if len(words) != mat.shape[0]:
        raise ValueError("different for length of dic and row of matrix: %d, %d" % (len(words), mat.shape[0]))
    for w, row in zip(words, mat):
        f.write("%s " % w + "".join("%f " % float(v) for v in row) + "\n")
