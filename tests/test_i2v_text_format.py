"""The word2vec text format (f3): the reference's own vectors of emb/embedding_test.go:30-100 and embutil_test.go,
and a save → load round trip in vector.Save's `%f ` layout (model/modelutil/vector/vector.go:40-67)."""
import io

import numpy as np
import pytest

import go_ctr_b200 as g


def test_load_reference_vectors():
    contents = "apple 1 1 1 1 1\n\t\t\tbanana 1 1 1 1 1\n\t\t\tchocolate 0 0 0 0 0\n\t\t\tdragon -1 -1 -1 -1 -1"   # embedding_test.go:36-41
    embs = g.LoadVectors(contents)
    assert len(embs) == 4 and [e.Word for e in embs] == ["apple", "banana", "chocolate", "dragon"]
    assert embs[3].Vector.tolist() == [-1.0] * 5


def test_parse_line_and_norm():
    e = g.ParseLine("apple 1 1 1 1 1")                                    # embedding_test.go:78-90
    assert e == g.Embedding("apple", [1, 1, 1, 1, 1]) and e.Dim == 5 and e.Norm == np.sqrt(5.0)
    assert g.Embedding("x", [1, 1, 1, 1, 0, 0]).Norm == 2.0               # embutil_test.go: norm
    with pytest.raises(ValueError):
        g.ParseLine("lonely")
    assert len(g.LoadVectors(" skipped 1 2 3\nkept 1 2 3")) == 1           # embedding.go:94-96: leading space = skipped


def test_save_layout_and_round_trip():
    rng = np.random.default_rng(0)
    mat = rng.standard_normal((5, 4)).astype(np.float32)
    words = ["17", "4", "900", "x", "y"]
    f = io.StringIO()
    g.SaveVectors(f, words, mat)
    text = f.getvalue()
    assert text.splitlines()[0] == "17 " + "".join("%f " % float(v) for v in mat[0])
    back = g.LoadVectors(text)
    assert [e.Word for e in back] == words
    np.testing.assert_allclose(np.stack([e.Vector for e in back]), mat, atol=5e-7)     # %f keeps six decimals
    with pytest.raises(ValueError, match="different"):
        g.SaveVectors(io.StringIO(), words[:3], mat)
