import numpy as np


def close_frac(a, b, rtol, atol):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.mean(np.abs(a - b) <= atol + rtol * np.abs(b)))


def assert_mostly_close(a, b, rtol, atol, frac=0.999, what=""):
    """Adam turns a gradient whose sign flips inside fp32 rounding noise into a ±lr weight step, so a
    handful of elements may legitimately differ after optimiser steps; everything else must agree."""
    f = close_frac(a, b, rtol, atol)
    assert f >= frac, "%s: only %.5f of elements within rtol=%g atol=%g" % (what, f, rtol, atol)


def make_tables(rng, U, I, uP, cF, D):
    uf = rng.random((U, uP), np.float32)
    itf = rng.random((I, cF), np.float32)
    emb = (rng.standard_normal((I, D)) / np.sqrt(D)).astype(np.float32)
    return uf, itf, emb


def make_batch(rng, U, I, B, S, pad_frac=0.2, zipf=False):
    ur = rng.integers(0, U, B).astype(np.int32)
    if zipf:
        ir = (rng.zipf(1.05, B) - 1) % I
        hist = (rng.zipf(1.05, (B, S)) - 1) % I
    else:
        ir = rng.integers(0, I, B)
        hist = rng.integers(0, I, (B, S))
    hist = hist.astype(np.int32); ir = ir.astype(np.int32)
    # most-recent-first with a -1 padded tail (prepare.go:49-51, rcmd.go:517-522)
    lens = np.where(rng.random(B) < pad_frac, rng.integers(0, S + 1, B), S)
    hist[np.arange(S)[None, :] >= lens[:, None]] = -1
    y = (rng.random(B) > 0.5).astype(np.float32)
    return ur, ir, hist, y


def scaled_init(orc, ocfg, seed, s0=0.05, s1=0.05, s2=0.3):
    """N(0,1) init (din.go:187-191) scaled so sigmoids are not saturated and parity is informative."""
    W0, W1, W2, att = orc.init_weights(ocfg, seed)
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray(W0 * s0), np.ascontiguousarray(W1 * s1), np.ascontiguousarray(W2 * s2),
            (att + 0.2 * rng.standard_normal(att.shape)).astype(np.float32)]
