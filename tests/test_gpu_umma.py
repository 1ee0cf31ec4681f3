"""The tcgen05 (3xTF32, TMEM-accumulated) GEMM engine against the oracle and against the exact-fp32
FFMA engine: same parity bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import go_ctr_b200 as g
from oracle import oracle as orc
from tests.test_gpu_parity import GRAD_ATOL, SCORE_RTOL, SHAPES, setup
from tests.util import assert_mostly_close, make_batch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
UM = dict(gemm=g.GEMM_TCGEN05_3XTF32)


@pytest.mark.parametrize("B", [128, 257, 1000])
@pytest.mark.parametrize("shape", ["ns", "ref", "odd"])
def test_umma_forward_scores(shape, B):
    eng, cfg, ocfg, W, (uf, itf, emb), (ur, ir, hist, y) = setup(g.MODEL_DIN_COS, shape, B, **UM)
    out = eng.debug_grads_idx(ur, ir, hist, y, training=False)
    X = orc.gather_rows(uf, itf, emb, ur, ir, hist)
    o0 = orc.make_cfg(orc.DIN_COS, ocfg.uP, ocfg.S, ocfg.D, ocfg.cF, 200, 80)
    p, z = orc.forward(o0, W, X, orc.make_ranges(ocfg.uP, ocfg.S, ocfg.D, ocfg.cF))
    np.testing.assert_allclose(out["logit"], z, rtol=SCORE_RTOL, atol=2e-5)
    np.testing.assert_allclose(out["p"], p, rtol=SCORE_RTOL, atol=1e-7)
    assert "umma_fwd0_sigmoid" not in eng.profile_dump()      # profiling is off by default


@pytest.mark.parametrize("model", [g.MODEL_DIN_COS, g.MODEL_YOUTUBE])
def test_umma_gradients_match_oracle_and_fp32_engine(model):
    eng, cfg, ocfg, W, (uf, itf, emb), (ur, ir, hist, y) = setup(model, "ns", 384, seed=3, **UM)
    ref_eng, *_ = setup(model, "ns", 384, seed=3, gemm=g.GEMM_FP32)
    out = eng.debug_grads_idx(ur, ir, hist, y, training=False)
    fp32 = ref_eng.debug_grads_idx(ur, ir, hist, y, training=False)
    X = orc.gather_rows(uf, itf, emb, ur, ir, hist)
    o0 = orc.make_cfg(model, ocfg.uP, ocfg.S, ocfg.D, ocfg.cF, 200, 80)
    ws = orc.Workspace(o0, len(y))
    orc.forward(o0, W, X, orc.make_ranges(ocfg.uP, ocfg.S, ocfg.D, ocfg.cF), ws=ws)
    ref = orc.backward(o0, W, ws, y)
    for k in ("dW0", "dW1", "dW2", "dIt"):
        scale = np.abs(ref[k]).max()
        np.testing.assert_allclose(out[k], ref[k], rtol=2e-3, atol=GRAD_ATOL * scale + 1e-12, err_msg=k)
        np.testing.assert_allclose(out[k], fp32[k], rtol=2e-3, atol=GRAD_ATOL * scale + 1e-12, err_msg=k + " vs fp32 engine")
    valid = hist >= 0
    np.testing.assert_allclose(out["dUb"][valid], ref["dUb"][valid], rtol=2e-3, atol=2e-5 * np.abs(ref["dUb"]).max() + 1e-12)


def test_umma_dropout_and_train_steps():
    eng, cfg, ocfg, W, (uf, itf, emb), (ur, ir, hist, y) = setup(g.MODEL_DIN_COS, "ref", 256, seed=6, table_opt=g.TABLE_SGD_DETERMINISTIC,
                                                                 table_lr=0.7, **UM)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(seed=cfg.seed), W, uf, itf, emb)
    rng = np.random.default_rng(66)
    for step in range(4):
        ur, ir, hist, y = make_batch(rng, uf.shape[0], itf.shape[0], 256, cfg.S, zipf=True)
        st = eng.train_step_idx(ur, ir, hist, y)
        ocost, _ = tr.step(ur, ir, hist, y, table_lr=0.7)
        assert abs(st.cost - ocost) <= 2e-4 * max(1.0, abs(ocost)), (step, st.cost, ocost)
    for a, b, name in zip(eng.get_weights(), tr.W, ("mlp0", "mlp1", "mlp2", "att0")):
        assert_mostly_close(a, b, 2e-3, 2e-4, 0.995, name)
    np.testing.assert_allclose(eng.table_download(g.TABLE_ITEM_EMB, *emb.shape), tr.emb, rtol=2e-4, atol=2e-6)


def test_umma_large_batch_matches_fp32_engine():
    """B=65536: scores of the tcgen05 engine vs the exact-fp32 engine on identical state."""
    res = []
    for gm in (g.GEMM_TCGEN05_3XTF32, g.GEMM_FP32):
        eng, cfg, ocfg, W, tabs, (ur, ir, hist, y) = setup(g.MODEL_DIN_COS, "ns", 65536, seed=2, U=5000, I=40000, gemm=gm)
        res.append(eng.debug_grads_idx(ur, ir, hist, y, training=False))
    np.testing.assert_allclose(res[0]["logit"], res[1]["logit"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(res[0]["dW0"], res[1]["dW0"], rtol=1e-3, atol=1e-5 * np.abs(res[1]["dW0"]).max())
    np.testing.assert_allclose(res[0]["dIt"], res[1]["dIt"], rtol=1e-3, atol=1e-5 * np.abs(res[1]["dIt"]).max())
