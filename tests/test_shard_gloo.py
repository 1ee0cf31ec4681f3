"""world_size-2 gloo test (CPU) of the row-sharded placement's data movement (tests/shard_model.py): sharded gather == direct gather, and the
gradient return leg == a single-process scatter-add over the concatenated batches."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import shard_model as sh


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # same table on every rank
    I, D, B, S = 101, 8, 17, 5
    table = rng.standard_normal((I, D)).astype(np.float32)
    brng = np.random.default_rng(1000 + rank)              # each rank its own batch
    hist = brng.integers(-1, I, (B, S)); item = brng.integers(0, I, B)
    shard = sh.shard_table(table, rank, world)
    rows_local, slot, plan = sh.fetch_rows(dist, torch, shard, hist, item, world)
    # forward leg: every lookup sees exactly the global row (bit-exact), missing rows have no slot
    lookups = np.concatenate([hist, item[:, None]], 1)
    for b in range(B):
        for s in range(S + 1):
            if lookups[b, s] < 0:
                assert slot[b, s] == -1
            else:
                assert rows_local[slot[b, s]].tobytes() == table[lookups[b, s]].tobytes()
    # backward leg: per-lookup gradients go home and accumulate on the owner
    grad = brng.standard_normal(rows_local.shape).astype(np.float32)
    new_shard = sh.return_grads(dist, torch, shard.copy(), grad, plan)
    np.save(os.path.join(out_dir, "shard%d.npy" % rank), new_shard)
    np.save(os.path.join(out_dir, "look%d.npy" % rank), lookups)
    np.save(os.path.join(out_dir, "grad%d.npy" % rank), grad)
    np.save(os.path.join(out_dir, "slot%d.npy" % rank), slot)
    dist.barrier()
    dist.destroy_process_group()


def test_plan_exchange_buckets_by_owner():
    hist = np.array([[4, -1, 7], [2, 2, 9]]); item = np.array([5, 0])
    send_rows, cnt, slot = sh.plan_exchange(hist, item, 2)
    assert cnt.tolist() == [4, 3]                          # owner 0: rows 4,2,2,0 ; owner 1: rows 7,5,9
    assert send_rows[:4].tolist() == [2, 1, 1, 0] and send_rows[4:].tolist() == [3, 2, 4]
    assert slot[0, 1] == -1 and sorted(slot[slot >= 0].tolist()) == list(range(7))


@pytest.mark.timeout(120)
def test_sharded_exchange_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    table = rng.standard_normal((101, 8)).astype(np.float32)
    want = table.astype(np.float64).copy()
    for r in range(world):
        look = np.load(tmp_path / ("look%d.npy" % r)); grad = np.load(tmp_path / ("grad%d.npy" % r)); slot = np.load(tmp_path / ("slot%d.npy" % r))
        for b in range(look.shape[0]):
            for s in range(look.shape[1]):
                if look[b, s] >= 0:
                    want[look[b, s]] += grad[slot[b, s]]
    got = np.empty_like(table)
    for r in range(world):
        got[r::world] = np.load(tmp_path / ("shard%d.npy" % r))
    np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-5, atol=1e-5)


# ---- replicated placement (tables <= 32 MB): local gather, one all-reduce, identical update on every rank -------------
def _repl_setup():
    from oracle import oracle as orc
    from tests.util import make_batch, make_tables, scaled_init
    uP, S, D, cF, U, I, B = 6, 4, 8, 5, 23, 31, 24
    rng = np.random.default_rng(5)
    tabs = make_tables(rng, U, I, uP, cF, D)
    cfg = orc.make_cfg(orc.DIN_COS, uP, S, D, cF, 200, 80)                # no dropout
    W = scaled_init(orc, cfg, 2)
    batch = make_batch(rng, U, I, 2 * B, S, zipf=True)
    return orc, cfg, W, tabs, batch, (uP, S, D, cF, U, I, B)


def _repl_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc, cfg, W, (uf, itf, emb), (ur, ir, hist, y), (uP, S, D, cF, U, I, B) = _repl_setup()
    sl = slice(rank * B, (rank + 1) * B)
    table_lr, lr, l2 = 0.3, 0.01, 1e-4
    W = [np.ascontiguousarray(w, np.float32).copy() for w in W]; emb = emb.copy()
    # local forward / backward on this rank's half of the global batch (local-mean gradients)
    X = orc.gather_rows(uf, itf, emb, ur[sl], ir[sl], hist[sl])
    ws = orc.Workspace(cfg, B)
    orc.forward(cfg, W, X, orc.make_ranges(uP, S, D, cF), ws=ws)
    g = orc.backward(cfg, W, ws, y[sl])
    # row gradients, pre-scaled by -lr/world, summed into a table-shaped buffer (k_attn_bwd_idx FUSED → comm.table_grad)
    tg = np.zeros_like(emb)
    for b in range(B):
        for s in range(S):
            if hist[sl][b, s] >= 0:
                tg[hist[sl][b, s]] += (-table_lr / world) * g["dUb"][b, s]
        if ir[sl][b] >= 0:
            tg[ir[sl][b]] += (-table_lr / world) * g["dIt"][b]
    dense = [g["dW0"], g["dW1"], g["dW2"], g["datt"]]
    cost = sh.allreduce_replicated(dist, torch, dense, tg, g["cost"] * B)
    emb += tg                                                             # k_apply_table_grad
    mv = [np.zeros_like(w) for w in W for _ in (0, 1)]
    for i, (w, gr) in enumerate(zip(W, dense)):                           # Adam on the summed gradients: gscale = 1/world, batch = world*B
        gr = np.ascontiguousarray(gr.reshape(w.shape) / world, np.float32)
        orc.adam_step(w, gr, mv[2 * i], mv[2 * i + 1], 1, lr=lr, l2=l2, batch=float(world * B))
    np.save(os.path.join(out_dir, "emb%d.npy" % rank), emb)
    np.save(os.path.join(out_dir, "cost%d.npy" % rank), np.array([cost / (world * B)]))
    for i, w in enumerate(W):
        np.save(os.path.join(out_dir, "w%d_%d.npy" % (i, rank)), w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_replicated_placement_two_ranks_gloo(tmp_path):
    """Two ranks with the whole table each + one all-reduce == one process training on the global batch."""
    world = 2
    mp.spawn(_repl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    orc, cfg, W, (uf, itf, emb), (ur, ir, hist, y), (uP, S, D, cF, U, I, B) = _repl_setup()
    ref = orc.IdxTrainer(cfg, orc.default_solver(), W, uf, itf, emb)
    cost, _ = ref.step(ur, ir, hist, y, table_lr=0.3)
    e0, e1 = np.load(tmp_path / "emb0.npy"), np.load(tmp_path / "emb1.npy")
    assert e0.tobytes() == e1.tobytes()                                   # identical replicas after the step
    assert np.abs(ref.emb - emb).max() > 1e-4
    np.testing.assert_allclose(e0, ref.emb, rtol=1e-4, atol=1e-6)
    assert abs(float(np.load(tmp_path / "cost0.npy")[0]) - cost) < 1e-5
    for i, w in enumerate(ref.W):
        w0 = np.load(tmp_path / ("w%d_0.npy" % i))
        assert w0.tobytes() == np.load(tmp_path / ("w%d_1.npy" % i)).tobytes()
        np.testing.assert_allclose(w0.reshape(w.shape), w, rtol=1e-3, atol=2e-5)
