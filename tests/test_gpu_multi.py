"""Two-rank test of the ITEM_EMB placements (needs >= 2 GPUs: `gpurun --gpus 2`): row-sharded with every lookup
served from the owner's HBM over NVLink peer mappings; the same with the most popular rows replicated on every rank;
replicated with gradient all-reduce.  A step on 2 GPUs with per-GPU batch B == a step on 1 GPU with batch 2B (same
samples), for scores, cost, dense weights and the updated embedding table; plus the pipelined epoch entry with a ragged
tail on the sharded path."""
import os
import socket

import numpy as np
import pytest
import torch

import go_ctr_b200 as g
from tests.util import assert_mostly_close, make_batch, make_tables

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

DIMS = dict(uP=52, S=10, D=16, cF=53)
U, I, B, STEPS = 97, 211, 256, 3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _data():
    rng = np.random.default_rng(77)
    uf, itf, emb = make_tables(rng, U, I, DIMS["uP"], DIMS["cF"], DIMS["D"])
    batches = [make_batch(rng, U, I, 2 * B, DIMS["S"], zipf=True) for _ in range(STEPS)]
    return uf, itf, emb, batches


def _worker(rank, world, port, out_dir, policy, hot):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    uf, itf, emb, batches = _data()
    cfg = g.engine.default_config(g.MODEL_DIN_COS, batch=B, pred_batch=B, table_opt=g.TABLE_SGD, table_lr=0.7, dropout0=0.0, dropout1=0.0,
                                  seed=3, device=rank, rank=rank, world=world, **DIMS)
    cfg.reserved[1] = policy                  # 1 = shard rows, 2 = replicate the table (auto would replicate: 13 KB)
    cfg.reserved[0] = hot                     # sharded: rows [0, hot) replicated on every rank (-1: none)
    eng = g.Engine(cfg)
    ids = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(ids[0])
    eng.table_upload(g.TABLE_USER_FEAT, uf); eng.table_upload(g.TABLE_ITEM_FEAT, itf); eng.table_upload(g.TABLE_ITEM_EMB, emb)
    costs = []
    sl = slice(rank * B, (rank + 1) * B)
    for k, (ur, ir, hist, y) in enumerate(batches):
        if k < STEPS - 1:
            costs.append(eng.train_step_idx(ur[sl], ir[sl], hist[sl], y[sl]).cost)
        else:           # the last batch goes through the epoch entry point (pageable buffers, pinned ring) — collective too
            costs.append(float(eng.train_idx(ur[sl], ir[sl], hist[sl], y[sl])[0]))
    ur, ir, hist, y = batches[0]
    p = eng.predict_idx(ur[sl], ir[sl], hist[sl])
    np.save(os.path.join(out_dir, "p%d.npy" % rank), p)
    np.save(os.path.join(out_dir, "cost%d.npy" % rank), np.array(costs, np.float32))
    np.save(os.path.join(out_dir, "emb%d.npy" % rank), eng.table_download(g.TABLE_ITEM_EMB, I, DIMS["D"]))
    for i, w in enumerate(eng.get_weights()):
        np.save(os.path.join(out_dir, "w%d_%d.npy" % (i, rank)), w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("policy,hot", [(1, -1), (1, 64), (2, 0)], ids=["sharded_peer", "sharded_hot_rows", "replicated"])
def test_two_gpu_step_equals_single_gpu_global_batch(tmp_path, policy, hot):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), policy, hot), nprocs=world, join=True)
    uf, itf, emb, batches = _data()
    cfg = g.engine.default_config(g.MODEL_DIN_COS, batch=2 * B, pred_batch=2 * B, table_opt=g.TABLE_SGD_DETERMINISTIC, table_lr=0.7,
                                  dropout0=0.0, dropout1=0.0, seed=3, **DIMS)
    ref = g.Engine(cfg)
    ref.table_upload(g.TABLE_USER_FEAT, uf); ref.table_upload(g.TABLE_ITEM_FEAT, itf); ref.table_upload(g.TABLE_ITEM_EMB, emb)
    want_cost = [ref.train_step_idx(*b).cost for b in batches]
    got_cost = np.load(tmp_path / "cost0.npy")
    np.testing.assert_allclose(got_cost, np.load(tmp_path / "cost1.npy"), rtol=0, atol=0)     # every rank reports the global mean
    np.testing.assert_allclose(got_cost, np.array(want_cost, np.float32), rtol=2e-4, atol=1e-6)
    for i, w in enumerate(ref.get_weights()):
        w0 = np.load(tmp_path / ("w%d_0.npy" % i))
        assert w0.tobytes() == np.load(tmp_path / ("w%d_1.npy" % i)).tobytes()               # identical Adam step on every rank
        assert_mostly_close(w0, w, 2e-3, 2e-4, 0.995, "weights %d" % i)
    got_emb = np.zeros_like(emb)
    for r in range(world):
        got_emb[r::world] = np.load(tmp_path / ("emb%d.npy" % r))[r::world]
    if policy == 2:      # every rank holds the whole, identical table
        assert np.load(tmp_path / "emb0.npy").tobytes() == np.load(tmp_path / "emb1.npy").tobytes()
        got_emb = np.load(tmp_path / "emb0.npy")
    want_emb = ref.table_download(g.TABLE_ITEM_EMB, I, DIMS["D"])
    assert np.abs(want_emb - emb).max() > 1e-5
    np.testing.assert_allclose(got_emb, want_emb, rtol=2e-3, atol=2e-5)
    p = np.concatenate([np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")])
    np.testing.assert_allclose(p, ref.predict_idx(*batches[0][:3]), rtol=2e-3, atol=1e-5)


def _i2v_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from tests.test_oracle_i2v import planted_corpus
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    toks, remap, nc, per = planted_corpus(np.random.default_rng(2))
    shard = np.array_split(toks, world)[rank]
    ids = [g.Engine(g.engine.default_config(g.MODEL_YOUTUBE, batch=1, pred_batch=1, device=rank)).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    cfg = g.i2v_default_config(dim=16, window=5, iter=3, seed=5, device=rank)
    emb, st = g.i2v_train_dist(shard, nc * per, rank, world, ids[0], sync_every=2000, cfg=cfg)
    np.save(os.path.join(out_dir, "i2v%d.npy" % rank), emb)
    np.save(os.path.join(out_dir, "i2v_doc%d.npy" % rank), np.array([st.doc_len, st.trained_positions]))
    dist.barrier(); dist.destroy_process_group()


def test_two_gpu_item2vec_replicas_average_to_one_table_of_equal_quality(tmp_path):
    """ctr_i2v_train_dist: each rank trains its half of the stream, dictionary counts are all-reduced (one Huffman tree),
    replicas are averaged every 2000 positions: both ranks return the identical table, and its neighbour structure is as
    good as the sequential float64 oracle's on the whole stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from tests.test_oracle_i2v import neighbour_purity, planted_corpus
    mp.spawn(_i2v_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    e0, e1 = np.load(tmp_path / "i2v0.npy"), np.load(tmp_path / "i2v1.npy")
    assert e0.tobytes() == e1.tobytes()
    toks, remap, nc, per = planted_corpus(np.random.default_rng(2))
    d0, d1 = np.load(tmp_path / "i2v_doc0.npy"), np.load(tmp_path / "i2v_doc1.npy")
    assert d0[0] + d1[0] == toks.size                                   # every rank filtered with the GLOBAL counts
    oemb, _ = orc.i2v_train(orc.i2v_cfg(dim=16, window=5, iters=3, seed=5, rng_mode=1), toks, nc * per)
    pg, po = neighbour_purity(e0, remap, nc, per), neighbour_purity(oemb, remap, nc, per)
    assert pg > 0.9 and abs(pg - po) < 0.08, (pg, po)
