"""world_size-2 gloo test (CPU) of the multi-GPU exchange plan: sharded gather == direct gather, and the
gradient return leg == a single-process scatter-add over the concatenated batches."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from go_ctr_b200 import shard as sh


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
