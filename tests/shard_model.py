"""Semantic model of the row-sharded placement (test infrastructure; SURVEY.md §8e, go-ctr_b200/csrc/comm.cuh):
ITEM_EMB row r lives on rank r % world at local row r // world; every lookup of a rank's batch is served from the OWNER's
shard, every row gradient (already scaled by -lr/world) is added into the OWNER's shard.  On the GPUs the kernels do this
with peer loads / red.add over NVLink mappings, without any exchange phase; here the same data movement is spelled out as
torch.distributed all-to-alls so that it runs over gloo on CPU with world_size 2 (tests/test_shard_gloo.py): what the
test pins is the owner mapping and the equivalence "sharded gather / scatter == the single-table gather / scatter-add".

The reference has no distributed code at all."""
import numpy as np


def owner(row, world):
    return row % world


def local_row(row, world):
    return row // world


def shard_table(table, rank, world):
    """Rows this rank owns, in local-row order (ctr_table_upload with world > 1)."""
    return np.ascontiguousarray(table[rank::world])


def plan_exchange(hist, item_row, world):
    """Lookup p = b*(S+1)+slot (slot S = the target item).  Returns
    send_rows [n_valid]  owner-local row ids, grouped by owner (ascending owner),
    counts    [world]    lookups per owner,
    slot      [B, S+1]   position of each lookup in send order, -1 for a missing row."""
    hist = np.asarray(hist); item_row = np.asarray(item_row)
    B, S = hist.shape
    rows = np.concatenate([hist, item_row[:, None]], axis=1).reshape(-1)
    valid = rows >= 0
    own = np.where(valid, rows % world, world)
    order = np.argsort(own, kind="stable")[: int(valid.sum())]
    counts = np.bincount(own[valid], minlength=world)[:world]
    slot = np.full(B * (S + 1), -1, np.int64)
    slot[order] = np.arange(order.size)
    send_rows = (rows[order] // world).astype(np.int32)
    return send_rows, counts.astype(np.int64), slot.reshape(B, S + 1)


def _all_to_all(dist, torch, send, scounts, rcounts, width=None):
    s = torch.from_numpy(np.ascontiguousarray(send))
    shape = (int(rcounts.sum()),) + tuple(s.shape[1:])
    r = torch.empty(shape, dtype=s.dtype)
    dist.all_to_all_single(r, s, output_split_sizes=[int(c) for c in rcounts], input_split_sizes=[int(c) for c in scounts])
    return r.numpy()


def fetch_rows(dist, torch, shard, hist, item_row, world):
    """Forward leg: returns (rows_local [n_valid, D] in send order, slot [B,S+1], plan) so that
    rows_local[slot[b, s]] is the embedding row of lookup (b, s)."""
    send_rows, scnt, slot = plan_exchange(hist, item_row, world)
    rcnt_t = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(rcnt_t, torch.from_numpy(scnt))
    rcnt = rcnt_t.numpy()
    recv_rows = _all_to_all(dist, torch, send_rows, scnt, rcnt)
    gathered = shard[recv_rows]                                   # the owner's rows (GPU: ld.global from the peer mapping)
    rows_local = _all_to_all(dist, torch, gathered, rcnt, scnt)
    return rows_local, slot, (send_rows, scnt, rcnt, recv_rows)


def return_grads(dist, torch, shard, grad_local, plan):
    """Backward leg: grad_local [n_valid, D] (already scaled by -lr/world) is added into the owners' shards
    (GPU: red.global.add.v4.f32 into the peer mapping)."""
    send_rows, scnt, rcnt, recv_rows = plan
    g = _all_to_all(dist, torch, grad_local, scnt, rcnt)
    np.add.at(shard, recv_rows, g)
    return shard


def allreduce_replicated(dist, torch, dense_grads, table_grad, cost_sum):
    """The replicated placement's one collective (comm_allreduce_grads with the table-shaped buffer, comm_impl.cuh; the
    sharded placement all-reduces its replicated hot rows the same way):
    dense gradients, the row-gradient buffer [I, D] and the cost sum, summed over the ranks, in place."""
    bufs = [torch.from_numpy(g) for g in dense_grads] + [torch.from_numpy(table_grad)]
    for t in bufs:
        dist.all_reduce(t)
    c = torch.tensor([cost_sum], dtype=torch.float64)
    dist.all_reduce(c)
    return float(c.item())
