"""Per-kernel device times of the row-sharded train step (ITEM_EMB / ITEM_FEAT sharded row % world, peers read and
updated over NVLink).  Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P
tests/tools/shard_time.py [rows] [per_gpu_batch] [zipf]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import go_ctr_b200 as g          # noqa: E402
from tests.util import make_batch          # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    I = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    zipf = len(sys.argv) > 3 and sys.argv[3] == "zipf"
    U, uP, S, D, cF = 138_493, 52, 50, 64, 53
    cfg = g.engine.default_config(g.MODEL_DIN_COS, uP=uP, S=S, D=D, cF=cF, batch=B, pred_batch=B, table_opt=g.TABLE_SGD, table_lr=0.01,
                                  device=local, rank=rank, world=world)
    cfg.reserved[1] = 1
    eng = g.Engine(cfg)
    ids = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(ids[0])
    eng.table_fill(g.TABLE_USER_FEAT, U, uP, 1, 0, 1.0); eng.table_fill(g.TABLE_ITEM_FEAT, I, cF, 2, 0, 1.0)
    eng.table_fill(g.TABLE_ITEM_EMB, I, D, 3, 1, 0.125)
    rng = np.random.default_rng(rank)
    batches = [make_batch(rng, U, I, B, S, pad_frac=0.2, zipf=zipf) for _ in range(4)]
    dev = torch.device("cuda", local)
    devb = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in batches]
    st = torch.cuda.ExternalStream(eng.stream, device=dev)

    def step(i):
        ur, ir, hist, y = devb[i % 4]
        eng.train_step_idx_dev(ur.data_ptr(), ir.data_ptr(), hist.data_ptr(), y.data_ptr(), B)
    for i in range(4):
        step(i)
    eng.sync(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    K = 20
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(K):
            step(i)
        e1.record(st)
    eng.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    cost = eng.last_cost()
    dist.barrier()
    eng.profile(True); eng.profile_reset()
    for i in range(8):
        step(i)
    eng.sync()
    prof = {k: round(v / n, 4) for k, (v, n) in eng.profile_dump().items()}
    eng.profile(False)
    rows = sum(int((b[2] >= 0).sum()) + B for b in batches) / 4
    remote = sum(int(((b[2] >= 0) & (b[2] % world != rank)).sum()) + int((b[1] % world != rank).sum()) for b in batches) / 4
    out = {"rank": rank, "world": world, "rows": I, "B": B, "zipf": zipf, "ms_per_step": round(ms, 4), "samples_per_s_global": round(B * world / ms * 1e3),
           "cost": cost, "ms": prof, "nvlink_MB_each_way": round(remote * D * 4 / 1e6, 1),
           "fwd_nvlink_GBs": round(remote * D * 4 / prof.get("attn_fwd_peer", 1) / 1e6), "bwd_nvlink_GBs": round(remote * D * 4 / prof.get("attn_bwd_peer", 1) / 1e6)}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
