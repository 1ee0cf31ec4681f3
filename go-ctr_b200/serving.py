"""Host mirror of the serving side of the boundary: recommend.Rank /
recommend.BatchPredict (rcmd.go:248-337) over sample keys, with the id→row maps, the ubcache window and
the forward pass all on the device (ctr_batch_predict_keys).  Names and argument order follow the
reference; `recSys` is an Engine that already holds the tables, both id maps and (optionally) the ubcache.
"""
import time
from dataclasses import dataclass

import numpy as np

from . import engine as _e


@dataclass
class Sample:                      # rcmd.go:189-194
    UserId: int
    ItemId: int
    Timestamp: int = 0
    Label: float = 0.0


@dataclass
class ItemScore:                   # rcmd.go:184-187
    ItemId: int
    Score: float


def BatchPredict(recSys, sampleKeys):
    """rcmd.go:277-337.  Returns float32 [n, 1] like the tensor the reference hands back.  An unknown
    user/item in key 0 raises (CtrError code ENOTFOUND), in later keys it scores as a zero X row."""
    if len(sampleKeys) == 0:
        raise ValueError("no sample keys")
    u = np.fromiter((s.UserId for s in sampleKeys), np.int64, len(sampleKeys))
    i = np.fromiter((s.ItemId for s in sampleKeys), np.int64, len(sampleKeys))
    t = np.fromiter((s.Timestamp for s in sampleKeys), np.int64, len(sampleKeys))
    return recSys.batch_predict_keys(u, i, t).reshape(-1, 1)


def Rank(recSys, userId, itemIds, now=None):
    """rcmd.go:248-275: score `itemIds` for one user at the current time; order is the caller's."""
    ts = int(time.time()) if now is None else int(now)
    y = BatchPredict(recSys, [Sample(userId, it, ts) for it in itemIds])
    return [ItemScore(int(it), float(y[k, 0])) for k, it in enumerate(itemIds)]


def load_id_maps(recSys, user_ids, item_ids):
    """user_ids[r] / item_ids[r] = external id of table row r (what the Go side's string-keyed caches hold,
    rcmd.go:472-505)."""
    recSys.idmap_build(_e.IDMAP_USER, user_ids)
    recSys.idmap_build(_e.IDMAP_ITEM, item_ids)
