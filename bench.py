#!/usr/bin/env python
"""bench.py — one "step" = one pass of the CTR hot path (gather → attention → MLP → BCE → backward →
scatter-add + SGD(rows) + Adam(dense)) over one batch of synthetic MovieLens-shaped input.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Prints ONE JSON line (see the task contract).  `value` = whole-job train samples/s with inputs
resident in HBM; `e2e` = same metric through the C-ABI entry point ctr_train_step_idx with pinned
HOST buffers (H2D of the indices/labels and D2H of the cost inside the timed region);
`roofline` = the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM
peak; `cpu_baseline` = the CPU oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: DIN on synthetic MovieLens-20M-shaped batches (138k users, 27k items, dim 64)
    "din_ml20m": dict(model="din", U=138493, I=27278, D=64, S=50, uP=52, cF=53, B=65536, zipf=True,
                      note="BASELINE.json configs[1]; item table 7 MB is L2-resident (reported, not an HBM reading)"),
    # BASELINE.json configs[2]: YouTube DNN, 10M-item table dim 64, batch 16384
    "youtube_10m": dict(model="youtube", U=138493, I=10_000_000, D=64, S=50, uP=52, cF=53, B=16384, zipf=False,
                        note="BASELINE.json configs[2]"),
    # BASELINE.json configs[3] per-GPU shard: DIN, 100M rows / 8 GPUs = 12.5M rows (3.2 GB), batch 65536, S=50
    # BASELINE.json configs[3] itself: 100M rows (25.6 GB) row-sharded over the GPUs, global batch 65536 at 8 GPUs
    "din_100m": dict(model="din", U=138493, I=100_000_000, D=64, S=50, uP=52, cF=53, B=8192, zipf=False,
                     note="BASELINE.json configs[3]: 100M-row table sharded row%world, 8192 samples per GPU (65536 global at 8 GPUs), uniform ids"),
    "din_100m_shard": dict(model="din", U=138493, I=12_500_000, D=64, S=50, uP=52, cF=53, B=65536, zipf=False,
                           note="BASELINE.json configs[3], one GPU's 1/8 row shard; uniform indices (worst case for caches)"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi polled every 10 ms from before the warm-up; stop() keeps the samples whose timestamp falls inside
    [t0, t1] = the timed region plus a short untimed tail of the same load (the timed region alone is often shorter
    than nvidia-smi's start-up + polling latency)."""
    Q = "timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "10"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[2]), float(c[3]), c[5:9]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 is None or (t0 - 0.005 <= r[0] <= t1 + 0.005)]
        sm, mx, reasons = [], [], set()
        for ts, a, m, flags in inside:
            sm.append(a); mx.append(m)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       window="timed region + untimed tail of the same steps, %.2f s" % (t1 - t0) if t0 is not None else "whole run")
        return out


def synth_batch(w, rng, B):
    from tests.util import make_batch
    return make_batch(rng, w["U"], w["I"], B, w["S"], pad_frac=0.2, zipf=w["zipf"])


def algorithmic_bytes(w, hist, ir):
    """fp32 table, int32 ids, duplicates counted, caches ignored; only rows that
    exist (hist >= 0) are counted."""
    B = hist.shape[0]
    rows = int((hist >= 0).sum()) + int((ir >= 0).sum())
    row_bytes = w["D"] * 4
    gather = rows * row_bytes + B * (w["S"] + 2) * 4 + B * (w["uP"] + w["cF"]) * 4
    scatter = 2 * rows * row_bytes
    return gather, scatter, rows


def run_reference(args, w, wname):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Go
    (gorgonia) and cannot be built here (no Go toolchain, deps not vendored), so this times the
    C port of its semantics (oracle/) with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    rng = np.random.default_rng(42)
    Bc = args.cpu_batch
    wc = dict(w); wc["I"] = min(w["I"], 2_000_000)            # the port allocates a dense f64 accumulator per step
    model = orc.DIN_COS if w["model"] == "din" else orc.YOUTUBE
    ocfg = orc.make_cfg(model, w["uP"], w["S"], w["D"], w["cF"], 200, 80, 0.005, 0.005)
    uf = rng.random((wc["U"], w["uP"]), dtype=np.float32); itf = rng.random((wc["I"], w["cF"]), dtype=np.float32)
    emb = (rng.standard_normal((wc["I"], w["D"]), dtype=np.float32) / np.sqrt(w["D"])).astype(np.float32)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(0), orc.init_weights(ocfg, 0), uf, itf, emb)
    from tests.util import make_batch
    batches = [make_batch(rng, wc["U"], wc["I"], Bc, w["S"], zipf=w["zipf"]) for _ in range(2)]
    cores = os.cpu_count() or 1
    for i in range(args.warmup):
        tr.step(*batches[i % 2], table_lr=0.05, nthreads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        tr.step(*batches[i % 2], table_lr=0.05, nthreads=cores)
    dt = time.perf_counter() - t0
    v = Bc * args.steps / dt
    line = {"impl": "reference", "metric": "ctr_train_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wname, "note": w["note"], "graph": w["model"], "users": w["U"], "items": w["I"], "D": w["D"], "S": w["S"],
                       "uP": w["uP"], "cF": w["cF"], "per_gpu_batch": w["B"], "global_batch": w["B"] * args.gpus,
                       "ids": "zipf(1.05)" if w["zipf"] else "uniform", "history_padding": "20% of samples have a -1 padded tail",
                       "sample": "each step = %d samples of this workload on the host cores (bounded CPU sample; item table capped at %d rows)" % (Bc, wc["I"])},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "%d steps x %d samples, OpenMP %d threads, C port of go-ctr semantics (Go reference unbuildable here)" % (args.steps, Bc, cores)},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(w, seconds=12.0, Bc=16384):
    from oracle import oracle as orc
    from tests.util import make_batch
    rng = np.random.default_rng(43)
    I = min(w["I"], 1_000_000)
    model = orc.DIN_COS if w["model"] == "din" else orc.YOUTUBE
    ocfg = orc.make_cfg(model, w["uP"], w["S"], w["D"], w["cF"], 200, 80, 0.005, 0.005)
    uf = rng.random((w["U"], w["uP"]), dtype=np.float32); itf = rng.random((I, w["cF"]), dtype=np.float32)
    emb = (rng.standard_normal((I, w["D"]), dtype=np.float32) / np.sqrt(w["D"])).astype(np.float32)
    tr = orc.IdxTrainer(ocfg, orc.default_solver(0), orc.init_weights(ocfg, 0), uf, itf, emb)
    batch = make_batch(rng, w["U"], I, Bc, w["S"], zipf=w["zipf"])
    cores = os.cpu_count() or 1
    tr.step(*batch, table_lr=0.05, nthreads=cores)           # warm
    n = 0; t0 = time.perf_counter()
    while True:
        tr.step(*batch, table_lr=0.05, nthreads=cores); n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds and n >= 2:
            break
    return {"value": Bc * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d steps x %d samples of the same workload shape in %.1f s, OpenMP %d threads" % (n, Bc, dt, cores)}


def run_item2vec(args):
    """BASELINE config 5 (item2vec SkipGram-HS, window 5): a step = one pass of ctr_i2v_train over a bounded
    synthetic token stream.  value = trained-document tokens / device time of the training kernels;
    e2e = stream tokens / wall time of the whole call (host dictionary + Huffman build, H2D, training, D2H
    of the embedding table).  The only throughput the reference publishes is for this loop: 555k words/s
    on an Apple M1 Max (README.md:140)."""
    import go_ctr_b200 as g
    from oracle import oracle as orc
    V, n, D = args.i2v_vocab, args.i2v_tokens, args.i2v_dim
    rng = np.random.default_rng(42)
    # Zipf(1.0)-like item popularity over V items, ids by first appearance are not required by the engine
    raw = (np.floor(np.exp(rng.random(n) * np.log(V))).astype(np.int64) - 1).clip(0, V - 1)
    uniq, toks = np.unique(raw, return_inverse=True)              # the dictionary holds only words that occur (dictionary.go:70-81)
    toks = toks.astype(np.int32); V = int(uniq.size)
    hbm_peak, peak_src = load_peaks()
    if args.impl == "reference":
        m = min(n, args.i2v_cpu_tokens)
        cfg = orc.i2v_cfg(dim=D, window=5, iters=1, seed=1, rng_mode=0)
        t0 = time.perf_counter(); emb, trained = orc.i2v_train(cfg, toks[:m], V); dt = time.perf_counter() - t0
        v = m / dt
        print(json.dumps({"impl": "reference", "metric": "item2vec_words_per_sec", "value": v, "unit": "words/s", "n_gpus": 1, "steps": 1, "warmup": 0,
                          "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": v / 555000.0, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "item2vec", "vocab": V, "dim": D, "window": 5, "sample": "%d tokens, single thread" % m},
                          "cpu_baseline": {"value": v, "unit": "words/s", "cores": 1, "kind": "port", "sample": "%d tokens of the same stream" % m},
                          "e2e": {"value": v, "unit": "words/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    best = None
    for i in range(max(1, args.warmup if args.warmup < 3 else 1) + max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        emb, st = g.i2v_train_ids(toks, V, dim=D, window=5, iter=1, seed=1)
        wall = time.perf_counter() - t0
        if i >= 1 and (best is None or st.ms_device < best[0].ms_device):
            best = (st, wall)
    st, wall = best
    value = st.doc_len / (st.ms_device * 1e-3)
    ach = st.algorithmic_bytes / (st.ms_device * 1e-3) / 1e9
    line = {"metric": "item2vec_words_per_sec", "value": value, "unit": "words/s", "n_gpus": 1, "steps": 1, "warmup": 1, "ms_per_step": st.ms_device,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / 555000.0, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "item2vec", "note": "BASELINE.json configs[4] on a bounded %d-token sample of the stream, 1 GPU" % n, "vocab": V, "dim": D, "window": 5,
                       "optimizer": "hierarchical softmax", "ids": "zipf(1.0)", "l2": "vector tables (%.0f MB) exceed L2; no flush" % (2.0 * V * D * 4 / 1e6),
                       "vs_baseline_note": "published 555k words/s is MovieLens-10M on an Apple M1 Max (README.md:140), different data and dim"},
            "e2e": {"value": n / wall, "unit": "words/s", "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": int(V * D * 4)},
            "gpu_launches": int(st.launches),
            "roofline": {"bound": "hbm", "kernel": "k_i2v_skipgram_hs", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": st.algorithmic_bytes, "ms_per_launch": st.ms_device,
                         "pairs": int(st.pairs), "node_visits": int(st.node_visits),
                         "note": "Zipf item popularity: most node/context rows are served by L2 and the top Huffman nodes by shared memory "
                                 "(ncu: DRAM traffic << algorithmic bytes, profiles/r01/ncu_item2vec_v2.md) - this is SURVEY 8(d)'s algorithmic-bytes figure, "
                                 "not a DRAM reading; the kernel is L2-latency / issue bound"},
            "stats": {"doc_len": int(st.doc_len), "trained_positions": int(st.trained_positions)}}
    if not args.no_cpu_baseline:
        m = min(n, args.i2v_cpu_tokens)
        cfg = orc.i2v_cfg(dim=D, window=5, iters=1, seed=1, rng_mode=0)
        t0 = time.perf_counter(); orc.i2v_train(cfg, toks[:m], V); dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "words/s", "cores": 1, "kind": "port", "sample": "%d tokens of the same stream, single-threaded float64 port" % m}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="din_ml20m", choices=list(WORKLOADS) + ["item2vec"])
    ap.add_argument("--i2v-vocab", type=int, default=10_000_000)
    ap.add_argument("--i2v-tokens", type=int, default=50_000_000)
    ap.add_argument("--i2v-dim", type=int, default=64)
    ap.add_argument("--i2v-cpu-tokens", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=16384)
    ap.add_argument("--table-opt", default="sgd", choices=["sgd", "det", "frozen"])
    ap.add_argument("--gemm", default="auto", choices=["auto", "fp32", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-leg", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wname = args.workload
    if wname == "item2vec":
        return run_item2vec(args)
    w = dict(WORKLOADS[wname])
    if args.batch:
        w["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, w, wname)

    import torch
    import go_ctr_b200 as g
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node N" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = load_peaks()

    def build_engine(w, B, placement=0):
        model = g.MODEL_DIN_COS if w["model"] == "din" else g.MODEL_YOUTUBE
        topt = {"sgd": g.TABLE_SGD, "det": g.TABLE_SGD_DETERMINISTIC, "frozen": g.TABLE_FROZEN}[args.table_opt]
        gm = {"auto": g.GEMM_AUTO, "fp32": g.GEMM_FP32, "tcgen05": g.GEMM_TCGEN05_3XTF32}[args.gemm]
        cfg = g.engine.default_config(model, uP=w["uP"], S=w["S"], D=w["D"], cF=w["cF"], batch=B, pred_batch=B,
                                      table_opt=topt, table_lr=0.05, gemm=gm, device=local, rank=rank, world=world, seed=1)
        cfg.reserved[1] = placement          # ITEM_EMB under world > 1: 0 = by size, 1 = row-sharded, 2 = replicated
        eng = g.Engine(cfg)
        if world > 1:
            ids = [None]
            if rank == 0:
                ids[0] = eng.comm_unique_id()
            dist.broadcast_object_list(ids, src=0)
            eng.comm_init(ids[0])
        # synthetic MovieLens-shaped tables generated on the device: features U[0,1), embeddings N(0, 1/D)
        eng.table_fill(g.TABLE_USER_FEAT, w["U"], w["uP"], seed=3, dist=0, scale=1.0)
        eng.table_fill(g.TABLE_ITEM_FEAT, w["I"], w["cF"], seed=4, dist=0, scale=1.0)
        eng.table_fill(g.TABLE_ITEM_EMB, w["I"], w["D"], seed=5, dist=1, scale=float(1.0 / np.sqrt(w["D"])))
        return eng

    def make_batches(w, B, nb, seed):
        rng = np.random.default_rng(seed + rank)
        out = []
        for _ in range(nb):
            out.append(synth_batch(w, rng, B))
        return out

    dev = torch.device("cuda", local)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def timed_leg(eng, w, B, steps, warmup, profile=False, clocks=True):
        """K timed steps, inputs resident in HBM, L2 flushed before every step (outside the events)."""
        host = make_batches(w, B, 4, 100)
        devb = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host]
        st = torch.cuda.ExternalStream(eng.stream, device=dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = eng.launch_count()

        def one(i, e=None):
            ur, ir, hist, y = devb[i % len(devb)]
            with torch.cuda.stream(st):
                flush_buf.zero_()                                           # L2 flush (write > L2 capacity)
                if e: e[0].record(st)
                eng.train_step_idx_dev(ur.data_ptr(), ir.data_ptr(), hist.data_ptr(), y.data_ptr(), B)
                if e: e[1].record(st)
        sampler = ClockSampler(local) if (clocks and not profile) else None      # polls from before the warm-up
        for i in range(warmup):
            one(i)
        eng.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l1 = eng.launch_count()
        if profile:
            eng.profile_reset(); eng.profile(True)
        t0 = time.perf_counter(); c0 = time.time()
        for i in range(steps):
            one(warmup + i, ev[i])
        eng.sync(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l2 = eng.launch_count()
        clocks = None
        if sampler:      # untimed tail: the same steps keep the load up until nvidia-smi has had >= 0.3 s to sample it
            k = 0
            # multi-rank steps contain collectives: every rank must run the same number of tail steps
            while (k < 300) if world > 1 else (time.time() - c0 < 0.3 and k < 400):
                one(k); k += 1
                if k % 8 == 0:
                    eng.sync()
            eng.sync(); torch.cuda.synchronize()
            clocks = sampler.stop(c0, time.time())
        if profile:
            eng.profile(False)
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = (l2 - l1)        # kernels launched inside the timed region only
        cost = eng.last_cost()
        return dict(ms=ms, wall=wall, clocks=clocks, launches=launches, host=host, cost=cost, prof=eng.profile_dump() if profile else None)

    def e2e_leg(eng, w, B, steps, warmup):
        """Same metric through the public host-buffer entry point ctr_train_idx (one model.Train pass over
        `steps` batches held in pinned HOST memory): every batch's ids/labels are copied H2D inside the call
        (overlapped with the previous batch's compute on a second stream) and every batch's cost is read
        back D2H before the call returns.  Single-GPU; with sharded tables (world > 1) the per-batch entry
        point ctr_train_step_idx is timed instead."""
        host = make_batches(w, B, 4, 200)
        if world > 1 and w["I"] * w["D"] * 4 > 32 * 2**20:     # row-sharded tables: per-batch entry point
            pinned = [tuple(torch.from_numpy(a).pin_memory() for a in b) for b in host]
            st = g.StepStats()
            for i in range(warmup):
                ur, ir, hist, y = pinned[i % 4]
                eng.train_step_idx_ptr(ur.data_ptr(), ir.data_ptr(), hist.data_ptr(), y.data_ptr(), B, st)
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                ur, ir, hist, y = pinned[(warmup + i) % 4]
                eng.train_step_idx_ptr(ur.data_ptr(), ir.data_ptr(), hist.data_ptr(), y.data_ptr(), B, st)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item()); last = st.cost
            h2d = sum(a.numel() * a.element_size() for a in pinned[0])
        else:
            cat = [torch.from_numpy(np.concatenate([host[i % 4][j] for i in range(steps)])).pin_memory() for j in range(4)]
            costs = torch.empty(steps, dtype=torch.float32).pin_memory()
            eng.train_idx_ptr(cat[0].data_ptr(), cat[1].data_ptr(), cat[2].data_ptr(), cat[3].data_ptr(), B * min(steps, max(warmup, 2)), costs.data_ptr())
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.train_idx_ptr(cat[0].data_ptr(), cat[1].data_ptr(), cat[2].data_ptr(), cat[3].data_ptr(), B * steps, costs.data_ptr())
            dt = time.perf_counter() - t0
            if world > 1:       # slowest rank
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            last = float(costs[-1])
            h2d = sum(a.numel() * a.element_size() for a in cat) // steps
        return dict(value=B * world * steps / dt, unit="samples/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=8,
                    ms_per_step=1e3 * dt / steps, last_cost=last)

    def roofline_of(w, leg, prof, B, workload_name=None):
        ur, ir, hist, y = leg["host"][0]
        gbytes, sbytes, rows = algorithmic_bytes(w, hist, ir)
        kern = {}
        for name, (ms, n) in prof.items():
            kern[name] = {"ms_per_launch": ms / max(n, 1), "launches_per_step": n / max(args.steps, 1)}
        fwd = next((k for k in kern if k.startswith("attn_fwd")), None)
        bwd = next((k for k in kern if k.startswith("attn_bwd")), None)
        if fwd:
            kern[fwd]["algorithmic_bytes"] = gbytes
            kern[fwd]["gbs"] = gbytes / (kern[fwd]["ms_per_launch"] * 1e-3) / 1e9
        if bwd:
            kern[bwd]["algorithmic_bytes"] = sbytes
            kern[bwd]["gbs"] = sbytes / (kern[bwd]["ms_per_launch"] * 1e-3) / 1e9
        cand = [k for k in (fwd, bwd) if k]
        if not cand:
            return None, kern
        dom = max(cand, key=lambda k: kern[k]["ms_per_launch"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(workload_name or wname, {}).get(dom)
        pair_ms = sum(kern[k]["ms_per_launch"] for k in cand)
        rl = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
              "frac": kern[dom]["gbs"] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
              "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"], "ms_per_launch": kern[dom]["ms_per_launch"],
              "fused_pair": {"kernels": cand, "algorithmic_bytes": gbytes + sbytes, "ms": pair_ms,
                             "achieved": (gbytes + sbytes) / (pair_ms * 1e-3) / 1e9,
                             "frac": (gbytes + sbytes) / (pair_ms * 1e-3) / 1e9 / hbm_peak},
              "rows_per_launch": rows,
              "how": "second pass of the same %d steps with every launch bracketed by CUDA events on the engine stream" % args.steps}
        return rl, kern

    B = w["B"]
    eng = build_engine(w, B)
    leg = timed_leg(eng, w, B, args.steps, args.warmup)
    prof_leg = timed_leg(eng, w, B, args.steps, 1, profile=True)
    rl, kern = roofline_of(w, prof_leg, prof_leg["prof"], B)
    e2e = e2e_leg(eng, w, B, max(3, min(args.steps, 64)), 2)
    value = B * world * args.steps / (leg["ms"] * 1e-3)
    step_ms = sum(v["ms_per_launch"] * v["launches_per_step"] for v in kern.values())
    for v in kern.values():
        v["share_of_step"] = v["ms_per_launch"] * v["launches_per_step"] / step_ms if step_ms else None
    line = {"metric": "ctr_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": leg["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wname, "note": w["note"], "graph": w["model"], "users": w["U"], "items": w["I"], "D": w["D"], "S": w["S"],
                       "uP": w["uP"], "cF": w["cF"], "per_gpu_batch": B, "global_batch": B * world, "table_opt": args.table_opt,
                       "gemm": args.gemm, "ids": "zipf(1.05)" if w["zipf"] else "uniform", "history_padding": "20% of samples have a -1 padded tail",
                       "l2": "256 MiB buffer written before every timed step (outside the event pair); per-step event pairs are summed",
                       "placement": ("single GPU" if world == 1 else
                                       "1 process per GPU; ITEM_EMB (%.0f MB) replicated, row + dense gradients all-reduced (NCCL)" % (w["I"] * w["D"] * 4 / 1e6)
                                       if w["I"] * w["D"] * 4 <= 32 * 2**20 else "1 process per GPU; ITEM_EMB rows sharded row%world, de-duplicated all-to-all exchange (NCCL)")},
            "clocks": leg["clocks"], "e2e": {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "e2e_ms_per_step": e2e["ms_per_step"], "gpu_launches": int(leg["launches"]), "last_cost": leg["cost"],
            "wall_s_timed_region": leg["wall"], "roofline": rl, "kernels": kern}
    if rl is not None and w["I"] * w["D"] * 4 / (1 if w["I"] * w["D"] * 4 <= 32 * 2**20 else world) < 126e6:
        rl["note"] = "table is L2-resident in the timed workload (DRAM traffic << algorithmic bytes): not an HBM reading"
    if world > 1 and w["I"] * w["D"] * 4 <= 32 * 2**20 and not args.no_hbm_leg:
        # the timed workload's table is small enough to be replicated; the same steps with the table forced onto
        # the row-sharded NCCL all-to-all path (BASELINE north_star's placement for large tables) — all ranks, collective
        del eng
        torch.cuda.empty_cache()
        eng_s = build_engine(w, B, placement=1)
        ks = max(5, args.steps // 4)
        leg_s = timed_leg(eng_s, w, B, ks, 3, clocks=False)
        if rank == 0:
            line["sharded_exchange"] = {"what": "same workload, ITEM_EMB forced row-sharded (row % world) with the de-duplicated all-to-all exchange",
                                        "value": B * world * ks / (leg_s["ms"] * 1e-3), "unit": "samples/s", "ms_per_step": leg_s["ms"] / ks}
        del eng_s
    if rank == 0 and world == 1 and not args.no_hbm_leg and wname != "din_100m_shard":
        # HBM-fair reading of the same kernels: a table far larger than L2
        del eng
        torch.cuda.empty_cache()
        w2 = dict(WORKLOADS["din_100m_shard"])
        eng2 = build_engine(w2, w2["B"])
        leg2 = timed_leg(eng2, w2, w2["B"], max(5, args.steps // 2), 3)
        p2 = timed_leg(eng2, w2, w2["B"], max(5, args.steps // 2), 1, profile=True)
        steps_saved = args.steps; args.steps = max(5, args.steps // 2)
        rl2, kern2 = roofline_of(w2, p2, p2["prof"], w2["B"], "din_100m_shard")
        args.steps = steps_saved
        line["hbm_roofline"] = {"workload": "din_100m_shard", "note": w2["note"], "items": w2["I"], "per_gpu_batch": w2["B"],
                                "samples_per_sec": w2["B"] * max(5, steps_saved // 2) / (leg2["ms"] * 1e-3),
                                "ms_per_step": leg2["ms"] / max(5, steps_saved // 2), "roofline": rl2,
                                "kernels": {k: v for k, v in kern2.items() if k.startswith("attn")}}
        del eng2
        if w["I"] * w["D"] * 4 < 126e6:
            # the timed workload's table fits the 126 MB L2, so its own GB/s is an L2 reading; the graded HBM roofline is
            # the same kernels, same batch shape, on the table that does not fit (measured just above, in this run)
            line["roofline_timed_workload"] = rl
            line["roofline"] = dict(rl2, workload="din_100m_shard",
                                    note="dominant kernel on a 12.5 M-row (3.2 GB) table, uniform ids; measured in this run after the timed region")
    if rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
