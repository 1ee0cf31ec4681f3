"""Host-side pieces of bench.py's CPU arm (no GPU): the usable-thread probe, the thread-count picker and a tiny
`--impl reference`-style step through oracle/cpu_fast.c."""
import os

import numpy as np

import bench


def test_host_threads_is_within_the_affinity_mask():
    n = bench.host_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0))


def test_pick_threads_keeps_the_fastest_candidate_and_reports_every_timing():
    calls = []

    def step(th):                       # pretend 8 threads are the sweet spot (oversubscription beyond)
        calls.append(th)
        return {64: 0.9, 32: 0.5, 16: 0.3, 8: 0.3}.get(th, 1.0) + (0.0 if th != 16 else -0.05)
    best, tried = bench.pick_threads(step, 64)
    assert best == 16
    assert set(tried) == {"64", "32", "16"} and all(v > 0 for v in tried.values())
    assert calls.count(64) == 2 and calls.count(16) == 2          # one warm + one timed step per candidate


def test_pick_threads_on_a_small_box_only_tries_counts_it_has():
    best, tried = bench.pick_threads(lambda th: 1.0 / th, 6)
    assert best == 6 and set(tried) == {"6", "3", "1"}


def test_cpu_arm_runs_the_fast_oracle_on_the_workload_shape():
    w = dict(bench.WORKLOADS["din_100m"]); w["I"] = 20_000
    state = {}
    v, n, dt, rows = bench.cpu_arm(w, 0.0, 512, 2, min_steps=1, state=state)
    assert n == 1 and v > 0 and rows == 20_000
    v2, n2, _, _ = bench.cpu_arm(w, 0.0, 512, 1, min_steps=2, state=state)      # the trainer is reused across thread counts
    assert n2 == 2 and state["Bc"] == 512
