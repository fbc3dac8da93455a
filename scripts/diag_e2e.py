#!/usr/bin/env python3
"""diagnostics: where does the end-to-end (host buffer) path lose time against the device-resident
one?  cfg2 geometry with fewer rows (same bytes per k-mer), host-side time per call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

rows = int(os.environ.get("ROWS", "1000003"))
nq = int(os.environ.get("NQ", "10000"))
thr = float(os.environ.get("THR", "0.1"))
g = cobs_b200.GpuIndex.procedural(0, 100_000, [rows], 3, fill_seed=bench.FILL_SEED)
g.set_option("max_batch", nq)
batches = [bench.make_batch(1000 + i, nq) for i in range(24)]
pinned = [torch.from_numpy(b).pin_memory().numpy() for b, _ in batches]
off = batches[0][1]

for depth in (1, 2, 3, 4):
    for timing in (0, 1):
        g.set_option("timing", timing)
        g.timers(reset=True)
        t_sub, t_col = 0.0, 0.0

        def run(idx):
            global t_sub, t_col
            pend = []
            for i in idx:
                a = time.perf_counter()
                pend.append(g.submit(pinned[i], off, thr, 0))
                b = time.perf_counter()
                t_sub += b - a
                if len(pend) == depth:
                    a = time.perf_counter()
                    g.collect(pend.pop(0), raw=True)
                    t_col += time.perf_counter() - a
            while pend:
                a = time.perf_counter()
                g.collect(pend.pop(0), raw=True)
                t_col += time.perf_counter() - a
        run(range(4))
        torch.cuda.synchronize()
        t_sub = t_col = 0.0
        g.timers(reset=True)
        t0 = time.perf_counter()
        run(range(4, 24))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = g.timers()
        print("depth %d timing %d: %.3f ms/step  host submit %.3f collect %.3f ms/step  phases %s" % (
            depth, timing, 1e3 * dt / 20, 1e3 * t_sub / 20, 1e3 * t_col / 20,
            {k: round(tm[k] / 20, 3) for k in ("hashes_ms", "score_ms", "select_ms", "h2d_ms", "d2h_ms")}
            if timing else ""), flush=True)

# single-query latency breakdown
q = pinned[0][:100]
o1 = off[:2]
for thr1, lim in ((0.8, 0), (0.0, 10)):
    for timing in (0, 1):
        g.set_option("timing", timing)
        for i in range(50):
            g.search_packed(q, o1, thr1, lim, raw=True)
        g.timers(reset=True)
        t0 = time.perf_counter()
        for i in range(500):
            g.search_packed(q, o1, thr1, lim, raw=True)
        dt = (time.perf_counter() - t0) / 500
        tm = g.timers()
        print("single query thr %.1f limit %d timing %d: %.1f us/call %s" % (
            thr1, lim, timing, 1e6 * dt,
            {k: round(1e3 * tm[k] / 500, 1) for k in ("hashes_ms", "score_ms", "select_ms", "h2d_ms", "d2h_ms")}
            if timing else ""), flush=True)
g.close()
