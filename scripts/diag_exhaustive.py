#!/usr/bin/env python3
"""Where the time of `benchmark-fpr`'s default pattern goes (1000-k-mer queries, threshold 0,
every document returned): wall time per call next to the device phases (K2, sort, copies), for
several batch sizes and copy granules (option pipe_kb; COBSGPU_NO_PIPE=1 = one pass)."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

n_docs, rows = 1_000_000, int(os.environ.get("ROWS", "100003"))
g = cobs_b200.GpuIndex.procedural(0, n_docs, [rows], 3, fill_seed=bench.FILL_SEED)
g.set_option("timing", 1)
for nq, qlen in ((16, 1030), (64, 1030), (64, 100)):
    blob, off = bench.make_batch(7100, nq, qlen)
    pin = torch.from_numpy(blob).pin_memory().numpy()
    for pipe_kb in (32768, 16384, 65536, 1 << 30):
        g.set_option("pipe_kb", pipe_kb)
        for _ in range(5):
            g.search_packed(pin, off, 0.0, 0, raw="view")
        g.timers(reset=True)
        reps = 4
        t0 = time.perf_counter()
        for _ in range(reps):
            roff, doc, score = g.search_packed(pin, off, 0.0, 0, raw="view")
        dt = (time.perf_counter() - t0) / reps
        t = g.timers(reset=True)
        print("nq %3d len %5d pipe_kb %10d: %.3f ms/query wall (%.1f GB/s of results) | per call: "
              "score %.3f select %.3f d2h %.3f hash %.3f h2d %.3f ms"
              % (nq, qlen, pipe_kb, 1e3 * dt / nq, (doc.nbytes + score.nbytes) / dt / 1e9,
                 t["score_ms"] / reps, t["select_ms"] / reps, t["d2h_ms"] / reps,
                 t["hashes_ms"] / reps, t["h2d_ms"] / reps), flush=True)
g.close()
