#!/usr/bin/env python3
"""single long query latency with and without the k-split kernel (COBSGPU_NO_KSPLIT=1)"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

for name, n_docs, rows in (("100k docs", 100_000, 1_000_003), ("1M docs", 1_000_000, 100_003)):
    g = cobs_b200.GpuIndex.procedural(0, n_docs, [rows], 3, fill_seed=bench.FILL_SEED)
    for qlen in (1030, 10_030):
        blob, off = bench.make_batch(7300, 24, qlen)
        ts = []
        for i in range(24):
            q = blob[i * qlen:(i + 1) * qlen].copy()
            t0 = time.perf_counter()
            g.search_packed(q, off[:2], 0.8, 0, raw=True)
            ts.append(time.perf_counter() - t0)
        print("%s ksplit=%s %d k-mers: p50 %.1f us" % (name, "off" if os.environ.get("COBSGPU_NO_KSPLIT") else "on",
                                                         qlen - 30, 1e6 * float(np.percentile(ts[4:], 50))), flush=True)
    g.close()
