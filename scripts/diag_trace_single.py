#!/usr/bin/env python3
"""device timeline of ONE search() per call (COBSGPU_TRACE=1) next to the host wall time:
which of upload / K1 / K2 / K3 / wake-up the ~50 us of a single short query are made of"""
import os
import sys
import time
os.environ["COBSGPU_TRACE"] = "1"
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

g = cobs_b200.GpuIndex.procedural(0, 1_000_000, [100003], 3, fill_seed=bench.FILL_SEED)
blob, off = bench.make_batch(7200, 64)
pin = torch.from_numpy(blob).pin_memory().numpy()
for thr, lim in ((0.8, 0), (0.0, 10)):
    g.set_option("timing", 0)
    ts = []
    for i in range(40):
        q = pin[i * 100:(i + 1) * 100]
        t0 = time.perf_counter()
        g.search_packed(q, off[:2], thr, lim, raw=True)
        ts.append(time.perf_counter() - t0)
    print("threshold %.1f limit %d: wall p50 %.1f us (timing off)" % (thr, lim, 1e6 * float(np.percentile(ts[8:], 50))), file=sys.stderr, flush=True)
    g.set_option("timing", 1)
    for i in range(40, 44):
        q = pin[i * 100:(i + 1) * 100]
        t0 = time.perf_counter()
        g.search_packed(q, off[:2], thr, lim, raw=True)
        print("---- call wall %.1f us (timing on)" % (1e6 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
g.close()
