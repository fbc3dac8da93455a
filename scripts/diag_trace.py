#!/usr/bin/env python3
"""device timeline of the pipelined host-buffer path (COBSGPU_TRACE=1): prints phase intervals"""
import os
import sys
os.environ["COBSGPU_TRACE"] = "1"
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

nq = 10000
g = cobs_b200.GpuIndex.procedural(0, 100_000, [1000003], 3, fill_seed=bench.FILL_SEED)
g.set_option("max_batch", nq)
batches = [bench.make_batch(1000 + i, nq) for i in range(12)]
pinned = [torch.from_numpy(b).pin_memory().numpy() for b, _ in batches]
off = batches[0][1]
g.set_option("timing", 0)
pend = []
for i in range(4):
    pend.append(g.submit(pinned[i], off, 0.1, 0))
    if len(pend) == 3:
        g.collect(pend.pop(0), raw=True)
while pend:
    g.collect(pend.pop(0), raw=True)
g.set_option("timing", 1)
print("---- timed", file=sys.stderr)
for i in range(4, 12):
    pend.append(g.submit(pinned[i], off, 0.1, 0))
    if len(pend) == 3:
        g.collect(pend.pop(0), raw=True)
while pend:
    g.collect(pend.pop(0), raw=True)
g.close()
