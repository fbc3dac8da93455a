#!/usr/bin/env python3
"""torchrun diagnostics: where does ShardedSearch's host-buffer path lose time against the
device-resident one?  Times submit_host/collect variants on cfg4 shards."""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200
from cobs_b200.dist import ShardedSearch

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
cfg = bench.WORKLOADS["cfg4"]
nq = cfg["nq"]
ix = cobs_b200.GpuIndex.procedural(0, cfg["n_docs"], cfg["sig"], 3, fill_seed=bench.FILL_SEED, device=lr,
                                   shard_index=rank, shard_count=world)
ix.set_option("max_batch", nq)
s = ShardedSearch(ix, rank, world, 128, overlap=True)
batches = [bench.make_batch(1000 + i, nq) for i in range(12)]
pinned = [torch.from_numpy(b).pin_memory() for b, _ in batches]
off = batches[0][1]
ix.set_option("prefetch", 1)
ix.set_option("inputs_ready", 0)
thr = cfg["thr_hits"]

def run(n, mode):
    pend = []
    t_sub = t_col = 0.0
    for i in range(n):
        a = time.perf_counter()
        pend.append(s.submit_host(pinned[i % 12], off, thr, 0))
        t_sub += time.perf_counter() - a
        if len(pend) == 3:
            a = time.perf_counter()
            t = pend.pop(0)
            if mode == "nokeys":
                t["done"].synchronize()
            else:
                s.collect(t)
            t_col += time.perf_counter() - a
    while pend:
        t = pend.pop(0)
        if mode == "nokeys":
            t["done"].synchronize()
        else:
            s.collect(t)
    return t_sub, t_col

for mode in ("full", "nokeys"):
    for steps in (20, 100):
        run(8, mode)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_sub, t_col = run(steps, mode)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print("world %d mode %-6s steps %3d: %.3f ms/step (submit %.3f, collect incl. wait %.3f)" % (
                world, mode, steps, 1e3 * dt / steps, 1e3 * t_sub / steps, 1e3 * t_col / steps), flush=True)
# device-resident for comparison
d_b = [p.cuda() for p in pinned]
ix.set_option("inputs_ready", 1)
for steps in (20, 100):
    for i in range(4):
        s.search_device(d_b[i], off, thr, 0)
    s.join(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        s.search_device(d_b[i % 12], off, thr, 0)
    s.join(); torch.cuda.synchronize()
    if rank == 0:
        print("world %d device-resident steps %3d: %.3f ms/step" % (world, steps, 1e3 * (time.perf_counter() - t0) / steps), flush=True)
ix.close()
dist.destroy_process_group()
