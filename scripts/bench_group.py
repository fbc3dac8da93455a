#!/usr/bin/env python3
"""Throughput of the single-process multi-GPU path (cobsgpu_group_search_batch, the C-ABI entry
point behind `cobs query --gpus N` and ClassicSearch with gopt_gpus > 1): cfg4 (1 M documents,
131 GB) sharded over the visible GPUs of ONE process, batches handed over as HOST buffers, result
lists returned to the host -- i.e. an end-to-end number.  Prints one JSON line per GPU count with
an oracle parity check of sampled queries.

    python scripts/bench_group.py [--gpus 2,4,8] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--workload", default="cfg4")
args = ap.parse_args()
n_dev = cobs_b200.lib().cobsgpu_device_count()
counts = [int(x) for x in args.gpus.split(",")] if args.gpus else [n for n in (1, 2, 4, 8) if n <= n_dev]
cfg = bench.WORKLOADS[args.workload]
nq = cfg["nq"]
T = bench.T_KMERS
batches = [bench.make_batch(1000 + i, nq) for i in range(4 + args.steps)]
off = batches[0][1]
try:
    import torch
    pinned = [torch.from_numpy(b).pin_memory().numpy() for b, _ in batches]
except Exception:
    pinned = [b for b, _ in batches]

for n in counts:
    g = cobs_b200.GpuGroup.procedural(cfg["kind"], cfg["n_docs"], cfg["sig"], cfg["h"], list(range(n)),
                                      page_size=cfg["page_size"], fill_seed=bench.FILL_SEED)
    g.set_option("max_batch", nq)
    out = {}
    for name, thr, limit in (("hits", cfg["thr_hits"], 0), ("threshold0_limit10", 0.0, 10)):
        for i in range(4):
            g.search_packed(pinned[i], off, thr, limit, raw=True)
        # one call carrying `steps` batches: the library pipelines them through its slot ring
        blob = np.concatenate(pinned[4:4 + args.steps])
        off_all = np.arange(nq * args.steps + 1, dtype=np.uint64) * bench.QUERY_LEN
        g.set_option("max_batch", nq)
        t0 = time.perf_counter()
        roff, doc, score = g.search_packed(blob, off_all, thr, limit, raw=True)
        dt = time.perf_counter() - t0
        out[name] = {"value": nq * args.steps * T / dt, "ms_per_batch": 1e3 * dt / args.steps,
                     "results_per_batch": int(roff[-1]) // args.steps}
        if name == "hits":
            last = (roff, doc, score)
    # parity of sampled queries of the last batch against the CPU oracle
    roff, doc, score = last
    base = nq * (args.steps - 1)
    sample = np.sort(np.random.default_rng(7).choice(nq, size=32, replace=False))
    raw = pinned[4 + args.steps - 1].tobytes()
    qs = [raw[q * bench.QUERY_LEN:(q + 1) * bench.QUERY_LEN] for q in sample]
    want = bench.oracle_lists(cfg, cfg["sig"], qs, cfg["thr_hits"], threads=os.cpu_count() or 1)
    bad = 0
    for q, w in zip(sample, want):
        a, b = int(roff[base + q]), int(roff[base + q + 1])
        if [(int(d), int(s)) for d, s in zip(doc[a:b], score[a:b])] != [(d, s) for _, d, s in w]:
            bad += 1
    line = {"path": "cobsgpu_group_search_batch (one process, host buffers in, host lists out)",
            "workload": args.workload, "n_gpus": n, "queries_per_batch": nq, "batches": args.steps,
            "legs": out, "parity_check": {"queries": len(sample), "mismatches": bad, "ok": bad == 0}}
    print(json.dumps(line), flush=True)
    g.close()
