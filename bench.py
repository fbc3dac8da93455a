#!/usr/bin/env python3
"""bench.py -- the COBS query hot path on N B200s (one process per GPU), BASELINE.json's metric:
query k-mers/s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (K1 hash -> K2 gather/AND/count -> K3 threshold/order, and
for N > 1 the NCCL all-gather + merge of the per-shard result lists) over one batch of
synthetic queries against a synthetic index resident in HBM.  Workloads (BASELINE.json
`configs`, SURVEY.md section 8d realisation):
  cfg2  classic, 100 000 docs, 8 388 593 signature rows (104.9 GB in HBM), h=3, k=31,
        10 000 random 100-bp queries per step, threshold 0.8 (the CLI default)
  cfg4  classic, 1 000 000 docs, 1 048 573 rows (131 GB), h=3, 2 048 queries per step
  cfg3  compact, 1 000 000 docs, 8 pages of 16 384 B, h=4, 512 queries per step
  cfg5  compact, 10 000 000 docs, 77 pages of 16 384 B (~600 GB): multi-GPU only
For N > 1 the SAME index is sharded along the document axis (strong scaling).

Every step uses a different query batch, and the row set a batch touches (>= 26 GB) is far
larger than the 126 MB L2, so no timed iteration can be served from cache.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "query_kmers_per_s"
UNIT = "k-mers/s"

WORKLOADS = {
    # name: kind, n_docs, signature sizes, page_size, h, queries per step, cpu rows
    "cfg2": dict(kind=0, n_docs=100_000, sig=[8_388_593], page_size=0, h=3, nq=10_000,
                 desc="classic index, 100000 docs, 8388593 rows (104.9 GB HBM), h=3, k=31, "
                      "10000 random 100-bp queries/step, threshold 0.8"),
    "cfg4": dict(kind=0, n_docs=1_000_000, sig=[1_048_573], page_size=0, h=3, nq=2_048,
                 desc="classic index, 1000000 docs, 1048573 rows (131 GB HBM), h=3, k=31, "
                      "2048 random 100-bp queries/step, threshold 0.8"),
    "cfg3": dict(kind=1, n_docs=1_000_000, page_size=16_384, h=4, nq=512,
                 sig=[int(196_613 * 1.5 ** p) for p in range(8)],
                 desc="compact index, 1000000 docs, 8 pages x 16384 B, h=4, k=31, "
                      "512 random 100-bp queries/step, threshold 0.8"),
    # BASELINE.json configs[4]: index larger than one GPU's HBM -> needs >= 4 GPUs
    "cfg5": dict(kind=1, n_docs=10_000_000, page_size=16_384, h=4, nq=256,
                 sig=[int(100_003 * 1.0345 ** p) for p in range(77)],
                 desc="compact index, 10000000 docs, 77 pages x 16384 B (~600 GB, document-sharded), "
                      "h=4, k=31, 256 random 100-bp queries/step, threshold 0.8"),
    # small variant for functional checks on any GPU
    "tiny": dict(kind=0, n_docs=100_000, sig=[65_521], page_size=0, h=3, nq=2_000,
                 desc="classic index, 100000 docs, 65521 rows (0.8 GB), h=3, k=31"),
}
QUERY_LEN = 100
K = 31
THRESHOLD = 0.8
FILL_SEED = 20260101


_REAL_STDOUT = None


def quiet_stdout():
    """Everything but the result line goes to stderr: libraries (NCCL prints its version line on
    stdout, the reference logs through tlx) must not pollute the ONE JSON line of the contract."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def make_batch(seed, nq):
    """nq random ACGT queries of QUERY_LEN bp as one uint8 blob + offsets"""
    rng = np.random.default_rng(seed)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nq * QUERY_LEN)]
    off = np.arange(nq + 1, dtype=np.uint64) * QUERY_LEN
    return np.ascontiguousarray(blob), off


# ------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU path on the host cores

def cpu_reference_run(wl, steps, warmup, queries_per_step, threads=None, budget_s=None):
    """Times cobs::ClassicSearch::search (oracle/_ref, the unmodified reference) -- or the
    oracle port when the reference library is absent -- on a reduced-row index with the SAME
    number of documents (same bytes per k-mer; the full-size matrix does not fit host RAM).
    Returns (kmers_per_s, info dict)."""
    from oracle import oracle, ref
    cfg = WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    # rows reduced so the file fits comfortably in host RAM / page cache
    row_bytes = (cfg["n_docs"] + 7) // 8 if cfg["kind"] == 0 else cfg["page_size"]
    budget = 768 << 20
    if cfg["kind"] == 0:
        sig = [max(1021, min(cfg["sig"][0], budget // row_bytes))]
    else:
        scale = min(1.0, budget / (sum(cfg["sig"]) * row_bytes))
        sig = [max(101, int(s * scale)) for s in cfg["sig"]]
    ix = oracle.Index.procedural(cfg["kind"], cfg["n_docs"], sig, cfg["h"],
                                 page_size=cfg["page_size"], fill_seed=FILL_SEED)
    T = QUERY_LEN - K + 1
    use_ref = ref.available()
    scratch = os.path.join(ROOT, "build", "tmp")
    os.makedirs(scratch, exist_ok=True)
    tmpdir = tempfile.mkdtemp(prefix="cobs_cpu_", dir=scratch)
    path = os.path.join(tmpdir, "cpu.cobs_" + ("classic" if cfg["kind"] == 0 else "compact"))
    try:
        if use_ref:
            ix.write(path)
            ref.set_threads(cores)
            ref.set_load_complete(True)
            s = ref.Search(path)

            def run(qs):
                sec, _ = s.bench(qs, THRESHOLD, 0)
                return sec
            kind = "reference"
        else:
            ix_m = oracle.Index.procedural(cfg["kind"], cfg["n_docs"], sig, cfg["h"],
                                           page_size=cfg["page_size"], fill_seed=FILL_SEED,
                                           materialize=True)
            cores = 1

            def run(qs):
                t0 = time.perf_counter()
                for q in qs:
                    oracle.search(ix_m, q, THRESHOLD, 0)
                return time.perf_counter() - t0
            kind = "port"

        def batch(seed, n):
            blob, _ = make_batch(seed, n)
            raw = blob.tobytes()
            return [raw[i * QUERY_LEN:(i + 1) * QUERY_LEN] for i in range(n)]

        if budget_s is not None:
            # size the sample from a short probe so the leg takes about budget_s seconds
            probe = 64
            sec = run(batch(999, probe))
            queries_per_step = int(max(probe, min(20000, budget_s / max(sec / probe, 1e-6))))
        for w in range(warmup):
            run(batch(5000 + w, max(16, queries_per_step // 8)))
        total_s, total_q = 0.0, 0
        for st in range(steps):
            total_s += run(batch(6000 + st, queries_per_step))
            total_q += queries_per_step
        rate = total_q * T / total_s
        info = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": "%d queries x %d k-mers/step x %d steps on a %d-doc index with %s rows "
                          "(full-size matrix does not fit host RAM; bytes per k-mer unchanged), "
                          "threshold %.1f, %d threads" %
                          (queries_per_step, T, steps, cfg["n_docs"], sig, THRESHOLD, cores),
                "ms_per_step": 1e3 * total_s / max(steps, 1),
                "queries_per_step": queries_per_step}
        if use_ref:
            s.close()
        return rate, info
    finally:
        try:
            if os.path.exists(path):
                os.unlink(path)
            os.rmdir(tmpdir)
        except OSError:
            pass


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return   # rank 0 alone runs the CPU arm
    cfg = WORKLOADS[args.workload]
    T = QUERY_LEN - K + 1
    rate, info = cpu_reference_run(args.workload, args.steps, args.warmup, args.ref_queries)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + cfg["desc"],
                   "cpu_sample": info["sample"], "kmers_per_query": T},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import cobs_b200
    from cobs_b200.dist import GridSearch, QuerySplitSearch, ShardedSearch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    cfg = WORKLOADS[args.workload]
    nq = args.nq or cfg["nq"]
    T = QUERY_LEN - K + 1
    sig = cfg["sig"]
    if args.rows:
        sig = [args.rows] * len(sig)
    split_queries = args.parallelism == "queries" and world > 1
    grid = args.parallelism == "grid" and world > 1
    rpq = args.results_per_query

    def open_index(shard_index, shard_count):
        ix = cobs_b200.GpuIndex.procedural(cfg["kind"], cfg["n_docs"], sig, cfg["h"],
                                           page_size=cfg["page_size"], fill_seed=FILL_SEED,
                                           device=local_rank, shard_index=shard_index,
                                           shard_count=shard_count)
        ix.set_option("max_batch", max(nq, 1))
        return ix

    if grid:
        # documents x queries: --doc-shards D ranks share one copy of the index, world / D copies
        sharded = GridSearch(open_index, rank, world, args.doc_shards, rpq,
                             overlap=not args.no_overlap)
        index = sharded.index
    else:
        if args.emulate_shards:
            index = open_index(0, args.emulate_shards)
        elif split_queries:
            index = open_index(0, 1)
        else:
            index = open_index(rank, world)
        cls = QuerySplitSearch if split_queries else ShardedSearch
        sharded = cls(index, rank, world, rpq, overlap=not args.no_overlap)
    info = index.info

    # bytes per k-mer of the WHOLE index (h * ceil(N/8), unpadded reference layout)
    if cfg["kind"] == 0:
        bytes_per_kmer = cfg["h"] * ((cfg["n_docs"] + 7) // 8)
    else:
        bytes_per_kmer = cfg["h"] * len(sig) * cfg["page_size"]

    n_batches = args.warmup + args.steps
    batches = [make_batch(1000 + i, nq) for i in range(n_batches)]
    pinned = [torch.from_numpy(b).pin_memory() for b, _ in batches]
    pinned_np = [p.numpy() for p in pinned]
    off = batches[0][1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs ("value") ----
    d_batches = [p.to(dev) for p in pinned]
    torch.cuda.synchronize()
    if not args.no_overlap:
        # steps are independent batches: upload + K1 of step i+1 (and, for N > 1, the exchange
        # of step i) overlap the score kernel of the neighbouring step
        index.set_option("prefetch", 1)
        index.set_option("inputs_ready", 1)     # every batch is already resident in HBM
    for i in range(args.warmup):
        sharded.search_device(d_batches[i], off, THRESHOLD, 0)
    sharded.join()
    barrier()
    index.set_option("timing", 1)
    index.timers(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    last = None
    for i in range(args.steps):
        last = sharded.search_device(d_batches[args.warmup + i], off, THRESHOLD, 0)
    sharded.join()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    tm = index.timers()
    score_ms = torch.tensor([tm["score_ms"] / max(tm["score_launches"], 1)], device=dev)
    phases = {k: tm[k] / args.steps for k in ("hashes_ms", "score_ms", "select_ms", "h2d_ms")}
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(score_ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    kmers_per_step = nq * T
    value = kmers_per_step * args.steps / (ms_total * 1e-3)
    launches = tm["kernel_launches"] + (args.steps if world > 1 else 0)
    index.set_option("timing", 0)
    index.set_option("inputs_ready", 0)     # the e2e leg uploads its queries itself
    n_results = int((last[-2].cpu().numpy().view(np.uint32).reshape(-1) % 0xFFFFFFFE).sum())

    # ---- leg 2: end to end from host buffers through the public API ("e2e") ----
    h2d = int(pinned[0].numel() + off.nbytes)
    if world == 1:
        def e2e_step(i):
            roff, doc, score = index.search_packed(pinned_np[i], off, THRESHOLD, 0, raw=True)
            return roff.nbytes + doc.nbytes + score.nbytes
    else:
        def e2e_step(i):
            c, k = sharded.search_host(pinned[i], off, THRESHOLD, 0)
            return c.nbytes + k.nbytes
    for i in range(args.warmup):
        e2e_step(i)
    if world > 1:   # warm the streaming pattern too (both buffer sets, side streams)
        tk = [sharded.submit_host(pinned[i], off, THRESHOLD, 0) for i in range(sharded.depth - 1)]
        for t in tk:
            sharded.collect(t)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    if world == 1:
        for i in range(args.steps):
            d2h = e2e_step(args.warmup + i)
    else:
        # streaming use of the public API: up to three batches in flight (submit ahead, collect
        # in order); every batch still pays its own H2D and D2H inside the timed region
        pending = []
        for i in range(args.steps):
            pending.append(sharded.submit_host(pinned[args.warmup + i], off, THRESHOLD, 0))
            if len(pending) == sharded.depth - 1:
                c, k = sharded.collect(pending.pop(0))
                d2h = c.nbytes + k.nbytes
        while pending:
            c, k = sharded.collect(pending.pop(0))
            d2h = c.nbytes + k.nbytes
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = kmers_per_step * args.steps / float(e2e_s.item())

    # ---- roofline of the score kernel (K2), measured with CUDA events in the timed loop ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak = float(json.load(f)["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    # DRAM bytes per K2 launch from the committed `ncu --set full` capture of this workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not args.nq and not args.rows and world == 1:
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    k2_ms = float(score_ms.item())
    # this rank's share of a step: its document shard, or its slice of the queries
    algo_bytes_per_launch = info.bytes_per_kmer * kmers_per_step
    if split_queries:
        per = (nq + world - 1) // world
        algo_bytes_per_launch = info.bytes_per_kmer * per * T
    if grid:
        lo, hi = sharded.slice_of(nq)
        algo_bytes_per_launch = info.bytes_per_kmer * (hi - lo) * T
    achieved = algo_bytes_per_launch / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_reference_run(args.workload, 1, 1, 0, budget_s=args.cpu_seconds)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:   # the baseline is reported, never load-bearing
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable",
                       "sample": "failed: %r" % (e,)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + cfg["desc"],
                       "parallelism": ("single GPU" if world == 1 else
                                       "index replicated x%d, queries split, NCCL all-gather of "
                                       "the result blocks" % world if split_queries else
                                       "%d document shards x %d query groups, NCCL all-gather + "
                                       "merge inside each group" % (args.doc_shards, world // args.doc_shards)
                                       if grid else
                                       "document-axis shards x%d, NCCL all-gather of per-rank "
                                       "result blocks + merge" % world),
                       "queries_per_step": nq, "kmers_per_query": T,
                       "bytes_per_kmer": bytes_per_kmer, "hbm_bytes_this_rank": info.hbm_bytes,
                       "l2_policy": "each step a different batch; rows touched per step "
                                    "(%.1f GB) >> 126 MB L2" % (algo_bytes_per_launch / 1e9),
                       "results_last_step": n_results},
            "queries_per_s": value / T,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "score_kernel<%d,CAND>" % cfg["h"],
                         "kernel_ms": k2_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                         "whole_step_frac": (bytes_per_kmer / world) * value / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "phases_ms_per_step_rank0": phases,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    index.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--nq", type=int, default=0, help="queries per step (default: workload's)")
    ap.add_argument("--rows", type=int, default=0, help="override signature rows (debug)")
    ap.add_argument("--results-per-query", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-queries", type=int, default=400,
                    help="queries per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="docs", choices=["docs", "queries", "grid"],
                    help="N > 1: shard the document axis (default, BASELINE.json's north_star), "
                         "replicate the index and split every query batch, or both (grid: "
                         "--doc-shards ranks share one copy of the index)")
    ap.add_argument("--doc-shards", type=int, default=2, help="document shards per copy (grid)")
    ap.add_argument("--emulate-shards", type=int, default=0,
                    help="debug: hold shard 0 of this many document shards on one GPU")
    ap.add_argument("--no-overlap", action="store_true",
                    help="disable the cross-step pipelining (K1 prefetch, side-stream exchange)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
