#!/usr/bin/env python3
"""bench.py -- the COBS query hot path on N B200s (one process per GPU), BASELINE.json's metric:
query k-mers/s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (K1 hash -> K2 gather/AND/count -> K3 threshold/order, and
for N > 1 the NCCL all-gather + merge of the per-shard result lists) over one batch of synthetic
queries against a synthetic index resident in HBM.  Workloads (BASELINE.json `configs`, SURVEY.md
section 8d realisation):
  cfg4  classic, 1 000 000 docs, 1 048 573 rows (131 GB), h=3, 16 384 queries per step: the
        north-star configuration and the HEADLINE of every run (document-axis shards for N > 1)
  cfg2  classic, 100 000 docs, 8 388 593 signature rows (104.9 GB in HBM), h=3, 10 000 queries
  cfg3  compact, 1 000 000 docs, 8 pages of 16 384 B, h=4, 2 048 queries per step
  cfg5  compact, 10 000 000 docs, 77 pages of 16 384 B (~600 GB): needs the HBM of >= 4 GPUs
One invocation times the headline workload and -- in `secondary` -- cfg2 (every N; for N > 1 with
the shard policy chosen from the index size), cfg3 (N = 1) and cfg5 (N = 8).

Every workload is timed on two legs: `hits` (a threshold that yields 10-100 documents per query,
so candidate append, K3, the result copies and the cross-rank merge all carry payload; this is the
leg `value` and `e2e` report) and `default_threshold` (the CLI's 0.8, which random queries never
reach).  After the timed loops a `parity_check` compares the merged results of sampled queries of
the last timed batch with the CPU oracle on the same procedural index; a mismatch fails the run.
`patterns` times the reference's other call patterns: threshold 0 with -l 10, 1000-k-mer queries
(benchmark-fpr's default), one search() per call.

Every step uses a different query batch, and the row set a batch touches (>= 18 GB) is far larger
than the 126 MB L2, so no timed iteration can be served from cache.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "query_kmers_per_s"
UNIT = "k-mers/s"

WORKLOADS = {
    # thr_hits: threshold of the `hits` leg (ceil(thr * 70) k-mers => 10-100 documents/query)
    "cfg2": dict(kind=0, n_docs=100_000, sig=[8_388_593], page_size=0, h=3, nq=10_000, thr_hits=0.1,
                 desc="classic index, 100000 docs, 8388593 rows (104.9 GB HBM), h=3, k=31, "
                      "10000 random 100-bp queries/step"),
    "cfg4": dict(kind=0, n_docs=1_000_000, sig=[1_048_573], page_size=0, h=3, nq=16_384, thr_hits=0.11,
                 desc="classic index, 1000000 docs, 1048573 rows (131 GB HBM), h=3, k=31, "
                      "16384 random 100-bp queries/step"),
    "cfg3": dict(kind=1, n_docs=1_000_000, page_size=16_384, h=4, nq=2_048, thr_hits=0.07,
                 sig=[int(196_613 * 1.5 ** p) for p in range(8)],
                 desc="compact index, 1000000 docs, 8 pages x 16384 B (158.7 GB HBM), h=4, k=31, "
                      "2048 random 100-bp queries/step"),
    # BASELINE.json configs[4]: index larger than one GPU's HBM -> needs >= 4 GPUs
    "cfg5": dict(kind=1, n_docs=10_000_000, page_size=16_384, h=4, nq=1_024, thr_hits=0.07,
                 sig=[int(100_003 * 1.0345 ** p) for p in range(77)],
                 desc="compact index, 10000000 docs, 77 pages x 16384 B (~600 GB, document-sharded), "
                      "h=4, k=31, 1024 random 100-bp queries/step"),
    # small variant for functional checks on any GPU
    "tiny": dict(kind=0, n_docs=100_000, sig=[65_521], page_size=0, h=3, nq=2_000, thr_hits=0.1,
                 desc="classic index, 100000 docs, 65521 rows (0.8 GB), h=3, k=31"),
}
QUERY_LEN = 100
K = 31
T_KMERS = QUERY_LEN - K + 1
THRESHOLD = 0.8          # the CLI default (src/cobs.cpp:478)
FILL_SEED = 20260101
HBM_BUDGET = 140e9       # bytes of index one GPU is asked to hold when the shard policy is "auto"


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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def make_batch(seed, nq, qlen=QUERY_LEN):
    """nq random ACGT queries of qlen bp as one uint8 blob + offsets"""
    rng = np.random.default_rng(seed)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nq * qlen)]
    off = np.arange(nq + 1, dtype=np.uint64) * qlen
    return np.ascontiguousarray(blob), off


def bytes_per_kmer_of(cfg, sig):
    """algorithmic bytes per query k-mer of the WHOLE index: h * ceil(N/8) (classic) or
    h * P * page_size (compact), the unpadded reference layout"""
    if cfg["kind"] == 0:
        return cfg["h"] * ((cfg["n_docs"] + 7) // 8)
    return cfg["h"] * len(sig) * cfg["page_size"]


def index_bytes_of(cfg, sig):
    row = (cfg["n_docs"] + 7) // 8 if cfg["kind"] == 0 else cfg["page_size"]
    return sum(sig) * row


def auto_doc_shards(index_bytes, world, budget=HBM_BUDGET):
    """document shards per copy of the index: the smallest divisor D of `world` whose shard fits
    the per-GPU budget; the world / D copies split every query batch (D == world: pure document
    sharding, D == 1: replicas).  Keeps the row slices per GPU as long as the HBM allows."""
    for d in range(1, world + 1):
        if world % d == 0 and index_bytes / d <= budget:
            return d
    return world


# ------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)

class ClockSampler:
    """nvidia-smi polled in the background.  It is started BEFORE the warm-up steps (its start-up
    takes driver locks that would stall the first timed launches) and only the samples that
    arrive between begin() and end() -- the timed region -- are summarised."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None

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
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.proc = None

    def summary(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)    # one more polling period: the sample covering the end of the region
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = (self.t1 if self.t1 is not None else time.perf_counter()) + 0.12
        snapshot = list(self.lines)
        inside = [l for t, l in snapshot if t0 <= t <= t1]
        lines = inside if inside else [l for _, l in snapshot]
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
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
                "samples": len(sm), "samples_in_timed_region": len(inside),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU path on the host cores

def cpu_reference_run(wl, steps, warmup, queries_per_step, threads=None, budget_s=None,
                      threshold=None):
    """Times cobs::ClassicSearch::search (oracle/_ref, the unmodified reference) -- or the
    oracle port when the reference library is absent -- on a reduced-row index with the SAME
    number of documents (same bytes per k-mer; the full-size matrix does not fit host RAM).
    Returns (kmers_per_s, info dict)."""
    from oracle import oracle, ref
    cfg = WORKLOADS[wl]
    thr = cfg["thr_hits"] if threshold is None else threshold
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
                sec, _ = s.bench(qs, thr, 0)
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
                    oracle.search(ix_m, q, thr, 0)
                return time.perf_counter() - t0
            kind = "port"

        def batch(seed, n):
            blob, _ = make_batch(seed, n)
            raw = blob.tobytes()
            return [raw[i * QUERY_LEN:(i + 1) * QUERY_LEN] for i in range(n)]

        if budget_s is not None:
            # size the sample from a short probe so the leg takes about budget_s seconds
            probe = 32
            sec = run(batch(999, probe))
            queries_per_step = int(max(probe, min(20000, budget_s / max(sec / probe, 1e-6))))
        for w in range(warmup):
            run(batch(5000 + w, max(16, queries_per_step // 8)))
        total_s, total_q = 0.0, 0
        for st in range(steps):
            total_s += run(batch(6000 + st, queries_per_step))
            total_q += queries_per_step
        rate = total_q * T_KMERS / total_s
        info = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": "%d queries x %d k-mers/step x %d steps on a %d-doc index with %s rows "
                          "(full-size matrix does not fit host RAM; bytes per k-mer unchanged), "
                          "threshold %.2f, %d threads" %
                          (queries_per_step, T_KMERS, steps, cfg["n_docs"], sig, thr, cores),
                "ms_per_step": 1e3 * total_s / max(steps, 1),
                "queries_per_step": queries_per_step}
        if use_ref:
            s.close()
        return rate, info
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return   # rank 0 alone runs the CPU arm
    wl = args.workload
    cfg = WORKLOADS[wl]
    rate, info = cpu_reference_run(wl, args.steps, args.warmup, args.ref_queries)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl + ": " + cfg["desc"] + ", threshold %.2f" % cfg["thr_hits"],
                   "cpu_sample": info["sample"], "kmers_per_query": T_KMERS},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm

class Env:
    """process-wide state of one bench run"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                self.peak = float(json.load(f)["hbm_gbs"])
            self.peak_src = "MEASURED_PEAKS.json hbm_gbs"
        # nvidia-smi polls for the whole run (its start-up takes driver locks for about a second
        # and must be long over when the timed region begins); begin()/end() mark that region
        self.sampler = None
        if self.rank == 0:
            self.sampler = ClockSampler(self.local_rank)
            self.sampler.start()
        self.traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                self.traffic = json.load(f)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def results_of(counts, keys, queries):
    """device (counts, keys) of a batch -> {query: (doc[], score[]) or None when flagged}"""
    import cobs_b200
    c = counts.reshape(-1).cpu().numpy().view(np.uint32)
    kmax = keys.shape[-1]
    k = keys.reshape(-1, kmax).cpu().numpy().view(np.uint64)
    out = {}
    for q in queries:
        if c[q] >= 0xFFFFFFFE:
            out[q] = None
        else:
            out[q] = cobs_b200.decode_keys(k[q, :c[q]])
    return out


def oracle_lists(cfg, sig, queries, thr, threads):
    """the CPU oracle's result lists for raw query strings on the procedural index"""
    from oracle import oracle
    o = oracle.Index.procedural(cfg["kind"], cfg["n_docs"], sig, cfg["h"],
                                page_size=cfg["page_size"], fill_seed=FILL_SEED)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(lambda q: oracle.search(o, q, thr, 0), queries))


def run_workload(env, wl, primary, parallelism):
    """Times one workload; returns the dict that becomes the headline (primary) or one
    `secondary` block.  Collective on all ranks."""
    import cobs_b200
    from cobs_b200.dist import GridSearch, ShardedSearch
    torch, args = env.torch, env.args
    world, rank = env.world, env.rank
    cfg = WORKLOADS[wl]
    nq = args.nq or cfg["nq"]
    sig = [args.rows] * len(cfg["sig"]) if args.rows else cfg["sig"]
    steps, warmup = args.steps, args.warmup
    rpq = args.results_per_query
    thr_hits = cfg["thr_hits"]
    bpk = bytes_per_kmer_of(cfg, sig)

    if parallelism == "auto":
        doc_shards = auto_doc_shards(index_bytes_of(cfg, sig), world)
    elif parallelism == "queries":
        doc_shards = 1
    elif parallelism == "grid":
        doc_shards = args.doc_shards
    else:
        doc_shards = world
    groups = world // doc_shards
    # slots per query in a rank's result block: a shard holds 1/doc_shards of a query's documents,
    # and the all-gather moves the whole padded block, so the block shrinks with the shard (a list
    # that still outgrew it would be flagged, never cut -- and fail the parity check below)
    rpq = max(32, rpq // doc_shards)

    def open_index(shard_index, shard_count):
        ix = cobs_b200.GpuIndex.procedural(cfg["kind"], cfg["n_docs"], sig, cfg["h"],
                                           page_size=cfg["page_size"], fill_seed=FILL_SEED,
                                           device=env.local_rank, shard_index=shard_index,
                                           shard_count=shard_count)
        ix.set_option("max_batch", max(nq, 1))
        return ix

    t_open = time.perf_counter()
    if groups > 1:
        search = GridSearch(open_index, rank, world, doc_shards, rpq, overlap=not args.no_overlap)
        index = search.index
        lo, hi = search.slice_of(nq)
    else:
        shards = args.emulate_shards or world
        index = open_index(0 if args.emulate_shards else rank, shards)
        search = ShardedSearch(index, rank, world, rpq, overlap=not args.no_overlap)
        lo, hi = 0, nq
    info = index.info
    open_s = time.perf_counter() - t_open
    my_nq = hi - lo
    if rank == 0:
        log("%s: index open %.1f s, %.1f GB on this rank, %d doc shards x %d query groups"
            % (wl, open_s, info.hbm_bytes / 1e9, doc_shards, groups))

    n_batches = warmup + steps
    batches = [make_batch(1000 + i, nq) for i in range(n_batches)]
    pinned = [torch.from_numpy(b).pin_memory() for b, _ in batches]
    off = batches[0][1]
    d_batches = [p.to(env.dev) for p in pinned]
    torch.cuda.synchronize()

    def device_call(i, thr, limit):
        r = search.search_device(d_batches[i], off, thr, limit)
        return r[-2:]          # (counts, keys); GridSearch prepends (lo, hi)

    def timed_device_leg(thr, limit, sample_clocks):
        """W warm-up + K timed device-resident steps; returns dict + the last step's results"""
        if not args.no_overlap:
            # steps are independent batches: upload + K1 of step i+1 (and, for N > 1, the exchange
            # of step i) overlap the score kernel of the neighbouring step
            index.set_option("prefetch", 1)
            index.set_option("inputs_ready", 1)     # every batch is already resident in HBM
        sampler = env.sampler if sample_clocks else None
        for i in range(warmup):
            device_call(i, thr, limit)
        search.join()
        env.barrier()
        index.set_option("timing", 1)
        index.timers(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        if sampler:
            sampler.begin()
        ev0.record()
        last = None
        for i in range(steps):
            last = device_call(warmup + i, thr, limit)
        search.join()
        ev1.record()
        env.barrier()
        if sampler:
            sampler.end()
        clocks = sampler.summary() if sampler else None
        ms_total = env.max_over_ranks(ev0.elapsed_time(ev1))
        tm = index.timers()
        k2_ms = env.max_over_ranks(tm["score_ms"] / max(tm["score_launches"], 1))
        index.set_option("timing", 0)
        index.set_option("inputs_ready", 0)
        index.set_option("prefetch", 0)
        kmers_per_step = nq * T_KMERS
        value = kmers_per_step * steps / (ms_total * 1e-3)
        # this rank's share of a step: its document shard x its slice of the queries
        algo = info.bytes_per_kmer * my_nq * T_KMERS
        achieved = algo / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
        c = last[0].reshape(-1).cpu().numpy().view(np.uint32)
        n_results = int(c[c < 0xFFFFFFFE].sum())
        n_flagged = int((c >= 0xFFFFFFFE).sum())
        leg = {"threshold": thr, "limit": limit, "value": value, "ms_per_step": ms_total / steps,
               "kernel_ms": k2_ms, "roofline_frac": achieved / env.peak,
               "achieved_gbs": achieved, "algorithmic_bytes_per_launch": algo,
               # every rank of a query group holds the same merged lists of the group's slice
               "results_last_step": int(env.sum_over_ranks(n_results) / doc_shards),
               "flagged_last_step": n_flagged,
               "phases_ms_per_step_rank0": {k: tm[k] / steps for k in
                                            ("hashes_ms", "score_ms", "select_ms", "h2d_ms")},
               "launches": int(tm["kernel_launches"] + (steps if doc_shards > 1 else 0))}
        if clocks:
            leg["clocks"] = clocks
        return leg, last

    def timed_e2e_leg(thr, limit):
        """the same K batches end to end from pinned host buffers through the public API: every
        batch pays its own H2D of the queries and D2H of its result inside the timed region"""
        h2d = int(pinned[0].numel() + off.nbytes)
        d2h = [0]
        if world == 1:
            pinned_np = [p.numpy() for p in pinned]

            def submit(i):
                return index.submit(pinned_np[i], off, thr, limit)

            def collect(t):
                roff, doc, score = index.collect(t, raw=True)
                d2h[0] = roff.nbytes + doc.nbytes + score.nbytes
            depth = 3
        else:
            if not args.no_overlap:
                # upload + K1 of batch i+1 run ahead on the library's input stream
                index.set_option("prefetch", 1)
                index.set_option("inputs_ready", 0)

            def submit(i):
                return search.submit_host(pinned[i], off, thr, limit)

            def collect(t):
                c, k = search.collect(t)
                d2h[0] = c.nbytes + k.nbytes
            depth = search.depth - 1

        def run(idx):
            pending = []
            for i in idx:
                pending.append(submit(i))
                if len(pending) == depth:
                    collect(pending.pop(0))
            while pending:
                collect(pending.pop(0))
        # (two turns of the 4-slot ring: every slot has its buffers at their steady-state size)
        run([i % (warmup + steps) for i in range(max(warmup, 8))])
        env.barrier()
        t0 = time.perf_counter()
        run(range(warmup, warmup + steps))
        env.torch.cuda.synchronize()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        env.barrier()
        if world > 1:
            index.set_option("prefetch", 0)
        return {"value": nq * T_KMERS * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(env.max_over_ranks(d2h[0]))}

    # ---- the two legs ----
    hits, last = timed_device_leg(thr_hits, 0, sample_clocks=primary)
    # sampled queries of the last timed batch, copied out before a later leg reuses the buffers;
    # every rank checks its share of the sample (all ranks of a query group hold the merged lists
    # of the group's slice)
    n_sample = min(args.parity_queries, nq)
    sample = np.sort(np.random.default_rng(4242).choice(nq, size=n_sample, replace=False))
    mine = [int(q) for j, q in enumerate(sample)
            if lo <= q < hi and j % doc_shards == rank % doc_shards]
    got = results_of(last[0], last[1], [q - lo for q in mine]) if not args.no_parity else {}
    e2e = timed_e2e_leg(thr_hits, 0)
    dflt, _ = timed_device_leg(THRESHOLD, 0, sample_clocks=False)

    # ---- parity of the last timed batch against the CPU oracle ----
    parity = None
    if not args.no_parity:
        t0 = time.perf_counter()
        raw = batches[warmup + steps - 1][0].tobytes()
        qstr = [raw[q * QUERY_LEN:(q + 1) * QUERY_LEN] for q in mine]
        want = oracle_lists(cfg, sig, qstr, thr_hits, threads=os.cpu_count() or 1)
        bad, flagged, n_docs_seen = 0, 0, 0
        for q, w in zip(mine, want):
            g = got[q - lo]
            if g is None:
                flagged += 1
                continue
            gl = [(int(d), int(s)) for d, s in zip(*g)]
            n_docs_seen += len(gl)
            if gl != [(d, s) for _, d, s in w]:
                bad += 1
        checked = int(env.sum_over_ranks(len(mine)))
        bad = int(env.sum_over_ranks(bad))
        flagged = int(env.sum_over_ranks(flagged))
        docs = int(env.sum_over_ranks(n_docs_seen))
        parity = {"queries": checked, "mismatches": bad, "flagged": flagged,
                  "documents_compared": docs, "threshold": thr_hits,
                  "ok": bool(bad == 0 and flagged == 0 and checked > 0 and docs > 0),
                  "oracle": "oracle/cobs_oracle.c on the same procedural index, full lists",
                  "seconds": round(time.perf_counter() - t0, 2)}

    # ---- the reference's other call patterns ----
    patterns = None
    if not args.no_patterns and (primary or wl == "cfg2"):
        patterns = run_patterns(env, cfg, sig, index, search, lo, hi, d_batches, off, warmup, steps)

    block = {
        "workload": wl + ": " + cfg["desc"],
        "value": hits["value"], "unit": UNIT, "ms_per_step": hits["ms_per_step"],
        "queries_per_s": hits["value"] / T_KMERS,
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": hits["achieved_gbs"], "peak": env.peak, "unit": "GB/s",
                     "frac": hits["roofline_frac"],
                     "traffic": (env.traffic.get(wl, {}).get("dram_bytes_per_launch")
                                 if world == 1 and not args.nq and not args.rows else None),
                     "traffic_source": (env.traffic.get(wl, {}).get("source")
                                        if world == 1 and not args.nq and not args.rows else None),
                     "kernel": "score_kernel<%d,CAND,8>" % cfg["h"],
                     "kernel_ms": hits["kernel_ms"], "peak_source": env.peak_src,
                     "algorithmic_bytes_per_launch": hits["algorithmic_bytes_per_launch"],
                     "whole_step_frac": (bpk / world) * hits["value"] / 1e9 / env.peak},
        "legs": {"hits": hits, "default_threshold": dflt},
        "parity_check": parity,
        "parallelism": ("single GPU" if world == 1 else
                        "%d document shards x %d query groups%s" %
                        (doc_shards, groups,
                         ", NCCL all-gather of the per-shard result blocks + merge" if doc_shards > 1 else
                         " (index replicated, queries split)")),
        "queries_per_step": nq, "kmers_per_query": T_KMERS, "bytes_per_kmer": bpk,
        "hbm_bytes_this_rank": info.hbm_bytes, "index_open_s": round(open_s, 2),
        "results_last_step": hits["results_last_step"],
        "gpu_launches": hits["launches"],
    }
    if patterns:
        block["patterns"] = patterns
    index.close()
    del d_batches
    torch.cuda.empty_cache()
    return block


def run_patterns(env, cfg, sig, index, search, lo, hi, d_batches, off, warmup, steps):
    """The reference's other call patterns: threshold 0 with -l 10 (served by
    the per-warp top-k epilogue), 1000-k-mer queries as `cobs benchmark-fpr` draws them
    (src/cobs.cpp:680; 16 bit-planes), and one search() per call (latency)."""
    torch = env.torch
    out = {}
    info = index.info
    nq = len(off) - 1
    steps = max(3, min(steps, 10))

    def leg(name, dq, off_, thr, limit, kmers_per_q):
        n = len(off_) - 1
        mine = n
        if hasattr(search, "slice_of"):
            a, b = search.slice_of(n)
            mine = b - a
        for i in range(4):      # every buffer set of the exchange ring sees this shape once
            search.search_device(dq[i % len(dq)], off_, thr, limit)
        search.join()
        env.barrier()
        index.set_option("timing", 1)
        index.timers(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        for i in range(steps):
            last = search.search_device(dq[i % len(dq)], off_, thr, limit)[-2:]
        search.join()
        ev1.record()
        env.barrier()
        ms = env.max_over_ranks(ev0.elapsed_time(ev1))
        tm = index.timers()
        index.set_option("timing", 0)
        k2 = env.max_over_ranks(tm["score_ms"] / max(tm["score_launches"], 1))
        algo = info.bytes_per_kmer * mine * kmers_per_q
        c = last[0].reshape(-1).cpu().numpy().view(np.uint32) if last[0] is not None else np.zeros(0, np.uint32)
        out[name] = {"threshold": thr, "limit": limit, "queries_per_step": n,
                     "kmers_per_query": kmers_per_q,
                     "value": n * kmers_per_q * steps / (ms * 1e-3), "unit": UNIT,
                     "ms_per_step": ms / steps, "kernel_ms": k2,
                     "roofline_frac": (algo / (k2 * 1e-3) / 1e9 / env.peak) if k2 > 0 else None,
                     "results_last_step": int(c[c < 0xFFFFFFFE].sum()),
                     "flagged_last_step": int((c >= 0xFFFFFFFE).sum()),
                     "phases_ms_per_step_rank0": {k: tm[k] / steps for k in
                                                  ("hashes_ms", "score_ms", "select_ms", "h2d_ms")}}

    # threshold 0 (Search::search's default, cobs/query/search.hpp:41) with -l 10
    leg("threshold0_limit10", d_batches, off, 0.0, 10, T_KMERS)
    # benchmark-fpr's query length: 1000 k-mers (1030 bp), with -l 10 so that the list is bounded
    nq_long = max(8 * env.world, min(nq, 1024) // 8)
    long_b = [make_batch(7000 + i, nq_long, 1030) for i in range(3)]
    d_long = [torch.from_numpy(b).to(env.dev) for b, _ in long_b]
    leg("kmers1000_threshold0_limit10", d_long, long_b[0][1], 0.0, 10, 1000)

    if env.world == 1:
        # benchmark-fpr's own default: 1000 k-mers, threshold 0, ALL documents returned (8 bytes
        # per document per query come back over PCIe: that copy is the bound, not HBM)
        for n_fpr, key in ((16, "benchmark_fpr_default"), (64, "benchmark_fpr_default_64q")):
            blob, off_f = make_batch(7100, n_fpr, 1030)
            pin = torch.from_numpy(blob).pin_memory().numpy()
            for _ in range(4):     # warm-up: every slot of the ring page-locks its result buffer once
                index.search_packed(pin, off_f, 0.0, 0, raw="view")
            torch.cuda.synchronize()
            blob, off_f = make_batch(7101, n_fpr, 1030)
            pin = torch.from_numpy(blob).pin_memory().numpy()
            t0 = time.perf_counter()
            # (arrays alias the library's result buffers: what the C++ callers get)
            roff, doc, score = index.search_packed(pin, off_f, 0.0, 0, raw="view")
            dt = time.perf_counter() - t0
            out[key] = {
                "threshold": 0.0, "limit": 0, "queries": n_fpr, "kmers_per_query": 1000,
                "value": n_fpr * 1000 / dt, "unit": UNIT, "ms_per_query": 1e3 * dt / n_fpr,
                "results_per_query": int(roff[-1]) // n_fpr,
                "d2h_gbs": (doc.nbytes + score.nbytes) / dt / 1e9,
                "bound": "result volume: 8 B per document per query, ordered on the device (counting "
                         "sort writes doc[] | score[]) and copied over PCIe into the pinned arrays "
                         "the caller reads, in sub-batches whose copies overlap the next pass; the "
                         "copy alone takes %.3f ms per query at 55 GB/s, K2 %.2f ms per query at the "
                         "HBM roofline"
                         % (int(roff[-1]) // n_fpr * 8 / 55e9 * 1e3,
                            info.bytes_per_kmer * 1000 / env.peak / 1e6)}
            del roff, doc, score
        # one search() per call, the `cobs query <string>` pattern (src/cobs.cpp:417-422)
        blob1, off1 = make_batch(7200, 200)
        pin1 = torch.from_numpy(blob1).pin_memory().numpy()
        lat = {}
        for name, thr, limit in (("threshold0.8", THRESHOLD, 0), ("threshold0_limit10", 0.0, 10)):
            ts = []
            for i in range(200):
                q = pin1[i * QUERY_LEN:(i + 1) * QUERY_LEN]
                t0 = time.perf_counter()
                index.search_packed(q, off1[:2], thr, limit, raw=True)
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts[20:]) * 1e6
            lat[name] = {"p50_us": float(np.percentile(ts, 50)), "p90_us": float(np.percentile(ts, 90)),
                         "p99_us": float(np.percentile(ts, 99))}
        # a single LONG query (a gene against the index): k-split score kernel + counting sort
        for name, qlen in (("kmers1000_threshold0.8", 1030), ("kmers10000_threshold0.8", 10_030)):
            blob_l, off_l = make_batch(7300, 24, qlen)
            pin_l = torch.from_numpy(blob_l).pin_memory().numpy()
            ts = []
            for i in range(24):
                q = pin_l[i * qlen:(i + 1) * qlen]
                t0 = time.perf_counter()
                index.search_packed(q, off_l[:2], THRESHOLD, 0, raw=True)
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts[4:]) * 1e6
            lat[name] = {"p50_us": float(np.percentile(ts, 50)), "p90_us": float(np.percentile(ts, 90))}
        lat["kernel_floor_us"] = ("one CTA streams %d row slices per tile; a lone CTA is "
                                  "latency-bound" % (T_KMERS * cfg["h"]))
        out["single_query_latency"] = lat
    return out


def run_load(env, gb):
    """Index loader throughput: a classic index of ~gb GB is written in the
    reference's file format to /dev/shm (page-cache speed, no disk in the way) and loaded with
    cobsgpu_index_open_file; reports GB/s of matrix bytes into HBM."""
    import cobs_b200
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    free = shutil.disk_usage(base).free
    gb = min(gb, free / 1e9 / 2.5)
    if gb < 0.5:
        return {"unavailable": "no scratch space under %s" % base}
    n_docs, h = 100_000, 3
    rows = int(gb * 1e9 / 12_500)
    path = os.path.join(base, "cobs_b200_load_%d.cobs_classic" % os.getpid())
    try:
        g = cobs_b200.GpuIndex.procedural(0, n_docs, [rows], h, fill_seed=5, device=env.local_rank)
        t0 = time.perf_counter()
        g.save(path)
        save_s = time.perf_counter() - t0
        probe = g.read_row(0, rows // 2, 0, 12_500)
        g.close()
        size = os.path.getsize(path)
        best, lib_s, threads = None, None, None
        for _ in range(2):
            t0 = time.perf_counter()
            g = cobs_b200.GpuIndex.open_file(path, device=env.local_rank)
            dt = time.perf_counter() - t0
            ok = bool(np.array_equal(g.read_row(0, rows // 2, 0, 12_500), probe))
            if best is None or dt < best:
                best, lib_s, threads = dt, g.info.load_seconds, g.info.load_threads
            g.close()
        return {"file_gb": size / 1e9, "seconds": best, "gbs": size / best / 1e9,
                "stream_seconds": lib_s, "stream_gbs": size / lib_s / 1e9 if lib_s else None,
                "save_gbs": size / save_s / 1e9, "roundtrip_ok": ok,
                "source": base + " (page cache)", "host_threads": threads,
                "note": "seconds = the whole cobsgpu_index_open_file call (header, cudaMalloc, "
                        "stream); stream_seconds = first read to last byte in HBM"}
    except Exception as e:   # reported, never load-bearing
        return {"unavailable": "failed: %r" % (e,)}
    finally:
        if os.path.exists(path):
            os.unlink(path)


def run_ours(args):
    env = Env(args)
    world, rank = env.world, env.rank
    wl = args.workload
    primary = run_workload(env, wl, True, args.parallelism)
    secondary = {}
    if not args.no_secondary and wl == "cfg4" and not args.nq and not args.rows:
        names = ["cfg2"]
        if world == 1:
            names.append("cfg3")
        if world == 8:
            names.append("cfg5")
        for name in names:
            try:
                secondary[name] = run_workload(env, name, False, "auto")
            except Exception as e:   # a secondary block must not take the headline down
                secondary[name] = {"unavailable": "failed: %r" % (e,)}
                env.torch.cuda.empty_cache()
    load = None
    if world == 1 and not args.no_load:
        load = run_load(env, args.load_gb)

    if rank == 0:
        cfg = WORKLOADS[wl]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, cpu = cpu_reference_run(wl, 1, 1, 0, budget_s=args.cpu_seconds)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:   # the baseline is reported, never load-bearing
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable",
                       "sample": "failed: %r" % (e,)}
        hits = primary["legs"]["hits"]
        line = {
            "metric": METRIC, "value": primary["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": primary["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": primary["workload"] + ", threshold %.2f (hits leg)" % cfg["thr_hits"],
                       "parallelism": primary["parallelism"],
                       "queries_per_step": primary["queries_per_step"],
                       "kmers_per_query": T_KMERS, "bytes_per_kmer": primary["bytes_per_kmer"],
                       "hbm_bytes_this_rank": primary["hbm_bytes_this_rank"],
                       "l2_policy": "each step a different batch; rows touched per step "
                                    "(%.1f GB) >> 126 MB L2" % (hits["algorithmic_bytes_per_launch"] / 1e9),
                       "results_last_step": primary["results_last_step"]},
            "queries_per_s": primary["queries_per_s"],
            "roofline": primary["roofline"],
            "e2e": primary["e2e"],
            "gpu_launches": primary["gpu_launches"],
            "legs": primary["legs"],
            "parity_check": primary["parity_check"],
            "clocks": hits.get("clocks"),
        }
        if "patterns" in primary:
            line["patterns"] = primary["patterns"]
        if secondary:
            line["secondary"] = secondary
        if load is not None:
            line["load"] = load
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    ok = True
    for blk in [primary] + [b for b in secondary.values() if "parity_check" in b]:
        pc = blk.get("parity_check")
        if pc is not None and not pc["ok"]:
            ok = False
    if env.sampler:
        env.sampler.stop()
    if world > 1:
        env.dist.destroy_process_group()
    if not ok:
        log("PARITY CHECK FAILED")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--nq", type=int, default=0, help="queries per step (default: workload's)")
    ap.add_argument("--rows", type=int, default=0, help="override signature rows (debug)")
    ap.add_argument("--results-per-query", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-queries", type=int, default=100,
                    help="queries per step of the --impl reference arm")
    ap.add_argument("--parity-queries", type=int, default=64)
    ap.add_argument("--load-gb", type=float, default=24.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-patterns", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-load", action="store_true")
    ap.add_argument("--parallelism", default="docs", choices=["docs", "queries", "grid", "auto"],
                    help="N > 1, headline workload: shard the document axis (default, BASELINE.json's "
                         "north_star), replicate the index and split every query batch, both "
                         "(grid: --doc-shards ranks share one copy), or auto (from the index size)")
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
