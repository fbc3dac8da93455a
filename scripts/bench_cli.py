#!/usr/bin/env python3
"""End-to-end wall time of `cobs query -f <fasta>` (the reference's CLI pattern for batched
queries) on a 24 GB classic index file in /dev/shm: index load + FASTA parse + GPU batches +
printing.  Prints the phases the CLI reports on stderr and the wall clock."""
import os
import subprocess
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cobs_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gb = float(os.environ.get("GB", "24"))
nq = int(os.environ.get("NQ", "500000"))
thr = os.environ.get("THR", "0.1")
rows = int(gb * 1e9 / 12_500)
idx = "/dev/shm/cli_bench.cobs_classic"
fa = "/dev/shm/cli_bench.fa"
g = cobs_b200.GpuIndex.procedural(0, 100_000, [rows], 3, fill_seed=5)
g.save(idx)
g.close()
blob, _ = bench.make_batch(99, nq)
raw = blob.tobytes()
with open(fa, "w") as f:
    for i in range(nq):
        f.write(">q%d\n%s\n" % (i, raw[i * 100:(i + 1) * 100].decode()))
env = dict(os.environ, COBS_CLI_TRACE="1")
for batch in os.environ.get("BATCHES", "16384,4096,65536,16384").split(","):
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "build", "cobs"), "query", "-i", idx, "-f", fa, "-t", thr,
                        "--batch", batch], stdout=open("/dev/shm/cli_bench.out", "w"), stderr=subprocess.PIPE,
                       text=True, env=env)
    dt = time.perf_counter() - t0
    out_bytes = os.path.getsize("/dev/shm/cli_bench.out")
    lines = [l for l in r.stderr.strip().splitlines() if l.startswith(("TIMER", "CLI "))]
    print("batch %s: wall %.2f s (%.0f queries/s incl. load), stdout %.1f MB, rc %d\n    %s" % (
        batch, dt, nq / dt, out_bytes / 1e6, r.returncode, "\n    ".join(l[:300] for l in lines)), flush=True)
for p in (idx, fa, "/dev/shm/cli_bench.out"):
    os.unlink(p)
