"""GPU test of `cobs benchmark-fpr`, the reference's query micro-benchmark (src/cobs.cpp:605-730),
re-hosted on the GPU path.  Kept in its own file so that it runs after the parity suites."""
import os
import subprocess

import pytest

from conftest import ROOT, golden_path

pytestmark = pytest.mark.gpu

COBS = os.path.join(ROOT, "build", "cobs")


def run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=120, **kw)


def test_cli_benchmark_fpr_matches_oracle_distribution():
    """`cobs benchmark-fpr` (src/cobs.cpp:605-730): same mt19937 query stream as the reference
    (warm-up queries drawn first), threshold 0, all results; the score histogram printed with -d
    must equal the oracle's over the same queries"""
    import collections
    import numpy as np
    from oracle import oracle

    def mt_queries(seed, n, length):
        bg = np.random.MT19937()
        bg._legacy_seeding(seed)
        raw = bg.random_raw(n * length)
        b = np.frombuffer(b"ACGT", dtype=np.uint8)[(raw % 4).astype(np.int64)]
        return [b[i * length:(i + 1) * length].tobytes() for i in range(n)]

    idx = golden_path("random203.cobs_classic")
    n_warm, n_q, kmers, seed = 3, 25, 70, 7
    qs = mt_queries(seed, n_warm + n_q, kmers + 30)[n_warm:]
    o = oracle.Index.load(idx)
    want = collections.Counter()
    for q in qs:
        for _, _, sc in oracle.search(o, q, 0.0, 0):
            want[sc] += 1
    for batch in ("1", "8"):
        r = run([COBS, "benchmark-fpr", "-k", str(kmers), "-q", str(n_q), "-w", str(n_warm),
                 "--seed", str(seed), "-d", "--batch", batch, idx])
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0].startswith("RESULT name=benchmark ")
        assert " kmer_queries=70 queries=25 warmup=3 results=203 " in lines[0]
        got = {}
        for l in lines[1:]:
            f = dict(x.split("=") for x in l.split()[1:])
            assert f["name"] == "benchmark_fpr"
            got[int(f["fpr"])] = int(f["dist"])
        assert got == dict(want)


def test_index_save_round_trips_reference_files(tmp_path, golden):
    """cobsgpu_index_save writes the reference's formats: loading a file written by the reference
    and saving it again must reproduce it byte for byte (classic and compact, incl. the compact
    header's padding rule, cobs/file/compact_index_header.cpp:20-42)"""
    from cobs_b200 import GpuIndex
    names = sorted({f for case in golden["cases"] for f in case["files"]})
    assert any(n.endswith(".cobs_compact") for n in names)
    for n in names:
        g = GpuIndex.open_file(golden_path(n))
        out = str(tmp_path / n)
        g.save(out)
        g.close()
        assert open(out, "rb").read() == open(golden_path(n), "rb").read(), n
