"""GPU parity on BASELINE.json's configurations that round 1 left open:

cfg1  literally: the reference's own `classic_construct_random` builds the 128-document,
      1 048 576-row, h=3 classic index (src/cobs.cpp:250-253); 100 random 100-bp queries; full
      result lists at -t 0 / -t 0.05 / -t 0.8 / -l 5 against the UNMODIFIED reference
      (oracle/_ref), through the C ABI, the C++ classes and -- stdout byte for byte -- against
      the reference's own `cobs query` binary.
cfg3  at the size bench.py times (196 613-row base, 158.7 GB).
cfg5  per-shard parity: shards 0, 3 and 7 of the 8-way document split of the 10 M-document,
      77-page compact index (75 GB each; every shard holds one 2048-byte column slice of every
      page), sampled column blocks against the procedural oracle.
"""
import os
import subprocess

import numpy as np
import pytest

import cobs_b200
from cobs_b200 import GpuIndex, KIND_COMPACT, _lib
from conftest import ROOT
from oracle import oracle, ref

pytestmark = pytest.mark.gpu

COBS = os.path.join(ROOT, "build", "cobs")
SEED = 20260101


def ref_cli():
    p = ref.lib_path().replace("libcobs_ref_", "cobs_ref_").replace(".so", "")
    return p if os.path.exists(p) else None


@pytest.fixture(scope="module")
def cfg1(tmp_path_factory):
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    d = tmp_path_factory.mktemp("cfg1")
    path = str(d / "cfg1.cobs_classic")
    # cobs classic-construct-random -n 128 -s 1048576 -m 100000 --num-hashes 3 --seed 1
    ref.classic_construct_random(path, 1_048_576, 128, 100_000, 3, 1)
    queries = [ref.random_sequence(100, 1000 + i) for i in range(100)]
    return path, queries, d


def test_cfg1_lists_match_the_reference(cfg1):
    path, queries, _ = cfg1
    assert os.path.getsize(path) > 16 * 1024 * 1024
    r = ref.Search(path)
    g = GpuIndex.open_file(path)
    assert g.n_docs == 128 and g.num_hashes == 3 and g.signature_size(0) == 1_048_576
    n_hits = 0
    for thr, k in ((0.0, 0), (0.05, 0), (0.8, 0), (0.0, 5), (0.03, 5)):
        got = g.search_batch(queries, thr, k)
        for q, (doc, score) in zip(queries, got):
            want = r.search(q, thr, k)
            assert [(0, int(d), int(s)) for d, s in zip(doc, score)] == want, (thr, k)
            n_hits += len(want)
    assert n_hits > 100 * 128
    # exhaustive per-document counts == the C restatement on the same file
    o = oracle.Index.load(path)
    sc = g.scores(queries[:10])
    for q, a in zip(queries[:10], sc):
        assert np.array_equal(a, o.scores(q))
    g.close()
    r.close()


def test_cfg1_cli_stdout_equals_the_reference_cli(cfg1):
    path, queries, d = cfg1
    exe = ref_cli()
    if exe is None:
        pytest.skip("reference CLI not built (make -C oracle ref)")
    fasta = str(d / "queries.fa")
    with open(fasta, "w") as f:
        for i, q in enumerate(queries):
            f.write(">query_%03d\n%s\n" % (i, q.decode()))
    for extra in (["-t", "0"], ["-t", "0.8"], ["-t", "0", "-l", "5"], ["-t", "0.05"], []):
        want = subprocess.run([exe, "query", "-i", path, "-f", fasta] + extra,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        got = subprocess.run([COBS, "query", "-i", path, "-f", fasta] + extra,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert want.returncode == 0 and got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, extra
        assert len(want.stdout.splitlines()) >= 100
    # and a verbatim query
    q = queries[0].decode()
    want = subprocess.run([exe, "query", "-i", path, "-t", "0", q], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=120)
    got = subprocess.run([COBS, "query", "-i", path, "-t", "0", q], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=120)
    assert got.stdout == want.stdout and len(got.stdout.splitlines()) == 128


def open_or_skip(*a, **kw):
    try:
        return GpuIndex.procedural(*a, **kw)
    except cobs_b200.CobsGpuError as e:
        if e.code == _lib.ERR_OOM:
            pytest.skip("not enough device memory for the full-size index: " + e.msg)
        raise


def check_blocks(g, o, queries, blocks):
    got = g.scores(queries)
    for q, a in zip(queries, got):
        for b0 in blocks:
            b0 -= b0 % 8
            b1 = min(b0 + 1024, o.counts_size)
            assert np.array_equal(a[b0:b1], o.scores(q, b0, b1)), b0
    return got


def test_cfg3_compact_at_the_benched_size():
    n_docs, ps, h = 1_000_000, 16_384, 4
    sig = [int(196_613 * 1.5 ** p) for p in range(8)]          # bench.py's cfg3: 158.7 GB
    g = open_or_skip(KIND_COMPACT, n_docs, sig, h, page_size=ps, fill_seed=SEED)
    o = oracle.Index.procedural(oracle.KIND_COMPACT, n_docs, sig, h, page_size=ps, fill_seed=SEED)
    assert g.info.hbm_bytes > 158e9 and g.info.bytes_per_kmer == 4 * 8 * 16_384
    queries = [oracle.random_query(i, 100) for i in range(2)]
    got = check_blocks(g, o, queries, (0, 131_072 - 512, 131_072, 5 * 131_072 + 4096,
                                       1_000_000 - 64, 8 * 131_072 - 1024))
    for q, a, (doc, score) in zip(queries, got, g.search_batch(queries, 0.07, 0)):
        keep = np.nonzero(a[:n_docs] >= 5)[0]
        order = sorted(keep.tolist(), key=lambda d: (-int(a[d]), d))
        assert doc.tolist() == order and score.tolist() == [int(a[d]) for d in order]
    # top-k epilogue on the compact layout
    for q, a, (doc, score) in zip(queries, got, g.search_batch(queries, 0.0, 10)):
        order = sorted(range(n_docs), key=lambda d: (-int(a[d]), d))[:10]
        assert doc.tolist() == order
    g.close()


@pytest.mark.parametrize("shard", [0, 3, 7])
def test_cfg5_shard_parity(shard):
    """one of the eight document shards of cfg5 on one GPU (~75 GB): the shard holds the same
    2048-byte column slice (16 384 documents) of each of the 77 pages"""
    n_docs, ps, h = 10_000_000, 16_384, 4
    sig = [int(100_003 * 1.0345 ** p) for p in range(77)]
    g = open_or_skip(KIND_COMPACT, n_docs, sig, h, page_size=ps, fill_seed=SEED,
                     shard_index=shard, shard_count=8)
    o = oracle.Index.procedural(oracle.KIND_COMPACT, n_docs, sig, h, page_size=ps, fill_seed=SEED)
    per_page, per_slice = 8 * ps, 8 * ps // 8
    assert g.info.bytes_per_kmer == h * 77 * (ps // 8)
    q = [oracle.random_query(50 + shard, 100)]
    a = g.scores(q)[0]
    other = (shard + 1) % 8
    for p in (0, 38, 76):
        s0 = p * per_page + shard * per_slice
        for b0 in (s0, s0 + per_slice // 2, s0 + per_slice - 1024):
            b1 = min(b0 + 1024, o.counts_size)
            assert np.array_equal(a[b0:b1], o.scores(q[0], b0, b1)), (p, b0)
        # columns of another shard are left untouched
        o0 = p * per_page + other * per_slice
        assert not a[o0:o0 + 1024].any()
    # lists of this shard: global document ids, only real documents, ordered
    doc, score = g.search_batch(q, 0.07, 0)[0]
    mine = (np.arange(o.counts_size) % per_page) // per_slice == shard
    keep = np.nonzero((a[:n_docs] >= 5) & mine[:n_docs])[0].tolist()
    order = sorted(keep, key=lambda d: (-int(a[d]), d))
    assert doc.tolist() == order and score.tolist() == [int(a[d]) for d in order]
    g.close()
