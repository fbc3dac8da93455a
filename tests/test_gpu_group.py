"""GPU tests of the multi-GPU group behind the C ABI (cobsgpu_group_*): one process, the
document axis sharded over several devices, the leader GPU merging the shards' result blocks.
The protocol is exercised on ONE device too (several shards of a group may share a GPU), so
these run on every box; the tests that need real peers skip on single-GPU boxes and are run with
`gpurun --gpus 2`."""
import os
import subprocess

import numpy as np
import pytest

import cobs_b200
from cobs_b200 import GpuGroup, GpuIndex, KIND_CLASSIC, KIND_COMPACT, _lib
from conftest import ROOT, golden_path
from oracle import oracle

pytestmark = pytest.mark.gpu

N_GPUS = cobs_b200.lib().cobsgpu_device_count()
COBS = os.path.join(ROOT, "build", "cobs")


def rq(seed, length):
    return oracle.random_query(seed, length)


def as_list(res):
    doc, score = res
    return [(0, int(d), int(s)) for d, s in zip(doc, score)]


def device_sets():
    sets = [[0, 0, 0]]
    if N_GPUS >= 2:
        sets.append([0, 1])
    if N_GPUS >= 4:
        sets.append([0, 1, 2, 3])
    return sets


@pytest.mark.parametrize("devices", device_sets())
@pytest.mark.parametrize("shape", [
    (KIND_CLASSIC, 20_000, [1009], 3, 0),
    (KIND_CLASSIC, 300, [211], 2, 0),            # fewer 128-document granules than shards
    (KIND_COMPACT, 40_000, [61, 97, 31, 43, 59], 3, 1024),
])
def test_group_lists_match_oracle(devices, shape):
    kind, n_docs, sig, h, ps = shape
    g = GpuGroup.procedural(kind, n_docs, sig, h, devices, page_size=ps, fill_seed=11)
    o = oracle.Index.procedural(kind, n_docs, sig, h, page_size=ps, fill_seed=11, materialize=True)
    qs = [rq(i, L) for i, L in enumerate([100, 100, 131, 31, 285, 286, 1030, 60])]
    # fused path (threshold), top-k epilogue (limit), per-shard exhaustive + host merge (0, all)
    for thr, k in ((0.05, 0), (0.3, 0), (0.0, 10), (0.02, 3), (0.0, 0), (0.0, 2000), (0.9, 0)):
        got = g.search_batch(qs, thr, k)
        for q, r in zip(qs, got):
            assert as_list(r) == oracle.search(o, q, thr, k), (devices, thr, k, len(q))
    # candidate overflow on the shards: flagged by the merge, redone exhaustively, nothing lost
    g.set_option("max_candidates", 8)
    for q, r in zip(qs, g.search_batch(qs, 0.02, 0)):
        assert as_list(r) == oracle.search(o, q, 0.02, 0)
    g.set_option("max_candidates", 1024)
    # many small batches through the ring
    g.set_option("max_batch", 3)
    many = [rq(100 + i, 100 + (i % 4) * 17) for i in range(40)]
    for q, r in zip(many, g.search_batch(many, 0.04, 0)):
        assert as_list(r) == oracle.search(o, q, 0.04, 0)
    # a query beyond 16 bit-planes takes the per-shard host path
    g.set_option("max_batch", 16384)
    huge = [rq(7, 66_000), rq(8, 100)]
    for q, r in zip(huge, g.search_batch(huge, 0.3, 0)):
        assert as_list(r) == oracle.search(o, q, 0.3, 0)
    # errors surface like on a single handle
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        g.search_batch([rq(1, 100), rq(2, 50) + b"N" + rq(3, 50)], 0.5, 0)
    assert e.value.code == _lib.ERR_INVALID_BASE and "(query 1)" in e.value.msg
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        g.search_batch([b"ACGT"], 0.5, 0)
    assert e.value.code == _lib.ERR_QUERY_TOO_SHORT
    assert as_list(g.search_batch([qs[0]], 0.05, 0)[0]) == oracle.search(o, qs[0], 0.05, 0)
    g.close()


@pytest.mark.parametrize("devices", device_sets())
def test_group_on_golden_files(golden, devices):
    """the reference's own result lists, index files loaded shard by shard"""
    n = 0
    for case in golden["cases"]:
        if len(case["files"]) != 1:
            continue
        g = GpuGroup.open_file(golden_path(case["files"][0]), devices)
        for c in case["cases"]:
            total_hashes = g.info.num_hashes * (len(c["query"]) - g.info.term_size + 1)
            if total_hashes <= 1:
                continue          # the no-sort quirk lives in the host classes
            got = g.search_batch([c["query"]], c["threshold"], c["num_results"])[0]
            assert as_list(got) == [tuple(x) for x in c["result"]], (case["name"], c["threshold"])
            n += 1
        g.close()
    assert n >= 60


@pytest.mark.skipif(N_GPUS < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_cli_gpus_2_matches_golden(golden):
    """`cobs query --gpus 2`: stdout byte-identical to the reference's lists"""
    n = 0
    for case in golden["cases"]:
        for c in case["cases"][::4]:
            cmd = [COBS, "query", "--gpus", "2"]
            for f in case["files"]:
                cmd += ["-i", golden_path(f)]
            cmd += ["-t", repr(c["threshold"]), "-l", str(c["num_results"]), c["query"]]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                               timeout=120)
            assert r.returncode == 0, r.stderr
            want = "".join("%s\t%d\n" % (case["doc_names"][f][d], s) for f, d, s in c["result"])
            assert r.stdout == want, (case["name"], c["threshold"], c["num_results"])
            n += 1
    assert n >= 25


@pytest.mark.skipif(N_GPUS < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_group_equals_single_gpu_at_size():
    """a 3 GB classic index: 2-GPU group == one GPU, hits-bearing threshold and top-k"""
    n_docs, sig, h = 200_000, [120_011], 3
    one = GpuIndex.procedural(KIND_CLASSIC, n_docs, sig, h, fill_seed=3)
    two = GpuGroup.procedural(KIND_CLASSIC, n_docs, sig, h, [0, 1], fill_seed=3)
    rng = np.random.default_rng(5)
    nq = 3000
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nq * 100)].copy()
    off = np.arange(nq + 1, dtype=np.uint64) * 100
    for thr, k in ((0.1, 0), (0.0, 10), (0.06, 0)):
        a = one.search_packed(blob, off, thr, k, raw=True)
        b = two.search_packed(blob, off, thr, k, raw=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (thr, k)
        assert a[0][-1] > 0
    one.close()
    two.close()
