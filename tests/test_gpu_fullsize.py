"""GPU tests at BASELINE.json's full sizes (SURVEY.md section 8d realisation): the matrices are
procedural (>= 100 GB, never materialised on the host), so parity is checked on sampled rows,
sampled queries over ALL documents, sampled column blocks, and through size-independent
properties (additivity over k-mer partitions, batch independence, threshold monotonicity,
host path == device path)."""
import numpy as np
import pytest

import cobs_b200
from cobs_b200 import GpuIndex, KIND_CLASSIC, KIND_COMPACT, _lib
from oracle import oracle

pytestmark = pytest.mark.gpu

SEED = 20260101


def rq(seed, length):
    return oracle.random_query(seed, length)


def open_or_skip(*a, **kw):
    try:
        return GpuIndex.procedural(*a, **kw)
    except cobs_b200.CobsGpuError as e:
        if e.code == _lib.ERR_OOM:
            pytest.skip("not enough device memory for the full-size index: " + e.msg)
        raise


def lists_equal(res, want):
    doc, score = res
    return [(0, int(d), int(s)) for d, s in zip(doc, score)] == want


def test_cfg2_classic_100k_docs_105GB():
    n_docs, sig, h = 100_000, [8_388_593], 3
    g = open_or_skip(KIND_CLASSIC, n_docs, sig, h, fill_seed=SEED)
    o = oracle.Index.procedural(oracle.KIND_CLASSIC, n_docs, sig, h, fill_seed=SEED)
    assert g.info.hbm_bytes > 100e9 and g.info.bytes_per_kmer == 37_500
    row = 12_500
    # sampled rows of the 105 GB matrix == the procedural definition
    for r in (0, 1, 4_194_304, 8_388_592, 7_777_777):
        got = g.read_row(0, r, 0, row)
        want = np.array([oracle.fill_word(SEED, 0, r, w) for w in range((row + 7) // 8)],
                        dtype="<u8").view(np.uint8)[:row]
        assert np.array_equal(got, want)
    # sampled queries, every document
    queries = [rq(i, 100) for i in range(6)] + [rq(77, 400)]
    got = g.scores(queries)
    for q, a in zip(queries, got):
        assert np.array_equal(a, o.scores(q))
    # ordered lists: thr 4/70 overflows the 1024 candidate slots (exhaustive + radix path),
    # thr 7/70 stays on the fused path
    for thr, k in ((0.05, 0), (0.1, 0), (0.05, 25), (0.0, 3)):
        for q, r in zip(queries[:3], g.search_batch(queries[:3], thr, k)):
            assert lists_equal(r, oracle.search(o, q, thr, k)), (thr, k)

    # additivity: the k-mers of q[0:170] are exactly those of q[0:100] plus those of q[70:170]
    q = rq(123, 170)
    s = g.scores([q, q[:100], q[70:]])
    assert np.array_equal(s[0], s[1] + s[2])

    # the benchmark batch: 10 000 queries; results do not depend on the batch they ride in,
    # are reproducible, and grow monotonically as the threshold drops
    rng = np.random.default_rng(1000)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=10_000 * 100)]
    off = np.arange(10_001, dtype=np.uint64) * 100
    off1, doc1, sc1 = g.search_packed(blob, off, 0.1, 0, raw=True)
    off2, doc2, sc2 = g.search_packed(blob, off, 0.1, 0, raw=True)
    assert np.array_equal(off1, off2) and np.array_equal(doc1, doc2) and np.array_equal(sc1, sc2)
    assert off1[-1] > 1000          # thr 7: a handful of documents per query
    raw = blob.tobytes()
    sample = [5, 1234, 9999]
    alone = g.search_batch([raw[i * 100:(i + 1) * 100] for i in sample], 0.1, 0)
    for i, (d, s_) in zip(sample, alone):
        a, b = int(off1[i]), int(off1[i + 1])
        assert np.array_equal(d, doc1[a:b]) and np.array_equal(s_, sc1[a:b])
        assert lists_equal((d, s_), oracle.search(o, raw[i * 100:(i + 1) * 100], 0.1, 0))
    off3, doc3, sc3 = g.search_packed(blob[:100 * 200], off[:201], 0.08, 0, raw=True)
    for i in range(200):
        hi = set(doc1[int(off1[i]):int(off1[i + 1])].tolist())
        lo = set(doc3[int(off3[i]):int(off3[i + 1])].tolist())
        assert hi <= lo

    # device-resident entry point == host entry point
    torch = pytest.importorskip("torch")
    nq, rpq = 2000, 64
    d_q = torch.from_numpy(blob[:nq * 100].copy()).cuda()
    counts = torch.zeros(nq, dtype=torch.int32, device="cuda")
    keys = torch.zeros((nq, rpq), dtype=torch.int64, device="cuda")
    g.search_device(d_q.data_ptr(), off[:nq + 1], 0.1, 0, rpq, counts.data_ptr(), keys.data_ptr(),
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    c = counts.cpu().numpy()
    kk = keys.cpu().numpy().view(np.uint64)
    for i in range(nq):
        a, b = int(off1[i]), int(off1[i + 1])
        assert c[i] == b - a
        d, s_ = cobs_b200.decode_keys(kk[i, :c[i]])
        assert np.array_equal(d, doc1[a:b]) and np.array_equal(s_, sc1[a:b])
    g.close()


def test_cfg4_classic_1M_docs_131GB():
    n_docs, sig, h = 1_000_000, [1_048_573], 3
    g = open_or_skip(KIND_CLASSIC, n_docs, sig, h, fill_seed=SEED)
    o = oracle.Index.procedural(oracle.KIND_CLASSIC, n_docs, sig, h, fill_seed=SEED)
    assert g.info.bytes_per_kmer == 375_000
    queries = [rq(i, 100) for i in range(3)]
    got = g.scores(queries)
    # sampled 1024-document column blocks incl. the first and the last (ragged) one
    for q, a in zip(queries, got):
        for b0 in (0, 524_288, 999_936 - 1024, 999_936):
            b1 = min(b0 + 1024, o.counts_size)
            assert np.array_equal(a[b0:b1], o.scores(q, b0, b1))
    q = rq(5, 170)
    s = g.scores([q, q[:100], q[70:]])
    assert np.array_equal(s[0], s[1] + s[2])
    # lists: every reported document has the oracle's score, order is (score desc, doc asc),
    # and the number of documents above the threshold matches the exhaustive scores
    for q, a, (doc, score) in zip(queries, got, g.search_batch(queries, 0.1, 0)):
        keep = np.nonzero(a[:n_docs] >= 7)[0]
        order = sorted(keep.tolist(), key=lambda d: (-int(a[d]), d))
        assert doc.tolist() == order and score.tolist() == [int(a[d]) for d in order]
    g.close()


def test_cfg3_compact_1M_docs_8_pages():
    # 8 pages of 16 384 B (131 072 documents each), growing signature sizes; base scaled to
    # ~80 GB so that it fits beside the workspaces
    n_docs, ps, h = 1_000_000, 16_384, 4
    sig = [int(98_307 * 1.5 ** p) for p in range(8)]
    g = open_or_skip(KIND_COMPACT, n_docs, sig, h, page_size=ps, fill_seed=SEED)
    o = oracle.Index.procedural(oracle.KIND_COMPACT, n_docs, sig, h, page_size=ps,
                                fill_seed=SEED)
    assert g.info.bytes_per_kmer == 4 * 8 * 16_384 and g.counts_size == 8 * 131_072
    queries = [rq(i, 100) for i in range(2)]
    got = g.scores(queries)
    for q, a in zip(queries, got):
        for b0 in (0, 131_072 - 512, 131_072, 5 * 131_072 + 4096, 1_000_000 - 64, 8 * 131_072 - 1024):
            b1 = b0 + 1024 if b0 + 1024 <= o.counts_size else o.counts_size
            b0 -= b0 % 8
            assert np.array_equal(a[b0:b1], o.scores(q, b0, b1))
    for q, a, (doc, score) in zip(queries, got, g.search_batch(queries, 0.05, 0)):
        keep = np.nonzero(a[:n_docs] >= 4)[0]      # padded columns (>= n_docs) never reported
        order = sorted(keep.tolist(), key=lambda d: (-int(a[d]), d))
        assert doc.tolist() == order and score.tolist() == [int(a[d]) for d in order]
    g.close()
