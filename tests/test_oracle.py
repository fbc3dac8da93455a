"""CPU tests: the oracle (oracle/cobs_oracle.c) against the reference's own known answers,
against the committed golden vectors (made by the real reference, tests/golden/make_golden.py)
and -- when oracle/_ref is present -- against the real reference live."""
import os
import random

import numpy as np
import pytest

from oracle import oracle, ref
from conftest import golden_path

PRIME = 2654435761


def _sanity_buffer():
    # extlib/xxhash/xxhsum.c:442-452
    buf, g = bytearray(), PRIME
    for _ in range(101):
        buf.append((g >> 24) & 0xFF)
        g = (g * g) & 0xFFFFFFFF
    return bytes(buf)


# extlib/xxhash/xxhsum.c:463-470
XXH64_KATS = [
    (0, 0, 0xEF46DB3751D8E999), (0, PRIME, 0xAC75FDA2929B17EF),
    (1, 0, 0x4FCE394CC88952D8), (1, PRIME, 0x739840CB819FA723),
    (14, 0, 0xCFFA8DB881BC3A3D), (14, PRIME, 0x5B9611585EFCC9CB),
    (101, 0, 0x0EAB543384F878AD), (101, PRIME, 0xCAA65939306F1E21),
]


def test_xxh64_known_answers():
    buf = _sanity_buffer()
    for n, seed, expect in XXH64_KATS:
        assert oracle.xxh64(buf[:n], seed) == expect


def test_canonicalize_reference_vectors():
    # tests/util.cpp:28-59
    vec = [
        ("AGTCAACGCTAAGGCATTTCCCCCCTGCCTC", "AGTCAACGCTAAGGCATTTCCCCCCTGCCTC"),
        ("GAGGCAGGGGGGAAATGCCTTAGCGTTGACT", "AGTCAACGCTAAGGCATTTCCCCCCTGCCTC"),
        ("AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"),
        ("TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"),
    ]
    for kmer, expect in vec:
        out, good = oracle.canonicalize_kmer(kmer)
        assert good and out.decode() == expect
    out, good = oracle.canonicalize_kmer("AGTCAACGCTAAGGCATTTCCCCCCTGCCTN")
    assert not good


def test_canonical_is_min_of_kmer_and_revcomp():
    # tests/parameters.cpp:107-122
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rng = random.Random(5)
    for _ in range(3000):
        k = rng.choice([1, 2, 15, 30, 31, 32, 33, 64])
        s = "".join(rng.choice("ACGT") for _ in range(k))
        rc = "".join(comp[c] for c in reversed(s))
        out, good = oracle.canonicalize_kmer(s)
        # only the outer k/2 base pairs are compared (cobs/util/query.cpp:155-189): when they
        # all tie -- possible for odd k only -- the forward strand is kept whatever the
        # centre base is ("palindrome to the centre", tests/util.cpp:52-58)
        expect = s if s[:k // 2] == rc[:k // 2] else min(s, rc)
        assert good and out.decode() == expect


def test_golden_kats(golden):
    for c in golden["kats"]["canonicalize"]:
        out, good = oracle.canonicalize_kmer(c["kmer"])
        expect = bytes.fromhex(c["out"]) if c.get("hex") else c["out"].encode()
        assert out == expect and good == c["good"]
    for c in golden["kats"]["xxh64"]:
        assert "%016x" % oracle.xxh64(c["data"], c["seed"]) == c["hash"]


def test_golden_result_lists(golden):
    """full (file, doc, score) lists of the real reference == oracle, every case"""
    n = 0
    for case in golden["cases"]:
        idx = [oracle.Index.load(golden_path(f)) for f in case["files"]]
        for f, ix in enumerate(idx):
            assert ix.doc_names == case["doc_names"][f]
        for c in case["cases"]:
            got = oracle.search(idx, c["query"], c["threshold"], c["num_results"])
            assert got == [tuple(r) for r in c["result"]], (case["name"], c["threshold"],
                                                           c["num_results"])
            n += 1
    assert n > 100


def test_python_known_answer(golden):
    # python/tests/test_cobs_index.py:36-40, 57-61
    for f in ("python_test.cobs_classic", "python_test.cobs_compact"):
        ix = oracle.Index.load(golden_path(f))
        r = oracle.search(ix, "AGTCAACGCTAAGGCATTTCCCCCCTGCCTCCTGCCTGCTGCCAAGCCCT")
        assert len(r) == 7
        assert ix.doc_names[r[0][1]] == "sample1" and r[0][2] == 20


def test_error_codes():
    ix = oracle.Index.load(golden_path("all160.cobs_classic"))
    with pytest.raises(oracle.OracleError) as e:
        oracle.search(ix, "ACGT")
    assert e.value.code == oracle.ERR_TOO_SHORT
    with pytest.raises(oracle.OracleError) as e:
        oracle.search(ix, "ACGTN" * 10)
    assert e.value.code == oracle.ERR_INVALID_BASE
    with pytest.raises(oracle.OracleError) as e:
        oracle.Index.load(golden_path("golden.json"))
    assert e.value.code == oracle.ERR_BAD_FILE


def test_procedural_matches_materialized(tmp_path):
    """procedural bits == materialised bits == bits written to a file and loaded back"""
    for kind, n_docs, sig, ps in ((oracle.KIND_CLASSIC, 203, [977], 0),
                                  (oracle.KIND_COMPACT, 100, [311, 57, 1000, 13], 4)):
        proc = oracle.Index.procedural(kind, n_docs, sig, 3, page_size=ps, fill_seed=11)
        mat = oracle.Index.procedural(kind, n_docs, sig, 3, page_size=ps, fill_seed=11,
                                      materialize=True)
        p = str(tmp_path / ("x%d.idx" % kind))
        proc.write(p)
        loaded = oracle.Index.load(p)
        assert loaded.signature_sizes == sig and loaded.n_docs == n_docs
        for seed in range(4):
            q = oracle.random_query(seed, 90)
            a = proc.scores(q)
            assert np.array_equal(a, mat.scores(q))
            assert np.array_equal(a, loaded.scores(q))
            assert oracle.search(proc, q, 0.0, 0) == oracle.search(loaded, q, 0.0, 0)
        bits = np.unpackbits(mat.page_array(0))
        assert 0.2 < bits.mean() < 0.3   # Bernoulli(1/4)


needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@needs_ref
def test_oracle_vs_reference_primitives():
    buf = _sanity_buffer()
    for n, seed, expect in XXH64_KATS:
        assert ref.xxh64(buf[:n], seed) == expect
    rng = random.Random(1)
    for _ in range(1500):
        n = rng.randrange(0, 200)
        d = bytes(rng.randrange(256) for _ in range(n))
        s = rng.randrange(2 ** 64)
        assert oracle.xxh64(d, s) == ref.xxh64(d, s)
    for _ in range(4000):
        n = rng.randrange(1, 70)
        d = bytes(rng.choice(b"ACGTACGTACGTNacgt") for _ in range(n))
        assert oracle.canonicalize_kmer(d) == ref.canonicalize_kmer(d)


@needs_ref
def test_oracle_vs_reference_live(tmp_path):
    """fresh random indices written by the oracle's writers, searched by both"""
    rng = np.random.default_rng(3)
    configs = [
        (oracle.KIND_CLASSIC, 1, [64], 0, 1, 1),
        (oracle.KIND_CLASSIC, 7, [100], 0, 2, 1),
        (oracle.KIND_CLASSIC, 129, [333], 0, 3, 0),
        (oracle.KIND_CLASSIC, 1000, [517], 0, 4, 1),
        (oracle.KIND_COMPACT, 50, [100, 200, 50, 77], 2, 3, 1),
        (oracle.KIND_COMPACT, 600, [331, 400, 123], 32, 1, 1),
        (oracle.KIND_COMPACT, 20, [97], 3, 2, 0),
    ]
    for ci, (kind, n_docs, sig, ps, h, canon) in enumerate(configs):
        ix = oracle.Index.procedural(kind, n_docs, sig, h, page_size=ps, canonicalize=canon,
                                     fill_seed=ci)
        p = str(tmp_path / ("c%d.cobs_%s" % (ci, "classic" if kind == 0 else "compact")))
        ix.write(p)
        s = ref.Search(p)
        assert s.counts_size() == ix.counts_size
        for qi in range(5):
            L = [31, 32, 100, 286, 300][qi]
            q = oracle.random_query(100 * ci + qi, L)
            for thr, k in [(0.0, 0), (0.0, 3), (0.1, 0), (0.5, 2)]:
                assert s.search(q, thr, k) == oracle.search(ix, q, thr, k), (ci, qi, thr, k)
        s.close()


@needs_ref
def test_reference_score_width_variants_agree(tmp_path):
    """the u8/u16/u32 (+SSE2) expansion variants all give the oracle's counts
    (tests/compact_index_query.cpp:93-154)"""
    ix = oracle.Index.procedural(oracle.KIND_COMPACT, 70, [211, 97, 150], 3, page_size=3,
                                 fill_seed=5)
    p = str(tmp_path / "v.cobs_compact")
    ix.write(p)
    s = ref.Search(p)
    q = oracle.random_query(1, 150)
    want = oracle.search(ix, q)
    try:
        for flags in [(False, True, True, True), (True, False, True, True),
                      (True, False, True, False), (True, True, False, True),
                      (True, True, False, False)]:
            ref.set_disable(*flags)
            assert s.search(q) == want
    finally:
        ref.set_disable(False, False, False, False)
        s.close()
