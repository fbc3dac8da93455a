"""Property tests (hypothesis) of the oracle itself: the size-independent laws the GPU full-size
tests rely on must hold for the CPU restatement on arbitrary small indices."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle


@st.composite
def index_and_query(draw):
    kind = draw(st.sampled_from([oracle.KIND_CLASSIC, oracle.KIND_COMPACT]))
    h = draw(st.integers(1, 4))
    k = draw(st.sampled_from([8, 15, 31, 32]))
    canon = draw(st.integers(0, 1))
    seed = draw(st.integers(0, 2 ** 32))
    if kind == oracle.KIND_CLASSIC:
        n_docs = draw(st.integers(1, 300))
        sig, ps = [draw(st.integers(2, 64))], 0
    else:
        ps = draw(st.sampled_from([1, 2, 3, 8]))
        pages = draw(st.integers(1, 4))
        n_docs = draw(st.integers(8 * ps * (pages - 1) + 1, 8 * ps * pages))
        sig = [draw(st.integers(2, 64)) for _ in range(pages)]
    ix = oracle.Index.procedural(kind, n_docs, sig, h, page_size=ps, term_size=k,
                                 canonicalize=canon, fill_seed=seed, materialize=True)
    length = draw(st.integers(k, k + 120))
    q = oracle.random_query(draw(st.integers(0, 2 ** 32)), length)
    return ix, q


@settings(max_examples=60, deadline=None)
@given(index_and_query(), st.data())
def test_scores_are_additive_over_kmer_partitions(iq, data):
    ix, q = iq
    k = ix.term_size
    T = len(q) - k + 1
    if T < 2:
        return
    a = data.draw(st.integers(1, T - 1))        # first a k-mers | the rest
    whole = ix.scores(q)
    left = ix.scores(q[:a + k - 1])
    right = ix.scores(q[a:])
    assert np.array_equal(whole, left + right)
    assert int(whole.max()) <= T


@settings(max_examples=60, deadline=None)
@given(index_and_query(), st.floats(0.0, 1.0), st.integers(0, 40))
def test_result_list_laws(iq, thr, limit):
    ix, q = iq
    T = len(q) - ix.term_size + 1
    if T * ix.num_hashes <= 1:
        return                                   # the reference's no-sort quirk, tested elsewhere
    scores = [int(x) for x in ix.scores(q)]
    full = oracle.search(ix, q, thr, 0)
    cut = oracle.search(ix, q, thr, limit)
    need = math.ceil(thr * T)
    # exactly the real documents at or above the threshold, best first, ties by document
    want = sorted((d for d in range(ix.n_docs) if scores[d] >= need), key=lambda d: (-scores[d], d))
    assert [d for _, d, _ in full] == want
    assert all(s == scores[d] for _, d, s in full)
    assert cut == (full if limit == 0 else full[:limit])
    # lowering the threshold can only add documents
    lower = oracle.search(ix, q, thr / 2, 0)
    assert {d for _, d, _ in full} <= {d for _, d, _ in lower}


@settings(max_examples=40, deadline=None)
@given(index_and_query(), st.integers(2, 5))
def test_column_blocks_tile_the_score_vector(iq, parts):
    """scores over [b0, b1) column blocks (what the sharded GPU path computes) concatenate to the
    full vector"""
    ix, q = iq
    cs = ix.counts_size
    cuts = sorted({0, cs} | {(cs * i // parts) // 8 * 8 for i in range(1, parts)})
    whole = ix.scores(q)
    got = np.concatenate([ix.scores(q, a, b) for a, b in zip(cuts, cuts[1:])])
    assert np.array_equal(whole, got)
