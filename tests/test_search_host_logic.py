"""CPU test of the host logic in cobs_b200.Search (multi-index merge, per-index limits, the
reference's no-sort quirk for single-hash queries, name lookup): the per-index device call is
replaced by a stand-in built on the oracle with the C ABI's semantics (thresholded list, ordered
(score desc, doc asc), cut at num_results -- never the quirk), and every golden case of the real
reference must come out right."""
import numpy as np

import cobs_b200
from cobs_b200 import api
from oracle import oracle
from conftest import golden_path


class OracleBackedIndex:
    """what GpuIndex exposes to Search, computed on the CPU by the oracle"""

    def __init__(self, path):
        self.ix = oracle.Index.load(path)
        self.term_size = self.ix.term_size
        self.num_hashes = self.ix.num_hashes
        self.counts_size = self.ix.counts_size
        self._names = self.ix.doc_names

    def doc_name(self, d):
        return self._names[d]

    def search_batch(self, queries, threshold=0.0, num_results=0):
        out = []
        for q in queries:
            sc = self.ix.scores(q)
            T = len(q) - self.term_size + 1
            need = int(np.ceil(threshold * T))
            docs = [d for d in range(self.ix.n_docs) if sc[d] >= need]
            docs.sort(key=lambda d: (-int(sc[d]), d))
            if num_results:
                docs = docs[:num_results]
            out.append((np.array(docs, dtype=np.uint32), np.array([sc[d] for d in docs], np.uint32)))
        return out

    def close(self):
        pass


def test_search_class_reproduces_every_golden_case(golden, monkeypatch):
    monkeypatch.setattr(api.GpuIndex, "open_file",
                        classmethod(lambda cls, p, device=0, **kw: OracleBackedIndex(p)))
    n = 0
    for case in golden["cases"]:
        s = cobs_b200.Search([golden_path(f) for f in case["files"]])
        # all cases of a fixture in one batch, grouped by (threshold, num_results)
        groups = {}
        for c in case["cases"]:
            groups.setdefault((c["threshold"], c["num_results"]), []).append(c)
        for (thr, k), cs in groups.items():
            got = s.search_batch([c["query"] for c in cs], thr, k)
            for c, g in zip(cs, got):
                want = [(case["doc_names"][f][d], sc) for f, d, sc in c["result"]]
                assert [(r.doc_name, r.score) for r in g] == want, (case["name"], thr, k)
                n += 1
        s.close()
    assert n >= 140
