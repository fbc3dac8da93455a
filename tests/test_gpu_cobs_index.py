"""The `Search` half of the reference's python/tests/test_cobs_index.py, run against the
`cobs_index` module of this repo (same import name, same calls, same expectations), plus the
construct-then-query flow of that test served by the device-side classic construction."""
import os
import unittest

import pytest

import cobs_index as cobs
from conftest import GOLDEN_DIR, golden_path

pytestmark = pytest.mark.gpu

cobs.disable_cache()


class MainTest(unittest.TestCase):
    # the index files below were written by the reference from its tests/data/fasta
    # (tests/golden/make_golden.py); the queries and expectations are the reference test's own
    def test_classic_query(self):
        index_file = golden_path("python_test.cobs_classic")
        self.assertTrue(os.path.isfile(index_file))
        s = cobs.Search(index_file)
        r = s.search("AGTCAACGCTAAGGCATTTCCCCCCTGCCTCCTGCCTGCTGCCAAGCCCT")
        self.assertEqual(len(r), 7)
        self.assertEqual(r[0].doc_name, "sample1")
        self.assertEqual(r[0].score, 20)

    def test_compact_query(self):
        index_file = golden_path("python_test.cobs_compact")
        self.assertTrue(os.path.isfile(index_file))
        s = cobs.Search(index_file)
        r = s.search("AGTCAACGCTAAGGCATTTCCCCCCTGCCTCCTGCCTGCTGCCAAGCCCT")
        self.assertEqual(len(r), 7)
        self.assertEqual(r[0].doc_name, "sample1")
        self.assertEqual(r[0].score, 20)

    def test_doc_list(self):
        l1 = cobs.DocumentList(os.path.join(GOLDEN_DIR, "construct_docs"))
        self.assertEqual(l1.size(), 9)
        l2 = cobs.DocumentList()
        l2.add_recursive(os.path.join(GOLDEN_DIR, "construct_docs"))
        self.assertEqual(l2.size(), 9)

    def test_classic_construct_query(self):
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            index_file = os.path.join(d, "python_test.cobs_classic")
            p = cobs.ClassicIndexParameters()
            p.clobber = True
            p.num_hashes = 3
            p.false_positive_rate = 0.1
            cobs.classic_construct(input=os.path.join(GOLDEN_DIR, "construct_docs"),
                                   out_file=index_file, index_params=p)
            self.assertTrue(os.path.isfile(index_file))
            # byte-identical to what the reference's classic_construct wrote from these files
            with open(index_file, "rb") as a, open(golden_path("construct_h3.cobs_classic"), "rb") as b:
                self.assertEqual(a.read(), b.read())
            s = cobs.Search(index_file)
            with open(os.path.join(GOLDEN_DIR, "construct_docs", "beta.fasta")) as f:
                seq = "".join(l.strip() for l in f if not l.startswith(">"))
            r = s.search(seq[40:140])
            self.assertEqual(r[0].doc_name, "beta")
            self.assertEqual(r[0].score, 70)
            r = s.search(seq[40:140], threshold=0.9, num_results=1)
            self.assertEqual([(x.doc_name, x.score) for x in r], [("beta", 70)])
