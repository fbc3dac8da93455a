"""GPU tests of the device-side classic construction (scope row f4): the index built by
cobsgpu_construct_classic from the committed FASTA fixtures must be BYTE-IDENTICAL to the file
the reference's classic_construct wrote from the same documents (tests/golden/make_golden.py),
including k-mers with non-ACGT characters, which the reference hashes with zero bytes."""
import os

import numpy as np
import pytest

import cobs_b200
from cobs_b200 import GpuIndex
from oracle import oracle
from conftest import GOLDEN_DIR, golden_path

pytestmark = pytest.mark.gpu


def read_fasta(path):
    """sequence records of a FASTA file the way the reference walks it
    (cobs/fasta_file.hpp:156-182): '>' / ';' lines and empty lines end a record, the other
    lines of a record are concatenated"""
    recs, cur = [], []
    for line in open(path, "rb").read().split(b"\n"):
        if not line or line[:1] in (b">", b";"):
            if cur:
                recs.append(b"".join(cur))
            cur = []
        else:
            cur.append(line)
    if cur:
        recs.append(b"".join(cur))
    return recs


def documents(names):
    d = os.path.join(GOLDEN_DIR, "construct_docs")
    return [(n, read_fasta(os.path.join(d, n + ".fasta"))) for n in names]


def test_constructed_index_is_byte_identical_to_the_reference(golden, tmp_path):
    assert len(golden["construct"]) == 3
    for c in golden["construct"]:
        want_path = golden_path(c["file"])
        want = oracle.Index.load(want_path)
        docs = documents(want.doc_names)          # the header's order is authoritative
        g = GpuIndex.construct_classic(docs, num_hashes=c["num_hashes"],
                                       false_positive_rate=c["false_positive_rate"],
                                       canonicalize=c["canonicalize"])
        # signature sizing: calc_signature_size on the largest document
        assert g.signature_size() == want.signature_sizes[0]
        assert g.n_docs == want.n_docs and g.num_hashes == c["num_hashes"]
        out = str(tmp_path / c["file"])
        g.save(out)
        assert open(out, "rb").read() == open(want_path, "rb").read(), c["file"]
        # and it is searchable right away, like the file loaded back
        q = docs[1][1][0][40:140]
        loaded = GpuIndex.open_file(out)
        a = g.search_batch([q], 0.0, 0)[0]
        b = loaded.search_batch([q], 0.0, 0)[0]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        got = [(0, int(d), int(s)) for d, s in zip(*a)]
        assert got == oracle.search(want, q, 0.0, 0)
        assert got[0][1] == 1 and got[0][2] == 70      # the source document matches fully
        g.close()
        loaded.close()


def test_construct_then_query_roundtrip_properties():
    """every k-mer of a document is found in that document (no false negatives), explicit
    signature_size, compact save is rejected only for shards"""
    rng = np.random.default_rng(4)
    docs = []
    for i in range(70):
        n = int(rng.integers(1, 4))
        docs.append(("doc%03d" % i,
                     [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(31, 400))).tolist())
                      for _ in range(n)]))
    g = GpuIndex.construct_classic(docs, num_hashes=3, signature_size=20011)
    assert g.signature_size() == 20011 and g.n_docs == 70
    for i in (0, 17, 69):
        for part in docs[i][1]:
            doc, score = g.search_batch([part], 1.0, 0)[0]
            assert i in doc.tolist()                    # threshold 1.0: all k-mers present
    # identical to the oracle's view of the saved file
    g.close()


def test_cli_classic_construct_is_byte_identical(golden, tmp_path):
    """`cobs classic-construct <dir> <out>` (the reference's subtool, src/cobs.cpp:117-190) on the
    committed FASTA fixtures writes the very bytes the reference wrote"""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "build", "cobs")
    docs = os.path.join(GOLDEN_DIR, "construct_docs")
    for c in golden["construct"]:
        out = str(tmp_path / c["file"])
        cmd = [exe, "classic-construct", docs, out, "-h", str(c["num_hashes"]),
               "-f", repr(c["false_positive_rate"])]
        if not c["canonicalize"]:
            cmd.append("--no-canonicalize")
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == open(golden_path(c["file"]), "rb").read(), c["file"]
        # refuses to overwrite without -C, like the reference
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode != 0 and "will not overwrite" in r.stderr
        r = subprocess.run(cmd + ["-C"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
    # and the result answers queries through the same CLI
    seq = "".join(l.strip() for l in open(os.path.join(docs, "beta.fasta")) if not l.startswith(">"))
    r = subprocess.run([exe, "query", "-i", out, "-t", "0.9", seq[40:140]], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout == "beta\t70\n", r.stdout + r.stderr
