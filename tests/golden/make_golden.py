#!/usr/bin/env python3
"""Generates tests/golden/* with the UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make ref`).  Run from the repo root, in the build container only:

    python tests/golden/make_golden.py

Every index file is written by the reference's own construction code
(cobs::classic_construct / compact_construct / classic_construct_random) and every expected
result list comes out of cobs::ClassicSearch::search through its public API.  The fixtures
are small (tens of KB) and committed; the GPU box never needs /root/reference.

Cases follow the reference's own query tests (tests/classic_index_query.cpp,
tests/compact_index_query.cpp, tests/test_util.hpp, python/tests/test_cobs_index.py).
"""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_DATA = "/root/reference/tests/data"


def docs_all(query, num_documents=33, num_terms=1000000):
    """tests/test_util.hpp:42-62: document j holds query k-mer i iff j % (i % (n-1) + 1) == 0"""
    pos = [[] for _ in range(num_documents)]
    for i in range(min(num_terms, len(query) - 31)):
        for j in range(num_documents):
            if j % (i % (num_documents - 1) + 1) == 0:
                pos[j].append(i)
    return pos


def docs_one(num_documents=33):
    """tests/test_util.hpp:66-84: document i holds the first k-mer, i*10+1 times"""
    return [[0] * (i * 10 + 1) for i in range(num_documents)]


def write_docs(dirname, seq, positions, prefix=""):
    os.makedirs(dirname, exist_ok=True)
    for i, pos in enumerate(positions):
        name = "%sdocument_%06d" % (prefix, i)
        ref.write_kmer_doc(os.path.join(dirname, name + ".cobs_doc"), name, seq, pos)


def run_cases(paths, queries, params):
    s = ref.Search(paths)
    names = [[s.doc_name(f, d) for d in range(s.n_docs[f])] for f in range(s.n_files)]
    results = []
    for q in queries:
        for thr, k in params:
            results.append({"query": q.decode(), "threshold": thr, "num_results": k,
                            "result": s.search(q, thr, k)})
    s.close()
    return names, results


def main():
    tmp = tempfile.mkdtemp(prefix="cobs_golden_")
    cases = []
    params = [(0.0, 0), (0.0, 5), (0.3, 0), (0.8, 0), (0.8, 3), (1.0, 0), (0.05, 1)]

    def emit(name, files, queries, prm=params):
        rel = []
        for f in files:
            dst = os.path.join(OUT, os.path.basename(f))
            shutil.copyfile(f, dst)
            rel.append(os.path.basename(f))
        names, results = run_cases(files, queries, prm)
        cases.append({"name": name, "files": rel, "doc_names": names, "cases": results})
        print(name, [os.path.getsize(f) for f in files], len(results))

    # 1. python/tests/test_cobs_index.py: the FASTA fixtures, default parameters
    fasta = os.path.join(tmp, "fasta")
    shutil.copytree(os.path.join(REF_DATA, "fasta"), fasta)
    os.system("chmod -R u+w " + fasta)
    py_classic = os.path.join(tmp, "python_test.cobs_classic")
    ref.classic_construct(fasta, py_classic, os.path.join(tmp, "t1"))
    q_py = b"AGTCAACGCTAAGGCATTTCCCCCCTGCCTCCTGCCTGCTGCCAAGCCCT"
    emit("python_classic", [py_classic], [q_py])
    py_compact = os.path.join(tmp, "python_test.cobs_compact")
    ref.compact_construct(fasta, py_compact, os.path.join(tmp, "t2"), page_size=16)
    emit("python_compact", [py_compact], [q_py])
    emit("python_two_indices", [py_classic, py_compact], [q_py])

    # 2. classic_index_query.all_included (33 docs, h=3, fpr 0.1): 8-bit and 16-bit score paths
    q160 = ref.random_sequence(160, 1)
    q1000 = ref.random_sequence(1000, 2)
    d = os.path.join(tmp, "all160")
    write_docs(d, q160, docs_all(q160))
    f = os.path.join(tmp, "all160.cobs_classic")
    ref.classic_construct(d, f, os.path.join(tmp, "t3"), num_hashes=3, fpr=0.1)
    emit("classic_all_160", [f], [q160, q160[:31], q160[5:100], ref.random_sequence(100, 7)])
    d = os.path.join(tmp, "all1000")
    write_docs(d, q1000, docs_all(q1000))
    f = os.path.join(tmp, "all1000.cobs_classic")
    ref.classic_construct(d, f, os.path.join(tmp, "t4"), num_hashes=3, fpr=0.1)
    emit("classic_all_1000", [f], [q1000, q1000[:300], q1000[100:131]])

    # 3. compact_index_query.all_included_mmap_small: page_size 2 -> 3 pages of 16 documents
    d = os.path.join(tmp, "call160")
    write_docs(d, q160, docs_all(q160))
    f = os.path.join(tmp, "all160.cobs_compact")
    ref.compact_construct(d, f, os.path.join(tmp, "t5"), num_hashes=3, fpr=0.1, page_size=2)
    emit("compact_all_160", [f], [q160, q160[3:90], ref.random_sequence(64, 3)])
    d = os.path.join(tmp, "call1000")
    write_docs(d, q1000, docs_all(q1000))
    f = os.path.join(tmp, "all1000.cobs_compact")
    ref.compact_construct(d, f, os.path.join(tmp, "t6"), num_hashes=3, fpr=0.1, page_size=2)
    emit("compact_all_1000", [f], [q1000, q1000[:255 + 30], q1000[:256 + 30]])

    # 4. one_included_large_batch_multi_index: 33/44/55 documents in three indices
    files = []
    for n in (33, 44, 55):
        d = os.path.join(tmp, "one%d" % n)
        write_docs(d, q160, docs_one(n))
        f = os.path.join(tmp, "one%d.cobs_classic" % n)
        ref.classic_construct(d, f, os.path.join(tmp, "t7_%d" % n), num_hashes=3, fpr=0.1)
        files.append(f)
    emit("multi_index_one", files, [q160, q160[:31]])

    # 5. single hash in total (h=1, query length == k): the reference skips the sort
    d = os.path.join(tmp, "one_h1")
    write_docs(d, q160, docs_one(20))
    f = os.path.join(tmp, "one_h1.cobs_classic")
    ref.classic_construct(d, f, os.path.join(tmp, "t8"), num_hashes=1, fpr=0.3)
    emit("single_hash_no_sort", [f], [q160[:31], q160[1:32], q160[:40]],
         [(0.0, 0), (0.0, 2), (0.5, 3), (1.0, 0)])

    # 6. classic_construct_random: dense random bits, many documents (ragged 203), canonicalize
    f = os.path.join(tmp, "random203.cobs_classic")
    ref.classic_construct_random(f, 1500, 203, 400, 3, 42)
    emit("classic_random_203", [f],
         [ref.random_sequence(100, 100 + i) for i in range(6)] + [ref.random_sequence(400, 9)],
         [(0.0, 0), (0.0, 10), (0.02, 0), (0.05, 7)])

    # 7. construction parity (scope row f4): synthetic FASTA documents -- several records per
    # file, wrapped lines, non-ACGT characters (hashed with zeros by the reference, not skipped),
    # lower case, a record shorter than k, ';' comments, empty lines -- indexed by the
    # reference's classic_construct.  The FASTA files themselves are committed as fixtures.
    docs_dir = os.path.join(OUT, "construct_docs")
    shutil.rmtree(docs_dir, ignore_errors=True)
    os.makedirs(docs_dir)
    import random
    rnd = random.Random(11)

    def seq(n, alphabet="ACGT"):
        return "".join(rnd.choice(alphabet) for _ in range(n))

    fasta_docs = {
        "alpha": [("r1", seq(400)), ("r2", seq(95))],
        "beta": [("only", seq(1500))],
        "gamma": [("with_n", seq(200) + "NNNN" + seq(150) + "RY" + seq(80)), ("short", seq(20))],
        "delta": [("lower", seq(120) + seq(60).lower() + seq(120))],
        "epsilon": [("a", seq(31)), ("b", seq(32)), ("c", seq(30)), ("d", seq(700))],
        "zeta": [("x", seq(64, "AC")), ("y", seq(900))],
        "eta": [("protein", seq(300, "ACDEFGHIKLMNPQRSTVWY"))],
        "theta": [("r", seq(333))],
        "iota": [("r", seq(2000))],
    }
    for name, recs in fasta_docs.items():
        with open(os.path.join(docs_dir, name + ".fasta"), "w") as f:
            for i, (rn, sq) in enumerate(recs):
                f.write(("%s%s %s\n" % (">" if i % 2 == 0 else ";", name, rn)))
                width = 60 if i % 2 == 0 else 71
                for j in range(0, len(sq), width):
                    f.write(sq[j:j + width] + "\n")
                if i % 2 == 1:
                    f.write("\n")
    tmp_docs = os.path.join(tmp, "construct_docs")
    shutil.copytree(docs_dir, tmp_docs)
    construct_cases = []
    for tag, h, fpr, canon in (("h3", 3, 0.1, 1), ("h1", 1, 0.3, 1), ("raw", 2, 0.2, 0)):
        f = os.path.join(tmp, "construct_%s.cobs_classic" % tag)
        ref.classic_construct(tmp_docs, f, os.path.join(tmp, "tc_" + tag), num_hashes=h, fpr=fpr,
                              canonicalize=canon)
        shutil.copyfile(f, os.path.join(OUT, os.path.basename(f)))
        construct_cases.append({"file": os.path.basename(f), "num_hashes": h,
                                "false_positive_rate": fpr, "canonicalize": canon})
        print("construct", tag, os.path.getsize(f))
    for fn in os.listdir(docs_dir):      # the reference drops .cobs_cache files next to inputs
        if fn.endswith(".cobs_cache"):
            os.unlink(os.path.join(docs_dir, fn))

    # known answers for K1: canonicalisation vectors and XXH64 of k-mers
    kats = {"canonicalize": [], "xxh64": []}
    for k in (15, 31, 32, 33, 64):
        seq = ref.random_sequence(k + 20, k)
        for i in range(0, 20, 3):
            km = seq[i:i + k]
            out, good = ref.canonicalize_kmer(km)
            kats["canonicalize"].append({"kmer": km.decode(), "out": out.decode(), "good": good})
            for seed in (0, 1, 2, 5):
                kats["xxh64"].append({"data": out.decode(), "seed": seed,
                                      "hash": "%016x" % ref.xxh64(out, seed)})
    for km in (b"AGTCAACGCTAAGGCATTTCCCCCCTGCCTN", b"NGTCAACGCTAAGGCATTTCCCCCCTGCCTC", b"acgt", b"AT",
               b"ACGT", b"TTTT", b"AAAAAAAAAAAAAAAATTTTTTTTTTTTTTT"):
        out, good = ref.canonicalize_kmer(km)
        kats["canonicalize"].append({"kmer": km.decode(), "out": out.hex(), "good": good,
                                     "hex": True})
    with open(os.path.join(OUT, "golden.json"), "w") as fp:
        json.dump({"cases": cases, "kats": kats, "construct": construct_cases}, fp, indent=0)
    shutil.rmtree(tmp)
    print("wrote", os.path.join(OUT, "golden.json"))


if __name__ == "__main__":
    main()
