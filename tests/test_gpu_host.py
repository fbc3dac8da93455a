"""GPU tests of the C++ drop-in layer (cobs::ClassicSearch & friends) and the `cobs query` CLI:
stdout must be byte-identical to what the reference prints for the golden cases
(format: src/cobs.cpp:418-462)."""
import os
import subprocess

import pytest

from conftest import ROOT, golden_path, GOLDEN_DIR

pytestmark = pytest.mark.gpu

COBS = os.path.join(ROOT, "build", "cobs")
HOST_TESTS = os.path.join(ROOT, "build", "host_tests")


def run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=120, **kw)


def expected_lines(case, c):
    return "".join("%s\t%d\n" % (case["doc_names"][f][d], s) for f, d, s in c["result"])


def test_host_api_checks():
    r = run([HOST_TESTS, GOLDEN_DIR])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_tests: ok" in r.stdout


def test_too_short_query_exits_like_the_reference():
    # classic_search.cpp:431-433: message on stderr, exit(EXIT_FAILURE)
    r = run([HOST_TESTS, GOLDEN_DIR, "too_short"])
    assert r.returncode == 1
    assert "query too short, needs to be at least 31 characters long" in r.stderr
    assert "not reached" not in r.stderr


def test_cli_single_queries_match_golden(golden):
    n = 0
    for case in golden["cases"]:
        # a spread of thresholds / limits per fixture, not all (each call boots a CUDA context)
        for c in case["cases"][::5]:
            cmd = [COBS, "query"]
            for f in case["files"]:
                cmd += ["-i", golden_path(f)]
            cmd += ["-t", repr(c["threshold"]), "-l", str(c["num_results"]), c["query"]]
            r = run(cmd)
            assert r.returncode == 0, r.stderr
            assert r.stdout == expected_lines(case, c), (case["name"], c["threshold"])
            assert "TIMER info=search" in r.stderr
            n += 1
    assert n >= 20


def test_cli_fasta_file_batches(golden, tmp_path):
    """-f <fasta>: '*comment\\tcount' + result lines per record, multi-line records joined,
    ';' comments accepted, default threshold 0.8"""
    case = next(c for c in golden["cases"] if c["name"] == "classic_all_1000")
    by_query = {}
    for c in case["cases"]:
        if c["threshold"] == 0.8 and c["num_results"] == 0:
            by_query[c["query"]] = c
    fasta = tmp_path / "q.fa"
    want = ""
    with open(fasta, "w") as f:
        for i, (q, c) in enumerate(by_query.items()):
            f.write(("%squery number %d\n" % (">" if i % 2 == 0 else ";", i)))
            for j in range(0, len(q), 60):       # wrapped sequence lines
                f.write(q[j:j + 60] + "\n")
            f.write("\n")
            want += "*query number %d\t%d\n" % (i, len(c["result"])) + expected_lines(case, c)
    for batch in ("1", "2", "4096"):
        r = run([COBS, "query", "-i", golden_path(case["files"][0]), "-f", str(fasta),
                 "--batch", batch])
        assert r.returncode == 0, r.stderr
        assert r.stdout == want


def test_cli_errors():
    r = run([COBS, "query", "-i", golden_path("golden.json"), "ACGT" * 10])
    assert r.returncode != 0 and "Could not open index path" in r.stderr
    r = run([COBS, "query", "-i", golden_path("all160.cobs_classic")])
    assert r.returncode != 0 and "Pass a verbatim query or a query file." in r.stderr
    r = run([COBS, "query", "-i", golden_path("all160.cobs_classic"), "ACGT"])
    assert r.returncode == 1 and "query too short" in r.stderr

