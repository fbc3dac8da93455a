"""Boundary proof (CPU, build container only): pieces of the REFERENCE's own callers are compiled
-- unmodified, extracted from /root/reference at test time, never copied into the repo -- against
the drop-in headers under cobs_b200/host/include and linked against libcobs_b200.so:

  * process_query() and the index-opening loop of query() from src/cobs.cpp (the `cobs query`
    driver, lines 410-469 and 505-525),
  * the query halves (everything after index construction) of three gtests from
    tests/classic_index_query.cpp and tests/compact_index_query.cpp, with a few-line stand-in for
    the gtest macros and the generated documents.

A maintainer who relinks those callers against this library gets the same source to compile.
Skipped on the GPU box (no /root/reference there)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cobs")),
                                reason="reference sources not present")

PRELUDE = r'''
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/query/search.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>
#include <cobs/util/fs.hpp>

#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// tlx/die.hpp's die(): a stream expression, then terminate
#define die(msg) do { std::ostringstream oss__; oss__ << msg; cobs::die_with_message(oss__.str()); } while (0)
// the three gtest macros the extracted bodies use
#define ASSERT_EQ(a, b) do { if (!((a) == (b))) std::abort(); } while (0)
#define ASSERT_GE(a, b) do { if (!((a) >= (b))) std::abort(); } while (0)
#define ASSERT_LE(a, b) do { if (!((a) <= (b))) std::abort(); } while (0)
namespace fs = cobs::fs;
// stand-ins for what the construction half of the tests leaves behind
struct FakeDoc { std::vector<int> d; const std::vector<int>& data() const { return d; } };
static std::vector<FakeDoc> documents(33);
static fs::path index_path = "index.cobs_classic";
static fs::path index_file = "index.cobs_compact";
static std::string query(50000, 'A');
static size_t num_documents = 33;
'''


def between(text, start, end, include_end=True):
    a = text.index(start)
    b = text.index(end, a)
    return text[a:b + (len(end) if include_end else 0)]


def compile_and_link(tmp_path, name, source):
    src = tmp_path / (name + ".cpp")
    src.write_text(source)
    exe = tmp_path / name
    cmd = ["g++", "-std=c++17", "-Wall", "-Wno-unused-variable", "-Wno-unused-function",
           "-I" + os.path.join(ROOT, "cobs_b200", "host", "include"), "-I" + os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L" + os.path.join(ROOT, "build"), "-lcobs_b200",
           "-L" + os.path.join(ROOT, "cobs_b200", "lib"), "-lcobsgpu",
           "-Wl,-rpath," + os.path.join(ROOT, "build"), "-Wl,-rpath," + os.path.join(ROOT, "cobs_b200", "lib")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_reference_process_query_compiles_against_the_drop_in_headers(tmp_path):
    cobs_cpp = open(os.path.join(REF, "src", "cobs.cpp")).read()
    process_query = between(cobs_cpp, "static inline\nvoid process_query(", "    s.timer().print(\"search\");\n}")
    opening = between(cobs_cpp, "    std::vector<std::shared_ptr<cobs::IndexSearchFile> > indices;",
                      "    process_query(s, threshold, num_results, query, query_file);")
    source = PRELUDE.replace("static std::string query(50000, 'A');", "") + process_query + '''

int main(int argc, char** argv) {
    std::vector<std::string> index_files(argv + 1, argv + argc);
    std::string query, query_file;
    double threshold = 0.8;
    unsigned num_results = 0;
''' + opening + '''
    return 0;
}
'''
    exe = compile_and_link(tmp_path, "ref_process_query", source)
    # no index, no query: the reference's own error path, no GPU needed
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode != 0 and "Pass a verbatim query or a query file." in r.stderr


def test_reference_query_gtest_bodies_compile_against_the_drop_in_headers(tmp_path):
    bodies = []
    classic = open(os.path.join(REF, "tests", "classic_index_query.cpp")).read()
    compact = open(os.path.join(REF, "tests", "compact_index_query.cpp")).read()
    for text, test, start in (
            (classic, "TEST_F(classic_index_query, all_included_small_batch)", "    cobs::ClassicSearch s_base("),
            (classic, "TEST_F(classic_index_query, one_included_small_batch)", "    cobs::ClassicSearch s_base("),
            (compact, "TEST_F(compact_index_query, one_included_mmap)", "    cobs::ClassicSearch s_base(")):
        t = text[text.index(test):]
        t = t[:t.index("\n}\n") + 3]
        body = t[t.index(start):]
        assert "search(query, result" in body and body.rstrip().endswith("}")
        bodies.append("void body_%d() {\n%s\n" % (len(bodies), body))
    source = PRELUDE + "\n".join(bodies) + "\nint main() { return 0; }\n"
    compile_and_link(tmp_path, "ref_query_gtests", source)


def test_custom_index_search_file_subclass_compiles(tmp_path):
    """IndexSearchFile is the reference's abstract interface again: a caller's own subclass
    compiles against it, and in-memory pages plug in through HbmIndexSearchFile(Pages)"""
    source = PRELUDE + r'''
class MyFile : public cobs::IndexSearchFile {
public:
    void read_from_disk(const std::vector<size_t>&, uint8_t*, size_t, size_t, size_t) override { }
    uint32_t term_size() const override { return 31; }
    uint8_t canonicalize() const override { return 1; }
    uint64_t row_size() const override { return 2; }
    uint64_t page_size() const override { return 1; }
    uint64_t num_hashes() const override { return 1; }
    uint64_t counts_size() const override { return 16; }
    const std::vector<std::string>& file_names() const override { return names; }
    std::vector<std::string> names;
};
int main() {
    std::shared_ptr<cobs::IndexSearchFile> f = std::make_shared<MyFile>();
    cobs::ClassicSearch s(f);                       // constructs; searching it dies with a message
    cobs::HbmIndexSearchFile::Pages p;              // the plug point for custom row sources
    p.file_names = { "a", "b" };
    return f->counts_size() == 16 && p.num_hashes == 1 ? 0 : 1;
}
'''
    exe = compile_and_link(tmp_path, "custom_subclass", source)
    r = subprocess.run([exe], timeout=60)
    assert r.returncode == 0
