// host_tests.cpp -- API-level checks of the C++ drop-in classes, in the style of the reference's
// query gtests (tests/classic_index_query.cpp, tests/compact_index_query.cpp).  Needs a GPU.
//   host_tests <golden dir>
#include <cobs/file/file_io_exception.hpp>
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

static int g_failed = 0;
#define CHECK(cond)                                                               \
    do {                                                                          \
        if (!(cond)) {                                                            \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++g_failed;                                                           \
        }                                                                         \
    } while (0)

static const char* kPyQuery = "AGTCAACGCTAAGGCATTTCCCCCCTGCCTCCTGCCTGCTGCCAAGCCCT";

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: host_tests <golden dir>\n");
        return 2;
    }
    const std::string dir = argv[1];
    using namespace cobs;

    // header sniffing (src/cobs.cpp:509-521 relies on it)
    CHECK(file_has_header<ClassicIndexHeader>(dir + "/python_test.cobs_classic"));
    CHECK(!file_has_header<CompactIndexHeader>(dir + "/python_test.cobs_classic"));
    CHECK(file_has_header<CompactIndexHeader>(dir + "/python_test.cobs_compact"));
    CHECK(!file_has_header<ClassicIndexHeader>(dir + "/golden.json"));

    // python/tests/test_cobs_index.py:36-40, 57-61 through the C++ API, both constructors
    for (const char* f : { "/python_test.cobs_classic", "/python_test.cobs_compact" }) {
        ClassicSearch s(dir + f);
        std::vector<SearchResult> r;
        s.search(kPyQuery, r);
        CHECK(r.size() == 7);
        CHECK(r.size() > 0 && std::string(r[0].doc_name) == "sample1" && r[0].score == 20);
        // defaults: threshold 0, all results; limit cuts the ordered list
        std::vector<SearchResult> r2;
        s.search(kPyQuery, r2, 0.0, 2);
        CHECK(r2.size() == 2 && r2[1].score == r[1].score &&
              std::strcmp(r2[1].doc_name, r[1].doc_name) == 0);
        std::ostringstream os;
        s.timer().print("search", os);
        CHECK(os.str().rfind("TIMER info=search", 0) == 0);
        CHECK(os.str().find("hashes=") != std::string::npos);
    }

    // geometry getters of the index classes (classic_index/search_file.hpp:22-31)
    {
        auto c = std::make_shared<ClassicIndexMMapSearchFile>(dir + "/all160.cobs_classic");
        CHECK(c->term_size() == 31 && c->canonicalize() == 1 && c->num_hashes() == 3);
        CHECK(c->page_size() == 1 && c->file_names().size() == 33);
        CHECK(c->row_size() == 5 && c->counts_size() == 40);
        auto k = std::make_shared<CompactIndexMMapSearchFile>(dir + "/all160.cobs_compact");
        CHECK(k->page_size() == 2 && k->row_size() == 6 && k->counts_size() == 48);
        CHECK(k->file_names().size() == 33);

        // wrong class for the file: FileIOException like the reference's magic check
        bool threw = false;
        try {
            ClassicIndexMMapSearchFile bad(dir + "/all160.cobs_compact");
        }
        catch (const FileIOException&) {
            threw = true;
        }
        CHECK(threw);

        // search_batch == search, query by query
        ClassicSearch s(c);
        std::vector<std::string> qs = {
            "ACGTACGTACGTACGTACGTACGTACGTACGTACGT",
            std::string(kPyQuery),
            "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT",
        };
        std::vector<std::vector<SearchResult> > batch;
        s.search_batch(qs, batch, 0.0, 0);
        CHECK(batch.size() == qs.size());
        for (size_t i = 0; i < qs.size(); ++i) {
            std::vector<SearchResult> one;
            s.search(qs[i], one);
            CHECK(one.size() == batch[i].size() && one.size() == 33);
            for (size_t j = 0; j < one.size() && j < batch[i].size(); ++j)
                CHECK(one[j].doc_name == batch[i][j].doc_name && one[j].score == batch[i][j].score);
        }
    }

    // multi-index search (tests/classic_index_query.cpp:148-197): 33 + 44 + 55 documents,
    // every one holds only the first k-mer of the query -> all scores are exactly 1
    {
        auto i1 = std::make_shared<ClassicIndexMMapSearchFile>(dir + "/one33.cobs_classic");
        auto i2 = std::make_shared<ClassicIndexMMapSearchFile>(dir + "/one44.cobs_classic");
        auto i3 = std::make_shared<ClassicIndexMMapSearchFile>(dir + "/one55.cobs_classic");
        ClassicSearch s({ i1, i2, i3 });
        // cobs::random_sequence(160, 1) as used by the fixture generator
        std::vector<SearchResult> r;
        // the first k-mer of that sequence is stored in the fixture's query list; use the
        // document names to check the merge order instead: file order, then document order
        s.search(std::string(160, 'A'), r);
        CHECK(r.size() == 33u + 44u + 55u);
        for (size_t i = 1; i < r.size(); ++i) CHECK(r[i - 1].score >= r[i].score);
    }

    // invalid base in a canonicalising index: die() -> DieException when enabled
    // (classic_search.cpp:93-96)
    {
        set_die_with_exception(true);
        ClassicSearch s(dir + "/all160.cobs_classic");
        std::vector<SearchResult> r;
        bool threw = false;
        try {
            s.search("ACGTACGTACGTACGTNCGTACGTACGTACGTACGTACGT", r);
        }
        catch (const DieException& e) {
            threw = std::string(e.what()).find("Invalid DNA base pair") != std::string::npos;
        }
        CHECK(threw);
        threw = false;
        try {
            ClassicSearch bad(dir + "/golden.json");
        }
        catch (const DieException& e) {
            threw = std::string(e.what()).find("Could not open index path") != std::string::npos;
        }
        CHECK(threw);
    }

    // the plug point for custom row sources: an index handed over as in-memory pages
    // (HbmIndexSearchFile::Pages) behaves like the file it was read from; a foreign
    // IndexSearchFile subclass is refused with a message instead of being searched wrongly
    {
        // rebuild all160.cobs_classic from rows read back through read_from_disk()
        auto file = std::make_shared<ClassicIndexMMapSearchFile>(dir + "/all160.cobs_classic");
        const uint64_t row = file->row_size();
        // signature size is not part of the interface: probe rows 0..S-1 via raw hashes
        // (hash % S == hash for hash < S); S comes from the file size here
        std::FILE* f = std::fopen((dir + "/all160.cobs_classic").c_str(), "rb");
        std::fseek(f, 0, SEEK_END);
        const long size = std::ftell(f);
        std::fclose(f);
        // header: magic(5+13) version(4) term(4) canon(1) n_docs(4) sig(8) hashes(8) names... magic(13)
        size_t names = 0;
        for (auto& n : file->file_names()) names += n.size() + 1;
        const uint64_t S = (uint64_t(size) - (5 + 13 + 4 + 4 + 1 + 4 + 8 + 8 + names + 13)) / row;
        std::vector<size_t> hashes(S);
        for (uint64_t i = 0; i < S; ++i) hashes[i] = i;
        std::vector<uint8_t> rows(S * row);
        file->read_from_disk(hashes, rows.data(), 0, row, row);
        HbmIndexSearchFile::Pages p;
        p.term_size = file->term_size();
        p.canonicalize = file->canonicalize();
        p.num_hashes = file->num_hashes();
        p.file_names = file->file_names();
        p.signature_sizes = { S };
        p.page_data = { rows.data() };
        auto mem = std::make_shared<HbmIndexSearchFile>(p);
        CHECK(mem->counts_size() == file->counts_size() && mem->row_size() == row);
        ClassicSearch a(file), b(mem);
        std::vector<SearchResult> ra, rb;
        a.search(kPyQuery, ra, 0.0, 0);
        b.search(kPyQuery, rb, 0.0, 0);
        CHECK(ra.size() == rb.size() && ra.size() == 33);
        for (size_t i = 0; i < ra.size() && i < rb.size(); ++i)
            CHECK(std::string(ra[i].doc_name) == rb[i].doc_name && ra[i].score == rb[i].score);

        struct Foreign : IndexSearchFile {
            std::vector<std::string> names { "x" };
            void read_from_disk(const std::vector<size_t>&, uint8_t*, size_t, size_t, size_t) override { }
            uint32_t term_size() const override { return 31; }
            uint8_t canonicalize() const override { return 1; }
            uint64_t row_size() const override { return 1; }
            uint64_t page_size() const override { return 1; }
            uint64_t num_hashes() const override { return 1; }
            uint64_t counts_size() const override { return 8; }
            const std::vector<std::string>& file_names() const override { return names; }
        };
        ClassicSearch foreign(std::make_shared<Foreign>());
        std::vector<SearchResult> r;
        bool threw = false;
        try {
            foreign.search(kPyQuery, r);
        }
        catch (const DieException& e) {
            threw = std::string(e.what()).find("HBM-resident") != std::string::npos;
        }
        CHECK(threw);
    }

    if (argc > 2 && std::string(argv[2]) == "too_short") {
        // assert_exit: message on stderr and exit(EXIT_FAILURE) (classic_search.cpp:431-433)
        ClassicSearch s(dir + "/all160.cobs_classic");
        std::vector<SearchResult> r;
        s.search("ACGT", r);
        std::fprintf(stderr, "not reached\n");
        return 0;
    }

    std::printf("host_tests: %s (%d failed checks)\n", g_failed ? "FAILED" : "ok", g_failed);
    return g_failed ? 1 : 0;
}
