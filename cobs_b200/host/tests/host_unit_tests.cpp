// host_unit_tests.cpp -- CPU-only checks of the host-side mirror (no GPU needed): Timer,
// header sniffing, the reference's error conventions, and that index classes refuse wrong
// files before any device work.
//   host_unit_tests <golden dir> [exit_error]
#include <cobs/file/file_io_exception.hpp>
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>
#include <cobs/util/timer.hpp>

#include <cstdio>
#include <sstream>
#include <string>

static int g_failed = 0;
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++g_failed;                                                                  \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    using namespace cobs;

    if (argc > 2 && std::string(argv[2]) == "exit_error") {
        assert_exit(false, "query too short, needs to be at least 31 characters long");
        return 0;   // not reached: assert_exit prints and exit(EXIT_FAILURE)s
    }

    // Timer: the reference's "TIMER info=<tag> <phase>=<sec> ... total=<sec>" line
    {
        Timer t;
        t.add("hashes", 0.25);
        t.add("and rows", 1.5);
        t.add("hashes", 0.25);
        CHECK(t.get("hashes") == 0.5 && t.get("and rows") == 1.5 && t.get("io") == 0.0);
        std::ostringstream os;
        t.print("search", os);
        CHECK(os.str() == "TIMER info=search hashes=0.5 and rows=1.5 io=0 total=2\n");
        Timer u;
        u.add("io", 1.0);
        t += u;
        CHECK(t.get("io") == 1.0);
        t.active("sort results");
        t.stop();
        CHECK(t.get("sort results") >= 0.0);
        t.reset();
        CHECK(t.get("hashes") == 0.0);
    }

    // header sniffing (what src/cobs.cpp:509-521 uses to pick the index class)
    CHECK(file_has_header<ClassicIndexHeader>(dir + "/python_test.cobs_classic"));
    CHECK(!file_has_header<CompactIndexHeader>(dir + "/python_test.cobs_classic"));
    CHECK(file_has_header<CompactIndexHeader>(dir + "/all160.cobs_compact"));
    CHECK(!file_has_header<ClassicIndexHeader>(dir + "/all160.cobs_compact"));
    CHECK(!file_has_header<ClassicIndexHeader>(dir + "/golden.json"));
    CHECK(!file_has_header<ClassicIndexHeader>(dir + "/does_not_exist"));
    CHECK(!file_has_header<ClassicIndexHeader>(dir));   // a directory
    CHECK(ClassicIndexHeader::file_extension == ".cobs_classic");
    CHECK(CompactIndexHeader::magic_word == "COMPACT_INDEX");

    // wrong class for a file: FileIOException("invalid file type"), raised before any GPU work
    {
        bool threw = false;
        try {
            ClassicIndexMMapSearchFile bad(dir + "/all160.cobs_compact");
        }
        catch (const FileIOException& e) {
            threw = std::string(e.what()) == "invalid file type";
        }
        CHECK(threw);
        threw = false;
        try {
            CompactIndexMMapSearchFile bad(dir + "/all160.cobs_classic");
        }
        catch (const FileIOException&) {
            threw = true;
        }
        CHECK(threw);
    }

    // die(): terminate by default, DieException when enabled (tlx::set_die_with_exception)
    {
        CHECK(set_die_with_exception(true) == false);
        bool threw = false;
        try {
            ClassicSearch s(dir + "/golden.json");
        }
        catch (const DieException& e) {
            threw = std::string(e.what()) == "Could not open index path \"" + dir + "/golden.json\"";
        }
        CHECK(threw);
        CHECK(set_die_with_exception(false) == true);
    }

    // an empty ClassicSearch returns without touching the result (classic_search.cpp:410-411)
    {
        ClassicSearch s(std::vector<std::shared_ptr<IndexSearchFile> >{});
        std::vector<SearchResult> r(3);
        s.search("ACGT", r);
        CHECK(r.size() == 3);
    }

    // settings exist under the reference's names
    CHECK(gopt_threads >= 1);
    CHECK(gopt_load_complete_index == false);
    CHECK(gopt_gpus >= 1);

    std::printf("host_unit_tests: %s (%d failed checks)\n", g_failed ? "FAILED" : "ok", g_failed);
    return g_failed ? 1 : 0;
}
