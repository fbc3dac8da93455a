/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" wrapper around the UNMODIFIED reference (compiled from the
 * sources where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libcobs_ref.so).  It drives the reference through its own public
 * API (cobs::ClassicSearch, cobs::classic_construct, ...) so that tests can
 * pin the C restatement (cobs_oracle.c) against the real thing, golden vectors
 * can be generated (tests/golden/make_golden.py), and bench.py can time the
 * reference's CPU path (cpu_baseline.kind == "reference").
 * Nothing under cobs_b200/ links or loads this.
 */
#include <cobs/construction/classic_index.hpp>
#include <cobs/construction/compact_index.hpp>
#include <cobs/document_list.hpp>
#include <cobs/kmer_buffer.hpp>
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/file.hpp>
#include <cobs/util/misc.hpp>
#include <cobs/util/query.hpp>

#include <tlx/die.hpp>
#include <tlx/logger/core.hpp>
#include <xxhash.h>

#include <chrono>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct RefHandle {
    std::vector<std::shared_ptr<cobs::IndexSearchFile> > files;
    std::unique_ptr<cobs::ClassicSearch> search;
    //! doc_name pointer -> (file, doc)
    std::unordered_map<const char*, std::pair<uint32_t, uint32_t> > name_map;
};

thread_local std::string g_error;

template <typename F>
int guarded(F&& f) {
    // like the reference's main() (src/cobs.cpp:1058): log lines go to stderr, stdout stays
    // clean for the one JSON line bench.py prints
    static bool once = (tlx::set_logger_to_stderr(), true);
    (void)once;
    tlx::set_die_with_exception(true);
    try {
        f();
        return 0;
    }
    catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
    catch (...) {
        g_error = "unknown exception";
        return -1;
    }
}

} // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

uint64_t ref_xxh64(const void* data, size_t len, uint64_t seed) {
    return XXH64(data, len, seed);
}

int ref_canonicalize_kmer(const char* in, char* out, size_t size) {
    return cobs::canonicalize_kmer(in, out, size) ? 1 : 0;
}

void ref_random_sequence(size_t size, size_t seed, char* out) {
    std::string s = cobs::random_sequence(size, seed);
    std::memcpy(out, s.data(), size);
}

//! same generator as `cobs benchmark-fpr` (src/cobs.cpp:709-720): one mt19937
//! stream, queries drawn back to back
void ref_random_queries_mt19937(size_t seed, size_t n, size_t len, char* out) {
    std::mt19937 rng(seed);
    for (size_t i = 0; i < n; ++i) {
        std::string s = cobs::random_sequence_rng(len, rng);
        std::memcpy(out + i * len, s.data(), len);
    }
}

void ref_set_threads(size_t n) { cobs::gopt_threads = n; }
void ref_set_load_complete(int b) { cobs::gopt_load_complete_index = b != 0; }
void ref_set_disable(int d8, int d16, int d32, int dsse2) {
    cobs::classic_search_disable_8bit = d8 != 0;
    cobs::classic_search_disable_16bit = d16 != 0;
    cobs::classic_search_disable_32bit = d32 != 0;
    cobs::classic_search_disable_sse2 = dsse2 != 0;
}

void* ref_open(const char* const* paths, size_t n) {
    RefHandle* h = new RefHandle;
    int rc = guarded([&] {
        for (size_t i = 0; i < n; ++i) {
            std::string p = paths[i];
            // same sniffing as src/cobs.cpp:509-521
            if (cobs::file_has_header<cobs::ClassicIndexHeader>(p))
                h->files.push_back(
                    std::make_shared<cobs::ClassicIndexMMapSearchFile>(p));
            else if (cobs::file_has_header<cobs::CompactIndexHeader>(p))
                h->files.push_back(
                    std::make_shared<cobs::CompactIndexMMapSearchFile>(p));
            else
                die("Could not open index path \"" << p << "\"");
        }
        h->search = std::make_unique<cobs::ClassicSearch>(h->files);
        for (size_t f = 0; f < h->files.size(); ++f) {
            const auto& names = h->files[f]->file_names();
            for (size_t d = 0; d < names.size(); ++d)
                h->name_map[names[d].c_str()] = { uint32_t(f), uint32_t(d) };
        }
    });
    if (rc != 0) {
        delete h;
        return nullptr;
    }
    return h;
}

void ref_close(void* handle) { delete static_cast<RefHandle*>(handle); }

uint32_t ref_num_docs(void* handle, size_t file) {
    return static_cast<RefHandle*>(handle)->files[file]->file_names().size();
}
const char* ref_doc_name(void* handle, size_t file, size_t doc) {
    return static_cast<RefHandle*>(handle)->files[file]->file_names()[doc].c_str();
}
uint64_t ref_counts_size(void* handle, size_t file) {
    return static_cast<RefHandle*>(handle)->files[file]->counts_size();
}

//! ClassicSearch::search through the public API; maps doc_name back to ids.
//! NOTE: "query too short" / "query too long" exit(1) inside the reference
//! (assert_exit) -- callers test those in a subprocess.
int ref_search(void* handle, const char* query, size_t len, double threshold,
               size_t num_results, uint32_t* out_file, uint32_t* out_doc,
               uint32_t* out_score, size_t cap, size_t* out_count) {
    RefHandle* h = static_cast<RefHandle*>(handle);
    return guarded([&] {
        std::vector<cobs::SearchResult> result;
        h->search->search(std::string(query, len), result, threshold,
                          num_results);
        *out_count = result.size();
        for (size_t i = 0; i < result.size() && i < cap; ++i) {
            auto it = h->name_map.find(result[i].doc_name);
            if (it == h->name_map.end()) die("doc_name not from this index");
            if (out_file) out_file[i] = it->second.first;
            out_doc[i] = it->second.second;
            out_score[i] = result[i].score;
        }
    });
}

//! wall-clock seconds for a loop of search() over nq queries (bench baseline,
//! BASELINE.md section 4).  *out_results accumulates result sizes so the loop
//! cannot be optimised away.
int ref_bench(void* handle, const char* blob, const uint64_t* offsets,
              size_t nq, double threshold, size_t num_results,
              double* out_seconds, uint64_t* out_results) {
    RefHandle* h = static_cast<RefHandle*>(handle);
    return guarded([&] {
        std::vector<cobs::SearchResult> result;
        uint64_t total = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < nq; ++i) {
            std::string q(blob + offsets[i], offsets[i + 1] - offsets[i]);
            h->search->search(q, result, threshold, num_results);
            total += result.size();
        }
        auto t1 = std::chrono::steady_clock::now();
        *out_seconds = std::chrono::duration<double>(t1 - t0).count();
        *out_results = total;
    });
}

int ref_classic_construct(const char* in_dir, const char* out_file,
                          const char* tmp_dir, unsigned term_size,
                          unsigned num_hashes, double fpr, int canonicalize,
                          uint64_t signature_size) {
    return guarded([&] {
        cobs::ClassicIndexParameters p;
        p.term_size = term_size;
        p.num_hashes = num_hashes;
        p.false_positive_rate = fpr;
        p.canonicalize = canonicalize;
        p.signature_size = signature_size;
        p.clobber = true;
        cobs::classic_construct(cobs::DocumentList(std::string(in_dir)),
                                out_file, tmp_dir, p);
    });
}

int ref_compact_construct(const char* in_dir, const char* out_file,
                          const char* tmp_dir, unsigned term_size,
                          unsigned num_hashes, double fpr, int canonicalize,
                          uint64_t page_size) {
    return guarded([&] {
        cobs::CompactIndexParameters p;
        p.term_size = term_size;
        p.num_hashes = num_hashes;
        p.false_positive_rate = fpr;
        p.canonicalize = canonicalize;
        p.page_size = page_size;
        p.clobber = true;
        cobs::compact_construct(cobs::DocumentList(std::string(in_dir)),
                                out_file, tmp_dir, p);
    });
}

int ref_classic_construct_random(const char* out_file, uint64_t signature_size,
                                 uint64_t num_documents, size_t document_size,
                                 uint64_t num_hashes, size_t seed) {
    return guarded([&] {
        cobs::classic_construct_random(out_file, signature_size, num_documents,
                                       document_size, num_hashes, seed);
    });
}

//! Writes one .cobs_doc k-mer buffer holding the canonical 31-mers found at the
//! given positions of `seq` -- the building block the reference's query tests
//! use for their synthetic documents (tests/test_util.hpp:42-95).
int ref_write_kmer_doc(const char* path, const char* name, const char* seq,
                       size_t seq_len, const uint64_t* positions, size_t n_pos) {
    return guarded([&] {
        cobs::KMerBuffer<31> doc;
        cobs::KMer<31> k;
        char buf[32];
        for (size_t i = 0; i < n_pos; ++i) {
            die_unless(positions[i] + 31 <= seq_len);
            bool good = cobs::canonicalize_kmer(seq + positions[i], buf, 31);
            die_unless(good);
            buf[31] = 0;
            k.init(buf);
            doc.data().push_back(k);
        }
        doc.serialize(cobs::fs::path(path), name);
    });
}

} // extern "C"
