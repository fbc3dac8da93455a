// ClassicSearch on the GPU: the semantics of cobs::ClassicSearch::search
// (reference: cobs/query/classic_search.cpp:403-505 and counts_to_result 109-202), with the
// per-index work (hash -> row AND -> per-document counts -> threshold -> order) done by
// cobsgpu_search_batch() and only the cross-index / cross-shard merge on the host.
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>

#include <cobsgpu.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <tuple>

namespace cobs {

bool classic_search_disable_8bit = false;
bool classic_search_disable_16bit = false;
bool classic_search_disable_32bit = false;
bool classic_search_disable_sse2 = false;

ClassicSearch::ClassicSearch(std::shared_ptr<IndexSearchFile> index)
    : index_files_({ std::move(index) }) { }

ClassicSearch::ClassicSearch(std::vector<std::shared_ptr<IndexSearchFile> > indices)
    : index_files_(std::move(indices)) { }

ClassicSearch::ClassicSearch(std::string path) {
    if (file_has_header<ClassicIndexHeader>(path))
        index_files_.emplace_back(std::make_shared<ClassicIndexMMapSearchFile>(path));
    else if (file_has_header<CompactIndexHeader>(path))
        index_files_.emplace_back(std::make_shared<CompactIndexMMapSearchFile>(path));
    else
        die_with_message("Could not open index path \"" + path + "\"");
}

void ClassicSearch::search(
    const std::string& query, std::vector<SearchResult>& result,
    double threshold, size_t num_results) {
    std::vector<std::vector<SearchResult> > results;
    search_batch(std::vector<std::string>{ query }, results, threshold, num_results);
    if (results.empty()) return;   // no index files: result is left untouched (reference 410-411)
    result.swap(results[0]);
}

namespace {

struct Entry {
    uint32_t score;
    uint32_t file;
    uint32_t doc;
};

//! runs one packed batch on every shard of one index and appends (score, file, doc) entries
void run_index(
    IndexSearchFile& index, uint32_t file_num, const std::string& blob,
    const std::vector<uint64_t>& offsets, const std::vector<uint32_t>& ids,
    double threshold, uint64_t limit, std::vector<std::vector<Entry> >& out, Timer& timer) {
    if (ids.empty()) return;
    // pack the selected queries
    std::string sub;
    std::vector<uint64_t> off(ids.size() + 1, 0);
    for (size_t i = 0; i < ids.size(); ++i) {
        sub.append(blob, offsets[ids[i]], offsets[ids[i] + 1] - offsets[ids[i]]);
        off[i + 1] = sub.size();
    }
    // one host thread per document shard (= per GPU); each handle has its own stream and buffers
    const auto& shards = index.gpu_shards();
    struct ShardOut {
        int rc = COBSGPU_OK;
        std::string err;
        std::vector<uint64_t> off;
        std::vector<uint32_t> doc, score;
        cobsgpu_timers tm{};
    };
    std::vector<ShardOut> outs(shards.size());
    auto work = [&](size_t s) {
        ShardOut& o = outs[s];
        cobsgpu_set_option(shards[s], "timing", 1);
        cobsgpu_reset_timers(shards[s]);
        cobsgpu_result res;
        o.rc = cobsgpu_search_batch(shards[s], sub.data(), off.data(), uint32_t(ids.size()),
                                    threshold, limit, &res);
        if (o.rc != COBSGPU_OK) {
            o.err = cobsgpu_last_error();   // thread-local in the library: read it here
            return;
        }
        o.off.assign(res.offsets, res.offsets + ids.size() + 1);
        o.doc.assign(res.doc, res.doc + o.off.back());
        o.score.assign(res.score, res.score + o.off.back());
        cobsgpu_get_timers(shards[s], &o.tm);
    };
    if (shards.size() == 1) {
        work(0);
    }
    else {
        std::vector<std::thread> threads;
        for (size_t s = 0; s < shards.size(); ++s) threads.emplace_back(work, s);
        for (auto& t : threads) t.join();
    }
    for (ShardOut& o : outs) {
        if (o.rc == COBSGPU_ERR_QUERY_TOO_SHORT) exit_error(o.err);
        if (o.rc == COBSGPU_ERR_INVALID_BASE)
            die_with_message("Invalid DNA base pair in query string. Only ACGT are allowed.");
        if (o.rc != COBSGPU_OK) die_with_message("GPU search failed: " + o.err);
        for (size_t i = 0; i < ids.size(); ++i) {
            std::vector<Entry>& dst = out[ids[i]];
            for (uint64_t e = o.off[i]; e < o.off[i + 1]; ++e)
                dst.push_back(Entry { o.score[e], file_num, o.doc[e] });
        }
        timer.add("hashes", o.tm.hashes_ms * 1e-3);
        timer.add("io", (o.tm.h2d_ms + o.tm.d2h_ms) * 1e-3);
        timer.add("and rows", o.tm.score_ms * 1e-3);   // gather + AND + add are one fused kernel
        timer.add("add rows", 0.0);
        timer.add("sort results", o.tm.select_ms * 1e-3);
    }
}

} // namespace

void ClassicSearch::search_batch(
    const std::vector<std::string>& queries,
    std::vector<std::vector<SearchResult> >& results,
    double threshold, size_t num_results) {
    results.clear();
    if (index_files_.empty()) return;
    const size_t nq = queries.size();
    results.resize(nq);

    // geometry over all indices (reference 413-451)
    size_t total_documents = 0;
    uint32_t max_term_size = 0;
    for (auto& f : index_files_) {
        total_documents += f->counts_size();
        max_term_size = std::max(max_term_size, f->term_size());
    }
    std::string blob;
    std::vector<uint64_t> offsets(nq + 1, 0);
    for (size_t i = 0; i < nq; ++i) {
        assert_exit(queries[i].size() >= max_term_size,
                    "query too short, needs to be at least "
                    + std::to_string(max_term_size) + " characters long");
        blob += queries[i];
        offsets[i + 1] = blob.size();
    }
    const size_t limit = num_results == 0 ? total_documents : std::min(num_results, total_documents);

    // The reference skips the sort when the query produced at most one hash in total
    // (classic_search.cpp:130/180, max_counts = total_hashes) and then returns the first
    // `limit` kept documents in column order; such queries need the untruncated lists.
    std::vector<uint32_t> normal, single_hash;
    for (size_t i = 0; i < nq; ++i) {
        size_t total_hashes = 0;
        for (auto& f : index_files_)
            total_hashes += f->num_hashes() * (queries[i].size() - f->term_size() + 1);
        (total_hashes > 1 ? normal : single_hash).push_back(uint32_t(i));
    }

    const bool sharded = std::any_of(index_files_.begin(), index_files_.end(),
                                     [](auto& f) { return f->gpu_shards().size() > 1; });
    std::vector<std::vector<Entry> > entries(nq);
    for (size_t f = 0; f < index_files_.size(); ++f) {
        // per-index lists are already ordered and cut at `limit`: the global top-k is
        // contained in the union of the per-index (and per-shard) top-ks
        run_index(*index_files_[f], uint32_t(f), blob, offsets, normal, threshold, limit,
                  entries, timer_);
        run_index(*index_files_[f], uint32_t(f), blob, offsets, single_hash, threshold, 0,
                  entries, timer_);
    }

    const bool merge = index_files_.size() > 1 || sharded;
    for (uint32_t i : normal) {
        std::vector<Entry>& e = entries[i];
        if (merge) {
            // score descending, then (file, doc) ascending (reference 139-143 / 173-177)
            std::sort(e.begin(), e.end(), [](const Entry& a, const Entry& b) {
                          return std::tie(b.score, a.file, a.doc) < std::tie(a.score, b.file, b.doc);
                      });
        }
    }
    for (uint32_t i : single_hash) {
        std::vector<Entry>& e = entries[i];
        std::sort(e.begin(), e.end(), [](const Entry& a, const Entry& b) {
                      return std::tie(a.file, a.doc) < std::tie(b.file, b.doc);
                  });
    }
    for (size_t i = 0; i < nq; ++i) {
        std::vector<Entry>& e = entries[i];
        const size_t n = std::min(limit, e.size());
        results[i].resize(n);
        for (size_t j = 0; j < n; ++j)
            results[i][j] = SearchResult(
                index_files_[e[j].file]->file_names()[e[j].doc].c_str(), e[j].score);
    }
}

} // namespace cobs
