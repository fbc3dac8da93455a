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

//! the HBM-resident form of an index file; anything else cannot be searched by this build
HbmIndexSearchFile& hbm(IndexSearchFile& index) {
    auto* h = dynamic_cast<HbmIndexSearchFile*>(&index);
    if (!h)
        die_with_message("this build searches HBM-resident index files only: wrap custom row "
                         "sources in cobs::HbmIndexSearchFile(Pages)");
    return *h;
}

//! folds the device-side phase times of one call into the reference's Timer keys
void account_timers(const std::vector<cobsgpu_index*>& shards, Timer& timer) {
    // shards run concurrently: a phase lasts as long as its slowest shard
    cobsgpu_timers tm{};
    for (cobsgpu_index* s : shards) {
        cobsgpu_timers t{};
        cobsgpu_get_timers(s, &t);
        tm.hashes_ms = std::max(tm.hashes_ms, t.hashes_ms);
        tm.score_ms = std::max(tm.score_ms, t.score_ms);
        tm.select_ms = std::max(tm.select_ms, t.select_ms);
        tm.h2d_ms = std::max(tm.h2d_ms, t.h2d_ms);
        tm.d2h_ms = std::max(tm.d2h_ms, t.d2h_ms);
    }
    timer.add("hashes", tm.hashes_ms * 1e-3);
    timer.add("io", (tm.h2d_ms + tm.d2h_ms) * 1e-3);
    timer.add("and rows", tm.score_ms * 1e-3);   // gather + AND + add are one fused kernel
    timer.add("add rows", 0.0);
    timer.add("sort results", tm.select_ms * 1e-3);
}

//! runs one packed batch on one index -- a single GPU handle, or a group of document-axis
//! shards merged on the leader GPU -- and appends (score, file, doc) entries
void run_index(
    IndexSearchFile& index_file, uint32_t file_num, const std::string& blob,
    const std::vector<uint64_t>& offsets, const std::vector<uint32_t>& ids,
    double threshold, uint64_t limit, std::vector<std::vector<Entry> >& out, Timer& timer) {
    if (ids.empty()) return;
    HbmIndexSearchFile& index = hbm(index_file);
    // pack the selected queries
    std::string sub;
    std::vector<uint64_t> off(ids.size() + 1, 0);
    for (size_t i = 0; i < ids.size(); ++i) {
        sub.append(blob, offsets[ids[i]], offsets[ids[i] + 1] - offsets[ids[i]]);
        off[i + 1] = sub.size();
    }
    const auto& shards = index.gpu_shards();
    for (cobsgpu_index* s : shards) {
        cobsgpu_set_option(s, "timing", 1);
        cobsgpu_reset_timers(s);
    }
    cobsgpu_result res;
    const int rc = index.gpu_group()
                   ? cobsgpu_group_search_batch(index.gpu_group(), sub.data(), off.data(),
                                                uint32_t(ids.size()), threshold, limit, &res)
                   : cobsgpu_search_batch(shards[0], sub.data(), off.data(), uint32_t(ids.size()),
                                          threshold, limit, &res);
    if (rc == COBSGPU_ERR_QUERY_TOO_SHORT) exit_error(cobsgpu_last_error());
    if (rc == COBSGPU_ERR_INVALID_BASE)
        die_with_message("Invalid DNA base pair in query string. Only ACGT are allowed.");
    if (rc != COBSGPU_OK) die_with_message(std::string("GPU search failed: ") + cobsgpu_last_error());
    for (size_t i = 0; i < ids.size(); ++i) {
        std::vector<Entry>& dst = out[ids[i]];
        for (uint64_t e = res.offsets[i]; e < res.offsets[i + 1]; ++e)
            dst.push_back(Entry { res.score[e], file_num, res.doc[e] });
    }
    account_timers(shards, timer);
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
    {
        size_t bytes = 0;
        for (const std::string& q : queries) bytes += q.size();
        blob.reserve(bytes);
    }
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

    // One index and no single-hash query -- the usual case: the library's lists are the result;
    // no intermediate entries, no re-packing of the queries.
    if (index_files_.size() == 1 && single_hash.empty()) {
        HbmIndexSearchFile& index = hbm(*index_files_[0]);
        const auto& shards = index.gpu_shards();
        for (cobsgpu_index* s : shards) {
            cobsgpu_set_option(s, "timing", 1);
            cobsgpu_reset_timers(s);
        }
        cobsgpu_result res;
        const int rc = index.gpu_group()
                       ? cobsgpu_group_search_batch(index.gpu_group(), blob.data(), offsets.data(),
                                                    uint32_t(nq), threshold, limit, &res)
                       : cobsgpu_search_batch(shards[0], blob.data(), offsets.data(), uint32_t(nq),
                                              threshold, limit, &res);
        if (rc == COBSGPU_ERR_QUERY_TOO_SHORT) exit_error(cobsgpu_last_error());
        if (rc == COBSGPU_ERR_INVALID_BASE)
            die_with_message("Invalid DNA base pair in query string. Only ACGT are allowed.");
        if (rc != COBSGPU_OK) die_with_message(std::string("GPU search failed: ") + cobsgpu_last_error());
        const std::vector<std::string>& names = index.file_names();
        for (size_t i = 0; i < nq; ++i) {
            const uint64_t a = res.offsets[i], b = res.offsets[i + 1];
            std::vector<SearchResult>& r = results[i];
            r.resize(b - a);
            for (uint64_t e = a; e < b; ++e)
                r[e - a] = SearchResult(names[res.doc[e]].c_str(), res.score[e]);
        }
        account_timers(shards, timer_);
        return;
    }

    std::vector<std::vector<Entry> > entries(nq);
    for (size_t f = 0; f < index_files_.size(); ++f) {
        // per-index lists are already ordered and cut at `limit`: the global top-k is
        // contained in the union of the per-index (and per-shard) top-ks
        run_index(*index_files_[f], uint32_t(f), blob, offsets, normal, threshold, limit,
                  entries, timer_);
        run_index(*index_files_[f], uint32_t(f), blob, offsets, single_hash, threshold, 0,
                  entries, timer_);
    }

    // (a sharded index arrives merged: the group's leader GPU did that)
    const bool merge = index_files_.size() > 1;
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
