// cobs/query/search.hpp -- abstract search interface, drop-in for the reference's
// cobs/query/search.hpp:17-47 (same members and defaults) plus a batched entry point.
#pragma once
#include <cobs/util/timer.hpp>

#include <cstdint>
#include <string>
#include <vector>

namespace cobs {

struct SearchResult {
    //! document name; borrows from the index object, valid while the index lives
    const char* doc_name;
    //! number of matched k-mers
    uint32_t score;

    SearchResult() = default;
    SearchResult(const char* doc_name, uint32_t score) : doc_name(doc_name), score(score) { }
};

class Search
{
public:
    virtual ~Search() = default;

    Timer& timer() { return timer_; }
    const Timer& timer() const { return timer_; }

    virtual void search(
        const std::string& query,
        std::vector<SearchResult>& result,
        double threshold = 0.0, size_t num_results = 0) = 0;

    //! extension: many queries per call -- one GPU batch instead of one launch per query.
    //! results[i] is exactly what search(queries[i], ...) returns.
    virtual void search_batch(
        const std::vector<std::string>& queries,
        std::vector<std::vector<SearchResult> >& results,
        double threshold = 0.0, size_t num_results = 0) = 0;

public:
    //! phases: "hashes", "io", "and rows", "add rows", "sort results"
    Timer timer_;
};

} // namespace cobs
