// cobs/query/search.hpp -- abstract search interface.  Source-compatible with the reference's
// cobs/query/search.hpp:17-47 (same type names, members, argument order and defaults), plus a
// batched entry point that maps onto one GPU batch.
#pragma once
#include <cobs/util/timer.hpp>

#include <cstdint>
#include <string>
#include <vector>

namespace cobs {

//! one hit: the document's name and the number of query k-mers found in it
struct SearchResult {
    const char* doc_name;   //!< borrows from the index object, valid while the index lives
    uint32_t score;

    SearchResult() = default;
    SearchResult(const char* name, uint32_t s) : doc_name(name), score(s) { }
};

class Search
{
public:
    virtual ~Search() = default;

    //! documents with at least ceil(threshold * #k-mers) hits, best first, at most num_results
    //! of them (0 = all)
    virtual void search(const std::string& query, std::vector<SearchResult>& result,
                        double threshold = 0.0, size_t num_results = 0) = 0;

    //! extension: many queries per call; results[i] is exactly what search(queries[i], ...) gives
    virtual void search_batch(const std::vector<std::string>& queries,
                              std::vector<std::vector<SearchResult> >& results,
                              double threshold = 0.0, size_t num_results = 0) = 0;

    Timer& timer() { return timer_; }
    const Timer& timer() const { return timer_; }

public:
    //! phases "hashes", "io", "and rows", "add rows", "sort results", filled from CUDA events
    Timer timer_;
};

} // namespace cobs
