// cobs/query/classic_search.hpp -- drop-in for the reference's ClassicSearch
// (cobs/query/classic_search.hpp:19-50): same constructors, same search() semantics, served
// by the CUDA path through the C ABI (include/cobsgpu.h).
#pragma once
#include <cobs/query/index_file.hpp>
#include <cobs/query/search.hpp>

#include <memory>
#include <string>
#include <vector>

namespace cobs {

class ClassicSearch : public Search
{
public:
    //! auto-detects classic vs compact (reference: classic_search.cpp:51-64)
    ClassicSearch(std::string path);
    ClassicSearch(std::shared_ptr<IndexSearchFile> index);
    ClassicSearch(std::vector<std::shared_ptr<IndexSearchFile> > indices);

    void search(const std::string& query, std::vector<SearchResult>& result,
                double threshold = 0.0, size_t num_results = 0) final;

    void search_batch(const std::vector<std::string>& queries,
                      std::vector<std::vector<SearchResult> >& results,
                      double threshold = 0.0, size_t num_results = 0) final;

protected:
    std::vector<std::shared_ptr<IndexSearchFile> > index_files_;
};

//! BASELINE.json names a CompactSearch; in the reference compact indices are served by
//! ClassicSearch over a CompactIndexMMapSearchFile (classic_search.cpp:57-60)
using CompactSearch = ClassicSearch;

// The reference's test-only toggles for its score-width variants (classic_search.cpp:207-211).
// All variants give identical results; the GPU path has one implementation, the toggles are
// accepted and ignored.
extern bool classic_search_disable_8bit;
extern bool classic_search_disable_16bit;
extern bool classic_search_disable_32bit;
extern bool classic_search_disable_sse2;

} // namespace cobs
