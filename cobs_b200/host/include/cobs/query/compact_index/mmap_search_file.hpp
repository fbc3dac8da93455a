// drop-in for cobs/query/compact_index/mmap_search_file.hpp of the reference (17-36)
#pragma once
#include <cobs/query/compact_index/search_file.hpp>

namespace cobs {

//! a compact index whose pages are resident in HBM (the reference mmaps them)
class CompactIndexMMapSearchFile : public CompactIndexSearchFile
{
public:
    explicit CompactIndexMMapSearchFile(const fs::path& path) : CompactIndexSearchFile(path) { }
};

} // namespace cobs
