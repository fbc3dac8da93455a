// drop-in for cobs/query/compact_index/{search_file,mmap_search_file}.hpp of the reference
#pragma once
#include <cobs/query/index_file.hpp>

namespace cobs {

class CompactIndexSearchFile : public IndexSearchFile
{
protected:
    explicit CompactIndexSearchFile(const fs::path& path) : IndexSearchFile(path, 1) { }
};

//! a compact index whose pages are resident in HBM (the reference mmaps them)
class CompactIndexMMapSearchFile : public CompactIndexSearchFile
{
public:
    explicit CompactIndexMMapSearchFile(const fs::path& path) : CompactIndexSearchFile(path) { }
};

} // namespace cobs
