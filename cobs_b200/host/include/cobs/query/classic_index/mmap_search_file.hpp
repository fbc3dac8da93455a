// drop-in for cobs/query/classic_index/mmap_search_file.hpp of the reference (17-33)
#pragma once
#include <cobs/query/classic_index/search_file.hpp>

namespace cobs {

//! a classic index whose matrix is resident in HBM (the reference mmaps it)
class ClassicIndexMMapSearchFile : public ClassicIndexSearchFile
{
public:
    explicit ClassicIndexMMapSearchFile(const fs::path& path) : ClassicIndexSearchFile(path) { }
};

} // namespace cobs
