// drop-in for cobs/query/classic_index/{search_file,mmap_search_file}.hpp of the reference
#pragma once
#include <cobs/query/index_file.hpp>

namespace cobs {

class ClassicIndexSearchFile : public IndexSearchFile
{
protected:
    explicit ClassicIndexSearchFile(const fs::path& path) : IndexSearchFile(path, 0) { }
};

//! a classic index whose matrix is resident in HBM (the reference mmaps it)
class ClassicIndexMMapSearchFile : public ClassicIndexSearchFile
{
public:
    explicit ClassicIndexMMapSearchFile(const fs::path& path) : ClassicIndexSearchFile(path) { }
};

} // namespace cobs
