// drop-in for cobs/query/compact_index/search_file.hpp of the reference (17-39)
#pragma once
#include <cobs/query/index_file.hpp>

namespace cobs {

class CompactIndexSearchFile : public HbmIndexSearchFile
{
protected:
    explicit CompactIndexSearchFile(const fs::path& path) : HbmIndexSearchFile(path, 1) { }

public:
    virtual ~CompactIndexSearchFile() = default;
};

} // namespace cobs
