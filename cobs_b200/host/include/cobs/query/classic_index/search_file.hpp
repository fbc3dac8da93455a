// drop-in for cobs/query/classic_index/search_file.hpp of the reference (17-36)
#pragma once
#include <cobs/query/index_file.hpp>

namespace cobs {

class ClassicIndexSearchFile : public HbmIndexSearchFile
{
protected:
    explicit ClassicIndexSearchFile(const fs::path& path) : HbmIndexSearchFile(path, 0) { }

public:
    virtual ~ClassicIndexSearchFile() = default;
};

} // namespace cobs
