// cobs/util/file.hpp -- header sniffing used by callers to pick the index class
// (reference: cobs/util/file.hpp:53-65 file_has_header<Header>)
#pragma once
#include <cobs/util/fs.hpp>
#include <string>

namespace cobs {

//! tag types carrying the magic words (reference: cobs/file/classic_index_header.cpp:15-17,
//! cobs/file/compact_index_header.cpp:13-15)
struct ClassicIndexHeader {
    static const std::string magic_word;
    static const uint32_t version;
    static const std::string file_extension;
};
struct CompactIndexHeader {
    static const std::string magic_word;
    static const uint32_t version;
    static const std::string file_extension;
};

bool file_has_magic(const fs::path& p, const std::string& magic_word, uint32_t version);

//! true if `p` is a regular file starting with "COBS:" + Header::magic_word + version
template <class Header>
bool file_has_header(const fs::path& p) {
    return file_has_magic(p, Header::magic_word, Header::version);
}

} // namespace cobs
