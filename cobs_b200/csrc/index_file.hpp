// cobs_b200/csrc/index_file.hpp -- read side of the reference's on-disk index formats.
//
// Host-only.  Parses `.cobs_classic` (cobs/file/classic_index_header.cpp:26-50) and
// `.cobs_compact` (cobs/file/compact_index_header.cpp:20-65) files and maps the bit matrix
// read-only (the analogue of initialize_mmap, cobs/util/query.cpp:38-88) so it can be
// streamed into HBM.  All fields little-endian, packed, no alignment.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace cobsgpu {

struct IndexFile {
    int kind = -1;   // 0 classic, 1 compact
    uint32_t term_size = 0;
    uint8_t canonicalize = 0;
    uint32_t num_hashes = 0;
    uint32_t n_docs = 0;
    uint64_t page_size = 0;   // bytes per row per page (classic: ceil(n_docs/8))
    std::vector<uint64_t> signature_sizes;
    std::vector<const uint8_t*> page_data;
    std::vector<std::string> doc_names;
    uint64_t data_pos = 0;    // file offset of the first matrix byte (stream_pos_.curr_pos)

    // mapping
    int fd = -1;
    uint8_t* map = nullptr;
    size_t map_size = 0;

    ~IndexFile() { close(); }
    void close() {
        if (map) munmap(map, map_size);
        if (fd >= 0) ::close(fd);
        map = nullptr;
        fd = -1;
    }

    // returns empty string on success, else an error message
    std::string open(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return "could not open index file " + path + ": " + std::strerror(errno);
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return "not a regular file: " + path;
        map_size = static_cast<size_t>(st.st_size);
        if (map_size == 0) return "invalid file type";
        void* m = mmap(nullptr, map_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return std::string("mmap failed: ") + std::strerror(errno);
        map = static_cast<uint8_t*>(m);
        // sequential streaming into HBM (the reference asks for MADV_RANDOM because it
        // gathers rows on the CPU; we read the matrix exactly once)
        madvise(map, map_size, MADV_SEQUENTIAL);
        return parse();
    }

private:
    size_t pos_ = 0;
    bool bad_ = false;

    template <typename T>
    T get() {
        T v{};
        if (bad_ || pos_ + sizeof(T) > map_size) {
            bad_ = true;
            return v;
        }
        std::memcpy(&v, map + pos_, sizeof(T));
        pos_ += sizeof(T);
        return v;
    }
    bool magic(const char* m) {
        size_t n = std::strlen(m);
        if (bad_ || pos_ + n > map_size) {
            bad_ = true;
            return false;
        }
        bool ok = std::memcmp(map + pos_, m, n) == 0;
        pos_ += n;
        return ok;
    }
    void names(uint32_t n) {
        // every name ends with '\n': a count beyond the remaining bytes is a corrupt header
        if (bad_ || n > map_size - pos_) {
            bad_ = true;
            return;
        }
        doc_names.resize(n);
        for (uint32_t i = 0; i < n && !bad_; ++i) {
            const void* nl = std::memchr(map + pos_, '\n', map_size - pos_);
            if (!nl) {
                bad_ = true;
                break;
            }
            size_t e = static_cast<const uint8_t*>(nl) - map;
            doc_names[i].assign(reinterpret_cast<const char*>(map + pos_), e - pos_);
            pos_ = e + 1;
        }
    }

    std::string parse() {
        if (!magic("COBS:")) return "invalid file type";
        size_t after = pos_;
        if (magic("CLASSIC_INDEX")) {
            uint32_t version = get<uint32_t>();
            if (bad_ || version != 1) return "invalid file version";
            term_size = get<uint32_t>();
            canonicalize = get<uint8_t>();
            n_docs = get<uint32_t>();
            uint64_t sig = get<uint64_t>();
            uint64_t nh = get<uint64_t>();
            if (bad_) return "input filestream broken";
            names(n_docs);
            if (bad_ || !magic("CLASSIC_INDEX")) return "invalid file type";
            kind = 0;
            num_hashes = static_cast<uint32_t>(nh);
            page_size = (static_cast<uint64_t>(n_docs) + 7) / 8;
            signature_sizes = { sig };
            data_pos = pos_;
            if (page_size != 0 && sig > (map_size - pos_) / page_size) return "index file truncated";
            page_data = { map + pos_ };
            return "";
        }
        pos_ = after;
        bad_ = false;
        if (!magic("COMPACT_INDEX")) return "invalid file type";
        uint32_t version = get<uint32_t>();
        if (bad_ || version != 1) return "invalid file version";
        term_size = get<uint32_t>();
        canonicalize = get<uint8_t>();
        uint32_t n_params = get<uint32_t>();
        n_docs = get<uint32_t>();
        page_size = get<uint64_t>();
        if (bad_ || n_params == 0 || page_size == 0) return "input filestream broken";
        if (n_params > (map_size - pos_) / 16) return "input filestream broken";
        signature_sizes.resize(n_params);
        for (uint32_t i = 0; i < n_params; ++i) {
            signature_sizes[i] = get<uint64_t>();
            uint64_t nh = get<uint64_t>();
            if (i == 0) num_hashes = static_cast<uint32_t>(nh);
            // one num_hashes for all pages (compact_index/search_file.cpp:23-27)
            else if (nh != num_hashes) return "compact index with differing num_hashes";
        }
        if (bad_) return "input filestream broken";
        names(n_docs);
        if (bad_) return "input filestream broken";
        // zero padding so that the matrix starts page_size-aligned
        // (compact_index_header.cpp:20-22, 62-63)
        {
            const uint64_t pad = (page_size - ((pos_ + 13) % page_size)) % page_size;
            if (pad > map_size - pos_) return "input filestream broken";
            pos_ += pad;
        }
        if (!magic("COMPACT_INDEX")) return "invalid file type";
        kind = 1;
        data_pos = pos_;
        page_data.resize(n_params);
        size_t p = pos_;
        for (uint32_t i = 0; i < n_params; ++i) {
            if (signature_sizes[i] > (map_size - p) / page_size) return "index file truncated";
            page_data[i] = map + p;
            p += signature_sizes[i] * page_size;
        }
        return "";
    }
};

}  // namespace cobsgpu
