// cobs/query/index_file.hpp -- the row provider behind a search, drop-in for the reference's
// cobs/query/index_file.hpp:19-35.  In the B200 build an IndexSearchFile owns the signature
// matrix in HBM (one handle per GPU when sharded along the document axis); the gather that
// read_from_disk() performs on the CPU is fused into the score kernel, so read_from_disk()
// only exists for source compatibility and debugging (it copies rows back from the device).
#pragma once
#include <cobs/util/fs.hpp>

#include <cstdint>
#include <string>
#include <vector>

struct cobsgpu_index;

namespace cobs {

class IndexSearchFile
{
public:
    virtual ~IndexSearchFile();
    IndexSearchFile(const IndexSearchFile&) = delete;
    IndexSearchFile& operator = (const IndexSearchFile&) = delete;

    virtual void read_from_disk(
        const std::vector<size_t>& hashes, uint8_t* rows,
        size_t begin, size_t size, size_t buffer_size);

    virtual uint32_t term_size() const { return term_size_; }
    virtual uint8_t canonicalize() const { return canonicalize_; }
    virtual uint64_t row_size() const { return row_size_; }
    virtual uint64_t page_size() const { return page_size_; }
    virtual uint64_t num_hashes() const { return num_hashes_; }
    virtual uint64_t counts_size() const { return counts_size_; }
    virtual const std::vector<std::string>& file_names() const { return file_names_; }

    //! the document-axis shards of this index, one per GPU
    const std::vector<cobsgpu_index*>& gpu_shards() const { return shards_; }

protected:
    //! loads `path` into HBM; kind: 0 classic, 1 compact, -1 auto-detect.
    //! Throws FileIOException on a wrong magic word / version.
    IndexSearchFile(const fs::path& path, int kind);

    uint32_t term_size_ = 0;
    uint8_t canonicalize_ = 0;
    uint64_t row_size_ = 0, page_size_ = 0, num_hashes_ = 0, counts_size_ = 0;
    std::vector<std::string> file_names_;
    std::vector<cobsgpu_index*> shards_;
    std::vector<uint64_t> signature_sizes_;
};

} // namespace cobs
