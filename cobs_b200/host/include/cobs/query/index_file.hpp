// cobs/query/index_file.hpp -- the row provider behind a search.  IndexSearchFile is the
// reference's abstract interface (cobs/query/index_file.hpp:19-35), member for member, so that
// callers holding std::shared_ptr<IndexSearchFile> compile unchanged.  In the B200 build the
// concrete index classes derive from HbmIndexSearchFile: the signature matrix lives in HBM (one
// cobsgpu handle, or one group of document-axis shards over several GPUs) and the gather that
// read_from_disk() performs on the CPU is fused into the score kernel.
#pragma once
#include <cobs/util/fs.hpp>

#include <cstdint>
#include <string>
#include <vector>

struct cobsgpu_index;
struct cobsgpu_group;

namespace cobs {

class IndexSearchFile
{
public:
    virtual ~IndexSearchFile() = default;

    virtual void read_from_disk(
        const std::vector<size_t>& hashes, uint8_t* rows,
        size_t begin, size_t size, size_t buffer_size) = 0;

    virtual uint32_t term_size() const = 0;
    virtual uint8_t canonicalize() const = 0;
    virtual uint64_t row_size() const = 0;
    virtual uint64_t page_size() const = 0;
    virtual uint64_t num_hashes() const = 0;
    virtual uint64_t counts_size() const = 0;
    virtual const std::vector<std::string>& file_names() const = 0;
};

//! An index resident in HBM.  The plug point for custom row sources: construct it from
//! in-memory pages (the raw matrix in the reference's row-major, LSB-first layout) instead of a
//! file and hand it to ClassicSearch like any other IndexSearchFile.
class HbmIndexSearchFile : public IndexSearchFile
{
public:
    //! description of an in-memory index (classic: one page, page_size = ceil(n_docs / 8))
    struct Pages {
        bool compact = false;
        uint32_t term_size = 31;
        uint8_t canonicalize = 1;
        uint64_t num_hashes = 1;
        uint64_t page_size = 0;                     //!< compact only
        std::vector<std::string> file_names;        //!< one per document
        std::vector<uint64_t> signature_sizes;      //!< rows per page
        std::vector<const uint8_t*> page_data;      //!< [signature_size][page_size] bytes each
    };
    explicit HbmIndexSearchFile(const Pages& pages);
    ~HbmIndexSearchFile() override;
    HbmIndexSearchFile(const HbmIndexSearchFile&) = delete;
    HbmIndexSearchFile& operator = (const HbmIndexSearchFile&) = delete;

    //! source compatibility / debugging only: copies rows back from HBM (single-GPU classic)
    void read_from_disk(
        const std::vector<size_t>& hashes, uint8_t* rows,
        size_t begin, size_t size, size_t buffer_size) override;

    uint32_t term_size() const override { return term_size_; }
    uint8_t canonicalize() const override { return canonicalize_; }
    uint64_t row_size() const override { return row_size_; }
    uint64_t page_size() const override { return page_size_; }
    uint64_t num_hashes() const override { return num_hashes_; }
    uint64_t counts_size() const override { return counts_size_; }
    const std::vector<std::string>& file_names() const override { return file_names_; }

    //! the document-axis shards of this index, one per GPU
    const std::vector<cobsgpu_index*>& gpu_shards() const { return shards_; }
    //! non-null when the index is spread over several GPUs (gopt_gpus > 1)
    cobsgpu_group* gpu_group() const { return group_; }

protected:
    //! loads `path` into HBM; kind: 0 classic, 1 compact.
    //! Throws FileIOException on a wrong magic word / version.
    HbmIndexSearchFile(const fs::path& path, int kind);
    void read_geometry();

    uint32_t term_size_ = 0;
    uint8_t canonicalize_ = 0;
    uint64_t row_size_ = 0, page_size_ = 0, num_hashes_ = 0, counts_size_ = 0;
    std::vector<std::string> file_names_;
    std::vector<cobsgpu_index*> shards_;
    cobsgpu_group* group_ = nullptr;
    std::vector<uint64_t> signature_sizes_;
};

} // namespace cobs
