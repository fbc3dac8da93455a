// HbmIndexSearchFile: loads an index (file or in-memory pages) into HBM through the C ABI and
// exposes the reference's getters (cobs/query/index_file.hpp:19-35,
// classic_index/search_file.cpp:15-23, compact_index/search_file.cpp:15-32).
#include <cobs/file/file_io_exception.hpp>
#include <cobs/query/index_file.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>

#include <cobsgpu.h>

#include <cstring>

namespace cobs {

namespace {

[[noreturn]] void open_failed(int rc) {
    std::string msg = cobsgpu_last_error();
    if (rc == COBSGPU_ERR_BAD_FILE) throw FileIOException(msg);
    if (rc == COBSGPU_ERR_IO) exit_error(msg);   // like open_file(): print + exit
    throw std::runtime_error("cobs GPU index: " + msg);
}

} // namespace

HbmIndexSearchFile::HbmIndexSearchFile(const fs::path& path, int kind) {
    // the reference throws FileIOException("invalid file type") when the magic word does not
    // match the class that was asked for (cobs/file/header.hpp:23-29)
    if (kind == 0 && !file_has_header<ClassicIndexHeader>(path))
        throw FileIOException("invalid file type");
    if (kind == 1 && !file_has_header<CompactIndexHeader>(path))
        throw FileIOException("invalid file type");
    const unsigned n = gopt_gpus ? gopt_gpus : 1;
    if (n == 1) {
        cobsgpu_index* h = nullptr;
        int rc = cobsgpu_index_open_file(path.string().c_str(), gopt_gpu_device, 0, 1, &h);
        if (rc != COBSGPU_OK) open_failed(rc);
        shards_.push_back(h);
    }
    else {
        // document-axis shards over n GPUs of this process, merged on the leader GPU
        std::vector<int32_t> devices(n);
        for (unsigned s = 0; s < n; ++s) devices[s] = gopt_gpu_device + int32_t(s);
        int rc = cobsgpu_group_open_file(path.string().c_str(), devices.data(), n, &group_);
        if (rc != COBSGPU_OK) open_failed(rc);
        for (unsigned s = 0; s < n; ++s) shards_.push_back(cobsgpu_group_shard(group_, s));
    }
    read_geometry();
    cobsgpu_index_info info;
    cobsgpu_index_get_info(shards_[0], &info);
    file_names_.resize(info.n_docs);
    for (uint32_t d = 0; d < info.n_docs; ++d) file_names_[d] = cobsgpu_index_doc_name(shards_[0], d);
}

HbmIndexSearchFile::HbmIndexSearchFile(const Pages& p) {
    if (p.signature_sizes.empty() || p.signature_sizes.size() != p.page_data.size())
        throw std::runtime_error("HbmIndexSearchFile: one signature size and one data pointer per page");
    cobsgpu_index_desc d;
    std::memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.kind = p.compact ? COBSGPU_KIND_COMPACT : COBSGPU_KIND_CLASSIC;
    d.term_size = p.term_size;
    d.canonicalize = p.canonicalize;
    d.num_hashes = uint32_t(p.num_hashes);
    d.n_docs = uint32_t(p.file_names.size());
    d.n_pages = uint32_t(p.signature_sizes.size());
    d.page_size = p.compact ? p.page_size : (p.file_names.size() + 7) / 8;
    d.signature_sizes = p.signature_sizes.data();
    d.page_data = p.page_data.data();
    d.device = gopt_gpu_device;
    d.shard_index = 0;
    d.shard_count = 1;
    const unsigned n = gopt_gpus ? gopt_gpus : 1;
    if (n == 1) {
        cobsgpu_index* h = nullptr;
        int rc = cobsgpu_index_open(&d, &h);
        if (rc != COBSGPU_OK) open_failed(rc);
        shards_.push_back(h);
    }
    else {
        std::vector<int32_t> devices(n);
        for (unsigned s = 0; s < n; ++s) devices[s] = gopt_gpu_device + int32_t(s);
        int rc = cobsgpu_group_open(&d, devices.data(), n, &group_);
        if (rc != COBSGPU_OK) open_failed(rc);
        for (unsigned s = 0; s < n; ++s) shards_.push_back(cobsgpu_group_shard(group_, s));
    }
    read_geometry();
    file_names_ = p.file_names;
}

void HbmIndexSearchFile::read_geometry() {
    cobsgpu_index_info info;
    cobsgpu_index_get_info(shards_[0], &info);
    term_size_ = info.term_size;
    canonicalize_ = static_cast<uint8_t>(info.canonicalize);
    row_size_ = info.row_size;
    page_size_ = info.page_size;
    num_hashes_ = info.num_hashes;
    counts_size_ = info.counts_size;
    for (uint32_t p = 0; p < info.n_pages; ++p)
        signature_sizes_.push_back(cobsgpu_index_signature_size(shards_[0], p));
}

HbmIndexSearchFile::~HbmIndexSearchFile() {
    if (group_) cobsgpu_group_close(group_);   // owns its shards
    else
        for (cobsgpu_index* h : shards_) cobsgpu_index_close(h);
}

//! Source-compatibility shim: copies `size` bytes starting at byte `begin` of the rows selected
//! by the raw hashes back from HBM, laid out like the reference's rows buffer.  Classic
//! indices only (the modulo is per page for compact ones); not on the search path.
void HbmIndexSearchFile::read_from_disk(
    const std::vector<size_t>& hashes, uint8_t* rows,
    size_t begin, size_t size, size_t buffer_size) {
    if (page_size_ != 1 || shards_.size() != 1)
        die_with_message("read_from_disk() is only kept for single-GPU classic indices");
    for (size_t i = 0; i < hashes.size(); ++i) {
        uint64_t row = hashes[i] % signature_sizes_[0];
        if (cobsgpu_debug_read_row(shards_[0], 0, row, begin, size, rows + i * buffer_size) != COBSGPU_OK)
            die_with_message(cobsgpu_last_error());
    }
}

} // namespace cobs
