/*
 * include/cobsgpu.h -- C ABI of libcobsgpu.so, the B200 (sm_100a) implementation of the
 * COBS query hot path:  per-k-mer XXH64 -> signature-row lookup -> AND of the h selected
 * bit-rows -> per-document hit counts -> threshold / ordered top-k.
 *
 * Plain C: opaque handles, pointers and sizes, `int` status codes plus
 * cobsgpu_last_error().  No exceptions, STL or torch types cross this boundary.
 * Each entry point cites the reference interface it replaces (paths relative to the
 * bingmann/cobs source tree).  The C++ drop-in classes (cobs::ClassicSearch, ...) in
 * cobs_b200/host/ and the Python/ctypes binding in cobs_b200/ sit on top of exactly these
 * symbols; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * There is NO CPU fallback: every compute entry point returns COBSGPU_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef COBSGPU_H
#define COBSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COBSGPU_VERSION 2

/* status codes */
#define COBSGPU_OK 0
#define COBSGPU_ERR_INVALID_ARG 1
#define COBSGPU_ERR_CUDA 2
#define COBSGPU_ERR_OOM 3
/* query shorter than term_size: the reference prints "query too short ..." and exits
 * (cobs/query/classic_search.cpp:431-433) */
#define COBSGPU_ERR_QUERY_TOO_SHORT 4
/* non-ACGT base in a canonicalising index: the reference die()s
 * (cobs/query/classic_search.cpp:93-96) */
#define COBSGPU_ERR_INVALID_BASE 5
/* bad magic / version: FileIOException in the reference (cobs/file/header.hpp:23-53) */
#define COBSGPU_ERR_BAD_FILE 6
#define COBSGPU_ERR_IO 7

/* per-query counts of the device-resident path that flag a query instead of a list */
#define COBSGPU_COUNT_OVERFLOW 0xFFFFFFFFu
#define COBSGPU_COUNT_INVALID 0xFFFFFFFEu

#define COBSGPU_KIND_CLASSIC 0
#define COBSGPU_KIND_COMPACT 1

typedef struct cobsgpu_index cobsgpu_index;

/*
 * Description of one index to load into HBM.  Mirrors what an IndexSearchFile exposes
 * (cobs/query/index_file.hpp:19-35) plus the raw bit matrix the mmap variants read
 * (cobs/query/classic_index/mmap_search_file.cpp:19-42,
 *  cobs/query/compact_index/mmap_search_file.cpp:16-67).
 * A classic index is described as ONE page with page_size = row_size = ceil(n_docs/8).
 */
typedef struct cobsgpu_index_desc {
    uint32_t struct_size;       /* sizeof(cobsgpu_index_desc) */
    int32_t kind;               /* COBSGPU_KIND_* */
    uint32_t term_size;         /* k */
    uint32_t canonicalize;      /* 0 or 1 */
    uint32_t num_hashes;        /* h */
    uint32_t n_docs;            /* real documents (file_names().size()) */
    uint32_t n_pages;           /* classic: 1 */
    uint32_t reserved0;
    uint64_t page_size;         /* bytes per row per page in the source layout */
    const uint64_t* signature_sizes;  /* [n_pages] rows per page */
    /* [n_pages] host pointers, row-major, row stride page_size; NULL => procedural
     * bits from fill_seed (synthetic benchmark indices; same function as
     * oracle_fill_word) */
    const uint8_t* const* page_data;
    uint64_t fill_seed;
    int32_t device;             /* CUDA device ordinal */
    /* document-axis shard held by this handle: shard_index of shard_count.
     * Classic: contiguous column ranges cut at multiples of 128 documents; compact: the same
     * column range of every page (pages of >= 64 bytes per shard), else whole pages dealt out
     * so that page count (work) and bytes (HBM) are both balanced.
     * Results carry GLOBAL document ids. */
    uint32_t shard_index;
    uint32_t shard_count;
    uint32_t reserved1;
} cobsgpu_index_desc;

typedef struct cobsgpu_index_info {
    int32_t kind;
    uint32_t term_size;
    uint32_t canonicalize;
    uint32_t num_hashes;
    uint32_t n_docs;        /* real documents of the whole index */
    uint32_t n_pages;
    uint64_t page_size;     /* IndexSearchFile::page_size(): classic 1, compact header page_size */
    uint64_t row_size;      /* IndexSearchFile::row_size() */
    uint64_t counts_size;   /* IndexSearchFile::counts_size() */
    uint32_t shard_index;
    uint32_t shard_count;
    uint32_t shard_doc_begin;   /* lowest global column held by this shard */
    uint32_t shard_doc_end;     /* one past the highest column held (padded columns included);
                                   compact shards need not be contiguous in between */
    uint64_t hbm_bytes;         /* bytes of signature matrix resident on the device */
    uint64_t bytes_per_kmer;    /* algorithmic bytes per query k-mer for THIS shard:
                                   h * (unpadded row bytes held) */
    /* loader statistics of a file / host-array open (0 for synthetic indices): matrix bytes read
     * from the source, wall-clock seconds from the first read to the last byte in HBM, host
     * threads used (replaces the progress lines of cobs/util/query.cpp:56-86) */
    double load_seconds;
    uint64_t load_bytes;
    uint32_t load_threads;
    uint32_t reserved;
} cobsgpu_index_info;

/* Result lists of one batch, CSR.  Query q owns entries [offsets[q], offsets[q+1]),
 * ordered like counts_to_result (cobs/query/classic_search.cpp:109-202): score
 * descending, then document ascending.  Host memory owned by the index handle, valid
 * until the next call on that handle. */
typedef struct cobsgpu_result {
    const uint64_t* offsets; /* [nq + 1] */
    const uint32_t* doc;     /* global document ids */
    const uint32_t* score;
} cobsgpu_result;

/* Accumulated device-side phase times (CUDA events) and launch counts since the last
 * reset.  Phase names follow the reference's Timer keys
 * (cobs/query/classic_search.cpp:329,380,386,392): "hashes" = K1, "io"+"and rows"+
 * "add rows" = the fused score kernel, "sort results" = select. */
typedef struct cobsgpu_timers {
    double hashes_ms;
    double score_ms;
    double select_ms;
    double h2d_ms;
    double d2h_ms;
    uint64_t kernel_launches;
    uint64_t score_launches;
    uint64_t kmers;          /* query k-mers processed */
    uint64_t queries;
} cobsgpu_timers;

const char* cobsgpu_last_error(void);
int cobsgpu_version(void);
/* number of usable CUDA devices (0 => every compute call fails with COBSGPU_ERR_CUDA) */
int cobsgpu_device_count(void);

/* replaces: IndexSearchFile construction + initialize_mmap (cobs/util/query.cpp:38-88) */
int cobsgpu_index_open(const cobsgpu_index_desc* desc, cobsgpu_index** out);
/* replaces: ClassicSearch(std::string path) auto-detection (cobs/query/classic_search.cpp:51-64),
 * header parsing (cobs/file/classic_index_header.cpp:38-50, compact_index_header.cpp:44-65) */
int cobsgpu_index_open_file(const char* path, int device, uint32_t shard_index,
                            uint32_t shard_count, cobsgpu_index** out);
/*
 * Classic index construction on the device ("next" row f4 of the scope table): documents are
 * given as in-memory sequences (a document = one or more sequences, e.g. the records of a
 * FASTA file; k-mers are all windows of each sequence and never span two sequences).
 * replaces: classic_construct / process_batch / process_term
 * (cobs/construction/classic_index.cpp:39-130, 565-600), signature sizing
 * (cobs/util/calc_signature_size.cpp:15-33).  Reading document files stays out of scope.
 * The result is an ordinary, immediately searchable index handle.
 */
typedef struct cobsgpu_construct_desc {
    uint32_t struct_size;       /* sizeof(cobsgpu_construct_desc) */
    uint32_t term_size;
    uint32_t canonicalize;
    uint32_t num_hashes;
    uint64_t signature_size;    /* rows; 0 => sized from false_positive_rate and the largest document */
    double false_positive_rate;
    uint32_t n_docs;
    uint32_t n_seqs;
    const char* const* doc_names;   /* [n_docs] */
    const char* sequences;          /* all sequences, concatenated */
    const uint64_t* seq_offsets;    /* [n_seqs + 1] */
    const uint32_t* seq_doc;        /* [n_seqs] document index of every sequence */
    int32_t device;
    uint32_t reserved;
} cobsgpu_construct_desc;

int cobsgpu_construct_classic(const cobsgpu_construct_desc* desc, cobsgpu_index** out);
/* writes an unsharded index in the reference's on-disk format
 * (cobs/file/classic_index_header.cpp:26-36, cobs/file/compact_index_header.cpp:24-42) */
int cobsgpu_index_save(const cobsgpu_index* idx, const char* path);
void cobsgpu_index_close(cobsgpu_index* idx);
int cobsgpu_index_get_info(const cobsgpu_index* idx, cobsgpu_index_info* out);
/* rows of global page `page` (the header's signature_size; classic: page 0); 0 if out of range */
uint64_t cobsgpu_index_signature_size(const cobsgpu_index* idx, uint32_t page);
/* IndexSearchFile::file_names()[doc].c_str(); NULL for synthetic indices */
const char* cobsgpu_index_doc_name(const cobsgpu_index* idx, uint32_t doc);

/* options: "max_candidates" (per-query candidate slots of the fused threshold path,
 * default 1024), "max_batch" (queries per device batch, default 16384),
 * "workspace_mb" (bound for the exhaustive path, default 1024), "pipe_kb" (result KB per
 * device-to-host copy when lists of every document are returned in pipelined sub-batches,
 * default 32768), "pinned_max_mb" (lists of every document are handed out from page-locked
 * memory up to this many MB per call, default 2048; larger results use pageable arrays),
 * "timing" (0/1),
 * "prefetch" (0/1: cobsgpu_search_batch_device runs the metadata upload + K1 of a call on an
 * internal stream, double-buffered, so that they overlap the previous call's K2),
 * "inputs_ready" (0/1: with prefetch, the caller guarantees d_queries is already complete --
 * otherwise the internal stream first waits for the caller's stream),
 * "input_stream" (with prefetch and inputs_ready 0: the cudaStream_t, passed as an integer, on
 * which the caller uploads d_queries -- K1 then waits for that stream instead of queueing
 * behind the caller's compute stream; -1 = none) */
int cobsgpu_set_option(cobsgpu_index* idx, const char* name, int64_t value);

/*
 * Queries are passed as one blob of ASCII characters plus nq+1 offsets
 * (query q = blob[offsets[q] .. offsets[q+1])).
 */

/* K1 only.  replaces: create_hashes (cobs/query/classic_search.cpp:66-107).
 * out receives, query after query, (len_q - k + 1) * h raw 64-bit XXH64 values. */
int cobsgpu_hash(cobsgpu_index* idx, const char* queries, const uint64_t* offsets,
                 uint32_t nq, uint64_t* out_hashes);

/* K1 + K2, exhaustive.  replaces: search_index_file (cobs/query/classic_search.cpp:309-401):
 * out_scores[q * counts_size + d] = hit count of column d (padded columns included, the
 * reference's score_list layout).  Columns outside this handle's shard are left untouched. */
int cobsgpu_scores(cobsgpu_index* idx, const char* queries, const uint64_t* offsets,
                   uint32_t nq, uint32_t* out_scores);

/* K1 + K2 + K3.  replaces: ClassicSearch::search (cobs/query/classic_search.cpp:403-505)
 * for a batch of queries over one index: documents with score >= ceil(threshold * T_q),
 * ordered (score desc, doc asc), at most num_results per query (0 = all). */
int cobsgpu_search_batch(cobsgpu_index* idx, const char* queries,
                         const uint64_t* offsets, uint32_t nq, double threshold,
                         uint64_t num_results, cobsgpu_result* out);

/*
 * Asynchronous pair for callers that stream batches from host buffers: cobsgpu_submit enqueues
 * one batch (upload + K1 on an input stream, K2 + K3 on the main stream, result download on an
 * output stream) and returns at once; cobsgpu_collect waits for it and hands out its lists.  Up
 * to 4 tickets may be outstanding per handle, so the upload of batch i+1 and the download of
 * batch i-1 overlap the score kernel of batch i (cobsgpu_search_batch pipelines its own
 * sub-batches the same way).  `queries` must stay valid until the ticket is collected when it
 * is page-locked memory (pageable memory is staged by the driver before submit returns).
 * Tickets are collected in any order; the result arrays are valid until the handle has
 * accepted 4 more batches.  Calls on one handle must come from one thread.
 * replaces: the per-query loop around Search::search in process_query (src/cobs.cpp:425-462).
 */
typedef uint64_t cobsgpu_ticket;
int cobsgpu_submit(cobsgpu_index* idx, const char* queries, const uint64_t* offsets,
                   uint32_t nq, double threshold, uint64_t num_results,
                   cobsgpu_ticket* ticket);
int cobsgpu_collect(cobsgpu_index* idx, cobsgpu_ticket ticket, cobsgpu_result* out);

/*
 * Device-resident variant used by the multi-GPU path and by bench.py's "value" leg:
 * d_queries is a DEVICE pointer (offsets stay on the host, they drive the launch
 * geometry).  Results stay on the device in caller-provided buffers:
 *   d_counts[nq]            number of results of query q (<= results_per_query)
 *   d_keys[nq * results_per_query]   sorted keys; key = (~score << 32) | doc, ascending
 * A query whose candidates overflowed the per-query candidate slots
 * (max(results_per_query, "max_candidates")) gets d_counts[q] = COBSGPU_COUNT_OVERFLOW
 * instead of an incomplete list (nothing is silently dropped; redo it through
 * cobsgpu_search_batch); a query with a non-ACGT base in a canonicalising index gets
 * COBSGPU_COUNT_INVALID (the host entry points report COBSGPU_ERR_INVALID_BASE for it).
 * A query whose result list is longer than results_per_query is flagged COBSGPU_COUNT_OVERFLOW
 * as well (a cut list never looks like a complete one).  With 1 <= num_results <= 1024 the
 * per-warp top-k epilogue bounds every list, so nothing overflows when results_per_query >=
 * num_results.  Queries are limited to 65 535 k-mers on this path.
 * All work is enqueued on `stream` (a cudaStream_t, may be 0) and is asynchronous; calls on
 * one handle must be issued from one thread, and the stream must be synchronised before the
 * host-buffer entry points are used on the same handle.
 */
int cobsgpu_search_batch_device(cobsgpu_index* idx, const char* d_queries,
                                const uint64_t* offsets, uint32_t nq,
                                double threshold, uint64_t num_results,
                                uint32_t results_per_query, uint32_t* d_counts,
                                uint64_t* d_keys, void* stream);

/* K3 merge for document-sharded indices: n_lists per-query lists (e.g. the all-gathered
 * outputs of cobsgpu_search_batch_device from every rank) are merged into one ordered list
 * per query of at most min(num_results or all, out_per_query) entries.  List l has its counts
 * [nq] at d_counts + l * counts_list_stride (uint32 elements) and its keys
 * [nq][results_per_query] at d_keys + l * keys_list_stride (uint64 elements); a stride of 0
 * means densely packed ([n_lists][nq] / [n_lists][nq][results_per_query]).  A flagged count
 * (COBSGPU_COUNT_*) in any list propagates to the output.  Device pointers; asynchronous on `stream`. */
int cobsgpu_merge_device(int device, uint32_t n_lists, uint32_t nq,
                         uint32_t results_per_query, const uint32_t* d_counts,
                         uint64_t counts_list_stride, const uint64_t* d_keys,
                         uint64_t keys_list_stride, uint64_t num_results,
                         uint32_t out_per_query, uint32_t* d_out_counts,
                         uint64_t* d_out_keys, void* stream);

/*
 * Multi-GPU inside ONE process ("a single process with one stream per GPU
 * is sufficient on one NVSwitch host"): a group holds the n_devices document-axis shards of one
 * index, shard g on devices[g] (classic: contiguous column ranges; compact: whole pages).
 * cobsgpu_group_search_batch runs K1-K3 on every shard concurrently, then the leader (shard 0)
 * merges the shards' fixed-size result blocks [nq][k] with ONE kernel that reads them straight
 * out of the peers' HBM over NVLink (peer access; staged copies when the devices cannot map
 * each other) and returns one ordered list per query -- same semantics and result layout as
 * cobsgpu_search_batch on an unsharded index.  A query that overflows a shard's candidate
 * slots is redone on every shard's exhaustive path and merged on the host: nothing is dropped.
 * replaces: the column-batch parallel_for of search_index_file
 * (cobs/query/classic_search.cpp:355-400) spread over GPUs instead of threads.
 */
typedef struct cobsgpu_group cobsgpu_group;
int cobsgpu_group_open_file(const char* path, const int32_t* devices, uint32_t n_devices,
                            cobsgpu_group** out);
/* desc->device / shard_index / shard_count are ignored (set per shard) */
int cobsgpu_group_open(const cobsgpu_index_desc* desc, const int32_t* devices, uint32_t n_devices,
                       cobsgpu_group** out);
void cobsgpu_group_close(cobsgpu_group* grp);
uint32_t cobsgpu_group_size(const cobsgpu_group* grp);
/* shard i of the group (borrowed): geometry, document names, timers, options */
cobsgpu_index* cobsgpu_group_shard(cobsgpu_group* grp, uint32_t i);
int cobsgpu_group_search_batch(cobsgpu_group* grp, const char* queries, const uint64_t* offsets,
                               uint32_t nq, double threshold, uint64_t num_results,
                               cobsgpu_result* out);

int cobsgpu_get_timers(const cobsgpu_index* idx, cobsgpu_timers* out);
int cobsgpu_reset_timers(cobsgpu_index* idx);

/* Raw access for tests and the loader: copy `bytes` bytes of row `row` of local page
 * `page` (shard-local column bytes, starting at byte `begin`) back to the host. */
int cobsgpu_debug_read_row(cobsgpu_index* idx, uint32_t page, uint64_t row,
                           uint64_t begin, uint64_t bytes, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* COBSGPU_H */
