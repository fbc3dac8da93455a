/*
 * oracle/cobs_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the COBS query hot path,
 * written from the reference's published behaviour.  It exists only to CHECK
 * the CUDA path: it may be imported/linked by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg -- never by cobs_b200/ (the product fails
 * loudly when the CUDA library is missing; there is no CPU fallback).
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this file against
 *   - the 8 XXH64 known-answer vectors of extlib/xxhash/xxhsum.c:442-470,
 *   - the 7 canonicalisation vectors of tests/util.cpp:38-59,
 *   - the python end-to-end known answer (python/tests/test_cobs_index.py:36-61),
 *   - golden (doc, score) lists produced by the real reference (oracle/_ref,
 *     built from /root/reference by oracle/Makefile) and committed under
 *     tests/golden/ together with the generating script.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef COBS_ORACLE_H
#define COBS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_OK 0
#define ORACLE_ERR_INVALID_BASE (-1) /* classic_search.cpp:93-96 die()          */
#define ORACLE_ERR_TOO_SHORT (-2)    /* classic_search.cpp:431-433 assert_exit  */
#define ORACLE_ERR_BAD_FILE (-3)     /* file/header.hpp:23-53 FileIOException   */
#define ORACLE_ERR_BAD_PARAM (-4)
#define ORACLE_ERR_IO (-5)

#define ORACLE_KIND_CLASSIC 0
#define ORACLE_KIND_COMPACT 1

/* Geometry + bit data of one index, host pointers.  Classic indices are
 * described as a single page whose page_size is the row size
 * (classic_index/search_file.hpp:17-36); compact indices carry P pages with
 * their own modulus (compact_index/search_file.cpp:15-32). */
typedef struct oracle_index {
    int kind;
    uint32_t term_size;
    uint8_t canonicalize;
    uint64_t num_hashes;
    uint32_t n_docs;     /* real documents (file_names().size()) */
    uint32_t n_pages;    /* classic: 1 */
    uint64_t page_size;  /* bytes per row per page; classic: ceil(n_docs/8) */
    uint64_t* signature_sizes; /* [n_pages] rows per page */
    uint8_t** page_data;       /* [n_pages] host pointers, row-major */
    char** doc_names;          /* [n_docs] or NULL for procedural indices */
    /* procedural indices: page_data == NULL and bits come from oracle_fill_word */
    int procedural;
    uint64_t fill_seed;
    /* ownership bookkeeping for oracle_index_free */
    uint8_t* file_blob;
    size_t file_size;
    int owns_pages;
} oracle_index;

/* extlib/xxhash/xxhash.c:665-878 (XXH64 v0.6.5) */
uint64_t oracle_xxh64(const void* data, size_t len, uint64_t seed);

/* cobs/util/query.cpp:143-199; returns 1 = good, 0 = contained non-ACGT */
int oracle_canonicalize_kmer(const char* input, char* output, size_t size);

/* cobs/query/classic_search.cpp:66-107.  out has num_hashes * (len-k+1)
 * entries, raw 64-bit values (no modulo). */
int oracle_create_hashes(const char* query, size_t len, uint32_t term_size,
                         uint64_t num_hashes, uint8_t canonicalize,
                         uint64_t* out);

/* counts_size(): classic 8*row_size, compact 8*P*page_size
 * (classic_index/search_file.cpp:21-23, compact_index/search_file.cpp:30-32) */
uint64_t oracle_counts_size(const oracle_index* idx);

/* Per-document hit counts of one query over one index:
 * read_from_disk + aggregate_rows + compute_counts
 * (classic_index/mmap_search_file.cpp:27-40, compact_index/mmap_search_file.cpp:34-67,
 *  classic_search.cpp:279-307 and 213-275).  scores has counts_size entries
 * (padded columns included, exactly like the reference's score_list). */
int oracle_scores(const oracle_index* idx, const char* query, size_t len,
                  uint32_t* scores);

/* Same, restricted to document columns [doc_begin, doc_end) (multiples of 8);
 * used for sampled column blocks on procedural full-size indices. */
int oracle_scores_range(const oracle_index* idx, const char* query, size_t len,
                        uint64_t doc_begin, uint64_t doc_end, uint32_t* scores);

/* ClassicSearch::search + counts_to_result (classic_search.cpp:403-505, 109-202)
 * over n_idx indices.  Outputs at most out_cap entries; *out_count receives the
 * number of results the reference would return.  out_file may be NULL. */
int oracle_search(const oracle_index* const* idx, size_t n_idx,
                  const char* query, size_t len, double threshold,
                  size_t num_results, uint32_t* out_file, uint32_t* out_doc,
                  uint32_t* out_score, size_t out_cap, size_t* out_count);

/* On-disk formats (file/classic_index_header.cpp:26-50,
 * file/compact_index_header.cpp:20-65).  Loads the whole file into memory. */
int oracle_index_load(const char* path, oracle_index* out);
void oracle_index_free(oracle_index* idx);

/* Test-fixture writers in the reference's formats. names may be NULL
 * ("doc_%06u" is used). data: classic = sig*row_size bytes; compact = pages
 * back to back (sig[p]*page_size each). */
int oracle_write_classic(const char* path, uint32_t term_size,
                         uint8_t canonicalize, uint32_t n_docs,
                         uint64_t signature_size, uint64_t num_hashes,
                         const char* const* names, const uint8_t* data);
int oracle_write_compact(const char* path, uint32_t term_size,
                         uint8_t canonicalize, uint32_t n_docs,
                         uint64_t page_size, uint32_t n_pages,
                         const uint64_t* signature_sizes, uint64_t num_hashes,
                         const char* const* names, const uint8_t* data);

/* Procedural index bits shared with the CUDA fill kernel
 * (cobs_b200/csrc/fill.cuh): 8 document-bytes (64 columns) of row `row`,
 * page `page`, starting at byte 8*word.  Density 1/4 (two mixed words ANDed). */
uint64_t oracle_fill_word(uint64_t seed, uint32_t page, uint64_t row,
                          uint64_t word);
/* Build a procedural index descriptor (no memory for bits). */
int oracle_index_procedural(oracle_index* out, int kind, uint32_t term_size,
                            uint8_t canonicalize, uint64_t num_hashes,
                            uint32_t n_docs, uint64_t page_size,
                            uint32_t n_pages, const uint64_t* signature_sizes,
                            uint64_t fill_seed);
/* Materialise a procedural index into host memory (small ones only). */
int oracle_index_materialize(oracle_index* idx);
/* Stream a procedural classic index to a file in the reference's format. */
int oracle_write_classic_procedural(const char* path, const oracle_index* idx);
int oracle_write_compact_procedural(const char* path, const oracle_index* idx);

/* cobs/util/misc.hpp:31-38 random_sequence_rng with a 64-bit LCG of our own
 * (not std::default_random_engine; only used for synthetic workloads). */
void oracle_random_query(uint64_t seed, size_t len, char* out);

#ifdef __cplusplus
}
#endif
#endif /* COBS_ORACLE_H */
