/*
 * oracle/cobs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See cobs_oracle.h for scope and parity status (PINNED against the
 * reference's golden vectors and against oracle/_ref).
 *
 * Plain C restatement of the COBS CPU query path; every function cites the
 * reference file:line (relative to /root/reference) whose behaviour it states.
 */
#define _POSIX_C_SOURCE 200809L
#include "cobs_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* XXH64 -- extlib/xxhash/xxhash.c (v0.6.5)                                  */

/* primes: xxhash.c:665-669 */
static const uint64_t P1 = 11400714785074694791ULL;
static const uint64_t P2 = 14029467366897019727ULL;
static const uint64_t P3 = 1609587929392839161ULL;
static const uint64_t P4 = 9650029242287828579ULL;
static const uint64_t P5 = 2870177450012600261ULL;

static inline uint64_t rotl64(uint64_t x, int r) {
    return (x << r) | (x >> (64 - r));
}
static inline uint64_t rd64(const uint8_t* p) { /* little-endian read */
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
static inline uint32_t rd32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
           ((uint32_t)p[3] << 24);
}
/* xxhash.c:671-677 */
static inline uint64_t xxh_round(uint64_t acc, uint64_t input) {
    acc += input * P2;
    acc = rotl64(acc, 31);
    return acc * P1;
}
/* xxhash.c:679-685 */
static inline uint64_t xxh_merge(uint64_t acc, uint64_t val) {
    val = xxh_round(0, val);
    acc ^= val;
    return acc * P1 + P4;
}

/* xxhash.c:811-852 (body), 701-805 (finalize), 687-695 (avalanche) */
uint64_t oracle_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t* limit = end - 32;
        do {
            v1 = xxh_round(v1, rd64(p));
            v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16));
            v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh_merge(h, v1);
        h = xxh_merge(h, v2);
        h = xxh_merge(h, v3);
        h = xxh_merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        h ^= xxh_round(0, rd64(p));
        h = rotl64(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)rd32(p) * P1;
        h = rotl64(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P5;
        h = rotl64(h, 11) * P1;
        ++p;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------------------------ */
/* canonicalize_kmer -- cobs/util/query.cpp:104-199                          */

static inline char fwd_map(uint8_t c) { /* query.cpp:104-121 */
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? (char)c : 0;
}
static inline char rev_map(uint8_t c) { /* query.cpp:124-141 */
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 0;
    }
}

int oracle_canonicalize_kmer(const char* input, char* output, size_t size) {
    /* Walk inwards from both ends comparing the k-mer with its reverse
     * complement (query.cpp:155-189).  The first differing position decides;
     * only the positions visited before the decision plus the copied bases
     * take part in the validity flag -- exactly as the reference, which stops
     * checking the reverse pointer once the forward strand has won. */
    const uint8_t* in = (const uint8_t*)input;
    int good = 1;
    size_t i = 0;
    for (; i < size / 2; ++i) {
        char f = fwd_map(in[i]);
        char r = rev_map(in[size - 1 - i]);
        output[i] = f;
        good = good && f != 0 && r != 0;
        if (f < r) {
            for (++i; i < size; ++i) {
                char g = fwd_map(in[i]);
                output[i] = g;
                good = good && g != 0;
            }
            return good;
        }
        if (f > r) {
            for (size_t j = 0; j < size; ++j) {
                char x = rev_map(in[j]);
                output[size - 1 - j] = x;
                good = good && x != 0;
            }
            return good;
        }
    }
    for (; i < size; ++i) { /* palindromic so far: keep forward (query.cpp:191-198) */
        char f = fwd_map(in[i]);
        output[i] = f;
        good = good && f != 0;
    }
    return good;
}

/* ------------------------------------------------------------------------ */
/* create_hashes -- cobs/query/classic_search.cpp:66-107                     */

int oracle_create_hashes(const char* query, size_t len, uint32_t term_size,
                         uint64_t num_hashes, uint8_t canonicalize,
                         uint64_t* out) {
    if (len < term_size) return ORACLE_ERR_TOO_SHORT;
    size_t num_terms = len - term_size + 1;
    if (canonicalize == 0) {
        for (size_t i = 0; i < num_terms; ++i)
            for (uint64_t j = 0; j < num_hashes; ++j)
                out[i * num_hashes + j] = oracle_xxh64(query + i, term_size, j);
        return ORACLE_OK;
    }
    if (canonicalize != 1) return ORACLE_ERR_BAD_PARAM;
    char* buf = (char*)malloc(term_size ? term_size : 1);
    for (size_t i = 0; i < num_terms; ++i) {
        if (!oracle_canonicalize_kmer(query + i, buf, term_size)) {
            free(buf);
            return ORACLE_ERR_INVALID_BASE;
        }
        for (uint64_t j = 0; j < num_hashes; ++j)
            out[i * num_hashes + j] = oracle_xxh64(buf, term_size, j);
    }
    free(buf);
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------ */
/* procedural bits                                                           */

static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ULL;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return z;
}

uint64_t oracle_fill_word(uint64_t seed, uint32_t page, uint64_t row,
                          uint64_t word) {
    uint64_t r = mix64(seed ^ mix64(row + ((uint64_t)page << 48)));
    uint64_t a = mix64(r ^ (word * 0xD6E8FEB86659FD93ULL));
    uint64_t b = mix64(a + 0x9E3779B97F4A7C15ULL);
    return a & b;
}

void oracle_random_query(uint64_t seed, size_t len, char* out) {
    static const char bp[4] = { 'A', 'C', 'G', 'T' };
    uint64_t s = seed;
    for (size_t i = 0; i < len; ++i) {
        s += 0x9E3779B97F4A7C15ULL;
        out[i] = bp[mix64(s) >> 62];
    }
}

/* ------------------------------------------------------------------------ */
/* scoring                                                                   */

uint64_t oracle_counts_size(const oracle_index* idx) {
    return 8 * (uint64_t)idx->n_pages * idx->page_size;
}

/* byte b of row `row` in page p */
static inline uint8_t row_byte(const oracle_index* idx, uint32_t p,
                               uint64_t row, uint64_t b) {
    if (idx->procedural)
        return (uint8_t)(oracle_fill_word(idx->fill_seed, p, row, b >> 3) >>
                         (8 * (b & 7)));
    return idx->page_data[p][row * idx->page_size + b];
}

int oracle_scores_range(const oracle_index* idx, const char* query, size_t len,
                        uint64_t doc_begin, uint64_t doc_end,
                        uint32_t* scores) {
    if (len < idx->term_size) return ORACLE_ERR_TOO_SHORT;
    if ((doc_begin & 7) || (doc_end & 7) || doc_end > oracle_counts_size(idx) ||
        doc_begin > doc_end)
        return ORACLE_ERR_BAD_PARAM;
    size_t T = len - idx->term_size + 1;
    uint64_t h = idx->num_hashes;
    uint64_t* hashes = (uint64_t*)malloc(sizeof(uint64_t) * (T * h + 1));
    int rc = oracle_create_hashes(query, len, idx->term_size, h,
                                  idx->canonicalize, hashes);
    if (rc != ORACLE_OK) {
        free(hashes);
        return rc;
    }
    memset(scores, 0, sizeof(uint32_t) * (doc_end - doc_begin));
    uint64_t byte_begin = doc_begin / 8, byte_end = doc_end / 8;
    for (size_t t = 0; t < T; ++t) {
        for (uint64_t gb = byte_begin; gb < byte_end; ++gb) {
            /* page p, byte b within the page row
             * (compact_index/mmap_search_file.cpp:56-66; classic: p = 0) */
            uint32_t p = (uint32_t)(gb / idx->page_size);
            uint64_t b = gb % idx->page_size;
            /* aggregate_rows: AND of the h selected rows (classic_search.cpp:279-307) */
            uint8_t v = 0xFF;
            for (uint64_t j = 0; j < h; ++j) {
                uint64_t row = hashes[t * h + j] % idx->signature_sizes[p];
                v &= row_byte(idx, p, row, b);
            }
            /* compute_counts: bit o of the byte -> document 8*gb + o, LSB first
             * (classic_search.cpp:643-655 and the expansion table 512-641) */
            uint32_t* s = scores + (gb - byte_begin) * 8;
            for (int o = 0; o < 8; ++o) s[o] += (v >> o) & 1u;
        }
    }
    free(hashes);
    return ORACLE_OK;
}

int oracle_scores(const oracle_index* idx, const char* query, size_t len,
                  uint32_t* scores) {
    return oracle_scores_range(idx, query, len, 0, oracle_counts_size(idx),
                               scores);
}

/* ------------------------------------------------------------------------ */
/* search + counts_to_result -- classic_search.cpp:403-505, 109-202          */

typedef struct {
    uint32_t score, file, doc;
} cand_t;

/* comparator of classic_search.cpp:139-143 / 173-177:
 * score descending, then (file, doc) ascending */
static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a;
    const cand_t* y = (const cand_t*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->file != y->file) return x->file < y->file ? -1 : 1;
    if (x->doc != y->doc) return x->doc < y->doc ? -1 : 1;
    return 0;
}

int oracle_search(const oracle_index* const* idx, size_t n_idx,
                  const char* query, size_t len, double threshold,
                  size_t num_results, uint32_t* out_file, uint32_t* out_doc,
                  uint32_t* out_score, size_t out_cap, size_t* out_count) {
    *out_count = 0;
    if (n_idx == 0) return ORACLE_OK; /* classic_search.cpp:410-411 */

    uint64_t total_documents = 0;
    uint32_t max_term = 0;
    for (size_t i = 0; i < n_idx; ++i) { /* 413-429 */
        total_documents += oracle_counts_size(idx[i]);
        if (idx[i]->term_size > max_term) max_term = idx[i]->term_size;
    }
    if (len < max_term) return ORACLE_ERR_TOO_SHORT; /* 431-433 */

    /* 450-451 */
    num_results = num_results == 0 ? total_documents
                  : (num_results < total_documents ? num_results
                                                   : total_documents);

    size_t total_hashes = 0, kept = 0, real_docs = 0;
    for (size_t i = 0; i < n_idx; ++i) real_docs += idx[i]->n_docs;
    cand_t* cands = (cand_t*)malloc(sizeof(cand_t) * (real_docs ? real_docs : 1));

    for (size_t k = 0; k < n_idx; ++k) {
        const oracle_index* ix = idx[k];
        size_t T = len - ix->term_size + 1;
        /* 444-449: threshold in double, ceil, converted to size_t */
        size_t thr = (size_t)ceil(threshold * (double)T);
        total_hashes += T * ix->num_hashes; /* 332 */
        uint64_t cs = oracle_counts_size(ix);
        uint32_t* scores = (uint32_t*)malloc(sizeof(uint32_t) * (cs ? cs : 1));
        int rc = oracle_scores(ix, query, len, scores);
        if (rc != ORACLE_OK) {
            free(scores);
            free(cands);
            return rc;
        }
        /* 121-126 / 164-175: only the real documents, padded columns never */
        for (uint32_t d = 0; d < ix->n_docs; ++d) {
            if ((size_t)scores[d] >= thr) {
                cands[kept].score = scores[d];
                cands[kept].file = (uint32_t)k;
                cands[kept].doc = d;
                ++kept;
            }
        }
        free(scores);
    }
    if (num_results > kept) num_results = kept; /* 128 / 178 */
    /* 130 / 180: the sort is skipped when the query produced <= 1 hash in total
     * (max_counts parameter is total_hashes, classic_search.cpp:468-469) */
    if (total_hashes > 1) qsort(cands, kept, sizeof(cand_t), cand_cmp);

    *out_count = num_results;
    for (size_t i = 0; i < num_results && i < out_cap; ++i) {
        if (out_file) out_file[i] = cands[i].file;
        out_doc[i] = cands[i].doc;
        out_score[i] = cands[i].score;
    }
    free(cands);
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------ */
/* on-disk formats                                                           */

static const char MAGIC0[] = "COBS:";
static const char MAGIC_CLASSIC[] = "CLASSIC_INDEX";
static const char MAGIC_COMPACT[] = "COMPACT_INDEX";

typedef struct {
    const uint8_t* p;
    size_t pos, size;
    int bad;
} rd_t;

static void rd_bytes(rd_t* r, void* dst, size_t n) {
    if (r->bad || r->pos + n > r->size) {
        r->bad = 1;
        return;
    }
    memcpy(dst, r->p + r->pos, n);
    r->pos += n;
}
static int rd_magic(rd_t* r, const char* m) { /* header.hpp:23-29 */
    char buf[32];
    size_t n = strlen(m);
    rd_bytes(r, buf, n);
    return !r->bad && memcmp(buf, m, n) == 0;
}
static char** rd_names(rd_t* r, uint32_t n) {
    /* std::getline per name (classic_index_header.cpp:44-47) */
    char** names = (char**)calloc(n ? n : 1, sizeof(char*));
    for (uint32_t i = 0; i < n; ++i) {
        size_t s = r->pos;
        while (s < r->size && r->p[s] != '\n') ++s;
        if (s >= r->size) {
            r->bad = 1;
            s = r->size;
        }
        size_t l = s - r->pos;
        names[i] = (char*)malloc(l + 1);
        memcpy(names[i], r->p + r->pos, l);
        names[i][l] = 0;
        r->pos = s < r->size ? s + 1 : s;
    }
    return names;
}

static uint8_t* slurp(const char* path, size_t* size) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc(sz > 0 ? (size_t)sz : 1);
    if (sz > 0 && fread(buf, 1, (size_t)sz, f) != (size_t)sz) {
        free(buf);
        fclose(f);
        return NULL;
    }
    fclose(f);
    *size = (size_t)sz;
    return buf;
}

int oracle_index_load(const char* path, oracle_index* out) {
    memset(out, 0, sizeof(*out));
    size_t size = 0;
    uint8_t* blob = slurp(path, &size);
    if (!blob) return ORACLE_ERR_IO;
    rd_t r = { blob, 0, size, 0 };
    uint32_t version = 0;
    if (!rd_magic(&r, MAGIC0)) goto bad;
    size_t after0 = r.pos;
    if (rd_magic(&r, MAGIC_CLASSIC)) {
        /* classic_index_header.cpp:38-50 */
        uint32_t n_docs = 0;
        uint64_t sig = 0, nh = 0;
        rd_bytes(&r, &version, 4);
        if (r.bad || version != 1) goto bad;
        rd_bytes(&r, &out->term_size, 4);
        rd_bytes(&r, &out->canonicalize, 1);
        rd_bytes(&r, &n_docs, 4);
        rd_bytes(&r, &sig, 8);
        rd_bytes(&r, &nh, 8);
        if (r.bad) goto bad;
        out->doc_names = rd_names(&r, n_docs);
        out->n_docs = n_docs;
        if (!rd_magic(&r, MAGIC_CLASSIC)) goto bad;
        out->kind = ORACLE_KIND_CLASSIC;
        out->num_hashes = nh;
        out->n_pages = 1;
        out->page_size = ((uint64_t)n_docs + 7) / 8;
        out->signature_sizes = (uint64_t*)malloc(8);
        out->signature_sizes[0] = sig;
        out->page_data = (uint8_t**)malloc(sizeof(uint8_t*));
        out->page_data[0] = blob + r.pos;
        if (r.pos + sig * out->page_size > size) goto bad;
    } else {
        r.pos = after0;
        r.bad = 0;
        if (!rd_magic(&r, MAGIC_COMPACT)) goto bad;
        /* compact_index_header.cpp:44-65 */
        uint32_t n_params = 0, n_docs = 0;
        rd_bytes(&r, &version, 4);
        if (r.bad || version != 1) goto bad;
        rd_bytes(&r, &out->term_size, 4);
        rd_bytes(&r, &out->canonicalize, 1);
        rd_bytes(&r, &n_params, 4);
        rd_bytes(&r, &n_docs, 4);
        rd_bytes(&r, &out->page_size, 8);
        if (r.bad || n_params == 0 || out->page_size == 0) goto bad;
        out->signature_sizes = (uint64_t*)malloc(8 * (size_t)n_params);
        for (uint32_t i = 0; i < n_params; ++i) {
            uint64_t nh = 0;
            rd_bytes(&r, &out->signature_sizes[i], 8);
            rd_bytes(&r, &nh, 8);
            /* compact_index/search_file.cpp:23-27: one num_hashes for all pages */
            if (i == 0) out->num_hashes = nh;
            else if (nh != out->num_hashes) goto bad;
        }
        if (r.bad) goto bad;
        out->doc_names = rd_names(&r, n_docs);
        out->n_docs = n_docs;
        /* padding so that the data starts page_size-aligned
         * (compact_index_header.cpp:20-22, 62-63) */
        size_t pad = (out->page_size -
                      ((r.pos + strlen(MAGIC_COMPACT)) % out->page_size)) %
                     out->page_size;
        r.pos += pad;
        if (!rd_magic(&r, MAGIC_COMPACT)) goto bad;
        out->kind = ORACLE_KIND_COMPACT;
        out->n_pages = n_params;
        out->page_data = (uint8_t**)malloc(sizeof(uint8_t*) * n_params);
        size_t pos = r.pos;
        for (uint32_t i = 0; i < n_params; ++i) {
            out->page_data[i] = blob + pos;
            pos += out->signature_sizes[i] * out->page_size;
        }
        if (pos > size) goto bad;
    }
    out->file_blob = blob;
    out->file_size = size;
    return ORACLE_OK;
bad:
    out->file_blob = blob;
    oracle_index_free(out);
    return ORACLE_ERR_BAD_FILE;
}

void oracle_index_free(oracle_index* idx) {
    if (idx->doc_names) {
        for (uint32_t i = 0; i < idx->n_docs; ++i) free(idx->doc_names[i]);
        free(idx->doc_names);
    }
    if (idx->owns_pages && idx->page_data)
        for (uint32_t i = 0; i < idx->n_pages; ++i) free(idx->page_data[i]);
    free(idx->page_data);
    free(idx->signature_sizes);
    free(idx->file_blob);
    memset(idx, 0, sizeof(*idx));
}

static void put_names(FILE* f, uint32_t n, const char* const* names) {
    for (uint32_t i = 0; i < n; ++i) {
        if (names) fprintf(f, "%s\n", names[i]);
        else fprintf(f, "doc_%06u\n", i);
    }
}

static FILE* begin_classic(const char* path, uint32_t term_size,
                           uint8_t canonicalize, uint32_t n_docs,
                           uint64_t signature_size, uint64_t num_hashes,
                           const char* const* names) {
    /* classic_index_header.cpp:26-36 */
    FILE* f = fopen(path, "wb");
    if (!f) return NULL;
    uint32_t version = 1;
    fwrite(MAGIC0, 1, 5, f);
    fwrite(MAGIC_CLASSIC, 1, 13, f);
    fwrite(&version, 4, 1, f);
    fwrite(&term_size, 4, 1, f);
    fwrite(&canonicalize, 1, 1, f);
    fwrite(&n_docs, 4, 1, f);
    fwrite(&signature_size, 8, 1, f);
    fwrite(&num_hashes, 8, 1, f);
    put_names(f, n_docs, names);
    fwrite(MAGIC_CLASSIC, 1, 13, f);
    return f;
}

static FILE* begin_compact(const char* path, uint32_t term_size,
                           uint8_t canonicalize, uint32_t n_docs,
                           uint64_t page_size, uint32_t n_pages,
                           const uint64_t* signature_sizes, uint64_t num_hashes,
                           const char* const* names) {
    /* compact_index_header.cpp:24-42 */
    FILE* f = fopen(path, "wb");
    if (!f) return NULL;
    uint32_t version = 1;
    fwrite(MAGIC0, 1, 5, f);
    fwrite(MAGIC_COMPACT, 1, 13, f);
    fwrite(&version, 4, 1, f);
    fwrite(&term_size, 4, 1, f);
    fwrite(&canonicalize, 1, 1, f);
    fwrite(&n_pages, 4, 1, f);
    fwrite(&n_docs, 4, 1, f);
    fwrite(&page_size, 8, 1, f);
    for (uint32_t i = 0; i < n_pages; ++i) {
        fwrite(&signature_sizes[i], 8, 1, f);
        fwrite(&num_hashes, 8, 1, f);
    }
    put_names(f, n_docs, names);
    long pos = ftell(f);
    size_t pad = (page_size - (((uint64_t)pos + 13) % page_size)) % page_size;
    for (size_t i = 0; i < pad; ++i) fputc(0, f);
    fwrite(MAGIC_COMPACT, 1, 13, f);
    return f;
}

int oracle_write_classic(const char* path, uint32_t term_size,
                         uint8_t canonicalize, uint32_t n_docs,
                         uint64_t signature_size, uint64_t num_hashes,
                         const char* const* names, const uint8_t* data) {
    FILE* f = begin_classic(path, term_size, canonicalize, n_docs,
                            signature_size, num_hashes, names);
    if (!f) return ORACLE_ERR_IO;
    uint64_t row_size = ((uint64_t)n_docs + 7) / 8;
    size_t n = signature_size * row_size;
    int ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    return ok ? ORACLE_OK : ORACLE_ERR_IO;
}

int oracle_write_compact(const char* path, uint32_t term_size,
                         uint8_t canonicalize, uint32_t n_docs,
                         uint64_t page_size, uint32_t n_pages,
                         const uint64_t* signature_sizes, uint64_t num_hashes,
                         const char* const* names, const uint8_t* data) {
    FILE* f = begin_compact(path, term_size, canonicalize, n_docs, page_size,
                            n_pages, signature_sizes, num_hashes, names);
    if (!f) return ORACLE_ERR_IO;
    size_t n = 0;
    for (uint32_t i = 0; i < n_pages; ++i) n += signature_sizes[i] * page_size;
    int ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    return ok ? ORACLE_OK : ORACLE_ERR_IO;
}

int oracle_index_procedural(oracle_index* out, int kind, uint32_t term_size,
                            uint8_t canonicalize, uint64_t num_hashes,
                            uint32_t n_docs, uint64_t page_size,
                            uint32_t n_pages, const uint64_t* signature_sizes,
                            uint64_t fill_seed) {
    memset(out, 0, sizeof(*out));
    if (n_pages == 0) return ORACLE_ERR_BAD_PARAM;
    if (kind == ORACLE_KIND_CLASSIC) {
        if (n_pages != 1) return ORACLE_ERR_BAD_PARAM;
        page_size = ((uint64_t)n_docs + 7) / 8;
    } else if ((uint64_t)n_docs > 8 * page_size * n_pages) {
        return ORACLE_ERR_BAD_PARAM;
    }
    out->kind = kind;
    out->term_size = term_size;
    out->canonicalize = canonicalize;
    out->num_hashes = num_hashes;
    out->n_docs = n_docs;
    out->n_pages = n_pages;
    out->page_size = page_size;
    out->signature_sizes = (uint64_t*)malloc(8 * (size_t)n_pages);
    memcpy(out->signature_sizes, signature_sizes, 8 * (size_t)n_pages);
    out->procedural = 1;
    out->fill_seed = fill_seed;
    return ORACLE_OK;
}

static void fill_row(const oracle_index* idx, uint32_t p, uint64_t row,
                     uint8_t* dst) {
    uint64_t words = idx->page_size / 8, tail = idx->page_size % 8;
    for (uint64_t w = 0; w < words; ++w) {
        uint64_t v = oracle_fill_word(idx->fill_seed, p, row, w);
        memcpy(dst + 8 * w, &v, 8); /* little-endian host */
    }
    if (tail) {
        uint64_t v = oracle_fill_word(idx->fill_seed, p, row, words);
        memcpy(dst + 8 * words, &v, tail);
    }
}

int oracle_index_materialize(oracle_index* idx) {
    if (!idx->procedural) return ORACLE_OK;
    idx->page_data = (uint8_t**)calloc(idx->n_pages, sizeof(uint8_t*));
    for (uint32_t p = 0; p < idx->n_pages; ++p) {
        size_t n = idx->signature_sizes[p] * idx->page_size;
        idx->page_data[p] = (uint8_t*)malloc(n ? n : 1);
        if (!idx->page_data[p]) return ORACLE_ERR_IO;
        for (uint64_t r = 0; r < idx->signature_sizes[p]; ++r)
            fill_row(idx, p, r, idx->page_data[p] + r * idx->page_size);
    }
    idx->procedural = 0;
    idx->owns_pages = 1;
    return ORACLE_OK;
}

static int stream_pages(FILE* f, const oracle_index* idx) {
    uint8_t* row = (uint8_t*)malloc(idx->page_size + 8);
    int ok = 1;
    for (uint32_t p = 0; p < idx->n_pages && ok; ++p)
        for (uint64_t r = 0; r < idx->signature_sizes[p] && ok; ++r) {
            if (idx->procedural) fill_row(idx, p, r, row);
            else memcpy(row, idx->page_data[p] + r * idx->page_size, idx->page_size);
            ok = fwrite(row, 1, idx->page_size, f) == idx->page_size;
        }
    free(row);
    return ok;
}

int oracle_write_classic_procedural(const char* path, const oracle_index* idx) {
    if (idx->kind != ORACLE_KIND_CLASSIC) return ORACLE_ERR_BAD_PARAM;
    FILE* f = begin_classic(path, idx->term_size, idx->canonicalize, idx->n_docs,
                            idx->signature_sizes[0], idx->num_hashes,
                            (const char* const*)idx->doc_names);
    if (!f) return ORACLE_ERR_IO;
    int ok = stream_pages(f, idx);
    fclose(f);
    return ok ? ORACLE_OK : ORACLE_ERR_IO;
}

int oracle_write_compact_procedural(const char* path, const oracle_index* idx) {
    if (idx->kind != ORACLE_KIND_COMPACT) return ORACLE_ERR_BAD_PARAM;
    FILE* f = begin_compact(path, idx->term_size, idx->canonicalize, idx->n_docs,
                            idx->page_size, idx->n_pages, idx->signature_sizes,
                            idx->num_hashes, (const char* const*)idx->doc_names);
    if (!f) return ORACLE_ERR_IO;
    int ok = stream_pages(f, idx);
    fclose(f);
    return ok ? ORACLE_OK : ORACLE_ERR_IO;
}
