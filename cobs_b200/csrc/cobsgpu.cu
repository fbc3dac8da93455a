// cobs_b200/csrc/cobsgpu.cu -- implementation of the C ABI declared in include/cobsgpu.h.
//
// Host orchestration of the three kernels (K1 hash.cuh, K2 score.cuh, K3 select.cuh) that
// replace cobs::ClassicSearch::search (cobs/query/classic_search.cpp:403-505) for one index,
// plus the loader that lays the signature matrix out in HBM.  No CPU compute path exists:
// without a usable CUDA device every entry point fails with COBSGPU_ERR_CUDA.
//
// HBM layout: every (local) page is a row-major matrix [signature_size][pitch] bytes, pitch =
// row bytes rounded up to 128 (16 for rows < 512 B) and zero padded, so each row slice a tile
// needs is one aligned contiguous run that a single cp.async.bulk can fetch.  Bit d of a row
// = byte d/8, bit d%8 (LSB first), exactly the reference's file layout
// (cobs/construction/classic_index.cpp:40-43).
#include "../../include/cobsgpu.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "construct.cuh"
#include "densesort.cuh"
#include "fill.cuh"
#include "hash.cuh"
#include "index_file.hpp"
#include "merge_runs.hpp"
#include "score.cuh"
#include "select.cuh"

using namespace cobsgpu;

namespace {

struct Err {
    int code;
    std::string msg;
};

thread_local std::string g_error;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            throw Err{ e_ == cudaErrorMemoryAllocation ? COBSGPU_ERR_OOM : COBSGPU_ERR_CUDA, \
                       std::string(#call) + ": " + cudaGetErrorString(e_) };              \
    } while (0)

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return COBSGPU_OK;
    }
    catch (const Err& e) {
        g_error = e.msg;
        return e.code;
    }
    catch (const std::bad_alloc&) {
        g_error = "host out of memory";
        return COBSGPU_ERR_OOM;
    }
    catch (const std::exception& e) {
        g_error = e.what();
        return COBSGPU_ERR_INVALID_ARG;
    }
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;              // owns its allocation
    DevBuf& operator=(const DevBuf&) = delete;
    void ensure(size_t n) {
        if (n <= cap) return;
        reserve(round_up<size_t>(n + n / 4, 256));
    }
    void reserve(size_t n) {   // exact capacity, contents are not kept
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        CK(cudaMalloc(&p, n));
        cap = n;
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;              // owns its allocation
    PinBuf& operator=(const PinBuf&) = delete;
    void ensure(size_t n) {
        if (n <= cap) return;
        reserve(round_up<size_t>(n + n / 4, 256));
    }
    void reserve(size_t n) {   // exact capacity, contents are not kept
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        CK(cudaMallocHost(&p, n));
        cap = n;
    }
    // grows the buffer, preserving its first `keep` bytes
    void ensure_keep(size_t n, size_t keep) {
        if (n <= cap) return;
        void* q = nullptr;
        n = round_up<size_t>(n + n / 4, 256);
        CK(cudaMallocHost(&q, n));
        if (p) {
            std::memcpy(q, p, std::min(keep, cap));
            cudaFreeHost(p);
        }
        p = q;
        cap = n;
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
};

// Result arrays (documents / scores) can run to hundreds of megabytes per call when a query
// returns every document: a grow-only buffer that never value-initialises (std::vector::resize
// would zero-fill what the decode overwrites a moment later) with the few vector operations the
// result assembly uses.
struct U32Buf {
    uint32_t* p = nullptr;
    size_t n = 0, cap = 0;
    bool owned = true;   // false: a view of pinned memory that a slot owns (see alias())
    U32Buf() = default;
    U32Buf(const U32Buf&) = delete;
    U32Buf(U32Buf&& o) noexcept : p(o.p), n(o.n), cap(o.cap), owned(o.owned) {
        o.p = nullptr;
        o.n = o.cap = 0;
        o.owned = true;
    }
    // view of `m` elements at `ptr` (not owned, never freed or grown in place)
    void alias(uint32_t* ptr, size_t m) {
        if (owned) std::free(p);
        p = ptr;
        n = m;
        cap = 0;
        owned = false;
    }
    U32Buf& operator=(const U32Buf& o) {
        if (this != &o) {
            resize(o.n);
            if (o.n) std::memcpy(p, o.p, o.n * 4);
        }
        return *this;
    }
    ~U32Buf() {
        if (owned) std::free(p);
    }
    void reserve(size_t m) {
        if (owned && m <= cap) return;
        const size_t want = owned ? std::max(m, cap + cap / 2) : std::max(m, n);
        if (want > (static_cast<size_t>(1) << 60)) throw std::bad_alloc();
        if (owned) {
            void* q = std::realloc(p, want * 4);
            if (!q) throw std::bad_alloc();
            p = static_cast<uint32_t*>(q);
        } else {   // leaving the view: own a copy of what it showed
            void* q = std::malloc(std::max<size_t>(want, 1) * 4);
            if (!q) throw std::bad_alloc();
            if (n) std::memcpy(q, p, n * 4);
            p = static_cast<uint32_t*>(q);
            owned = true;
        }
        cap = want;
    }
    void resize(size_t m) {   // new elements are NOT initialised
        reserve(m);
        n = m;
    }
    void clear() {
        if (!owned) {
            p = nullptr;
            owned = true;
        }
        n = 0;
    }
    size_t size() const { return n; }
    uint32_t* data() { return p; }
    uint32_t* begin() { return p; }
    uint32_t* end() { return p + n; }
    const uint32_t* begin() const { return p; }
    // append [first, last); `where` must be end()
    void insert(const uint32_t*, const uint32_t* first, const uint32_t* last) {
        const size_t m = static_cast<size_t>(last - first);
        reserve(n + m);
        if (m) std::memcpy(p + n, first, m * 4);
        n += m;
    }
    void swap(U32Buf& o) {
        std::swap(p, o.p);
        std::swap(n, o.n);
        std::swap(cap, o.cap);
        std::swap(owned, o.owned);
    }
};

struct LocalPage {
    uint32_t global_page;
    uint64_t sig;
    uint32_t pitch;
    uint32_t row_bytes;     // source-row bytes held by this shard
    uint32_t rb16;          // row_bytes rounded up to 16 (what the tiles cover)
    uint64_t byte_begin;    // first source-row byte held
    uint32_t doc_base;      // global document id of bit 0
    uint32_t n_real;        // real documents among the 8*row_bytes columns held
    uint32_t dense_off;     // first column in the shard-local dense layout
    uint8_t* d_base;
};

// Everything one batch needs from upload to collected result.  The handle rotates through a
// small ring of these so that the upload + K1 of batch i+1 and the download of batch i-1 overlap
// the score kernel of batch i.
struct Slot {
    // host geometry of the batch
    std::vector<uint64_t> qoff;          // [nq+1] byte offsets inside the batch blob
    std::vector<uint32_t> koff, thr;     // [nq+1] k-mer prefix, [nq] ceil(threshold * T_q)
    uint32_t nq = 0, total_kmers = 0, uniform_T = 0, max_T = 0;
    std::vector<uint64_t> geo_offsets;   // the caller's offsets this geometry was derived from
    double geo_threshold = 0;
    bool geo_valid = false;
    bool meta_resident = false;          // d_meta holds this geometry: only re-arm the flags
    void* meta_ptr = nullptr;
    // device: d_meta = [flags 2 x int | bad nq x u32 | qoff | koff | thr] in one block,
    // d_out = [flags 2 x int | offsets (nq+1) x u64 | cand_count nq x u32 | keys ...]
    DevBuf d_queries, d_meta, d_hashes, d_cand, d_scratch, d_res_count, d_out, d_qlist, d_dense;
    DevBuf d_hist, d_total;              // counting sort of the exhaustive lists (aux slot only)
    DevBuf d_out_alt;                    // second result area of the pipelined exhaustive passes
    cudaEvent_t ev_pipe[4] = { nullptr, nullptr, nullptr, nullptr };   // [done 0/1, copied 0/1]
    size_t meta_qoff = 0, meta_koff = 0, meta_thr = 0, meta_bad = 0;
    size_t out_off = 0, out_cc = 0, out_keys = 0;
    const char* dev_queries = nullptr;
    uint32_t* small_cc = nullptr;        // small batches: zeroed candidate counters inside d_meta
    PinBuf h_meta, h_out;
    PinBuf h_res;                        // exhaustive results land here and are handed out as they are
    cudaEvent_t ev_meta = nullptr, ev_in = nullptr, ev_main = nullptr, ev_out = nullptr;
    // the submitted batch
    bool busy = false;
    uint64_t ticket = 0;
    uint32_t q0 = 0;                     // first query of the batch inside the caller's call
    double threshold = 0;
    uint64_t limit = 0;
    int mode = 0;                        // MODE_CAND / MODE_TOPK, or -1: exhaustive at collect
    bool lng = false;
    bool ksplit = false;                 // few long queries: k-split score kernel at collect
    uint32_t cap = 0;
    uint64_t spec_keys = 0;              // keys copied back speculatively with the header
    std::vector<uint32_t> main_ids;      // queries of the main pass when not all of them
    std::vector<uint32_t> huge_ids;      // queries of more than 65 535 k-mers
    // collected result (valid until the slot is submitted again)
    std::vector<uint64_t> r_off;
    U32Buf r_doc, r_score;

    int* d_flags() const { return d_meta.as<int>(); }
    uint32_t* d_bad() const { return reinterpret_cast<uint32_t*>(d_meta.as<char>() + meta_bad); }
    uint64_t* d_qoff() const { return reinterpret_cast<uint64_t*>(d_meta.as<char>() + meta_qoff); }
    uint32_t* d_koff() const { return reinterpret_cast<uint32_t*>(d_meta.as<char>() + meta_koff); }
    uint32_t* d_thr() const { return reinterpret_cast<uint32_t*>(d_meta.as<char>() + meta_thr); }
    // result header / keys inside d_out for n query slots
    void layout_out(uint32_t n) {
        out_off = 8;
        out_cc = out_off + (static_cast<size_t>(n) + 1) * 8;
        out_keys = round_up<size_t>(out_cc + static_cast<size_t>(n) * 4, 8);
    }
    int* o_flags() const { return d_out.as<int>(); }
    uint64_t* o_off() const { return reinterpret_cast<uint64_t*>(d_out.as<char>() + out_off); }
    uint32_t* o_cc() const { return reinterpret_cast<uint32_t*>(d_out.as<char>() + out_cc); }
    uint64_t* o_keys() const { return reinterpret_cast<uint64_t*>(d_out.as<char>() + out_keys); }
    cudaEvent_t& ev(cudaEvent_t& e) {
        if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
    // total capacity of the buffers: changes exactly when one of them grew
    size_t footprint() const {
        size_t n = h_meta.cap + h_out.cap;
        for (const DevBuf* b : { &d_queries, &d_meta, &d_hashes, &d_cand, &d_scratch, &d_res_count,
                                 &d_out, &d_qlist, &d_dense })
            n += b->cap;
        return n;
    }
    // grows this (idle) slot's buffers to the capacities of `o`: allocations synchronise the
    // device, so the whole ring is sized during the first batch instead of once per slot
    void match(const Slot& o) {
        DevBuf* mine[] = { &d_queries, &d_meta, &d_hashes, &d_cand, &d_scratch, &d_res_count,
                           &d_out, &d_qlist, &d_dense };
        const DevBuf* theirs[] = { &o.d_queries, &o.d_meta, &o.d_hashes, &o.d_cand, &o.d_scratch,
                                   &o.d_res_count, &o.d_out, &o.d_qlist, &o.d_dense };
        for (size_t i = 0; i < sizeof(mine) / sizeof(mine[0]); ++i) {
            if (mine[i]->cap < theirs[i]->cap) {
                if (mine[i] == &d_meta) meta_resident = false;
                mine[i]->reserve(theirs[i]->cap);
            }
        }
        h_meta.reserve(o.h_meta.cap);
        h_out.reserve(o.h_out.cap);
    }
    void release() {
        for (cudaEvent_t* e : { &ev_meta, &ev_in, &ev_main, &ev_out, &ev_pipe[0], &ev_pipe[1], &ev_pipe[2],
                                &ev_pipe[3] })
            if (*e) {
                cudaEventDestroy(*e);
                *e = nullptr;
            }
    }
};

enum Phase { PH_H2D = 0, PH_HASH, PH_SCORE, PH_SELECT, PH_D2H };

}  // namespace

struct cobsgpu_index {
    // description
    int kind = 0;
    uint32_t term_size = 0, canonicalize = 0, num_hashes = 0, n_docs = 0, n_pages_global = 0;
    uint64_t page_size_src = 0;   // bytes per row per page in the source layout
    int device = 0;
    uint32_t shard_index = 0, shard_count = 1;
    std::vector<std::string> doc_names;
    std::vector<uint64_t> signature_sizes;   // per GLOBAL page

    // HBM layout
    std::vector<LocalPage> pages;
    uint8_t* d_arena = nullptr;
    uint64_t hbm_bytes = 0;
    uint32_t ncw = 1;             // consumer warps per CTA -> tile width 512*ncw
    std::vector<TileDesc> tiles;
    TileDesc* d_tiles = nullptr;
    uint64_t dense_pitch = 128;
    uint32_t shard_real_docs = 0;
    uint32_t shard_doc_begin = 0, shard_doc_end = 0;
    uint64_t bytes_per_kmer = 0;
    uint32_t* d_seg = nullptr;    // [3][n_local_pages]: dense_off, n_real, doc_base
    // score kernel work counters: WORK_RING pairs {next item, CTAs finished}, zero between launches
    static constexpr uint32_t WORK_RING = 16;
    unsigned long long* d_work = nullptr;
    uint32_t work_slot = 0;

    // loader statistics of the last open (matrix bytes read from the source, wall seconds)
    double load_seconds = 0;
    uint64_t load_bytes = 0;
    uint32_t load_threads = 0;

    // device properties
    int sm_count = 0;
    size_t smem_optin = 0;

    // options
    uint32_t max_candidates = 1024;
    uint32_t max_batch = 16384;
    uint64_t workspace_bytes = 1024ull << 20;
    uint64_t pipe_bytes = 32ull << 20;    // result bytes per copy of the pipelined exhaustive passes
    uint64_t pinned_result_max = 2ull << 30;   // result bytes per call handed out from pinned memory
    bool timing = false;

    // execution state: three streams so that consecutive batches overlap -- s_in uploads the
    // queries + metadata and runs K1, `stream` (main) runs K2 + K3, s_out copies results back
    // (s_fix: follow-up copies of a collect, which must not queue behind later batches' downloads)
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr, s_fix = nullptr;
    uint64_t spec_hint = 0;               // keys the last collected batch returned
    static constexpr int N_SLOTS = 4;     // batches in flight (submit/collect tickets)
    Slot slots[N_SLOTS];
    Slot aux;                             // workspace of the exhaustive passes / cobsgpu_scores
    int next_slot = 0;
    uint64_t ticket_counter = 0;
    bool prefetch = false, inputs_ready = false, input_stream_set = false;
    cudaStream_t input_stream = nullptr;  // device path: where the caller uploads d_queries
    cudaEvent_t ev_caller = nullptr;      // device path: caller's stream -> s_in ordering
    // cached launch configuration of the score kernel per mode
    struct ScoreCfg {
        bool valid = false;
        uint32_t n_stages = 0;
        size_t smem = 0;
        int occupancy = 0;
    } score_cfg[SCORE_MODES][2];   // [mode][8 or 16 bit-planes]
    uint32_t warps_per_query = 0;         // consumer warps that see one query (TOPK candidate bound)

    // results of the last cobsgpu_search_batch call
    std::vector<uint64_t> r_off;
    U32Buf r_doc, r_score;

    // timers
    cobsgpu_timers tm{};
    struct Ev {
        cudaEvent_t a, b;
        int phase;
    };
    std::vector<Ev> pending;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t trace_base = nullptr;

    ~cobsgpu_index() {
        cudaSetDevice(device);
        for (cudaStream_t st : { s_in, stream, s_out, s_fix })
            if (st) cudaStreamSynchronize(st);
        for (auto& e : pending) {
            cudaEventDestroy(e.a);
            cudaEventDestroy(e.b);
        }
        for (auto e : ev_pool) cudaEventDestroy(e);
        for (Slot& sl : slots) sl.release();
        aux.release();
        if (ev_caller) cudaEventDestroy(ev_caller);
        if (d_arena) cudaFree(d_arena);
        if (d_tiles) cudaFree(d_tiles);
        if (d_seg) cudaFree(d_seg);
        if (d_work) cudaFree(d_work);
        for (cudaStream_t st : { s_in, stream, s_out, s_fix })
            if (st) cudaStreamDestroy(st);
    }
};

namespace {

// ---------------------------------------------------------------------------------------
// phase timing with CUDA events (only when the "timing" option is on)

struct PhaseScope {
    cobsgpu_index* ix;
    cudaEvent_t a = nullptr, b = nullptr;
    int phase;
    cudaStream_t st;
    PhaseScope(cobsgpu_index* ix_, int phase_, cudaStream_t st_) : ix(ix_), phase(phase_), st(st_) {
        if (!ix->timing) return;
        a = take();
        b = take();
        cudaEventRecord(a, st);
    }
    cudaEvent_t take() {
        if (!ix->ev_pool.empty()) {
            cudaEvent_t e = ix->ev_pool.back();
            ix->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    ~PhaseScope() {
        if (!ix->timing) return;
        cudaEventRecord(b, st);
        ix->pending.push_back({ a, b, phase });
    }
};

void resolve_timers(cobsgpu_index* ix, bool wait = true) {
    // experiments only: COBSGPU_TRACE=1 prints every phase's start/end on the device timeline
    static const bool trace = std::getenv("COBSGPU_TRACE") != nullptr;
    static const char* names[] = { "h2d", "hash", "score", "select", "d2h" };
    std::vector<cobsgpu_index::Ev> later;
    for (auto& e : ix->pending) {
        float ms = 0;
        // (a collect in the middle of a pipeline must not wait for the batches behind it)
        if (!wait && cudaEventQuery(e.b) != cudaSuccess) {
            cudaGetLastError();
            later.push_back(e);
            continue;
        }
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            if (trace) {
                if (!ix->trace_base) {
                    ix->trace_base = e.a;   // (kept out of the pool below)
                }
                float t0 = 0, t1 = 0;
                cudaEventElapsedTime(&t0, ix->trace_base, e.a);
                cudaEventElapsedTime(&t1, ix->trace_base, e.b);
                std::fprintf(stderr, "TRACE %-6s %10.3f %10.3f us\n", names[e.phase], 1e3 * t0, 1e3 * t1);
            }
            switch (e.phase) {
            case PH_H2D: ix->tm.h2d_ms += ms; break;
            case PH_HASH: ix->tm.hashes_ms += ms; break;
            case PH_SCORE: ix->tm.score_ms += ms; break;
            case PH_SELECT: ix->tm.select_ms += ms; break;
            case PH_D2H: ix->tm.d2h_ms += ms; break;
            }
        }
        if (e.a != ix->trace_base) ix->ev_pool.push_back(e.a);
        ix->ev_pool.push_back(e.b);
    }
    ix->pending.swap(later);
}

// ---------------------------------------------------------------------------------------
// index layout

void check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Err{ COBSGPU_ERR_CUDA,
                   std::string("no CUDA device available (libcobsgpu has no CPU fallback): ") +
                       cudaGetErrorString(e) };
    if (device < 0 || device >= n)
        throw Err{ COBSGPU_ERR_INVALID_ARG, "device ordinal out of range" };
    CK(cudaSetDevice(device));
}

void build_layout(cobsgpu_index* ix, const std::vector<uint64_t>& sig) {
    const uint32_t P = ix->n_pages_global;
    const uint64_t ps = ix->page_size_src;
    const uint32_t S = ix->shard_count, g = ix->shard_index;
    ix->pages.clear();
    if (ix->kind == COBSGPU_KIND_CLASSIC) {
        // contiguous column ranges cut at multiples of 128 documents (16 bytes)
        const uint64_t gran = div_ceil<uint64_t>(ps, 16);
        const uint64_t lo = gran * g / S, hi = gran * (g + 1) / S;
        const uint64_t b0 = lo * 16, b1 = std::min<uint64_t>(hi * 16, ps);
        if (b1 > b0) {
            LocalPage lp{};
            lp.global_page = 0;
            lp.sig = sig[0];
            lp.byte_begin = b0;
            lp.row_bytes = static_cast<uint32_t>(b1 - b0);
            lp.doc_base = static_cast<uint32_t>(8 * b0);
            ix->pages.push_back(lp);
        }
    } else if (div_ceil<uint64_t>(ps, 16) >= 4ull * S) {
        // Compact, pages wide enough: every shard holds the SAME column range of EVERY page (cut
        // at multiples of 128 documents, at least 64 bytes per page and shard).  Work per k-mer
        // (h * bytes) and memory are then equal across shards to within one granule -- dealing
        // out whole pages (below) left 77 pages over 8 shards at 9 or 10 pages, and the step
        // waits for the shards with 10.
        const uint64_t gran = div_ceil<uint64_t>(ps, 16);
        const uint64_t lo = gran * g / S, hi = gran * (g + 1) / S;
        const uint64_t b0 = lo * 16, b1 = std::min<uint64_t>(hi * 16, ps);
        for (uint32_t p = 0; p < P && b1 > b0; ++p) {
            LocalPage lp{};
            lp.global_page = p;
            lp.sig = sig[p];
            lp.byte_begin = b0;
            lp.row_bytes = static_cast<uint32_t>(b1 - b0);
            lp.doc_base = static_cast<uint32_t>(static_cast<uint64_t>(p) * 8 * ps + 8 * b0);
            ix->pages.push_back(lp);
        }
    } else {
        // Narrow pages: whole pages per shard.  Every page costs the same per query k-mer
        std::vector<uint32_t> order(P);
        for (uint32_t p = 0; p < P; ++p) order[p] = p;
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return sig[a] > sig[b]; });
        std::vector<uint32_t> mine;
        for (uint32_t i = 0; i < P; ++i) {
            const uint32_t round = i / S, pos = i % S;
            const uint32_t owner = (round & 1) ? S - 1 - pos : pos;
            if (owner == g) mine.push_back(order[i]);
        }
        std::sort(mine.begin(), mine.end());
        for (uint32_t p : mine) {
            LocalPage lp{};
            lp.global_page = p;
            lp.sig = sig[p];
            lp.byte_begin = 0;
            lp.row_bytes = static_cast<uint32_t>(ps);
            lp.doc_base = static_cast<uint32_t>(static_cast<uint64_t>(p) * 8 * ps);
            ix->pages.push_back(lp);
        }
    }
    uint32_t max_rb16 = 16;
    uint64_t dense = 0, arena = 0;
    ix->shard_real_docs = 0;
    ix->bytes_per_kmer = 0;
    ix->shard_doc_begin = ix->pages.empty() ? 0 : ix->pages.front().doc_base;
    ix->shard_doc_end = ix->shard_doc_begin;
    for (auto& lp : ix->pages) {
        lp.rb16 = round_up<uint32_t>(lp.row_bytes, 16);
        lp.pitch = lp.row_bytes >= 512 ? round_up<uint32_t>(lp.row_bytes, 128) : lp.rb16;
        const uint64_t cols = static_cast<uint64_t>(lp.row_bytes) * 8;
        lp.n_real = ix->n_docs > lp.doc_base
                        ? static_cast<uint32_t>(std::min<uint64_t>(ix->n_docs - lp.doc_base, cols))
                        : 0;
        if (dense + static_cast<uint64_t>(lp.rb16) * 8 > 0xFFFFFF00ull)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "shard holds more than 2^32 columns" };
        lp.dense_off = static_cast<uint32_t>(dense);
        dense += static_cast<uint64_t>(lp.rb16) * 8;
        arena += round_up<uint64_t>(lp.sig * lp.pitch, 256);
        max_rb16 = std::max(max_rb16, lp.rb16);
        ix->shard_real_docs += lp.n_real;
        ix->bytes_per_kmer += static_cast<uint64_t>(ix->num_hashes) * lp.row_bytes;
        ix->shard_doc_end = std::max<uint32_t>(ix->shard_doc_end,
                                               lp.doc_base + static_cast<uint32_t>(cols));
    }
    ix->dense_pitch = std::max<uint64_t>(128, round_up<uint64_t>(dense, 128));
    ix->ncw = std::min<uint32_t>(4, std::max<uint32_t>(1, div_ceil<uint32_t>(max_rb16, 512)));
    ix->hbm_bytes = arena;
}

void build_tiles(cobsgpu_index* ix) {
    const uint32_t W = ix->ncw * 512;
    ix->tiles.clear();
    for (auto& lp : ix->pages) {
        const uint32_t n = div_ceil<uint32_t>(lp.rb16, W);
        // Even split: an item costs about the same whatever its width (its k-mers are fetched
        // one row slice at a time), so a ragged last tile wastes its share of the time -- a
        // 15.6 KB shard row cut into 7 x 2048 + 1296 bytes ran 4.5 % below 8 x ~1950.  Widths are
        // multiples of 64 bytes (whole DRAM sectors) when that still fits in W, else of 16.
        uint32_t tb = round_up<uint32_t>(div_ceil<uint32_t>(lp.rb16, n), 64);
        if (tb > W) tb = round_up<uint32_t>(div_ceil<uint32_t>(lp.rb16, n), 16);
        for (uint32_t off = 0; off < lp.rb16; off += tb) {
            TileDesc t{};
            t.base = lp.d_base + off;
            t.sig = lp.sig;
            t.pitch = lp.pitch;
            t.bytes = std::min<uint32_t>(tb, lp.rb16 - off);
            t.doc_base = lp.doc_base + off * 8;
            const uint64_t cols = static_cast<uint64_t>(t.bytes) * 8;
            t.n_real = ix->n_docs > t.doc_base
                           ? static_cast<uint32_t>(std::min<uint64_t>(ix->n_docs - t.doc_base, cols))
                           : 0;
            // columns beyond the real row bytes (16-byte padding) are never real
            const uint64_t held = static_cast<uint64_t>(lp.row_bytes) * 8;
            const uint64_t first = static_cast<uint64_t>(off) * 8;
            const uint64_t in_row = held > first ? std::min<uint64_t>(held - first, cols) : 0;
            t.n_real = static_cast<uint32_t>(std::min<uint64_t>(t.n_real, in_row));
            t.dense_off = lp.dense_off + off * 8;
            ix->tiles.push_back(t);
        }
    }
    ix->warps_per_query = 0;
    for (const TileDesc& t : ix->tiles) ix->warps_per_query += div_ceil<uint32_t>(t.bytes, 512);
    if (!ix->tiles.empty()) {
        CK(cudaMalloc(&ix->d_tiles, ix->tiles.size() * sizeof(TileDesc)));
        CK(cudaMemcpy(ix->d_tiles, ix->tiles.data(), ix->tiles.size() * sizeof(TileDesc),
                      cudaMemcpyHostToDevice));
    }
    const size_t np = ix->pages.size();
    if (np) {
        std::vector<uint32_t> seg(3 * np);
        for (size_t i = 0; i < np; ++i) {
            seg[i] = ix->pages[i].dense_off;
            seg[np + i] = ix->pages[i].n_real;
            seg[2 * np + i] = ix->pages[i].doc_base;
        }
        CK(cudaMalloc(&ix->d_seg, seg.size() * 4));
        CK(cudaMemcpy(ix->d_seg, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice));
    }
}

// Fills dst with source rows [row0, row0 + nrows) of GLOBAL page `page`, full source rows of
// page_size_src bytes each (memcpy from host pointers, or pread from the index file).  Called
// concurrently from several loader threads.
using RowReader = std::function<void(uint32_t page, uint64_t row0, uint64_t nrows, uint8_t* dst)>;

// Streams the pages of this shard into HBM: the analogue of the reference's --load-complete read
// loop (cobs/util/query.cpp:56-86), with HBM as the destination.  The matrix is cut into chunks
// of whole rows; LOADER_THREADS host threads each own two pinned staging buffers and a stream and
// take chunks round-robin: pread (page-cache copy, the slow part: a few GB/s per thread) into one
// buffer while the DMA engine re-pitches the previous chunk out of the other (source row stride
// page_size -> device pitch, only this shard's column bytes).  Threads never wait for each other,
// so the copies of all of them keep the PCIe link busy.  Only the padding columns are zeroed.
// Pinned staging buffers outlive a load: page-locking memory costs ~0.7 s per GB, far more than
// copying through it, so the buffers are kept for the next index (or the next shard) to load.
struct StagePool {
    std::mutex m;
    std::vector<PinBuf*> idle;
    PinBuf* take(size_t bytes) {
        {
            std::lock_guard<std::mutex> lock(m);
            for (size_t i = 0; i < idle.size(); ++i)
                if (idle[i]->cap >= bytes) {
                    PinBuf* b = idle[i];
                    idle.erase(idle.begin() + i);
                    return b;
                }
        }
        std::unique_ptr<PinBuf> b(new PinBuf);
        b->reserve(bytes);
        return b.release();
    }
    void give(PinBuf* b) {
        std::lock_guard<std::mutex> lock(m);
        if (idle.size() < 64) idle.push_back(b);
        else delete b;
    }
};
StagePool g_stage_pool;

struct LoaderStats {
    double seconds = 0;
    uint64_t bytes = 0;
    uint32_t threads = 0;
};

void stream_pages(cobsgpu_index* ix, const RowReader& read, LoaderStats* stats) {
    const uint64_t ps = ix->page_size_src;
    struct Chunk {
        const LocalPage* lp;
        uint64_t row0, nrows;
    };
    std::vector<Chunk> chunks;
    uint64_t target = 4ull << 20;
    if (const char* e = std::getenv("COBSGPU_LOADER_CHUNK_MB"))
        target = static_cast<uint64_t>(std::max(1, std::atoi(e))) << 20;
    const uint64_t chunk_rows = std::max<uint64_t>(1, target / std::max<uint64_t>(1, ps));
    uint64_t total = 0;
    for (const LocalPage& lp : ix->pages) {
        if (lp.pitch > lp.row_bytes)   // zero padding columns (everything else is overwritten)
            CK(cudaMemset2DAsync(lp.d_base + lp.row_bytes, lp.pitch, 0, lp.pitch - lp.row_bytes, lp.sig,
                                 ix->stream));
        for (uint64_t r = 0; r < lp.sig; r += chunk_rows)
            chunks.push_back({ &lp, r, std::min<uint64_t>(chunk_rows, lp.sig - r) });
        total += lp.sig * ps;
    }
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    // measured on a 16-vCPU host: 8 threads read ~45 GB/s out of the page cache, more only fight
    // over memory bandwidth with the DMA engine
    uint32_t nthreads = static_cast<uint32_t>(std::min<size_t>(std::max(4u, std::min(hw / 2, 16u)), chunks.size()));
    if (const char* e = std::getenv("COBSGPU_LOADER_THREADS"))
        nthreads = static_cast<uint32_t>(std::max(1, std::min(64, std::atoi(e))));
    nthreads = std::max<uint32_t>(1, std::min<uint32_t>(nthreads, static_cast<uint32_t>(std::max<size_t>(chunks.size(), 1))));
    const size_t buf_bytes = static_cast<size_t>(chunk_rows * ps);
    // experiments only: COBSGPU_LOADER_MODE=read skips the DMA, =dma skips the reads
    int dbg_mode = 0;
    if (const char* e = std::getenv("COBSGPU_LOADER_MODE"))
        dbg_mode = std::strcmp(e, "read") == 0 ? 1 : (std::strcmp(e, "dma") == 0 ? 2 : 0);
    std::atomic<size_t> next{ 0 };
    std::vector<Err> errors(nthreads, Err{ COBSGPU_OK, "" });
    const auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](uint32_t t) {
        try {
            CK(cudaSetDevice(ix->device));
            PinBuf* stage[2] = { nullptr, nullptr };
            cudaEvent_t ev[2] = { nullptr, nullptr };
            cudaStream_t st = nullptr;
            CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            struct Guard {
                cudaEvent_t (&ev)[2];
                cudaStream_t& st;
                PinBuf* (&stage)[2];
                ~Guard() {
                    if (st) {
                        cudaStreamSynchronize(st);
                        cudaStreamDestroy(st);
                    }
                    for (cudaEvent_t e : ev)
                        if (e) cudaEventDestroy(e);
                    for (PinBuf* b : stage)
                        if (b) g_stage_pool.give(b);
                }
            } guard{ ev, st, stage };
            int slot = 0;
            for (;;) {
                const size_t c = next.fetch_add(1);
                if (c >= chunks.size()) break;
                const Chunk& ch = chunks[c];
                const LocalPage& lp = *ch.lp;
                if (!stage[slot]) stage[slot] = g_stage_pool.take(buf_bytes);
                if (!ev[slot]) CK(cudaEventCreateWithFlags(&ev[slot], cudaEventDisableTiming));
                else CK(cudaEventSynchronize(ev[slot]));   // DMA out of this buffer has finished
                uint8_t* buf = stage[slot]->as<uint8_t>();
                if (dbg_mode != 2) read(lp.global_page, ch.row0, ch.nrows, buf);
                if (dbg_mode != 1)
                    CK(cudaMemcpy2DAsync(lp.d_base + ch.row0 * lp.pitch, lp.pitch, buf + lp.byte_begin, ps,
                                         lp.row_bytes, ch.nrows, cudaMemcpyHostToDevice, st));
                CK(cudaEventRecord(ev[slot], st));
                slot ^= 1;
            }
            CK(cudaStreamSynchronize(st));
        } catch (const Err& e) {
            errors[t] = e;
        }
    };
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    CK(cudaSetDevice(ix->device));
    for (const Err& e : errors)
        if (e.code != COBSGPU_OK) throw e;
    CK(cudaStreamSynchronize(ix->stream));
    if (stats) {
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->bytes = total;
        stats->threads = nthreads;
    }
}

void open_common(cobsgpu_index* ix, const std::vector<uint64_t>& sig,
                 const RowReader* reader, uint64_t fill_seed, bool zero_fill = false) {
    check_device(ix->device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ix->device));
    if (prop.major < 10)   // the library is built for sm_100a only (Makefile)
        throw Err{ COBSGPU_ERR_CUDA, "unsupported device: libcobsgpu is built for sm_100 (B200) only" };
    ix->sm_count = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    if (ix->num_hashes == 0 || ix->num_hashes > 32)
        throw Err{ COBSGPU_ERR_INVALID_ARG, "num_hashes must be in 1..32" };
    if (ix->term_size == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "term_size must be > 0" };
    if (ix->canonicalize > 1)
        throw Err{ COBSGPU_ERR_INVALID_ARG,
                   "Unknown canonicalize value " + std::to_string(ix->canonicalize) };
    if (ix->shard_count == 0 || ix->shard_index >= ix->shard_count)
        throw Err{ COBSGPU_ERR_INVALID_ARG, "bad shard spec" };
    for (uint64_t s : sig)
        if (s == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "signature_size must be > 0" };

    ix->signature_sizes = sig;
    build_layout(ix, sig);
    {
        // uploads + K1 and the result copies get the higher priority: they are short and the
        // next score kernel waits for them
        int lo_p = 0, hi_p = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        CK(cudaStreamCreateWithPriority(&ix->stream, cudaStreamNonBlocking, lo_p));
        CK(cudaStreamCreateWithPriority(&ix->s_in, cudaStreamNonBlocking, hi_p));
        CK(cudaStreamCreateWithPriority(&ix->s_out, cudaStreamNonBlocking, hi_p));
        CK(cudaStreamCreateWithPriority(&ix->s_fix, cudaStreamNonBlocking, hi_p));
    }
    if (ix->hbm_bytes) {
        CK(cudaMalloc(reinterpret_cast<void**>(&ix->d_arena), ix->hbm_bytes));
        uint64_t off = 0;
        for (auto& lp : ix->pages) {
            lp.d_base = ix->d_arena + off;
            off += round_up<uint64_t>(lp.sig * lp.pitch, 256);
        }
        if (reader) {
            LoaderStats ls;
            stream_pages(ix, *reader, &ls);
            ix->load_seconds = ls.seconds;
            ix->load_bytes = ls.bytes;
            ix->load_threads = ls.threads;
        }
        for (auto& lp : ix->pages) {
            if (reader) {
                continue;
            } else if (zero_fill) {
                CK(cudaMemsetAsync(lp.d_base, 0, lp.sig * lp.pitch, ix->stream));
            } else {
                FillParams fp{ lp.d_base, lp.sig, lp.pitch, lp.row_bytes, lp.byte_begin, fill_seed,
                               lp.global_page };
                const uint32_t grid = static_cast<uint32_t>(
                    std::min<uint64_t>(lp.sig, static_cast<uint64_t>(ix->sm_count) * 16));
                fill_page_kernel<<<grid, 256, 0, ix->stream>>>(fp);
                CK(cudaGetLastError());
            }
        }
        CK(cudaStreamSynchronize(ix->stream));
    }
    build_tiles(ix);
    CK(cudaMalloc(reinterpret_cast<void**>(&ix->d_work), 16 * cobsgpu_index::WORK_RING));
    CK(cudaMemset(ix->d_work, 0, 16 * cobsgpu_index::WORK_RING));
}

// ---------------------------------------------------------------------------------------
// score kernel dispatch

using ScoreFn = void (*)(const ScoreParams);

template <int MODE, int NP>
ScoreFn pick_h(uint32_t h) {
    switch (h) {
    case 1: return score_kernel<1, MODE, NP>;
    case 2: return score_kernel<2, MODE, NP>;
    case 3: return score_kernel<3, MODE, NP>;
    case 4: return score_kernel<4, MODE, NP>;
    default: return score_kernel<0, MODE, NP>;
    }
}

// lng = 16 bit-planes (queries of up to 65 535 k-mers); the dense modes exist for 8 planes
// only (DENSE32 flushes into u32 scores before the planes overflow, whatever the query length)
ScoreFn pick_score(uint32_t h, int mode, bool lng) {
    switch (mode) {
    case MODE_CAND: return lng ? pick_h<MODE_CAND, 16>(h) : pick_h<MODE_CAND, 8>(h);
    case MODE_TOPK: return lng ? pick_h<MODE_TOPK, 16>(h) : pick_h<MODE_TOPK, 8>(h);
    case MODE_DENSE8: return pick_h<MODE_DENSE8, 8>(h);
    case MODE_DENSE16: return pick_h<MODE_DENSE16, 16>(h);
    case MODE_KSPLIT: return pick_h<MODE_KSPLIT, 8>(h);
    default: return pick_h<MODE_DENSE32, 8>(h);
    }
}

void launch_score(cobsgpu_index* ix, ScoreParams sp, int mode, bool lng, cudaStream_t st) {
    if (sp.nq_items == 0 || sp.n_tiles == 0) return;
    const uint32_t h = ix->num_hashes;
    ScoreFn fn = pick_score(h, mode, lng);
    const int threads = static_cast<int>((ix->ncw + 1) * 32);
    if (mode == MODE_DENSE8 || mode == MODE_DENSE32 || mode == MODE_KSPLIT) lng = false;
    if (mode == MODE_DENSE16) lng = true;
    cobsgpu_index::ScoreCfg& cfg = ix->score_cfg[mode][lng ? 1 : 0];
    if (!cfg.valid) {
        const uint32_t W = ix->ncw * 512;
        const uint32_t stage = h * W;
        // ring depth from the shared-memory budget: aim for 3 CTAs/SM, fall back to 2 or 1
        // when a stage (h row slices) is large
        const size_t avail = ix->smem_optin;   // 227 KB on B200
        uint32_t ns = 0;
        // tuning knobs for experiments: COBSGPU_OCC (CTAs per SM to aim for), COBSGPU_STAGES (cap)
        const char* e_occ = std::getenv("COBSGPU_OCC");
        const char* e_st = std::getenv("COBSGPU_STAGES");
        const int occ0 = e_occ ? std::max(1, std::min(4, std::atoi(e_occ))) : 3;
        for (int occ = occ0; occ >= 1; --occ) {
            const size_t budget = avail / occ - (occ > 1 ? 1024 : 0);
            if (budget <= 2048 + stage) continue;
            ns = static_cast<uint32_t>((budget - 2048) / stage);
            if (ns >= 4 || occ == 1) break;
        }
        if (ns == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "num_hashes too large for shared memory" };
        ns = std::min<uint32_t>(ns, 64);
        if (e_st) ns = std::max<uint32_t>(1, std::min<uint32_t>(ns, static_cast<uint32_t>(std::atoi(e_st))));
        cfg.n_stages = ns;
        cfg.smem = score_smem_header(ns) + static_cast<size_t>(ns) * stage;
        // The attribute belongs to the kernel function (per device), not to this handle: every
        // handle sets the same device-wide maximum, so handles with different tile widths (or
        // host threads racing on one device) can never lower each other's limit.
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(ix->smem_optin)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occupancy, fn, threads, cfg.smem));
        if (cfg.occupancy < 1) throw Err{ COBSGPU_ERR_CUDA, "score kernel does not fit on an SM" };
        cfg.valid = true;
    }
    sp.n_stages = cfg.n_stages;
    // work counters: one pair per launch out of a ring, so that score kernels of this handle
    // running concurrently on different streams never share a pair (the last CTA of a launch
    // re-arms its own pair)
    sp.work = ix->d_work + 2 * (ix->work_slot++ % cobsgpu_index::WORK_RING);
    if (mode != MODE_KSPLIT) sp.n_kchunks = 1;
    const uint64_t items = static_cast<uint64_t>(sp.nq_items) * sp.n_tiles * sp.n_kchunks;
    const uint64_t max_ctas = static_cast<uint64_t>(cfg.occupancy) * ix->sm_count;
    uint64_t grid64 = std::min<uint64_t>(items, max_ctas);
    // A few waves of long items (128 queries of 1000 k-mers on 7 tiles: 896 items for 444 CTAs):
    // the last, nearly empty wave would run a handful of CTAs alone, each latency-bound
    // (~0.18 us per k-mer) -- 0.18 of 0.84 ms there.  Fewer CTAs that all take the same number
    // of items finish together; two CTAs per SM still keep > 100 KB of row data in flight.
    static const bool no_balance = std::getenv("COBSGPU_NO_BALANCE") != nullptr;   // experiments only
    if (!no_balance && items > max_ctas) {
        const uint64_t waves = div_ceil<uint64_t>(items, max_ctas);
        if (waves <= 16) grid64 = div_ceil<uint64_t>(items, waves);
    }
    const uint32_t grid = static_cast<uint32_t>(grid64);
    PhaseScope ps(ix, PH_SCORE, st);
    fn<<<grid, threads, cfg.smem, st>>>(sp);
    CK(cudaGetLastError());
    ix->tm.kernel_launches++;
    ix->tm.score_launches++;
}

// ---------------------------------------------------------------------------------------
// batch preparation: host geometry, uploads, K1

// d_meta starts with two flags: [0] first query with an invalid base ([1] unused), FLAG_CLEAR
// when clear; the per-query `bad` words behind them are armed with the same byte pattern
static constexpr int FLAG_CLEAR = 0x7F7F7F7F;
static constexpr uint32_t MAX_T_SHORT = 255;      // 8 bit-planes
static constexpr uint32_t MAX_T_LONG = 65535;     // 16 bit-planes
static constexpr uint32_t TOPK_MAX_K = 1024;      // largest -l served by the per-warp top-k epilogue

// Fills slot `sl` with queries [q0, q1) of the caller's batch: host geometry, upload of the
// queries (unless they already live on the device) and of the metadata block, K1 -- all on `st`.
void prepare_batch(cobsgpu_index* ix, Slot& sl, const char* queries, bool dev_queries,
                   const uint64_t* offsets, uint32_t q0, uint32_t q1, double threshold,
                   cudaStream_t st, bool pack_small = false) {
    const uint32_t nq = q1 - q0;
    const uint32_t k = ix->term_size;
    if (std::isnan(threshold)) throw Err{ COBSGPU_ERR_INVALID_ARG, "threshold is NaN" };
    // Streaming workloads (fixed-length reads) present the same offsets batch after batch:
    // the derived geometry is cached per slot and, when the device copy is current, only the
    // flag words are re-armed on the device instead of a rebuild + upload.
    const bool same = sl.geo_valid && sl.geo_offsets.size() == static_cast<size_t>(nq) + 1 &&
                      sl.geo_threshold == threshold &&
                      std::memcmp(sl.geo_offsets.data(), offsets + q0, (static_cast<size_t>(nq) + 1) * 8) == 0;
    if (!same) {
        sl.geo_valid = false;   // stays invalid if one of the checks below throws
        sl.meta_resident = false;
        sl.nq = nq;
        sl.qoff.resize(nq + 1);
        sl.koff.resize(nq + 1);
        sl.thr.resize(nq);
        const uint64_t base0 = offsets[q0];
        uint64_t km = 0, last_T = ~0ull, max_T = 0;
        uint32_t uT = 0, last_thr = 0;
        bool uniform = true;
        for (uint32_t i = 0; i < nq; ++i) {
            if (offsets[q0 + i + 1] < offsets[q0 + i])
                throw Err{ COBSGPU_ERR_INVALID_ARG, "query offsets must be non-decreasing" };
            const uint64_t len = offsets[q0 + i + 1] - offsets[q0 + i];
            if (len < k)
                throw Err{ COBSGPU_ERR_QUERY_TOO_SHORT,
                           "query too short, needs to be at least " + std::to_string(k) +
                               " characters long (query " + std::to_string(q0 + i) + ")" };
            const uint64_t T = len - k + 1;
            if (T >= 0xFFFFFFFFull) throw Err{ COBSGPU_ERR_INVALID_ARG, "query too long" };
            sl.qoff[i] = offsets[q0 + i] - base0;
            sl.koff[i] = static_cast<uint32_t>(km);
            km += T;
            max_T = std::max(max_T, T);
            if (i == 0) uT = static_cast<uint32_t>(T);
            else if (T != uT) uniform = false;
            if (T != last_T) {
                // thresholds[i] = ceil(threshold * num_terms) in double (classic_search.cpp:444-449)
                const double th = std::ceil(threshold * static_cast<double>(T));
                last_thr = th <= 0.0 ? 0u : (th >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(th));
                last_T = T;
            }
            sl.thr[i] = last_thr;
        }
        if (km > 0x7FFFFFFFull)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "batch holds more than 2^31 k-mers" };
        sl.qoff[nq] = offsets[q1] - base0;
        sl.koff[nq] = static_cast<uint32_t>(km);
        sl.total_kmers = static_cast<uint32_t>(km);
        sl.uniform_T = uniform ? uT : 0;
        sl.max_T = static_cast<uint32_t>(max_T);
        sl.geo_offsets.assign(offsets + q0, offsets + q1 + 1);
        sl.geo_threshold = threshold;
        sl.geo_valid = true;
    }
    const uint64_t base = offsets[q0];
    const uint64_t kmers = sl.total_kmers;
    const uint64_t blob_bytes = sl.qoff[nq];

    sl.small_cc = nullptr;
    if (pack_small && !dev_queries) {
        // A handful of queries: ONE upload carries everything -- flags | bad | qoff | koff | thr |
        // zeroed candidate counters | the query bytes -- instead of a copy, a second copy and a
        // memset, each of which costs a few microseconds of pure latency.
        PhaseScope ps(ix, PH_H2D, st);
        const size_t nq1 = std::max<size_t>(nq, 1);
        sl.meta_bad = 8;
        sl.meta_qoff = round_up<size_t>(sl.meta_bad + nq1 * 4, 8);
        sl.meta_koff = sl.meta_qoff + (static_cast<size_t>(nq) + 1) * 8;
        sl.meta_thr = sl.meta_koff + round_up<size_t>((static_cast<size_t>(nq) + 1) * 4, 8);
        const size_t meta_cc = round_up<size_t>(sl.meta_thr + nq1 * 4, 8);
        const size_t meta_q = round_up<size_t>(meta_cc + nq1 * 4, 16);
        const size_t total = meta_q + blob_bytes;
        sl.d_meta.ensure(total + 16);
        if (sl.ev_meta) CK(cudaEventSynchronize(sl.ev_meta));   // previous upload out of h_meta done
        sl.h_meta.ensure(total);
        char* hm = sl.h_meta.as<char>();
        std::memset(hm, 0x7F, sl.meta_bad + nq1 * 4);
        std::memcpy(hm + sl.meta_qoff, sl.qoff.data(), (static_cast<size_t>(nq) + 1) * 8);
        std::memcpy(hm + sl.meta_koff, sl.koff.data(), (static_cast<size_t>(nq) + 1) * 4);
        if (nq) std::memcpy(hm + sl.meta_thr, sl.thr.data(), static_cast<size_t>(nq) * 4);
        std::memset(hm + meta_cc, 0, nq1 * 4);
        std::memcpy(hm + meta_q, queries + base, blob_bytes);
        CK(cudaMemcpyAsync(sl.d_meta.p, hm, total, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(sl.ev(sl.ev_meta), st));
        sl.meta_resident = false;   // (the block layout differs from the cached one)
        sl.dev_queries = sl.d_meta.as<char>() + meta_q;
        sl.small_cc = reinterpret_cast<uint32_t*>(sl.d_meta.as<char>() + meta_cc);
    } else {
        PhaseScope ps(ix, PH_H2D, st);
        if (dev_queries) {
            sl.dev_queries = queries + base;
        } else {
            sl.d_queries.ensure(blob_bytes + 16);
            CK(cudaMemcpyAsync(sl.d_queries.p, queries + base, blob_bytes, cudaMemcpyHostToDevice, st));
            sl.dev_queries = sl.d_queries.as<char>();
        }
        // one block: flags (2 x int) | bad (nq x u32) | qoff | koff | thr.  flags and bad are
        // "clear" when every byte is 0x7F, so one memset re-arms them.
        const size_t nq1 = std::max<size_t>(nq, 1);
        sl.meta_bad = 8;
        sl.meta_qoff = round_up<size_t>(sl.meta_bad + nq1 * 4, 8);
        sl.meta_koff = sl.meta_qoff + (static_cast<size_t>(nq) + 1) * 8;
        sl.meta_thr = sl.meta_koff + round_up<size_t>((static_cast<size_t>(nq) + 1) * 4, 8);
        const size_t meta_bytes = sl.meta_thr + nq1 * 4;
        const size_t arm_bytes = sl.meta_bad + nq1 * 4;
        sl.d_meta.ensure(meta_bytes);
        if (sl.meta_resident && sl.meta_ptr == sl.d_meta.p) {
            CK(cudaMemsetAsync(sl.d_meta.p, 0x7F, arm_bytes, st));
        } else {
            if (sl.ev_meta) CK(cudaEventSynchronize(sl.ev_meta));   // previous upload out of h_meta done
            sl.h_meta.ensure(meta_bytes);
            char* hm = sl.h_meta.as<char>();
            std::memset(hm, 0x7F, arm_bytes);
            std::memcpy(hm + sl.meta_qoff, sl.qoff.data(), (static_cast<size_t>(nq) + 1) * 8);
            std::memcpy(hm + sl.meta_koff, sl.koff.data(), (static_cast<size_t>(nq) + 1) * 4);
            if (nq) std::memcpy(hm + sl.meta_thr, sl.thr.data(), static_cast<size_t>(nq) * 4);
            CK(cudaMemcpyAsync(sl.d_meta.p, hm, meta_bytes, cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(sl.ev(sl.ev_meta), st));
            sl.meta_resident = true;
            sl.meta_ptr = sl.d_meta.p;
        }
    }
    sl.d_hashes.ensure(std::max<uint64_t>(1, kmers) * ix->num_hashes * 8);
    if (kmers) {
        HashParams hp{};
        hp.queries = sl.dev_queries;
        hp.qoff = sl.d_qoff();
        hp.koff = sl.d_koff();
        hp.nq = nq;
        hp.total_kmers = sl.total_kmers;
        hp.uniform_T = sl.uniform_T;
        hp.k = k;
        hp.h = ix->num_hashes;
        hp.canonicalize = ix->canonicalize;
        hp.hashes = sl.d_hashes.as<uint64_t>();
        hp.first_bad = sl.d_flags();
        hp.bad = sl.d_bad();
        PhaseScope ps(ix, PH_HASH, st);
        const uint32_t grid = div_ceil<uint32_t>(sl.total_kmers, 128);
        // k = 31 is what COBS indices use in practice: k-mer bytes held in registers
        if (k == 31) hash_kmers_kernel<31><<<grid, 128, 0, st>>>(hp);
        else hash_kmers_kernel<0><<<grid, 128, 0, st>>>(hp);
        CK(cudaGetLastError());
        ix->tm.kernel_launches++;
    }
    ix->tm.kmers += kmers;
    ix->tm.queries += nq;
}

// `src` holds the batch (hashes + metadata), the candidate buffers may belong to another slot
ScoreParams base_params(cobsgpu_index* ix, const Slot& src, const uint32_t* d_qlist, uint32_t n_slots) {
    ScoreParams sp{};
    sp.tiles = ix->d_tiles;
    sp.n_tiles = static_cast<uint32_t>(ix->tiles.size());
    sp.h = ix->num_hashes;
    sp.hashes = src.d_hashes.as<uint64_t>();
    sp.koff = src.d_koff();
    sp.qlist = d_qlist;
    sp.nq_items = n_slots;
    sp.thr = src.d_thr();
    sp.dense_pitch = ix->dense_pitch;
    sp.kchunk = 0;
    sp.n_kchunks = 1;
    return sp;
}

uint32_t bits_for(uint32_t v) {
    uint32_t b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return std::max<uint32_t>(b, 1);
}

uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// How one pass turns scores into candidate lists.
struct PassPlan {
    int mode = MODE_CAND;     // MODE_CAND / MODE_TOPK / MODE_DENSE32 (queries beyond 16 planes)
    bool lng = false;         // 16 bit-planes
    uint32_t cap = 1;         // candidate slots per query
    uint64_t limit = 0;       // results wanted per query (0 = all)
};

// K3 of one pass: (radix sort of the long lists) + finalize.  Candidates of slot i are at
// cand + i * cap with their number in cand_count[i]; the first min(n, limit) sorted keys end up
// at out_keys + i * stride, the count (or a COUNT_* flag) in out_counts[i].
void launch_select(cobsgpu_index* ix, const Slot& src, Slot& work, const uint32_t* d_qlist,
                   uint32_t n_slots, const PassPlan& pl, uint32_t max_T, uint32_t* cand_count,
                   uint64_t* out_keys, uint32_t* out_counts, uint32_t stride, bool report_bad,
                   cudaStream_t st, bool fuse_csr = false, char* host_out = nullptr) {
    PhaseScope ps(ix, PH_SELECT, st);
    const uint32_t cap = pl.cap;
    uint64_t* cand = work.d_cand.as<uint64_t>();
    bool large_in_scratch = false;
    // (limits of at most 32 are served by finalize_kernel's streaming selection, whatever the
    // length of the candidate list)
    if (cap > FIN_SORT_MAX && !(pl.limit >= 1 && pl.limit <= 32)) {
        work.d_scratch.ensure(static_cast<uint64_t>(n_slots) * cap * 8);
        SortLargeParams lp{};
        lp.cand = cand;
        lp.scratch = work.d_scratch.as<uint64_t>();
        lp.cand_count = cand_count;
        lp.cap = cap;
        lp.min_n = FIN_SORT_MAX;
        // digits over the varying key bits only: document id (low word) then ~score
        const uint32_t doc_bits = bits_for(ix->shard_doc_end ? ix->shard_doc_end - 1 : 0);
        const uint32_t score_bits = bits_for(max_T);
        uint32_t n = 0;
        for (uint32_t b = 0; b < doc_bits; b += 8) lp.digit_shift[n++] = b;
        for (uint32_t b = 0; b < score_bits; b += 8) lp.digit_shift[n++] = 32 + b;
        lp.n_pass = n;
        large_in_scratch = (n & 1) != 0;
        sort_large_kernel<<<n_slots, SORT_LARGE_THREADS, 0, st>>>(lp);
        CK(cudaGetLastError());
        ix->tm.kernel_launches++;
    }
    FinalizeParams fp{};
    fp.cand = cand;
    fp.scratch = work.d_scratch.as<uint64_t>();
    fp.cand_count = cand_count;
    fp.bad = report_bad ? src.d_bad() : nullptr;
    fp.qlist = d_qlist;
    fp.cap = cap;
    fp.nq = n_slots;
    fp.limit = pl.limit;
    fp.out_keys = out_keys;
    fp.out_counts = out_counts;
    fp.stride = stride;
    fp.fin_sort_max = std::min<uint32_t>(FIN_SORT_MAX, pow2_ceil(std::max<uint32_t>(cap, 32)));
    fp.large_in_scratch = large_in_scratch ? 1 : 0;
    if (fuse_csr) {   // small batch: offsets + keys + flags written by the same (single) CTA
        fp.csr_off = work.o_off();
        fp.csr_keys = work.o_keys();
        fp.flags_src = src.d_flags();
        fp.flags_dst = work.o_flags();
        if (host_out) {
            // ... straight into the (device-mapped) pinned result buffer: no copy afterwards
            fp.csr_off = reinterpret_cast<uint64_t*>(host_out + work.out_off);
            fp.csr_keys = reinterpret_cast<uint64_t*>(host_out + work.out_keys);
            fp.flags_dst = reinterpret_cast<int*>(host_out);
            fp.cc_dst = reinterpret_cast<uint32_t*>(host_out + work.out_cc);
        }
    }
    const size_t smem = static_cast<size_t>(fp.fin_sort_max) * 8;
    static std::atomic<bool> attr_set[64];   // (shards of a group collect on threads of their own)
    if (!attr_set[ix->device & 63].load()) {
        CK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(FIN_SORT_MAX * 8)));
        attr_set[ix->device & 63].store(true);
    }
    finalize_kernel<<<div_ceil<uint32_t>(n_slots, FIN_WARPS), FIN_WARPS * 32, smem, st>>>(fp);
    CK(cudaGetLastError());
    ix->tm.kernel_launches++;
}

// K2 of one pass over `n_slots` queries of the batch in `src` (slot i = batch query qlist[i],
// or i when d_qlist is null): fills work.d_cand / cand_count (zeroed here).
void launch_pass_score(cobsgpu_index* ix, const Slot& src, Slot& work, const uint32_t* d_qlist,
                       uint32_t n_slots, const PassPlan& pl, uint32_t* cand_count, cudaStream_t st,
                       bool counters_zeroed = false) {
    const uint32_t cap = pl.cap;
    work.d_cand.ensure(static_cast<uint64_t>(n_slots) * cap * 8);
    if (!counters_zeroed) CK(cudaMemsetAsync(cand_count, 0, static_cast<size_t>(n_slots) * 4, st));
    ScoreParams sp = base_params(ix, src, d_qlist, n_slots);
    sp.cand_count = cand_count;
    sp.cand = work.d_cand.as<uint64_t>();
    sp.cap = cap;
    sp.topk = static_cast<uint32_t>(std::min<uint64_t>(pl.limit, 0xFFFFFFFFull));
    if (pl.mode != MODE_DENSE32) {
        launch_score(ix, sp, pl.mode, pl.lng, st);
        return;
    }
    // queries beyond the 16 bit-planes: u32 scores flushed to global memory, then thresholded
    work.d_dense.ensure(static_cast<uint64_t>(n_slots) * ix->dense_pitch * 4);
    CK(cudaMemsetAsync(work.d_dense.p, 0, static_cast<uint64_t>(n_slots) * ix->dense_pitch * 4, st));
    sp.dense32 = work.d_dense.as<uint32_t>();
    launch_score(ix, sp, MODE_DENSE32, false, st);
    if (ix->pages.empty()) return;
    PhaseScope ps(ix, PH_SELECT, st);
    DenseToCandParams dp{};
    dp.dense32 = sp.dense32;
    dp.dense_pitch = ix->dense_pitch;
    dp.qlist = d_qlist;
    dp.thr = src.d_thr();
    const uint32_t np = static_cast<uint32_t>(ix->pages.size());
    dp.seg_dense_off = ix->d_seg;
    dp.seg_n_real = ix->d_seg + np;
    dp.seg_doc_base = ix->d_seg + 2 * np;
    dp.n_seg = np;
    dp.cand_count = sp.cand_count;
    dp.cand = sp.cand;
    dp.cap = cap;
    uint32_t max_real = 1;
    for (auto& lp : ix->pages) max_real = std::max(max_real, lp.n_real);
    dim3 grid(std::min<uint32_t>(div_ceil<uint32_t>(max_real, 256), 64), n_slots);
    dense_to_cand_kernel<<<grid, 256, 0, st>>>(dp);
    CK(cudaGetLastError());
    ix->tm.kernel_launches++;
}

// CSR formatting in `work.d_out` (layout_out(n_slots) done by the caller): offsets + keys
void launch_csr(cobsgpu_index* ix, const Slot& src, Slot& work, uint32_t n_slots, uint32_t cap,
                cudaStream_t st) {
    PhaseScope ps(ix, PH_SELECT, st);
    scan_offsets_kernel<<<1, 1024, 0, st>>>(work.d_res_count.as<uint32_t>(), n_slots, work.o_off());
    CK(cudaGetLastError());
    GatherKeysParams gp{ work.d_cand.as<uint64_t>(), work.d_res_count.as<uint32_t>(), work.o_off(),
                         cap, work.o_keys(), src.d_flags(), work.o_flags() };
    gather_keys_kernel<<<n_slots, 256, 0, st>>>(gp);
    CK(cudaGetLastError());
    ix->tm.kernel_launches += 2;
}

// bytes of d_out a pass over n_slots queries can need: header + every list at its longest
size_t out_bytes(Slot& work, uint32_t n_slots, const PassPlan& pl) {
    work.layout_out(n_slots);
    const uint64_t per = pl.limit ? std::min<uint64_t>(pl.limit, pl.cap) : pl.cap;
    return work.out_keys + static_cast<size_t>(n_slots) * per * 8;
}

struct HostList {
    std::vector<uint64_t> off;   // [n_slots + 1]
    U32Buf doc, score;
};

// fn(a, b) over [0, n) in a few host threads: result lists of large indices are memory-bound on
// the host, one thread does not saturate it
template <typename F>
void split_over_threads(uint64_t n, F&& fn) {
    const uint64_t per_thread = 1ull << 20;
    unsigned hw = std::thread::hardware_concurrency();
    const uint64_t nt = std::min<uint64_t>(std::max(1u, std::min(hw, 8u)), n / per_thread);
    if (nt <= 1) {
        fn(static_cast<uint64_t>(0), n);
        return;
    }
    std::vector<std::thread> th;
    for (uint64_t t = 0; t < nt; ++t) th.emplace_back(fn, n * t / nt, n * (t + 1) / nt);
    for (auto& t : th) t.join();
}

void decode_keys(const uint64_t* keys, uint64_t n, uint32_t* doc, uint32_t* score) {
    split_over_threads(n, [=](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            doc[i] = key_doc(keys[i]);
            score[i] = key_score(keys[i]);
        }
    });
}

void copy_u32(uint32_t* dst, const uint32_t* src, uint64_t n) {
    split_over_threads(n, [=](uint64_t a, uint64_t b) { std::memcpy(dst + a, src + a, (b - a) * 4); });
}

void throw_bad_base(uint32_t query) {
    throw Err{ COBSGPU_ERR_INVALID_BASE,
               "Invalid DNA base pair in query string. Only ACGT are allowed. (query " +
                   std::to_string(query) + ")" };
}

const uint32_t* upload_qlist(Slot& work, const std::vector<uint32_t>& ids, size_t begin, size_t n,
                             cudaStream_t st) {
    work.d_qlist.ensure(std::max<size_t>(n, 1) * 4);
    CK(cudaMemcpyAsync(work.d_qlist.p, ids.data() + begin, n * 4, cudaMemcpyHostToDevice, st));
    return work.d_qlist.as<uint32_t>();
}

// One sub-batch of the exhaustive path for queries of at most 65 535 k-mers: K2 stores one count
// per document (DENSE8 / DENSE16), then a stable multi-CTA counting sort on the score
// (densesort.cuh) writes the ordered keys straight into the CSR area of work.d_out.
// Returns the candidate slots per query when the k-split reduce emitted thresholded candidates
// (the header's cand_count words then tell whether a query overflowed them and stage 1 -- the
// counting sort over the dense vector that is still there -- has to follow), else 0.
uint32_t exhaustive_dense(cobsgpu_index* ix, const Slot& src, Slot& work, const uint32_t* d_ql,
                          uint32_t n, uint32_t max_T, uint64_t limit, bool ksplit, int stage,
                          cudaStream_t st) {
    // (the counting sort writes documents[] and scores[] as two u32 arrays -- what the C ABI
    // hands out -- where the candidate paths write 64-bit keys; see run_exhaustive)
    // k-split: the k-mers of every query are cut into chunks that become work items of their own
    // (partial counts added into the u16 vector), sized so that the items fill the GPU
    uint32_t kchunk = 0, n_kchunks = 1;
    if (ksplit) {
        const uint64_t per_chunk = std::max<uint64_t>(1, static_cast<uint64_t>(n) * ix->tiles.size());
        const uint64_t want = div_ceil<uint64_t>(2ull * 3 * ix->sm_count, per_chunk);   // ~2 waves of CTAs
        kchunk = round_up<uint32_t>(std::max<uint32_t>(8, static_cast<uint32_t>(div_ceil<uint64_t>(max_T, want))), 8);
        kchunk = std::min<uint32_t>(kchunk, 248);
        n_kchunks = div_ceil<uint32_t>(max_T, kchunk);
        if (n_kchunks <= 1) ksplit = false;
    }
    if (ksplit) {
        // one dense8 vector per (query, chunk), summed into the u16 vector afterwards
        work.d_scratch.ensure(static_cast<uint64_t>(n) * n_kchunks * ix->dense_pitch);
    }
    const bool two = ksplit || max_T > MAX_T_SHORT;
    const uint32_t cap = std::max<uint32_t>(ix->shard_real_docs, 1);
    const size_t esz = two ? 2 : 1;
    const uint32_t dense_cols = static_cast<uint32_t>(ix->dense_pitch);
    const uint32_t n_chunks = div_ceil<uint32_t>(std::max<uint32_t>(dense_cols, cap), DS_CHUNK);
    work.d_dense.ensure(static_cast<uint64_t>(n) * ix->dense_pitch * esz);
    work.d_hist.ensure(static_cast<uint64_t>(n) * n_chunks * 256 * 4);
    work.d_total.ensure(static_cast<size_t>(n) * 8);
    uint32_t* total1 = work.d_total.as<uint32_t>();
    uint32_t* total2 = total1 + n;
    const uint32_t np = static_cast<uint32_t>(ix->pages.size());
    ScoreParams sp = base_params(ix, src, d_ql, n);
    sp.dense8 = work.d_dense.as<uint8_t>();
    sp.dense16 = work.d_dense.as<uint16_t>();
    CK(cudaMemsetAsync(work.o_cc(), 0, static_cast<size_t>(n) * 4, st));
    if (stage == 0) {
        if (!ksplit)
            CK(cudaMemsetAsync(work.d_dense.p, 0, static_cast<uint64_t>(n) * ix->dense_pitch * esz, st));
        sp.kchunk = kchunk;
        sp.n_kchunks = n_kchunks;
        if (ksplit) sp.dense8 = work.d_scratch.as<uint8_t>();
        launch_score(ix, sp, ksplit ? MODE_KSPLIT : (two ? MODE_DENSE16 : MODE_DENSE8), two, st);
        sp.dense8 = work.d_dense.as<uint8_t>();
        if (ksplit && !ix->pages.empty()) {
            // threshold > 0: the reduce appends the documents that pass, like the CAND epilogue
            const bool emit = src.threshold > 0.0;
            const uint32_t ccap = std::min<uint32_t>(ix->max_candidates, cap);
            KsplitReduceParams rp{};
            rp.part8 = work.d_scratch.as<uint8_t>();
            rp.out16 = work.d_dense.as<uint16_t>();
            rp.dense_pitch = ix->dense_pitch;
            rp.koff = src.d_koff();
            rp.qlist = d_ql;
            rp.kchunk = kchunk;
            rp.n_kchunks = n_kchunks;
            if (emit) {
                work.d_cand.ensure(static_cast<uint64_t>(n) * ccap * 8);
                work.d_res_count.ensure(static_cast<size_t>(n) * 4);
                rp.thr = src.d_thr();
                rp.seg_dense_off = ix->d_seg;
                rp.seg_n_real = ix->d_seg + np;
                rp.seg_doc_base = ix->d_seg + 2 * np;
                rp.n_seg = np;
                rp.cand_count = work.o_cc();
                rp.cand = work.d_cand.as<uint64_t>();
                rp.cap = ccap;
            }
            const dim3 rgrid(div_ceil<uint32_t>(static_cast<uint32_t>(ix->dense_pitch), 256 * 16), n);
            ksplit_reduce_kernel<<<rgrid, 256, 0, st>>>(rp);
            CK(cudaGetLastError());
            ix->tm.kernel_launches++;
            if (emit) {
                PassPlan cp;
                cp.mode = MODE_CAND;
                cp.cap = ccap;
                cp.limit = limit;
                launch_select(ix, src, work, d_ql, n, cp, max_T, work.o_cc(), work.d_cand.as<uint64_t>(),
                              work.d_res_count.as<uint32_t>(), ccap, false, st);
                launch_csr(ix, src, work, n, ccap, st);
                return ccap;
            }
        }
    }

    PhaseScope ps(ix, PH_SELECT, st);
    DenseSortParams dp{};
    dp.dense8 = sp.dense8;
    dp.dense16 = sp.dense16;
    dp.dense_pitch = ix->dense_pitch;
    dp.dense_cols = dense_cols;
    dp.qlist = d_ql;
    dp.thr = src.d_thr();
    dp.seg_dense_off = ix->d_seg;
    dp.seg_n_real = ix->d_seg + np;
    dp.seg_doc_base = ix->d_seg + 2 * np;
    dp.n_seg = np;
    dp.cap = cap;
    dp.n_chunks = n_chunks;
    dp.hist = work.d_hist.as<uint32_t>();
    dp.limit = limit;
    dp.n_slots = n;
    dp.soa = 1;
    const dim3 grid(n_chunks, n);
    auto final_pass = [&](uint32_t* totals) {
        // slot totals -> CSR offsets -> scatter into the result area
        dp.slot_total = totals;
        ds_hist_kernel<<<grid, DS_THREADS, 0, st>>>(dp);
        ds_scan_kernel<<<n, 256 * DS_SCAN_GROUPS, 0, st>>>(dp);
        scan_offsets_kernel<<<1, 1024, 0, st>>>(totals, n, work.o_off());
        dp.keys_out = work.o_keys();
        dp.csr_off = work.o_off();
        ds_scatter_kernel<<<grid, DS_THREADS, 0, st>>>(dp);
        CK(cudaGetLastError());
        ix->tm.kernel_launches += 4;
    };
    if (ix->pages.empty()) {
        CK(cudaMemsetAsync(work.o_off(), 0, (static_cast<size_t>(n) + 1) * 8, st));
    } else if (!two) {
        dp.pass = 0;
        final_pass(total1);
    } else {
        work.d_cand.ensure(static_cast<uint64_t>(n) * cap * 8);
        dp.pass = 1;
        dp.slot_total = total1;
        dp.keys_out = work.d_cand.as<uint64_t>();
        ds_hist_kernel<<<grid, DS_THREADS, 0, st>>>(dp);
        ds_scan_kernel<<<n, 256 * DS_SCAN_GROUPS, 0, st>>>(dp);
        ds_scatter_kernel<<<grid, DS_THREADS, 0, st>>>(dp);
        CK(cudaGetLastError());
        ix->tm.kernel_launches += 3;
        dp.pass = 2;
        dp.keys_in = work.d_cand.as<uint64_t>();
        dp.in_count = total1;
        final_pass(total2);
    }
    CK(cudaMemcpyAsync(work.o_flags(), src.d_flags(), 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// Exhaustive lists whose lengths are known beforehand (threshold <= 0: every real document is a
// result, cut at `limit`): the result volume -- 8 bytes per document per query over PCIe -- is
// the bound, so the batch is cut into sub-batches of a few tens of megabytes and the copy of
// one overlaps K2 + the counting sort of the next (two result areas, copies on s_fix).  The
// doc[] | score[] arrays are assembled directly in the slot's pinned buffer and handed out.
// Returns false when the batch is not worth splitting (the caller's one-pass path serves it).
bool exhaustive_pipelined(cobsgpu_index* ix, const Slot& src, const std::vector<uint32_t>& ids,
                          uint64_t limit, size_t sub_max, std::vector<HostList>* lists,
                          std::vector<std::pair<uint32_t, uint32_t>>* where, bool ksplit,
                          PinBuf* res_pin) {
    static const bool off = std::getenv("COBSGPU_NO_PIPE") != nullptr;   // experiments only
    const uint64_t real = ix->shard_real_docs;
    const uint64_t per_q = limit ? std::min<uint64_t>(limit, real) : real;
    const size_t nq = ids.size();
    if (off || per_q == 0 || ix->pages.empty()) return false;
    // the arrays handed out are page-locked (~0.7 s per GB the first time a slot grows to a
    // size): beyond a few GB per call the general path's pageable arrays are the better deal
    if (static_cast<uint64_t>(ids.size()) * per_q * 8 > ix->pinned_result_max) return false;
    // Sub-batch size: every pass costs a few launches of fixed latency, and the first pass and the
    // last copy overlap nothing -- so at least "pipe_kb" (32 MB) of results per copy, and at most
    // about eight sub-batches.  A batch that fits one pass and is too short to split is left to
    // the caller's one-pass path.
    const size_t sub = std::min<size_t>(
        sub_max, std::max<uint64_t>(div_ceil<uint64_t>(nq, 8),
                                    std::max<uint64_t>(1, div_ceil<uint64_t>(ix->pipe_bytes, per_q * 8))));
    if (nq < 2 * sub && nq <= sub_max) return false;
    cudaStream_t st = ix->stream, sc = ix->s_fix;
    Slot& work = ix->aux;
    PassPlan pl;
    pl.cap = static_cast<uint32_t>(std::max<uint64_t>(real, 1));
    pl.limit = limit;
    pl.mode = MODE_DENSE32;
    const size_t ob = out_bytes(work, static_cast<uint32_t>(sub), pl);
    work.d_out.ensure(ob);
    work.d_out_alt.ensure(ob);
    const size_t n_sub = div_ceil<size_t>(nq, sub);
    const uint64_t total = static_cast<uint64_t>(nq) * per_q;
    res_pin->ensure(total * 8);
    work.h_out.ensure(std::max<size_t>(work.out_keys, n_sub * 8));
    uint32_t* res_doc = res_pin->as<uint32_t>();
    uint32_t* res_score = res_doc + total;
    uint64_t* h_tot = work.h_out.as<uint64_t>();
    for (size_t i = 0, b = 0; b < nq; ++i, b += sub) {
        const uint32_t n = static_cast<uint32_t>(std::min(sub, nq - b));
        const int buf = static_cast<int>(i & 1);
        // work.d_out is result area `buf` from here on
        std::swap(work.d_out.p, work.d_out_alt.p);
        std::swap(work.d_out.cap, work.d_out_alt.cap);
        if (i >= 2) CK(cudaStreamWaitEvent(st, work.ev_pipe[2 + buf], 0));   // its last copy has left
        const uint32_t* d_ql = upload_qlist(work, ids, b, n, st);
        uint32_t max_T = 1;
        for (size_t j = 0; j < n; ++j) max_T = std::max(max_T, src.koff[ids[b + j] + 1] - src.koff[ids[b + j]]);
        out_bytes(work, n, pl);   // (header layout for n queries)
        if (exhaustive_dense(ix, src, work, d_ql, n, max_T, limit, ksplit, 0, st) != 0)
            throw Err{ COBSGPU_ERR_CUDA, "exhaustive pass emitted candidates without a threshold" };
        CK(cudaEventRecord(work.ev(work.ev_pipe[buf]), st));
        CK(cudaStreamWaitEvent(sc, work.ev_pipe[buf], 0));
        {
            PhaseScope ps(ix, PH_D2H, sc);
            const uint64_t m = static_cast<uint64_t>(n) * per_q;
            const uint32_t* d_doc = reinterpret_cast<const uint32_t*>(work.o_keys());
            CK(cudaMemcpyAsync(res_doc + b * per_q, d_doc, m * 4, cudaMemcpyDeviceToHost, sc));
            CK(cudaMemcpyAsync(res_score + b * per_q, d_doc + m, m * 4, cudaMemcpyDeviceToHost, sc));
            CK(cudaMemcpyAsync(h_tot + i, work.o_off() + n, 8, cudaMemcpyDeviceToHost, sc));
        }
        CK(cudaEventRecord(work.ev(work.ev_pipe[2 + buf]), sc));
    }
    CK(cudaStreamSynchronize(sc));
    for (size_t i = 0, b = 0; b < nq; ++i, b += sub)
        if (h_tot[i] != std::min(sub, nq - b) * per_q)
            throw Err{ COBSGPU_ERR_CUDA, "exhaustive list length differs from the number of documents" };
    lists->emplace_back();
    HostList& L = lists->back();
    L.off.resize(nq + 1);
    for (size_t i = 0; i <= nq; ++i) L.off[i] = i * per_q;
    L.doc.alias(res_doc, total);
    L.score.alias(res_score, total);
    for (size_t i = 0; i < nq; ++i)
        (*where)[ids[i]] = { static_cast<uint32_t>(lists->size() - 1), static_cast<uint32_t>(i) };
    return true;
}

// Synchronous exhaustive pass over the given queries of the batch in `src`, in workspace-bounded
// sub-batches through the aux buffers: every real document of the shard can become a result
// (cap = shard_real_docs), so nothing is ever dropped.  Used for threshold <= 0 without a small
// limit, for queries whose candidates overflowed the fused path, and for queries beyond 16 planes.
void run_exhaustive(cobsgpu_index* ix, const Slot& src, const std::vector<uint32_t>& ids,
                    uint64_t limit, std::vector<HostList>* lists,
                    std::vector<std::pair<uint32_t, uint32_t>>* where, bool ksplit = false,
                    PinBuf* res_pin = nullptr) {
    if (ids.empty()) return;
    cudaStream_t st = ix->stream;
    Slot& work = ix->aux;
    // queries beyond 16 bit-planes keep the u32 score vector + candidate keys + radix sort
    std::vector<uint32_t> dense_ids, huge_ids;
    for (uint32_t q : ids)
        (src.koff[q + 1] - src.koff[q] > MAX_T_LONG ? huge_ids : dense_ids).push_back(q);
    PassPlan pl;
    pl.cap = std::max<uint32_t>(ix->shard_real_docs, 1);
    pl.limit = limit;
    const uint64_t out_per_q = (limit ? std::min<uint64_t>(limit, pl.cap) : pl.cap) * 8;
    // the whole batch, one kind of pass: the sub-batches append to ONE list in query order, which
    // collect_batch then takes over as it is (no second copy of what may be gigabytes)
    const bool join = res_pin != nullptr && (huge_ids.empty() || dense_ids.empty());
    size_t joined = static_cast<size_t>(-1);
    for (int huge = 0; huge < 2; ++huge) {
        const std::vector<uint32_t>& list = huge ? huge_ids : dense_ids;
        if (list.empty()) continue;
        uint64_t per_q;
        if (huge) {
            per_q = static_cast<uint64_t>(pl.cap) * 8 * (pl.cap > FIN_SORT_MAX ? 2 : 1) + out_per_q +
                    ix->dense_pitch * 4;
        } else {
            const uint64_t n_chunks = div_ceil<uint64_t>(std::max<uint64_t>(ix->dense_pitch, pl.cap), DS_CHUNK);
            per_q = ix->dense_pitch * 2 + n_chunks * 1024 + out_per_q + static_cast<uint64_t>(pl.cap) * 8;
            if (ksplit) per_q += ix->dense_pitch * 64;   // (chunk vectors; a k-split batch is a few queries)
        }
        const size_t sub = static_cast<size_t>(
            std::max<uint64_t>(1, std::min<uint64_t>(list.size(), ix->workspace_bytes / std::max<uint64_t>(per_q, 1))));
        if (!huge && res_pin != nullptr && src.threshold <= 0.0 && list.size() == ids.size() &&
            exhaustive_pipelined(ix, src, list, limit, sub, lists, where, ksplit, res_pin))
            continue;
        for (size_t b = 0; b < list.size(); b += sub) {
            const uint32_t n = static_cast<uint32_t>(std::min(sub, list.size() - b));
            const uint32_t* d_ql = upload_qlist(work, list, b, n, st);
            uint32_t max_T = 1;
            for (size_t i = 0; i < n; ++i)
                max_T = std::max(max_T, src.koff[list[b + i] + 1] - src.koff[list[b + i]]);
            pl.lng = max_T > MAX_T_SHORT;
            pl.mode = MODE_DENSE32;
            work.d_out.ensure(out_bytes(work, n, pl));
            uint32_t cand_cap = 0;
            bool soa = false;   // result area holds doc[] | score[] instead of keys
            if (!huge) {
                cand_cap = exhaustive_dense(ix, src, work, d_ql, n, max_T, limit, ksplit, 0, st);
                soa = cand_cap == 0;
            } else {
                work.d_res_count.ensure(static_cast<size_t>(n) * 4);
                launch_pass_score(ix, src, work, d_ql, n, pl, work.o_cc(), st);
                launch_select(ix, src, work, d_ql, n, pl, max_T, work.o_cc(), work.d_cand.as<uint64_t>(),
                              work.d_res_count.as<uint32_t>(), pl.cap, false, st);
                launch_csr(ix, src, work, n, pl.cap, st);
            }
            // header first (it tells how many keys there are), then the keys
            work.h_out.ensure(work.out_keys);
            {
                PhaseScope ps(ix, PH_D2H, st);
                CK(cudaMemcpyAsync(work.h_out.p, work.d_out.p, work.out_keys, cudaMemcpyDeviceToHost, st));
            }
            CK(cudaStreamSynchronize(st));
            if (cand_cap) {
                // k-split batch with a threshold: did a query overflow its candidate slots?  Then
                // the counting sort over the dense vector (still there) produces the lists.
                const uint32_t* cc = reinterpret_cast<const uint32_t*>(work.h_out.as<char>() + work.out_cc);
                bool over = false;
                for (uint32_t i = 0; i < n; ++i) over = over || cc[i] > cand_cap;
                if (over) {
                    exhaustive_dense(ix, src, work, d_ql, n, max_T, limit, ksplit, 1, st);
                    soa = true;
                    CK(cudaMemcpyAsync(work.h_out.p, work.d_out.p, work.out_keys, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                }
            }
            const uint64_t* off = reinterpret_cast<const uint64_t*>(work.h_out.as<char>() + work.out_off);
            const uint64_t total = off[n];
            if (join && joined != static_cast<size_t>(-1)) {
                // a later sub-batch of a joined batch: append
                HostList& J = (*lists)[joined];
                const uint64_t base = J.off.back(), have = J.doc.size();
                for (uint32_t i = 1; i <= n; ++i) J.off.push_back(base + off[i]);
                J.doc.resize(have + total);
                J.score.resize(have + total);
                if (total) {
                    work.h_out.ensure(work.out_keys + total * 8);
                    {
                        PhaseScope ps(ix, PH_D2H, st);
                        CK(cudaMemcpyAsync(work.h_out.as<char>() + work.out_keys, work.o_keys(), total * 8,
                                           cudaMemcpyDeviceToHost, st));
                    }
                    CK(cudaStreamSynchronize(st));
                    const char* got = work.h_out.as<char>() + work.out_keys;
                    if (soa) {
                        copy_u32(J.doc.data() + have, reinterpret_cast<const uint32_t*>(got), total);
                        copy_u32(J.score.data() + have, reinterpret_cast<const uint32_t*>(got) + total, total);
                    } else {
                        decode_keys(reinterpret_cast<const uint64_t*>(got), total, J.doc.data() + have,
                                    J.score.data() + have);
                    }
                }
                for (size_t i = 0; i < n; ++i)
                    (*where)[list[b + i]] = { static_cast<uint32_t>(joined), static_cast<uint32_t>(b + i) };
                continue;
            }
            lists->emplace_back();
            HostList& L = lists->back();
            L.off.assign(off, off + n + 1);
            if (join) joined = lists->size() - 1;
            // One pass that covers the whole batch and came out as doc[] | score[]: the arrays land
            // in the batch slot's own pinned buffer and are handed out as they are -- for lists of
            // every document of a large index, unpacking and copying cost more than the search.
            const bool hand_out = soa && res_pin != nullptr && sub >= list.size() && list.size() == ids.size() &&
                                  total * 8 <= ix->pinned_result_max;
            if (hand_out && total) {
                res_pin->ensure(total * 8);
                {
                    PhaseScope ps(ix, PH_D2H, st);
                    CK(cudaMemcpyAsync(res_pin->p, work.o_keys(), total * 8, cudaMemcpyDeviceToHost, st));
                }
                CK(cudaStreamSynchronize(st));
                L.doc.alias(res_pin->as<uint32_t>(), total);
                L.score.alias(res_pin->as<uint32_t>() + total, total);
            } else if (total) {
                L.doc.resize(total);
                L.score.resize(total);
                work.h_out.ensure(work.out_keys + total * 8);
                {
                    PhaseScope ps(ix, PH_D2H, st);
                    CK(cudaMemcpyAsync(work.h_out.as<char>() + work.out_keys, work.o_keys(), total * 8,
                                       cudaMemcpyDeviceToHost, st));
                }
                CK(cudaStreamSynchronize(st));
                const char* got = work.h_out.as<char>() + work.out_keys;
                if (soa) {
                    copy_u32(L.doc.data(), reinterpret_cast<const uint32_t*>(got), total);
                    copy_u32(L.score.data(), reinterpret_cast<const uint32_t*>(got) + total, total);
                } else {
                    decode_keys(reinterpret_cast<const uint64_t*>(got), total, L.doc.data(), L.score.data());
                }
            }
            for (size_t i = 0; i < n; ++i)
                (*where)[list[b + i]] = { static_cast<uint32_t>(lists->size() - 1), static_cast<uint32_t>(i) };
        }
    }
}

// The plan of a batch's main pass, from the request and the batch geometry:
//   a small limit        -> TOPK (bounded candidate lists, can never overflow)
//   threshold > 0        -> CAND with `max_candidates` slots per query (overflow -> redone)
//   otherwise            -> no main pass: every document passes, exhaustive at collect
// Queries beyond 16 bit-planes never take part in the main pass.
bool plan_main_pass(const cobsgpu_index* ix, const Slot& sl, double threshold, uint64_t limit,
                    uint32_t main_max_T, PassPlan* pl) {
    pl->lng = main_max_T > MAX_T_SHORT;
    pl->limit = limit;
    const uint32_t real = std::max<uint32_t>(ix->shard_real_docs, 1);
    if (limit >= 1 && limit <= TOPK_MAX_K) {
        const uint64_t cap = std::min<uint64_t>(static_cast<uint64_t>(ix->warps_per_query) * limit, real);
        // the candidate buffer of one batch must fit the workspace (batches are cut accordingly)
        if (cap * 8 <= ix->workspace_bytes) {
            pl->mode = MODE_TOPK;
            pl->cap = static_cast<uint32_t>(std::max<uint64_t>(cap, 1));
            return true;
        }
    }
    if (threshold > 0.0) {
        pl->mode = MODE_CAND;
        pl->cap = std::min<uint32_t>(ix->max_candidates, real);
        return true;
    }
    (void)sl;
    return false;
}

// On scope exit: when this slot had to grow a buffer, the idle slots of the ring follow at once.
// Allocations synchronise the device, so the whole ring is sized during the first batch instead
// of stalling once per slot.
struct WarmRing {
    cobsgpu_index* ix;
    Slot& sl;
    size_t before;
    ~WarmRing() {
        if (sl.footprint() == before) return;
        try {
            for (Slot& o : ix->slots)
                if (&o != &sl && !o.busy) o.match(sl);
        } catch (const Err&) {   // best effort: the slot allocates on its own turn instead
            cudaGetLastError();
        }
    }
};

Slot& take_slot(cobsgpu_index* ix) {
    Slot& sl = ix->slots[ix->next_slot];
    if (sl.busy)
        throw Err{ COBSGPU_ERR_INVALID_ARG,
                   "too many batches in flight: collect a ticket before submitting another" };
    ix->next_slot = (ix->next_slot + 1) % cobsgpu_index::N_SLOTS;
    return sl;
}

// Enqueues queries [q0, q1) of the caller's batch: upload + K1 on s_in, K2 + K3 on the main
// stream, header + a speculative prefix of the keys back to pinned memory on s_out.  Nothing
// here waits for the device.
Slot& submit_batch(cobsgpu_index* ix, const char* queries, const uint64_t* offsets, uint32_t q0,
                   uint32_t q1, double threshold, uint64_t limit) {
    Slot& sl = take_slot(ix);
    WarmRing warm{ ix, sl, sl.footprint() };
    const uint32_t nq = q1 - q0;
    // A handful of queries (the `cobs query <string>` pattern) is latency-bound: everything runs
    // on ONE stream -- no cross-stream hand-offs -- and K3 formats the result in its own launch.
    const bool small = nq <= FIN_WARPS;
    cudaStream_t st_in = small ? ix->stream : ix->s_in;
    cudaStream_t st_out = small ? ix->stream : ix->s_out;
    prepare_batch(ix, sl, queries, false, offsets, q0, q1, threshold, st_in, small);
    if (!small) CK(cudaEventRecord(sl.ev(sl.ev_in), st_in));
    sl.q0 = q0;
    sl.threshold = threshold;
    sl.limit = limit;
    sl.main_ids.clear();
    sl.huge_ids.clear();
    uint32_t main_max_T = sl.max_T;
    if (sl.max_T > MAX_T_LONG) {
        main_max_T = 1;
        for (uint32_t i = 0; i < nq; ++i) {
            const uint32_t T = sl.koff[i + 1] - sl.koff[i];
            if (T > MAX_T_LONG) sl.huge_ids.push_back(i);
            else {
                sl.main_ids.push_back(i);
                main_max_T = std::max(main_max_T, T);
            }
        }
    }
    const uint32_t n_main = sl.huge_ids.empty() ? nq : static_cast<uint32_t>(sl.main_ids.size());
    PassPlan pl;
    sl.mode = -1;
    // the slot only becomes a ticket once everything below has been enqueued: an error on the
    // way (a CUDA failure, an allocation) must not leave it marked busy for ever
    struct Commit {
        cobsgpu_index* ix;
        Slot& sl;
        bool ok = false;
        ~Commit() {
            if (!ok) return;
            sl.busy = true;
            sl.ticket = ++ix->ticket_counter;
        }
    } commit{ ix, sl };
    cudaStream_t st = ix->stream;
    if (!small) CK(cudaStreamWaitEvent(st, sl.ev_in, 0));
    // A few LONG queries (a gene or a plasmid against the index): with one work item per (query,
    // tile) only n_tiles CTAs would run, each walking a latency chain of thousands of k-mers.
    // Such batches take the k-split score kernel + the dense counting sort at collect instead.
    static const bool no_ksplit = std::getenv("COBSGPU_NO_KSPLIT") != nullptr;   // experiments only
    // (measured: an unsplit item walks ~0.18 us per k-mer; the split path costs ~40 us of extra
    // launches plus, without a threshold, a counting sort of ~0.3 us per 1000 documents -- split
    // only when the chain is clearly longer)
    sl.ksplit = !no_ksplit && sl.huge_ids.empty() && main_max_T >= 256 &&
                static_cast<uint64_t>(n_main) * ix->tiles.size() < 2ull * 3 * ix->sm_count &&
                0.18 * main_max_T > 1.3 * (40.0 + (threshold > 0.0 ? 0.00002 : 0.0003) * ix->shard_real_docs);
    if (n_main == 0 || sl.ksplit || !plan_main_pass(ix, sl, threshold, limit, main_max_T, &pl)) {
        // exhaustive at collect; only the invalid-base flag is needed from the device
        sl.layout_out(0);
        sl.d_out.ensure(sl.out_keys);
        sl.h_out.ensure(sl.out_keys);
        CK(cudaMemcpyAsync(sl.h_out.p, sl.d_flags(), 8, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(sl.ev(sl.ev_out), st));
        commit.ok = true;
        return sl;
    }
    sl.mode = pl.mode;
    sl.lng = pl.lng;
    sl.cap = pl.cap;
    const uint32_t* d_ql = sl.huge_ids.empty() ? nullptr : upload_qlist(sl, sl.main_ids, 0, n_main, st);
    const size_t ob = out_bytes(sl, n_main, pl);
    sl.d_out.ensure(ob);
    sl.d_res_count.ensure(static_cast<size_t>(n_main) * 4);
    const bool fuse_csr = small && sl.huge_ids.empty();
    const uint64_t max_keys = (ob - sl.out_keys) / 8;
    if (fuse_csr && sl.small_cc && ob <= (64u << 20)) {
        // Zero-copy result: K3 writes offsets, counts, flags and keys straight into the pinned
        // host buffer (device-mapped), so the batch is upload -> K1 -> K2 -> K3 and nothing else.
        sl.h_out.ensure(ob);
        char* mapped = nullptr;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&mapped), sl.h_out.p, 0));
        launch_pass_score(ix, sl, sl, d_ql, n_main, pl, sl.small_cc, st, true);
        launch_select(ix, sl, sl, d_ql, n_main, pl, main_max_T, sl.small_cc, sl.d_cand.as<uint64_t>(),
                      sl.d_res_count.as<uint32_t>(), pl.cap, false, st, true, mapped);
        sl.spec_keys = max_keys;
        CK(cudaEventRecord(sl.ev(sl.ev_out), st));
        commit.ok = true;
        return sl;
    }
    launch_pass_score(ix, sl, sl, d_ql, n_main, pl, sl.o_cc(), st);
    launch_select(ix, sl, sl, d_ql, n_main, pl, main_max_T, sl.o_cc(), sl.d_cand.as<uint64_t>(),
                  sl.d_res_count.as<uint32_t>(), pl.cap, false, st, fuse_csr);
    if (!fuse_csr) launch_csr(ix, sl, sl, n_main, pl.cap, st);
    if (!small) CK(cudaEventRecord(sl.ev(sl.ev_main), st));
    // ONE device-to-host copy in the common case: the header and the first spec_keys keys
    // (sized from what the previous batch returned, with headroom)
    sl.spec_keys = std::min<uint64_t>(
        max_keys, std::max<uint64_t>(std::max<uint64_t>(8192, 8ull * n_main), ix->spec_hint + ix->spec_hint / 4));
    sl.h_out.ensure(sl.out_keys + sl.spec_keys * 8);
    if (!small) CK(cudaStreamWaitEvent(st_out, sl.ev_main, 0));
    {
        PhaseScope ps(ix, PH_D2H, st_out);
        CK(cudaMemcpyAsync(sl.h_out.p, sl.d_out.p, sl.out_keys + sl.spec_keys * 8, cudaMemcpyDeviceToHost,
                           st_out));
    }
    CK(cudaEventRecord(sl.ev(sl.ev_out), st_out));
    commit.ok = true;
    return sl;
}

// Waits for a submitted batch and assembles its result lists (query order) in the slot.
void collect_batch(cobsgpu_index* ix, Slot& sl) {
    if (!sl.busy) throw Err{ COBSGPU_ERR_INVALID_ARG, "ticket is not in flight" };
    struct Release {
        Slot& sl;
        ~Release() { sl.busy = false; }
    } release{ sl };
    CK(cudaEventSynchronize(sl.ev_out));
    const uint32_t nq = sl.nq;
    const int* flags = sl.h_out.as<int>();
    if (flags[0] != FLAG_CLEAR) throw_bad_base(sl.q0 + static_cast<uint32_t>(flags[0]));
    sl.r_off.assign(1, 0);
    sl.r_doc.clear();
    sl.r_score.clear();

    std::vector<HostList> lists;
    std::vector<std::pair<uint32_t, uint32_t>> where(nq, { 0u, 0u });   // query -> (list, slot)
    std::vector<uint32_t> redo = sl.huge_ids;
    const bool all_main = sl.huge_ids.empty();
    if (sl.mode < 0) {
        if (all_main) {
            redo.resize(nq);
            for (uint32_t i = 0; i < nq; ++i) redo[i] = i;
        } else {
            redo.insert(redo.end(), sl.main_ids.begin(), sl.main_ids.end());
            std::sort(redo.begin(), redo.end());
        }
    } else {
        const uint32_t n_main = all_main ? nq : static_cast<uint32_t>(sl.main_ids.size());
        const char* h = sl.h_out.as<char>();
        const uint64_t* off = reinterpret_cast<const uint64_t*>(h + sl.out_off);
        const uint32_t* cc = reinterpret_cast<const uint32_t*>(h + sl.out_cc);
        const uint64_t total = off[n_main];
        ix->spec_hint = total;
        if (total > sl.spec_keys) {
            // more results than the speculative copy carried: fetch the rest
            sl.h_out.ensure_keep(sl.out_keys + total * 8, sl.out_keys + sl.spec_keys * 8);
            h = sl.h_out.as<char>();
            off = reinterpret_cast<const uint64_t*>(h + sl.out_off);
            cc = reinterpret_cast<const uint32_t*>(h + sl.out_cc);
            // (on its own stream: s_out already holds the downloads of the batches behind this one)
            CK(cudaMemcpyAsync(sl.h_out.as<char>() + sl.out_keys + sl.spec_keys * 8,
                               sl.o_keys() + sl.spec_keys, (total - sl.spec_keys) * 8,
                               cudaMemcpyDeviceToHost, ix->s_fix));
            CK(cudaStreamSynchronize(ix->s_fix));
        }
        lists.emplace_back();
        HostList& L = lists.back();
        L.off.assign(off, off + n_main + 1);
        L.doc.resize(total);
        L.score.resize(total);
        decode_keys(reinterpret_cast<const uint64_t*>(h + sl.out_keys), total, L.doc.data(), L.score.data());
        for (uint32_t i = 0; i < n_main; ++i) {
            const uint32_t q = all_main ? i : sl.main_ids[i];
            where[q] = { 0u, i };
            // more candidates than slots: redone exhaustively, nothing is ever dropped silently
            if (cc[i] > sl.cap) redo.push_back(q);
        }
        if (redo.empty() && all_main) {
            // the common case: the pass's CSR is the batch's result
            sl.r_off.swap(L.off);
            sl.r_doc.swap(L.doc);
            sl.r_score.swap(L.score);
            return;
        }
    }
    run_exhaustive(ix, sl, redo, sl.limit, &lists, &where, sl.ksplit,
                   redo.size() == nq ? &sl.h_res : nullptr);
    if (lists.size() == 1 && redo.size() == nq) {
        // the whole batch came out of one exhaustive pass, in query order: its CSR is the result
        // (no second copy of what may be hundreds of megabytes)
        HostList& L = lists[0];
        sl.r_off.swap(L.off);
        sl.r_doc.swap(L.doc);
        sl.r_score.swap(L.score);
        return;
    }
    uint64_t run = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        const HostList& L = lists[where[i].first];
        run += L.off[where[i].second + 1] - L.off[where[i].second];
    }
    sl.r_doc.reserve(run);     // one allocation, not a chain of growing copies
    sl.r_score.reserve(run);
    run = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        const HostList& L = lists[where[i].first];
        const uint32_t s = where[i].second;
        const uint64_t a = L.off[s], b = L.off[s + 1];
        sl.r_doc.insert(sl.r_doc.end(), L.doc.begin() + a, L.doc.begin() + b);
        sl.r_score.insert(sl.r_score.end(), L.score.begin() + a, L.score.begin() + b);
        run += b - a;
        sl.r_off.push_back(run);
    }
}

// batch boundaries: at most max_batch queries, a bounded number of k-mers, and candidate
// buffers that fit the workspace
uint32_t next_batch_end(const cobsgpu_index* ix, const uint64_t* offsets, uint32_t q0, uint32_t nq,
                        uint64_t limit) {
    const uint64_t kmer_budget = 64ull << 20;
    uint64_t max_q = ix->max_batch;
    if (limit >= 1 && limit <= TOPK_MAX_K) {
        const uint64_t cap = std::max<uint64_t>(1, static_cast<uint64_t>(ix->warps_per_query) * limit);
        max_q = std::max<uint64_t>(1, std::min<uint64_t>(max_q, ix->workspace_bytes / (cap * 8)));
    }
    uint64_t kmers = 0;
    uint32_t q = q0;
    while (q < nq && q - q0 < max_q) {
        const uint64_t len = offsets[q + 1] - offsets[q];
        const uint64_t T = len >= ix->term_size ? len - ix->term_size + 1 : 0;
        if (q > q0 && kmers + T > kmer_budget) break;
        kmers += T;
        ++q;
    }
    return q;
}

void check_idle(const cobsgpu_index* ix) {
    for (const Slot& sl : ix->slots)
        if (sl.busy)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "batches are in flight: collect every ticket first" };
}

void drop_tickets(cobsgpu_index* ix) {
    for (cudaStream_t st : { ix->s_in, ix->stream, ix->s_out, ix->s_fix })
        if (st) cudaStreamSynchronize(st);
    for (Slot& sl : ix->slots) sl.busy = false;
}

// ---------------------------------------------------------------------------------------
// CSR formatting with explicit list pointers (the group's merged lists live beside the
// leader's own candidates)
void launch_csr_lists(cobsgpu_index* ix, const Slot& src, Slot& work, uint32_t n_slots,
                      const uint64_t* lists, const uint32_t* counts, uint32_t stride,
                      cudaStream_t st) {
    PhaseScope ps(ix, PH_SELECT, st);
    scan_offsets_kernel<<<1, 1024, 0, st>>>(counts, n_slots, work.o_off());
    CK(cudaGetLastError());
    GatherKeysParams gp{ lists, counts, work.o_off(), stride, work.o_keys(), src.d_flags(),
                         work.o_flags() };
    gather_keys_kernel<<<n_slots, 256, 0, st>>>(gp);
    CK(cudaGetLastError());
    ix->tm.kernel_launches += 2;
}

}  // namespace

// A document-sharded index spread over several GPUs of ONE process: shard g lives on
// devices[g]; shard 0's device is the leader that merges.  Every shard runs K1-K3 on its own
// streams; the leader's merge kernel then reads the peers' fixed-size result blocks
// [nq][rpq] straight out of their HBM over NVLink (peer access) -- the one exchange of the
// path, k * 8 bytes per query per shard -- and returns ONE ordered list
// per query to the host.
struct cobsgpu_group {
    std::vector<cobsgpu_index*> shards;
    std::vector<int> devices;
    std::vector<char> peer_ok;       // leader can map shard g's memory
    struct GSlot {
        bool busy = false;
        int mode = 0;                // 0: merged on the leader, -1: per-shard host path
        uint32_t q0 = 0, nq = 0;
        uint32_t rpq = 0, out_stride = 0;
        double threshold = 0;
        uint64_t limit = 0;
        uint64_t spec_keys = 0;
        std::vector<Slot*> sl;       // the shards' slots of this batch
        DevBuf d_mkeys, d_mcount;    // merged lists on the leader
        DevBuf d_stage_keys, d_stage_counts;   // copies of peers' blocks when peer access is missing
        std::vector<uint64_t> r_off;
        U32Buf r_doc, r_score;
    } gs[cobsgpu_index::N_SLOTS];
    int next = 0;
    // result of the last cobsgpu_group_search_batch call
    std::vector<uint64_t> r_off;
    U32Buf r_doc, r_score;
    // a copy of the caller's queries that stays valid while batches are in flight is not
    // needed: cudaMemcpyAsync from pageable memory is staged before it returns

    ~cobsgpu_group() {
        for (cobsgpu_index* ix : shards) delete ix;   // (synchronises the shards' streams)
        // the slots' merge buffers on the leader are released by their own destructors
    }
};

namespace {

using GSlot = cobsgpu_group::GSlot;
static_assert(MERGE_RUNS_MAX_LISTS >= MERGE_MAX_LISTS, "merge_runs holds one cursor per shard");

void group_finish_open(cobsgpu_group* grp) {
    const size_t n = grp->shards.size();
    grp->peer_ok.assign(n, 1);
    CK(cudaSetDevice(grp->devices[0]));
    for (size_t g = 1; g < n; ++g) {
        if (grp->devices[g] == grp->devices[0]) continue;   // (tests: several shards on one GPU)
        int can = 0;
        cudaDeviceCanAccessPeer(&can, grp->devices[0], grp->devices[g]);
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(grp->devices[g], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            cudaGetLastError();
        }
        grp->peer_ok[g] = static_cast<char>(can);
    }
}

void group_drop(cobsgpu_group* grp) {
    for (cobsgpu_index* ix : grp->shards) {
        cudaSetDevice(ix->device);
        drop_tickets(ix);
    }
    for (auto& g : grp->gs) g.busy = false;
}

// Enqueues queries [q0, q1) on every shard and the merge on the leader.  Nothing waits.
GSlot& group_submit(cobsgpu_group* grp, const char* queries, const uint64_t* offsets, uint32_t q0,
                    uint32_t q1, double threshold, uint64_t limit) {
    GSlot& gsl = grp->gs[grp->next];
    if (gsl.busy) throw Err{ COBSGPU_ERR_INVALID_ARG, "too many group batches in flight" };
    grp->next = (grp->next + 1) % cobsgpu_index::N_SLOTS;
    const uint32_t n = static_cast<uint32_t>(grp->shards.size());
    const uint32_t nq = q1 - q0;
    gsl.q0 = q0;
    gsl.nq = nq;
    gsl.threshold = threshold;
    gsl.limit = limit;
    gsl.sl.assign(n, nullptr);
    gsl.mode = 0;
    // one stride for every shard: a small limit bounds each list (top-k epilogue), otherwise
    // the candidate slots, cut so that the merge fits its shared memory
    const bool topk = limit >= 1 && limit <= TOPK_MAX_K && limit * n <= MERGE_MAX;
    if (!topk && !(threshold > 0.0)) gsl.mode = -1;   // every document passes: per-shard exhaustive
    // (queries beyond 16 bit-planes are rare: such a batch takes the per-shard host path too)
    const uint64_t k = grp->shards[0]->term_size;
    for (uint32_t q = q0; q < q1 && gsl.mode == 0; ++q)
        if (offsets[q + 1] >= offsets[q] && offsets[q + 1] - offsets[q] > MAX_T_LONG + k - 1) gsl.mode = -1;
    uint32_t rpq = 1;
    if (gsl.mode == 0) {
        uint32_t mc = 1;
        for (cobsgpu_index* ix : grp->shards) mc = std::max(mc, ix->max_candidates);
        rpq = topk ? static_cast<uint32_t>(limit) : std::min<uint32_t>(mc, MERGE_MAX / n);
    }
    gsl.rpq = rpq;
    for (uint32_t g = 0; g < n; ++g) {
        cobsgpu_index* ix = grp->shards[g];
        CK(cudaSetDevice(ix->device));
        if (gsl.mode < 0) {
            gsl.sl[g] = &submit_batch(ix, queries, offsets, q0, q1, threshold, limit);
            continue;
        }
        Slot& sl = take_slot(ix);
        WarmRing warm{ ix, sl, sl.footprint() };
        sl.busy = true;
        sl.ticket = ++ix->ticket_counter;
        gsl.sl[g] = &sl;
        prepare_batch(ix, sl, queries, false, offsets, q0, q1, threshold, ix->s_in);
        CK(cudaEventRecord(sl.ev(sl.ev_in), ix->s_in));
        cudaStream_t st = ix->stream;
        CK(cudaStreamWaitEvent(st, sl.ev_in, 0));
        PassPlan pl;
        pl.lng = sl.max_T > MAX_T_SHORT;
        pl.limit = limit;
        const uint32_t real = std::max<uint32_t>(ix->shard_real_docs, 1);
        if (topk) {
            pl.mode = MODE_TOPK;
            pl.cap = static_cast<uint32_t>(std::max<uint64_t>(
                1, std::min<uint64_t>(static_cast<uint64_t>(ix->warps_per_query) * limit, real)));
        } else {
            pl.mode = MODE_CAND;
            pl.cap = std::min<uint32_t>(std::max<uint32_t>(rpq, ix->max_candidates), real);
        }
        sl.layout_out(nq);
        sl.d_out.ensure(sl.out_keys);
        sl.d_res_count.ensure(static_cast<size_t>(nq) * 4);
        sl.d_scratch.ensure(1);
        // the shard's result block [nq][rpq] lives in d_dense (unused on this path)
        sl.d_dense.ensure(static_cast<uint64_t>(nq) * rpq * 8);
        launch_pass_score(ix, sl, sl, nullptr, nq, pl, sl.o_cc(), st);
        launch_select(ix, sl, sl, nullptr, nq, pl, sl.max_T, sl.o_cc(), sl.d_dense.as<uint64_t>(),
                      sl.d_res_count.as<uint32_t>(), rpq, true, st);
        CK(cudaEventRecord(sl.ev(sl.ev_main), st));
    }
    gsl.busy = true;
    if (gsl.mode < 0) return gsl;

    // ---- the exchange + merge on the leader ----
    cobsgpu_index* lead = grp->shards[0];
    Slot& ls = *gsl.sl[0];
    CK(cudaSetDevice(lead->device));
    cudaStream_t st = lead->stream;
    MergeParams mp{};
    for (uint32_t g = 0; g < n; ++g) {
        Slot& sl = *gsl.sl[g];
        if (g) CK(cudaStreamWaitEvent(st, sl.ev_main, 0));
        const uint32_t* counts = sl.d_res_count.as<uint32_t>();
        const uint64_t* keys = sl.d_dense.as<uint64_t>();
        if (g && !grp->peer_ok[g]) {
            // no peer mapping: copy the block into a staging area on the leader instead
            gsl.d_stage_counts.ensure(static_cast<size_t>(n) * nq * 4);
            gsl.d_stage_keys.ensure(static_cast<size_t>(n) * nq * rpq * 8);
            uint32_t* dc = gsl.d_stage_counts.as<uint32_t>() + static_cast<size_t>(g) * nq;
            uint64_t* dk = gsl.d_stage_keys.as<uint64_t>() + static_cast<size_t>(g) * nq * rpq;
            CK(cudaMemcpyPeerAsync(dc, lead->device, counts, grp->shards[g]->device,
                                   static_cast<size_t>(nq) * 4, st));
            CK(cudaMemcpyPeerAsync(dk, lead->device, keys, grp->shards[g]->device,
                                   static_cast<size_t>(nq) * rpq * 8, st));
            counts = dc;
            keys = dk;
        }
        mp.counts[g] = counts;
        mp.keys[g] = keys;
    }
    const uint64_t all = static_cast<uint64_t>(n) * rpq;
    // merged lists are short in practice: the slots per query are bounded so that a batch's merge
    // buffers stay within ~256 MB on the leader (a merged list that outgrows them is flagged by
    // merge_kernel and redone like a candidate overflow, never cut)
    const uint64_t by_mem = std::max<uint64_t>(rpq, (256ull << 20) / (8ull * std::max<uint32_t>(nq, 1)));
    gsl.out_stride = static_cast<uint32_t>(limit ? std::min<uint64_t>(limit, all) : std::min<uint64_t>(all, by_mem));
    gsl.d_mkeys.ensure(static_cast<uint64_t>(nq) * gsl.out_stride * 8);
    gsl.d_mcount.ensure(static_cast<size_t>(nq) * 4);
    mp.n_lists = n;
    mp.nq = nq;
    mp.stride = rpq;
    mp.limit = limit;
    mp.out_stride = gsl.out_stride;
    mp.out_counts = gsl.d_mcount.as<uint32_t>();
    mp.out_keys = gsl.d_mkeys.as<uint64_t>();
    {
        PhaseScope ps(lead, PH_SELECT, st);
        const size_t smem = static_cast<size_t>(pow2_ceil(static_cast<uint32_t>(all))) * 8;
        CK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(MERGE_MAX * 8)));
        merge_kernel<<<nq, 256, smem, st>>>(mp);
        CK(cudaGetLastError());
        lead->tm.kernel_launches++;
    }
    // CSR + one copy back, like the single-GPU path; the header's cand_count words carry the
    // merged counts so that flagged queries are visible to the host
    ls.layout_out(nq);
    const size_t ob = ls.out_keys + static_cast<size_t>(nq) * gsl.out_stride * 8;
    ls.d_out.ensure(ob);
    launch_csr_lists(lead, ls, ls, nq, gsl.d_mkeys.as<uint64_t>(), gsl.d_mcount.as<uint32_t>(),
                     gsl.out_stride, st);
    CK(cudaMemcpyAsync(ls.o_cc(), gsl.d_mcount.p, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaEventRecord(ls.ev(ls.ev_main), st));
    gsl.spec_keys = std::min<uint64_t>(
        static_cast<uint64_t>(nq) * gsl.out_stride,
        std::max<uint64_t>(std::max<uint64_t>(8192, 8ull * nq), lead->spec_hint + lead->spec_hint / 4));
    ls.h_out.ensure(ls.out_keys + gsl.spec_keys * 8);
    CK(cudaStreamWaitEvent(lead->s_out, ls.ev_main, 0));
    {
        PhaseScope ps(lead, PH_D2H, lead->s_out);
        CK(cudaMemcpyAsync(ls.h_out.p, ls.d_out.p, ls.out_keys + gsl.spec_keys * 8, cudaMemcpyDeviceToHost,
                           lead->s_out));
    }
    CK(cudaEventRecord(ls.ev(ls.ev_out), lead->s_out));
    return gsl;
}

// A group batch whose lists hold every document (threshold <= 0 without a small limit): every
// shard collects on a host thread of its own -- the exhaustive passes run inside collect, so the
// GPUs work side by side instead of one after the other -- and the shards' lists are merged run
// by run, queries spread over host threads.
void group_collect_exhaustive(cobsgpu_group* grp, GSlot& gsl) {
    const uint32_t n = static_cast<uint32_t>(grp->shards.size());
    const uint32_t nq = gsl.nq;
    std::vector<Err> errs(n, Err{ COBSGPU_OK, "" });
    auto collect_one = [&](uint32_t g) {
        try {
            cobsgpu_index* ix = grp->shards[g];
            CK(cudaSetDevice(ix->device));
            collect_batch(ix, *gsl.sl[g]);
        } catch (const Err& e) {
            errs[g] = e;
        } catch (const std::bad_alloc&) {
            errs[g] = Err{ COBSGPU_ERR_OOM, "host out of memory" };
        } catch (const std::exception& e) {
            errs[g] = Err{ COBSGPU_ERR_INVALID_ARG, e.what() };
        }
    };
    {
        std::vector<std::thread> th;
        for (uint32_t g = 1; g < n; ++g) th.emplace_back(collect_one, g);
        collect_one(0);
        for (auto& t : th) t.join();
    }
    for (uint32_t g = 0; g < n; ++g)
        if (errs[g].code != COBSGPU_OK) {
            // an invalid base is reported by every shard: keep the message of the first
            throw errs[g];
        }
    gsl.r_off.resize(static_cast<size_t>(nq) + 1);
    gsl.r_off[0] = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        uint64_t len = 0;
        for (uint32_t g = 0; g < n; ++g) len += gsl.sl[g]->r_off[i + 1] - gsl.sl[g]->r_off[i];
        if (gsl.limit && len > gsl.limit) len = gsl.limit;
        gsl.r_off[i + 1] = gsl.r_off[i] + len;
    }
    const uint64_t total = gsl.r_off[nq];
    gsl.r_doc.resize(total);
    gsl.r_score.resize(total);
    auto merge_range = [&](uint32_t q0, uint32_t q1) {
        const uint32_t* doc[MERGE_MAX_LISTS];
        const uint32_t* score[MERGE_MAX_LISTS];
        uint64_t len[MERGE_MAX_LISTS];
        for (uint32_t i = q0; i < q1; ++i) {
            for (uint32_t g = 0; g < n; ++g) {
                const Slot& sl = *gsl.sl[g];
                doc[g] = sl.r_doc.begin() + sl.r_off[i];
                score[g] = sl.r_score.begin() + sl.r_off[i];
                len[g] = sl.r_off[i + 1] - sl.r_off[i];
            }
            merge_runs(n, doc, score, len, gsl.r_off[i + 1] - gsl.r_off[i], gsl.r_doc.data() + gsl.r_off[i],
                       gsl.r_score.data() + gsl.r_off[i]);
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    const uint32_t nt = static_cast<uint32_t>(std::min<uint64_t>(
        std::min<uint32_t>(nq, std::max(1u, std::min(hw, 8u))), std::max<uint64_t>(1, total >> 18)));
    if (nt <= 1) {
        merge_range(0, nq);
    } else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < nt; ++t)
            th.emplace_back(merge_range, static_cast<uint32_t>(static_cast<uint64_t>(nq) * t / nt),
                            static_cast<uint32_t>(static_cast<uint64_t>(nq) * (t + 1) / nt));
        for (auto& t : th) t.join();
    }
}

void group_collect(cobsgpu_group* grp, GSlot& gsl) {
    if (!gsl.busy) throw Err{ COBSGPU_ERR_INVALID_ARG, "group batch is not in flight" };
    const uint32_t n = static_cast<uint32_t>(grp->shards.size());
    const uint32_t nq = gsl.nq;
    struct Release {
        cobsgpu_group* grp;
        GSlot& g;
        ~Release() {
            g.busy = false;
            for (Slot* s : g.sl)
                if (s) s->busy = false;
        }
    } release{ grp, gsl };
    gsl.r_off.assign(1, 0);
    gsl.r_doc.clear();
    gsl.r_score.clear();
    if (gsl.mode < 0) {   // every list holds every document: per-shard collects + host merge
        group_collect_exhaustive(grp, gsl);
        return;
    }
    // queries that need the per-shard host path: the flagged ones
    std::vector<uint32_t> redo;
    std::vector<uint64_t> off_main;
    const uint64_t* keys_main = nullptr;
    {
        cobsgpu_index* lead = grp->shards[0];
        Slot& ls = *gsl.sl[0];
        CK(cudaSetDevice(lead->device));
        CK(cudaEventSynchronize(ls.ev_out));
        // the peers' streams have finished too: the merge waited for them
        const char* h = ls.h_out.as<char>();
        const int* flags = ls.h_out.as<int>();
        if (flags[0] != FLAG_CLEAR) throw_bad_base(gsl.q0 + static_cast<uint32_t>(flags[0]));
        const uint64_t* off = reinterpret_cast<const uint64_t*>(h + ls.out_off);
        const uint64_t total = off[nq];
        lead->spec_hint = total;
        if (total > gsl.spec_keys) {
            ls.h_out.ensure_keep(ls.out_keys + total * 8, ls.out_keys + gsl.spec_keys * 8);
            h = ls.h_out.as<char>();
            off = reinterpret_cast<const uint64_t*>(h + ls.out_off);
            CK(cudaMemcpyAsync(ls.h_out.as<char>() + ls.out_keys + gsl.spec_keys * 8,
                               ls.o_keys() + gsl.spec_keys, (total - gsl.spec_keys) * 8,
                               cudaMemcpyDeviceToHost, lead->s_fix));
            CK(cudaStreamSynchronize(lead->s_fix));
        }
        const uint32_t* cc = reinterpret_cast<const uint32_t*>(h + ls.out_cc);
        for (uint32_t i = 0; i < nq; ++i)
            if (cc[i] == COUNT_OVERFLOW) redo.push_back(i);
        off_main.assign(off, off + nq + 1);
        keys_main = reinterpret_cast<const uint64_t*>(h + ls.out_keys);
        if (redo.empty()) {
            gsl.r_off.swap(off_main);
            gsl.r_doc.resize(total);
            gsl.r_score.resize(total);
            decode_keys(keys_main, total, gsl.r_doc.data(), gsl.r_score.data());
            return;
        }
    }
    // flagged queries: re-run them through every shard's own host path (which falls back to its
    // exhaustive pass), straight from the hashes still resident in the shard's slot, and merge
    // the shards' lists on the host
    std::vector<std::vector<HostList>> lists(n);
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> where(
        n, std::vector<std::pair<uint32_t, uint32_t>>(nq, { 0u, 0u }));
    for (uint32_t g = 0; g < n; ++g) {
        cobsgpu_index* ix = grp->shards[g];
        CK(cudaSetDevice(ix->device));
        run_exhaustive(ix, *gsl.sl[g], redo, gsl.limit, &lists[g], &where[g]);
    }
    size_t rj = 0;
    uint64_t run = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        if (rj < redo.size() && redo[rj] == i) {
            const uint32_t* doc[MERGE_MAX_LISTS];
            const uint32_t* score[MERGE_MAX_LISTS];
            uint64_t len[MERGE_MAX_LISTS];
            uint64_t sum = 0;
            for (uint32_t g = 0; g < n; ++g) {
                const HostList& L = lists[g][where[g][i].first];
                const uint32_t s = where[g][i].second;
                doc[g] = L.doc.begin() + L.off[s];
                score[g] = L.score.begin() + L.off[s];
                len[g] = L.off[s + 1] - L.off[s];
                sum += len[g];
            }
            if (gsl.limit && sum > gsl.limit) sum = gsl.limit;
            gsl.r_doc.resize(run + sum);
            gsl.r_score.resize(run + sum);
            merge_runs(n, doc, score, len, sum, gsl.r_doc.data() + run, gsl.r_score.data() + run);
            run += sum;
            ++rj;
        } else {
            const uint64_t a = off_main[i], m = off_main[i + 1] - off_main[i];
            gsl.r_doc.resize(run + m);
            gsl.r_score.resize(run + m);
            decode_keys(keys_main + a, m, gsl.r_doc.data() + run, gsl.r_score.data() + run);
            run += m;
        }
        gsl.r_off.push_back(run);
    }
}

}  // namespace

// =========================================================================================
// C ABI

extern "C" {

const char* cobsgpu_last_error(void) { return g_error.c_str(); }
int cobsgpu_version(void) { return COBSGPU_VERSION; }

int cobsgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cobsgpu_index_open(const cobsgpu_index_desc* d, cobsgpu_index** out) {
    return guarded([&] {
        if (!d || !out || d->struct_size != sizeof(cobsgpu_index_desc))
            throw Err{ COBSGPU_ERR_INVALID_ARG, "bad cobsgpu_index_desc" };
        if (d->n_pages == 0 || !d->signature_sizes)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "index needs at least one page" };
        std::unique_ptr<cobsgpu_index> ix(new cobsgpu_index);
        ix->kind = d->kind;
        ix->term_size = d->term_size;
        ix->canonicalize = d->canonicalize;
        ix->num_hashes = d->num_hashes;
        ix->n_docs = d->n_docs;
        ix->n_pages_global = d->n_pages;
        ix->device = d->device;
        ix->shard_index = d->shard_index;
        ix->shard_count = d->shard_count ? d->shard_count : 1;
        if (d->kind == COBSGPU_KIND_CLASSIC) {
            if (d->n_pages != 1) throw Err{ COBSGPU_ERR_INVALID_ARG, "classic index has one page" };
            ix->page_size_src = (static_cast<uint64_t>(d->n_docs) + 7) / 8;
        } else if (d->kind == COBSGPU_KIND_COMPACT) {
            ix->page_size_src = d->page_size;
            if (d->page_size == 0 ||
                static_cast<uint64_t>(d->n_docs) > 8 * d->page_size * d->n_pages)
                throw Err{ COBSGPU_ERR_INVALID_ARG, "compact geometry does not cover n_docs" };
        } else {
            throw Err{ COBSGPU_ERR_INVALID_ARG, "unknown index kind" };
        }
        if (ix->page_size_src * 8 * d->n_pages > 0xFFFFFFFFull)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "more than 2^32 document columns" };
        std::vector<uint64_t> sig(d->signature_sizes, d->signature_sizes + d->n_pages);
        if (d->page_data) {
            const uint8_t* const* pd = d->page_data;
            const uint64_t ps = ix->page_size_src;
            RowReader rd = [pd, ps](uint32_t page, uint64_t row0, uint64_t nrows, uint8_t* dst) {
                std::memcpy(dst, pd[page] + row0 * ps, nrows * ps);
            };
            open_common(ix.get(), sig, &rd, 0);
        } else {
            open_common(ix.get(), sig, nullptr, d->fill_seed);
        }
        *out = ix.release();
    });
}

int cobsgpu_index_open_file(const char* path, int device, uint32_t shard_index,
                            uint32_t shard_count, cobsgpu_index** out) {
    return guarded([&] {
        if (!path || !out) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        std::unique_ptr<IndexFile> f(new IndexFile);
        std::string err = f->open(path);
        if (!err.empty()) {
            const bool io = err.rfind("could not open", 0) == 0 || err.rfind("mmap", 0) == 0;
            throw Err{ io ? COBSGPU_ERR_IO : COBSGPU_ERR_BAD_FILE, err };
        }
        std::unique_ptr<cobsgpu_index> ix(new cobsgpu_index);
        ix->kind = f->kind;
        ix->term_size = f->term_size;
        ix->canonicalize = f->canonicalize;
        ix->num_hashes = f->num_hashes;
        ix->n_docs = f->n_docs;
        ix->n_pages_global = static_cast<uint32_t>(f->signature_sizes.size());
        ix->page_size_src = f->page_size;
        ix->device = device;
        ix->shard_index = shard_index;
        ix->shard_count = shard_count ? shard_count : 1;
        ix->doc_names = f->doc_names;
        if (ix->page_size_src * 8 * ix->n_pages_global > 0xFFFFFFFFull)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "more than 2^32 document columns" };
        // matrix bytes are pread() straight into the pinned staging buffers
        IndexFile* fp = f.get();
        const uint64_t ps = ix->page_size_src;
        RowReader rd = [fp, ps](uint32_t page, uint64_t row0, uint64_t nrows, uint8_t* dst) {
            uint64_t pos = static_cast<uint64_t>(fp->page_data[page] - fp->map) + row0 * ps;
            uint64_t left = nrows * ps;
            while (left) {
                ssize_t n = pread(fp->fd, dst, left, static_cast<off_t>(pos));
                if (n <= 0) throw Err{ COBSGPU_ERR_IO, std::string("read failed: ") + std::strerror(errno) };
                dst += n;
                pos += static_cast<uint64_t>(n);
                left -= static_cast<uint64_t>(n);
            }
        };
        open_common(ix.get(), f->signature_sizes, &rd, 0);
        f->close();   // the matrix now lives in HBM
        *out = ix.release();
    });
}

int cobsgpu_construct_classic(const cobsgpu_construct_desc* d, cobsgpu_index** out) {
    return guarded([&] {
        if (!d || !out || d->struct_size != sizeof(cobsgpu_construct_desc))
            throw Err{ COBSGPU_ERR_INVALID_ARG, "bad cobsgpu_construct_desc" };
        if (d->n_docs == 0 || (d->n_seqs && (!d->sequences || !d->seq_offsets || !d->seq_doc)))
            throw Err{ COBSGPU_ERR_INVALID_ARG, "construction needs documents and sequences" };
        const uint32_t k = d->term_size;
        if (k == 0 || d->num_hashes == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "term_size and num_hashes must be > 0" };
        // windows (k-mers) per sequence and per document
        std::vector<uint64_t> win_off(static_cast<size_t>(d->n_seqs) + 1, 0), doc_terms(d->n_docs, 0);
        for (uint32_t s = 0; s < d->n_seqs; ++s) {
            if (d->seq_offsets[s + 1] < d->seq_offsets[s] || d->seq_doc[s] >= d->n_docs)
                throw Err{ COBSGPU_ERR_INVALID_ARG, "bad sequence table" };
            const uint64_t len = d->seq_offsets[s + 1] - d->seq_offsets[s];
            const uint64_t w = len >= k ? len - k + 1 : 0;
            win_off[s + 1] = win_off[s] + w;
            doc_terms[d->seq_doc[s]] += w;
        }
        uint64_t sig = d->signature_size;
        if (sig == 0) {
            // classic_construct: size the filters for the largest document
            // (cobs/construction/classic_index.cpp:571-575, cobs/util/calc_signature_size.cpp:15-33)
            const uint64_t max_doc = *std::max_element(doc_terms.begin(), doc_terms.end());
            const double hh = static_cast<double>(d->num_hashes);
            const double ratio = -hh / std::log(1 - std::pow(d->false_positive_rate, 1 / hh));
            if (!(ratio > 0)) throw Err{ COBSGPU_ERR_INVALID_ARG, "bad false_positive_rate" };
            sig = static_cast<uint64_t>(std::ceil(static_cast<double>(max_doc) * ratio));
        }
        if (sig == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "signature_size is 0 (empty documents?)" };

        std::unique_ptr<cobsgpu_index> ix(new cobsgpu_index);
        ix->kind = COBSGPU_KIND_CLASSIC;
        ix->term_size = k;
        ix->canonicalize = d->canonicalize;
        ix->num_hashes = d->num_hashes;
        ix->n_docs = d->n_docs;
        ix->n_pages_global = 1;
        ix->page_size_src = (static_cast<uint64_t>(d->n_docs) + 7) / 8;
        ix->device = d->device;
        ix->shard_index = 0;
        ix->shard_count = 1;
        ix->doc_names.resize(d->n_docs);
        for (uint32_t i = 0; i < d->n_docs; ++i)
            ix->doc_names[i] = d->doc_names && d->doc_names[i] ? d->doc_names[i] : "";
        open_common(ix.get(), { sig }, nullptr, 0, /*zero_fill=*/true);

        const uint64_t total = win_off.back();
        if (total) {
            const uint64_t nchar = d->seq_offsets[d->n_seqs];
            DevBuf d_seq, d_off, d_win, d_doc;
            d_seq.ensure(nchar + 16);
            d_off.ensure((d->n_seqs + 1) * 8);
            d_win.ensure((d->n_seqs + 1) * 8);
            d_doc.ensure(d->n_seqs * 4);
            cudaStream_t st = ix->stream;
            CK(cudaMemcpyAsync(d_seq.p, d->sequences, nchar, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_off.p, d->seq_offsets, (d->n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_win.p, win_off.data(), (d->n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_doc.p, d->seq_doc, d->n_seqs * 4, cudaMemcpyHostToDevice, st));
            const LocalPage& lp = ix->pages[0];
            ConstructParams cp{ d_seq.as<char>(), d_off.as<uint64_t>(), d_win.as<uint64_t>(),
                                d_doc.as<uint32_t>(), d->n_seqs, total, k, d->num_hashes,
                                d->canonicalize, sig, lp.d_base, lp.pitch };
            const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>(
                div_ceil<uint64_t>(total, 128), static_cast<uint64_t>(ix->sm_count) * 32));
            construct_classic_kernel<<<grid, 128, 0, st>>>(cp);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
        }
        *out = ix.release();
    });
}

int cobsgpu_index_save(const cobsgpu_index* ix, const char* path) {
    return guarded([&] {
        if (!ix || !path) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        if (ix->shard_count != 1)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "only an unsharded index can be saved" };
        CK(cudaSetDevice(ix->device));
        FILE* f = std::fopen(path, "wb");
        if (!f) throw Err{ COBSGPU_ERR_IO, std::string("could not open ") + path + " for writing" };
        struct Closer {
            FILE* f;
            ~Closer() { std::fclose(f); }
        } closer{ f };
        auto put = [&](const void* p, size_t n) {
            if (n && std::fwrite(p, 1, n, f) != n) throw Err{ COBSGPU_ERR_IO, "write failed" };
        };
        const bool classic = ix->kind == COBSGPU_KIND_CLASSIC;
        const uint32_t version = 1, n_docs = ix->n_docs;
        const uint8_t canon = static_cast<uint8_t>(ix->canonicalize);
        const uint64_t nh = ix->num_hashes;
        // headers: cobs/file/classic_index_header.cpp:26-36, compact_index_header.cpp:24-42
        put("COBS:", 5);
        put(classic ? "CLASSIC_INDEX" : "COMPACT_INDEX", 13);
        put(&version, 4);
        put(&ix->term_size, 4);
        put(&canon, 1);
        if (classic) {
            put(&n_docs, 4);
            put(&ix->signature_sizes[0], 8);
            put(&nh, 8);
        } else {
            const uint32_t np = ix->n_pages_global;
            put(&np, 4);
            put(&n_docs, 4);
            put(&ix->page_size_src, 8);
            for (uint32_t p = 0; p < np; ++p) {
                put(&ix->signature_sizes[p], 8);
                put(&nh, 8);
            }
        }
        for (uint32_t i = 0; i < n_docs; ++i) {
            const std::string& nm = i < ix->doc_names.size() ? ix->doc_names[i] : std::string();
            put(nm.data(), nm.size());
            put("\n", 1);
        }
        if (!classic) {
            const uint64_t pos = static_cast<uint64_t>(std::ftell(f));
            const uint64_t ps = ix->page_size_src;
            std::vector<char> pad((ps - ((pos + 13) % ps)) % ps, 0);
            put(pad.data(), pad.size());
        }
        put(classic ? "CLASSIC_INDEX" : "COMPACT_INDEX", 13);
        // matrix: device pitch -> the file's unpadded rows
        const uint64_t ps = ix->page_size_src;
        const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / std::max<uint64_t>(1, ps));
        std::vector<uint8_t> buf;
        for (const LocalPage& lp : ix->pages) {
            for (uint64_t r = 0; r < lp.sig; r += chunk_rows) {
                const uint64_t nr = std::min<uint64_t>(chunk_rows, lp.sig - r);
                buf.resize(nr * ps);
                CK(cudaMemcpy2D(buf.data(), ps, lp.d_base + r * lp.pitch, lp.pitch, ps, nr,
                                cudaMemcpyDeviceToHost));
                put(buf.data(), buf.size());
            }
        }
        if (std::fflush(f) != 0) throw Err{ COBSGPU_ERR_IO, "write failed" };
    });
}

void cobsgpu_index_close(cobsgpu_index* idx) { delete idx; }

int cobsgpu_index_get_info(const cobsgpu_index* ix, cobsgpu_index_info* o) {
    return guarded([&] {
        if (!ix || !o) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        o->kind = ix->kind;
        o->term_size = ix->term_size;
        o->canonicalize = ix->canonicalize;
        o->num_hashes = ix->num_hashes;
        o->n_docs = ix->n_docs;
        o->n_pages = ix->n_pages_global;
        const bool classic = ix->kind == COBSGPU_KIND_CLASSIC;
        o->page_size = classic ? 1 : ix->page_size_src;
        o->row_size = classic ? ix->page_size_src : ix->page_size_src * ix->n_pages_global;
        o->counts_size = 8 * ix->page_size_src * ix->n_pages_global;
        o->shard_index = ix->shard_index;
        o->shard_count = ix->shard_count;
        o->shard_doc_begin = ix->shard_doc_begin;
        o->shard_doc_end = ix->shard_doc_end;
        o->hbm_bytes = ix->hbm_bytes;
        o->bytes_per_kmer = ix->bytes_per_kmer;
        o->load_seconds = ix->load_seconds;
        o->load_bytes = ix->load_bytes;
        o->load_threads = ix->load_threads;
        o->reserved = 0;
    });
}

uint64_t cobsgpu_index_signature_size(const cobsgpu_index* ix, uint32_t page) {
    if (!ix || page >= ix->signature_sizes.size()) return 0;
    return ix->signature_sizes[page];
}

const char* cobsgpu_index_doc_name(const cobsgpu_index* ix, uint32_t doc) {
    if (!ix || doc >= ix->doc_names.size()) return nullptr;
    return ix->doc_names[doc].c_str();
}

int cobsgpu_set_option(cobsgpu_index* ix, const char* name, int64_t value) {
    return guarded([&] {
        if (!ix || !name) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        const std::string n = name;
        if (n == "max_candidates" && value >= 1) ix->max_candidates = static_cast<uint32_t>(std::min<int64_t>(value, 1 << 24));
        else if (n == "max_batch" && value >= 1) ix->max_batch = static_cast<uint32_t>(std::min<int64_t>(value, 1 << 22));
        else if (n == "workspace_mb" && value >= 1) ix->workspace_bytes = static_cast<uint64_t>(value) << 20;
        else if (n == "pipe_kb" && value >= 1) ix->pipe_bytes = static_cast<uint64_t>(value) << 10;
        else if (n == "pinned_max_mb" && value >= 0) ix->pinned_result_max = static_cast<uint64_t>(value) << 20;
        else if (n == "timing") ix->timing = value != 0;
        else if (n == "prefetch") ix->prefetch = value != 0;
        else if (n == "inputs_ready") ix->inputs_ready = value != 0;
        else if (n == "input_stream") {
            // a cudaStream_t passed as an integer; -1 = back to the caller's compute stream
            ix->input_stream_set = value != -1;
            ix->input_stream = value == -1 ? nullptr : reinterpret_cast<cudaStream_t>(static_cast<intptr_t>(value));
        }
        else throw Err{ COBSGPU_ERR_INVALID_ARG, "unknown option or bad value: " + n };
    });
}

int cobsgpu_hash(cobsgpu_index* ix, const char* queries, const uint64_t* offsets, uint32_t nq,
                 uint64_t* out) {
    return guarded([&] {
        if (!ix || !offsets || (!queries && nq)) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        CK(cudaSetDevice(ix->device));
        check_idle(ix);
        Slot& sl = ix->aux;
        cudaStream_t st = ix->stream;
        uint64_t done = 0;
        for (uint32_t q0 = 0; q0 < nq;) {
            const uint32_t q1 = next_batch_end(ix, offsets, q0, nq, 0);
            prepare_batch(ix, sl, queries, false, offsets, q0, q1, 0.0, st);
            const uint64_t n = static_cast<uint64_t>(sl.total_kmers) * ix->num_hashes;
            int flags[2];
            CK(cudaMemcpyAsync(out + done, sl.d_hashes.p, n * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(flags, sl.d_flags(), sizeof(flags), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (flags[0] != FLAG_CLEAR) throw_bad_base(q0 + static_cast<uint32_t>(flags[0]));
            done += n;
            q0 = q1;
        }
        resolve_timers(ix);
    });
}

int cobsgpu_scores(cobsgpu_index* ix, const char* queries, const uint64_t* offsets, uint32_t nq,
                   uint32_t* out) {
    return guarded([&] {
        if (!ix || !offsets || (!queries && nq) || !out) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        CK(cudaSetDevice(ix->device));
        check_idle(ix);
        Slot& sl = ix->aux;
        cudaStream_t st = ix->stream;
        const uint64_t counts_size = 8 * ix->page_size_src * ix->n_pages_global;
        for (uint32_t q0 = 0; q0 < nq;) {
            const uint32_t q1 = next_batch_end(ix, offsets, q0, nq, 0);
            prepare_batch(ix, sl, queries, false, offsets, q0, q1, 0.0, st);
            std::vector<uint32_t> ids[2];   // [0] short (u8 planes suffice), [1] long
            for (uint32_t i = 0; i < q1 - q0; ++i)
                ids[sl.koff[i + 1] - sl.koff[i] > MAX_T_SHORT ? 1 : 0].push_back(i);
            for (int lm = 0; lm < 2; ++lm) {
                const size_t esz = lm ? 4 : 1;
                const size_t sub = static_cast<size_t>(std::max<uint64_t>(
                    1, std::min<uint64_t>(ids[lm].size(), ix->workspace_bytes / (ix->dense_pitch * esz))));
                for (size_t b = 0; b < ids[lm].size(); b += sub) {
                    const size_t n = std::min(sub, ids[lm].size() - b);
                    const uint32_t* d_ql = upload_qlist(sl, ids[lm], b, n, st);
                    const size_t bytes = n * ix->dense_pitch * esz;
                    sl.d_dense.ensure(bytes);
                    CK(cudaMemsetAsync(sl.d_dense.p, 0, bytes, st));
                    ScoreParams sp = base_params(ix, sl, d_ql, static_cast<uint32_t>(n));
                    sp.dense8 = sl.d_dense.as<uint8_t>();
                    sp.dense32 = sl.d_dense.as<uint32_t>();
                    launch_score(ix, sp, lm ? MODE_DENSE32 : MODE_DENSE8, false, st);
                    sl.h_out.ensure(bytes);
                    CK(cudaMemcpyAsync(sl.h_out.p, sl.d_dense.p, bytes, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    // shard-local dense layout -> the reference's score_list layout
                    for (size_t i = 0; i < n; ++i) {
                        uint32_t* dst = out + static_cast<uint64_t>(q0 + ids[lm][b + i]) * counts_size;
                        for (auto& lp : ix->pages) {
                            const uint64_t gcol = static_cast<uint64_t>(lp.global_page) * 8 * ix->page_size_src +
                                                  8 * lp.byte_begin;
                            const uint64_t cols = static_cast<uint64_t>(lp.row_bytes) * 8;
                            if (lm) {
                                const uint32_t* s = sl.h_out.as<uint32_t>() + i * ix->dense_pitch + lp.dense_off;
                                std::memcpy(dst + gcol, s, cols * 4);
                            } else {
                                const uint8_t* s = sl.h_out.as<uint8_t>() + i * ix->dense_pitch + lp.dense_off;
                                for (uint64_t c = 0; c < cols; ++c) dst[gcol + c] = s[c];
                            }
                        }
                    }
                }
            }
            int flags[2];
            CK(cudaMemcpyAsync(flags, sl.d_flags(), sizeof(flags), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (flags[0] != FLAG_CLEAR) throw_bad_base(q0 + static_cast<uint32_t>(flags[0]));
            q0 = q1;
        }
        resolve_timers(ix);
    });
}

int cobsgpu_search_batch(cobsgpu_index* ix, const char* queries, const uint64_t* offsets,
                         uint32_t nq, double threshold, uint64_t num_results,
                         cobsgpu_result* out) {
    return guarded([&] {
        if (!ix || !offsets || (!queries && nq) || !out) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        CK(cudaSetDevice(ix->device));
        check_idle(ix);
        ix->r_off.assign(1, 0);
        ix->r_doc.clear();
        ix->r_score.clear();
        // sub-batches are pipelined through the slot ring: up to N_SLOTS - 1 in flight
        std::vector<Slot*> inflight;
        size_t head = 0, n_batches = 0;
        Slot* last = nullptr;
        auto drain_one = [&] {
            Slot& sl = *inflight[head++];
            collect_batch(ix, sl);
            last = &sl;
            if (++n_batches == 1 && head == inflight.size()) return;   // maybe the only batch
            if (n_batches == 2) {
                // first batch was left in its slot: move it over now
                Slot& f = *inflight[0];
                ix->r_off.insert(ix->r_off.end(), f.r_off.begin() + 1, f.r_off.end());
                ix->r_doc = f.r_doc;
                ix->r_score = f.r_score;
            }
            if (n_batches >= 2) {
                const uint64_t base = ix->r_off.back();
                for (size_t i = 1; i < sl.r_off.size(); ++i) ix->r_off.push_back(base + sl.r_off[i]);
                ix->r_doc.insert(ix->r_doc.end(), sl.r_doc.begin(), sl.r_doc.end());
                ix->r_score.insert(ix->r_score.end(), sl.r_score.begin(), sl.r_score.end());
            }
        };
        try {
            for (uint32_t q0 = 0; q0 < nq;) {
                const uint32_t q1 = next_batch_end(ix, offsets, q0, nq, num_results);
                if (inflight.size() - head == cobsgpu_index::N_SLOTS - 1) drain_one();
                inflight.push_back(&submit_batch(ix, queries, offsets, q0, q1, threshold, num_results));
                q0 = q1;
            }
            while (head < inflight.size()) drain_one();
        } catch (...) {
            drop_tickets(ix);
            throw;
        }
        resolve_timers(ix);
        if (n_batches == 1) {
            // single batch: hand out the slot's own arrays (valid until the next call)
            out->offsets = last->r_off.data();
            out->doc = last->r_doc.data();
            out->score = last->r_score.data();
            return;
        }
        out->offsets = ix->r_off.data();
        out->doc = ix->r_doc.data();
        out->score = ix->r_score.data();
    });
}

int cobsgpu_submit(cobsgpu_index* ix, const char* queries, const uint64_t* offsets, uint32_t nq,
                   double threshold, uint64_t num_results, cobsgpu_ticket* ticket) {
    return guarded([&] {
        if (!ix || !offsets || (!queries && nq) || !ticket) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        if (nq == 0) throw Err{ COBSGPU_ERR_INVALID_ARG, "empty batch" };
        CK(cudaSetDevice(ix->device));
        if (next_batch_end(ix, offsets, 0, nq, num_results) != nq)
            throw Err{ COBSGPU_ERR_INVALID_ARG,
                       "batch too large for one ticket (see the max_batch / workspace_mb options)" };
        Slot& sl = submit_batch(ix, queries, offsets, 0, nq, threshold, num_results);
        *ticket = sl.ticket;
    });
}

int cobsgpu_collect(cobsgpu_index* ix, cobsgpu_ticket ticket, cobsgpu_result* out) {
    return guarded([&] {
        if (!ix || !out) throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        CK(cudaSetDevice(ix->device));
        for (Slot& sl : ix->slots) {
            if (!sl.busy || sl.ticket != ticket) continue;
            collect_batch(ix, sl);
            resolve_timers(ix, false);
            out->offsets = sl.r_off.data();
            out->doc = sl.r_doc.data();
            out->score = sl.r_score.data();
            return;
        }
        throw Err{ COBSGPU_ERR_INVALID_ARG, "unknown ticket" };
    });
}

int cobsgpu_group_open_file(const char* path, const int32_t* devices, uint32_t n_devices,
                            cobsgpu_group** out) {
    return guarded([&] {
        if (!path || !devices || !out || n_devices == 0 || n_devices > MERGE_MAX_LISTS)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "a group holds 1..16 shards" };
        std::unique_ptr<cobsgpu_group> grp(new cobsgpu_group);
        // the shards load concurrently: one host thread per GPU, each streaming its own columns
        std::vector<cobsgpu_index*> ix(n_devices, nullptr);
        std::vector<int> rc(n_devices, COBSGPU_OK);
        std::vector<std::string> err(n_devices);
        std::vector<std::thread> th;
        for (uint32_t g = 0; g < n_devices; ++g)
            th.emplace_back([&, g] {
                rc[g] = cobsgpu_index_open_file(path, devices[g], g, n_devices, &ix[g]);
                if (rc[g] != COBSGPU_OK) err[g] = g_error;   // thread-local: read it here
            });
        for (auto& t : th) t.join();
        for (uint32_t g = 0; g < n_devices; ++g)
            if (ix[g]) {   // (the group's destructor closes whatever did open)
                grp->shards.push_back(ix[g]);
                grp->devices.push_back(devices[g]);
            }
        for (uint32_t g = 0; g < n_devices; ++g)
            if (rc[g] != COBSGPU_OK) throw Err{ rc[g], err[g] };
        group_finish_open(grp.get());
        *out = grp.release();
    });
}

int cobsgpu_group_open(const cobsgpu_index_desc* desc, const int32_t* devices, uint32_t n_devices,
                       cobsgpu_group** out) {
    return guarded([&] {
        if (!desc || !devices || !out || n_devices == 0 || n_devices > MERGE_MAX_LISTS)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "a group holds 1..16 shards" };
        std::unique_ptr<cobsgpu_group> grp(new cobsgpu_group);
        for (uint32_t g = 0; g < n_devices; ++g) {
            cobsgpu_index_desc d = *desc;
            d.device = devices[g];
            d.shard_index = g;
            d.shard_count = n_devices;
            cobsgpu_index* ix = nullptr;
            const int rc = cobsgpu_index_open(&d, &ix);
            if (rc != COBSGPU_OK) throw Err{ rc, g_error };
            grp->shards.push_back(ix);
            grp->devices.push_back(devices[g]);
        }
        group_finish_open(grp.get());
        *out = grp.release();
    });
}

void cobsgpu_group_close(cobsgpu_group* grp) { delete grp; }

uint32_t cobsgpu_group_size(const cobsgpu_group* grp) {
    return grp ? static_cast<uint32_t>(grp->shards.size()) : 0;
}

cobsgpu_index* cobsgpu_group_shard(cobsgpu_group* grp, uint32_t i) {
    return (grp && i < grp->shards.size()) ? grp->shards[i] : nullptr;
}

int cobsgpu_group_search_batch(cobsgpu_group* grp, const char* queries, const uint64_t* offsets,
                               uint32_t nq, double threshold, uint64_t num_results,
                               cobsgpu_result* out) {
    return guarded([&] {
        if (!grp || !offsets || (!queries && nq) || !out || grp->shards.empty())
            throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        for (cobsgpu_index* ix : grp->shards) check_idle(ix);
        grp->r_off.assign(1, 0);
        grp->r_doc.clear();
        grp->r_score.clear();
        // batch boundaries: the tightest of the shards' (workspace, max_batch); pipelined
        // through the ring like cobsgpu_search_batch
        std::vector<GSlot*> inflight;
        size_t head = 0;
        auto drain_one = [&] {
            GSlot& g = *inflight[head++];
            group_collect(grp, g);
            const uint64_t base = grp->r_off.back();
            for (size_t i = 1; i < g.r_off.size(); ++i) grp->r_off.push_back(base + g.r_off[i]);
            grp->r_doc.insert(grp->r_doc.end(), g.r_doc.begin(), g.r_doc.end());
            grp->r_score.insert(grp->r_score.end(), g.r_score.begin(), g.r_score.end());
        };
        try {
            for (uint32_t q0 = 0; q0 < nq;) {
                uint32_t q1 = nq;
                for (cobsgpu_index* ix : grp->shards)
                    q1 = std::min(q1, next_batch_end(ix, offsets, q0, nq, num_results));
                if (inflight.size() - head == cobsgpu_index::N_SLOTS - 1) drain_one();
                inflight.push_back(&group_submit(grp, queries, offsets, q0, q1, threshold, num_results));
                q0 = q1;
            }
            while (head < inflight.size()) drain_one();
        } catch (...) {
            group_drop(grp);
            throw;
        }
        for (cobsgpu_index* ix : grp->shards) resolve_timers(ix);
        out->offsets = grp->r_off.data();
        out->doc = grp->r_doc.data();
        out->score = grp->r_score.data();
    });
}

int cobsgpu_search_batch_device(cobsgpu_index* ix, const char* d_queries, const uint64_t* offsets,
                                uint32_t nq, double threshold, uint64_t num_results,
                                uint32_t results_per_query, uint32_t* d_counts, uint64_t* d_keys,
                                void* stream) {
    return guarded([&] {
        if (!ix || !offsets || (!d_queries && nq) || !d_counts || !d_keys || results_per_query == 0)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        if (nq == 0) return;
        CK(cudaSetDevice(ix->device));
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        Slot& sl = take_slot(ix);
        WarmRing warm{ ix, sl, sl.footprint() };
        // "prefetch": metadata upload + K1 run ahead on s_in, in a slot no earlier call still
        // reads, and only K2/K3 are ordered on the caller's stream
        cudaStream_t ks = st;
        if (ix->prefetch) {
            ks = ix->s_in;
            if (!ix->inputs_ready) {
                // d_queries may still be in the making: on the stream named by the "input_stream"
                // option (an upload stream of the caller's, so that K1 does not queue behind the
                // previous batch's K2 on the caller's compute stream), else on `stream` itself
                cudaStream_t src = ix->input_stream_set ? ix->input_stream : st;
                if (!ix->ev_caller) CK(cudaEventCreateWithFlags(&ix->ev_caller, cudaEventDisableTiming));
                CK(cudaEventRecord(ix->ev_caller, src));
                CK(cudaStreamWaitEvent(ks, ix->ev_caller, 0));
            }
            // the slot's previous K2/K3 (N_SLOTS calls ago) must have finished with its buffers
            if (sl.ev_main) CK(cudaStreamWaitEvent(ks, sl.ev_main, 0));
        }
        prepare_batch(ix, sl, d_queries, true, offsets, 0, nq, threshold, ks);
        if (ks != st) {
            CK(cudaEventRecord(sl.ev(sl.ev_in), ks));
            CK(cudaStreamWaitEvent(st, sl.ev_in, 0));
        }
        struct MainDone {   // marks the slot's buffers free once K2/K3 of this call are enqueued
            Slot& sl;
            cudaStream_t st;
            ~MainDone() { cudaEventRecord(sl.ev(sl.ev_main), st); }
        } main_done{ sl, st };
        if (sl.max_T > MAX_T_LONG)
            throw Err{ COBSGPU_ERR_INVALID_ARG,
                       "device-resident path handles queries of at most 65535 k-mers" };
        PassPlan pl;
        if (!plan_main_pass(ix, sl, threshold, num_results, sl.max_T, &pl) || pl.mode == MODE_CAND) {
            // (threshold <= 0 without a small limit: every query overflows and is flagged)
            pl.mode = MODE_CAND;
            pl.lng = sl.max_T > MAX_T_SHORT;
            pl.limit = num_results;
            pl.cap = std::min<uint32_t>(std::max<uint32_t>(results_per_query, ix->max_candidates),
                                        std::max<uint32_t>(ix->shard_real_docs, 1));
        }
        sl.layout_out(nq);
        sl.d_out.ensure(sl.out_keys);
        launch_pass_score(ix, sl, sl, nullptr, nq, pl, sl.o_cc(), st);
        launch_select(ix, sl, sl, nullptr, nq, pl, sl.max_T, sl.o_cc(), d_keys, d_counts,
                      results_per_query, true, st);
    });
}

int cobsgpu_merge_device(int device, uint32_t n_lists, uint32_t nq, uint32_t results_per_query,
                         const uint32_t* d_counts, uint64_t counts_list_stride,
                         const uint64_t* d_keys, uint64_t keys_list_stride, uint64_t num_results,
                         uint32_t out_per_query, uint32_t* d_out_counts, uint64_t* d_out_keys,
                         void* stream) {
    return guarded([&] {
        if (!d_counts || !d_keys || !d_out_counts || !d_out_keys || n_lists == 0 || out_per_query == 0)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "null argument" };
        if (nq == 0) return;
        check_device(device);
        const uint64_t maxk = static_cast<uint64_t>(n_lists) * results_per_query;
        if (maxk > MERGE_MAX)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "merge handles at most 8192 keys per query" };
        uint32_t np2 = 1;
        while (np2 < maxk) np2 <<= 1;
        const size_t smem = static_cast<size_t>(np2) * 8;
        CK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
        if (n_lists > MERGE_MAX_LISTS)
            throw Err{ COBSGPU_ERR_INVALID_ARG, "merge handles at most 16 lists" };
        const uint64_t cs = counts_list_stride ? counts_list_stride : nq;
        const uint64_t ks = keys_list_stride ? keys_list_stride : static_cast<uint64_t>(nq) * results_per_query;
        MergeParams mp{};
        for (uint32_t l = 0; l < n_lists; ++l) {
            mp.counts[l] = d_counts + l * cs;
            mp.keys[l] = d_keys + l * ks;
        }
        mp.n_lists = n_lists;
        mp.nq = nq;
        mp.stride = results_per_query;
        mp.limit = num_results;
        mp.out_stride = out_per_query;
        mp.out_counts = d_out_counts;
        mp.out_keys = d_out_keys;
        merge_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(mp);
        CK(cudaGetLastError());
    });
}

int cobsgpu_get_timers(const cobsgpu_index* ix, cobsgpu_timers* out) {
    if (!ix || !out) return COBSGPU_ERR_INVALID_ARG;
    resolve_timers(const_cast<cobsgpu_index*>(ix));
    *out = ix->tm;
    return COBSGPU_OK;
}

int cobsgpu_reset_timers(cobsgpu_index* ix) {
    if (!ix) return COBSGPU_ERR_INVALID_ARG;
    resolve_timers(ix);
    ix->tm = cobsgpu_timers{};
    return COBSGPU_OK;
}

int cobsgpu_debug_read_row(cobsgpu_index* ix, uint32_t page, uint64_t row, uint64_t begin,
                           uint64_t bytes, uint8_t* out) {
    return guarded([&] {
        if (!ix || !out || page >= ix->pages.size()) throw Err{ COBSGPU_ERR_INVALID_ARG, "bad page" };
        const LocalPage& lp = ix->pages[page];
        if (row >= lp.sig || begin + bytes > lp.pitch) throw Err{ COBSGPU_ERR_INVALID_ARG, "out of range" };
        CK(cudaSetDevice(ix->device));
        CK(cudaMemcpy(out, lp.d_base + row * lp.pitch + begin, bytes, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
