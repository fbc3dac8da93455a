// cobs_b200/csrc/select.cuh -- K3: ordering / top-k of the per-query candidate lists.
//
// Replaces the sort + truncation of counts_to_result (cobs/query/classic_search.cpp:128-157,
// 178-201): candidates were already filtered by the threshold (score kernel epilogue or
// dense_to_cand_kernel); here each query's keys are sorted ascending, which by construction
// of make_key() is (score descending, document ascending) -- the reference's comparator
// (classic_search.cpp:139-143) is a strict total order, so any correct sort reproduces it.
#pragma once

#include "common.cuh"

namespace cobsgpu {

static constexpr uint32_t SORT_LARGE_THREADS = 1024;
static constexpr uint32_t MERGE_MAX = 8192;        // keys per query in the shard merge
// per-query result counts of the device-resident path that flag a query instead of a list
static constexpr uint32_t COUNT_OVERFLOW = 0xFFFFFFFFu;   // more candidates than slots
static constexpr uint32_t COUNT_INVALID = 0xFFFFFFFEu;    // non-ACGT base, canonicalising index

// in-place ascending bitonic sort of n_pow2 keys in shared memory by the whole CTA
__device__ __forceinline__ void block_bitonic_sort(uint64_t* s, uint32_t n_pow2) {
    for (uint32_t k = 2; k <= n_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = s[i], b = s[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s[i] = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

// dense u32 scores -> candidate keys (queries with more than 255 k-mers).
struct DenseToCandParams {
    const uint32_t* dense32;   // [nq_items * dense_pitch]
    uint64_t dense_pitch;
    const uint32_t* qlist;     // optional
    const uint32_t* thr;       // [nq] by batch query index
    // page segments of the shard-local dense layout
    const uint32_t* seg_dense_off;
    const uint32_t* seg_n_real;
    const uint32_t* seg_doc_base;
    uint32_t n_seg;
    uint32_t* cand_count;
    uint64_t* cand;
    uint32_t cap;
};

__global__ void __launch_bounds__(256) dense_to_cand_kernel(DenseToCandParams p) {
    const uint32_t qi = blockIdx.y;
    const uint32_t q = p.qlist ? p.qlist[qi] : qi;
    const uint32_t thr = p.thr[q];
    const uint32_t* sc = p.dense32 + static_cast<uint64_t>(qi) * p.dense_pitch;
    uint64_t* out = p.cand + static_cast<uint64_t>(qi) * p.cap;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t sgi = 0; sgi < p.n_seg; ++sgi) {
        const uint32_t n = p.seg_n_real[sgi], off = p.seg_dense_off[sgi], db = p.seg_doc_base[sgi];
        const uint32_t n_round = round_up<uint32_t>(n, 32);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
             i += gridDim.x * blockDim.x) {
            uint32_t v = 0;
            bool keep = false;
            if (i < n) {
                v = sc[off + i];
                keep = v >= thr;
            }
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (bal) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&p.cand_count[qi], __popc(bal));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (keep) {
                    const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
                    if (pos < p.cap) out[pos] = make_key(v, db + i);
                }
            }
        }
    }
}

// exclusive prefix sum of res_count into 64-bit offsets[nq+1]; single CTA
__global__ void __launch_bounds__(1024) scan_offsets_kernel(const uint32_t* res_count, uint32_t nq,
                                                            uint64_t* offsets) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nq; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint64_t v = i < nq ? res_count[i] : 0;
        if (v >= COUNT_INVALID) v = 0;   // flagged query: no list
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sum[lane];
            uint64_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sum[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        const uint64_t c = carry;
        const uint64_t excl = c + warp_sum[warp] + incl - v;
        if (i < nq) offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[nq] = carry;
}

// one CTA per query with more than `min_n` candidates (what finalize_kernel cannot sort in
// shared memory): LSD radix sort (8-bit digits)
// over the varying key bits, ping-ponging between `cand` and `scratch` (same layout).
// digit_shift[pass] lists the bit offsets to sort on, lowest significance first.
struct SortLargeParams {
    uint64_t* cand;
    uint64_t* scratch;
    const uint32_t* cand_count;
    uint32_t cap;
    uint32_t min_n;        // lists of at most this many keys are left to finalize_kernel
    uint32_t n_pass;
    uint32_t digit_shift[8];
};

__global__ void __launch_bounds__(SORT_LARGE_THREADS) sort_large_kernel(SortLargeParams p) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t warp_cnt[32][256];
    const uint32_t qi = blockIdx.x;
    uint32_t n = p.cand_count[qi];
    if (n > p.cap) n = p.cap;
    if (n <= p.min_n) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* src = p.cand + static_cast<uint64_t>(qi) * p.cap;
    uint64_t* dst = p.scratch + static_cast<uint64_t>(qi) * p.cap;
    for (uint32_t pass = 0; pass < p.n_pass; ++pass) {
        const uint32_t shift = p.digit_shift[pass];
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            atomicAdd(&hist[(src[i] >> shift) & 0xFFu], 1u);
        __syncthreads();
        // exclusive scan of the 256 bins by warp 0 (8 bins per lane)
        if (warp == 0) {
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                loc[b] = hist[lane * 8 + b];
                sum += loc[b];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                hist[lane * 8 + b] = run;
                run += loc[b];
            }
        }
        __syncthreads();
        // stable scatter, one chunk of blockDim.x keys at a time, in input order
        for (uint32_t base = 0; base < n; base += blockDim.x) {
            for (uint32_t i = threadIdx.x; i < 32 * 256; i += blockDim.x)
                (&warp_cnt[0][0])[i] = 0;
            __syncthreads();
            const uint32_t i = base + threadIdx.x;
            const bool valid = i < n;
            uint64_t key = 0;
            uint32_t dig = 0, rank = 0;
            if (valid) {
                key = src[i];
                dig = static_cast<uint32_t>(key >> shift) & 0xFFu;
            }
            const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid) {
                const uint32_t peers = __match_any_sync(act, dig);
                rank = __popc(peers & ((1u << lane) - 1u));
                if (rank == 0) warp_cnt[warp][dig] = __popc(peers);
            }
            __syncthreads();
            // per digit: running offsets across the warps of this chunk
            if (threadIdx.x < 256) {
                uint32_t run = hist[threadIdx.x];
                for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
                    const uint32_t c = warp_cnt[w][threadIdx.x];
                    warp_cnt[w][threadIdx.x] = run;
                    run += c;
                }
                hist[threadIdx.x] = run;
            }
            __syncthreads();
            if (valid) dst[warp_cnt[warp][dig] + rank] = key;
            __syncthreads();
        }
        uint64_t* tmp = src;
        src = dst;
        dst = tmp;
        __syncthreads();
    }
}

// Fused K3: result count + sort + output in ONE kernel.  A CTA owns FIN_WARPS queries: lists of
// <= 32 candidates (the common case at the CLI's default threshold) are sorted by one warp with
// shuffles; longer ones (<= FIN_SORT_MAX, dynamic shared memory) by the whole CTA afterwards;
// lists longer than fin_sort_max were sorted by sort_large_kernel before and are only counted
// (and copied when the output is not in place).  Output: the first min(n, limit) sorted keys of
// query q at out_keys[q * stride] (out_keys == cand, stride == cap sorts in place) and
// out_counts[q] -- or a COUNT_* flag instead of an incomplete list.
static constexpr uint32_t FIN_WARPS = 8;
static constexpr uint32_t FIN_SORT_MAX = 8192;   // 64 KB of dynamic shared memory

struct FinalizeParams {
    uint64_t* cand;
    const uint64_t* scratch;   // where sort_large left its result when large_in_scratch
    const uint32_t* cand_count;
    const uint32_t* bad;       // optional [nq] by batch query, == 1: non-ACGT base (canonical index)
    const uint32_t* qlist;     // optional: slot -> batch query (only used to look up `bad`)
    uint32_t cap;
    uint32_t nq;
    uint64_t limit;
    uint64_t* out_keys;        // [nq * stride]
    uint32_t* out_counts;      // [nq]
    uint32_t stride;
    uint32_t fin_sort_max;     // keys the CTA sort may hold (<= FIN_SORT_MAX, power of two)
    uint32_t large_in_scratch;
    // small batches (nq <= FIN_WARPS, one CTA): CSR formatting fused in -- offsets, keys and the
    // batch's invalid-base flag land in the result area without two more launches
    uint64_t* csr_off;         // optional [nq + 1]
    uint64_t* csr_keys;
    const int* flags_src;
    int* flags_dst;
    uint32_t* cc_dst;          // optional: copy of cand_count next to the offsets (the targets may
                               // be device-mapped HOST memory: no copy engine involved at all)
};

static constexpr uint32_t FIN_SEL_UNROLL = 8;   // chunks of 32 candidates in flight per warp
static constexpr uint32_t FIN_COOP_MAX_Q = 2;   // batches this small share the CTA's warps per query

// Streaming selection for small limits (`-l 10`), one warp: returns the k smallest of
// keys[begin, end) sorted over the lanes (lane i = i-th smallest, KEY_PAD where there are fewer;
// lanes >= k hold larger keys or KEY_PAD).  A chunk of 32 candidates is tested against the
// current k-th key with one ballot and only the few that beat it are inserted (expected
// k * ln(n / k) insertions): no shared memory, no sort of the whole list.  FIN_SEL_UNROLL chunks
// are loaded up front -- with one load per chunk the loop is a chain of L2 latencies, 77 of them
// for the 2450 per-warp candidates of a single `-l 10` query on a million documents.
__device__ __forceinline__ uint64_t select_stream(const uint64_t* keys, uint32_t begin, uint32_t end,
                                                  uint32_t k, uint32_t lane) {
    uint64_t best = KEY_PAD;
    for (uint32_t base = begin; base < end; base += 32 * FIN_SEL_UNROLL) {
        uint64_t chunk[FIN_SEL_UNROLL];
#pragma unroll
        for (uint32_t j = 0; j < FIN_SEL_UNROLL; ++j) {
            const uint32_t i = base + j * 32 + lane;
            chunk[j] = i < end ? keys[i] : KEY_PAD;
        }
#pragma unroll
        for (uint32_t j = 0; j < FIN_SEL_UNROLL; ++j) {
            const uint64_t key = chunk[j];
            const uint64_t kth = __shfl_sync(0xFFFFFFFFu, best, k - 1);
            uint32_t m = __ballot_sync(0xFFFFFFFFu, key < kth);
            while (m) {
                const uint32_t src = __ffs(m) - 1;
                m &= m - 1;
                const uint64_t x = __shfl_sync(0xFFFFFFFFu, key, src);
                const uint32_t pos = __popc(__ballot_sync(0xFFFFFFFFu, best < x));
                const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, best, 1);
                if (lane > pos) best = up;
                else if (lane == pos) best = x;
            }
        }
    }
    return best;
}

__global__ void __launch_bounds__(FIN_WARPS * 32) finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) uint64_t fin_s[];
    __shared__ uint32_t big_n[FIN_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * FIN_WARPS + warp;
    const bool in_place = p.out_keys == p.cand;
    if (lane == 0) big_n[warp] = 0;
    // One or two queries (the `cobs query <string>` pattern): a lone warp walking a query's
    // candidates is most of K3's time, so every warp of the CTA first selects from its share of
    // each query's list and the query's own warp then selects among those partial results.
    __shared__ uint64_t coop_s[FIN_COOP_MAX_Q][FIN_WARPS * 32];
    const bool coop = p.nq <= FIN_COOP_MAX_Q && p.limit != 0 && p.limit <= 32;   // uniform
    if (coop) {
        for (uint32_t cq = 0; cq < p.nq; ++cq) {
            const uint32_t c = p.cand_count[cq];
            const uint32_t n = c < p.cap ? c : p.cap;
            const uint32_t per = ((n + FIN_WARPS - 1) / FIN_WARPS + 31) & ~31u;
            const uint32_t b = warp * per < n ? warp * per : n;
            const uint32_t e = b + per < n ? b + per : n;
            coop_s[cq][warp * 32 + lane] = select_stream(p.cand + static_cast<uint64_t>(cq) * p.cap, b, e,
                                                         static_cast<uint32_t>(p.limit), lane);
        }
        __syncthreads();
    }
    if (q < p.nq) {
        const uint32_t c = p.cand_count[q];
        uint32_t n = c < p.cap ? c : p.cap;
        const uint32_t r = (p.limit != 0 && n > p.limit) ? static_cast<uint32_t>(p.limit) : n;
        const bool invalid = p.bad && p.bad[p.qlist ? p.qlist[q] : q] == 1;
        // more candidates than slots, or a list longer than the caller's stride: flagged like
        // an overflow -- a cut list must never look like a complete one
        const bool over = c > p.cap || r > p.stride;
        if (lane == 0) p.out_counts[q] = invalid ? COUNT_INVALID : (over ? COUNT_OVERFLOW : r);
        if (over || invalid) n = 0;
        if (n > 32 && p.limit != 0 && p.limit <= 32) {
            // the k best of a long list (the union of the per-warp top-k lists): streaming
            // selection -- over the whole list, or over the partial selections of the CTA's warps
            const uint32_t k = static_cast<uint32_t>(p.limit);
            const uint64_t best = coop ? select_stream(coop_s[q], 0, FIN_WARPS * 32, k, lane)
                                       : select_stream(p.cand + static_cast<uint64_t>(q) * p.cap, 0, n, k, lane);
            if (lane < r) p.out_keys[static_cast<uint64_t>(q) * p.stride + lane] = best;
        } else if (n > 32) {
            if (lane == 0) big_n[warp] = n;
        } else if (n > 0) {
            uint64_t key = lane < n ? p.cand[static_cast<uint64_t>(q) * p.cap + lane] : KEY_PAD;
#pragma unroll
            for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    const uint64_t other = __shfl_xor_sync(0xFFFFFFFFu, key, j);
                    const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
                    key = take_min ? (key < other ? key : other) : (key > other ? key : other);
                }
            }
            if (lane < r) p.out_keys[static_cast<uint64_t>(q) * p.stride + lane] = key;
        }
    }
    __syncthreads();
    for (uint32_t w = 0; w < FIN_WARPS; ++w) {
        const uint32_t bn = big_n[w];
        if (bn == 0) continue;   // uniform across the CTA
        const uint32_t bq = blockIdx.x * FIN_WARPS + w;
        uint32_t br = (p.limit != 0 && bn > p.limit) ? static_cast<uint32_t>(p.limit) : bn;
        uint64_t* o = p.out_keys + static_cast<uint64_t>(bq) * p.stride;
        if (bn > p.fin_sort_max) {
            // already sorted by sort_large_kernel
            const uint64_t* keys = (p.large_in_scratch ? p.scratch : p.cand) + static_cast<uint64_t>(bq) * p.cap;
            if (keys != o)
                for (uint32_t i = threadIdx.x; i < br; i += blockDim.x) o[i] = keys[i];
            continue;
        }
        const uint64_t* keys = p.cand + static_cast<uint64_t>(bq) * p.cap;
        const uint32_t np2 = next_pow2(bn);
        for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) fin_s[i] = i < bn ? keys[i] : KEY_PAD;
        __syncthreads();
        block_bitonic_sort(fin_s, np2);
        for (uint32_t i = threadIdx.x; i < br; i += blockDim.x) o[i] = fin_s[i];
        __syncthreads();
    }
    (void)in_place;
    if (p.csr_off != nullptr) {   // (host guarantees gridDim.x == 1 and nq <= FIN_WARPS)
        __shared__ uint64_t off_s[FIN_WARPS + 1];
        __syncthreads();          // the lists and counts of this CTA are complete
        if (threadIdx.x == 0) {
            uint64_t run = 0;
            for (uint32_t i = 0; i < p.nq; ++i) {
                off_s[i] = run;
                const uint32_t c = p.out_counts[i];
                run += c >= COUNT_INVALID ? 0 : c;
            }
            off_s[p.nq] = run;
            for (uint32_t i = 0; i <= p.nq; ++i) p.csr_off[i] = off_s[i];
            p.flags_dst[0] = p.flags_src[0];
            p.flags_dst[1] = p.flags_src[1];
            if (p.cc_dst)
                for (uint32_t i = 0; i < p.nq; ++i) p.cc_dst[i] = p.cand_count[i];
        }
        __syncthreads();
        if (q < p.nq) {
            const uint32_t c = static_cast<uint32_t>(off_s[warp + 1] - off_s[warp]);
            const uint64_t* src = p.out_keys + static_cast<uint64_t>(q) * p.stride;
            uint64_t* dst = p.csr_keys + off_s[warp];
            for (uint32_t i = lane; i < c; i += 32) dst[i] = src[i];
        }
    }
}

// CSR formatting of finalized lists: keys of slot qi (sorted in place at cand[qi * cap], count
// res_count[qi]) -> out_keys[offsets[qi] ...]; CTA 0 also copies the batch's invalid-base flag
// next to the offsets so that ONE device-to-host copy returns everything.
struct GatherKeysParams {
    const uint64_t* cand;
    const uint32_t* res_count;
    const uint64_t* offsets;
    uint32_t cap;
    uint64_t* out_keys;
    const int* flags_src;   // d_meta flags
    int* flags_dst;         // d_out header
};

__global__ void __launch_bounds__(256) gather_keys_kernel(GatherKeysParams p) {
    const uint32_t qi = blockIdx.x;
    if (qi == 0 && threadIdx.x < 2 && p.flags_dst) p.flags_dst[threadIdx.x] = p.flags_src[threadIdx.x];
    uint32_t r = p.res_count[qi];
    if (r >= COUNT_INVALID) r = 0;
    const uint64_t* src = p.cand + static_cast<uint64_t>(qi) * p.cap;
    uint64_t* dst = p.out_keys + p.offsets[qi];
    for (uint32_t i = threadIdx.x; i < r; i += blockDim.x) dst[i] = src[i];
}

// shard merge: per query concatenate n_lists sorted lists, sort, keep the first `limit`.  The
// lists are addressed by pointer, so they may live on other GPUs of the process: the leader's
// merge then reads its peers' result blocks directly over NVLink (peer access enabled).
static constexpr uint32_t MERGE_MAX_LISTS = 16;

struct MergeParams {
    const uint32_t* counts[MERGE_MAX_LISTS];   // list l: [nq]
    const uint64_t* keys[MERGE_MAX_LISTS];     // list l: [nq][stride]
    uint32_t n_lists, nq, stride;
    uint64_t limit;
    uint32_t out_stride;
    uint32_t* out_counts;     // [nq]
    uint64_t* out_keys;       // [nq][out_stride]
};

__global__ void __launch_bounds__(256) merge_kernel(MergeParams p) {
    extern __shared__ __align__(16) uint64_t ms[];
    __shared__ uint32_t tot, ovf;
    const uint32_t q = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t t = 0, o = 0;
        for (uint32_t l = 0; l < p.n_lists; ++l) {
            const uint32_t c = p.counts[l][q];
            if (c >= COUNT_INVALID) o = o > c ? o : c;   // a shard flagged the query
            else t += c < p.stride ? c : p.stride;
        }
        tot = t;
        ovf = o;
    }
    __syncthreads();
    if (ovf) {
        if (threadIdx.x == 0) p.out_counts[q] = ovf;
        return;
    }
    const uint32_t total = tot;
    // lists are short; every thread walks the list table
    uint32_t start = 0;
    for (uint32_t l = 0; l < p.n_lists; ++l) {
        uint32_t c = p.counts[l][q];
        if (c > p.stride) c = p.stride;   // (flagged lists returned above)
        const uint64_t* src = p.keys[l] + static_cast<uint64_t>(q) * p.stride;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) ms[start + i] = src[i];
        start += c;
    }
    const uint32_t np2 = next_pow2(total);
    for (uint32_t i = total + threadIdx.x; i < np2; i += blockDim.x) ms[i] = KEY_PAD;
    __syncthreads();
    if (total > 1) block_bitonic_sort(ms, np2);
    uint32_t r = total;
    if (p.limit != 0 && r > p.limit) r = static_cast<uint32_t>(p.limit);
    // a merged list longer than the output stride is flagged, never cut silently
    if (r > p.out_stride) {
        if (threadIdx.x == 0) p.out_counts[q] = COUNT_OVERFLOW;
        return;
    }
    uint64_t* o = p.out_keys + static_cast<uint64_t>(q) * p.out_stride;
    for (uint32_t i = threadIdx.x; i < r; i += blockDim.x) o[i] = ms[i];
    if (threadIdx.x == 0) p.out_counts[q] = r;
}

}  // namespace cobsgpu
