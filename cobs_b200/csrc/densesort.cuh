// cobs_b200/csrc/densesort.cuh -- K3 for EXHAUSTIVE result lists: every (real) document of the
// shard with score >= threshold, ordered (score desc, doc asc).
//
// This is what counts_to_result (cobs/query/classic_search.cpp:109-157) produces when the
// threshold lets (nearly) everything through -- Search::search's default arguments (threshold
// 0.0, all results; cobs/query/search.hpp:39-42) and `cobs benchmark-fpr` (src/cobs.cpp:620-626).
// The score kernel stores one count per document (DENSE8 / DENSE16, document order); the order
// wanted is then a STABLE counting sort on the score alone, because document ids already ascend:
// one pass of {histogram per 2048-document chunk, scan per query, stable scatter} for counts of
// one byte, two passes (low byte, then high byte: LSD) for 16-bit counts.  All passes are spread
// over chunks x queries CTAs; nothing sorts 64-bit keys and no CTA walks a whole list alone (the
// radix sort this replaces spent ~15 ms per million-document query in ONE CTA).
#pragma once

#include "common.cuh"

namespace cobsgpu {

static constexpr uint32_t DS_CHUNK = 2048;     // documents (or keys) per CTA
static constexpr uint32_t DS_THREADS = 256;

struct DenseSortParams {
    // source A (pass 0 / 1): dense per-document counts, shard-local dense layout
    const uint8_t* dense8;      // [n_slots][dense_pitch]   (pass 0)
    const uint16_t* dense16;    // [n_slots][dense_pitch]   (pass 1)
    uint64_t dense_pitch;
    uint32_t dense_cols;        // columns of the dense layout that tiles cover
    const uint32_t* qlist;      // optional: slot -> batch query
    const uint32_t* thr;        // [nq] by batch query: documents below are dropped
    const uint32_t* seg_dense_off;   // page segments of the dense layout
    const uint32_t* seg_n_real;
    const uint32_t* seg_doc_base;
    uint32_t n_seg;
    // source B (pass 2): keys written by pass 1, [n_slots][cap], in_count[slot] of them
    const uint64_t* keys_in;
    const uint32_t* in_count;
    uint32_t cap;
    // work
    uint32_t n_chunks;          // chunks per slot
    uint32_t* hist;             // [n_slots][n_chunks][256]; after the scan: start offsets
    uint32_t* slot_total;       // [n_slots] entries kept (cut at `limit` in the last pass)
    uint64_t limit;             // last pass only, 0 = all
    // destination: pass 1 -> keys_out[slot * cap ...]; pass 0 / 2 -> keys_out[csr_off[slot] ...],
    // or, with soa != 0, two u32 arrays in the same place: documents [total], then scores [total]
    // (total = csr_off[n_slots]) -- the form the C ABI hands out, so the host copies nothing
    uint64_t* keys_out;
    const uint64_t* csr_off;
    uint32_t n_slots;
    uint32_t soa;
    int pass;                   // 0: one byte, final; 1: low byte of 16-bit counts; 2: high byte, final
};

// element i of a slot: is it kept, its digit bin (0 = largest digit: bins ascend as scores fall)
// and its key
__device__ __forceinline__ bool ds_fetch(const DenseSortParams& p, uint32_t slot, uint32_t i,
                                         uint32_t* bin, uint64_t* key) {
    if (p.pass == 2) {
        if (i >= p.in_count[slot]) return false;
        const uint64_t k = p.keys_in[static_cast<uint64_t>(slot) * p.cap + i];
        *key = k;
        *bin = 255u - ((key_score(k) >> 8) & 0xFFu);
        return true;
    }
    if (i >= p.dense_cols) return false;
    // page segment of column i: the last one starting at or before it
    uint32_t lo = 0, hi = p.n_seg;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.seg_dense_off[mid] <= i) lo = mid;
        else hi = mid;
    }
    const uint32_t rel = i - p.seg_dense_off[lo];
    if (rel >= p.seg_n_real[lo]) return false;          // padded column
    const uint64_t at = static_cast<uint64_t>(slot) * p.dense_pitch + i;
    const uint32_t score = p.pass == 0 ? p.dense8[at] : p.dense16[at];
    const uint32_t q = p.qlist ? p.qlist[slot] : slot;
    if (score < p.thr[q]) return false;
    *key = make_key(score, p.seg_doc_base[lo] + rel);
    *bin = 255u - (score & 0xFFu);
    return true;
}

// grid (n_chunks, n_slots): per-chunk histogram of the digit
__global__ void __launch_bounds__(DS_THREADS) ds_hist_kernel(DenseSortParams p) {
    __shared__ uint32_t h[256];
    const uint32_t slot = blockIdx.y, chunk = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t r = 0; r < DS_CHUNK / DS_THREADS; ++r) {
        const uint32_t i = chunk * DS_CHUNK + r * DS_THREADS + threadIdx.x;
        uint32_t bin = 0;
        uint64_t key;
        const bool keep = ds_fetch(p, slot, i, &bin, &key);
        // warp-aggregated: scores cluster in a few bins
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
            const uint32_t peers = __match_any_sync(act, bin);
            if ((threadIdx.x & 31) == static_cast<uint32_t>(__ffs(peers) - 1)) atomicAdd(&h[bin], __popc(peers));
        }
    }
    __syncthreads();
    p.hist[(static_cast<uint64_t>(slot) * p.n_chunks + chunk) * 256 + threadIdx.x] = h[threadIdx.x];
}

// grid (n_slots): counts -> start offsets in (bin, chunk) order; slot total.
// Thread (g, b) owns bin b over the g-th quarter of the chunks.  The walk over the chunks is a
// latency chain (one small load per chunk), so the loads are issued DS_SCAN_UNROLL at a time:
// with one thread per bin and a load-store-load chain this kernel alone cost ~0.1 ms per pass on
// a million documents -- more than the histogram and the scatter together.
static constexpr uint32_t DS_SCAN_GROUPS = 4;
static constexpr uint32_t DS_SCAN_UNROLL = 16;
__global__ void __launch_bounds__(256 * DS_SCAN_GROUPS) ds_scan_kernel(DenseSortParams p) {
    __shared__ uint32_t group_sum[DS_SCAN_GROUPS][256];
    __shared__ uint32_t bin_base[256];
    __shared__ uint32_t warp_sum[8];
    const uint32_t slot = blockIdx.x, b = threadIdx.x & 255u, g = threadIdx.x >> 8;
    uint32_t* h = p.hist + static_cast<uint64_t>(slot) * p.n_chunks * 256 + b;
    const uint32_t per = (p.n_chunks + DS_SCAN_GROUPS - 1) / DS_SCAN_GROUPS;
    const uint32_t c0 = min(g * per, p.n_chunks), c1 = min(c0 + per, p.n_chunks);
    uint32_t sum = 0;
    for (uint32_t c = c0; c < c1; c += DS_SCAN_UNROLL) {
        uint32_t v[DS_SCAN_UNROLL];
#pragma unroll
        for (uint32_t j = 0; j < DS_SCAN_UNROLL; ++j) v[j] = c + j < c1 ? h[static_cast<uint64_t>(c + j) * 256] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < DS_SCAN_UNROLL; ++j) sum += v[j];
    }
    group_sum[g][b] = sum;
    __syncthreads();
    // exclusive scan of the 256 bin totals (threads of group 0)
    uint32_t run = 0;
    if (g == 0) {
#pragma unroll
        for (uint32_t k = 0; k < DS_SCAN_GROUPS; ++k) run += group_sum[k][b];
        const uint32_t lane = b & 31, warp = b >> 5;
        uint32_t incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sum[warp] = incl;
        bin_base[b] = incl - run;   // (within the warp; the warps before are added below)
    }
    __syncthreads();
    uint32_t base = bin_base[b];
    for (uint32_t w = 0; w < (b >> 5); ++w) base += warp_sum[w];
    if (g == 0 && b == 255) {
        uint64_t total = static_cast<uint64_t>(base) + run;
        if (p.pass != 1 && p.limit != 0 && total > p.limit) total = p.limit;
        p.slot_total[slot] = static_cast<uint32_t>(total);
    }
    // start of (bin b, chunk c0): the bin's base + the groups before this one
    uint32_t at = base;
    for (uint32_t k = 0; k < g; ++k) at += group_sum[k][b];
    for (uint32_t c = c0; c < c1; c += DS_SCAN_UNROLL) {
        uint32_t v[DS_SCAN_UNROLL];
#pragma unroll
        for (uint32_t j = 0; j < DS_SCAN_UNROLL; ++j) v[j] = c + j < c1 ? h[static_cast<uint64_t>(c + j) * 256] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < DS_SCAN_UNROLL; ++j) {
            if (c + j < c1) h[static_cast<uint64_t>(c + j) * 256] = at;
            at += v[j];
        }
    }
}

// grid (n_chunks, n_slots): stable scatter of the chunk's entries to their final positions
__global__ void __launch_bounds__(DS_THREADS) ds_scatter_kernel(DenseSortParams p) {
    __shared__ uint32_t bin_off[256];
    __shared__ uint32_t warp_cnt[DS_THREADS / 32][256];
    const uint32_t slot = blockIdx.y, chunk = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bin_off[threadIdx.x] = p.hist[(static_cast<uint64_t>(slot) * p.n_chunks + chunk) * 256 + threadIdx.x];
    uint64_t* out = p.pass == 1 ? p.keys_out + static_cast<uint64_t>(slot) * p.cap
                                : p.keys_out + p.csr_off[slot];
    const bool soa = p.pass != 1 && p.soa != 0;
    uint32_t* out_doc = reinterpret_cast<uint32_t*>(p.keys_out) + (soa ? p.csr_off[slot] : 0);
    uint32_t* out_score = out_doc + (soa ? p.csr_off[p.n_slots] : 0);
    const uint64_t limit = (p.pass != 1 && p.limit != 0) ? p.limit : ~0ull;
    for (uint32_t r = 0; r < DS_CHUNK / DS_THREADS; ++r) {
        for (uint32_t i = threadIdx.x; i < (DS_THREADS / 32) * 256; i += DS_THREADS) (&warp_cnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t i = chunk * DS_CHUNK + r * DS_THREADS + threadIdx.x;
        uint32_t bin = 0, rank = 0;
        uint64_t key = 0;
        const bool keep = ds_fetch(p, slot, i, &bin, &key);
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
            const uint32_t peers = __match_any_sync(act, bin);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) warp_cnt[warp][bin] = __popc(peers);
        }
        __syncthreads();
        {   // per bin: running offsets across the warps of this round (thread b owns bin b)
            uint32_t run = bin_off[threadIdx.x];
#pragma unroll
            for (uint32_t w = 0; w < DS_THREADS / 32; ++w) {
                const uint32_t c = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += c;
            }
            bin_off[threadIdx.x] = run;
        }
        __syncthreads();
        if (keep) {
            const uint64_t pos = static_cast<uint64_t>(warp_cnt[warp][bin]) + rank;
            if (pos < limit) {
                if (soa) {
                    out_doc[pos] = key_doc(key);
                    out_score[pos] = key_score(key);
                } else {
                    out[pos] = key;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace cobsgpu
