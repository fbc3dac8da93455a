// cobs_b200/csrc/score.cuh -- K2: fused row gather + AND + per-document counting (+ threshold).
//
// Replaces, in ONE kernel, the reference's per-batch pipeline
//   read_from_disk   (cobs/query/classic_index/mmap_search_file.cpp:27-40,
//                     cobs/query/compact_index/mmap_search_file.cpp:34-67)
//   aggregate_rows   (cobs/query/classic_search.cpp:279-307)
//   compute_counts_* (cobs/query/classic_search.cpp:213-275, 643-1022)
// and, in CAND mode, the threshold filter of counts_to_result (classic_search.cpp:121-126).
//
// Design (B200, HBM-bound integer work, no tensor cores):
//   * work item = (query, column tile); a tile is <= W = 512*NCW contiguous bytes of a
//     signature row (4096*NCW documents).  Persistent CTAs pull items from a global atomic
//     counter (the producer warp fetches, consumers follow through a 2-slot shared queue), so
//     no CTA is ever more than one item behind at the end: a straggler CTA alone is
//     latency-bound (~37 GB/s), which made static striding cost ~0.1 ms per launch.
//   * one PRODUCER warp per CTA turns raw hashes into row addresses (hash % signature_size,
//     one modulus per page) and streams the h row slices of every k-mer into a shared-memory
//     ring with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP);
//     the `rows` scratch buffer of the reference is never materialised.
//   * NCW CONSUMER warps: each thread owns 16 bytes = 128 documents of the tile, reads the h
//     slices with 128-bit shared loads, ANDs them, and adds the 128 result bits into
//     BIT-SLICED (vertical) counters: 8 bit-planes per 32-document word, fed through a
//     carry-save adder tree over groups of 8 k-mers (~3 LOP3 per word per k-mer) -- a
//     per-bit extract+add could not keep up with HBM.
//   * epilogue per item: CAND  -> bit-sliced compare against ceil(threshold*T), survivors are
//                                 appended (warp-aggregated atomics) as sort keys;
//                         DENSE8 / DENSE16 -> the 128 counts are transposed out of the planes and
//                                   stored, one byte / one u16 per document (exhaustive lists);
//                         DENSE32-> counts are added into a u32 score vector (flushed before the
//                                   planes could overflow; any query length);
//                         TOPK  -> every consumer warp keeps only the best `topk` of its 4096
//                                  documents under the reference's order (score desc, doc asc):
//                                  a bit-sliced binary search for the cutoff score plus a
//                                  prefix-limited tie mask.  The union of the per-warp lists
//                                  contains the query's global top-k, never overflows, and is
//                                  tiny, so `-t 0 -l k` no longer materialises every document.
//                         KSPLIT -> a work item is (query, tile, CHUNK of the query's k-mers): the
//                                  partial counts of the chunk are stored as bytes in the chunk's
//                                  own vector and summed by ksplit_reduce_kernel.  For a few long
//                                  queries (one gene against the index) the chunks spread one query
//                                  over every SM instead of leaving it to n_tiles CTAs whose k-mers
//                                  form a latency chain.
//   * NP = number of bit-planes per word: 8 (queries of <= 255 k-mers) or 16 (<= 65 535).
#pragma once

#include "common.cuh"

namespace cobsgpu {

enum ScoreMode : int { MODE_CAND = 0, MODE_DENSE8 = 1, MODE_DENSE32 = 2, MODE_TOPK = 3, MODE_DENSE16 = 4,
                       MODE_KSPLIT = 5 };
static constexpr int SCORE_MODES = 6;

// one column tile of one page of the shard held by this device
struct TileDesc {
    const uint8_t* base;  // device address of (row 0, first byte of the tile)
    uint64_t sig;         // rows of the page = modulus for this tile
    uint32_t pitch;       // bytes between consecutive rows of the page
    uint32_t bytes;       // bytes of this tile, multiple of 16, <= W
    uint32_t doc_base;    // global document id of bit 0 of the tile
    uint32_t n_real;      // real documents in the tile counted from doc_base
    uint32_t dense_off;   // first column of the tile in the shard-local dense score layout
    uint32_t pad;
};

struct ScoreParams {
    const TileDesc* tiles;
    uint32_t n_tiles;
    uint32_t h;
    const uint64_t* hashes;   // [total_kmers * h] from K1
    const uint32_t* koff;     // [nq+1] k-mer prefix per query of the batch
    const uint32_t* qlist;    // optional [nq_items]: batch query index per slot; NULL = identity
    uint32_t nq_items;
    uint32_t n_stages;        // ring depth (k-mers in flight per CTA)
    // MODE_CAND
    const uint32_t* thr;      // [nq] ceil(threshold * T_q), indexed by batch query
    uint32_t* cand_count;     // [nq_items]
    uint64_t* cand;           // [nq_items * cap]
    uint32_t cap;
    uint32_t topk;            // MODE_TOPK: results wanted per query (>= 1)
    // MODE_DENSE8 / MODE_DENSE32
    uint8_t* dense8;          // [nq_items * dense_pitch]
    uint16_t* dense16;        // [nq_items * dense_pitch] (16 bit-planes, queries of <= 65 535 k-mers)
    uint32_t* dense32;        // [nq_items * dense_pitch] (zero-initialised)
    uint64_t dense_pitch;     // columns per slot, multiple of 128
    // MODE_KSPLIT: k-mers per chunk (multiple of 8, <= 248) and chunks per query; 1 chunk otherwise
    uint32_t kchunk, n_kchunks;
    // dynamic work distribution: [0] next item, [1] CTAs finished (both zero between launches;
    // the last CTA to finish resets them)
    unsigned long long* work;
};

// shared-memory header of a CTA: full[NS] + empty[NS] + item-queue barriers and slots
__host__ __device__ constexpr uint32_t score_smem_header(uint32_t ns) {
    return ((2 * ns + 6) * 8 + 127) / 128 * 128;
}
static constexpr uint32_t SCORE_ITEM_Q = 2;   // item ids the producer may run ahead

static constexpr int SCORE_MAX_THREADS = 160;   // 4 consumer warps + 1 producer warp
static constexpr int SCORE_PLANES = 8;          // short queries: counts up to 255
static constexpr int SCORE_PLANES_LONG = 16;    // long queries: counts up to 65 535

// full adder over 32 lanes of documents
#define COBS_CSA(sum, carry, a, b, c)            \
    {                                            \
        uint32_t u_ = (a) ^ (b);                 \
        uint32_t c_ = ((a) & (b)) | (u_ & (c));  \
        (sum) = u_ ^ (c);                        \
        (carry) = c_;                            \
    }

// documents of a 32-bit word whose bit-sliced count is >= thr
template <int NP>
__host__ __device__ __forceinline__ uint32_t planes_ge(const uint32_t (&pl)[NP], uint32_t thr) {
    if (thr == 0) return 0xFFFFFFFFu;
    if ((thr >> NP) != 0) return 0u;
    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
        if ((thr >> i) & 1u) {
            eq &= pl[i];
        } else {
            gt |= eq & pl[i];
            eq &= ~pl[i];
        }
    }
    return gt | eq;
}

template <int NP>
__host__ __device__ __forceinline__ uint32_t planes_count(const uint32_t (&pl)[NP], uint32_t bit) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) c |= ((pl[i] >> bit) & 1u) << i;
    return c;
}

// bits [8*B, 8*B+8) of the counts of documents 4g..4g+3 of a word, one per byte (8x4 bit-matrix
// transpose by multiply)
template <int B = 0, int NP>
__host__ __device__ __forceinline__ uint32_t planes_pack4(const uint32_t (&pl)[NP], uint32_t g) {
    static_assert(8 * B + 8 <= NP, "plane byte out of range");
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t nib = (pl[8 * B + i] >> (4 * g)) & 0xFu;
        out |= ((nib * 0x00204081u) & 0x01010101u) << i;
    }
    return out;
}

template <int H, int MODE, int NP>
__global__ void __launch_bounds__(SCORE_MAX_THREADS, 3) score_kernel(const ScoreParams p) {
    static_assert(NP == SCORE_PLANES || NP == SCORE_PLANES_LONG, "8 or 16 bit-planes");
    static_assert(MODE != MODE_DENSE8 || NP == 8, "DENSE8 stores one byte per document");
    static_assert(MODE != MODE_DENSE16 || NP == 16, "DENSE16 stores 16-bit counts");
    static_assert(MODE != MODE_KSPLIT || NP == 8, "a chunk holds at most 248 k-mers");
    constexpr uint32_t MAXC = (1u << NP) - 1u;   // largest count the planes can hold
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t h = H > 0 ? static_cast<uint32_t>(H) : p.h;
    const uint32_t ncw = (blockDim.x >> 5) - 1;   // consumer warps; the last warp produces
    const uint32_t W = ncw * 512;                 // tile width in bytes
    const uint32_t NS = p.n_stages;
    const uint32_t stage_bytes = h * W;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + NS;
    uint64_t* iq_full = empty + NS;                 // item queue: slot published by the producer
    uint64_t* iq_empty = iq_full + SCORE_ITEM_Q;    // slot read by every consumer warp
    volatile uint64_t* item_q = iq_empty + SCORE_ITEM_Q;
    uint8_t* data = smem + score_smem_header(NS);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NS; ++s) {
            mbar_init(&full[s], h);      // one arrive.expect_tx per row slice
            mbar_init(&empty[s], ncw);   // one arrive per consumer warp
        }
        for (uint32_t s = 0; s < SCORE_ITEM_Q; ++s) {
            mbar_init(&iq_full[s], 1);
            mbar_init(&iq_empty[s], ncw);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint64_t n_items = static_cast<uint64_t>(p.nq_items) * p.n_tiles * (MODE == MODE_KSPLIT ? p.n_kchunks : 1u);
    uint32_t s = 0, par = 0;   // ring position / phase parity, advanced identically by both roles
    uint32_t iq = 0, iq_par = 0;   // item-queue position / parity, likewise

    if (warp == ncw) {
        // ------------------------------ producer warp ------------------------------
        uint32_t kpr = 32 / h;             // k-mers issued per round, one lane per row slice
        if (kpr > NS) kpr = NS;
        const uint32_t my_k = lane / h, my_j = lane - my_k * h;
        const bool lane_used = my_k < kpr;
        const uint64_t policy = l2_policy_evict_first();
        for (;;) {
            // next item from the global counter, published to the consumer warps
            unsigned long long item = 0;
            if (lane == 0) {
                mbar_wait(&iq_empty[iq], iq_par ^ 1);
                item = atomicAdd(p.work, 1ull);
                item_q[iq] = item;
                mbar_arrive(&iq_full[iq]);
            }
            item = __shfl_sync(0xFFFFFFFFu, item, 0);
            if (++iq == SCORE_ITEM_Q) {
                iq = 0;
                iq_par ^= 1;
            }
            if (item >= n_items) break;
            uint64_t qt = item;
            uint32_t kc = 0;
            if (MODE == MODE_KSPLIT) {
                qt = item / p.n_kchunks;
                kc = static_cast<uint32_t>(item - qt * p.n_kchunks);
            }
            const uint32_t qi = static_cast<uint32_t>(qt / p.n_tiles);
            const uint32_t tile = static_cast<uint32_t>(qt - static_cast<uint64_t>(qi) * p.n_tiles);
            const uint32_t q = p.qlist ? p.qlist[qi] : qi;
            const TileDesc td = p.tiles[tile];
            uint32_t k0 = p.koff[q], T = p.koff[q + 1] - k0;
            if (MODE == MODE_KSPLIT) {   // this item's chunk of the query's k-mers
                const uint32_t kb = kc * p.kchunk;
                T = kb < T ? (T - kb < p.kchunk ? T - kb : p.kchunk) : 0;
                k0 += kb;
            }
            const uint64_t* hq = p.hashes + static_cast<uint64_t>(k0) * h;
            for (uint32_t t0 = 0; t0 < T; t0 += kpr) {
                const uint32_t t = t0 + my_k;
                uint32_t ms = s + my_k, mpar = par;
                if (ms >= NS) {
                    ms -= NS;
                    mpar ^= 1;
                }
                if (lane_used && t < T) {
                    const uint64_t hv = hq[static_cast<uint64_t>(t) * h + my_j];
                    const uint64_t row = hv % td.sig;   // classic: signature_size; compact: per page
                    const uint8_t* src = td.base + row * td.pitch;
                    mbar_wait(&empty[ms], mpar ^ 1);
                    mbar_arrive_expect_tx(&full[ms], td.bytes);
                    bulk_g2s(data + ms * stage_bytes + my_j * W, src, td.bytes, &full[ms], policy);
                }
                const uint32_t adv = (T - t0 < kpr) ? (T - t0) : kpr;
                s += adv;
                if (s >= NS) {
                    s -= NS;
                    par ^= 1;
                }
                __syncwarp();
            }
        }
        // the last CTA to run out of work re-arms the counters for the next launch
        if (lane == 0) {
            const unsigned long long done = atomicAdd(p.work + 1, 1ull);
            if (done == gridDim.x - 1) {
                p.work[0] = 0;
                p.work[1] = 0;
            }
        }
        return;
    }

    // -------------------------------- consumer warps --------------------------------
    const uint8_t* my = data + threadIdx.x * 16;
    for (;;) {
        mbar_wait(&iq_full[iq], iq_par);
        const uint64_t item = item_q[iq];
        __syncwarp();
        if (lane == 0) mbar_arrive(&iq_empty[iq]);
        if (++iq == SCORE_ITEM_Q) {
            iq = 0;
            iq_par ^= 1;
        }
        if (item >= n_items) break;
        uint64_t qt = item;
        uint32_t kc = 0;
        if (MODE == MODE_KSPLIT) {
            qt = item / p.n_kchunks;
            kc = static_cast<uint32_t>(item - qt * p.n_kchunks);
        }
        const uint32_t qi = static_cast<uint32_t>(qt / p.n_tiles);
        const uint32_t tile = static_cast<uint32_t>(qt - static_cast<uint64_t>(qi) * p.n_tiles);
        const uint32_t q = p.qlist ? p.qlist[qi] : qi;
        const TileDesc td = p.tiles[tile];
        uint32_t T = p.koff[q + 1] - p.koff[q];
        if (MODE == MODE_KSPLIT) {
            const uint32_t kb = kc * p.kchunk;
            T = kb < T ? (T - kb < p.kchunk ? T - kb : p.kchunk) : 0;
        }
        const bool active = threadIdx.x * 16 < td.bytes;

        uint32_t pl[4][NP];
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int i = 0; i < NP; ++i) pl[w][i] = 0;

        // AND of the h row slices of the next k-mer, 128 documents per thread
        auto fetch = [&](uint32_t (&x)[4]) {
            mbar_wait(&full[s], par);
            const uint8_t* src = my + s * stage_bytes;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (active) {
                v = lds128(src);
#pragma unroll 4
                for (uint32_t j = 1; j < h; ++j) {
                    uint4 u = lds128(src + j * W);
                    v.x &= u.x;
                    v.y &= u.y;
                    v.z &= u.z;
                    v.w &= u.w;
                }
            }
            x[0] = v.x;
            x[1] = v.y;
            x[2] = v.z;
            x[3] = v.w;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);   // slot may be refilled
            if (++s == NS) {
                s = 0;
                par ^= 1;
            }
        };

        // transposes the planes into per-document counts: store (DENSE8) or add (DENSE32)
        auto flush_dense = [&]() {
            if (!active) return;
            const uint64_t slot = MODE == MODE_KSPLIT ? static_cast<uint64_t>(qi) * p.n_kchunks + kc : qi;
            const uint64_t col = slot * p.dense_pitch + td.dense_off + threadIdx.x * 128;
            if (MODE == MODE_DENSE8 || MODE == MODE_KSPLIT) {
                uint4* dst = reinterpret_cast<uint4*>(p.dense8 + col);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint4 lo, hi;
                    lo.x = planes_pack4(pl[w], 0);
                    lo.y = planes_pack4(pl[w], 1);
                    lo.z = planes_pack4(pl[w], 2);
                    lo.w = planes_pack4(pl[w], 3);
                    hi.x = planes_pack4(pl[w], 4);
                    hi.y = planes_pack4(pl[w], 5);
                    hi.z = planes_pack4(pl[w], 6);
                    hi.w = planes_pack4(pl[w], 7);
                    dst[2 * w] = lo;
                    dst[2 * w + 1] = hi;
                }
            } else if (MODE == MODE_DENSE16) {
                uint4* dst = reinterpret_cast<uint4*>(p.dense16 + col);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
#pragma unroll
                    for (int gp = 0; gp < 4; ++gp) {
                        // low and high count bytes of documents 8gp..8gp+7, interleaved to u16
                        const uint32_t lo0 = planes_pack4(pl[w], 2 * gp);
                        const uint32_t hi0 = planes_pack4<(NP > 8 ? 1 : 0)>(pl[w], 2 * gp);
                        const uint32_t lo1 = planes_pack4(pl[w], 2 * gp + 1);
                        const uint32_t hi1 = planes_pack4<(NP > 8 ? 1 : 0)>(pl[w], 2 * gp + 1);
                        uint4 v;
                        v.x = __byte_perm(lo0, hi0, 0x5140);
                        v.y = __byte_perm(lo0, hi0, 0x7362);
                        v.z = __byte_perm(lo1, hi1, 0x5140);
                        v.w = __byte_perm(lo1, hi1, 0x7362);
                        dst[4 * w + gp] = v;
                    }
                }
            } else {
                uint4* dst = reinterpret_cast<uint4*>(p.dense32 + col);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const uint32_t pk = planes_pack4(pl[w], g);
                        uint32_t ph = 0;   // (DENSE32 is instantiated with 8 planes: ph stays 0)
                        if (NP > 8) ph = planes_pack4<(NP > 8 ? 1 : 0)>(pl[w], g);
                        uint4 a = dst[8 * w + g];
                        a.x += (pk & 0xFFu) | ((ph & 0xFFu) << 8);
                        a.y += ((pk >> 8) & 0xFFu) | (((ph >> 8) & 0xFFu) << 8);
                        a.z += ((pk >> 16) & 0xFFu) | (((ph >> 16) & 0xFFu) << 8);
                        a.w += (pk >> 24) | ((ph >> 24) << 8);
                        dst[8 * w + g] = a;
                    }
                }
            }
        };
        auto clear_planes = [&]() {
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int i = 0; i < NP; ++i) pl[w][i] = 0;
        };

        uint32_t t = 0, acc = 0;   // acc = k-mers held in the planes since the last flush
        // groups of 8 k-mers through a carry-save adder tree: planes 0..2 are the tree's
        // ones/twos/fours, the carry of weight 8 ripples into planes 3..NP-1
        for (; t + 8 <= T; t += 8) {
            if (MODE == MODE_DENSE32 && acc + 8 > MAXC) {
                flush_dense();
                clear_planes();
                acc = 0;
            }
            uint32_t xa[4], xb[4], t2a[4], t2b[4], t4a[4], t4b[4];
            fetch(xa);
            fetch(xb);
#pragma unroll
            for (int w = 0; w < 4; ++w) COBS_CSA(pl[w][0], t2a[w], pl[w][0], xa[w], xb[w]);
            fetch(xa);
            fetch(xb);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                COBS_CSA(pl[w][0], t2b[w], pl[w][0], xa[w], xb[w]);
                COBS_CSA(pl[w][1], t4a[w], pl[w][1], t2a[w], t2b[w]);
            }
            fetch(xa);
            fetch(xb);
#pragma unroll
            for (int w = 0; w < 4; ++w) COBS_CSA(pl[w][0], t2a[w], pl[w][0], xa[w], xb[w]);
            fetch(xa);
            fetch(xb);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t t8;
                COBS_CSA(pl[w][0], t2b[w], pl[w][0], xa[w], xb[w]);
                COBS_CSA(pl[w][1], t4b[w], pl[w][1], t2a[w], t2b[w]);
                COBS_CSA(pl[w][2], t8, pl[w][2], t4a[w], t4b[w]);
#pragma unroll
                for (int i = 3; i < 8; ++i) {
                    const uint32_t n = pl[w][i] & t8;
                    pl[w][i] ^= t8;
                    t8 = n;
                }
                // planes 8..15 only see a carry when a count crosses a multiple of 256: rare
                if (NP > 8 && t8 != 0) {
#pragma unroll
                    for (int i = 8; i < NP; ++i) {
                        const uint32_t n = pl[w][i] & t8;
                        pl[w][i] ^= t8;
                        t8 = n;
                    }
                }
            }
            acc += 8;
        }
        // tail (< 8 k-mers): plain ripple-carry increment
        for (; t < T; ++t) {
            if (MODE == MODE_DENSE32 && acc + 1 > MAXC) {
                flush_dense();
                clear_planes();
                acc = 0;
            }
            uint32_t x[4];
            fetch(x);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t c = x[w];
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const uint32_t n = pl[w][i] & c;
                    pl[w][i] ^= c;
                    c = n;
                }
            }
            acc += 1;
        }

        if (MODE == MODE_KSPLIT) {
            // partial counts of this chunk (<= 248 fit a byte) go to the chunk's own dense8 vector,
            // slot qi * n_kchunks + kc; ksplit_reduce_kernel adds the chunks up afterwards.  (A
            // first version added into one u16 vector with packed atomics: 20 M atomics for one
            // 10 000-k-mer query on a million documents cost more than the row traffic.)
            if (T != 0) flush_dense();
        } else if (MODE == MODE_DENSE8 || MODE == MODE_DENSE16 || MODE == MODE_DENSE32) {
            flush_dense();
        } else {
            // threshold in bit-sliced form, then append the survivors as sort keys
            const uint32_t thr = p.thr[q];
            uint32_t m[4], valid[4], cnt = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t d0 = threadIdx.x * 128 + w * 32;   // first document of the word
                const uint32_t n = td.n_real > d0 ? td.n_real - d0 : 0;   // real ones (no padding)
                valid[w] = !active ? 0u : (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
            }
            if (MODE == MODE_TOPK) {
                // Per-warp top-k under (score desc, doc asc).  Documents of a warp are ordered
                // (lane, word, bit) == ascending id.  count_ge(t) = documents of this warp with
                // score >= t; the cutoff c is the largest t >= thr with count_ge(t) >= k: keep
                // every document above c and the first k - count_ge(c+1) documents equal to c.
                const uint32_t k = p.topk;
                auto count_ge = [&](uint32_t t) {
                    uint32_t c = 0;
#pragma unroll
                    for (int w = 0; w < 4; ++w) c += __popc(planes_ge(pl[w], t) & valid[w]);
                    return __reduce_add_sync(0xFFFFFFFFu, c);
                };
                uint32_t lo = thr;
                if (count_ge(lo) <= k) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) m[w] = planes_ge(pl[w], lo) & valid[w];
                } else {
                    uint32_t hi = (T < MAXC ? T : MAXC) + 1;   // count_ge(hi) == 0: scores <= T
                    while (hi - lo > 1) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if (count_ge(mid) >= k) lo = mid;
                        else hi = mid;
                    }
                    uint32_t eq[4], n_gt = 0, n_eq = 0;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        m[w] = planes_ge(pl[w], lo + 1) & valid[w];
                        eq[w] = planes_ge(pl[w], lo) & valid[w] & ~m[w];
                        n_gt += __popc(m[w]);
                        n_eq += __popc(eq[w]);
                    }
                    const uint32_t need = k - __reduce_add_sync(0xFFFFFFFFu, n_gt);   // >= 1
                    uint32_t pre = n_eq;   // exclusive prefix of the tie counts over the lanes
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, pre, d);
                        if (lane >= d) pre += o;
                    }
                    pre -= n_eq;
                    uint32_t allow = need > pre ? need - pre : 0;   // ties this lane may keep
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t c = __popc(eq[w]);
                        if (c > allow) {
                            uint32_t rest = eq[w], kept = 0;
                            for (uint32_t i = 0; i < allow; ++i) {
                                kept |= rest & (0u - rest);
                                rest &= rest - 1;
                            }
                            eq[w] = kept;
                            allow = 0;
                        } else {
                            allow -= c;
                        }
                        m[w] |= eq[w];
                    }
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) cnt += __popc(m[w]);
            } else {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    m[w] = planes_ge(pl[w], thr) & valid[w];
                    cnt += __popc(m[w]);
                }
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total != 0) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&p.cand_count[qi], total);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                uint32_t pos = base + incl - cnt;
                uint64_t* out = p.cand + static_cast<uint64_t>(qi) * p.cap;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t mm = m[w];
                    while (mm) {
                        const uint32_t b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        if (pos < p.cap)
                            out[pos] = make_key(planes_count(pl[w], b),
                                                td.doc_base + threadIdx.x * 128 + w * 32 + b);
                        ++pos;
                    }
                }
            }
        }
    }
}

#undef COBS_CSA

// Sums the per-chunk byte counts of MODE_KSPLIT into the u16 score vector of each query:
// out16[q][c] = sum over the chunks kc < ceil(T_q / kchunk) of part8[(q * n_kchunks + kc)][c].
// grid (dense_pitch / (256 * 16), n_slots); a thread owns 16 columns.
struct KsplitReduceParams {
    const uint8_t* part8;     // [n_slots * n_kchunks][dense_pitch]
    uint16_t* out16;          // [n_slots][dense_pitch]
    uint64_t dense_pitch;     // multiple of 128
    const uint32_t* koff;     // [nq + 1]
    const uint32_t* qlist;    // optional: slot -> batch query
    uint32_t kchunk, n_kchunks;
    // optional fused threshold (threshold > 0): documents with score >= thr[q] are appended as
    // sort keys right here, so a batch with few hits never needs the counting sort over all
    // documents; cand == nullptr switches it off
    const uint32_t* thr;             // [nq] by batch query
    const uint32_t* seg_dense_off;   // page segments of the dense layout
    const uint32_t* seg_n_real;
    const uint32_t* seg_doc_base;
    uint32_t n_seg;
    uint32_t* cand_count;            // [n_slots]
    uint64_t* cand;                  // [n_slots * cap]
    uint32_t cap;
};

__global__ void __launch_bounds__(256) ksplit_reduce_kernel(KsplitReduceParams p) {
    const uint32_t slot = blockIdx.y;
    const uint64_t c0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
    // (threads past the end stay for the warp-wide append below)
    const bool in_range = c0 < p.dense_pitch;
    const uint32_t q = p.qlist ? p.qlist[slot] : slot;
    const uint32_t T = p.koff[q + 1] - p.koff[q];
    const uint32_t chunks = (T + p.kchunk - 1) / p.kchunk;   // chunks this query really has
    uint32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    if (in_range) {
        for (uint32_t kc = 0; kc < chunks && kc < p.n_kchunks; ++kc) {
            const uint4 v = *reinterpret_cast<const uint4*>(
                p.part8 + (static_cast<uint64_t>(slot) * p.n_kchunks + kc) * p.dense_pitch + c0);
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        }
        uint4 lo, hi;
        lo.x = acc[0] | (acc[1] << 16);
        lo.y = acc[2] | (acc[3] << 16);
        lo.z = acc[4] | (acc[5] << 16);
        lo.w = acc[6] | (acc[7] << 16);
        hi.x = acc[8] | (acc[9] << 16);
        hi.y = acc[10] | (acc[11] << 16);
        hi.z = acc[12] | (acc[13] << 16);
        hi.w = acc[14] | (acc[15] << 16);
        uint4* dst = reinterpret_cast<uint4*>(p.out16 + static_cast<uint64_t>(slot) * p.dense_pitch + c0);
        dst[0] = lo;
        dst[1] = hi;
    }
    if (p.cand == nullptr) return;
    uint32_t mask = 0, doc0 = 0;
    if (in_range) {
        // the 16 columns of a thread lie in one page segment (segments start at multiples of 128)
        uint32_t lo_s = 0, hi_s = p.n_seg;
        while (hi_s - lo_s > 1) {
            const uint32_t mid = (lo_s + hi_s) >> 1;
            if (p.seg_dense_off[mid] <= c0) lo_s = mid;
            else hi_s = mid;
        }
        const uint32_t rel0 = static_cast<uint32_t>(c0) - p.seg_dense_off[lo_s];
        const uint32_t n_real = p.seg_n_real[lo_s];
        doc0 = p.seg_doc_base[lo_s] + rel0;
        const uint32_t thr = p.thr[q];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (rel0 + i < n_real && acc[i] >= thr) mask |= 1u << i;
    }
    // warp-aggregated append
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&p.cand_count[slot], total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    uint32_t pos = base + incl - cnt;
    uint64_t* out = p.cand + static_cast<uint64_t>(slot) * p.cap;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((mask >> i) & 1u) {
            if (pos < p.cap) out[pos] = make_key(acc[i], doc0 + i);
            ++pos;
        }
    }
}

}  // namespace cobsgpu
