// cobs_b200/csrc/hash.cuh -- K1: canonicalise + XXH64 every k-mer of a query batch on-device.
//
// Replaces create_hashes (cobs/query/classic_search.cpp:66-107), canonicalize_kmer
// (cobs/util/query.cpp:143-199) and XXH64 (extlib/xxhash/xxhash.c:665-878, v0.6.5) of the
// reference.  One thread per (query, k-mer); the h seeds 0..h-1 are evaluated by the same
// thread.  Output: raw 64-bit hashes, k-mer-major ([kmer][j]); the modulo by the signature
// size is deferred to the score kernel because compact indices use one modulus per page.
#pragma once

#include "common.cuh"

namespace cobsgpu {

struct HashParams {
    const char* queries;      // device: concatenated ASCII queries
    const uint64_t* qoff;     // device [nq+1]: byte offset of each query in `queries`
    const uint32_t* koff;     // device [nq+1]: prefix sum of k-mers per query
    uint32_t nq;
    uint32_t total_kmers;
    uint32_t uniform_T;       // != 0: every query has exactly this many k-mers
    uint32_t k;               // term_size
    uint32_t h;               // num_hashes
    uint32_t canonicalize;
    uint64_t* hashes;         // device [total_kmers * h]
    int* first_bad;           // device: min query index holding a non-ACGT base (canonical only)
    uint32_t* bad;            // device [nq], armed to 0x7F7F7F7F: set to 1 for such queries
};

namespace xxh {
static constexpr uint64_t P1 = 11400714785074694791ULL;
static constexpr uint64_t P2 = 14029467366897019727ULL;
static constexpr uint64_t P3 = 1609587929392839161ULL;
static constexpr uint64_t P4 = 9650029242287828579ULL;
static constexpr uint64_t P5 = 2870177450012600261ULL;

__host__ __device__ __forceinline__ uint64_t rotl(uint64_t x, int r) {
    return (x << r) | (x >> (64 - r));
}
__host__ __device__ __forceinline__ uint64_t round(uint64_t acc, uint64_t input) {
    acc += input * P2;
    acc = rotl(acc, 31);
    return acc * P1;
}
__host__ __device__ __forceinline__ uint64_t merge(uint64_t acc, uint64_t val) {
    acc ^= round(0, val);
    return acc * P1 + P4;
}

// XXH64 of `len` bytes produced by get(i), i = 0..len-1 (little-endian lane assembly).
template <typename Get>
__host__ __device__ __forceinline__ uint64_t hash64(Get get, uint32_t len, uint64_t seed) {
    auto rd64 = [&](uint32_t p) {
        uint64_t v = 0;
#pragma unroll
        for (int b = 7; b >= 0; --b) v = (v << 8) | static_cast<uint64_t>(get(p + b));
        return v;
    };
    uint32_t p = 0;
    uint64_t hsh;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = round(v1, rd64(p));
            v2 = round(v2, rd64(p + 8));
            v3 = round(v3, rd64(p + 16));
            v4 = round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= len);
        hsh = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        hsh = merge(hsh, v1);
        hsh = merge(hsh, v2);
        hsh = merge(hsh, v3);
        hsh = merge(hsh, v4);
    } else {
        hsh = seed + P5;
    }
    hsh += len;
    while (p + 8 <= len) {
        hsh ^= round(0, rd64(p));
        hsh = rotl(hsh, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= len) {
        uint64_t v = 0;
#pragma unroll
        for (int b = 3; b >= 0; --b) v = (v << 8) | static_cast<uint64_t>(get(p + b));
        hsh ^= v * P1;
        hsh = rotl(hsh, 23) * P2 + P3;
        p += 4;
    }
    while (p < len) {
        hsh ^= static_cast<uint64_t>(get(p)) * P5;
        hsh = rotl(hsh, 11) * P1;
        ++p;
    }
    hsh ^= hsh >> 33;
    hsh *= P2;
    hsh ^= hsh >> 29;
    hsh *= P3;
    hsh ^= hsh >> 32;
    return hsh;
}
}  // namespace xxh

namespace xxh {
// XXH64 of LEN bytes already packed little-endian into 64-bit words (compile-time length: every
// lane read is a register pick, the whole function unrolls to ~150 instructions per seed)
template <int LEN>
__host__ __device__ __forceinline__ uint64_t hash64_words(const uint64_t* w, uint64_t seed) {
    int p = 0;
    uint64_t hsh;
    if (LEN >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
#pragma unroll
        for (; p + 32 <= LEN; p += 32) {
            v1 = round(v1, w[p / 8]);
            v2 = round(v2, w[p / 8 + 1]);
            v3 = round(v3, w[p / 8 + 2]);
            v4 = round(v4, w[p / 8 + 3]);
        }
        hsh = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        hsh = merge(hsh, v1);
        hsh = merge(hsh, v2);
        hsh = merge(hsh, v3);
        hsh = merge(hsh, v4);
    } else {
        hsh = seed + P5;
    }
    hsh += static_cast<uint64_t>(LEN);
#pragma unroll
    for (; p + 8 <= LEN; p += 8) {
        hsh ^= round(0, w[p / 8]);
        hsh = rotl(hsh, 27) * P1 + P4;
    }
    if (p + 4 <= LEN) {
        hsh ^= ((w[p / 8] >> (8 * (p % 8))) & 0xFFFFFFFFull) * P1;
        hsh = rotl(hsh, 23) * P2 + P3;
        p += 4;
    }
#pragma unroll
    for (; p < LEN; ++p) {
        hsh ^= ((w[p / 8] >> (8 * (p % 8))) & 0xFFull) * P5;
        hsh = rotl(hsh, 11) * P1;
    }
    hsh ^= hsh >> 33;
    hsh *= P2;
    hsh ^= hsh >> 29;
    hsh *= P3;
    hsh ^= hsh >> 32;
    return hsh;
}
}  // namespace xxh

// A C G T -> themselves, everything else -> 0
__host__ __device__ __forceinline__ uint8_t base_fwd(uint8_t c) {
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : 0;
}
// complement, everything else -> 0
__host__ __device__ __forceinline__ uint8_t base_rev(uint8_t c) {
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 0;
}

// Hashes of one k-mer at `s`.  Returns false when canonicalising and a non-ACGT base occurs.
// KT > 0 fixes the k-mer length at compile time (the bytes then live in registers and are
// shared by the h seeds); KT == 0 takes it from `k_rt`.
template <int KT, typename Emit>
__host__ __device__ __forceinline__ bool hash_kmer(const uint8_t* s, uint32_t k_rt, uint32_t h,
                                                   uint32_t canonicalize, Emit emit) {
    const uint32_t k = KT > 0 ? static_cast<uint32_t>(KT) : k_rt;
    if (!canonicalize) {
        for (uint32_t j = 0; j < h; ++j)
            emit(j, xxh::hash64([&](uint32_t i) { return s[i]; }, k, j));
        return true;
    }
    // the lexicographically smaller of (k-mer, reverse complement): first differing
    // position from the outside in decides; a tie over the whole first half keeps the
    // forward strand (cobs/util/query.cpp:155-198)
    bool reverse = false, decided = false, good = true;
#pragma unroll
    for (uint32_t i = 0; i < k / 2; ++i) {
        const uint8_t f = base_fwd(s[i]), r = base_rev(s[k - 1 - i]);
        if (!decided && f != r) {
            reverse = f > r;
            decided = true;
        }
    }
#pragma unroll
    for (uint32_t i = 0; i < k; ++i) good = good && base_fwd(s[i]) != 0;
    if (!good) return false;
    for (uint32_t j = 0; j < h; ++j)
        emit(j, xxh::hash64([&](uint32_t i) { return reverse ? base_rev(s[k - 1 - i]) : s[i]; },
                            k, j));
    return true;
}

// Fixed-length fast path: validity, strand choice and the canonical bytes are computed once,
// packed into 64-bit words, and every seed hashes the words.
//   valid base  <=> bit (c - 'A') of 0x80045 (A, C, G, T)
//   complement   =  c ^ 0x15 ^ (bit1(c) * 0x11)      (A<->T differ by 0x15, C<->G by 0x04)
template <int KT, typename Emit>
__host__ __device__ __forceinline__ bool hash_kmer_fixed(const uint8_t (&b)[KT], uint32_t h,
                                                uint32_t canonicalize, Emit emit) {
    constexpr int NW = (KT + 7) / 8;
    uint64_t w[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = 0;
    if (canonicalize) {
        bool good = true;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const uint32_t idx = static_cast<uint32_t>(b[i]) - 'A';
            good = good && idx < 20 && ((0x80045u >> idx) & 1u);
        }
        if (!good) return false;
        auto comp = [](uint8_t c) -> uint8_t {
            return static_cast<uint8_t>(c ^ 0x15 ^ (((c >> 1) & 1) * 0x11));
        };
        bool reverse = false, decided = false;
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) {
            const uint8_t f = b[i], r = comp(b[KT - 1 - i]);
            if (!decided && f != r) {
                reverse = f > r;
                decided = true;
            }
        }
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const uint8_t c = reverse ? comp(b[KT - 1 - i]) : b[i];
            w[i / 8] |= static_cast<uint64_t>(c) << (8 * (i % 8));
        }
    } else {
#pragma unroll
        for (int i = 0; i < KT; ++i) w[i / 8] |= static_cast<uint64_t>(b[i]) << (8 * (i % 8));
    }
    for (uint32_t j = 0; j < h; ++j) emit(j, xxh::hash64_words<KT>(w, j));
    return true;
}

template <int KT>
__global__ void __launch_bounds__(128) hash_kmers_kernel(HashParams p) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.total_kmers) return;
    // query owning k-mer gid: largest q with koff[q] <= gid
    uint32_t q, t;
    if (p.uniform_T) {
        q = gid / p.uniform_T;
        t = gid - q * p.uniform_T;
    } else {
        uint32_t lo = 0, hi = p.nq;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.koff[mid] <= gid) lo = mid;
            else hi = mid;
        }
        q = lo;
        t = gid - p.koff[q];
    }
    const uint8_t* s = reinterpret_cast<const uint8_t*>(p.queries) + p.qoff[q] + t;
    uint64_t* out = p.hashes + static_cast<uint64_t>(gid) * p.h;
    bool good;
    if (KT > 0) {
        // stage the bytes in registers once; the canonical form and all h seeds reuse them
        uint8_t b[KT > 0 ? KT : 1];
#pragma unroll
        for (int i = 0; i < KT; ++i) b[i] = s[i];
        good = hash_kmer_fixed<(KT > 0 ? KT : 1)>(
            b, p.h, p.canonicalize, [&](uint32_t j, uint64_t v) { out[j] = v; });
    } else {
        good = hash_kmer<0>(s, p.k, p.h, p.canonicalize,
                            [&](uint32_t j, uint64_t v) { out[j] = v; });
    }
    if (!good) {
        atomicMin(p.first_bad, static_cast<int>(q));
        p.bad[q] = 1;
    }
}

}  // namespace cobsgpu
