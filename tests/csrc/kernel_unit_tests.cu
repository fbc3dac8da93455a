// kernel_unit_tests.cu -- CPU-side unit tests of the __host__ __device__ arithmetic the kernels
// are built from (compiled by nvcc, executed on the host, no GPU needed): XXH64 in its three
// forms (byte getter, word-packed fixed length, canonicalising k-mer front-ends), the bit-sliced
// counter helpers of the score kernel, the sort-key encoding and the procedural fill function.
// The expected values come from the oracle (oracle/liboracle.so, test infrastructure).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../cobs_b200/csrc/common.cuh"
#include "../../cobs_b200/csrc/hash.cuh"
#include "../../cobs_b200/csrc/merge_runs.hpp"
#include "../../cobs_b200/csrc/score.cuh"
#include "../../oracle/cobs_oracle.h"

using namespace cobsgpu;

static int g_failed = 0;
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++g_failed;                                                                  \
        }                                                                                \
    } while (0)

template <int K>
static void check_fixed(std::mt19937_64& rng, int canon) {
    static const char alphabet[] = "ACGTACGTACGTN";
    for (int it = 0; it < 400; ++it) {
        uint8_t s[K + 1];
        const int na = it % 5 == 0 ? 13 : 12;   // every fifth k-mer may contain an N
        for (int i = 0; i < K; ++i) s[i] = alphabet[rng() % na];
        uint64_t a[3] = { 0, 0, 0 }, b[3] = { 1, 1, 1 };
        const bool ga = hash_kmer<0>(s, K, 3, canon, [&](uint32_t j, uint64_t v) { a[j] = v; });
        uint8_t c[K];
        std::memcpy(c, s, K);
        const uint8_t(&cr)[K] = c;
        const bool gb = hash_kmer_fixed<K>(cr, 3, canon, [&](uint32_t j, uint64_t v) { b[j] = v; });
        CHECK(ga == gb);
        // oracle
        char buf[K + 1];
        int good = 1;
        if (canon) good = oracle_canonicalize_kmer(reinterpret_cast<const char*>(s), buf, K);
        else std::memcpy(buf, s, K);
        CHECK((good != 0) == ga);
        if (ga)
            for (int j = 0; j < 3; ++j) {
                const uint64_t want = oracle_xxh64(buf, K, j);
                CHECK(a[j] == want);
                CHECK(b[j] == want);
            }
    }
}


// merge_runs (group path, lists of every document): random partitions of a scored document set
// into shards -- contiguous blocks, interleaved blocks, single documents -- merged back and
// compared with a plain sort under the reference order (score desc, document asc)
static void check_merge_runs(std::mt19937_64& rng) {
    for (int it = 0; it < 300; ++it) {
        const uint32_t n_docs = it < 10 ? it : 1 + rng() % 5000;
        const uint32_t n_lists = 1 + rng() % 16;
        const uint32_t max_score = it % 3 == 0 ? 1 : (it % 3 == 1 ? 7 : 70000);
        const uint32_t block = it % 4 == 0 ? 1 : 1 + rng() % 700;   // documents per column block
        std::vector<uint64_t> all;
        std::vector<std::vector<uint64_t>> part(n_lists);
        for (uint32_t d = 0; d < n_docs; ++d) {
            const uint64_t k = make_key(static_cast<uint32_t>(rng() % (max_score + 1)), d * 3 + 1);
            all.push_back(k);
            const uint32_t owner = it % 2 ? (d / block) % n_lists : static_cast<uint32_t>(static_cast<uint64_t>(d) * n_lists / n_docs);
            part[owner].push_back(k);
        }
        std::sort(all.begin(), all.end());
        std::vector<std::vector<uint32_t>> docs(n_lists), scores(n_lists);
        const uint32_t* dp[16];
        const uint32_t* sp[16];
        uint64_t len[16];
        for (uint32_t g = 0; g < n_lists; ++g) {
            std::sort(part[g].begin(), part[g].end());
            for (uint64_t k : part[g]) {
                docs[g].push_back(key_doc(k));
                scores[g].push_back(key_score(k));
            }
            dp[g] = docs[g].data();
            sp[g] = scores[g].data();
            len[g] = part[g].size();
        }
        for (uint64_t want : { static_cast<uint64_t>(n_docs), static_cast<uint64_t>(n_docs / 3), static_cast<uint64_t>(1), static_cast<uint64_t>(0) }) {
            want = std::min<uint64_t>(want, n_docs);
            std::vector<uint32_t> od(want + 1, 0xABABABABu), os(want + 1, 0xABABABABu);
            merge_runs(n_lists, dp, sp, len, want, od.data(), os.data());
            bool same = od[want] == 0xABABABABu && os[want] == 0xABABABABu;   // nothing written beyond
            for (uint64_t i = 0; i < want && same; ++i)
                same = od[i] == key_doc(all[i]) && os[i] == key_score(all[i]);
            CHECK(same);
        }
    }
}

int main() {
    std::mt19937_64 rng(12345);
    check_merge_runs(rng);

    // XXH64 through the byte getter: all lengths 0..200, random seeds
    for (uint32_t len = 0; len <= 200; ++len) {
        std::vector<uint8_t> d(len + 1);
        for (auto& x : d) x = static_cast<uint8_t>(rng());
        const uint64_t seed = rng();
        const uint64_t got = xxh::hash64([&](uint32_t i) { return d[i]; }, len, seed);
        CHECK(got == oracle_xxh64(d.data(), len, seed));
    }
    // the known answer of xxhsum.c:463 (empty input, seed 0)
    CHECK(xxh::hash64([](uint32_t) { return uint8_t(0); }, 0, 0) == 0xEF46DB3751D8E999ULL);

    // k-mer front-ends: generic (run-time k) == fixed (word-packed) == oracle
    for (int canon = 0; canon < 2; ++canon) {
        check_fixed<31>(rng, canon);
        check_fixed<32>(rng, canon);
        check_fixed<15>(rng, canon);
        check_fixed<64>(rng, canon);
        check_fixed<1>(rng, canon);
    }

    // bit-sliced counters: planes hold per-document counts of 32 documents
    for (int it = 0; it < 2000; ++it) {
        uint32_t cnt[32], pl[SCORE_PLANES] = { 0 };
        for (int d = 0; d < 32; ++d) {
            cnt[d] = static_cast<uint32_t>(rng() % 256);
            for (int i = 0; i < SCORE_PLANES; ++i) pl[i] |= ((cnt[d] >> i) & 1u) << d;
        }
        for (uint32_t thr : { 0u, 1u, 2u, 56u, 70u, 127u, 128u, 255u, 256u, 1000u,
                              static_cast<uint32_t>(rng() % 256) }) {
            uint32_t want = 0;
            for (int d = 0; d < 32; ++d) want |= (cnt[d] >= thr ? 1u : 0u) << d;
            CHECK(planes_ge(pl, thr) == want);
        }
        for (uint32_t d = 0; d < 32; ++d) CHECK(planes_count(pl, d) == cnt[d]);
        for (uint32_t g = 0; g < 8; ++g) {
            const uint32_t pk = planes_pack4(pl, g);
            for (uint32_t b = 0; b < 4; ++b) CHECK(((pk >> (8 * b)) & 0xFFu) == cnt[4 * g + b]);
        }
    }

    // sort keys: ascending key order == (score descending, document ascending)
    for (int it = 0; it < 5000; ++it) {
        const uint32_t s1 = rng() % 300, s2 = rng() % 300, d1 = rng(), d2 = rng();
        const uint64_t k1 = make_key(s1, d1), k2 = make_key(s2, d2);
        const bool before = s1 != s2 ? s1 > s2 : d1 < d2;
        if (s1 != s2 || d1 != d2) CHECK((k1 < k2) == before);
        CHECK(key_score(k1) == s1 && key_doc(k1) == d1);
        CHECK(k1 < KEY_PAD);
    }

    // procedural index bits: device fill function == oracle_fill_word
    for (int it = 0; it < 5000; ++it) {
        const uint64_t seed = rng(), row = rng() % (1ull << 40), word = rng() % (1ull << 20);
        const uint32_t page = rng() % 100;
        CHECK(fill_word_from_key(fill_row_key(seed, page, row), word) ==
              oracle_fill_word(seed, page, row, word));
    }

    // shared-memory header of the score kernel: barriers + item queue fit, 128-byte aligned
    for (uint32_t ns = 1; ns <= 64; ++ns) {
        CHECK(score_smem_header(ns) % 128 == 0);
        CHECK(score_smem_header(ns) >= (2 * ns + 2 * SCORE_ITEM_Q + SCORE_ITEM_Q) * 8);
    }

    std::printf("kernel_unit_tests: %s (%d failed checks)\n", g_failed ? "FAILED" : "ok", g_failed);
    return g_failed ? 1 : 0;
}
