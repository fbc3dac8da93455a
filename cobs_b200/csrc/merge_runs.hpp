// merge_runs.hpp -- host-side merge of the shards' ordered lists of one query (group path,
// lists of every document).  Plain C++ so that the unit tests can run it without a GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace cobsgpu {

static constexpr uint32_t MERGE_RUNS_MAX_LISTS = 16;

// Merges the ordered lists of one query held by several shards (list g = n[g] entries at
// doc[g] / score[g]) into out_doc / out_score, at most `want` entries, under the global order
// (score descending, document ascending).  Shards hold disjoint document ranges -- whole column
// blocks per page -- so the merged list is made of long runs out of one shard each: the shard with
// the smallest head is copied for as long as it stays below the second smallest head.
inline void merge_runs(uint32_t n_lists, const uint32_t* const* doc, const uint32_t* const* score,
                const uint64_t* n, uint64_t want, uint32_t* out_doc, uint32_t* out_score) {
    uint64_t at[MERGE_RUNS_MAX_LISTS] = {};
    uint64_t done = 0;
    while (done < want) {
        uint32_t a = n_lists;
        uint64_t ka = ~0ull, kb = ~0ull;   // smallest and second smallest head keys
        for (uint32_t g = 0; g < n_lists; ++g) {
            if (at[g] >= n[g]) continue;
            const uint64_t k = make_key(score[g][at[g]], doc[g][at[g]]);
            if (a == n_lists || k < ka) {
                kb = ka;
                ka = k;
                a = g;
            } else if (k < kb) {
                kb = k;
            }
        }
        if (a == n_lists) break;
        const uint32_t* d = doc[a];
        const uint32_t* sc = score[a];
        uint64_t i = at[a];
        const uint64_t end = std::min<uint64_t>(n[a], i + (want - done));
        const uint64_t first = i;
        // (the head itself is below kb by construction)
        do ++i; while (i < end && make_key(sc[i], d[i]) < kb);
        std::memcpy(out_doc + done, d + first, (i - first) * 4);
        std::memcpy(out_score + done, sc + first, (i - first) * 4);
        done += i - first;
        at[a] = i;
    }
}

}  // namespace cobsgpu
