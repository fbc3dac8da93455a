// cobs_b200/csrc/fill.cuh -- procedural signature bits for synthetic benchmark indices.
//
// the full-size configurations (>= 100 GB of matrix) cannot be built or
// even held by the host, so their bits are a pure function of (seed, page, row, column word)
// that the device fill kernel and the CPU oracle (oracle_fill_word) evaluate identically.
// Density 1/4, close to `cobs classic-construct-random` (0.249).
#pragma once

#include "common.cuh"

namespace cobsgpu {

struct FillParams {
    uint8_t* base;          // device page base
    uint64_t sig;           // rows
    uint32_t pitch;         // bytes between rows (multiple of 16)
    uint32_t row_bytes;     // bytes of the source row held by this shard
    uint64_t byte_begin;    // first source-row byte held (multiple of 16)
    uint64_t seed;
    uint32_t page;          // GLOBAL page index
};

__global__ void __launch_bounds__(256) fill_page_kernel(FillParams p) {
    const uint32_t wpr = p.pitch / 8;
    for (uint64_t row = blockIdx.x; row < p.sig; row += gridDim.x) {
        const uint64_t key = fill_row_key(p.seed, p.page, row);
        uint64_t* dst = reinterpret_cast<uint64_t*>(p.base + row * p.pitch);
        for (uint32_t w = threadIdx.x; w < wpr; w += blockDim.x) {
            const uint32_t b = w * 8;
            uint64_t v = 0;
            if (b < p.row_bytes) {
                v = fill_word_from_key(key, p.byte_begin / 8 + w);
                const uint32_t rem = p.row_bytes - b;   // bytes of this word inside the row
                if (rem < 8) v &= (1ULL << (8 * rem)) - 1ULL;
            }
            dst[w] = v;
        }
    }
}

}  // namespace cobsgpu
