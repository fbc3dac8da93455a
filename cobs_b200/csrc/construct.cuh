// cobs_b200/csrc/construct.cuh -- classic index construction on the device.
//
// The inverse of the query path: every k-mer of every document sets h bits,
//   row = XXH64(canonical k-mer, seed j) % signature_size,  column = document
// (reference: process_term / set_bit, cobs/construction/classic_index.cpp:39-73; k-mers of a
// document are all windows of each of its sequences, cobs/fasta_file.hpp:156-182).
// One thread per (sequence, window); bits are set with atomicOr on the 32-bit word holding the
// column, in the same row-major LSB-first layout the score kernel reads.
//
// Bit-exactness detail: with canonicalize == 1 the reference does NOT skip k-mers containing
// non-ACGT characters during construction -- canonicalize_kmer() writes a binary zero for each
// of them and the buffer is hashed anyway (classic_index.cpp:57-71, only a warning is logged).
// hash_kmer_any() below reproduces exactly that buffer.
#pragma once

#include "common.cuh"
#include "hash.cuh"

namespace cobsgpu {

struct ConstructParams {
    const char* sequences;     // device: concatenated characters
    const uint64_t* seq_off;   // device [n_seqs+1]
    const uint64_t* win_off;   // device [n_seqs+1]: prefix sum of windows (k-mers) per sequence
    const uint32_t* seq_doc;   // device [n_seqs]
    uint32_t n_seqs;
    uint64_t total_windows;
    uint32_t k, h, canonicalize;
    uint64_t sig;
    uint8_t* base;             // device matrix [sig][pitch]
    uint32_t pitch;            // multiple of 16
};

// hashes of the k-mer at s under the reference's construction rules (invalid bases -> 0)
template <typename Emit>
__device__ __forceinline__ void hash_kmer_any(const uint8_t* s, uint32_t k, uint32_t h,
                                              uint32_t canonicalize, Emit emit) {
    if (!canonicalize) {
        for (uint32_t j = 0; j < h; ++j)
            emit(xxh::hash64([&](uint32_t i) { return s[i]; }, k, j));
        return;
    }
    bool reverse = false;
    for (uint32_t i = 0; i < k / 2; ++i) {
        const uint8_t f = base_fwd(s[i]), r = base_rev(s[k - 1 - i]);
        if (f != r) {
            reverse = f > r;
            break;
        }
    }
    for (uint32_t j = 0; j < h; ++j)
        emit(xxh::hash64(
            [&](uint32_t i) { return reverse ? base_rev(s[k - 1 - i]) : base_fwd(s[i]); }, k, j));
}

__global__ void __launch_bounds__(128) construct_classic_kernel(ConstructParams p) {
    for (uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
         gid < p.total_windows; gid += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        // sequence owning window gid: largest s with win_off[s] <= gid
        uint32_t lo = 0, hi = p.n_seqs;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.win_off[mid] <= gid) lo = mid;
            else hi = mid;
        }
        const uint64_t w = gid - p.win_off[lo];
        const uint8_t* s = reinterpret_cast<const uint8_t*>(p.sequences) + p.seq_off[lo] + w;
        const uint32_t doc = p.seq_doc[lo];
        hash_kmer_any(s, p.k, p.h, p.canonicalize, [&](uint64_t hv) {
            const uint64_t row = hv % p.sig;
            uint32_t* word = reinterpret_cast<uint32_t*>(p.base + row * p.pitch) + (doc >> 5);
            atomicOr(word, 1u << (doc & 31));
        });
    }
}

}  // namespace cobsgpu
