// cobs_b200/csrc/common.cuh -- shared device helpers (sm_100a): mbarrier / bulk-copy PTX,
// 64-bit mixing, small utilities.  No reference code; see DESIGN.md for the kernel map.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cobsgpu {

// ---------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (TMA 1-D bulk copy, SASS: UBLKCP) wrappers

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}

// make barrier inits visible to the async proxy before the first bulk copy uses them
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

// L2 eviction policy for data that is read exactly once (signature rows)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared bulk copy, completion signalled on an mbarrier (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint4 lds128(const void* smem_ptr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(smem_u32(smem_ptr)));
    return v;
}

// ---------------------------------------------------------------------------------------

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ULL;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return z;
}

// procedural index bits: 8 document-bytes (64 columns) of `row` in `page` at byte 8*word.
// Must stay identical to oracle_fill_word (oracle/cobs_oracle.c), which the parity tests use.
__host__ __device__ __forceinline__ uint64_t fill_row_key(uint64_t seed, uint32_t page,
                                                           uint64_t row) {
    return mix64(seed ^ mix64(row + (static_cast<uint64_t>(page) << 48)));
}
__host__ __device__ __forceinline__ uint64_t fill_word_from_key(uint64_t row_key, uint64_t word) {
    uint64_t a = mix64(row_key ^ (word * 0xD6E8FEB86659FD93ULL));
    uint64_t b = mix64(a + 0x9E3779B97F4A7C15ULL);
    return a & b;
}

template <typename T>
__host__ __device__ __forceinline__ T div_ceil(T a, T b) {
    return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ __forceinline__ T round_up(T a, T b) {
    return div_ceil(a, b) * b;
}

// sort key of one candidate: ascending order == (score desc, doc asc)
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t score, uint32_t doc) {
    return (static_cast<uint64_t>(~score) << 32) | doc;
}
__host__ __device__ __forceinline__ uint32_t key_score(uint64_t key) {
    return ~static_cast<uint32_t>(key >> 32);
}
__host__ __device__ __forceinline__ uint32_t key_doc(uint64_t key) {
    return static_cast<uint32_t>(key);
}
static constexpr uint64_t KEY_PAD = ~0ULL;  // sorts after every real candidate

}  // namespace cobsgpu
