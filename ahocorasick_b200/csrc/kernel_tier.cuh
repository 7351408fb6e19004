// Tables and probe helpers shared by the AhoCorasick kernels for narrow alphabets (kernel_mask.cuh, kernel_emit.cuh).
//
// Every haystack position q is an END anchor (all keywords ending at q, longest first — AhoCorasickSet.java:522-535):
//   * the context of q = the last max_len classes, b bits each, packs into one 64-bit word;
//   * levels 1..K are direct-indexed bit tables in SHARED memory (row layout, see TierTables in builder.hpp);
//   * level K+1 existence comes from a per-level-K-entry child mask (one L2-resident word), so the deep table is
//     only probed for contexts that really continue;
//   * levels > K are path-compressed: one 16-byte entry per chain head (keyed by the packed context of the head)
//     holds the unbranched chain below it, so one sector load resolves a whole keyword tail.
#pragma once
#include "builder.hpp"
#include "device_tables.cuh"

namespace acgpu {

struct DevTier {
    const uint32_t *row_words;   // direct-indexed level tables in row layout (copied to shared memory by every CTA)
    const uint32_t *prow_words;  // the same levels in PAIR layout ({fwd, back} per row, TierTables::prow_words) for k_tier_pair; nullptr: more than 31 classes
    uint32_t n_prow_words;
    uint32_t prow_off[10];
    uint32_t pair_gate_bit;
    uint32_t pair_low_bit;       // 0: no LOW bit (more than 30 classes), level K-1 takes its own pair-row load
    const uint32_t *cls8;        // 64 words: class of code units 0..255, one byte each
    const uint32_t *kidmask;     // [C^K][2] exact continuation masks of the level-K contexts: backward (this position), forward (the next one); nullptr: no deeper levels
    cudaTextureObject_t kid_tex; // the same table as a linear uint2 texture (k_tier_mask gathers it through the TEX pipe)
    const uint4 *buckets;        // deep table: 2 entries per 32-byte bucket, see TierTables in builder.hpp
    const uint4 *vbuckets;       // Map values: {key lo, key hi, value, 0}, 2 per bucket, key = packed classes | (length - 1) << 60
    unsigned long long vseed;
    uint32_t n_vbuckets;
    unsigned long long hash_seed;
    uint32_t n_row_words;
    uint32_t row_off[10];
    uint32_t n_buckets;
    uint32_t inv_b;              // ceil(65536 / b): (t * inv_b) >> 16 == t / b for t < 64
    uint32_t term_levels;
    int32_t b, C, K;
    uint32_t pow_c[10];
};

// streaming 128-bit haystack load, predicated (rows at the edges of the haystack)
__device__ __forceinline__ uint4 ldcs_v4_if(const void *p, bool on) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];\n}"
                 : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"((uint32_t)on));
    return v;
}

__device__ __forceinline__ unsigned long long deep_hash64_d(unsigned long long key, unsigned long long seed) {
    unsigned long long h = key ^ seed;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    return h;
}

// One deep probe: the entry of the chain head whose packed context is `key`.  Every probe is for a key that exists
// (child masks are exact) and the builder keeps tags unique along a probe path, so the first entry whose tag
// matches is the head.
__device__ __forceinline__ uint4 deep_probe(const DevTier &T, unsigned long long key, uint32_t &slot) {
    const unsigned long long h = deep_hash64_d(key, T.hash_seed);
    uint32_t bucket = __umulhi((uint32_t)h, T.n_buckets);
    const uint32_t want = ((uint32_t)(h >> 36) << 4) | 8u;
    while (true) {
        const uint4 *q = T.buckets + (size_t)bucket * 2;
        const uint4 e0 = __ldg(q);
        const uint4 e1 = __ldg(q + 1);
        if ((e0.x & ~7u) == want) { slot = bucket * 2u; return e0; }
        if ((e1.x & ~7u) == want) { slot = bucket * 2u + 1u; return e1; }
        bucket = bucket + 1u == T.n_buckets ? 0u : bucket + 1u;  // overflowed bucket: the key sits further along
    }
}

// Levels K+1.. of a context whose level-(K+1) node exists: bit (d - K - 1) set = a keyword of length d ends here.
// One probe per chain head; the chain below a head is compared against the context in registers.
template <int K>
__device__ __forceinline__ uint32_t deep_bits(const DevTier &T, unsigned long long ctx, uint32_t cm, int max_len) {
    uint32_t bits = 0;
    int d = K + 1;  // depth of the current head
    while (true) {
        uint32_t slot;
        const uint4 e = deep_probe(T, ctx & ((1ull << (T.b * d)) - 1ull), slot);
        const unsigned long long zw = ((unsigned long long)e.w << 32) | e.z;
        const int L = (int)(e.w >> 8) & 15;
        const uint32_t term = (e.w >> 12) & 0x1FFu;
        const unsigned long long x = ((ctx >> (T.b * d)) ^ zw) & ((1ull << (T.b * L)) - 1ull);
        const int m = x ? (int)(((uint32_t)(__ffsll((long long)x) - 1) * T.inv_b) >> 16) : L;  // chain steps that match
        bits |= (term & ((2u << m) - 1u)) << (d - K - 1);
        d += L;
        if (m < L || d >= max_len) break;
        const uint32_t c = ((uint32_t)(ctx >> (T.b * d))) & cm;
        if (!((e.y >> c) & 1u)) break;  // exact: no such child (class 0 never has one)
        d += 1;
    }
    return bits;
}

}  // namespace acgpu
