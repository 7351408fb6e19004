// Generation-2 AhoCorasick kernel for narrow alphabets: k_ac_tier.
//
// Every haystack position q is still an END anchor (all keywords ending at q, longest first —
// AhoCorasickSet.java:522-535), but the anchored trie is not walked node by node:
//   * the context of q = the last max_len classes, b bits each, lives in ONE 64-bit register and rolls
//     from position to position (a lane owns 8 consecutive positions, its left context arrives from the
//     two lanes before it by warp shuffle — the haystack is read exactly once, 128-bit coalesced, straight
//     into registers);
//   * levels 1..K are direct-indexed bit tables in SHARED memory (mixed-radix index of the last j classes);
//   * levels > K are 8-byte slots (context << 4 | flags) in an L2-resident open-addressing table keyed by the
//     context itself, so a probe does not depend on the previous level's result except for its has-children bit;
//   * the depths that hit are kept as a bitmask per position: one pass, count = popc, then block scan +
//     decoupled look-back, then the records are written longest first from the mask.
#pragma once
#include "device_tables.cuh"

namespace acgpu {

struct DevTier {
    const uint32_t *smem_words;  // direct-indexed level tables (copied to shared memory by every CTA)
    const uint32_t *cls8;        // 64 words: class of code units 0..255, one byte each
    const unsigned long long *deep;
    const uint32_t *shallow_val;
    const uint32_t *deep_val;
    uint32_t n_words;
    uint32_t deep_mask;
    uint32_t term_levels;
    int32_t b, C, K;
    uint32_t lvl_off[10];
    uint32_t pow_c[10];
    unsigned long long val_off[10];
};

constexpr int kTierThreads = 1024;
constexpr int kTierWarps = kTierThreads / 32;
constexpr int kTierPer = 8;                         // consecutive positions per lane
constexpr int kTierTile = kTierThreads * kTierPer;  // 8192 end positions per tile

__device__ __forceinline__ uint32_t deep_hash_d(unsigned long long key) {
    key ^= key >> 29;
    key *= 0xBF58476D1CE4E5B9ull;
    key ^= key >> 32;
    return (uint32_t)key;
}

__device__ __forceinline__ bool deep_find(const DevTier &T, unsigned long long key, uint32_t &flags, uint32_t &slot) {
    uint32_t i = deep_hash_d(key) & T.deep_mask;
    while (true) {
        unsigned long long s = __ldg(&T.deep[i]);
        if ((s >> 4) == key) {
            flags = (uint32_t)s & 15u;
            slot = i;
            return true;
        }
        if (s == 0) return false;
        i = (i + 1) & T.deep_mask;
    }
}

// classes of the 8 chars at [p0, p0+8), first char in the HIGHEST field; positions outside [0, n) give class 0
__device__ __forceinline__ unsigned long long pack8(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t p0,
                                                    const uint8_t *s_cls8, int b) {
    uint32_t ch[8];
    if (p0 >= 0 && p0 + 8 <= n && ((reinterpret_cast<uintptr_t>(hay + p0) & 15) == 0)) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(hay + p0));
        ch[0] = v.x & 0xFFFFu; ch[1] = v.x >> 16; ch[2] = v.y & 0xFFFFu; ch[3] = v.y >> 16;
        ch[4] = v.z & 0xFFFFu; ch[5] = v.z >> 16; ch[6] = v.w & 0xFFFFu; ch[7] = v.w >> 16;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int64_t p = p0 + j;
            ch[j] = (p >= 0 && p < n) ? (uint32_t)__ldg(&hay[p]) : 0x10000u;  // 0x10000 = outside
        }
    }
    unsigned long long P = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t c = ch[j] < 256u ? (uint32_t)s_cls8[ch[j]] : (ch[j] < 0x10000u ? (uint32_t)__ldg(&A.cls[ch[j]]) : 0u);
        P = (P << b) | c;
    }
    return P;
}

template <int K>
__device__ __forceinline__ uint32_t tier_hits(const DevAutomaton &A, const DevTier &T, const uint32_t *s_tab,
                                              unsigned long long ctx, uint32_t cm) {
    uint32_t m = 0, idx = 0, deeper = 0;
#pragma unroll
    for (int i = 1; i <= K; i++) {
        uint32_t ci = (uint32_t)(ctx >> (T.b * (i - 1))) & cm;
        idx += ci * T.pow_c[i];
        if (i < K) {
            if ((T.term_levels >> i) & 1u) m |= ((s_tab[T.lvl_off[i] + (idx >> 5)] >> (idx & 31u)) & 1u) << i;
        } else {
            uint32_t f = (s_tab[T.lvl_off[K] + (idx >> 4)] >> ((idx & 15u) * 2u)) & 3u;
            m |= (f & 1u) << K;
            deeper = f & 2u;
        }
    }
    if (deeper) {
        for (int d = K + 1; d <= A.max_len; d++) {
            // class 0 (char in no keyword, or before the haystack start) ends the walk; without this test a
            // depth-d key with a zero top field would alias the depth-(d-1) entry
            if ((((uint32_t)(ctx >> (T.b * (d - 1)))) & cm) == 0u) break;
            unsigned long long key = ctx & ((1ull << (T.b * d)) - 1ull);
            uint32_t fl, slot;
            if (!deep_find(T, key, fl, slot)) break;
            m |= (fl & 1u) << d;
            if (!(fl & 2u)) break;
        }
    }
    return m;
}

template <int K>
__device__ __forceinline__ uint32_t tier_value(const DevAutomaton &A, const DevTier &T, unsigned long long ctx,
                                               uint32_t cm, int d) {
    if (d <= K) {
        uint32_t idx = 0;
#pragma unroll
        for (int i = 1; i <= K; i++) {
            if (i <= d) idx += ((uint32_t)(ctx >> (T.b * (i - 1))) & cm) * T.pow_c[i];
        }
        return __ldg(&T.shallow_val[T.val_off[d] + idx]);
    }
    uint32_t fl, slot = 0;
    deep_find(T, ctx & ((1ull << (T.b * d)) - 1ull), fl, slot);
    return __ldg(&T.deep_val[slot]);
}

template <int K, bool kIsMap>
__global__ void __launch_bounds__(kTierThreads, 1) k_ac_tier(const DevAutomaton A, const DevTier T, const AcArgs P) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint8_t *s_cls8 = reinterpret_cast<const uint8_t *>(s_mem);
    const uint32_t *s_tab = s_mem + 64;
    __shared__ uint32_t s_warp_tot[kTierWarps];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u;

    for (uint32_t i = tid; i < 64; i += kTierThreads) s_mem[i] = __ldg(&T.cls8[i]);
    for (uint32_t i = tid; i < T.n_words; i += kTierThreads) s_mem[64 + i] = __ldg(&T.smem_words[i]);
    __syncthreads();

    while (true) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t q_lo = P.emit_from + tile * kTierTile;
        const int64_t q_hi = min(P.emit_to, q_lo + (int64_t)kTierTile);
        const int64_t p0 = q_lo + (int64_t)tid * kTierPer;

        // my 8 classes + the 16 classes before them (lanes 0/1 of a warp fetch theirs from memory)
        const unsigned long long Pk = pack8(A, P.hay, P.n, p0, s_cls8, b);
        unsigned long long Q1 = __shfl_up_sync(0xFFFFFFFFu, Pk, 1);
        unsigned long long Q2 = __shfl_up_sync(0xFFFFFFFFu, Pk, 2);
        if (lane < 2) {
            // lane 0 needs chunks -1 and -2 of the warp's row, lane 1 needs chunk -1 as its Q2
            const int64_t w0 = p0 - (int64_t)lane * kTierPer;  // first position of the warp's row
            unsigned long long m1 = pack8(A, P.hay, P.n, w0 - 8, s_cls8, b);
            if (lane == 0) {
                Q1 = m1;
                Q2 = pack8(A, P.hay, P.n, w0 - 16, s_cls8, b);
            } else {
                Q2 = m1;
            }
        }
        const unsigned long long ctx0 = (Q2 << (8 * b)) | Q1;

        uint32_t masks[kTierPer];
        uint32_t my_cnt = 0;
        {
            unsigned long long ctx = ctx0;
#pragma unroll
            for (int j = 0; j < kTierPer; j++) {
                ctx = (ctx << b) | ((Pk >> (b * (kTierPer - 1 - j))) & cm);
                const int64_t q = p0 + j;
                uint32_t m = 0;
                if (q < q_hi) m = tier_hits<K>(A, T, s_tab, ctx, cm);
                masks[j] = m;
                my_cnt += __popc(m);
            }
        }

        // ordered offsets: thread order == position order
        uint32_t inc = my_cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_warp_tot[warp] = inc;
        __syncthreads();
        uint32_t block_total = 0;
        if (warp == 0) {
            uint32_t t = s_warp_tot[lane];
            uint32_t ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ti, o);
                if (lane >= o) ti += y;
            }
            block_total = __shfl_sync(0xFFFFFFFFu, ti, 31);
            s_warp_tot[lane] = ti - t;  // exclusive warp bases
            unsigned long long excl = lookback_exclusive(P.status, tile, block_total);
            if (lane == 0) {
                s_base = excl;
                if (tile == P.n_tiles - 1) *P.total_out = excl + block_total;
            }
        }
        __syncthreads();
        unsigned long long idx = s_base + s_warp_tot[warp] + (inc - my_cnt);

        if (my_cnt) {
            unsigned long long ctx = ctx0;
#pragma unroll
            for (int j = 0; j < kTierPer; j++) {
                ctx = (ctx << b) | ((Pk >> (b * (kTierPer - 1 - j))) & cm);
                uint32_t m = masks[j];
                const int32_t e = (int32_t)(p0 + j + 1);
                while (m) {
                    const int d = 31 - __clz(m);
                    m ^= 1u << d;
                    if (idx < (unsigned long long)P.cap) {
                        P.pos_out[idx] = make_int2(e - d + P.pos_base, e + P.pos_base);
                        if (kIsMap) P.val_out[idx] = tier_value<K>(A, T, ctx, cm, d);
                    }
                    ++idx;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace acgpu
