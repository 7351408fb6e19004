// Generation-2 AhoCorasick kernel for narrow alphabets: k_ac_tier.
//
// Every haystack position q is an END anchor (all keywords ending at q, longest first —
// AhoCorasickSet.java:522-535), but the anchored trie is not walked node by node:
//   * the context of q = the last max_len classes, b bits each, lives in ONE 64-bit register and rolls from
//     position to position (a lane owns 8 consecutive positions, its left context arrives from the two lanes
//     before it by warp shuffle — the haystack is read exactly once, 128-bit coalesced, streaming, straight
//     into registers);
//   * levels 1..K are direct-indexed bit tables in SHARED memory (mixed-radix index of the last j classes);
//   * level K+1 existence comes from a per-level-K-entry child mask (one L2-resident word), so the deep table is
//     only probed for contexts that really continue;
//   * levels > K are 8-byte slots (context << 4 | flags) in an open-addressing table keyed by the context itself;
//   * loads of a lane's 8 positions are issued back to back (memory-level parallelism) before any is consumed;
//   * the depths that hit are kept as a bitmask per position: one pass, count = popc.
// Ordered emission without block barriers: one WARP row (256 positions) is one tile of the decoupled look-back;
// rows are assigned statically to the persistent, fully resident grid, so a warp only ever waits for rows that
// are running or finished.  Records are staged in a per-warp shared-memory window and flushed with coalesced
// streaming stores.
#pragma once
#include "device_tables.cuh"

namespace acgpu {

struct DevTier {
    const uint32_t *smem_words;  // direct-indexed level tables (copied to shared memory by every CTA)
    const uint32_t *cls8;        // 64 words: class of code units 0..255, one byte each
    const uint32_t *kidmask;     // [C^K] exact child masks of the level-K entries (nullptr: no deeper levels)
    const uint4 *buckets;        // deep table: 2 x uint4 per 32-byte bucket = 4 entries {tag<<4 | 8 | flags, child mask}
    const uint32_t *shallow_val;
    const uint32_t *deep_val;    // [bucket * 4 + entry]
    unsigned long long hash_seed;
    uint32_t n_words;
    uint32_t bucket_mask;
    uint32_t term_levels;
    int32_t b, C, K;
    uint32_t lvl_off[10];
    uint32_t pow_c[10];
    unsigned long long val_off[10];
};

constexpr int kTierThreads = 768;
constexpr int kTierWarps = kTierThreads / 32;
constexpr int kTierPer = 8;                  // consecutive positions per lane
constexpr int kTierRow = 30 * kTierPer;      // 240 emitting positions per warp row = one look-back tile
constexpr int kTierStage = 256;              // records per warp staging window

__host__ __device__ constexpr size_t tier_stage_bytes(bool is_map) {
    return (size_t)kTierWarps * kTierStage * (is_map ? 12 : 8);
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

// A deep probe in two halves so that several can be in flight: issue the sector load, consume it later.
// Every probe is for a key that exists (child masks are exact) and the builder keeps tags unique along a probe
// path, so the first entry whose tag matches is the node.
struct DeepProbe {
    uint4 a, b;
    uint32_t bucket, want;
};

__device__ __forceinline__ void deep_issue(const DevTier &T, unsigned long long key, DeepProbe &p) {
    const unsigned long long h = deep_hash64_d(key, T.hash_seed);
    p.bucket = (uint32_t)h & T.bucket_mask;
    p.want = ((uint32_t)(h >> 36) << 4) | 8u;
    const uint4 *q = T.buckets + (size_t)p.bucket * 2;
    p.a = __ldg(q);
    p.b = __ldg(q + 1);
}

__device__ __forceinline__ void deep_consume(const DevTier &T, DeepProbe &p, uint32_t &flags, uint32_t &kids, uint32_t &slot) {
    while (true) {
        if ((p.a.x & ~7u) == p.want) { flags = p.a.x & 7u; kids = p.a.y; slot = p.bucket * 4u; return; }
        if ((p.a.z & ~7u) == p.want) { flags = p.a.z & 7u; kids = p.a.w; slot = p.bucket * 4u + 1u; return; }
        if ((p.b.x & ~7u) == p.want) { flags = p.b.x & 7u; kids = p.b.y; slot = p.bucket * 4u + 2u; return; }
        if ((p.b.z & ~7u) == p.want) { flags = p.b.z & 7u; kids = p.b.w; slot = p.bucket * 4u + 3u; return; }
        p.bucket = (p.bucket + 1u) & T.bucket_mask;  // overflowed bucket: the key sits further along
        const uint4 *q = T.buckets + (size_t)p.bucket * 2;
        p.a = __ldg(q);
        p.b = __ldg(q + 1);
    }
}

// classes of the 8 chars at [p0, p0+8), first char in the HIGHEST field; positions outside [0, n) give class 0
__device__ __forceinline__ unsigned long long pack8(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t p0,
                                                    const uint8_t *s_cls8, int b) {
    uint32_t ch[8];
    if (p0 >= 0 && p0 + 8 <= n && ((reinterpret_cast<uintptr_t>(hay + p0) & 15) == 0)) {
        uint4 v = __ldcs(reinterpret_cast<const uint4 *>(hay + p0));  // streaming: read once
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
    DeepProbe pr;
    uint32_t fl, kids, slot;
    deep_issue(T, ctx & ((1ull << (T.b * d)) - 1ull), pr);
    deep_consume(T, pr, fl, kids, slot);
    return __ldg(&T.deep_val[slot]);
}

// Levels d0.. for a context whose level (d0-1) node has flags `fl` and child mask `kids`; returns hit bits.
__device__ __forceinline__ uint32_t deep_walk_from(const DevAutomaton &A, const DevTier &T, unsigned long long ctx,
                                                   uint32_t cm, int d0, uint32_t fl, uint32_t kids) {
    uint32_t m = 0;
    for (int d = d0; d <= A.max_len && (fl & 2u); d++) {
        const uint32_t c = ((uint32_t)(ctx >> (T.b * (d - 1)))) & cm;
        if (!((kids >> c) & 1u)) break;  // exact: no such child (class 0 never has one)
        DeepProbe pr;
        uint32_t slot;
        deep_issue(T, ctx & ((1ull << (T.b * d)) - 1ull), pr);
        deep_consume(T, pr, fl, kids, slot);
        m |= (fl & 1u) << d;
    }
    return m;
}

// AcArgs.n_tiles = number of rows; a row = 240 emitting positions (lanes 2..31) preceded by 16 context positions
// (lanes 0,1), so one 128-bit load per lane covers the row and its left context.  tile_counter is unused.
template <int K, bool kIsMap>
__global__ void __launch_bounds__(kTierThreads, 1) k_ac_tier(const DevAutomaton A, const DevTier T, const AcArgs P) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint8_t *s_cls8 = reinterpret_cast<const uint8_t *>(s_mem);
    const uint32_t *s_tab = s_mem + 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u;
    // per-warp staging window behind the tables
    int2 *s_stage = reinterpret_cast<int2 *>(s_mem + 64 + ((T.n_words + 3u) & ~3u)) + warp * kTierStage;
    uint32_t *s_stage_val = reinterpret_cast<uint32_t *>(reinterpret_cast<int2 *>(s_mem + 64 + ((T.n_words + 3u) & ~3u)) +
                                                         kTierWarps * kTierStage) + warp * kTierStage;

    for (uint32_t i = tid; i < 64; i += kTierThreads) s_mem[i] = __ldg(&T.cls8[i]);
    for (uint32_t i = tid; i < T.n_words; i += kTierThreads) s_mem[64 + i] = __ldg(&T.smem_words[i]);
    __syncthreads();

    const int64_t n_rows = P.n_tiles;
    // software pipeline: row i's records are resolved and written after row i+1 has been counted and its
    // aggregate published, so the look-back of row i finds its predecessors ready
    bool pv = false;
    int64_t p_row = 0, p_p0 = 0;
    unsigned long long p_Pk = 0, p_ctx0 = 0;
    uint32_t p_masks[kTierPer], p_cnt = 0, p_inc = 0, p_total = 0;
#pragma unroll
    for (int j = 0; j < kTierPer; j++) p_masks[j] = 0;

    for (int64_t row = (int64_t)blockIdx.x * kTierWarps + warp;; row += (int64_t)gridDim.x * kTierWarps) {
        const bool have = row < n_rows;
        int64_t p0 = 0;
        unsigned long long Pk = 0, ctx0 = 0;
        uint32_t masks[kTierPer], my_cnt = 0, inc = 0, row_total = 0;
#pragma unroll
        for (int j = 0; j < kTierPer; j++) masks[j] = 0;
        if (have) {
            const int64_t q_lo = P.emit_from + row * kTierRow;
            const int64_t q_hi = min(P.emit_to, q_lo + (int64_t)kTierRow);
            p0 = q_lo - 16 + (int64_t)lane * kTierPer;
            Pk = pack8(A, P.hay, P.n, p0, s_cls8, b);
            const unsigned long long Q1 = __shfl_up_sync(0xFFFFFFFFu, Pk, 1);
            const unsigned long long Q2 = __shfl_up_sync(0xFFFFFFFFu, Pk, 2);
            ctx0 = (Q2 << (8 * b)) | Q1;  // lanes 0,1 hold garbage here and never emit

            // ---- phase A: shared-memory levels for the 8 positions; remember which contexts may continue
            uint32_t ki[kTierPer];
            {
                unsigned long long ctx = ctx0;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) {
                    ctx = (ctx << b) | ((Pk >> (b * (kTierPer - 1 - j))) & cm);
                    uint32_t m = 0, idx = 0, f = 0;
#pragma unroll
                    for (int i = 1; i <= K; i++) {
                        idx += ((uint32_t)(ctx >> (b * (i - 1))) & cm) * T.pow_c[i];
                        if (i < K) {
                            if ((T.term_levels >> i) & 1u) m |= ((s_tab[T.lvl_off[i] + (idx >> 5)] >> (idx & 31u)) & 1u) << i;
                        } else {
                            f = (s_tab[T.lvl_off[K] + (idx >> 4)] >> ((idx & 15u) * 2u)) & 3u;
                            m |= (f & 1u) << K;
                        }
                    }
                    const bool valid = lane >= 2 && (p0 + j) < q_hi;
                    masks[j] = valid ? m : 0u;
                    ki[j] = (valid && (f & 2u)) ? idx : 0xFFFFFFFFu;
                }
            }
            // ---- phase B: exact child masks of the level-K entries (all 8 loads in flight together) say which
            //      contexts continue to level K+1
            uint32_t probe_mask = 0;
            if (T.kidmask) {
#pragma unroll
                for (int j = 0; j < kTierPer; j++) ki[j] = ki[j] != 0xFFFFFFFFu ? __ldg(&T.kidmask[ki[j]]) : 0u;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) {
                    // context of position j in closed form: ctx0 shifted by j+1 classes | the first j+1 classes of Pk
                    const unsigned long long ctx = (ctx0 << (b * (j + 1))) | (Pk >> (b * (kTierPer - 1 - j)));
                    const uint32_t c = (uint32_t)(ctx >> (b * K)) & cm;
                    probe_mask |= ((ki[j] >> c) & 1u) << j;
                }
            }
            // ---- phase C: deep levels.  Converged loop: every lane takes up to two of its continuing positions per
            //      round, both sector loads are issued before either is consumed.  Hit bits go to a packed word
            //      (8 bits per position, bit = depth - K - 1; 16 levels in two words) to avoid indexed registers.
            unsigned long long dlo = 0, dhi = 0;
            while (__any_sync(0xFFFFFFFFu, probe_mask != 0)) {
                int j1 = -1, j2 = -1;
                if (probe_mask) {
                    j1 = __ffs(probe_mask) - 1;
                    probe_mask &= probe_mask - 1;
                }
                if (probe_mask) {
                    j2 = __ffs(probe_mask) - 1;
                    probe_mask &= probe_mask - 1;
                }
                DeepProbe pr1, pr2;
                unsigned long long c1 = 0, c2 = 0;
                if (j1 >= 0) {
                    c1 = (ctx0 << (b * (j1 + 1))) | (Pk >> (b * (kTierPer - 1 - j1)));
                    deep_issue(T, c1 & ((1ull << (b * (K + 1))) - 1ull), pr1);
                }
                if (j2 >= 0) {
                    c2 = (ctx0 << (b * (j2 + 1))) | (Pk >> (b * (kTierPer - 1 - j2)));
                    deep_issue(T, c2 & ((1ull << (b * (K + 1))) - 1ull), pr2);
                }
                if (j1 >= 0) {
                    uint32_t fl, kids, slot;
                    deep_consume(T, pr1, fl, kids, slot);
                    uint32_t bits = (fl & 1u) | (deep_walk_from(A, T, c1, cm, K + 2, fl, kids) >> (K + 1));
                    dlo |= (unsigned long long)(bits & 0xFFu) << (8 * j1);
                    dhi |= (unsigned long long)((bits >> 8) & 0xFFu) << (8 * j1);
                }
                if (j2 >= 0) {
                    uint32_t fl, kids, slot;
                    deep_consume(T, pr2, fl, kids, slot);
                    uint32_t bits = (fl & 1u) | (deep_walk_from(A, T, c2, cm, K + 2, fl, kids) >> (K + 1));
                    dlo |= (unsigned long long)(bits & 0xFFu) << (8 * j2);
                    dhi |= (unsigned long long)((bits >> 8) & 0xFFu) << (8 * j2);
                }
            }
#pragma unroll
            for (int j = 0; j < kTierPer; j++) {
                masks[j] |= ((uint32_t)(dlo >> (8 * j)) & 0xFFu) << (K + 1);
                masks[j] |= ((uint32_t)(dhi >> (8 * j)) & 0xFFu) << (K + 9);
            }
            // ---- ordered offsets inside the row; publish the row aggregate right away
#pragma unroll
            for (int j = 0; j < kTierPer; j++) my_cnt += __popc(masks[j]);
            inc = my_cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (lane == 0) lookback_publish(P.status, row, row_total);
        }

        // ---- finish the previous row of this warp: look-back, stage, flush
        if (pv) {
            const unsigned long long base = lookback_resolve(P.status, p_row, p_total);
            if (lane == 0 && p_row == n_rows - 1) *P.total_out = base + p_total;
            const uint32_t my_off = p_inc - p_cnt;
            for (uint32_t win = 0; win < p_total; win += kTierStage) {
                if (p_cnt && my_off < win + kTierStage && my_off + p_cnt > win) {
                    unsigned long long ctx = p_ctx0;
                    uint32_t o = my_off;
#pragma unroll
                    for (int j = 0; j < kTierPer; j++) {
                        ctx = (ctx << b) | ((p_Pk >> (b * (kTierPer - 1 - j))) & cm);
                        uint32_t m = p_masks[j];
                        const int32_t e = (int32_t)(p_p0 + j + 1);
                        while (m) {
                            const int d = 31 - __clz(m);
                            m ^= 1u << d;
                            if (o >= win && o < win + kTierStage) {
                                s_stage[o - win] = make_int2(e - d + P.pos_base, e + P.pos_base);
                                if (kIsMap) s_stage_val[o - win] = tier_value<K>(A, T, ctx, cm, d);
                            }
                            ++o;
                        }
                    }
                }
                __syncwarp();
                const uint32_t cnt = min((uint32_t)kTierStage, p_total - win);
                for (uint32_t r = lane; r < cnt; r += 32) {
                    const unsigned long long g = base + win + r;
                    if (g < (unsigned long long)P.cap) {
                        __stcs(&P.pos_out[g], s_stage[r]);
                        if (kIsMap) __stcs(&P.val_out[g], s_stage_val[r]);
                    }
                }
                __syncwarp();
            }
        }
        if (!have) break;
        pv = true;
        p_row = row;
        p_p0 = p0;
        p_Pk = Pk;
        p_ctx0 = ctx0;
        p_cnt = my_cnt;
        p_inc = inc;
        p_total = row_total;
#pragma unroll
        for (int j = 0; j < kTierPer; j++) p_masks[j] = masks[j];
    }
}

}  // namespace acgpu
