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
//   * levels > K are path-compressed: one 16-byte entry per chain head (keyed by the packed context of the head)
//     holds the unbranched chain below it, so one sector load resolves a whole keyword tail;
//   * the contexts of a warp row that continue past level K are compacted into a queue and walked 32 at a time;
//   * the depths that hit are kept as a bitmask per position: one pass, count = popc.
// Ordered emission: a CTA tile = 24 warp rows of 240 positions.  Worker warps count their row, hand the count to
// a dedicated RESOLVER warp through shared memory + named barriers (bar.arrive / bar.sync, no __syncthreads in the
// loop), and go on to count their next row; the resolver does ONE decoupled look-back per CTA tile and hands every
// worker its global record offset, which the worker picks up one iteration later (software pipeline), stages its
// records in a per-warp shared-memory window and flushes them with coalesced streaming stores.  Tiles are assigned
// statically to the persistent, fully resident (cooperatively launched) grid, so a look-back only ever waits for
// tiles that are running or finished.
#pragma once
#include "device_tables.cuh"

namespace acgpu {

struct DevTier {
    const uint32_t *smem_words;  // direct-indexed level tables (copied to shared memory by every CTA)
    const uint32_t *cls8;        // 64 words: class of code units 0..255, one byte each
    const uint32_t *kidmask;     // [C^K] exact child masks of the level-K entries (nullptr: no deeper levels)
    const uint4 *buckets;        // deep table: 2 entries per 32-byte bucket, see TierTables in builder.hpp
    const uint32_t *shallow_val;
    const uint32_t *deep_valbase;  // [bucket * 2 + entry] -> first value of the entry's chain in deep_val
    const uint32_t *deep_val;
    unsigned long long hash_seed;
    uint32_t n_words;
    uint32_t n_buckets;
    uint32_t inv_b;              // ceil(65536 / b): (t * inv_b) >> 16 == t / b for t < 64
    uint32_t term_levels;
    int32_t b, C, K;
    uint32_t lvl_off[10];
    uint32_t pow_c[10];
    unsigned long long val_off[10];
};

constexpr int kTierWorkers = 24;                         // worker warps per CTA
constexpr int kTierWarps = kTierWorkers;                 // rows per CTA tile
constexpr int kTierThreads = (kTierWorkers + 1) * 32;    // + the resolver warp
constexpr int kTierPer = 8;                              // consecutive positions per lane
constexpr int kTierRow = 30 * kTierPer;                  // 240 emitting positions per warp row
constexpr int kTierTile = kTierWorkers * kTierRow;       // positions per CTA tile = one look-back tile
constexpr int kTierCtrlWords = 256;                      // control block: counts + per-warp bases, double buffered

__host__ __device__ constexpr int tier_stage_records(bool is_map) { return is_map ? 256 : 384; }
__host__ __device__ constexpr size_t tier_stage_bytes(bool is_map) {
    return (size_t)kTierWorkers * tier_stage_records(is_map) * (is_map ? 12 : 8);
}
// dynamic shared memory of k_ac_tier for a table of n_words words
__host__ __device__ constexpr size_t tier_smem_bytes(size_t n_words, bool is_map) {
    return (64 + ((n_words + 3) & ~size_t(3)) + kTierCtrlWords) * sizeof(uint32_t) + tier_stage_bytes(is_map);
}

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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

// classes of the 8 chars at [p0, p0+8); positions outside [0, n) give class 0
__device__ __forceinline__ void load_classes8(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t p0,
                                              const uint8_t *s_cls8, uint32_t (&c)[kTierPer]) {
    if (p0 >= 0 && p0 + 8 <= n) {  // rows are laid out so that hay + p0 is 16-byte aligned
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(hay + p0));  // streaming: read once
        if (((v.x | v.y | v.z | v.w) & 0xFF00FF00u) == 0u) {
            c[0] = s_cls8[v.x & 0xFFu]; c[1] = s_cls8[v.x >> 16];
            c[2] = s_cls8[v.y & 0xFFu]; c[3] = s_cls8[v.y >> 16];
            c[4] = s_cls8[v.z & 0xFFu]; c[5] = s_cls8[v.z >> 16];
            c[6] = s_cls8[v.w & 0xFFu]; c[7] = s_cls8[v.w >> 16];
        } else {
            const uint32_t ch[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                                    v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
#pragma unroll
            for (int j = 0; j < 8; j++) c[j] = ch[j] < 256u ? (uint32_t)s_cls8[ch[j]] : (uint32_t)__ldg(&A.cls[ch[j]]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t p = p0 + j;
            c[j] = (p >= 0 && p < n) ? (uint32_t)__ldg(&A.cls[__ldg(&hay[p])]) : 0u;
        }
    }
}

// value index of the keyword of length d that ends the context (Maps)
template <int K>
__device__ __forceinline__ uint32_t tier_value(const DevTier &T, unsigned long long ctx, uint32_t cm, int d) {
    if (d <= K) {
        uint32_t idx = 0;
#pragma unroll
        for (int i = 1; i <= K; i++) {
            if (i <= d) idx += ((uint32_t)(ctx >> (T.b * (i - 1))) & cm) * T.pow_c[i];
        }
        return __ldg(&T.shallow_val[T.val_off[d] + idx]);
    }
    int hd = K + 1;  // walk the chain heads down to the one whose chain covers depth d
    while (true) {
        uint32_t slot;
        const uint4 e = deep_probe(T, ctx & ((1ull << (T.b * hd)) - 1ull), slot);
        const int L = (int)(e.w >> 8) & 15;
        if (d <= hd + L) {
            const uint32_t term = (e.w >> 12) & 0x1FFu;
            return __ldg(&T.deep_val[__ldg(&T.deep_valbase[slot]) + __popc(term & ((1u << (d - hd)) - 1u))]);
        }
        hd += L + 1;
    }
}

// AcArgs: n_tiles = number of CTA tiles; a row = 240 emitting positions (lanes 2..31) preceded by 16 context
// positions (lanes 0,1), so one 128-bit load per lane covers the row and its left context.  Rows start at
// P.origin (<= emit_from, chosen by the host so that every lane's load is 16-byte aligned).
// LOW selects how the levels below K are looked up: 0 = every level that has keywords (predicated), 1 = only level
// K-1 has keywords, 2 = no keyword is shorter than K.
template <int K, int LOW, bool kIsMap>
__global__ void __launch_bounds__(kTierThreads, 1) k_ac_tier(const DevAutomaton A, const DevTier T, const AcArgs P) {
    constexpr int kStage = tier_stage_records(kIsMap);
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint8_t *s_cls8 = reinterpret_cast<const uint8_t *>(s_mem);
    const uint32_t *s_tab = s_mem + 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u;
    uint32_t *s_ctrl = s_mem + 64 + ((T.n_words + 3u) & ~3u);
    uint32_t *s_cnt = s_ctrl;                                                             // [2][32]
    unsigned long long *s_wbase = reinterpret_cast<unsigned long long *>(s_ctrl + 64);   // [2][32]
    int2 *s_stage_all = reinterpret_cast<int2 *>(s_ctrl + kTierCtrlWords);
    int2 *s_stage = s_stage_all + warp * kStage;
    uint32_t *s_stage_val = reinterpret_cast<uint32_t *>(s_stage_all + kTierWorkers * kStage) + warp * kStage;

    for (uint32_t i = tid; i < 64; i += kTierThreads) s_mem[i] = __ldg(&T.cls8[i]);
    for (uint32_t i = tid; i < T.n_words; i += kTierThreads) s_mem[64 + i] = __ldg(&T.smem_words[i]);
    __syncthreads();

    const int64_t n_tiles = P.n_tiles;
    const int n_iter = (int)((n_tiles - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);

    // ================================================================= resolver warp
    if (warp == kTierWorkers) {
        for (int it = 0; it < n_iter; ++it) {
            const int buf = it & 1;
            named_sync(1 + buf, kTierThreads);  // the 24 row counts of this tile are in s_cnt[buf]
            const uint32_t c = lane < kTierWorkers ? s_cnt[buf * 32 + lane] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            if (lane == 0) lookback_publish(P.status, tile, total);
            const unsigned long long excl = lookback_resolve(P.status, tile, total);
            if (lane < kTierWorkers) s_wbase[buf * 32 + lane] = excl + inc - c;
            if (lane == 0 && tile == n_tiles - 1) *P.total_out = excl + total;
            __threadfence_block();
            named_arrive(3 + buf, kTierThreads);
        }
        return;
    }

    // ================================================================= worker warps
    const uint32_t C = (uint32_t)T.C;
    const uint32_t sh = 1u << b;
    const bool deeper = T.kidmask != nullptr;  // some keyword is longer than K
    // row state carried to the next iteration (emission is one iteration behind counting)
    uint32_t p_masks[kTierPer], p_cnt = 0, p_inc = 0, p_total = 0, p_hi = 0, p_lo = 0;
    int32_t p_e0 = 0;
#pragma unroll
    for (int j = 0; j < kTierPer; j++) p_masks[j] = 0;

    for (int it = 0; it <= n_iter; ++it) {
        uint32_t masks[kTierPer], my_cnt = 0, inc = 0, row_total = 0, hi = 0, lo = 0;
        int32_t e0 = 0;
#pragma unroll
        for (int j = 0; j < kTierPer; j++) masks[j] = 0;

        if (it < n_iter) {
            const int64_t row = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kTierWorkers + warp;
            const int64_t q_lo = P.origin + row * kTierRow;
            const int64_t q_hi = min(P.emit_to, q_lo + (int64_t)kTierRow);
            const int64_t p0 = q_lo - 16 + (int64_t)lane * kTierPer;
            e0 = (int32_t)(p0 + 1) + P.pos_base;
            // bit j set: position p0 + j reports matches
            uint32_t vm = 0;
            if (lane >= 2) {
                const int64_t lo_j = max(P.emit_from, q_lo) - p0, hi_j = q_hi - p0;
                const uint32_t a = lo_j <= 0 ? 0xFFu : (lo_j >= 8 ? 0u : (0xFFu << (int)lo_j) & 0xFFu);
                const uint32_t z = hi_j >= 8 ? 0xFFu : (hi_j <= 0 ? 0u : (0xFFu >> (8 - (int)hi_j)));
                vm = a & z;
            }
            uint32_t c[kTierPer];
            load_classes8(A, P.hay, P.n, p0, s_cls8, c);
            hi = ((c[0] * sh + c[1]) * sh + c[2]) * sh + c[3];
            lo = ((c[4] * sh + c[5]) * sh + c[6]) * sh + c[7];
            const uint32_t lo1 = __shfl_up_sync(0xFFFFFFFFu, lo, 1);
            const uint32_t hi1 = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
            // class i positions before this lane's first one (i = 1..8); lane 0 reads garbage and never reports
            auto prev_class = [&](int i) -> uint32_t {
                return i <= 4 ? (lo1 >> (b * (i - 1))) & cm : (hi1 >> (b * (i - 5))) & cm;
            };

            // ---- phase A: rolling mixed-radix indices of the last 1..K classes; shared-memory level tables
            uint32_t r[K + 1];
            {
                uint32_t acc = 0;
                r[0] = 0;
#pragma unroll
                for (int k = 1; k < K; k++) {
                    acc += prev_class(k) * T.pow_c[k];
                    r[k] = acc;
                }
                r[K] = 0;
            }
            uint32_t ki[kTierPer];
#pragma unroll
            for (int j = 0; j < kTierPer; j++) {
#pragma unroll
                for (int k = K; k >= 2; k--) r[k] = r[k - 1] * C + c[j];
                r[1] = c[j];
                uint32_t m = 0;
#pragma unroll
                for (int i = 1; i < K; i++) {
                    if (LOW == 2 || (LOW == 1 && i != K - 1)) continue;
                    const bool on = LOW == 1 || ((T.term_levels >> i) & 1u);
                    const uint32_t w = on ? s_tab[T.lvl_off[i] + (r[i] >> 5)] : 0u;
                    m |= (__funnelshift_r(w, 0u, r[i]) & 1u) << i;
                }
                const uint32_t w = s_tab[T.lvl_off[K] + (r[K] >> 4)];
                const uint32_t f = __funnelshift_r(w, 0u, (r[K] & 15u) * 2u) & 3u;
                m |= (f & 1u) << K;
                const bool valid = (vm >> j) & 1u;
                masks[j] = valid ? m : 0u;
                ki[j] = (valid && (f & 2u)) ? r[K] : 0xFFFFFFFFu;
            }

            if (deeper) {
                // ---- phase B: exact child masks of the level-K entries (all 8 loads in flight together) say which
                //      contexts continue to level K+1
                uint32_t pm = 0;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) ki[j] = ki[j] != 0xFFFFFFFFu ? __ldg(&T.kidmask[ki[j]]) : 0u;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) {
                    const uint32_t ck = j >= K ? c[j >= K ? j - K : 0] : prev_class(K - j);
                    pm |= ((ki[j] >> ck) & 1u) << j;
                }
                // ---- phase C: the continuing contexts of the whole row are compacted into a queue (ids in the warp's
                //      staging window, free at this point) and walked 32 at a time with all lanes busy
                uint32_t n_mine = __popc(pm), q_inc = n_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, q_inc, o);
                    if (lane >= o) q_inc += y;
                }
                const uint32_t q_total = __shfl_sync(0xFFFFFFFFu, q_inc, 31);
                if (q_total) {
                    uint8_t *s_q = reinterpret_cast<uint8_t *>(s_stage);                 // [256] ids = lane * 8 + j
                    uint16_t *s_r = reinterpret_cast<uint16_t *>(s_stage) + 128;         // [256] deep bits per id
                    reinterpret_cast<uint4 *>(s_r)[lane] = make_uint4(0u, 0u, 0u, 0u);
                    uint32_t o = q_inc - n_mine;
                    for (uint32_t t = pm; t; t &= t - 1u) s_q[o++] = (uint8_t)(lane * 8 + (__ffs(t) - 1));
                    __syncwarp();
                    for (uint32_t base = 0; base < q_total; base += 32) {
                        const bool act = base + lane < q_total;
                        const uint32_t id = act ? (uint32_t)s_q[base + lane] : 64u;
                        const int owner = (int)(id >> 3), j = (int)(id & 7u);
                        const uint32_t h0 = __shfl_sync(0xFFFFFFFFu, hi, owner), l0 = __shfl_sync(0xFFFFFFFFu, lo, owner);
                        const uint32_t h1 = __shfl_sync(0xFFFFFFFFu, hi, owner - 1), l1 = __shfl_sync(0xFFFFFFFFu, lo, owner - 1);
                        const uint32_t h2 = __shfl_sync(0xFFFFFFFFu, hi, owner - 2), l2 = __shfl_sync(0xFFFFFFFFu, lo, owner - 2);
                        if (act) {
                            const unsigned long long P0 = ((unsigned long long)h0 << (4 * b)) | l0;
                            const unsigned long long P1 = ((unsigned long long)h1 << (4 * b)) | l1;
                            const unsigned long long P2 = ((unsigned long long)h2 << (4 * b)) | l2;
                            const unsigned long long ctx = ((((P2 << (8 * b)) | P1) << (b * (j + 1)))) | (P0 >> (b * (7 - j)));
                            s_r[id] = (uint16_t)deep_bits<K>(T, ctx, cm, A.max_len);
                        }
                    }
                    __syncwarp();
                    const uint4 rr = reinterpret_cast<const uint4 *>(s_r)[lane];
                    masks[0] |= (rr.x & 0xFFFFu) << (K + 1); masks[1] |= (rr.x >> 16) << (K + 1);
                    masks[2] |= (rr.y & 0xFFFFu) << (K + 1); masks[3] |= (rr.y >> 16) << (K + 1);
                    masks[4] |= (rr.z & 0xFFFFu) << (K + 1); masks[5] |= (rr.z >> 16) << (K + 1);
                    masks[6] |= (rr.w & 0xFFFFu) << (K + 1); masks[7] |= (rr.w >> 16) << (K + 1);
                    __syncwarp();
                }
            }
            // ---- ordered offsets inside the row; hand the row total to the resolver
#pragma unroll
            for (int j = 0; j < kTierPer; j++) my_cnt += __popc(masks[j]);
            inc = my_cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (lane == 0) s_cnt[(it & 1) * 32 + warp] = row_total;
            __threadfence_block();
            named_arrive(1 + (it & 1), kTierThreads);
        }

        // ---- finish the previous row of this warp: pick up its global offset, stage, flush
        if (it > 0) {
            const int pb = (it - 1) & 1;
            named_sync(3 + pb, kTierThreads);
            const unsigned long long base = s_wbase[pb * 32 + warp];
            unsigned long long ctx0 = 0;
            if (kIsMap) {
                const uint32_t h1 = __shfl_up_sync(0xFFFFFFFFu, p_hi, 1), l1 = __shfl_up_sync(0xFFFFFFFFu, p_lo, 1);
                const uint32_t h2 = __shfl_up_sync(0xFFFFFFFFu, p_hi, 2), l2 = __shfl_up_sync(0xFFFFFFFFu, p_lo, 2);
                ctx0 = (((((unsigned long long)h2 << (4 * b)) | l2)) << (8 * b)) | (((unsigned long long)h1 << (4 * b)) | l1);
            }
            const unsigned long long Pk = ((unsigned long long)p_hi << (4 * b)) | p_lo;
            const uint32_t my_off = p_inc - p_cnt;
            if (!kIsMap && p_total <= (uint32_t)kStage) {
                // ---- common case: the whole row fits the staging window.  Longest first: deep levels (rare, loop),
                //      then the two shared-memory levels that carry almost all matches as predicated straight-line
                //      stores, then shallower levels (rare, loop)
                if (p_cnt) {
                    int2 *dst = s_stage + my_off;
#pragma unroll
                    for (int j = 0; j < kTierPer; j++) {
                        const uint32_t m = p_masks[j];
                        const int32_t e = p_e0 + j;
                        if (m >> (K + 1)) {
                            for (uint32_t t = m >> (K + 1); t;) {
                                const int d = 31 - __clz(t);
                                t ^= 1u << d;
                                *dst++ = make_int2(e - (d + K + 1), e);
                            }
                        }
                        const bool tk = (m >> K) & 1u;
                        if (tk) *dst = make_int2(e - K, e);
                        dst += tk;
                        if (K >= 2 && LOW != 2) {
                            const bool tk1 = (m >> (K - 1)) & 1u;
                            if (tk1) *dst = make_int2(e - (K - 1), e);
                            dst += tk1;
                        }
                        if (K >= 3 && LOW == 0) {
                            for (uint32_t t = m & ((1u << (K - 1)) - 1u); t;) {
                                const int d = 31 - __clz(t);
                                t ^= 1u << d;
                                *dst++ = make_int2(e - d, e);
                            }
                        }
                    }
                }
                __syncwarp();
                for (uint32_t rr = lane; rr < p_total; rr += 32) {
                    const unsigned long long g = base + rr;
                    if (g < (unsigned long long)P.cap) __stcs(&P.pos_out[g], s_stage[rr]);
                }
                __syncwarp();
            } else {
                for (uint32_t win = 0; win < p_total; win += kStage) {
                    if (p_cnt && my_off < win + kStage && my_off + p_cnt > win) {
                        uint32_t o = my_off - win;  // wraps below zero for records of an earlier window
#pragma unroll 1
                        for (int j = 0; j < kTierPer; j++) {
                            const int32_t e = p_e0 + j;
                            const unsigned long long ctx = kIsMap ? (ctx0 << (b * (j + 1))) | (Pk >> (b * (kTierPer - 1 - j))) : 0ull;
                            uint32_t m = 0;
#pragma unroll
                            for (int jj = 0; jj < kTierPer; jj++) m = jj == j ? p_masks[jj] : m;
                            while (m) {  // longest first
                                const int d = 31 - __clz(m);
                                m ^= 1u << d;
                                if (o < (uint32_t)kStage) {
                                    s_stage[o] = make_int2(e - d, e);
                                    if (kIsMap) s_stage_val[o] = tier_value<K>(T, ctx, cm, d);
                                }
                                ++o;
                            }
                        }
                    }
                    __syncwarp();
                    const uint32_t cnt = min((uint32_t)kStage, p_total - win);
                    for (uint32_t rr = lane; rr < cnt; rr += 32) {
                        const unsigned long long g = base + win + rr;
                        if (g < (unsigned long long)P.cap) {
                            __stcs(&P.pos_out[g], s_stage[rr]);
                            if (kIsMap) __stcs(&P.val_out[g], s_stage_val[rr]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        p_e0 = e0;
        p_hi = hi;
        p_lo = lo;
        p_cnt = my_cnt;
        p_inc = inc;
        p_total = row_total;
#pragma unroll
        for (int j = 0; j < kTierPer; j++) p_masks[j] = masks[j];
    }
}

}  // namespace acgpu
