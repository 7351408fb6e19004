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
// records in a per-warp shared-memory window and flushes them with coalesced streaming stores.  Tiles are handed out
// by a global ticket counter (SMs differ in speed; a static split leaves the fast ones waiting at the look-back):
// warp 0 draws the CTA's tickets two iterations ahead and publishes them in shared memory.  The grid is persistent
// and fully resident (cooperative launch) and tickets ascend, so a look-back only ever waits for tiles that are
// running or finished.
#pragma once
#include "device_tables.cuh"

namespace acgpu {

struct DevTier {
    const uint32_t *smem_words;  // direct-indexed level tables (copied to shared memory by every CTA)
    const uint32_t *row_words;   // the same levels in row layout (k_tier_mask), see TierTables in builder.hpp
    const uint32_t *cls8;        // 64 words: class of code units 0..255, one byte each
    const uint32_t *kidmask;     // [C^K] exact child masks of the level-K entries (nullptr: no deeper levels)
    cudaTextureObject_t kid_tex; // the same table as a linear texture (k_tier_mask gathers it through the TEX pipe)
    const uint4 *buckets;        // deep table: 2 entries per 32-byte bucket, see TierTables in builder.hpp
    const uint32_t *shallow_val;
    const uint32_t *deep_valbase;  // [bucket * 2 + entry] -> first value of the entry's chain in deep_val
    const uint32_t *deep_val;
    unsigned long long hash_seed;
    uint32_t n_words;
    uint32_t n_row_words;
    uint32_t row_off[10];
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
constexpr int kTierCtrlWords = 416;                      // control block: row counts + per-warp bases (4 phases), 8 tile tickets, 4 mbarriers

// Emission runs tier_depth() iterations behind counting: a row's hit masks wait in a per-warp shared-memory ring
// until the resolver has produced the row's global record offset.
__host__ __device__ constexpr int tier_stage_records(bool is_map) { return is_map ? 160 : 256; }
__host__ __device__ constexpr int tier_depth(bool is_map) { return is_map ? 2 : 3; }
// ring slot: 32 x uint4 hit masks (8 positions x 16 depth bits), 32 x u16 inclusive record offsets, Maps: 32 x (hi, lo),
// the tile index (16 bytes reserved)
__host__ __device__ constexpr size_t tier_slot_bytes(bool is_map) { return 512 + 64 + (is_map ? 256 : 0) + 16; }
__host__ __device__ constexpr size_t tier_warp_bytes(bool is_map) {
    return tier_depth(is_map) * tier_slot_bytes(is_map) + (size_t)tier_stage_records(is_map) * (is_map ? 12 : 8);
}
// dynamic shared memory of k_ac_tier for a table of n_words words
__host__ __device__ constexpr size_t tier_smem_bytes(size_t n_words, bool is_map) {
    return (64 + ((n_words + 3) & ~size_t(3)) + kTierCtrlWords) * sizeof(uint32_t) + kTierWorkers * tier_warp_bytes(is_map);
}

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// mbarrier (one signaller, many waiters: unlike bar.sync the waiters do not wait for each other)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra LAB_WAIT;\n DONE:\n}"
        ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Loads that must be ISSUED where they are written (their consumers sit behind barriers and other work): volatile asm
// keeps the compiler from sinking them to the first use.
__device__ __forceinline__ uint32_t ldg_u32_early(const uint32_t *p) {
    uint32_t x;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(p));
    return x;
}
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

// classes of the 8 chars at [p0, p0+8) given the prefetched vector v (valid when the 8 chars lie inside [0, n));
// positions outside [0, n) give class 0
__device__ __forceinline__ void classify8(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t p0, const uint4 v,
                                          const uint8_t *s_cls8, uint32_t (&c)[kTierPer]) {
    if (p0 >= 0 && p0 + 8 <= n) {
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
    constexpr int kDepth = tier_depth(kIsMap);
    constexpr int kSlot = (int)tier_slot_bytes(kIsMap);
    extern __shared__ __align__(16) uint32_t s_mem[];
    const uint8_t *s_cls8 = reinterpret_cast<const uint8_t *>(s_mem);
    const uint32_t *s_tab = s_mem + 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u;
    uint32_t *s_ctrl = s_mem + 64 + ((T.n_words + 3u) & ~3u);
    uint32_t *s_cnt = s_ctrl;                                                              // [4][32]
    unsigned long long *s_wbase = reinterpret_cast<unsigned long long *>(s_ctrl + 128);   // [4][32]
    // tile tickets: slot (it & 7) = (it + 1) << 32 | tile of iteration it (tile >= n_tiles: the corpus is exhausted)
    volatile unsigned long long *s_ticket = reinterpret_cast<volatile unsigned long long *>(s_ctrl + 384);
    // mbarrier ph: "the record offsets of iteration it (it & 3 == ph) are in s_wbase[ph]"; phase parity (it >> 2) & 1
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_ctrl + 400);
    const int64_t n_tiles = P.n_tiles;

    for (uint32_t i = tid; i < 64; i += kTierThreads) s_mem[i] = __ldg(&T.cls8[i]);
    for (uint32_t i = tid; i < T.n_words; i += kTierThreads) s_mem[64 + i] = __ldg(&T.smem_words[i]);
    if (tid < 8) {
        unsigned long long w = 0;
        if (tid < 2) w = ((unsigned long long)(tid + 1) << 32) | (unsigned long long)atomicAdd(P.tile_counter, 1u);
        s_ticket[tid] = w;
    }
    if (tid >= 32 && tid < 36) mbar_init(s_bar + (tid - 32), 1);
    __syncthreads();
    auto wait_ticket = [&](int it) -> uint32_t {
        unsigned long long w;
        do {
            w = s_ticket[it & 7];
        } while ((uint32_t)(w >> 32) != (uint32_t)(it + 1));
        return (uint32_t)w;
    };

    // ================================================================= resolver warp
    // named barriers 1..4: "row counts of iteration it are in s_cnt[it & 3]" (workers bar.arrive, the resolver
    // bar.sync's); mbarriers 0..3: "offsets of iteration it are in s_wbase[it & 3]" (the resolver arrives, every worker
    // waits on its own).  Four phases: a worker is at most kDepth + 1 <= 4 iterations ahead of the resolver.
    if (warp == kTierWorkers) {
        for (int it = 0;; ++it) {
            const int64_t tile = wait_ticket(it);
            if (tile >= n_tiles) break;
            const int ph = it & 3;
            named_sync(1 + ph, kTierThreads);
            const uint32_t c = lane < kTierWorkers ? s_cnt[ph * 32 + lane] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (lane == 0) lookback_publish(P.status, tile, total);
            const unsigned long long excl = lookback_resolve(P.status, tile, total);
            if (lane < kTierWorkers) s_wbase[ph * 32 + lane] = excl + inc - c;
            if (lane == 0 && tile == n_tiles - 1) *P.total_out = excl + total;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_bar + ph);
        }
        return;
    }

    // ================================================================= worker warps
    unsigned char *s_warp = reinterpret_cast<unsigned char *>(s_ctrl + kTierCtrlWords) + (size_t)warp * tier_warp_bytes(kIsMap);
    int2 *s_stage = reinterpret_cast<int2 *>(s_warp + kDepth * kSlot);
    uint32_t *s_stage_val = reinterpret_cast<uint32_t *>(s_stage + kStage);
    const uint32_t C = (uint32_t)T.C;
    const uint32_t sh = 1u << b;
    const bool deeper = T.kidmask != nullptr;  // some keyword is longer than K
    const uint32_t lt_mask = (1u << lane) - 1u;
    auto row_p0 = [&](uint32_t tile) -> int64_t {  // first char of this lane in this warp's row of the tile
        const int64_t row = (int64_t)tile * kTierWorkers + warp;
        return P.origin + row * kTierRow - 16 + (int64_t)lane * kTierPer;
    };
    auto fetch = [&](uint32_t tile) -> uint4 {  // rows are laid out so that hay + p0 is 16-byte aligned
        const int64_t p0 = row_p0(tile);
        return ldcs_v4_if(P.hay + p0, (int64_t)tile < n_tiles && p0 >= 0 && p0 + 8 <= P.n);
    };

    uint32_t tile_cur = wait_ticket(0);
    uint4 v = fetch(tile_cur);
    int slot_w = 0, slot_r = 0;  // ring slots written by counting / read by emission
    int end_it = 0x7FFFFFFF;     // first iteration without a tile
    for (int it = 0; it - kDepth < end_it; ++it) {
        const bool counting = (int64_t)tile_cur < n_tiles;
        if (!counting && end_it == 0x7FFFFFFF) end_it = it;
        uint32_t tile_next = 0xFFFFFFFFu;
        uint32_t c[kTierPer], masks[kTierPer], ki[kTierPer];
        uint32_t hi = 0, lo = 0, lo1 = 0, hi1 = 0;
#pragma unroll
        for (int j = 0; j < kTierPer; j++) c[j] = masks[j] = ki[j] = 0;

        if (counting) {
            // warp 0 draws the ticket of iteration it + 2 (published further down, once the atomic has returned)
            unsigned int drawn = 0;
            if (warp == 0 && lane == 0) drawn = atomicAdd(P.tile_counter, 1u);
            tile_next = wait_ticket(it + 1);
            const int64_t p0 = row_p0(tile_cur);
            const int64_t q_lo = p0 + 16 - (int64_t)lane * kTierPer;
            const int64_t q_hi = min(P.emit_to, q_lo + (int64_t)kTierRow);
            // bit j set: position p0 + j reports matches
            uint32_t vm = 0;
            if (lane >= 2) {
                const int64_t lo_j = max(P.emit_from, q_lo) - p0, hi_j = q_hi - p0;
                const uint32_t a = lo_j <= 0 ? 0xFFu : (lo_j >= 8 ? 0u : (0xFFu << (int)lo_j) & 0xFFu);
                const uint32_t z = hi_j >= 8 ? 0xFFu : (hi_j <= 0 ? 0u : (0xFFu >> (8 - (int)hi_j)));
                vm = a & z;
            }
            classify8(A, P.hay, P.n, p0, v, s_cls8, c);
            v = fetch(tile_next);  // next row's chars, in flight while this row is processed
            hi = ((c[0] * sh + c[1]) * sh + c[2]) * sh + c[3];
            lo = ((c[4] * sh + c[5]) * sh + c[6]) * sh + c[7];
            lo1 = __shfl_up_sync(0xFFFFFFFFu, lo, 1);
            hi1 = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
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
                // ---- phase B, issue: exact child mask of the level-K entry.  The 8 loads stay in flight across the
                //      emission block below and are consumed after it.
                // (entry 0 = the all-"other" context has no children: positions that need no mask read it, an L1 hit)
                if (deeper) ki[j] = ldg_u32_early(T.kidmask + ((valid && (f & 2u)) ? r[K] : 0u));
            }
            if (warp == 0 && lane == 0) s_ticket[(it + 2) & 7] = ((unsigned long long)(it + 3) << 32) | drawn;
        }

        // ---- emission of the row counted kDepth iterations ago: pick up its global offset, stage, flush
        if (it >= kDepth) {  // it - kDepth < end_it: that iteration had a tile
            const int eit = it - kDepth;
            mbar_wait(s_bar + (eit & 3), (uint32_t)(eit >> 2) & 1u);
            const unsigned long long base = s_wbase[(eit & 3) * 32 + warp];
            const unsigned char *slot = s_warp + slot_r * kSlot;
            slot_r = slot_r + 1 == kDepth ? 0 : slot_r + 1;
            const uint4 mm = reinterpret_cast<const uint4 *>(slot)[lane];
            const uint16_t *s_inc = reinterpret_cast<const uint16_t *>(slot + 512);
            const uint32_t p_inc = s_inc[lane], p_total = s_inc[31];
            uint32_t p_masks[kTierPer];  // bit d = a keyword of length d ends here (stored shifted down by one)
            p_masks[0] = (mm.x & 0xFFFFu) << 1; p_masks[1] = (mm.x >> 16) << 1;
            p_masks[2] = (mm.y & 0xFFFFu) << 1; p_masks[3] = (mm.y >> 16) << 1;
            p_masks[4] = (mm.z & 0xFFFFu) << 1; p_masks[5] = (mm.z >> 16) << 1;
            p_masks[6] = (mm.w & 0xFFFFu) << 1; p_masks[7] = (mm.w >> 16) << 1;
            const uint32_t p_cnt = __popc(mm.x) + __popc(mm.y) + __popc(mm.z) + __popc(mm.w);
            const int32_t p_e0 = (int32_t)(row_p0(*reinterpret_cast<const uint32_t *>(slot + kSlot - 16)) + 1) + P.pos_base;
            unsigned long long ctx0 = 0, Pk = 0;
            if (kIsMap) {
                const uint2 hl = reinterpret_cast<const uint2 *>(slot + 576)[lane];
                const uint32_t h1 = __shfl_up_sync(0xFFFFFFFFu, hl.x, 1), l1 = __shfl_up_sync(0xFFFFFFFFu, hl.y, 1);
                const uint32_t h2 = __shfl_up_sync(0xFFFFFFFFu, hl.x, 2), l2 = __shfl_up_sync(0xFFFFFFFFu, hl.y, 2);
                ctx0 = (((((unsigned long long)h2 << (4 * b)) | l2)) << (8 * b)) | (((unsigned long long)h1 << (4 * b)) | l1);
                Pk = ((unsigned long long)hl.x << (4 * b)) | hl.y;
            }
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
                {
                    // records of this row that fit the caller's buffer (cap = 0: count only)
                    const unsigned long long room = base < (unsigned long long)P.cap ? (unsigned long long)P.cap - base : 0ull;
                    const uint32_t n_out = (uint32_t)min((unsigned long long)p_total, room);
                    int2 *gp = P.pos_out + base + lane;
                    const int2 *sp = s_stage + lane;
                    for (uint32_t rr = lane; rr < n_out; rr += 32, gp += 32, sp += 32) __stcs(gp, *sp);
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

        if (counting) {
            if (deeper) {
                // ---- phase B, consume: which contexts continue to level K+1
                auto prev_class = [&](int i) -> uint32_t {
                    return i <= 4 ? (lo1 >> (b * (i - 1))) & cm : (hi1 >> (b * (i - 5))) & cm;
                };
                uint32_t pm = 0;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) {
                    const uint32_t ck = j >= K ? c[j >= K ? j - K : 0] : prev_class(K - j);
                    pm |= ((ki[j] >> ck) & 1u) << j;
                }
                // ---- phase C: the continuing contexts of the whole row are compacted into a queue (ids in the warp's
                //      staging window, free at this point) and walked 32 at a time with all lanes busy
                uint8_t *s_q = reinterpret_cast<uint8_t *>(s_stage);                 // [256] ids = lane * 8 + j
                uint16_t *s_r = reinterpret_cast<uint16_t *>(s_stage) + 128;         // [256] deep bits per id
                uint32_t q_total = 0;
#pragma unroll
                for (int j = 0; j < kTierPer; j++) {
                    const bool on = (pm >> j) & 1u;
                    const uint32_t bj = __ballot_sync(0xFFFFFFFFu, on);
                    if (on) s_q[q_total + __popc(bj & lt_mask)] = (uint8_t)(lane * 8 + j);
                    q_total += __popc(bj);
                }
                if (q_total) {
                    reinterpret_cast<uint4 *>(s_r)[lane] = make_uint4(0u, 0u, 0u, 0u);
                    __syncwarp();
                    for (uint32_t qb = 0; qb < q_total; qb += 32) {
                        const bool act = qb + lane < q_total;
                        const uint32_t id = act ? (uint32_t)s_q[qb + lane] : 64u;
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
                }
            }
            // ---- ordered offsets inside the row; park the row in the ring; hand the row total to the resolver
            uint32_t my_cnt = 0;
#pragma unroll
            for (int j = 0; j < kTierPer; j++) my_cnt += __popc(masks[j]);
            uint32_t inc = my_cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            const uint32_t row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            unsigned char *slot = s_warp + slot_w * kSlot;
            slot_w = slot_w + 1 == kDepth ? 0 : slot_w + 1;
            __syncwarp();
            reinterpret_cast<uint4 *>(slot)[lane] =
                make_uint4((masks[0] >> 1) | (masks[1] >> 1) << 16, (masks[2] >> 1) | (masks[3] >> 1) << 16,
                           (masks[4] >> 1) | (masks[5] >> 1) << 16, (masks[6] >> 1) | (masks[7] >> 1) << 16);
            reinterpret_cast<uint16_t *>(slot + 512)[lane] = (uint16_t)inc;
            if (kIsMap) reinterpret_cast<uint2 *>(slot + 576)[lane] = make_uint2(hi, lo);
            if (lane == 0) {
                *reinterpret_cast<uint32_t *>(slot + kSlot - 16) = tile_cur;
                s_cnt[(it & 3) * 32 + warp] = row_total;
            }
            __threadfence_block();
            named_arrive(1 + (it & 3), kTierThreads);
        }
        tile_cur = tile_next;
    }
}

}  // namespace acgpu
