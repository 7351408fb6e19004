// AhoCorasick family for dictionaries OUTSIDE the narrow-alphabet envelope of kernel_mask.cuh (more than 32 character
// classes - mixed-case, digits, Latin-1, ... - or keywords of 13..32 chars with 5-bit classes): the same three launches
//
//   k_wide_mask   every haystack position q is an END anchor (AhoCorasickSet.java:522-535: own match, then the
//                 suffixMatch chain = the terminal nodes on the path of the reversed-keyword trie spelled by h[q], h[q-1], ...);
//                 one 32-BIT hit mask per position (bit 32 - d = a keyword of length d ends here) + one count per 256-position row
//   k_row_scan    (kernel_emit.cuh) exclusive scan of the row counts
//   k_wide_emit   masks -> (start, end[, value]) records at their final offsets, position-major, longest first
//
// k_wide_mask: a warp owns chunks of 32 rows (tickets); a lane owns 8 consecutive positions of a row (one streaming 128-bit
// load) and walks them IN LOCKSTEP, level by level, so that the 8 dependent gather chains of a lane overlap.  Levels 1 and 2
// come from a direct-indexed class-pair table in shared memory (<= 64 classes), deeper levels from the open-addressing
// edge table (device_tables.cuh::trie_step's layout; L2-resident).  Classes of the row and of the 32 chars before it sit in
// a per-warp shared-memory window.
#pragma once
#include "kernel_emit.cuh"

namespace acgpu {

#ifndef ACGPU_WIDE_MIN_CTAS
#define ACGPU_WIDE_MIN_CTAS 3   // resident CTAs per SM the register allocation must allow (measured on config 5: 3 -> 47.7, 4 -> 47.3, 5 -> 44.5, 6 -> 35.3 GB/s)
#endif
constexpr int kWideWarps = 8;
constexpr int kWideThreads = kWideWarps * 32;
constexpr int kWideMaxLen = 32;      // hit masks are 32 bits
constexpr int kWideLockLevels = 5;   // levels walked in lockstep by the owning lane; deeper walks are compacted over the warp
constexpr int kWidePairMax = 64;
constexpr int kWideChainMaxD = 8;   // = kWideChainMax of builder.hpp: chain steps per path-compressed entry     // class-pair table: C * C * 8 bytes <= 32 KB of shared memory (3 CTAs per SM stay below the L1 carve-out cliff)

struct DevWide {
    const uint2 *pair;   // [C * C] (c0 * C + c1) -> {level-2 node or kNoneD, info2 | info1 << 8 | level-1 node exists << 16}; nullptr: no table
    int32_t C;
    const uint4 *chain;   // path-compressed edges below level 2 (HostAutomaton::wide_chain): two uint4 per 32-byte entry; nullptr: none
    const uint4 *pair16;  // [C * C] k_wide_tile's table of levels 1 and 2 (HostAutomaton::wide_pair16)
    const uint4 *vals;    // Map values by keyword hash (HostAutomaton::wide_vals), two entries per bucket; nullptr: walk the trie
    uint32_t n_vbuckets;
};

struct WideArgs {
    const uint16_t *hay;
    int64_t n;
    int64_t emit_from;      // positions (index of a keyword's last char) in [emit_from, emit_to) report
    int64_t emit_to;
    int64_t origin;         // first position of row 0: <= emit_from and hay + origin is 16-byte aligned
    uint32_t *masks;        // [n_rows * 256] one word per position
    uint32_t *row_count;    // [n_rows]
    unsigned int *ticket;
    int64_t n_rows;
    // k_wide_tile -> k_wide_tail: walks a tile hands over instead of finishing them in thin rounds {entry, position, depth, 0}
    uint4 *tail;
    unsigned int *tail_count;   // [0] walks handed over
    uint32_t tail_cap;
    uint32_t min_rounds, hand_over;  // a tile runs min_rounds rounds itself and hands its walks over once at most hand_over are left
};

__host__ __device__ constexpr size_t wide_smem_bytes(int C, bool pair) {
    // classes 0..255 | pair table | per warp: class window (32 + 256 halfwords), deep-walk queue (256 nodes, 256 hit words, 256 positions)
    return 512 + (pair ? (size_t)C * C * 8 : 0) + (size_t)kWideWarps * ((32 + kMaskRow) * 2 + kMaskRow * 9);
}

template <bool PAIR>
__global__ void __launch_bounds__(kWideThreads, ACGPU_WIDE_MIN_CTAS) k_wide_mask(const DevAutomaton A, const DevWide Wd, const WideArgs P) {
    extern __shared__ __align__(16) unsigned char s_wide[];
    uint16_t *s_cls8 = reinterpret_cast<uint16_t *>(s_wide);                    // classes of code units 0..255
    const uint2 *s_pair = reinterpret_cast<const uint2 *>(s_wide + 512);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = Wd.C;
    unsigned char *s_warp = s_wide + 512 + (PAIR ? (size_t)C * C * 8 : 0) + (size_t)warp * ((32 + kMaskRow) * 2 + kMaskRow * 9);
    uint32_t *q_node = reinterpret_cast<uint32_t *>(s_warp);
    uint32_t *q_bits = q_node + kMaskRow;
    uint16_t *w = reinterpret_cast<uint16_t *>(q_bits + kMaskRow);
    uint8_t *q_pos = reinterpret_cast<uint8_t *>(w + 32 + kMaskRow);
    for (int i = tid; i < 256; i += kWideThreads) s_cls8[i] = __ldg(&A.cls[i]);
    if (PAIR) {
        uint2 *dst = reinterpret_cast<uint2 *>(s_wide + 512);
        for (int i = tid; i < C * C; i += kWideThreads) dst[i] = __ldg(&Wd.pair[i]);
    }
    __syncthreads();
    const int max_len = min(A.max_len, kWideMaxLen);

    auto class_of = [&](uint32_t ch) -> uint32_t { return ch < 256u ? (uint32_t)s_cls8[ch] : (uint32_t)__ldg(&A.cls[ch]); };

    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(P.ticket, 1u);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        const int64_t row0 = (int64_t)chunk * kMaskChunkRows;
        if (row0 >= P.n_rows) break;
        const int n_cr = (int)min((int64_t)kMaskChunkRows, P.n_rows - row0);
        const int64_t c_lo = P.origin + row0 * kMaskRow;
        // the 32 classes before the chunk
        {
            const int64_t p = c_lo - 32 + lane;
            __syncwarp();
            w[lane] = (p >= 0 && p < P.n) ? (uint16_t)class_of(__ldg(&P.hay[p])) : (uint16_t)0;
        }
        for (int r = 0; r < n_cr; ++r) {
            const int64_t p0 = c_lo + (int64_t)r * kMaskRow + (int64_t)lane * 8;
            uint32_t ch[8];
            if (p0 >= 0 && p0 + 8 <= P.n) {
                const uint4 v = ldcs_v4_if(P.hay + p0, true);
                ch[0] = v.x & 0xFFFFu; ch[1] = v.x >> 16; ch[2] = v.y & 0xFFFFu; ch[3] = v.y >> 16;
                ch[4] = v.z & 0xFFFFu; ch[5] = v.z >> 16; ch[6] = v.w & 0xFFFFu; ch[7] = v.w >> 16;
#pragma unroll
                for (int j = 0; j < 8; j++) w[32 + lane * 8 + j] = (uint16_t)class_of(ch[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int64_t p = p0 + j;
                    w[32 + lane * 8 + j] = (p >= 0 && p < P.n) ? (uint16_t)class_of(__ldg(&P.hay[p])) : (uint16_t)0;
                }
            }
            __syncwarp();
            // ---- the 8 anchored walks of the lane, in lockstep
            const uint16_t *wp = w + 32 + lane * 8;  // wp[j - i] = class i positions before position j
            uint32_t m[8], node[8];
            uint32_t alive = 0;
            int d0;  // first level the edge table serves
            if (PAIR) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t c0 = wp[j], c1 = wp[j - 1];
                    const uint2 e = s_pair[c0 * (uint32_t)C + c1];
                    uint32_t mj = 0;
                    if ((e.y >> 8) & kTerm) mj |= 1u << 31;
                    node[j] = e.x;
                    if (e.x != kNoneD && max_len >= 2) {
                        if (e.y & kTerm) mj |= 1u << 30;
                        if (e.y & kKids) alive |= 1u << j;
                    }
                    m[j] = mj;
                }
                d0 = 3;
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t c0 = wp[j];
                    uint32_t mj = 0;
                    node[j] = kNoneD;
                    if (c0 != 0u) {
                        const uint2 rt = __ldg(&A.root[c0]);
                        if (rt.x != kNoneD) {
                            node[j] = rt.x;
                            if (rt.y & kTerm) mj |= 1u << 31;
                            if (rt.y & kKids) alive |= 1u << j;
                        }
                    }
                    m[j] = mj;
                }
                d0 = 2;
            }
            const int d_lock = min(max_len, kWideLockLevels);
            for (int d = d0; alive != 0u && d <= d_lock; ++d) {
                // first probe of every live walk, issued back to back
                uint4 e[8];
                uint32_t slot[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t c = wp[j - d + 1];
                    if (c == 0u) alive &= ~(1u << j);  // a char that is in no keyword ends every walk
                    slot[j] = edge_hash_d(node[j], c) & A.edge_mask;
                    e[j] = make_uint4(kNoneD, 0u, 0u, 0u);
                    if ((alive >> j) & 1u) e[j] = __ldg(&A.edges[slot[j]]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (!((alive >> j) & 1u)) continue;
                    const uint32_t c = wp[j - d + 1];
                    uint4 ej = e[j];
                    uint32_t i = slot[j];
                    while (!(ej.x == node[j] && ej.y == c) && ej.x != kNoneD) {  // open addressing: the key sits further along
                        i = (i + 1u) & A.edge_mask;
                        ej = __ldg(&A.edges[i]);
                    }
                    if (ej.x == kNoneD) {
                        alive &= ~(1u << j);
                        continue;
                    }
                    node[j] = ej.z;
                    if (ej.w & kTerm) m[j] |= 1u << (32 - d);
                    if (!(ej.w & kKids)) alive &= ~(1u << j);
                }
            }
            // ---- walks that are still alive (inside a long keyword: few) are compacted over the warp and finished one
            //      walk per lane; their hits come back through the warp's shared-memory slots
            if (max_len > d_lock) {
                const uint32_t n_mine = (uint32_t)__popc(alive);
                uint32_t inc = n_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += y;
                }
                const uint32_t n_deep = __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (n_deep) {
                    uint32_t at = inc - n_mine;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if ((alive >> j) & 1u) {
                            q_node[at] = node[j];
                            q_pos[at] = (uint8_t)(lane * 8 + j);
                            ++at;
                        }
                    }
                    __syncwarp();
                    for (uint32_t k = lane; k < n_deep; k += 32) {
                        uint32_t nd = q_node[k], info = 0, bits = 0;
                        const uint32_t pos = q_pos[k];
                        const uint16_t *cp = w + 32 + pos;
                        for (int d = d_lock + 1; d <= max_len; ++d) {
                            const uint32_t c = cp[1 - d];
                            if (c == 0u || !trie_step(A, nd, c, info)) break;
                            if (info & kTerm) bits |= 1u << (32 - d);
                            if (!(info & kKids)) break;
                        }
                        q_bits[pos] = bits;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if ((alive >> j) & 1u) m[j] |= q_bits[lane * 8 + j];
                }
            }
            // ---- positions outside [emit_from, emit_to) report nothing
            if (p0 < P.emit_from || p0 + 8 > P.emit_to) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (p0 + j < P.emit_from || p0 + j >= P.emit_to) m[j] = 0u;
            }
            uint4 *mp = reinterpret_cast<uint4 *>(P.masks + ((size_t)(row0 + r) * kMaskRow + (size_t)lane * 8));
            mp[0] = make_uint4(m[0], m[1], m[2], m[3]);
            mp[1] = make_uint4(m[4], m[5], m[6], m[7]);
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) cnt += __popc(m[j]);
            const uint32_t row_total = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) P.row_count[row0 + r] = row_total;
            // ---- the last 32 classes of this row are the left context of the next
            const uint16_t keep = w[kMaskRow + lane];
            __syncwarp();
            w[lane] = keep;
        }
    }
}

// k_wide_tile: the mask kernel of dictionaries with a class-pair table (at most 64 classes) - generation 2 of the wide path.
//
// k_wide_mask is latency-bound: a warp's row waits for its longest walk, one dependent L2 gather per level, and a walk
// that dies pays a last gather (plus open-addressing probes) to learn it (ncu: 36 % of the lanes active, 10 long-scoreboard
// stalls per issue, 16 % of the L2 bandwidth).  Here a CTA owns a TILE of 4 096 positions:
//   * levels 1 and 2 of every position come from the pair table in shared memory, which also holds the level-2 node's
//     exact child mask: only walks whose next edge exists are pushed to the queue of the CTA;
//   * the queue is worked off in ROUNDS: every thread takes up to 8 walks, issues their gathers back to back and advances
//     each to its next branch point.  Edges are PATH-COMPRESSED and addressed without hashing (HostAutomaton::wide_chain:
//     one 32-byte entry = a child and the unbranched chain below it, the child mask of the chain's end and the index of
//     that node's first child entry; the next entry is first + popcount(mask below the class)), so a walk costs exactly
//     one gathered sector per branch point and never a gather that finds nothing;
//   * hits are OR-ed into the tile's masks in shared memory; masks and row counts leave with coalesced 128-bit stores.
// The number of rounds does not depend on the tile size, so a large tile amortises the latency of the thin late rounds,
// and the rounds of the two resident CTAs of an SM overlap.
constexpr int kWtThreads = 512;
constexpr int kWtTile = 4096;                       // positions per tile = 16 rows
constexpr int kWtRows = kWtTile / kMaskRow;
constexpr int kWtPer = kWtTile / kWtThreads;        // positions (and at most walks per round) per thread
constexpr int kWtIlp = 4;                           // gathers a thread keeps in flight
constexpr int kWtHalo = 40;                         // class bytes kept before the tile (walks look 31 back, the 8-byte compare 7 more)
constexpr int kWtMinRounds = 3;                     // rounds a tile always runs itself (default of WideArgs::min_rounds)
constexpr int kWtHandOver = 768;                    // walks left at which a tile hands them to k_wide_tail (default of WideArgs::hand_over)

__host__ __device__ constexpr size_t wide_tile_smem_bytes(int C) {
    // classes 0..255 | pair table (16 bytes per class pair) | class window (halo + tile bytes) | masks | queue (8 bytes per walk) | counters
    return 512 + (size_t)C * C * 16 + (kWtHalo + kWtTile + 8) + kWtTile * 4 + kWtTile * 8 + 32;
}

__device__ __forceinline__ uint32_t rank64(uint32_t lo, uint32_t hi, uint32_t c) {  // set bits of (hi:lo) below bit c
    return c < 32u ? (uint32_t)__popc(lo & ((1u << c) - 1u)) : (uint32_t)__popc(lo) + (uint32_t)__popc(hi & ((1u << (c - 32u)) - 1u));
}
__device__ __forceinline__ bool bit64(uint32_t lo, uint32_t hi, uint32_t c) { return ((c < 32u ? lo >> c : hi >> (c - 32u)) & 1u) != 0u; }

__global__ void __launch_bounds__(kWtThreads, 2) k_wide_tile(const DevAutomaton A, const DevWide Wd, const WideArgs P) {
    extern __shared__ __align__(16) unsigned char s_wide[];
    uint16_t *s_cls8 = reinterpret_cast<uint16_t *>(s_wide);
    const uint4 *s_pair = reinterpret_cast<const uint4 *>(s_wide + 512);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = Wd.C;
    unsigned char *s_after = s_wide + 512 + (size_t)C * C * 16;
    uint8_t *w = s_after;                                                       // w[kWtHalo + p] = class of tile position p (at most 64 classes)
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_after + (kWtHalo + kWtTile + 8));
    uint2 *s_q = reinterpret_cast<uint2 *>(s_mask + kWtTile);                   // {entry to load, position | depth << 16}
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_q + kWtTile);              // [0], [1]: walks queued for even / odd rounds; [2]: tile; [3]: hand-over slot
    for (int i = tid; i < 256; i += kWtThreads) s_cls8[i] = __ldg(&A.cls[i]);
    {
        uint4 *dst = reinterpret_cast<uint4 *>(s_wide + 512);
        for (int i = tid; i < C * C; i += kWtThreads) dst[i] = __ldg(&Wd.pair16[i]);
    }
    if (tid < 2) s_cnt[tid] = 0u;
    const int max_len = min(A.max_len, kWideMaxLen);
    const int64_t n_tiles = (P.n_rows + kWtRows - 1) / kWtRows;
    auto class_of = [&](uint32_t ch) -> uint32_t { return ch < 256u ? (uint32_t)s_cls8[ch] : (uint32_t)__ldg(&A.cls[ch]); };
    // a warp appends its lanes' walks (flag per lane) to the queue counted by *cnt
    auto push = [&](bool on, uint32_t entry, uint32_t pd, uint32_t *cnt) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, on);
        if (!bal) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cnt, (uint32_t)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (on) s_q[base + __popc(bal & ((1u << lane) - 1u))] = make_uint2(entry, pd);
    };

    bool tail_full = false;  // k_wide_tail's list had no room for a hand-over of this CTA
    while (true) {
        __syncthreads();  // the previous tile's masks are out, the tables are in
        if (tid == 0) s_cnt[2] = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const int64_t tile = s_cnt[2];
        if (tile >= n_tiles) break;
        const int64_t row0 = tile * kWtRows;
        const int64_t t_lo = P.origin + row0 * kMaskRow;
        // ---- classes of the tile and of the chars before it
        uint2 own;  // the thread's 8 classes, position j in byte j
        {
            const int64_t p0 = t_lo + (int64_t)tid * kWtPer;
            uint32_t c[8];
            if (p0 >= 0 && p0 + 8 <= P.n) {
                const uint4 v = ldcs_v4_if(P.hay + p0, true);
                c[0] = class_of(v.x & 0xFFFFu); c[1] = class_of(v.x >> 16); c[2] = class_of(v.y & 0xFFFFu); c[3] = class_of(v.y >> 16);
                c[4] = class_of(v.z & 0xFFFFu); c[5] = class_of(v.z >> 16); c[6] = class_of(v.w & 0xFFFFu); c[7] = class_of(v.w >> 16);
            } else {
#pragma unroll
                for (int j = 0; j < kWtPer; j++) {
                    const int64_t p = p0 + j;
                    c[j] = (p >= 0 && p < P.n) ? class_of(__ldg(&P.hay[p])) : 0u;
                }
            }
            own = make_uint2(c[0] | c[1] << 8 | c[2] << 16 | c[3] << 24, c[4] | c[5] << 8 | c[6] << 16 | c[7] << 24);
            *reinterpret_cast<uint2 *>(w + kWtHalo + tid * kWtPer) = own;  // kWtHalo is a multiple of 8
            if (tid < kWtHalo) {
                const int64_t p = t_lo - kWtHalo + tid;
                w[tid] = (p >= 0 && p < P.n) ? (uint8_t)class_of(__ldg(&P.hay[p])) : (uint8_t)0;
            }
        }
        __syncthreads();
        // ---- levels 1 and 2; a walk whose level-3 edge exists is queued for round 0
        {
            const uint32_t before = *reinterpret_cast<const uint32_t *>(w + kWtHalo + tid * kWtPer - 4);  // the 4 classes before the thread's
            uint32_t m[kWtPer];
#pragma unroll
            for (int j = 0; j < kWtPer; j++) {
                // classes of positions j, j - 1, j - 2
                auto cls_rel = [&](int r) -> uint32_t {
                    return r >= 4 ? (own.y >> (8 * (r - 4))) & 0xFFu : (r >= 0 ? (own.x >> (8 * r)) & 0xFFu : (before >> (8 * (4 + r))) & 0xFFu);
                };
                const uint32_t c0 = cls_rel(j), c1 = cls_rel(j - 1), c2 = cls_rel(j - 2);
                const uint4 e = s_pair[c0 * (uint32_t)C + c1];
                uint32_t mj = 0;
                if ((e.y >> 8) & kTerm) mj |= 1u << 31;
                if (((e.y >> 17) & 1u) && max_len >= 2 && (e.y & kTerm)) mj |= 1u << 30;
                const bool on = ((e.y >> 17) & 1u) && max_len >= 3 && bit64(e.z, e.w, c2);  // class 0 is in no mask
                m[j] = mj;
                push(on, e.x + rank64(e.z, e.w, c2), (uint32_t)(tid * kWtPer + j) | 2u << 16, &s_cnt[0]);
            }
            uint4 *mp = reinterpret_cast<uint4 *>(s_mask + tid * kWtPer);
            mp[0] = make_uint4(m[0], m[1], m[2], m[3]);
            mp[1] = make_uint4(m[4], m[5], m[6], m[7]);
        }
        __syncthreads();
        // ---- rounds
        for (uint32_t round = 0;; ++round) {
            const uint32_t n_q = s_cnt[round & 1u];
            if (n_q == 0u) break;
            // a thin round costs the whole CTA a gather latency: once few walks are left they are handed to k_wide_tail
            if (round >= P.min_rounds && n_q <= P.hand_over && P.tail_cap && !tail_full) {
                // one fetch-and-add per hand-over (a compare-and-swap loop convoys: 60 ms instead of 17 per 10^9 chars).  The
                // count is never taken back: a reservation that does not fit fills what it got of the list with EMPTY
                // entries (k_wide_tail skips them), and from then on every tile finishes its walks itself.
                if (tid == 0) s_cnt[3] = atomicAdd(P.tail_count, n_q);
                __syncthreads();
                const uint32_t at = s_cnt[3];
                if (at + n_q <= P.tail_cap && at + n_q >= at) {
                    const uint32_t pos0 = (uint32_t)(row0 * kMaskRow);
                    for (uint32_t k = tid; k < n_q; k += kWtThreads) {
                        const uint2 q = s_q[k];
                        P.tail[at + k] = make_uint4(q.x, pos0 + (q.y & 0xFFFFu), q.y >> 16, 0u);
                    }
                    __syncthreads();
                    if (tid == 0) s_cnt[round & 1u] = 0u;
                    break;
                }
                for (uint32_t k = tid; k < n_q && at + k < P.tail_cap && at + k >= at; k += kWtThreads)
                    P.tail[at + k] = make_uint4(kNoneD, 0u, 0u, 0u);
                tail_full = true;  // the count only grows: this CTA does not ask again
            }
            uint2 mine[kWtPer];
#pragma unroll
            for (int i = 0; i < kWtPer; i++) {
                const uint32_t k = (uint32_t)tid + (uint32_t)i * kWtThreads;
                mine[i] = k < n_q ? s_q[k] : make_uint2(kNoneD, 0u);
            }
            __syncthreads();  // every walk of this round is in registers: the queue can take the survivors
            if (tid == 0) s_cnt[round & 1u] = 0u;
            uint32_t alive = 0;  // bit i: walk i of this thread goes on (its next entry and depth replace mine[i])
#pragma unroll
            for (int i0 = 0; i0 < kWtPer; i0 += kWtIlp) {
                if ((uint32_t)(i0 * kWtThreads) >= n_q) break;  // uniform over the CTA
                uint4 ea[kWtIlp], eb[kWtIlp];
#pragma unroll
                for (int i = 0; i < kWtIlp; i++) {
                    const uint2 q = mine[i0 + i];
                    ea[i] = make_uint4(0u, 0u, 0u, 0u);
                    eb[i] = ea[i];
                    if (q.x != kNoneD) {
                        ea[i] = __ldg(Wd.chain + (size_t)q.x * 2);
                        eb[i] = __ldg(Wd.chain + (size_t)q.x * 2 + 1);  // the same sector
                    }
                }
#pragma unroll
                for (int i = 0; i < kWtIlp; i++) {
                    const uint2 q = mine[i0 + i];
                    if (q.x == kNoneD) continue;
                    const int pos = (int)(q.y & 0xFFFFu);
                    int d = (int)(q.y >> 16) + 1;  // depth of the entry's child
                    const uint4 e = ea[i];
                    const int L = (int)(e.w & 15u);
                    const uint32_t term = e.w >> 4;
                    // the L chain classes against the context, all at once: the 8 class bytes that end at the class step 0
                    // must equal, most recent in the top byte - the order the entry keeps its chain in
                    const int a7 = kWtHalo + pos - d - 7;  // >= 2
                    const uint32_t *wp = reinterpret_cast<const uint32_t *>(w + (a7 & ~3));
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], sh = (uint32_t)(a7 & 3) * 8u;
                    const uint32_t x_lo = __funnelshift_r(w0, w1, sh) ^ eb[i].x, x_hi = __funnelshift_r(w1, w2, sh) ^ eb[i].y;
                    const int mtc = min(L, (x_hi ? __clz((int)x_hi) : 32 + __clz((int)x_lo)) >> 3);  // steps that match before the first that does not
                    // terminal flags of the child and of the matched steps -> hit bits 32 - d, 32 - d - 1, ..
                    const uint32_t t_ok = term & ((2u << mtc) - 1u);
                    const uint32_t hits = (__brev(t_ok) >> (d - 1));
                    if (hits) atomicOr(&s_mask[pos], hits);
                    d += mtc;
                    if (mtc == L && d < max_len) {
                        const uint32_t c = w[kWtHalo + pos - d];
                        if (bit64(e.x, e.y, c)) {
                            alive |= 1u << (i0 + i);
                            mine[i0 + i] = make_uint2(e.z + rank64(e.x, e.y, c), (uint32_t)pos | (uint32_t)d << 16);
                        }
                    }
                }
            }
            // ---- survivors -> queue of the next round: one reservation per warp
            {
                const uint32_t n_mine = (uint32_t)__popc(alive);
                uint32_t inc = n_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += y;
                }
                const uint32_t n_warp = __shfl_sync(0xFFFFFFFFu, inc, 31);
                uint32_t base = 0;
                if (lane == 31 && n_warp) base = atomicAdd(&s_cnt[(round + 1u) & 1u], n_warp);
                base = __shfl_sync(0xFFFFFFFFu, base, 31) + inc - n_mine;
#pragma unroll
                for (int i = 0; i < kWtPer; i++)
                    if ((alive >> i) & 1u) s_q[base++] = mine[i];
            }
            __syncthreads();  // the survivors are queued
        }
        // ---- masks and row counts leave the tile
        {
            const int64_t p0 = t_lo + (int64_t)tid * kWtPer;
            uint32_t m[kWtPer];
#pragma unroll
            for (int j = 0; j < kWtPer; j++) m[j] = s_mask[tid * kWtPer + j];
            if (p0 < P.emit_from || p0 + kWtPer > P.emit_to) {
#pragma unroll
                for (int j = 0; j < kWtPer; j++)
                    if (p0 + j < P.emit_from || p0 + j >= P.emit_to) m[j] = 0u;
            }
            const int64_t row = row0 + warp;  // a warp's 256 positions are one row
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < kWtPer; j++) cnt += __popc(m[j]);
            const uint32_t row_total = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (row < P.n_rows) {
                uint4 *mp = reinterpret_cast<uint4 *>(P.masks + ((size_t)row * kMaskRow + (size_t)lane * kWtPer));
                mp[0] = make_uint4(m[0], m[1], m[2], m[3]);
                mp[1] = make_uint4(m[4], m[5], m[6], m[7]);
                if (lane == 0) P.row_count[row] = row_total;
            }
        }
    }
}

// k_wide_tail: the walks k_wide_tile handed over, one per lane, refilled from the list as lanes finish.  A step = the
// walk's entry (one sector) + the up to 9 haystack chars it is compared with; hits are OR-ed into the masks the tile
// already stored and added to the row counts (k_row_scan runs after this kernel).
__global__ void __launch_bounds__(256) k_wide_tail(const DevAutomaton A, const DevWide Wd, const WideArgs P) {
    __shared__ uint16_t s_cls8[256];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 256; i += 256) s_cls8[i] = __ldg(&A.cls[i]);
    __syncthreads();
    const int max_len = min(A.max_len, kWideMaxLen);
    const uint32_t n_tail = min(P.tail_count[0], P.tail_cap);
    auto class_at = [&](int64_t p) -> uint32_t {
        if (p < 0 || p >= P.n) return 0u;
        const uint32_t ch = __ldg(&P.hay[p]);
        return ch < 256u ? (uint32_t)s_cls8[ch] : (uint32_t)__ldg(&A.cls[ch]);
    };
    // every warp owns an equal slice of the list (no shared ticket: one atomic per refill convoys on a single address)
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5), wid = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    const uint32_t per = (n_tail + n_warps - 1u) / n_warps;
    uint32_t next = min(n_tail, wid * per);
    const uint32_t last = min(n_tail, next + per);
    bool have = false;
    uint32_t entry = 0, pos = 0;
    int d = 0;
    while (true) {
        // ---- lanes without a walk take the next ones of the warp's slice
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !have);
        if (need) {
            if (!have) {
                const uint32_t k = next + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (k < last) {
                    const uint4 t = P.tail[k];
                    entry = t.x;
                    pos = t.y;
                    d = (int)t.z;
                    have = t.x != kNoneD;  // an EMPTY entry: the slot of a hand-over that did not fit
                }
            }
            next = min(last, next + (uint32_t)__popc(need));
            if (!__ballot_sync(0xFFFFFFFFu, have)) break;  // the slice is empty and every lane is done
        }
        if (!have) continue;
        const uint4 e = __ldg(Wd.chain + (size_t)entry * 2), ch = __ldg(Wd.chain + (size_t)entry * 2 + 1);
        const int64_t q = P.origin + (int64_t)pos;  // haystack position of the walk's END anchor
        d += 1;
        const int L = (int)(e.w & 15u);
        const uint32_t term = e.w >> 4;
        uint32_t same = 0;
#pragma unroll
        for (int k = 0; k < kWideChainMaxD; k++) {
            const uint32_t want = ((k < 4 ? ch.y >> (8 * (3 - k)) : ch.x >> (8 * (7 - k))) & 0xFFu);  // step k in byte 7 - k
            same |= (k < L && class_at(q - d - k) == want) ? 1u << k : 0u;
        }
        const int mtc = __ffs(~same) - 1;
        const uint32_t t_ok = term & ((2u << mtc) - 1u);
        const uint32_t hits = (q >= P.emit_from && q < P.emit_to) ? (__brev(t_ok) >> (d - 1)) : 0u;
        if (hits) {
            atomicOr(P.masks + pos, hits);
            atomicAdd(P.row_count + (pos >> 8), (uint32_t)__popc(hits));
        }
        d += mtc;
        have = false;
        if (mtc == L && d < max_len) {
            const uint32_t c = class_at(q - d);
            if (bit64(e.x, e.y, c)) {
                entry = e.z + rank64(e.x, e.y, c);
                have = true;
            }
        }
    }
}

// value index of the keyword of length d whose last char is position q (Maps).  With the keyword-hash table: hash the d
// classes, one gathered bucket (the record is a real match, so the key is there and (hash, length) names it); without:
// the walk again, d steps.
__device__ __forceinline__ uint32_t wide_value(const DevAutomaton &A, const DevWide &Wd, const uint16_t *hay, int64_t q, int d) {
    if (Wd.vals) {
        WideValHash h;
        for (int i = 0; i < d; i++) h.add((uint32_t)__ldg(&A.cls[__ldg(&hay[q - i])]));
        const unsigned long long x = h.finish((uint32_t)d);
        const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32), tag = (uint32_t)d | 0x80000000u;
        uint32_t bk = __umulhi(hi, Wd.n_vbuckets);
        for (uint32_t tries = 0; tries < Wd.n_vbuckets; tries++) {
            const uint4 e0 = __ldg(Wd.vals + (size_t)bk * 2), e1 = __ldg(Wd.vals + (size_t)bk * 2 + 1);
            if (e0.x == lo && e0.y == hi && e0.z == tag) return e0.w;
            if (e1.x == lo && e1.y == hi && e1.z == tag) return e1.w;
            bk = bk + 1u == Wd.n_vbuckets ? 0u : bk + 1u;
        }
        return kNoneD;
    }
    uint32_t node = 0, info = 0;
    for (int i = 0; i < d; i++) {
        if (!trie_step(A, node, (uint32_t)__ldg(&A.cls[__ldg(&hay[q - i])]), info)) return kNoneD;
    }
    return __ldg(&A.node_value[node]);
}

// code = row position << 5 | 32 - length  ->  (start, end); e_row = end of a keyword whose last char is row position 0
__device__ __forceinline__ int2 decode_rec32(uint32_t code, int32_t e_row) {
    const int32_t e = e_row + (int32_t)(code >> 5);
    return make_int2(e - 32 + (int32_t)(code & 31u), e);
}

// Masks -> records, the 32-bit twin of k_tier_emit (kernel_emit.cuh): a warp takes rows round-robin, every lane expands the
// bits of its 8 positions into 16-bit codes in shared memory (code = index of the bit in the row's 8 192-bit mask), then
// the warp decodes two codes per lane into one 16-byte streaming store.
template <bool kIsMap>
__global__ void __launch_bounds__(kEmitWarps * 32) k_wide_emit(const DevAutomaton A, const DevWide Wd, const EmitArgs E) {
    __shared__ __align__(16) int2 s_stage_all[kEmitWarps][kEmitStage + 2];
    __shared__ uint32_t s_val_all[kIsMap ? kEmitWarps : 1][kIsMap ? kEmitStage : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int2 *s_stage = s_stage_all[warp];
    const unsigned short *s_code = reinterpret_cast<const unsigned short *>(s_stage);
    uint32_t *s_val = s_val_all[kIsMap ? warp : 0];
    const int64_t stride = (int64_t)gridDim.x * kEmitWarps;
    const uint32_t stage_sa = (uint32_t)__cvta_generic_to_shared(s_stage);
    const uint32_t out_par = (uint32_t)(reinterpret_cast<uintptr_t>(E.pos_out) >> 3) & 1u;

    for (int64_t row = (int64_t)blockIdx.x * kEmitWarps + warp; row < E.n_rows; row += stride) {
        const uint4 *mrow = reinterpret_cast<const uint4 *>(E.masks + ((size_t)row * kMaskRow + (size_t)lane * 8));
        const uint4 ma = __ldcs(mrow), mb = __ldcs(mrow + 1);
        const unsigned long long base = __ldg(E.block_excl + row / kScanRows) + __ldg(E.row_excl + row);
        const uint32_t words[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) cnt += __popc(words[j]);
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (total == 0) continue;
        const uint32_t my_off = inc - cnt;
        const int64_t q_row = E.origin + row * kMaskRow;                       // position of the row's first char
        const int32_t e_row = (int32_t)q_row + 1 + E.pos_base;
        __syncwarp();
        if (total <= (uint32_t)kEmitStage && base + total <= (unsigned long long)E.cap) {
            const uint32_t par = ((uint32_t)base + out_par) & 1u;
            uint32_t sa = stage_sa + (my_off + par) * 2u;
            uint32_t cb = (uint32_t)lane * 256u;
#pragma unroll
            for (int wi = 0; wi < 8; wi++) {
                uint32_t wv = words[wi];
                while (wv) {
                    const uint32_t t = (uint32_t)__clz((int)__brev(wv));
                    wv &= wv - 1u;
                    sts_u16(sa, cb + t);
                    sa += 2u;
                }
                cb += 32u;
            }
            __syncwarp();
            if (kIsMap) {
                for (uint32_t r = lane; r < total; r += 32) {
                    const uint32_t code = s_code[r + par];
                    __stcs(E.val_out + base + r, wide_value(A, Wd, E.hay, q_row + (int64_t)(code >> 5), 32 - (int)(code & 31u)));
                }
            }
            int2 *g = E.pos_out + (base - par);  // 16-byte aligned
            const uint32_t end = par + total, k_hi = end >> 1;
            if (lane == 0 && par) __stcs(g + 1, decode_rec32(s_code[1], e_row));
            if (lane == 1 && (end & 1u)) __stcs(g + (end - 1u), decode_rec32(s_code[end - 1u], e_row));
            int4 *gp = reinterpret_cast<int4 *>(g) + par + lane;
            const uint32_t *sp = reinterpret_cast<const uint32_t *>(s_code) + par + lane;
            for (uint32_t k = par + lane; k < k_hi; k += 32, gp += 32, sp += 32) {
                const uint32_t cc = *sp;
                const int2 r0 = decode_rec32(cc & 0xFFFFu, e_row), r1 = decode_rec32(cc >> 16, e_row);
                __stcs(gp, make_int4(r0.x, r0.y, r1.x, r1.y));
            }
            __syncwarp();
            continue;
        }
        // ---- dense rows (more than kEmitStage records) and rows that cross the caller's capacity: windows of kEmitStage
        for (uint32_t win = 0; win < total; win += kEmitStage) {
            if (cnt && my_off < win + kEmitStage && my_off + cnt > win) {
                uint32_t o = my_off - win;  // wraps below zero for records of an earlier window
#pragma unroll
                for (int wi = 0; wi < 8; wi++) {
                    uint32_t wv = words[wi];
                    while (wv) {
                        const int t = __ffs(wv) - 1;
                        wv &= wv - 1u;
                        if (o < (uint32_t)kEmitStage) {
                            const int32_t e = e_row + lane * 8 + wi;
                            s_stage[o] = make_int2(e - (32 - t), e);
                            if (kIsMap) s_val[o] = wide_value(A, Wd, E.hay, q_row + lane * 8 + wi, 32 - t);
                        }
                        ++o;
                    }
                }
            }
            __syncwarp();
            const uint32_t n_win = min((uint32_t)kEmitStage, total - win);
            const unsigned long long g0 = base + win;
            const unsigned long long room = g0 < (unsigned long long)E.cap ? (unsigned long long)E.cap - g0 : 0ull;
            const uint32_t n_out = (uint32_t)min((unsigned long long)n_win, room);
            for (uint32_t rr = lane; rr < n_out; rr += 32) {
                __stcs(&E.pos_out[g0 + rr], s_stage[rr]);
                if (kIsMap) __stcs(&E.val_out[g0 + rr], s_val[rr]);
            }
            __syncwarp();
        }
    }
}

}  // namespace acgpu
