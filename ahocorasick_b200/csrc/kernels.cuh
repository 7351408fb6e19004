// sm_100a kernels of the matching path (generation 1: anchored trie walks + ordered single-pass emission).
//
//   k_ac_scan      AhoCorasick family.  Every haystack position q is an END anchor: walk the reversed-keyword
//                  trie backwards from q; every terminal node met is one match (q+1-depth, q+1).  Two walks
//                  (count, then write) around a warp/block scan and a decoupled look-back across tiles give the
//                  reference's listener order (end ascending, longest first — AhoCorasickSet.java:522-535)
//                  in ONE pass over the haystack with no sort and no global count pass.
//   k_fwd_v        Longest / Shortest / WholeWord.  Every position s is a START anchor: walk the forward trie and
//                  record v[s] = longest keyword at s (Longest), first/shortest keyword at s (Shortest), or the
//                  word length if the whole word-char run starting at s is a keyword (WholeWord), else 0.
//   k_sel_*        Non-overlapping selection (LongestMatchSet.java:192-265 + SetMatchQueue.java:45-95,
//                  ShortestMatchSet.java:182-260): the sequential left-to-right choice is a chain
//                  pos -> J(pos) with J local to a window of v[]; tiles resolve their part of the chain for every
//                  possible entry offset by pointer doubling in shared memory ("exit maps"), maps are composed
//                  over groups of tiles, then every tile marks the positions the true chain visits and emits them
//                  in order (same look-back compaction).
#pragma once
#include "device_tables.cuh"

namespace acgpu {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------ small block primitives

__device__ __forceinline__ uint32_t warp_inclusive_sum(uint32_t x) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// exclusive prefix sum over the block's threads (thread order), also returns the block total.
// s_tmp: kWarps + 1 words of shared memory; contains a barrier pair.
__device__ __forceinline__ uint32_t block_exclusive_sum(uint32_t x, uint32_t *s_tmp, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_sum(x);
    if (lane == 31) s_tmp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        uint32_t t = s_tmp[w];
        if (w < warp) wbase += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return wbase + inc - x;
}

// min over all threads with a HIGHER thread index (exclusive suffix min), identity = 0xFFFFFFFF.
__device__ __forceinline__ uint32_t block_suffix_min_exclusive(uint32_t x, uint32_t *s_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = x;  // inclusive suffix min within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_down_sync(0xFFFFFFFFu, inc, o);
        if (lane + o < 32) inc = min(inc, y);
    }
    if (lane == 0) s_tmp[warp] = inc;
    __syncthreads();
    uint32_t right = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        if (w > warp) right = min(right, s_tmp[w]);
    }
    __syncthreads();
    uint32_t excl = __shfl_down_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 31) excl = 0xFFFFFFFFu;
    return min(excl, right);
}

__device__ __forceinline__ uint32_t block_min(uint32_t x, uint32_t *s_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xFFFFFFFFu, x, o));
    if (lane == 0) s_tmp[warp] = x;
    __syncthreads();
    uint32_t r = 0xFFFFFFFFu;
#pragma unroll
    for (int w = 0; w < kWarps; w++) r = min(r, s_tmp[w]);
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------ AhoCorasick: end-anchored scan

constexpr int kAcTile = 4096;                  // end positions per tile
constexpr int kAcRows = kAcTile / kThreads;    // 16 rows of 32 per warp
constexpr int kAcHaloMax = 2048;               // classes staged to the left of a tile


template <bool kIsMap>
__global__ void __launch_bounds__(kThreads) k_ac_scan(const DevAutomaton A, const AcArgs P) {
    __shared__ uint16_t s_cls[kAcTile + kAcHaloMax];
    __shared__ uint32_t s_warp_tot[kWarps];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int halo = min(A.max_len > 0 ? A.max_len - 1 : 0, kAcHaloMax);

    while (true) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t q_lo = P.emit_from + tile * kAcTile;
        const int64_t q_hi = min(P.emit_to, q_lo + kAcTile);
        const int64_t win_lo = max((int64_t)0, q_lo - halo);
        for (int64_t i = win_lo + tid; i < q_hi; i += kThreads) s_cls[i - win_lo] = __ldg(&A.cls[__ldg(&P.hay[i])]);
        __syncthreads();

        auto cls_at = [&](int64_t i) -> uint32_t {
            return i >= win_lo ? (uint32_t)s_cls[i - win_lo] : (uint32_t)__ldg(&A.cls[__ldg(&P.hay[i])]);
        };

        // ---- walk 1: count keywords ending at each position
        uint32_t cnt[kAcRows], off[kAcRows];
        uint32_t run = 0;
#pragma unroll
        for (int j = 0; j < kAcRows; j++) {
            const int64_t q = q_lo + warp * (kAcRows * 32) + j * 32 + lane;
            uint32_t c_hits = 0;
            if (q < q_hi) {
                uint32_t node = 0, info = 0;
                const int64_t lim = max((int64_t)0, q - A.max_len + 1);
                for (int64_t i = q; i >= lim; --i) {
                    uint32_t c = cls_at(i);
                    if (A.has_other && c == 0) break;
                    if (!trie_step_sig(A, node, c, info)) break;
                    c_hits += info & kTerm;
                    if (!(info & kKids)) break;
                }
            }
            cnt[j] = c_hits;
            uint32_t inc = warp_inclusive_sum(c_hits);
            off[j] = run + inc - c_hits;
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) s_warp_tot[warp] = run;
        __syncthreads();
        uint32_t warp_base = 0, block_total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            uint32_t t = s_warp_tot[w];
            if (w < warp) warp_base += t;
            block_total += t;
        }
        if (warp == 0) {
            unsigned long long excl = lookback_exclusive(P.status, tile, block_total);
            if (lane == 0) {
                s_base = excl;
                if (tile == P.n_tiles - 1) *P.total_out = excl + block_total;
            }
        }
        __syncthreads();
        const unsigned long long base = s_base + warp_base;

        // ---- walk 2: write (start, end[, value]) longest first
#pragma unroll
        for (int j = 0; j < kAcRows; j++) {
            if (cnt[j] == 0) continue;
            const int64_t q = q_lo + warp * (kAcRows * 32) + j * 32 + lane;
            uint32_t node = 0, info = 0, k = 0;
            const int64_t lim = max((int64_t)0, q - A.max_len + 1);
            const unsigned long long last = base + off[j] + cnt[j] - 1;
            for (int64_t i = q; i >= lim; --i) {
                uint32_t c = cls_at(i);
                if (A.has_other && c == 0) break;
                if (!trie_step_sig(A, node, c, info)) break;
                if (info & kTerm) {
                    unsigned long long idx = last - k;
                    if (idx < (unsigned long long)P.cap) {
                        P.pos_out[idx] = make_int2((int32_t)i + P.pos_base, (int32_t)(q + 1) + P.pos_base);
                        if (kIsMap) P.val_out[idx] = __ldg(&A.node_value[node]);
                    }
                    ++k;
                }
                if (!(info & kKids)) break;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ forward anchored walk: v[s]

constexpr int kFwTile = 4096;
constexpr int kFwHaloMax = 2048;

struct FwArgs {
    const uint16_t *hay;
    int64_t n;        // chars available in the window; treated as end of input by the walks
    int64_t p_lo;     // compute v for start positions [p_lo, p_hi)
    int64_t p_hi;
    uint16_t *v;      // v[s], indexed by window position
    int64_t abs0;     // WholeWordLongest: window position of the first char of the input (always a walk start), -1 = not here
};

template <int kFamily>
__global__ void __launch_bounds__(kThreads) k_fwd_v(const DevAutomaton A, const FwArgs P) {
    __shared__ uint16_t s_cls[kFwTile + kFwHaloMax];
    const int tid = threadIdx.x;
    const int halo = min(A.max_len, kFwHaloMax);
    const int64_t n_tiles = (P.p_hi - P.p_lo + kFwTile - 1) / kFwTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s_lo = P.p_lo + tile * kFwTile;
        const int64_t s_hi = min(P.p_hi, s_lo + kFwTile);
        const int64_t win_hi = min(P.n, s_hi + halo);
        __syncthreads();
        for (int64_t i = s_lo + tid; i < win_hi; i += kThreads) s_cls[i - s_lo] = __ldg(&A.cls[__ldg(&P.hay[i])]);
        __syncthreads();
        auto cls_at = [&](int64_t i) -> uint32_t {
            return i < win_hi ? (uint32_t)s_cls[i - s_lo] : (uint32_t)__ldg(&A.cls[__ldg(&P.hay[i])]);
        };
        for (int64_t s = s_lo + tid; s < s_hi; s += kThreads) {
            uint32_t best = 0;
            if (kFamily == 3) {
                // WholeWord: s must start a maximal run of word characters (WholeWordMatchSet.java:58-93)
                bool start = is_word_char(A, __ldg(&P.hay[s])) && (s == 0 || !is_word_char(A, __ldg(&P.hay[s - 1])));
                if (start) {
                    uint32_t node = 0, info = 0;
                    int64_t i = s;
                    while (true) {
                        if (i >= P.n || !is_word_char(A, __ldg(&P.hay[i]))) {
                            best = (info & kTerm) ? (uint32_t)(i - s) : 0u;
                            break;
                        }
                        uint32_t c = cls_at(i);
                        if ((A.has_other && c == 0) || !trie_step_sig(A, node, c, info)) break;
                        ++i;
                    }
                }
            } else if (kFamily == 4) {
                // WholeWordLongest (WholeWordLongestMatchSet.java:47-182): a walk starts at the first char of the input
                // and at every word start; it follows the trie over word AND non-word chars until there is no
                // transition (position idx).  Reported: the longest keyword on the path that is followed by a non-word
                // char or the end of the input (the node's own match or its carried "fail match", :224-240).  The next
                // walk starts at the first word start after idx.  v = length | (idx + 1 - s) << 8.
                const bool start = s == P.abs0 ||
                                   (is_word_char(A, __ldg(&P.hay[s])) && s > 0 && !is_word_char(A, __ldg(&P.hay[s - 1])));
                if (start) {
                    uint32_t node = 0, info = 0, len = 0;
                    int64_t i = s;
                    while (i < P.n) {
                        const uint32_t c = cls_at(i);
                        if ((A.has_other && c == 0) || !trie_step_sig(A, node, c, info)) break;
                        ++i;
                        if ((info & kTerm) && (i == P.n || !is_word_char(A, __ldg(&P.hay[i])))) len = (uint32_t)(i - s);
                        if (!(info & kKids)) break;
                    }
                    // i = first position without a transition (a leaf has none for any char)
                    best = len | ((uint32_t)(i - s + 1) << 8);
                }
            } else {
                uint32_t node = 0, info = 0;
                const int64_t lim = min(P.n, s + A.max_len);
                for (int64_t i = s; i < lim;) {
                    uint32_t c = cls_at(i);
                    if (A.has_other && c == 0) break;
                    if (!trie_step_sig(A, node, c, info)) break;
                    ++i;
                    if (info & kTerm) {
                        best = (uint32_t)(i - s);
                        if (kFamily == 2) break;  // Shortest: first keyword met
                    }
                    if (!(info & kKids)) break;
                }
            }
            P.v[s] = (uint16_t)best;
        }
    }
}

// ------------------------------------------------------------------ non-overlapping selection

constexpr int kSelTile = 2048;                 // chain positions per tile
constexpr int kSelPer = kSelTile / kThreads;   // 8 consecutive positions per thread
constexpr int kSelLevels = 11;                 // 2^11 = kSelTile
constexpr int kSelGroup = 256;                 // tiles per composition group
constexpr int kSelMaxLen = 2048;               // longest keyword the selection kernels accept
constexpr uint32_t kSkip = 0xFFFFu;

enum { kModeLongest = 1, kModeShortest = 2, kModeWholeWord = 3, kModeWholeWordLongest = 4 };

struct SelArgs {
    const uint16_t *v;   // per-start values, window positions [0, n_v)
    int64_t n_v;
    int64_t n;           // chain domain [0, n): chain positions >= n are not emitted (carried over)
    int32_t M;           // number of entry offsets per tile (> largest possible entry offset) = max_len + 1
    int32_t halo;        // Shortest: candidates up to halo positions right of a tile can win (= max_len - 1)
    int32_t mode;
    int64_t dom_lo;      // positions < dom_lo are left context only (streaming), never emitted
    int64_t n_tiles;
    uint16_t *exit1;     // [n_tiles * M]  tile exit offset (relative to nominal tile end) per entry offset
    int64_t n_groups;
    uint16_t *exit2;     // [n_groups * M]
    int64_t *entry2;     // [n_groups] absolute entry position of each group, -1 = skipped
    int32_t *entry1;     // [n_tiles] entry offset inside each tile, -1 = skipped
    int64_t entry0;      // chain position at the start (0 for a fresh haystack)
    const int32_t *wpos; // WholeWordLongest over COMPACTED walk starts (k_wwl_starts): chain position i stands for haystack position wpos[i]; nullptr: the chain runs over haystack positions
    long long *carry_out;  // [0] first chain position >= n (left untouched when the chain never crosses n)
    // emission
    const uint16_t *hay;
    int64_t n_hay;
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
    unsigned long long *total_out;
    unsigned int *tile_counter;
    unsigned long long *status;
};

// Fills s_nxt[p] (next chain position if the chain stands at tile position p, in [p+1, kSelTile + M)) and
// s_st[p] (tile-relative start of the match emitted from p, or kSkip when p only skips to the tile end).
// s_v needs kSelTile entries; s_tmp kWarps+1 words.
__device__ __forceinline__ void sel_build_nxt(const SelArgs &P, int64_t tile_start, uint16_t *s_v, uint16_t *s_nxt,
                                              uint16_t *s_st, uint32_t *s_tmp) {
    const int tid = threadIdx.x;
    const int p0 = tid * kSelPer;
    uint32_t vloc[kSelPer];
#pragma unroll
    for (int k = 0; k < kSelPer; k++) {
        int64_t g = tile_start + p0 + k;
        vloc[k] = (g < P.n_v) ? (uint32_t)__ldg(&P.v[g]) : 0u;
        s_v[p0 + k] = (uint16_t)vloc[k];
    }
    if (P.mode == kModeLongest || P.mode == kModeWholeWordLongest) {
        // first start >= p with a keyword (suffix "min position with v > 0"); WholeWordLongest: first walk start >= p,
        // v = reported length | jump << 8 (the chain moves by the jump, the record ends at start + length)
        const bool wwl = P.mode == kModeWholeWordLongest;
        uint32_t mine = 0xFFFFFFFFu;
#pragma unroll
        for (int k = kSelPer - 1; k >= 0; k--) {
            if (vloc[k]) mine = p0 + k;
        }
        uint32_t right = block_suffix_min_exclusive(mine, s_tmp);  // barrier inside: s_v now visible
        uint32_t fs = right;
#pragma unroll
        for (int k = kSelPer - 1; k >= 0; k--) {
            if (vloc[k]) fs = p0 + k;
            if (fs == 0xFFFFFFFFu) {
                s_nxt[p0 + k] = kSelTile;
                s_st[p0 + k] = kSkip;
            } else {
                s_nxt[p0 + k] = (uint16_t)(fs + (wwl ? (uint32_t)s_v[fs] >> 8 : (uint32_t)s_v[fs]));
                s_st[p0 + k] = (uint16_t)fs;
            }
        }
    } else {
        // earliest end among starts >= p, leftmost start on ties: suffix-min of (end << 16 | start)
        uint32_t mine = 0xFFFFFFFFu;
#pragma unroll
        for (int k = kSelPer - 1; k >= 0; k--) {
            if (vloc[k]) mine = min(mine, ((uint32_t)(p0 + k) + vloc[k]) << 16 | (uint32_t)(p0 + k));
        }
        uint32_t right = block_suffix_min_exclusive(mine, s_tmp);
        // candidates to the right of the tile: only their overall minimum matters
        uint32_t hmin = 0xFFFFFFFFu;
        for (int h = tid; h < P.halo; h += kThreads) {
            int64_t g = tile_start + kSelTile + h;
            uint32_t vv = (g < P.n_v) ? (uint32_t)__ldg(&P.v[g]) : 0u;
            if (vv) hmin = min(hmin, ((uint32_t)(kSelTile + h) + vv) << 16 | (uint32_t)(kSelTile + h));
        }
        hmin = block_min(hmin, s_tmp);
        uint32_t suf = right;
#pragma unroll
        for (int k = kSelPer - 1; k >= 0; k--) {
            if (vloc[k]) suf = min(suf, ((uint32_t)(p0 + k) + vloc[k]) << 16 | (uint32_t)(p0 + k));
            if (suf == 0xFFFFFFFFu) {
                s_nxt[p0 + k] = kSelTile;
                s_st[p0 + k] = kSkip;
            } else {
                uint32_t key = min(suf, hmin);
                s_nxt[p0 + k] = (uint16_t)(key >> 16);
                s_st[p0 + k] = (uint16_t)(key & 0xFFFFu);
            }
        }
    }
    __syncthreads();
}

// pass 1: per-tile exit maps
__global__ void __launch_bounds__(kThreads) k_sel_map(const SelArgs P) {
    __shared__ uint16_t s_v[kSelTile];
    __shared__ uint16_t s_a[kSelTile];
    __shared__ uint16_t s_b[kSelTile];
    __shared__ uint16_t s_st[kSelTile];
    __shared__ uint32_t s_tmp[kWarps + 1];
    const int tid = threadIdx.x;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int64_t tile_start = tile * kSelTile;
        __syncthreads();
        sel_build_nxt(P, tile_start, s_v, s_a, s_st, s_tmp);
        uint16_t *cur = s_a, *nw = s_b;
        for (int r = 0; r < kSelLevels; r++) {
#pragma unroll
            for (int k = 0; k < kSelPer; k++) {
                int p = tid + k * kThreads;
                uint32_t a = cur[p];
                nw[p] = a < kSelTile ? cur[a] : (uint16_t)a;
            }
            __syncthreads();
            uint16_t *t = cur;
            cur = nw;
            nw = t;
        }
        for (int o = tid; o < P.M && o < kSelTile; o += kThreads) P.exit1[tile * P.M + o] = (uint16_t)(cur[o] - kSelTile);
    }
}

// pass 2: compose the tile maps of one group for every entry offset (one thread per offset)
__global__ void __launch_bounds__(kThreads) k_sel_group(const SelArgs P) {
    const int64_t g = blockIdx.x;
    const int64_t t_lo = g * kSelGroup, t_hi = min(P.n_tiles, t_lo + kSelGroup);
    const int64_t g_start = t_lo * kSelTile, g_end = t_hi * kSelTile;
    for (int o = threadIdx.x; o < P.M; o += kThreads) {
        int64_t cur = g_start + o;
        for (int64_t t = t_lo; t < t_hi && cur < g_end; t++) {
            const int64_t t_end = (t + 1) * kSelTile;
            if (cur < t_end) cur = t_end + __ldg(&P.exit1[t * P.M + (cur - t * kSelTile)]);
        }
        // cur >= g_end here unless the group is the last one and shorter than an entry offset
        P.exit2[g * P.M + o] = (uint16_t)(cur >= g_end ? cur - g_end : 0);
    }
}

// pass 3: walk the groups sequentially (tiny), one thread
__global__ void k_sel_top(const SelArgs P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t cur = P.entry0;
    for (int64_t g = 0; g < P.n_groups; g++) {
        const int64_t g_start = g * kSelGroup * (int64_t)kSelTile;
        const int64_t g_end = min(P.n_tiles, (g + 1) * kSelGroup) * (int64_t)kSelTile;
        if (cur < g_end) {
            P.entry2[g] = cur;
            cur = g_end + P.exit2[g * P.M + (cur - g_start)];
        } else {
            P.entry2[g] = -1;
        }
    }
}

// pass 4: per group, hand every tile its true entry offset
__global__ void k_sel_entries(const SelArgs P) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.n_groups) return;
    const int64_t t_lo = g * kSelGroup, t_hi = min(P.n_tiles, t_lo + kSelGroup);
    int64_t cur = P.entry2[g];
    for (int64_t t = t_lo; t < t_hi; t++) {
        const int64_t t_start = t * kSelTile, t_end = t_start + kSelTile;
        if (cur >= 0 && cur < t_end) {
            P.entry1[t] = (int32_t)(cur - t_start);
            cur = t_end + __ldg(&P.exit1[t * P.M + (cur - t_start)]);
        } else {
            P.entry1[t] = -1;
        }
    }
}

// pass 5: mark the chain inside every tile and emit the selected matches in order
template <bool kIsMap>
__global__ void __launch_bounds__(kThreads) k_sel_emit(const DevAutomaton A, const SelArgs P) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint16_t *s_lvl = reinterpret_cast<uint16_t *>(s_dyn);                 // [kSelLevels][kSelTile]
    uint16_t *s_v = s_lvl + kSelLevels * kSelTile;                         // [kSelTile]
    uint16_t *s_st = s_v + kSelTile;                                       // [kSelTile]
    unsigned char *s_mark = reinterpret_cast<unsigned char *>(s_st + kSelTile);  // [kSelTile]
    __shared__ uint32_t s_tmp[kWarps + 1];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    while (true) {
        if (tid == 0) {
            s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        }
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t tile_start = tile * kSelTile;
        const int p0 = tid * kSelPer;
        uint32_t my_cnt = 0;
        unsigned char mk[kSelPer];
        uint16_t *s_nxt = s_lvl;  // level 0

        if (P.mode == kModeWholeWord) {
            // no chain: every word start with v > 0 is a match (WholeWordMatchSet.java:58-93)
#pragma unroll
            for (int k = 0; k < kSelPer; k++) {
                int64_t g = tile_start + p0 + k;
                uint32_t vv = (g >= P.dom_lo && g < P.n && g < P.n_v) ? (uint32_t)__ldg(&P.v[g]) : 0u;
                s_nxt[p0 + k] = 0;
                s_st[p0 + k] = (uint16_t)(p0 + k);
                s_v[p0 + k] = (uint16_t)vv;
                mk[k] = vv != 0;
                my_cnt += mk[k];
            }
        } else {
            const int32_t entry = P.entry1[tile];
            sel_build_nxt(P, tile_start, s_v, s_nxt, s_st, s_tmp);
            // doubling tables: level r+1 = level r applied twice
            for (int r = 0; r + 1 < kSelLevels; r++) {
                const uint16_t *cur = s_lvl + r * kSelTile;
                uint16_t *nw = s_lvl + (r + 1) * kSelTile;
#pragma unroll
                for (int k = 0; k < kSelPer; k++) {
                    int p = tid + k * kThreads;
                    uint32_t a = cur[p];
                    nw[p] = a < kSelTile ? cur[a] : (uint16_t)a;
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < kSelPer; k++) s_mark[tid + k * kThreads] = 0;
            __syncthreads();
            if (entry >= 0 && tid == 0) s_mark[entry] = 1;
            __syncthreads();
            // positions reachable from the entry: apply jump tables from the largest stride down
            // (reads and writes of a round are separated by a barrier: a mark set in round r is only looked at from
            // round r - 1 on, so the set of marks after every round is deterministic)
            for (int r = kSelLevels - 1; r >= 0; r--) {
                const uint16_t *lv = s_lvl + r * kSelTile;
                uint32_t tgt[kSelPer];
#pragma unroll
                for (int k = 0; k < kSelPer; k++) {
                    int p = tid + k * kThreads;
                    tgt[k] = s_mark[p] ? (uint32_t)lv[p] : (uint32_t)kSelTile;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < kSelPer; k++) {
                    if (tgt[k] < kSelTile) s_mark[tgt[k]] = 1;
                }
                __syncthreads();
            }
            // Chain positions at/after the end of the chain domain are carried to the next block, not emitted.
            // Exactly one chain position p < n (in whichever tile) jumps to/over n: it defines the carry.  A jump that
            // only skips, or whose match would start at/after n, restarts at n (no candidate lies in [p, n)).
#pragma unroll
            for (int k = 0; k < kSelPer; k++) {
                int p = p0 + k;
                int64_t g = tile_start + p;
                bool on = s_mark[p] != 0 && g < P.n;
                bool emits = on && s_st[p] != kSkip && tile_start + (int64_t)s_st[p] < P.n;
                if (on && tile_start + (int64_t)s_nxt[p] >= P.n && P.carry_out) {
                    P.carry_out[0] = emits ? tile_start + (int64_t)s_nxt[p] : (long long)P.n;
                }
                // WholeWordLongest: a walk moves the chain even when it reports nothing
                if (P.mode == kModeWholeWordLongest && emits) emits = (s_v[s_st[p]] & 0xFFu) != 0u;
                mk[k] = emits;
                my_cnt += mk[k];
            }
        }

        uint32_t block_total;
        uint32_t my_off = block_exclusive_sum(my_cnt, s_tmp, block_total);
        if (warp == 0) {
            unsigned long long excl = lookback_exclusive(P.status, tile, block_total);
            if (lane == 0) {
                s_base = excl;
                if (tile == P.n_tiles - 1) {
                    *P.total_out = excl + block_total;
                }
            }
        }
        __syncthreads();
        unsigned long long idx = s_base + my_off;
#pragma unroll
        for (int k = 0; k < kSelPer; k++) {
            if (!mk[k]) continue;
            const int p = p0 + k;
            const int64_t st = P.wpos ? (int64_t)__ldg(&P.wpos[tile_start + s_st[p]]) : tile_start + s_st[p];
            const int64_t en = (P.mode == kModeWholeWord) ? st + s_v[p]
                               : (P.mode == kModeWholeWordLongest ? st + (s_v[s_st[p]] & 0xFFu) : tile_start + s_nxt[p]);
            if (idx < (unsigned long long)P.cap) {
                P.pos_out[idx] = make_int2((int32_t)st + P.pos_base, (int32_t)en + P.pos_base);
                if (kIsMap) {
                    // value of the keyword hay[st, en): re-walk its trie path
                    uint32_t node = 0, info = 0;
                    for (int64_t i = st; i < en; i++) trie_step(A, node, __ldg(&A.cls[__ldg(&P.hay[i])]), info);
                    P.val_out[idx] = __ldg(&A.node_value[node]);
                }
            }
            ++idx;
        }
        __syncthreads();
    }
}

constexpr size_t kSelEmitSmem = (size_t)(kSelLevels + 2) * kSelTile * sizeof(uint16_t) + kSelTile;

}  // namespace acgpu
