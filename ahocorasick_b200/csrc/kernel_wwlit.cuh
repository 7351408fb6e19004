// k_segments<0> (the former k_ww_literal): WholeWordMatchSet / Map for the one input class where "maximal runs of word chars" is not the reference's
// behaviour - a case-insensitive matcher whose word-character table is not closed under Character.toLowerCase (quirk Q7).
//
// The reference's loop (WholeWordMatchSet.java:47-132; Map :155-240, Readable :55-153 + scroll :325-339):
//     walk the trie from a walk start over the LOWER-CASED chars until there is no transition, at position idx;
//     if the lower-cased char at idx is no word char: report the node's keyword, if it has one; otherwise scroll over
//     word chars;  then scroll over non-word chars;  the next walk starts where that stops.
// Both scrolls test the RAW char in the String overloads and the lower-cased one in the Readable overload, and the trie
// can hold chars that are no word chars (a keyword is checked against the table before it is folded), so a walk may run
// across "non-word" chars.  This kernel keeps the loop as it is and finds parallelism at SYNCHRONISATION POINTS: a
// position t whose previous char is in no keyword (class 0 - no node has a transition for it), is no word char in either
// view, and whose own char is a word char in the scroll view.  Whatever the loop was doing when it reached t - 1 (a walk:
// it fails there and reports; the first scroll: it stops there; the second scroll: it goes on), its next walk starts at
// t.  One thread per synchronisation point runs the loop until it reaches the next one; position 0 starts the first
// segment.  In ordinary text every space or punctuation mark is such a point; a text without any is one segment.
//
// Two launches around k_row_scan: counts per 256-position row, then records at their final offsets (the segment is
// walked again - the matchers this serves are a corner case, exactness is the point).
#pragma once
#include "kernel_emit.cuh"

namespace acgpu {

struct WwLitArgs {
    const uint16_t *hay;
    int64_t n;
    int64_t n_rows;
    uint32_t *row_count;                    // counting launch: records of the segments that START in the row
    const uint32_t *row_excl;               // writing launch: exclusive prefix inside the row's scan block ...
    const unsigned long long *block_excl;   // ... and of the block
    int32_t pos_base;
    int32_t scroll_folded;                  // 1: the scroll loops test the lower-cased char (Readable overloads)
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
};

__device__ __forceinline__ bool bit_of(const uint32_t *bits, uint32_t c) { return (__ldg(&bits[c >> 5]) >> (c & 31u)) & 1u; }

struct WwLitView {
    const DevAutomaton &A;
    const WwLitArgs &P;
    // wordbits_fold is only there when the two views differ
    __device__ __forceinline__ bool word_fold(int64_t i) const { return bit_of(A.wordbits_fold ? A.wordbits_fold : A.wordbits, __ldg(&P.hay[i])); }
    __device__ __forceinline__ bool word_scroll(int64_t i) const {
        return bit_of(P.scroll_folded && A.wordbits_fold ? A.wordbits_fold : A.wordbits, __ldg(&P.hay[i]));
    }
    __device__ __forceinline__ uint32_t cls(int64_t i) const { return __ldg(&A.cls[__ldg(&P.hay[i])]); }
    // a walk starts at t in every execution of the loop
    __device__ __forceinline__ bool sync_start(int64_t t) const {
        if (t == 0) return true;
        return A.has_other && cls(t - 1) == 0u && !word_scroll(t - 1) && !word_fold(t - 1) && word_scroll(t);
    }
};

// The loop from walk start s to the next synchronisation point (or the end of the input).  Returns the number of
// records; kWrite: record k goes to slot first + k.
template <bool kWrite, bool kIsMap>
__device__ __forceinline__ uint32_t wwlit_segment(const WwLitView &V, int64_t s, unsigned long long first) {
    const DevAutomaton &A = V.A;
    const WwLitArgs &P = V.P;
    uint32_t count = 0;
    auto report = [&](int64_t from, int64_t to, uint32_t node) {
        if (kWrite) {
            const unsigned long long at = first + count;
            if (at < (unsigned long long)P.cap) {
                P.pos_out[at] = make_int2((int32_t)from + P.pos_base, (int32_t)to + P.pos_base);
                if (kIsMap) P.val_out[at] = __ldg(&A.node_value[node]);
            }
        }
        ++count;
    };
    int64_t idx = s, walk = s;   // walk: where the current walk started (a node's keyword length = its depth = idx - walk)
    uint32_t node = 0, info = 0;
    while (idx < P.n) {
        const uint32_t c = V.cls(idx);
        uint32_t nx = node, ni = 0;
        const bool step = !(A.has_other && c == 0u) && trie_step(A, nx, c, ni);
        if (step) {
            ++idx;
            node = nx;
            info = ni;
            continue;
        }
        if (!V.word_fold(idx)) {
            if (node != 0u && (info & kTerm)) report(walk, idx, node);
        } else {
            while (++idx < P.n && V.word_scroll(idx)) {
            }
        }
        while (++idx < P.n && !V.word_scroll(idx)) {
        }
        node = 0;
        info = 0;
        walk = idx;
        if (idx < P.n && V.sync_start(idx)) return count;  // the next segment's thread goes on from here
    }
    if (idx == P.n && node != 0u && (info & kTerm)) report(walk, idx, node);  // the scrolls can leave idx at n + 1: nothing is pending then
    return count;
}

// ---------------------------------------------------------------- literal loops for keywords the selection kernels cannot hold
//
// The generation-1 selection kernels keep one exit entry per possible keyword length in shared memory (Longest / Shortest:
// keywords up to 2 047 chars) and pack a walk into 16 bits (WholeWordLongest: up to 254 chars).  Longer keywords - legal
// for the reference, absurd in practice - take the same segment scheme as above: a char that occurs in no keyword sends
// every one of these automata back to its root (LongestMatchSet.java:227 flushes the queue there, a Shortest walk ends, a
// WholeWordLongest walk fails), so the position after it starts an independent chain; one thread per such position follows
// the chain (SURVEY A.2 / A.3 / DESIGN "WholeWordLongest as a chain") with its own trie walks, to the next such position.

// longest (kFirst = false) or first = shortest (kFirst = true) keyword that starts at s: its length, 0 if none
template <bool kFirst>
__device__ __forceinline__ int64_t seg_walk(const DevAutomaton &A, const WwLitArgs &P, int64_t s, uint32_t &node_out) {
    uint32_t node = 0, info = 0;
    int64_t best = 0;
    for (int64_t i = s; i < P.n;) {
        const uint32_t c = __ldg(&A.cls[__ldg(&P.hay[i])]);
        if ((A.has_other && c == 0u) || !trie_step_sig(A, node, c, info)) break;
        ++i;
        if (info & kTerm) {
            best = i - s;
            node_out = node;
            if (kFirst) break;
        }
        if (!(info & kKids)) break;
    }
    return best;
}

struct SegReport {
    const DevAutomaton &A;
    const WwLitArgs &P;
    unsigned long long first;
    uint32_t count;
    template <bool kWrite, bool kIsMap>
    __device__ __forceinline__ void put(int64_t from, int64_t to, uint32_t node) {
        if (kWrite) {
            const unsigned long long at = first + count;
            if (at < (unsigned long long)P.cap) {
                P.pos_out[at] = make_int2((int32_t)from + P.pos_base, (int32_t)to + P.pos_base);
                if (kIsMap) P.val_out[at] = __ldg(&A.node_value[node]);
            }
        }
        ++count;
    }
};

// kFamily: 0 = the literal WholeWord loop above, ACGPU-style 1 Longest, 2 Shortest, 4 WholeWordLongest
template <int kFamily>
__device__ __forceinline__ bool seg_start(const WwLitView &V, int64_t t) {
    if (kFamily == 0) return V.sync_start(t);
    if (t == 0) return true;
    if (!V.A.has_other || V.cls(t - 1) != 0u) return false;
    if (kFamily == 4) return !V.word_scroll(t - 1) && !V.word_fold(t - 1) && V.word_scroll(t);
    return true;
}

template <int kFamily, bool kWrite, bool kIsMap>
__device__ __forceinline__ uint32_t seg_run(const WwLitView &V, int64_t t, unsigned long long first) {
    if (kFamily == 0) return wwlit_segment<kWrite, kIsMap>(V, t, first);
    const DevAutomaton &A = V.A;
    const WwLitArgs &P = V.P;
    SegReport R{A, P, first, 0u};
    if (kFamily == 1) {
        // Longest (A.2): from chain position p the first start with a keyword wins with its longest keyword; the chain goes on behind it
        int64_t s = t;
        while (s < P.n) {
            if (A.has_other && V.cls(s) == 0u) return R.count;  // the next segment starts behind this char
            uint32_t node = 0;
            const int64_t L = seg_walk<false>(A, P, s, node);
            if (L) {
                R.put<kWrite, kIsMap>(s, s + L, node);
                s += L;
            } else {
                ++s;
            }
        }
        return R.count;
    }
    if (kFamily == 2) {
        // Shortest (A.3): among the starts >= p the keyword that ENDS first wins (leftmost start on ties); the chain goes on at its end
        int64_t p = t;
        while (p < P.n) {
            int64_t best_end = P.n + 1, best_start = 0;
            uint32_t best_node = 0;
            for (int64_t s = p; s < best_end && s < P.n; ++s) {
                if (A.has_other && V.cls(s) == 0u) {
                    if (best_end > P.n) return R.count;  // nothing pending: the next segment starts behind this char
                    break;
                }
                uint32_t node = 0;
                const int64_t L = seg_walk<true>(A, P, s, node);
                if (L && s + L < best_end) {
                    best_end = s + L;
                    best_start = s;
                    best_node = node;
                }
            }
            if (best_end > P.n) return R.count;
            R.put<kWrite, kIsMap>(best_start, best_end, best_node);
            p = best_end;
        }
        return R.count;
    }
    // WholeWordLongest (WholeWordLongestMatchSet.java:47-182): walk from a walk start until there is no transition (idx);
    // report the longest keyword on the path that is followed by a non-word char or the end; scroll to the next word start.
    // "Followed by a non-word char" is decided on the LOWER-CASED char (the node's own match: :62-66 tests the folded char
    // at idx; a fail match: :224-240 tests the trie key, which is folded), the scrolls test the raw char in the String
    // overloads and the folded one in the Readable overload (Map :401-414) - the two views differ only for case-insensitive
    // matchers whose word-char table is not closed under toLowerCase (quirk Q7).
    auto word = [&](int64_t i) { return V.word_scroll(i); };
    int64_t idx = t;
    while (idx < P.n) {
        const int64_t s = idx;
        uint32_t node = 0, info = 0, rep_node = 0;
        int64_t len = 0;
        while (idx < P.n) {
            const uint32_t c = V.cls(idx);
            if ((A.has_other && c == 0u) || !trie_step_sig(A, node, c, info)) break;
            ++idx;
            if ((info & kTerm) && (idx == P.n || !V.word_fold(idx))) {
                len = idx - s;
                rep_node = node;
            }
            if (!(info & kKids)) break;
        }
        if (len) R.put<kWrite, kIsMap>(s, s + len, rep_node);
        if (idx >= P.n) break;
        // idx: the first position without a transition (a leaf has none for any char)
        if (V.word_fold(idx)) {
            while (++idx < P.n && word(idx)) {
            }
        }
        while (++idx < P.n && !word(idx)) {
        }
        if (idx < P.n && seg_start<4>(V, idx)) return R.count;
    }
    return R.count;
}

// kWrite = false: row_count[row] = records of the segments starting in the row.  kWrite = true: the records.
template <int kFamily, bool kWrite, bool kIsMap>
__global__ void __launch_bounds__(kMaskRow) k_segments(const DevAutomaton A, const WwLitArgs P) {
    __shared__ uint32_t s_warp[kMaskRow / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WwLitView V{A, P};
    for (int64_t row = blockIdx.x; row < P.n_rows; row += gridDim.x) {
        const int64_t t = row * kMaskRow + tid;
        const bool start = t < P.n && seg_start<kFamily>(V, t);
        const uint32_t cnt = start ? seg_run<kFamily, false, kIsMap>(V, t, 0ull) : 0u;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        __syncthreads();  // the previous row's readers are done with s_warp
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kMaskRow / 32; w++) {
            const uint32_t x = s_warp[w];
            if (w < warp) before += x;
            total += x;
        }
        if (!kWrite) {
            if (tid == 0) P.row_count[row] = total;
        } else if (cnt) {
            const unsigned long long first = __ldg(P.block_excl + (row >> 12)) + __ldg(P.row_excl + row) + before + (inc - cnt);
            seg_run<kFamily, true, kIsMap>(V, t, first);
        }
    }
    static_assert(kScanRows == 4096, "row >> 12 is the scan block of a row");
}

}  // namespace acgpu
