// WholeWord (generation 2): one launch, no trie.
//
// WholeWordMatchSet.java:47-132 walks a trie from every word start and reports the word when the walk ends on a keyword
// exactly at the word's end.  With a word-character table closed under toLowerCase (checked by the builder)
// that is: split the haystack into maximal runs of word characters and report a run iff its folded text is a keyword -
// a set-membership test.  k_ww_scan does it per tile of 4 096 positions:
//   A  16 chars per thread (two 128-bit loads) -> classes in shared memory plus two bitmaps (word char, keyword char),
//      with max_len + 1 positions of right context (a run that starts in the tile may end beyond it);
//   B  word starts (word char preceded by a non-word char: three bit operations per 16 positions) compacted IN ORDER
//      into a queue;
//   C  one thread per queued start: run length and "only keyword chars" from the bitmaps, then hash the run's classes,
//      probe the keyword table, compare the stored class string exactly;
//   D  ordered emission: block scan of the hits + decoupled look-back across tiles (tiles are taken in ticket order).
#pragma once
#include "builder.hpp"
#include "kernels.cuh"

namespace acgpu {

constexpr int kWwThreads = 256;
constexpr int kWwPer = 16;                       // positions per thread in phase B
constexpr int kWwTile = kWwThreads * kWwPer;     // 4 096 positions per tile
constexpr int kWwQueue = kWwTile / 2;            // word starts are at least two positions apart

struct DevWw {
    const uint16_t *wcls;     // [65536]
    const uint4 *buckets;     // 2 entries per bucket
    const uint16_t *pool;
    uint32_t n_buckets;
    int32_t max_len;
    const uint32_t *bloom;    // generation 3 (kernel_ww3.cuh): one-hash Bloom filter over the keys, bloom_bits = 0: none
    uint32_t bloom_bits;
};

struct WwArgs {
    const uint16_t *hay;
    int64_t n;              // chars in the window; the end of the window is the end of the input for the runs
    int64_t dom_lo;         // words starting in [dom_lo, dom_hi) are reported
    int64_t dom_hi;
    int64_t origin;         // first position of tile 0: <= dom_lo and hay + origin is 16-byte aligned
    int64_t n_tiles;        // tiles cover [origin, dom_hi)
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
    unsigned long long *total_out;
    unsigned int *tile_counter;
    unsigned long long *status;
};

// 16 consecutive positions starting at window position p0 (hay + p0 is 16-byte aligned when the group lies inside the
// window): classes to s_c[0..16), returns word-char bits | keyword-char bits << 16
__device__ __forceinline__ uint32_t ww_classify16(const DevWw &W, const uint16_t *hay, int64_t n, int64_t p0, const uint16_t *s_tab,
                                                  uint16_t *s_c) {
    uint32_t ch[16];
    if (p0 >= 0 && p0 + 16 <= n) {
        const uint4 v0 = __ldcs(reinterpret_cast<const uint4 *>(hay + p0)), v1 = __ldcs(reinterpret_cast<const uint4 *>(hay + p0) + 1);
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) { ch[2 * k] = w[k] & 0xFFFFu; ch[2 * k + 1] = w[k] >> 16; }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) ch[j] = (p0 + j >= 0 && p0 + j < n) ? (uint32_t)__ldg(&hay[p0 + j]) : 0x10000u;  // outside: no word char
    }
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) any |= ch[j];
    uint32_t wc = 0, kc = 0;
    if (any < 256u) {  // the common case: 16 Latin-1 chars, classes straight from the shared-memory table
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t x0 = s_tab[ch[2 * k]], x1 = s_tab[ch[2 * k + 1]];
            const uint32_t pair = x0 | (x1 << 16);
            reinterpret_cast<uint32_t *>(s_c)[k] = pair & 0x7FFF7FFFu;
            wc |= ((pair >> 15) & 1u) << (2 * k) | (pair >> 31) << (2 * k + 1);
            kc |= ((pair & 0x7FFFu) ? 1u : 0u) << (2 * k) | ((pair & 0x7FFF0000u) ? 1u : 0u) << (2 * k + 1);
        }
        return wc | (kc << 16);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t x[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t c = ch[2 * k + h];
            x[h] = c < 256u ? (uint32_t)s_tab[c] : (c < 0x10000u ? (uint32_t)__ldg(&W.wcls[c]) : 0u);
            wc |= (x[h] >> 15) << (2 * k + h);
            kc |= ((x[h] & 0x7FFFu) ? 1u : 0u) << (2 * k + h);
        }
        reinterpret_cast<uint32_t *>(s_c)[k] = (x[0] & 0x7FFFu) | ((x[1] & 0x7FFFu) << 16);
    }
    return wc | (kc << 16);
}

// 32 bits of a shared-memory bitmap starting at bit `at`
__device__ __forceinline__ uint32_t ww_bits32(const uint32_t *bm, uint32_t at) {
    return __funnelshift_r(bm[at >> 5], bm[(at >> 5) + 1], at & 31u);
}

template <bool kIsMap>
__global__ void __launch_bounds__(kWwThreads) k_ww_scan(const DevWw W, const WwArgs P) {
    extern __shared__ __align__(16) uint16_t s_ww_dyn[];
    // s_c[i] = class of position t0 + i, i < kWwTile + 16 * n_halo (n_halo groups of 16 cover max_len + 1 positions)
    uint16_t *s_c = s_ww_dyn;
    __shared__ uint16_t s_tab[256];
    __shared__ uint32_t s_wc[kWwTile / 32 + 18], s_kc[kWwTile / 32 + 18];  // word-char / keyword-char bits of the same positions
    __shared__ uint32_t s_prev;                                           // word-char bit of position t0 - 1
    __shared__ uint16_t s_q[kWwQueue];
    __shared__ uint16_t s_len[kWwQueue];
    __shared__ uint32_t s_val[kIsMap ? kWwQueue : 1];
    __shared__ uint32_t s_tmp[kWarps + 1];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_halo = (W.max_len + 1 + 15) / 16;  // <= 16
    for (int i = tid; i < 256; i += kWwThreads) s_tab[i] = __ldg(&W.wcls[i]);
    if (tid < 10) { s_wc[kWwTile / 32 + 8 + tid] = 0u; s_kc[kWwTile / 32 + 8 + tid] = 0u; }  // read by ww_bits32 past the last halo group

    while (true) {
        __syncthreads();  // previous tile done (and s_tab written)
        if (tid == 0) s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t t0 = P.origin + tile * kWwTile;

        // ---- A: classes and bitmaps of [t0, t0 + kWwTile + 16 * n_halo), word-char bit of t0 - 1
        {
            uint32_t bits = ww_classify16(W, P.hay, P.n, t0 + tid * 16, s_tab, s_c + tid * 16);
            uint32_t other = __shfl_down_sync(0xFFFFFFFFu, bits, 1);
            if (!(lane & 1)) {
                s_wc[tid >> 1] = (bits & 0xFFFFu) | (other << 16);
                s_kc[tid >> 1] = (bits >> 16) | (other & 0xFFFF0000u);
            }
            bits = 0;
            if (tid < n_halo) bits = ww_classify16(W, P.hay, P.n, t0 + kWwTile + tid * 16, s_tab, s_c + kWwTile + tid * 16);
            if (warp == 0) {  // n_halo <= 16: the first warp holds every halo group
                other = __shfl_down_sync(0xFFFFFFFFu, bits, 1);
                if (!(lane & 1) && lane < 16) {
                    s_wc[kWwTile / 32 + (lane >> 1)] = (bits & 0xFFFFu) | (other << 16);
                    s_kc[kWwTile / 32 + (lane >> 1)] = (bits >> 16) | (other & 0xFFFF0000u);
                }
            }
            if (tid == kWwThreads - 1) {
                const int64_t p = t0 - 1;
                uint32_t w = 0;
                if (p >= 0 && p < P.n) {
                    const uint32_t ch = __ldg(&P.hay[p]);
                    w = ch < 256u ? (uint32_t)s_tab[ch] : (uint32_t)__ldg(&W.wcls[ch]);
                }
                s_prev = w >> 15;
            }
        }
        __syncthreads();

        // ---- B: word starts of the thread's 16 positions, compacted in order
        uint32_t starts;
        {
            const uint32_t word = s_wc[tid >> 1];
            const uint32_t mine = (tid & 1) ? word >> 16 : word & 0xFFFFu;
            const uint32_t prev = (tid & 1) ? (word >> 15) & 1u : (tid ? s_wc[(tid >> 1) - 1] >> 31 : s_prev);
            starts = mine & ~((mine << 1) | prev) & 0xFFFFu;
            // only words starting in [dom_lo, dom_hi) report (edge tiles)
            const int64_t p0 = t0 + tid * 16;
            if (p0 < P.dom_lo) starts &= P.dom_lo - p0 >= 16 ? 0u : (0xFFFFu << (int)(P.dom_lo - p0));
            if (p0 + 16 > P.dom_hi) starts &= P.dom_hi <= p0 ? 0u : (0xFFFFu >> (int)(p0 + 16 - P.dom_hi));
        }
        uint32_t nq;
        uint32_t qoff = block_exclusive_sum(__popc(starts), s_tmp, nq);
        while (starts) {
            const int j = __ffs(starts) - 1;
            starts &= starts - 1u;
            s_q[qoff++] = (uint16_t)(tid * 16 + j);
        }
        __syncthreads();

        // ---- C: one queued word per thread
        for (uint32_t q = tid; q < nq; q += kWwThreads) {
            const uint32_t p = s_q[q];
            // length of the run and whether all of it is keyword chars, 32 positions per step
            uint32_t L = 0;
            bool ok = true;
            for (uint32_t at = p;; at += 32) {
                const uint32_t x = ww_bits32(s_wc, at), k = ww_bits32(s_kc, at);
                const uint32_t r = x == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~x) - 1u;
                const uint32_t m = r == 32u ? 0xFFFFFFFFu : (1u << r) - 1u;
                ok = ok && (k & m) == m;
                L += r;
                if (r < 32u || L > (uint32_t)W.max_len) break;
            }
            uint32_t hit_len = 0, hit_val = kNoneD;
            if (ok && L <= (uint32_t)W.max_len) {
                const uint16_t *run = s_c + p;
                WwHash h;
                {
                    // two classes per step: aligned words of the class array, shifted into place when the word starts odd
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(s_c) + (p >> 1);
                    const uint32_t sh = (p & 1u) * 16u;
                    uint32_t w0 = cw[0];
                    for (uint32_t i = 0; i < L; i += 2) {
                        const uint32_t w1 = cw[(i >> 1) + 1];
                        uint32_t pair = __funnelshift_r(w0, w1, sh);
                        if (i + 1 >= L) pair &= 0xFFFFu;  // a lone last class
                        h.add_pair(pair);
                        w0 = w1;
                    }
                }
                h.finish(L);
                uint32_t bk = __umulhi(h.spread(), W.n_buckets);
                for (uint32_t tries = 0; tries < W.n_buckets && !hit_len; tries++) {
                    const uint4 e0 = __ldg(W.buckets + (size_t)bk * 2), e1 = __ldg(W.buckets + (size_t)bk * 2 + 1);
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const uint4 e = k ? e1 : e0;
                        if (e.x != h.h1 || e.y != L || e.z == 0xFFFFFFFFu || hit_len) continue;
                        const uint16_t *kw = W.pool + e.z;
                        bool same = true;
                        for (uint32_t i = 0; i < L && same; i++) same = __ldg(kw + i) == run[i];
                        if (same) { hit_len = L; hit_val = e.w; }
                    }
                    if (e0.z == 0xFFFFFFFFu || e1.z == 0xFFFFFFFFu) break;  // a free slot on the probe path: not in the table
                    bk = bk + 1u == W.n_buckets ? 0u : bk + 1u;
                }
            }
            s_len[q] = (uint16_t)hit_len;
            if (kIsMap) s_val[q] = hit_val;
        }
        __syncthreads();

        // ---- D: ordered emission
        const uint32_t per = (nq + kWwThreads - 1) / kWwThreads;
        const uint32_t q_lo = min(nq, (uint32_t)tid * per), q_hi = min(nq, q_lo + per);
        uint32_t mine = 0;
        for (uint32_t q = q_lo; q < q_hi; q++) mine += s_len[q] != 0;
        uint32_t block_total;
        const uint32_t my_off = block_exclusive_sum(mine, s_tmp, block_total);
        if (warp == 0) {
            const unsigned long long excl = lookback_exclusive(P.status, tile, block_total);
            if (lane == 0) {
                s_base = excl;
                if (tile == P.n_tiles - 1) *P.total_out = excl + block_total;
            }
        }
        __syncthreads();
        unsigned long long idx = s_base + my_off;
        for (uint32_t q = q_lo; q < q_hi; q++) {
            const uint32_t L = s_len[q];
            if (!L) continue;
            if (idx < (unsigned long long)P.cap) {
                const int32_t st = (int32_t)(t0 + s_q[q]) + P.pos_base;
                P.pos_out[idx] = make_int2(st, st + (int32_t)L);
                if (kIsMap) P.val_out[idx] = s_val[q];
            }
            ++idx;
        }
    }
}

// ---------------------------------------------------------------- WholeWordLongest with phrases: compacted walk starts
//
// WholeWordLongestMatchSet.java:47-182 walks the trie from the first char of the input and from every word start, over
// word AND non-word chars, until there is no transition (position idx), reports the longest keyword on the path that
// is followed by a non-word char or the end, and goes on at the first word start after idx ("WholeWordLongest
// as a chain").  Generation 1 ran that chain over EVERY haystack position (k_fwd_v<4> + k_sel_*: 34.6 ms per 10^9 chars
// although only one position in six is a walk start).  k_wwl_starts finds the walk starts with k_ww_scan's bitmaps, walks
// from each, and writes them COMPACTED and in order:
//     wpos[i] = haystack position of walk start i
//     v[i]    = reported length | (number of walk starts the chain moves on by) << 8     (1 + starts inside (wpos[i], idx])
// so the selection kernels resolve the chain over n / 6 candidates in index space (SelArgs::wpos maps back).
struct WwlArgs {
    const uint16_t *hay;
    int64_t n;
    int64_t origin;         // first position of tile 0: <= 0 and hay + origin is 16-byte aligned
    int64_t n_tiles;        // tiles cover [origin, n)
    int32_t *wpos;          // [cap]
    uint16_t *v;            // [cap]
    int64_t cap;
    unsigned long long *total_out;   // number of walk starts
    unsigned int *tile_counter;
    unsigned long long *status;
};

__global__ void __launch_bounds__(kWwThreads) k_wwl_starts(const DevAutomaton A, const DevWw W, const WwlArgs P) {
    extern __shared__ __align__(16) uint16_t s_ww_dyn[];
    uint16_t *s_c = s_ww_dyn;  // classes of [t0, t0 + kWwTile + 16 * n_halo)
    __shared__ uint16_t s_tab[256];
    __shared__ uint32_t s_wc[kWwTile / 32 + 18];  // word-char bits of the window
    __shared__ uint32_t s_prev;
    __shared__ uint16_t s_q[kWwQueue + 1];
    __shared__ uint16_t s_v[kWwQueue + 1];
    __shared__ uint32_t s_tmp[kWarps + 1];
    __shared__ long long s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_halo = (W.max_len + 1 + 15) / 16;  // <= 16
    for (int i = tid; i < 256; i += kWwThreads) s_tab[i] = __ldg(&W.wcls[i]);
    if (tid < 10) s_wc[kWwTile / 32 + 8 + tid] = 0u;

    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t t0 = P.origin + tile * kWwTile;
        // ---- A: classes and word-char bits of the tile and its halo (k_ww_scan's phase A)
        {
            uint32_t bits = ww_classify16(W, P.hay, P.n, t0 + tid * 16, s_tab, s_c + tid * 16);
            uint32_t other = __shfl_down_sync(0xFFFFFFFFu, bits, 1);
            if (!(lane & 1)) s_wc[tid >> 1] = (bits & 0xFFFFu) | (other << 16);
            bits = 0;
            if (tid < n_halo) bits = ww_classify16(W, P.hay, P.n, t0 + kWwTile + tid * 16, s_tab, s_c + kWwTile + tid * 16);
            if (warp == 0) {
                other = __shfl_down_sync(0xFFFFFFFFu, bits, 1);
                if (!(lane & 1) && lane < 16) s_wc[kWwTile / 32 + (lane >> 1)] = (bits & 0xFFFFu) | (other << 16);
            }
            if (tid == kWwThreads - 1) {
                const int64_t p = t0 - 1;
                uint32_t w = 0;
                if (p >= 0 && p < P.n) {
                    const uint32_t ch = __ldg(&P.hay[p]);
                    w = ch < 256u ? (uint32_t)s_tab[ch] : (uint32_t)__ldg(&W.wcls[ch]);
                }
                s_prev = w >> 15;
            }
        }
        __syncthreads();
        // ---- B: walk starts of the thread's 16 positions (word starts, and the first char of the input), in order
        uint32_t starts;
        {
            const uint32_t word = s_wc[tid >> 1];
            const uint32_t mine = (tid & 1) ? word >> 16 : word & 0xFFFFu;
            const uint32_t prev = (tid & 1) ? (word >> 15) & 1u : (tid ? s_wc[(tid >> 1) - 1] >> 31 : s_prev);
            starts = mine & ~((mine << 1) | prev) & 0xFFFFu;
            const int64_t p0 = t0 + tid * 16;
            if (p0 <= 0 && p0 + 16 > 0) starts |= 1u << (int)(-p0);                       // position 0 always starts a walk
            if (p0 < 0) starts &= -p0 >= 16 ? 0u : (0xFFFFu << (int)(-p0));
            if (p0 + 16 > P.n) starts &= P.n <= p0 ? 0u : (0xFFFFu >> (int)(p0 + 16 - P.n));
        }
        uint32_t nq;
        uint32_t qoff = block_exclusive_sum(__popc(starts), s_tmp, nq);
        while (starts) {
            const int j = __ffs(starts) - 1;
            starts &= starts - 1u;
            s_q[qoff++] = (uint16_t)(tid * 16 + j);
        }
        __syncthreads();
        // ---- C: one walk per thread
        const int win = kWwTile + 16 * n_halo;  // positions of the window
        for (uint32_t q = tid; q < nq; q += kWwThreads) {
            const int p = s_q[q];
            uint32_t node = 0, info = 0, len = 0;
            int i = p;
            while (t0 + i < P.n) {
                const uint32_t c = i < win ? (uint32_t)s_c[i] : (uint32_t)__ldg(&A.cls[__ldg(&P.hay[t0 + i])]);
                if ((A.has_other && c == 0u) || !trie_step_sig(A, node, c, info)) break;
                ++i;
                if (info & kTerm) {
                    const bool word_next = t0 + i < P.n && (i < win ? ((s_wc[i >> 5] >> (i & 31)) & 1u) != 0u
                                                                     : (__ldg(&W.wcls[__ldg(&P.hay[t0 + i])]) >> 15) != 0u);
                    if (!word_next) len = (uint32_t)(i - p);
                }
                if (!(info & kKids)) break;
            }
            // walk starts inside (p, i]: word-start bits of those positions, 32 at a time (i - p <= max_len: inside the window)
            uint32_t skipped = 0;
            for (int at = p + 1; at <= i; at += 32) {
                const uint32_t x = ww_bits32(s_wc, (uint32_t)at), xp = ww_bits32(s_wc, (uint32_t)(at - 1));
                uint32_t st = x & ~xp;
                const int left = i - at + 1;  // positions at .. i
                if (left < 32) st &= (1u << left) - 1u;
                skipped += (uint32_t)__popc(st);
            }
            s_v[q] = (uint16_t)(len | (1u + skipped) << 8);
        }
        __syncthreads();
        // ---- D: the tile's walk starts go out in order
        if (warp == 0) {
            const unsigned long long excl = lookback_exclusive(P.status, tile, nq);
            if (lane == 0) {
                s_base = excl;
                if (tile == P.n_tiles - 1) *P.total_out = excl + nq;
            }
        }
        __syncthreads();
        const unsigned long long base = s_base;
        for (uint32_t q = tid; q < nq; q += kWwThreads) {
            if (base + q < (unsigned long long)P.cap) {
                P.wpos[base + q] = (int32_t)(t0 + s_q[q]);
                P.v[base + q] = s_v[q];
            }
        }
    }
}

}  // namespace acgpu
