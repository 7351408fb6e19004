// WholeWord (generation 3): warp-autonomous, no block barrier, no per-word hashing loop.
//
// Same reformulation as generation 2 (kernel_ww.cuh; WholeWordMatchSet.java:47-132 with a word-character table closed
// under toLowerCase): split the haystack into maximal runs of word chars, report a run iff its folded text is a keyword.
// Generation 2 spent 87 thread-instructions per char in four block-wide phases (ncu: 9 barrier stalls per issue, a
// divergent hashing loop per word).  Here the hash of EVERY run comes out of one running polynomial over the haystack:
//     v[i] = class + 1 of a word char, 0 of any other char        G[i] = G[i-1] * B + v[i]   (mod 2^32)
//     hash of the run [s, t) = G[t-1] - G[s-1] * B^(t-s)
// A warp owns a chunk of 32 rows of 256 positions (tickets); a lane owns 8 consecutive chars of a row (one streaming
// 128-bit load, prefetched a row ahead).  Per row: 8 table look-ups and 8 multiply-adds per lane, a 5-step warp scan that
// composes the lanes' polynomials, G and the word-char bits go to a two-row ring in shared memory (a run is at most 255
// chars).  Run ENDS (a non-word char after a word char) are compacted in order; one lane per end finds the run's start with
// count-leading-zeros over the ring's bitmap, takes the hash from two ring loads, tests a one-hash Bloom filter in shared
// memory (50 k keywords: 12 % pass) and queues the survivors; the queue is probed against the keyword table 32 at a time
// (one gathered 32-byte bucket; the stored class string is compared exactly on a tag match).
//
//   k_ww3_hits   two bitmaps on zeroed scratch (1 bit per position each): a keyword run STARTS here / ENDS before here; Maps
//                also leave the run's value at slot end / 2 of a scratch array (run ends are at least 2 positions apart)
//   k_ww3_count  records per row = set bits of the row's end bitmap
//   k_row_scan   row counts -> offsets (kernel_emit.cuh)
//   k_ww3_emit   end bits -> (start, end[, value]) records in position order; the start is the last start bit before the end
// (The first version counted rows with one atomicAdd per hit, probed the table a second time to verify, and let the emit
// kernel scan back over the word's chars and hash it again for a Map's value: on text where EVERY word is a keyword that
// was 10.4 / 15.9 ms per 10^9 chars against generation 2's 6.3 / 6.7 - tools/bench_ww_dense.py.)
#pragma once
#include "builder.hpp"
#include "kernel_emit.cuh"
#include "kernel_ww.cuh"

namespace acgpu {

constexpr int kW3Warps = 32;                 // one CTA per SM
constexpr int kW3Row = kMaskRow;             // 256 positions: k_row_scan's rows
#ifndef ACGPU_W3_CHUNK
#define ACGPU_W3_CHUNK 64     // rows per ticket (32: +1 %: one context row per chunk)
#endif
#ifndef ACGPU_W3_SWZ
#define ACGPU_W3_SWZ 0      // experiment, lost 2 %: spread the class table over the banks (upper / lower case and digits share banks in ASCII order)
#endif
#ifndef ACGPU_W3_STS
#define ACGPU_W3_STS 1      // conflict-free order of the two 128-bit stores of G (+2.3 %)
#endif
constexpr int kW3ChunkRows = ACGPU_W3_CHUNK;
// shared memory per warp, in words.  kShort (keywords < 32 chars: a run and the char before it are at most 33 positions back):
// G of the row + the previous row's last 64, word-char bits likewise, no index wraps.  Otherwise a two-row ring.  Then the end
// queue (128 x u16) and the probe and verify queues (64 x 8 bytes each).  (With 64 KB of Bloom filter the kShort layout keeps
// the CTA under the 164 KB carve-out; 182 KB - the two-row ring for every dictionary - left 28 KB of L1 and cost 5 %.)
__host__ __device__ constexpr int w3_g_words(bool shortk) { return shortk ? 64 + 256 : 64 + 512; }
__host__ __device__ constexpr int w3_wb_words(bool shortk) { return shortk ? 12 : 20; }
__host__ __device__ constexpr int w3_warp_words(bool shortk) { return w3_g_words(shortk) + w3_wb_words(shortk) + 64 + 128 + 128; }
static_assert(kW3Row == 256 && kWwMaxLen <= 255, "a run and the char before it fit the two-row ring; lengths fit 8 bits");

__host__ __device__ constexpr uint32_t w3_pow(uint32_t b, int e) {
    uint32_t r = 1;
    for (int i = 0; i < e; i++) r *= b;
    return r;
}
__host__ __device__ constexpr size_t ww3_smem_bytes(uint32_t bloom_bits, bool shortk) {
    return (size_t)(256 + 256 + bloom_bits / 32 + kW3Warps * w3_warp_words(shortk)) * 4;
}

struct Ww3Args {
    const uint16_t *hay;
    int64_t n;              // chars in the window; its end is the end of the input for the runs
    int64_t dom_lo;         // words starting in [dom_lo, dom_hi) are reported
    int64_t dom_hi;
    int64_t origin;         // first position of row 0: <= dom_lo and hay + origin is 16-byte aligned
    int64_t n_rows;         // rows cover [origin, min(n, dom_hi + max_len)] - a run that ends with the window is reported at n
    uint32_t *hitbits;      // [n_rows * 8], zeroed: bit t - origin = a keyword run ends at t (exclusive)
    uint32_t *startbits;    // [n_rows * 8], zeroed: bit s - origin = a keyword run starts at s
    uint32_t *val_scratch;  // Maps: [n_rows * 128] value of the run that ends at t, at slot (t - origin) / 2
    unsigned int *ticket;
    int32_t chunk_rows;     // rows per ticket, 1 .. kW3ChunkRows: short windows take small tickets so that every warp gets one
};

struct Ww3EmitArgs {
    const uint16_t *hay;
    int64_t n;
    int64_t origin;
    int64_t n_rows;
    const uint32_t *hitbits;
    const uint32_t *startbits;
    const uint32_t *val_scratch;
    const uint32_t *row_excl;
    const unsigned long long *block_excl;
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
};

// x of a wcls entry (class | word-char flag << 15): a word char gives (class + 1) | 1 << 31, any other char 0 - the top bit is
// the word-char bit (one funnel shift per char collects the bits) and the whole word is the polynomial's digit
__device__ __forceinline__ uint32_t w3_enc(uint32_t x) { return (x >> 15) ? ((x & 0x7FFFu) + 1u) | 0x80000000u : 0u; }
// slot of a Latin-1 code unit in the shared-memory table
__device__ __forceinline__ uint32_t w3_slot(uint32_t ch) { return ACGPU_W3_SWZ ? ch ^ (ch >> 5) : ch; }
__device__ __forceinline__ uint32_t w3_v(const DevWw &W, const uint32_t *s_tab, uint32_t ch) {
    return ch < 256u ? s_tab[w3_slot(ch)] : w3_enc(__ldg(&W.wcls[ch]));
}
__device__ __forceinline__ uint32_t w3_bucket(const DevWw &W, uint32_t key) { return __umulhi(ww_poly_spread(key), W.n_buckets); }

// The slot (bucket * 2 + entry) of an entry with this key and length on the key's probe path, 0xFFFFFFFF if there is none.
// (No string comparison: the survivors are verified 32 at a time.)
__device__ __forceinline__ uint32_t ww3_tag_probe(const DevWw &W, uint32_t key, uint32_t len) {
    uint32_t bk = w3_bucket(W, key);
    for (uint32_t tries = 0; tries < W.n_buckets; tries++) {
        const uint4 e0 = __ldg(W.buckets + (size_t)bk * 2), e1 = __ldg(W.buckets + (size_t)bk * 2 + 1);
        if (e0.x == key && e0.y == len && e0.z != 0xFFFFFFFFu) return bk * 2u;
        if (e1.x == key && e1.y == len && e1.z != 0xFFFFFFFFu) return bk * 2u + 1u;
        if (e0.z == 0xFFFFFFFFu || e1.z == 0xFFFFFFFFu) return 0xFFFFFFFFu;  // a free slot on the probe path: not in the table
        bk = bk + 1u == W.n_buckets ? 0u : bk + 1u;
    }
    return 0xFFFFFFFFu;
}

// exact comparison of the run hay[s, s + len) with the class string of a table entry, four chars per step (their eight loads
// are independent: on text full of keywords the one-char-per-step loop was a chain of L2 round trips)
__device__ __forceinline__ bool ww3_same(const DevWw &W, const uint16_t *hay, const uint32_t *s_tab, uint32_t pool_off, uint32_t len, int64_t s) {
    const uint16_t *kw = W.pool + pool_off, *h = hay + s;
    bool same = true;
    for (uint32_t i = 0; i < len && same; i += 4) {
        uint32_t k[4], c[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool on = i + j < len;
            k[j] = on ? (uint32_t)__ldg(kw + i + j) : 0u;
            c[j] = on ? (uint32_t)__ldg(h + i + j) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) same = same && (i + j >= len || (k[j] + 1u | 0x80000000u) == w3_v(W, s_tab, c[j]));
    }
    return same;
}

// Is the run hay[s, s + len) with key `key` a keyword?  One bucket per step of the probe path; exact comparison of the
// stored class string on a tag match.
__device__ __forceinline__ bool ww3_lookup(const DevWw &W, const uint16_t *hay, const uint32_t *s_tab, uint32_t key, uint32_t len, int64_t s,
                                           uint32_t &val) {
    uint32_t bk = w3_bucket(W, key);
    for (uint32_t tries = 0; tries < W.n_buckets; tries++) {
        const uint4 e0 = __ldg(W.buckets + (size_t)bk * 2), e1 = __ldg(W.buckets + (size_t)bk * 2 + 1);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint4 e = k ? e1 : e0;
            if (e.x != key || e.y != len || e.z == 0xFFFFFFFFu) continue;
            if (ww3_same(W, hay, s_tab, e.z, len, s)) {
                val = e.w;
                return true;
            }
        }
        if (e0.z == 0xFFFFFFFFu || e1.z == 0xFFFFFFFFu) return false;
        bk = bk + 1u == W.n_buckets ? 0u : bk + 1u;
    }
    return false;
}

// a queue of 8-byte entries in shared memory, filled by ballot compaction, drained 32 at a time by the whole warp
struct W3Queue {
    uint2 *q;
    uint32_t n;
    __device__ __forceinline__ void push(bool on, uint2 e, uint32_t lt_mask) {
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, on);
        if (on) q[n + __popc(ballot & lt_mask)] = e;
        n += (uint32_t)__popc(ballot);
        __syncwarp();
    }
    // after the first 32 entries were consumed
    __device__ __forceinline__ void pop32(int lane) {
        n -= 32u;
        uint2 e = make_uint2(0u, 0u);
        if ((uint32_t)lane < n) e = q[32 + lane];
        __syncwarp();
        if ((uint32_t)lane < n) q[lane] = e;
        __syncwarp();
    }
};

// kBloom: the dictionary has a Bloom filter; kShort: max_len < 32 (a run's length comes from one 32-bit window of the bitmap)
template <bool kBloom, bool kShort>
__global__ void __launch_bounds__(kW3Warps * 32, 1) k_ww3_hits(const DevWw W, const Ww3Args P) {
    extern __shared__ __align__(16) uint32_t s_w3[];
    uint32_t *s_tab = s_w3;                                 // [256] x of the Latin-1 code units
    uint32_t *s_pow = s_w3 + 256;                           // [256] B^len
    uint32_t *s_bloom = s_w3 + 512;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_mine = s_bloom + (W.bloom_bits >> 5) + warp * w3_warp_words(kShort);
    uint32_t *s_G = s_mine;                                  // running polynomial: [64 + 256] (kShort) or a ring of two rows at [64, 64 + 512)
    uint32_t *s_wb = s_mine + w3_g_words(kShort);            // word-char bits of the same positions
    uint16_t *s_q = reinterpret_cast<uint16_t *>(s_wb + w3_wb_words(kShort));   // [128] run ends of the row, row-relative
    W3Queue probe_q{reinterpret_cast<uint2 *>(s_wb + w3_wb_words(kShort) + 64), 0u};    // [64] {key, (start - chunk context base) << 8 | length}
    W3Queue verify_q{reinterpret_cast<uint2 *>(s_wb + w3_wb_words(kShort) + 192), 0u};  // [64] the same entries after a tag match
    constexpr uint32_t B = kWwPolyB;

    if (tid < 256) {
        s_tab[w3_slot(tid)] = w3_enc(__ldg(&W.wcls[tid]));
        uint32_t p = 1, b = B;
        for (int e = tid; e; e >>= 1, b *= b)
            if (e & 1) p *= b;
        s_pow[tid] = p;
    }
    if (kBloom)
        for (uint32_t i = tid; i < (W.bloom_bits >> 5); i += kW3Warps * 32) s_bloom[i] = __ldg(&W.bloom[i]);
    __syncthreads();

    uint32_t pow_lane = 1;   // B^(8 * lane)
    {
        uint32_t b = w3_pow(B, 8);
        for (int e = lane; e; e >>= 1, b *= b)
            if (e & 1) pow_lane *= b;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t max_len = (uint32_t)W.max_len;
    const uint32_t bloom_shift = 32u - (uint32_t)__ffs((int)W.bloom_bits) + 1u;   // key >> shift = the filter's bit (bloom_bits is a power of two)

    while (true) {
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(P.ticket, 1u);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        const int64_t row0 = (int64_t)ticket * P.chunk_rows;
        if (row0 >= P.n_rows) break;
        const int n_chunk_rows = (int)(min(row0 + P.chunk_rows, P.n_rows) - row0) + 1;   // with the context row
        const int64_t ctx_base = P.origin + (row0 - 1) * kW3Row;   // the chunk's context row: the 256 positions before it
        // positions are chunk-relative (0 = ctx_base) from here on
        const int32_t dom_lo = (int32_t)max((int64_t)-1, min(P.dom_lo - ctx_base, (int64_t)1 << 20));
        const int32_t dom_hi = (int32_t)max((int64_t)-1, min(P.dom_hi - ctx_base, (int64_t)1 << 20));
        const bool whole = ctx_base >= 0 && ctx_base + (int64_t)n_chunk_rows * kW3Row <= P.n;   // every load of the chunk lies inside the window
        uint32_t carry_row = 0, prev_bit = 0;

        auto verify = [&](uint32_t count) {   // the first `count` entries of the verify queue {packed, table slot}, one per lane
            if ((uint32_t)lane < count) {
                const uint2 q = verify_q.q[lane];
                const uint32_t len = q.x & 0xFFu;
                const int64_t s = ctx_base + (int64_t)(q.x >> 8);
                const uint4 e = __ldg(W.buckets + q.y);   // {key, length, pool offset, value}
                uint32_t val = e.w;
                // a tag match that is another string (a 32-bit collision) sends the search down the rest of the probe path
                if (ww3_same(W, P.hay, s_tab, e.z, len, s) || ww3_lookup(W, P.hay, s_tab, e.x, len, s, val)) {
                    const int64_t rs = (row0 - 1) * kW3Row + (int64_t)(q.x >> 8), rt = rs + len;   // relative to the origin
                    atomicOr(&P.startbits[rs >> 5], 1u << (uint32_t)(rs & 31));
                    atomicOr(&P.hitbits[rt >> 5], 1u << (uint32_t)(rt & 31));
                    if (P.val_scratch) P.val_scratch[rt >> 1] = val;
                }
            }
            __syncwarp();
        };
        auto probe = [&](uint32_t count) {   // the first `count` entries of the probe queue {key, packed}
            uint2 e = make_uint2(0u, 0u);
            uint32_t slot = 0xFFFFFFFFu;
            if ((uint32_t)lane < count) {
                e = probe_q.q[lane];
                slot = ww3_tag_probe(W, e.x, e.y & 0xFFu);
            }
            verify_q.push(slot != 0xFFFFFFFFu, make_uint2(e.y, slot), lt_mask);
            if (verify_q.n >= 32u) {
                verify(32u);
                verify_q.pop32(lane);
            }
        };

        const uint4 *src = reinterpret_cast<const uint4 *>(P.hay + ctx_base) + lane;   // dereferenced only inside the window
        bool in_next = whole || (ctx_base + lane * 8 >= 0 && ctx_base + lane * 8 + 8 <= P.n);
        uint4 nxt = ldcs_v4_if(src, in_next);
        for (int ri = 0; ri < n_chunk_rows; ri++) {
            const uint4 cur = nxt;
            const bool inside = in_next;
            if (ri + 1 < n_chunk_rows) {
                src += kW3Row / 8;
                if (!whole) {
                    const int64_t p_next = ctx_base + (int64_t)(ri + 1) * kW3Row + lane * 8;
                    in_next = p_next >= 0 && p_next + 8 <= P.n;
                }
                nxt = ldcs_v4_if(src, in_next);
            }
            uint32_t v[8];
            if (inside) {
                if (((cur.x | cur.y | cur.z | cur.w) & 0xFF00FF00u) == 0u) {
                    v[0] = s_tab[w3_slot(cur.x & 0xFFu)]; v[1] = s_tab[w3_slot(cur.x >> 16)];
                    v[2] = s_tab[w3_slot(cur.y & 0xFFu)]; v[3] = s_tab[w3_slot(cur.y >> 16)];
                    v[4] = s_tab[w3_slot(cur.z & 0xFFu)]; v[5] = s_tab[w3_slot(cur.z >> 16)];
                    v[6] = s_tab[w3_slot(cur.w & 0xFFu)]; v[7] = s_tab[w3_slot(cur.w >> 16)];
                } else {
                    const uint32_t ch[8] = {cur.x & 0xFFFFu, cur.x >> 16, cur.y & 0xFFFFu, cur.y >> 16,
                                            cur.z & 0xFFFFu, cur.z >> 16, cur.w & 0xFFFFu, cur.w >> 16};
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = w3_v(W, s_tab, ch[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int64_t p = ctx_base + (int64_t)ri * kW3Row + lane * 8 + j;
                    v[j] = (p >= 0 && p < P.n) ? w3_v(W, s_tab, __ldg(&P.hay[p])) : 0u;   // outside the window: no word char
                }
            }
            // the lane's polynomial, the warp's composition, G of the lane's 8 positions
            uint32_t g[8], wb = 0;
            {
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc = acc * B + v[j];
                    g[j] = acc;
                    wb = __funnelshift_l(v[j], wb, 1);   // (wb << 1) | word-char bit of char j
                }
                wb = __brev(wb) >> 24;                   // bit j = char j
                uint32_t x = acc, y;
                y = __shfl_up_sync(0xFFFFFFFFu, x, 1);  if (lane >= 1) x += y * w3_pow(B, 8);
                y = __shfl_up_sync(0xFFFFFFFFu, x, 2);  if (lane >= 2) x += y * w3_pow(B, 16);
                y = __shfl_up_sync(0xFFFFFFFFu, x, 4);  if (lane >= 4) x += y * w3_pow(B, 32);
                y = __shfl_up_sync(0xFFFFFFFFu, x, 8);  if (lane >= 8) x += y * w3_pow(B, 64);
                y = __shfl_up_sync(0xFFFFFFFFu, x, 16); if (lane >= 16) x += y * w3_pow(B, 128);
                y = __shfl_up_sync(0xFFFFFFFFu, x, 1);
                const uint32_t carry = (lane ? y : 0u) + carry_row * pow_lane;
#pragma unroll
                for (int j = 0; j < 8; j++) g[j] += carry * w3_pow(B, j + 1);
                carry_row = __shfl_sync(0xFFFFFFFFu, g[7], 31);
            }
            // kShort: the row at s_G[64 ..), bit 64 + i of s_wb; otherwise ring index i (= position mod 512) at s_G[64 + i]
            const uint32_t slot256 = kShort ? 0u : (uint32_t)(ri & 1) * 256u;
            {
                uint4 *dst = reinterpret_cast<uint4 *>(s_G + 64 + slot256 + lane * 8);
                if (ACGPU_W3_STS) {
                    // a quarter warp's 8 stores of 16 bytes at a stride of 32 bytes hit every bank twice; lanes 4-7 of each
                    // quarter store their other half first
                    const bool hi_first = (lane & 4) != 0;
                    const uint4 lo = make_uint4(g[0], g[1], g[2], g[3]), hi = make_uint4(g[4], g[5], g[6], g[7]);
                    dst[hi_first ? 1 : 0] = hi_first ? hi : lo;
                    dst[hi_first ? 0 : 1] = hi_first ? lo : hi;
                } else {
                    dst[0] = make_uint4(g[0], g[1], g[2], g[3]);
                    dst[1] = make_uint4(g[4], g[5], g[6], g[7]);
                }
                reinterpret_cast<uint8_t *>(s_wb)[8 + (slot256 >> 3) + lane] = (uint8_t)wb;
            }
            uint32_t before = __shfl_up_sync(0xFFFFFFFFu, wb >> 7, 1);
            if (lane == 0) before = prev_bit;
            prev_bit = __shfl_sync(0xFFFFFFFFu, wb >> 7, 31);
            auto keep_tail = [&]() {   // kShort: the next row finds this row's last 64 positions below its own
                if (kShort) {
                    __syncwarp();
                    if (lane >= 24) {
                        const uint4 a = *reinterpret_cast<const uint4 *>(s_G + 64 + lane * 8), b = *reinterpret_cast<const uint4 *>(s_G + 68 + lane * 8);
                        *reinterpret_cast<uint4 *>(s_G + (lane - 24) * 8) = a;
                        *reinterpret_cast<uint4 *>(s_G + (lane - 24) * 8 + 4) = b;
                        reinterpret_cast<uint8_t *>(s_wb)[lane - 24] = (uint8_t)wb;
                    }
                }
                __syncwarp();   // the next row overwrites the row (the other half of the ring) and the end queue
            };
            if (ri == 0) {   // the context row reports nothing
                keep_tail();
                continue;
            }
            // run ends (a non-word char after a word char; exclusive end), chunk-relative, compacted in order
            const uint32_t row_rel = (uint32_t)ri * kW3Row;
            uint32_t ends = ~wb & ((wb << 1) | before) & 0xFFu;
            const uint32_t inc = warp_inclusive_sum((uint32_t)__popc(ends));
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            uint32_t off = inc - (uint32_t)__popc(ends);
            while (ends) {
                const int j = __ffs(ends) - 1;
                ends &= ends - 1u;
                s_q[off++] = (uint16_t)(lane * 8 + j);
            }
            __syncwarp();
            // (Carrying the last partial batch over to the next row - an end may wait one row if its run lies in this row - was
            // measured: 15 % fewer candidate rounds, but the bookkeeping cost more instructions than the rounds saved.)
            const uint32_t now = total;
            for (uint32_t q0 = 0; q0 < now; q0 += 32) {
                bool pass = false;
                uint32_t key = 0, packed = 0;
                if (q0 + lane < now) {
                    const uint32_t tq = s_q[q0 + lane], t = row_rel + tq, rt = slot256 + tq;   // chunk-relative end of the run, its (ring) index
                    // word-char bits of the 32 positions before the end, most recent highest: the run's length
                    uint32_t len;
                    if (kShort) {
                        const uint32_t b0 = rt + 32u;
                        len = (uint32_t)__clz((int)~__funnelshift_r(s_wb[b0 >> 5], s_wb[(b0 >> 5) + 1u], b0 & 31u));
                    } else {
                        uint32_t b0 = (rt - 32u) & 511u;
                        len = (uint32_t)__clz((int)~__funnelshift_r(s_wb[2u + (b0 >> 5)], s_wb[2u + (((b0 >> 5) + 1u) & 15u)], b0 & 31u));
                        uint32_t run = len;
                        while (run == 32u && len <= max_len) {
                            b0 = (b0 - 32u) & 511u;
                            run = (uint32_t)__clz((int)~__funnelshift_r(s_wb[2u + (b0 >> 5)], s_wb[2u + (((b0 >> 5) + 1u) & 15u)], b0 & 31u));
                            len += run;
                        }
                    }
                    const int32_t s = (int32_t)(t - len);
                    if (len <= max_len && s >= dom_lo && s < dom_hi) {
                        const uint32_t poly = kShort ? s_G[63u + rt] - s_G[63u + rt - len] * s_pow[len]
                                                     : s_G[64u + ((rt - 1u) & 511u)] - s_G[64u + ((rt - len - 1u) & 511u)] * s_pow[len];
                        key = ww_poly_key(poly, len);
                        pass = !kBloom || ((s_bloom[key >> (bloom_shift + 5u)] >> ((key >> bloom_shift) & 31u)) & 1u);
                        packed = (uint32_t)s << 8 | len;
                    }
                }
                probe_q.push(pass, make_uint2(key, packed), lt_mask);
                if (probe_q.n >= 32u) {
                    probe(32u);
                    probe_q.pop32(lane);
                }
            }
            keep_tail();
        }
        probe(probe_q.n);
        probe_q.n = 0;
        verify(verify_q.n);
        verify_q.n = 0;
    }
}

// records of a row = set bits of its end bitmap
__global__ void __launch_bounds__(256) k_ww3_count(const uint32_t *hitbits, uint32_t *row_count, int64_t n_rows) {
    for (int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x; row < n_rows; row += (int64_t)gridDim.x * 256) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(hitbits + row * 8)), b = __ldg(reinterpret_cast<const uint4 *>(hitbits + row * 8) + 1);
        row_count[row] = (uint32_t)(__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w));
    }
}

// hit bits -> records.  A warp takes 32 rows at a time (a lane loads one row's 256 hit bits), the rows' hits are numbered
// by a warp scan, and every lane resolves one hit per step - 32 independent scans back to the word start in flight (one row
// per warp and one hit per lane-step left the kernel waiting on a single chain of dependent loads: 1.0 ms per 10^9 chars
// for 2 hits per 1 000 chars).
template <bool kIsMap>
__global__ void __launch_bounds__(256) k_ww3_emit(const DevWw W, const Ww3EmitArgs E) {
    __shared__ uint32_t s_bits[8][32 * 8 + 4];
    __shared__ uint32_t s_inc[8][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_groups = (E.n_rows + 31) / 32;
    for (int64_t group = (int64_t)blockIdx.x * 8 + warp; group < n_groups; group += (int64_t)gridDim.x * 8) {
        const int64_t row = group * 32 + lane;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (row < E.n_rows) {
            a = __ldg(reinterpret_cast<const uint4 *>(E.hitbits + row * 8));
            b = __ldg(reinterpret_cast<const uint4 *>(E.hitbits + row * 8) + 1);
        }
        const uint32_t cnt = (uint32_t)(__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w));
        const uint32_t inc = warp_inclusive_sum(cnt);
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (total == 0u) continue;
        __syncwarp();
        {
            uint32_t *dst = &s_bits[warp][lane * 8];
            dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
            dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
            s_inc[warp][lane] = inc;
        }
        __syncwarp();
        for (uint32_t h = lane; h < total; h += 32) {
            // the row of hit h: the first lane whose inclusive count exceeds h
            uint32_t r = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1)
                if (s_inc[warp][r + step - 1] <= h) r += step;
            uint32_t k = h - (r ? s_inc[warp][r - 1] : 0u);   // rank inside the row
            const uint32_t *bits = &s_bits[warp][r * 8];
            uint32_t j = 0, word = bits[0];
            while ((uint32_t)__popc(word) <= k) {
                k -= (uint32_t)__popc(word);
                word = bits[++j];
            }
            const uint32_t bit = __fns(word, 0u, (int)k + 1);
            const int64_t hit_row = group * 32 + r;
            const int64_t t = E.origin + hit_row * kW3Row + j * 32 + bit;
            // the run's start: the last start bit before the end (a keyword run is at most 255 chars: nine words back at most)
            int64_t rel = t - E.origin - 1, w_at = rel >> 5;
            uint32_t sw = __ldg(&E.startbits[w_at]) & (0xFFFFFFFFu >> (31u - (uint32_t)(rel & 31)));
            while (sw == 0u && w_at > 0) sw = __ldg(&E.startbits[--w_at]);
            const int64_t s = E.origin + w_at * 32 + (31 - __clz((int)sw));
            const unsigned long long idx = __ldg(E.block_excl + (hit_row >> 12)) + __ldg(E.row_excl + hit_row) + (h - (r ? s_inc[warp][r - 1] : 0u));
            if (idx < (unsigned long long)E.cap) {
                E.pos_out[idx] = make_int2((int32_t)s + E.pos_base, (int32_t)t + E.pos_base);
                if (kIsMap) E.val_out[idx] = __ldg(&E.val_scratch[(t - E.origin) >> 1]);
            }
        }
        __syncwarp();
    }
    static_assert(kScanRows == 4096, "row >> 12 is the scan block of a row");
}

}  // namespace acgpu
