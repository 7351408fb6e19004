// Generation-3 AhoCorasick path for narrow alphabets: three launches per match, no ordered coupling inside a kernel.
//
//   k_tier_mask<K, LOW>  every haystack position q is an END anchor (all keywords ending at q, longest first —
//                        AhoCorasickSet.java:522-535).  The kernel writes one 16-bit HIT MASK per position (bit 16 - d
//                        set = a keyword of length d ends here) and one record count per 256-position row.
//   k_row_scan           exclusive scan of the row counts (the reference's output order is position-major, so a row's
//                        records start at the sum of the counts of all rows before it).
//   k_tier_emit<isMap>   expands the masks into (start, end[, value]) records at their final offsets — the same ordered
//                        stream the reference's listener sees (end ascending, longest first).
//
// k_tier_mask: a warp owns a CHUNK of 32 consecutive rows (dynamic tickets); a lane owns 8 consecutive positions of a row
// (one streaming 128-bit load, prefetched two rows ahead) and the classes of the 16 positions before them arrive by warp
// shuffle — from the previous row of the same warp for lanes 0 and 1, so nothing is loaded twice.  Levels 1..K come from
// row-layout bit tables in shared memory (row = mixed-radix number of the previous classes, kept pre-scaled and
// rolling in registers; bit = current class; one rotate drops the bit onto its place in the hit mask).  Contexts
// that continue past level K (exact child mask, one L2-resident word) are queued with their packed context and probed
// in full batches of 32 against the path-compressed deep table; their hits are OR-ed into the already stored masks.
#pragma once
#include "kernel_tier.cuh"

namespace acgpu {

#ifndef ACGPU_MASK_WARPS
#define ACGPU_MASK_WARPS 32
#endif
constexpr int kMaskWarps = ACGPU_MASK_WARPS;
#ifndef ACGPU_ABL
#define ACGPU_ABL 0       // ablation builds (tools/gpu_exp.sh VARIANTS): 1 no continuation-mask gather, 2 no deep probes
#endif
#ifndef ACGPU_KID_TEX
#define ACGPU_KID_TEX 1   // child masks are gathered through the texture pipe (the LSU data pipe is the kernel's bottleneck)
#endif
// bytes per queue entry: the 64-bit context and the position.  (Keeping the packed classes instead and shifting the
// context together at probe time - 20-byte entries, 201 KB of shared memory - pushed the carve-out to 228 KB and the
// kernel from 2.80 to 4.79 ms: the 28 KB of L1 left do not hold the gathers' working set.  Measured, session 5.)
constexpr int kMaskQueueEntry = 12;
constexpr int kMaskThreads = kMaskWarps * 32;
constexpr int kMaskRow = 256;          // positions per warp row
constexpr int kMaskChunkRows = 32;     // rows per ticket
constexpr int kMaskQueue = 64;         // deep-probe queue entries per warp (a round adds at most 32 to at most 31)
constexpr int kScanRows = 4096;        // rows per k_row_scan block
constexpr int kEmitWarps = 8;
constexpr int kEmitStage = 384;        // records staged per warp before a coalesced flush

struct MaskArgs {
    const uint16_t *hay;
    int64_t n;
    int64_t emit_from;      // positions (index of a keyword's last char) in [emit_from, emit_to) report
    int64_t emit_to;
    int64_t origin;         // first position of row 0: <= emit_from and hay + origin is 16-byte aligned
    uint32_t *masks;        // [n_rows * 128] words, two positions per word
    uint32_t *row_count;    // [n_rows]
    unsigned int *ticket;
    int64_t n_rows;
    int32_t chunk_rows;     // rows per ticket, 1 .. kMaskChunkRows (0 = kMaskChunkRows): short haystacks take small tickets so that every warp of the grid gets one
};

struct ScanArgs {
    uint32_t *row_count;              // in: counts, out: exclusive prefix inside the row's block of kScanRows rows
    unsigned long long *block_excl;   // [n_blocks] exclusive prefix of the block
    unsigned int *done;               // blocks finished
    unsigned long long *total_out;
    int64_t n_rows;
    const unsigned long long *base_in;   // records before row 0 (slab runs: the total of the slabs before this one); nullptr: 0
};

struct EmitArgs {
    const uint16_t *hay;
    int64_t n;
    const uint32_t *masks;
    const uint32_t *row_excl;
    const unsigned long long *block_excl;
    int64_t n_rows;
    int64_t origin;
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
};

__host__ __device__ constexpr size_t mask_smem_bytes(size_t n_row_words) {
    return (64 + ((n_row_words + 3) & ~size_t(3))) * sizeof(uint32_t) + (size_t)kMaskWarps * kMaskQueue * kMaskQueueEntry;
}

__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t s) { return __funnelshift_r(x, x, s); }

__device__ __forceinline__ uint32_t ldg_u32_if(const void *p, bool on) {
    uint32_t x = 0;
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.global.nc.u32 %0, [%1];\n}" : "+r"(x) : "l"(p), "r"((uint32_t)on));
    return x;
}

// scaled classes (class * SCALE; s_cls4 holds the first 256 code units at the same scale) of the 8 chars at kernel
// positions [p0, p0 + 8); v is the prefetched vector, valid when `inside` (the 8 chars lie inside [0, n)); positions
// outside [0, n) give class 0
template <bool MIR = false, int SCALE = 4>
__device__ __forceinline__ void classify8x4(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t p0, bool inside, const uint4 v,
                                            const uint8_t *s_cls4, uint32_t (&c4)[8]) {
    if (inside) {
        if (((v.x | v.y | v.z | v.w) & 0xFF00FF00u) == 0u) {
            c4[0] = s_cls4[v.x & 0xFFu]; c4[1] = s_cls4[v.x >> 16];
            c4[2] = s_cls4[v.y & 0xFFu]; c4[3] = s_cls4[v.y >> 16];
            c4[4] = s_cls4[v.z & 0xFFu]; c4[5] = s_cls4[v.z >> 16];
            c4[6] = s_cls4[v.w & 0xFFu]; c4[7] = s_cls4[v.w >> 16];
        } else {
            const uint32_t ch[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                                    v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
#pragma unroll
            for (int j = 0; j < 8; j++) c4[j] = (uint32_t)__ldg(&A.cls[ch[j]]) * (uint32_t)SCALE;  // no test per char: the Latin-1 part of the table stays in L1 (a select between the two tables cost 17 instructions per char: 15 % of k_tier_emit<1> on configs[1])
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t p = p0 + j;
            c4[j] = (p >= 0 && p < n) ? (uint32_t)__ldg(&A.cls[__ldg(&hay[MIR ? n - 1 - p : p])]) * (uint32_t)SCALE : 0u;
        }
    }
}

// The 8 chars of kernel positions [p0, p0 + 8).  MIR: kernel position q' is haystack position n - 1 - q' (the haystack is
// read right to left, so an END anchor of the kernel is a START anchor of the haystack and the tables are those of the
// forward trie): one aligned 128-bit load of hay[n - 8 - p0, n - p0), chars reversed in registers.
__device__ __forceinline__ uint32_t swap16(uint32_t x) { return __byte_perm(x, 0u, 0x1032u); }
template <bool MIR>
__device__ __forceinline__ uint4 load8(const uint16_t *hay, int64_t n, int64_t p0, bool on) {
    if (MIR) {
        const uint4 v = ldcs_v4_if(hay + (n - 8 - p0), on);
        return make_uint4(swap16(v.w), swap16(v.z), swap16(v.y), swap16(v.x));
    }
    return ldcs_v4_if(hay + p0, on);
}

// The classes of 8 consecutive positions packed b bits each, most recent lowest, as two words of 4 classes.
struct Pack8 {
    uint32_t hi, lo;  // hi: positions 0..3, lo: positions 4..7 (position 7 = most recent = lowest bits of lo)
};
template <int SCALE = 4>
__device__ __forceinline__ Pack8 pack8(const uint32_t (&c4)[8], uint32_t sh) {
    Pack8 p;
    // every c4 is a multiple of SCALE, so the scaled sum is SCALE x the packed classes
    p.hi = (((c4[0] * sh + c4[1]) * sh + c4[2]) * sh + c4[3]) / (uint32_t)SCALE;
    p.lo = (((c4[4] * sh + c4[5]) * sh + c4[6]) * sh + c4[7]) / (uint32_t)SCALE;
    return p;
}
__device__ __forceinline__ unsigned long long pack64(const Pack8 &p, int b) { return ((unsigned long long)p.hi << (4 * b)) | p.lo; }

// Context (the last classes, most recent lowest) of the lane's position j given the lane's own 8 classes P0 and the 16
// before them (P1 most recent).  Only the low b * max_len bits matter to the callers.
__device__ __forceinline__ unsigned long long context_of(const Pack8 &P0, const Pack8 &P1, const Pack8 &P2, int j, int b) {
    const unsigned long long q = (pack64(P2, b) << (8 * b)) | pack64(P1, b);
    return (q << (b * (j + 1))) | (pack64(P0, b) >> (b * (7 - j)));
}

// One queued context: levels K+1.. against the deep table; hits are OR-ed into the stored mask of the position and
// added to its row's count.
template <int K>
__device__ __noinline__ uint32_t deep_resolve(const uint4 *buckets, unsigned long long hash_seed, uint32_t n_buckets, int b,
                                              uint32_t inv_b, unsigned long long ctx, uint32_t pos, int max_len, uint32_t *masks,
                                              uint32_t *row_count) {
    DevTier T;  // only the fields deep_bits reads
    T.buckets = buckets;
    T.hash_seed = hash_seed;
    T.n_buckets = n_buckets;
    T.b = b;
    T.inv_b = inv_b;
    const uint32_t bits = deep_bits<K>(T, ctx, (1u << b) - 1u, max_len);
    if (bits) {
        atomicOr(masks + (pos >> 1), (__brev(bits) >> (16 + K)) << ((pos & 1u) * 16u));
        if (row_count) atomicAdd(row_count + (pos >> 8), (uint32_t)__popc(bits));  // fused runs (kernel_fuse.cuh) count per ticket instead
    }
    return (uint32_t)__popc(bits);
}

// How a warp of the mask kernel gets its work.  PlainSched: tickets of kMaskChunkRows rows from one counter, masks and row
// counts at the rows' own places (k_tier_mask, three launches per match).  kernel_fuse.cuh has the scheduler of the
// single-launch path, where a warp alternates between making masks and expanding them.
struct PlainSched {
    static constexpr bool kFused = false;
    unsigned int *ticket;
    int chunk_rows;
    __device__ __forceinline__ uint32_t next(int lane) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(ticket, 1u);
        return __shfl_sync(0xFFFFFFFFu, chunk, 0);
    }
    __device__ __forceinline__ int rows() const { return chunk_rows; }
    __device__ __forceinline__ int64_t mask_row0(uint32_t, int64_t row0) const { return row0; }
    __device__ __forceinline__ void finish(uint32_t, uint32_t, int) {}
};

template <int K, int LOW, bool MIR, class Sched>
__device__ __forceinline__ void tier_mask_body(const DevAutomaton &A, const DevTier &T, const MaskArgs &P, Sched &S, uint32_t *s_mem) {
    uint8_t *s_cls4 = reinterpret_cast<uint8_t *>(s_mem);
    const unsigned char *s_tab = reinterpret_cast<const unsigned char *>(s_mem + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u, sh = 1u << b, C = (uint32_t)T.C;
    // per-warp queue of contexts that continue past level K
    unsigned char *s_q = reinterpret_cast<unsigned char *>(s_mem + 64 + ((T.n_row_words + 3u) & ~3u)) + (size_t)warp * kMaskQueue * kMaskQueueEntry;
    unsigned long long *s_qctx = reinterpret_cast<unsigned long long *>(s_q);
    uint32_t *s_qpos = reinterpret_cast<uint32_t *>(s_q + kMaskQueue * 8);

    for (uint32_t i = tid; i < 256; i += kMaskThreads) s_cls4[i] = (uint8_t)(((__ldg(&T.cls8[i >> 2]) >> ((i & 3) * 8)) & 0xFFu) * 4u);
    {
        const uint32_t n4 = T.n_row_words >> 2;  // the table is 16-byte aligned on both sides
        const uint4 *src = reinterpret_cast<const uint4 *>(T.row_words);
        uint4 *dst = reinterpret_cast<uint4 *>(s_mem + 64);
#pragma unroll 4
        for (uint32_t i = tid; i < n4; i += kMaskThreads) dst[i] = __ldg(src + i);
        for (uint32_t i = (n4 << 2) + tid; i < T.n_row_words; i += kMaskThreads) s_mem[64 + i] = __ldg(&T.row_words[i]);
    }
    __syncthreads();

    const bool deeper = T.kidmask != nullptr;  // some keyword is longer than K
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t roff[K + 1];
#pragma unroll
    for (int i = 1; i <= K; i++) roff[i] = T.row_off[i] * 4u;
#if !ACGPU_KID_TEX
    const unsigned char *kid_bytes = reinterpret_cast<const unsigned char *>(T.kidmask);
#endif
    uint32_t q_cnt = 0;
    uint32_t deep_hits = 0, shallow_hits = 0;  // fused runs: records of the current ticket (per lane / warp-uniform)

    // entries [first, first + count) of the queue (count <= 32) against the deep table.  (A separate gather kernel fed
    // from a candidate list measured 3% slower end to end and costs 1.5 GB of scratch per 10^9 chars.)
    auto probe = [&](uint32_t first, uint32_t count) {
        if ((uint32_t)lane < count)
            deep_hits += deep_resolve<K>(T.buckets, T.hash_seed, T.n_buckets, T.b, T.inv_b, s_qctx[first + lane], s_qpos[first + lane], A.max_len,
                                         P.masks, Sched::kFused ? nullptr : P.row_count);
    };
    const int t_rows = S.rows();
    while (true) {
        const uint32_t chunk = S.next(lane);
        const int64_t row0 = (int64_t)chunk * t_rows;
        if (chunk == 0xFFFFFFFFu || row0 >= P.n_rows) break;
        const int n_cr = (int)min((int64_t)t_rows, P.n_rows - row0);  // rows of this chunk
        const int64_t m_row0 = S.mask_row0(chunk, row0);               // where the chunk's masks go (fused runs: a ring)
        // 64-bit position arithmetic once per chunk; rows use 32-bit offsets from here
        const int64_t c_lo = P.origin + row0 * kMaskRow, c_hi = c_lo + (int64_t)n_cr * kMaskRow;
        const bool chunk_in = c_lo - 16 >= 0 && c_hi <= P.n;                    // every load of the chunk is inside the haystack
        const bool chunk_full = c_lo >= P.emit_from && c_hi <= P.emit_to;        // no row is cut by the emit range
        const int64_t l_lo = c_lo + (int64_t)lane * 8;                           // the lane's first position in row 0 of the chunk
        const uint16_t *lp = MIR ? P.hay + (P.n - 8 - l_lo) : P.hay + l_lo;
        auto fetch = [&](int r) -> uint4 {
            const bool live = r < n_cr;
            if (chunk_in) {
                const uint4 x = ldcs_v4_if(MIR ? lp - r * kMaskRow : lp + r * kMaskRow, live);
                return MIR ? make_uint4(swap16(x.w), swap16(x.z), swap16(x.y), swap16(x.x)) : x;
            }
            const int64_t p0 = P.origin + (row0 + r) * kMaskRow + (int64_t)lane * 8;
            return load8<MIR>(P.hay, P.n, p0, live && p0 >= 0 && p0 + 8 <= P.n);
        };
        uint4 v = fetch(0), vn = fetch(1);
        // left context of the chunk: lanes 0, 1 classify the 16 chars before it
        Pack8 car0, car1;  // the 8 classes ending 8 positions before the row / right before the row
        {
            Pack8 h{0u, 0u};
            if (lane < 2) {
                const int64_t p0 = c_lo - 16 + (int64_t)lane * 8;
                const bool in = p0 >= 0 && p0 + 8 <= P.n;
                const uint4 hv = load8<MIR>(P.hay, P.n, p0, in);
                uint32_t h4[8];
                classify8x4<MIR>(A, P.hay, P.n, p0, in, hv, s_cls4, h4);
                h = pack8(h4, sh);
            }
            car0.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 0); car0.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 0);
            car1.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 1); car1.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 1);
        }
        // per-chunk bases of the outputs
        uint4 *mp = reinterpret_cast<uint4 *>(P.masks) + (MIR ? (size_t)(P.n_rows - 1 - row0) * 32 + (31 - lane) : (size_t)m_row0 * 32 + lane);
        const uint32_t q0 = (uint32_t)(m_row0 * kMaskRow) + lane * 8;
        for (int r = 0; r < n_cr; ++r) {
            const uint4 vnn = fetch(r + 2);
            uint32_t c4[8];
            if (chunk_in) {
                classify8x4<MIR>(A, P.hay, P.n, 0, true, v, s_cls4, c4);
            } else {
                const int64_t p0 = P.origin + (row0 + r) * kMaskRow + (int64_t)lane * 8;
                classify8x4<MIR>(A, P.hay, P.n, p0, p0 >= 0 && p0 + 8 <= P.n, v, s_cls4, c4);
            }
            v = vn;
            vn = vnn;
            const Pack8 P0 = pack8(c4, sh);
            Pack8 P1, P2;
            P1.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 1); P1.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 1);
            P2.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 2); P2.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 2);
            if (lane == 0) { P1 = car1; P2 = car0; }
            if (lane == 1) P2 = car1;
            car0.hi = __shfl_sync(0xFFFFFFFFu, P0.hi, 30); car0.lo = __shfl_sync(0xFFFFFFFFu, P0.lo, 30);
            car1.hi = __shfl_sync(0xFFFFFFFFu, P0.hi, 31); car1.lo = __shfl_sync(0xFFFFFFFFu, P0.lo, 31);
            // pc[i] = class i positions before this lane's first one (i = 1..K)
            uint32_t pc[K + 1];
            pc[0] = 0;
#pragma unroll
            for (int i = 1; i <= K; i++) pc[i] = i <= 4 ? (P1.lo >> (b * (i - 1))) & cm : (P1.hi >> (b * (i - 5))) & cm;

            // ---- levels 1..K.  rs[k] = 4 * (mixed-radix number of the k classes before the current position)
            uint32_t rs[K + 1];
            rs[0] = 0;
            {
                uint32_t acc = 0;
#pragma unroll
                for (int k = 1; k < K; k++) {
                    acc += pc[k] * (T.pow_c[k] * 4u);
                    rs[k] = acc;
                }
            }
            // one gather per PAIR of positions: the entry of position j's context holds the backward continuation mask
            // (position j) and the forward one (position j + 1) - 0.275 instead of 0.49 gathers per position on configs[4]
            uint32_t m[8], need = 0, rkv[4];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t cj = c4[j] >> 2;
                uint32_t mj = 0;
#pragma unroll
                for (int i = 1; i < K; i++) {
                    if (LOW != 0) continue;  // LOW 1: level K-1 rides in the level-K rows (below); LOW 2: nothing there
                    if ((T.term_levels >> i) & 1u) {
                        const uint32_t w = *reinterpret_cast<const uint32_t *>(s_tab + roff[i] + rs[i - 1]);
                        mj |= rotr32(w, cj) & (1u << (16 - i));
                    }
                }
                const uint2 wk = *reinterpret_cast<const uint2 *>(s_tab + roff[K] + rs[K - 1] * 2u);
                mj |= rotr32(wk.x, cj) & (1u << (16 - K));
                // bit 31 of the second word: the K-1 classes that name this row are a keyword (it ends one position back)
                if (LOW == 1 && j >= 1) m[j - 1] |= (wk.y >> (14 + K)) & (1u << (17 - K));
                // class K positions back: the child the context needs (class 0 = "in no keyword" never has one)
                const uint32_t ck = j >= K ? c4[j >= K ? j - K : 0] >> 2 : pc[j >= K ? 0 : K - j];
                if (((wk.y >> cj) & 1u) && ck != 0u) need |= 1u << j;
                if (!(j & 1)) rkv[j >> 1] = rs[K - 1] * C + c4[j];  // byte offset / 2 of the context's entry
#pragma unroll
                for (int k = K - 1; k >= 2; k--) rs[k] = rs[k - 1] * C + c4[j];
                if (K >= 2) rs[1] = c4[j];
                m[j] = mj;
            }
            uint2 kq[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const bool g = deeper && ((need >> (2 * p)) & 3u) != 0u;
#if ACGPU_ABL & 1
                kq[p] = make_uint2(0u, 0u);
#elif ACGPU_KID_TEX
                kq[p] = g ? tex1Dfetch<uint2>(T.kid_tex, (int)(rkv[p] >> 2)) : make_uint2(0u, 0u);
#else
                kq[p] = make_uint2(0u, 0u);
                if (g) kq[p] = __ldg(reinterpret_cast<const uint2 *>(kid_bytes + rkv[p] * 2u));
#endif
            }
            if (LOW == 1) {  // position 7's level K-1 bit sits in the row the NEXT position would read
                const uint32_t y = *reinterpret_cast<const uint32_t *>(s_tab + roff[K] + rs[K - 1] * 2u + 4u);
                m[7] |= (y >> (14 + K)) & (1u << (17 - K));
            }
            // ---- positions outside [emit_from, emit_to) report nothing (edge rows only)
            uint32_t vm = 0xFFu;
            if (!chunk_full) {
                const int64_t p0 = P.origin + (row0 + r) * kMaskRow + (int64_t)lane * 8;
                const int64_t lo_j = P.emit_from - p0, hi_j = P.emit_to - p0;
                const uint32_t a = lo_j <= 0 ? 0xFFu : (lo_j >= 8 ? 0u : (0xFFu << (int)lo_j) & 0xFFu);
                const uint32_t z = hi_j >= 8 ? 0xFFu : (hi_j <= 0 ? 0u : (0xFFu >> (8 - (int)hi_j)));
                vm = a & z;
#pragma unroll
                for (int j = 0; j < 8; j++) m[j] = (vm >> j) & 1u ? m[j] : 0u;
            }
            // ---- which contexts continue to level K + 1
            uint32_t pm = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int j = 2 * p;
                const uint32_t ck = j >= K ? c4[j >= K ? j - K : 0] >> 2 : pc[j >= K ? 0 : K - j];
                pm |= ((kq[p].x >> ck) & 1u) << j;
                pm |= ((kq[p].y >> (c4[j + 1] >> 2)) & 1u) << (j + 1);
            }
            pm &= vm;
#if ACGPU_ABL & 2
            m[0] |= (pm * 0x01010101u) >> 31;  // keep the gathers alive, never queue
            pm = 0u;
#endif
            const unsigned long long own8 = pack64(P0, b), prev16 = (pack64(P2, b) << (8 * b)) | pack64(P1, b);
            // ---- store the shallow masks and the row count; deep hits are OR-ed in later by this same warp
            const uint4 mw = make_uint4(m[0] | m[1] << 16, m[2] | m[3] << 16, m[4] | m[5] << 16, m[6] | m[7] << 16);
            if (MIR)  // masks are stored in haystack order: the array is the kernel's, reversed
                *(mp - r * 32) = make_uint4(swap16(mw.w), swap16(mw.z), swap16(mw.y), swap16(mw.x));
            else
                *(mp + r * 32) = mw;
            const uint32_t cnt = __popc(mw.x) + __popc(mw.y) + __popc(mw.z) + __popc(mw.w);
            const uint32_t row_total = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (Sched::kFused)
                shallow_hits += row_total;
            else if (lane == 0)
                P.row_count[MIR ? P.n_rows - 1 - (row0 + r) : row0 + r] = row_total;
            __syncwarp();
            // ---- queue the continuing contexts; probe whenever 32 are waiting
            while (true) {
                const bool has = pm != 0u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
                if (!bal) break;
                if (has) {
                    const int j = __ffs(pm) - 1;
                    pm &= pm - 1u;
                    const uint32_t slot = q_cnt + __popc(bal & lt_mask);
                    s_qctx[slot] = (prev16 << (b * (j + 1))) | (own8 >> (b * (7 - j)));
                    const uint32_t qp = q0 + (uint32_t)(r * kMaskRow) + j;
                    s_qpos[slot] = MIR ? (uint32_t)(P.n_rows * kMaskRow) - 1u - qp : qp;
                }
                q_cnt += __popc(bal);
                __syncwarp();
                if (q_cnt >= 32u) {
                    q_cnt -= 32u;
                    probe(q_cnt, 32u);
                    __syncwarp();
                }
            }
        }
        if (Sched::kFused) {
            // the ticket's masks must be final before it is handed to the emit side: drain the queue now
            if (q_cnt) probe(0u, q_cnt);
            q_cnt = 0;
            const uint32_t total = shallow_hits + __reduce_add_sync(0xFFFFFFFFu, deep_hits);
            shallow_hits = 0;
            deep_hits = 0;
            S.finish(chunk, total, lane);
        }
    }
    if (q_cnt) probe(0u, q_cnt);
}

template <int K, int LOW, bool MIR>
__global__ void __launch_bounds__(kMaskThreads, 1) k_tier_mask(const DevAutomaton A, const DevTier T, const MaskArgs P) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    PlainSched S{P.ticket, P.chunk_rows > 0 ? P.chunk_rows : kMaskChunkRows};
    tier_mask_body<K, LOW, MIR>(A, T, P, S, s_mem);
}

}  // namespace acgpu
