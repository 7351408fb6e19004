// k_tier_pair<K, LOW, MIR>: generation 4 of the hit-mask kernel (same contract as k_tier_mask in kernel_mask.cuh: one
// 16-bit HIT MASK per position - bit 16 - d = a keyword of length d ends here, AhoCorasickSet.java:522-535 - and one
// record count per 256-position row).
//
// What limits k_tier_mask on a saturated dictionary (ncu, profiles/r02_s1_*): the L1 -> crossbar request port (one missed
// sector per cycle and SM: 82 % busy) and the issue slots (66 %); per position it pays one random 8-byte shared load and
// one gathered sector of continuation masks.  Here every table look-up answers a PAIR of neighbouring positions:
//   * levels 1..K: row = the j-1 classes the two level-j contexts share, {fwd, back} words (TierTables::prow_words) -
//     fwd is indexed by the class AFTER the shared ones (position q + 1), back by the class BEFORE them (position q);
//   * level K + 1: one gathered entry of kidmask at the level-K context of q = {children of that node (position q),
//     level-(K+1) nodes whose trailing K classes are that context (position q + 1)}.
// So a pair costs one 8-byte shared load per level that holds keywords and at most one gathered sector, the rolling
// mixed-radix row numbers are built once per pair, and the has-children pre-filter of k_tier_mask (useless when the
// level-K nodes of a big dictionary nearly all have children) shrinks to one GATE bit per row.  Continuing contexts are
// compacted into the warp's probe queue with ONE warp scan per row instead of one ballot round per queued position.
#pragma once
#include "kernel_mask.cuh"

namespace acgpu {

// class of row position i of the lane (i >= 0: own char, i < 0: the -i-th char before the lane's first)
#define ACGPU_CL(i) ((i) >= 0 ? c[(i) >= 0 ? (i) : 0] : pc[(i) < 0 ? -(i) : 0])

template <int K, int LOW, bool MIR>
__global__ void __launch_bounds__(kMaskThreads, 1) k_tier_pair(const DevAutomaton A, const DevTier T, const MaskArgs P) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint8_t *s_cls = reinterpret_cast<uint8_t *>(s_mem);
    const unsigned char *s_tab = reinterpret_cast<const unsigned char *>(s_mem + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u, sh = 1u << b;
    unsigned char *s_q = reinterpret_cast<unsigned char *>(s_mem + 64 + ((T.n_prow_words + 3u) & ~3u)) + (size_t)warp * kMaskQueue * kMaskQueueEntry;
    unsigned long long *s_qctx = reinterpret_cast<unsigned long long *>(s_q);
    uint32_t *s_qpos = reinterpret_cast<uint32_t *>(s_q + kMaskQueue * 8);

    for (uint32_t i = tid; i < 256; i += kMaskThreads) s_cls[i] = (uint8_t)((__ldg(&T.cls8[i >> 2]) >> ((i & 3) * 8)) & 0xFFu);
    {
        const uint32_t n4 = T.n_prow_words >> 2;
        const uint4 *src = reinterpret_cast<const uint4 *>(T.prow_words);
        uint4 *dst = reinterpret_cast<uint4 *>(s_mem + 64);
#pragma unroll 4
        for (uint32_t i = tid; i < n4; i += kMaskThreads) dst[i] = __ldg(src + i);
        for (uint32_t i = (n4 << 2) + tid; i < T.n_prow_words; i += kMaskThreads) s_mem[64 + i] = __ldg(&T.prow_words[i]);
    }
    __syncthreads();

    const bool deeper = T.kidmask != nullptr;  // some keyword is longer than K
    const uint32_t gate_bit = T.pair_gate_bit, low_bit = T.pair_low_bit;
    uint32_t poff[K + 1];  // byte offsets of the levels' rows ([0]: the compact fwd words of level K-1)
#pragma unroll
    for (int i = 0; i <= K; i++) poff[i] = T.prow_off[i] * 4u;
    uint32_t q_cnt = 0;

    auto probe = [&](uint32_t first, uint32_t count) {
        if ((uint32_t)lane < count)
            deep_resolve<K>(T.buckets, T.hash_seed, T.n_buckets, T.b, T.inv_b, s_qctx[first + lane], s_qpos[first + lane], A.max_len,
                            P.masks, P.row_count);
    };
    const int t_rows = P.chunk_rows > 0 ? P.chunk_rows : kMaskChunkRows;
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(P.ticket, 1u);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        const int64_t row0 = (int64_t)chunk * t_rows;
        if (row0 >= P.n_rows) break;
        const int n_cr = (int)min((int64_t)t_rows, P.n_rows - row0);
        const int64_t c_lo = P.origin + row0 * kMaskRow, c_hi = c_lo + (int64_t)n_cr * kMaskRow;
        const bool chunk_in = c_lo - 16 >= 0 && c_hi <= P.n;
        const bool chunk_full = c_lo >= P.emit_from && c_hi <= P.emit_to;
        const int64_t l_lo = c_lo + (int64_t)lane * 8;
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
        Pack8 car0, car1;  // the 8 classes ending 8 positions before the row / right before the row
        {
            Pack8 h{0u, 0u};
            if (lane < 2) {
                const int64_t p0 = c_lo - 16 + (int64_t)lane * 8;
                const bool in = p0 >= 0 && p0 + 8 <= P.n;
                const uint4 hv = load8<MIR>(P.hay, P.n, p0, in);
                uint32_t h1[8];
                classify8x4<MIR, 1>(A, P.hay, P.n, p0, in, hv, s_cls, h1);
                h = pack8<1>(h1, sh);
            }
            car0.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 0); car0.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 0);
            car1.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 1); car1.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 1);
        }
        uint4 *mp = reinterpret_cast<uint4 *>(P.masks) + (MIR ? (size_t)(P.n_rows - 1 - row0) * 32 + (31 - lane) : (size_t)row0 * 32 + lane);
        const uint32_t q0 = (uint32_t)(row0 * kMaskRow) + lane * 8;
        for (int r = 0; r < n_cr; ++r) {
            const uint4 vnn = fetch(r + 2);
            uint32_t c[8];  // classes of the lane's 8 positions
            if (chunk_in) {
                classify8x4<MIR, 1>(A, P.hay, P.n, 0, true, v, s_cls, c);
            } else {
                const int64_t p0 = P.origin + (row0 + r) * kMaskRow + (int64_t)lane * 8;
                classify8x4<MIR, 1>(A, P.hay, P.n, p0, p0 >= 0 && p0 + 8 <= P.n, v, s_cls, c);
            }
            v = vn;
            vn = vnn;
            const Pack8 P0 = pack8<1>(c, sh);
            Pack8 P1, P2;
            P1.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 1); P1.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 1);
            P2.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 2); P2.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 2);
            if (lane == 0) { P1 = car1; P2 = car0; }
            if (lane == 1) P2 = car1;
            car0.hi = __shfl_sync(0xFFFFFFFFu, P0.hi, 30); car0.lo = __shfl_sync(0xFFFFFFFFu, P0.lo, 30);
            car1.hi = __shfl_sync(0xFFFFFFFFu, P0.hi, 31); car1.lo = __shfl_sync(0xFFFFFFFFu, P0.lo, 31);
            uint32_t pc[K + 1];  // pc[i] = class i positions before the lane's first one
            pc[0] = 0;
#pragma unroll
            for (int i = 1; i <= K; i++) pc[i] = i <= 4 ? (P1.lo >> (b * (i - 1))) & cm : (P1.hi >> (b * (i - 5))) & cm;

            uint32_t m[8];
#pragma unroll
            for (int j = 0; j < 8; j++) m[j] = 0u;
            uint32_t pm = 0;  // bit j: the context of position j continues to level K + 1
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int q = 2 * p;  // the pair is (q, q + 1)
                // row[j] = mixed-radix number of c[q], .., c[q-j+2] (c[q] lowest digit); the level-j row of the pair
                uint32_t row[K + 1];
                row[0] = 0;
                row[1] = 0;
#pragma unroll
                for (int j = 2; j <= K; j++) row[j] = j == 2 ? ACGPU_CL(q) : row[j - 1] + ACGPU_CL(q - j + 2) * T.pow_c[j - 1];
                uint2 wk = make_uint2(0u, 0u);
#pragma unroll
                for (int j = 1; j <= K; j++) {
                    if (j < K) {
                        if (LOW != 0) continue;  // LOW 1: level K-1 rides in the level-K row and the compact table (below)
                        if (!((T.term_levels >> j) & 1u)) continue;  // warp-uniform
                    }
                    const uint2 w = *reinterpret_cast<const uint2 *>(s_tab + poff[j] + row[j] * 8u);
                    m[q + 1] |= rotr32(w.x, ACGPU_CL(q + 1)) & (1u << (16 - j));
                    m[q] |= rotr32(w.y, ACGPU_CL(q - j + 1)) & (1u << (16 - j));
                    if (j == K) wk = w;
                }
                if (LOW == 1) {
                    // level K-1: position q from the LOW bit of the row its K-1 classes name, position q + 1 from the
                    // compact fwd table (one 4-byte load)
                    m[q] |= (wk.x & low_bit) ? (1u << (17 - K)) : 0u;
                    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(s_tab + poff[0] + row[K >= 2 ? K - 1 : 0] * 4u);
                    m[q + 1] |= rotr32(w1, ACGPU_CL(q + 1)) & (1u << (17 - K));
                }
                // one gather for the pair: the entry of the level-K context of q
                const uint32_t cy = ACGPU_CL(q - K + 1);
                const bool g = deeper && (wk.x & gate_bit) != 0u && cy != 0u;
                uint2 kq = make_uint2(0u, 0u);
#if !(ACGPU_ABL & 1)
                if (g) kq = tex1Dfetch<uint2>(T.kid_tex, (int)(row[K] + cy * T.pow_c[K]));
#endif
                pm |= ((kq.x >> ACGPU_CL(q - K)) & 1u) << q;
                pm |= ((kq.y >> ACGPU_CL(q + 1)) & 1u) << (q + 1);
            }
            // ---- positions outside [emit_from, emit_to) report nothing (edge rows only)
            if (!chunk_full) {
                const int64_t p0 = P.origin + (row0 + r) * kMaskRow + (int64_t)lane * 8;
                const int64_t lo_j = P.emit_from - p0, hi_j = P.emit_to - p0;
                const uint32_t a = lo_j <= 0 ? 0xFFu : (lo_j >= 8 ? 0u : (0xFFu << (int)lo_j) & 0xFFu);
                const uint32_t z = hi_j >= 8 ? 0xFFu : (hi_j <= 0 ? 0u : (0xFFu >> (8 - (int)hi_j)));
                const uint32_t vm = a & z;
#pragma unroll
                for (int j = 0; j < 8; j++) m[j] = (vm >> j) & 1u ? m[j] : 0u;
                pm &= vm;
            }
#if ACGPU_ABL & 2
            m[0] |= (pm * 0x01010101u) >> 31;
            pm = 0u;
#endif
            // ---- store the shallow masks and the row count; deep hits are OR-ed in later by this same warp
            const uint4 mw = make_uint4(m[0] | m[1] << 16, m[2] | m[3] << 16, m[4] | m[5] << 16, m[6] | m[7] << 16);
            if (MIR)
                *(mp - r * 32) = make_uint4(swap16(mw.w), swap16(mw.z), swap16(mw.y), swap16(mw.x));
            else
                *(mp + r * 32) = mw;
            const uint32_t cnt = __popc(mw.x) + __popc(mw.y) + __popc(mw.z) + __popc(mw.w);
            const uint32_t row_total = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) P.row_count[MIR ? P.n_rows - 1 - (row0 + r) : row0 + r] = row_total;
            __syncwarp();
            // ---- queue the continuing contexts (one warp scan gives every lane its slots); probe in batches of 32
            if (__ballot_sync(0xFFFFFFFFu, pm != 0u)) {
                const unsigned long long own8 = pack64(P0, b), prev16 = (pack64(P2, b) << (8 * b)) | pack64(P1, b);
                const uint32_t qb = q0 + (uint32_t)(r * kMaskRow);
                const uint32_t n_mine = (uint32_t)__popc(pm);
                uint32_t inc = n_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += y;
                }
                const uint32_t n_all = __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (q_cnt + n_all <= (uint32_t)kMaskQueue) {
                    uint32_t slot = q_cnt + inc - n_mine;
                    while (pm) {
                        const int j = __ffs(pm) - 1;
                        pm &= pm - 1u;
                        s_qctx[slot] = (prev16 << (b * (j + 1))) | (own8 >> (b * (7 - j)));
                        const uint32_t qp = qb + (uint32_t)j;
                        s_qpos[slot] = MIR ? (uint32_t)(P.n_rows * kMaskRow) - 1u - qp : qp;
                        ++slot;
                    }
                    q_cnt += n_all;
                    __syncwarp();
                    while (q_cnt >= 32u) {
                        q_cnt -= 32u;
                        probe(q_cnt, 32u);
                        __syncwarp();
                    }
                } else {
                    // a row with more continuing contexts than the queue has room for: one ballot round per position
                    while (true) {
                        const bool has = pm != 0u;
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
                        if (!bal) break;
                        if (has) {
                            const int j = __ffs(pm) - 1;
                            pm &= pm - 1u;
                            const uint32_t slot = q_cnt + __popc(bal & ((1u << lane) - 1u));
                            s_qctx[slot] = (prev16 << (b * (j + 1))) | (own8 >> (b * (7 - j)));
                            const uint32_t qp = qb + (uint32_t)j;
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
            }
        }
    }
    if (q_cnt) probe(0u, q_cnt);
}

#undef ACGPU_CL

}  // namespace acgpu
