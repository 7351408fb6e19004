// k_row_scan and k_tier_emit: the second and third launch of the generation-3 AhoCorasick path (see kernel_mask.cuh).
#pragma once
#include "kernel_mask.cuh"

namespace acgpu {

// Exclusive scan of the row counts: every block scans kScanRows rows in place; the last block to finish scans the
// block totals.
static __global__ void __launch_bounds__(1024, 1) k_row_scan(const ScanArgs S) {
    __shared__ unsigned long long s_warp[32];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanRows + (int64_t)tid * 4;
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = base + k < S.n_rows ? S.row_count[base + k] : 0u;
    const uint32_t mine = c[0] + c[1] + c[2] + c[3];
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (uint32_t)s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += y;
        }
        s_warp[lane] = ((unsigned long long)wi << 32) | (wi - w);  // inclusive | exclusive
    }
    __syncthreads();
    uint32_t ex = (uint32_t)s_warp[warp] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < S.n_rows) S.row_count[base + k] = ex;
        ex += c[k];
    }
    if (tid == 0) {
        S.block_excl[blockIdx.x] = s_warp[31] >> 32;  // block total for now
        __threadfence();
        s_last = atomicAdd(S.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- block totals -> exclusive prefixes (one block; each thread takes a contiguous slice)
    const int nb = (int)gridDim.x;
    const int per = (nb + 1023) / 1024;
    const int lo = min(nb, tid * per), hi = min(nb, lo + per);
    unsigned long long sum = 0;
    volatile unsigned long long *be = S.block_excl;
    for (int i = lo; i < hi; i++) sum += be[i];
    unsigned long long sinc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, sinc, o);
        if (lane >= o) sinc += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = sinc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = s_warp[lane];
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += y;
        }
        const unsigned long long before = S.base_in ? *S.base_in : 0ull;
        s_warp[lane] = before + wi - w;
        if (lane == 31) *S.total_out = before + wi;
    }
    __syncthreads();
    unsigned long long run = s_warp[warp] + sinc - sum;
    for (int i = lo; i < hi; i++) {
        const unsigned long long t = be[i];
        be[i] = run;
        run += t;
    }
}

// value index of the keyword of length d that ends the context (Maps): one probe of the keyword -> value table (the
// record is a real match, so the key exists)
__device__ __forceinline__ uint32_t tier_value_rt(const DevTier &T, unsigned long long ctx, uint32_t cm, int d) {
    const unsigned long long key = (ctx & ((1ull << (T.b * d)) - 1ull)) | ((unsigned long long)(d - 1) << 60);
    const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
    uint32_t bucket = __umulhi(value_hash32(klo, khi), T.n_vbuckets);
    for (uint32_t tries = 0; tries < T.n_vbuckets; tries++) {
        const uint4 *q = T.vbuckets + (size_t)bucket * 2;
        const uint4 e0 = __ldg(q);
        if (e0.x == klo && e0.y == khi) return e0.z;
        const uint4 e1 = __ldg(q + 1);
        if (e1.x == klo && e1.y == khi) return e1.z;
        bucket = bucket + 1u == T.n_vbuckets ? 0u : bucket + 1u;
    }
    return kNoneD;
}

// Masks -> records.  A warp takes rows round-robin (the next row's masks and offsets are prefetched while the current
// row is expanded); a lane expands its 8 positions in order (ascending bit scan of the four mask words: position-major,
// longest keyword first) into the warp's staging window, which is flushed with coalesced 128-bit streaming stores.
// Maps re-read the row's chars to rebuild the contexts their value look-ups need.
__device__ __forceinline__ void sts_v2(uint32_t saddr, int32_t x, int32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t x) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)x) : "memory");
}
// index of the highest set bit (0xFFFFFFFF: none) and a shift that is defined for every count (>= 32 gives 0)
__device__ __forceinline__ uint32_t flo32(uint32_t x) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t shl32(uint32_t x, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
// code = row position << 4 | 16 - length  ->  (start, end); e_row = end of a keyword whose last char is row position 0
__device__ __forceinline__ int2 decode_rec(uint32_t code, int32_t e_row) {
    const int32_t e = e_row + (int32_t)(code >> 4);
    return make_int2(e - 16 + (int32_t)(code & 15u), e);
}

#ifndef ACGPU_EMIT_MIN_CTAS
#define ACGPU_EMIT_MIN_CTAS 6  // resident CTAs per SM the register allocation must allow (63 registers gave 4 = half the warps)
#endif
// The common path stages 2-byte codes (kEmitStage + 2 of them per warp); rows that do not fit take windows of kEmitWin
// 8-byte records through the same bytes.
constexpr int kEmitWin = (kEmitStage + 2) / 4;

// What a warp needs to expand rows: its staging window (and, for Maps, the value window and the packed classes of the
// row) in shared memory plus a few constants.  Shared by k_tier_emit and the emit role of k_tier_fused (kernel_fuse.cuh).
struct EmitWarp {
    int2 *s_stage;           // kEmitWin + 1 records = kEmitStage + 2 codes (+ 2 spare)
    uint32_t *s_val;         // Maps: kEmitWin values
    uint2 *s_pack;           // Maps: 34 entries, [0,1] = the 16 chars before the row
    const uint8_t *s_cls4;   // Maps: class * 4 of code units 0..255
    uint32_t stage_sa, out_par, lane_code, cm, sh;
    int b, lane;
};
constexpr size_t kEmitWarpBytesSet = sizeof(int2) * (kEmitWin + 1);
constexpr size_t kEmitWarpBytesMap = kEmitWarpBytesSet + 4 * ((kEmitWin + 3) & ~3) + 8 * 34;

// Records of one row: mm = the lane's four mask words, base = records of all rows before it.  Returns the row's count.
template <bool kIsMap>
__device__ __forceinline__ uint32_t emit_row(const DevAutomaton &A, const DevTier &T, const EmitArgs &E, const EmitWarp &W, int row, const uint4 mm,
                                             const unsigned long long base) {
    const int lane = W.lane, b = W.b;
    const uint32_t cm = W.cm, sh = W.sh;
    int2 *s_stage = W.s_stage;
    const unsigned short *s_code = reinterpret_cast<const unsigned short *>(s_stage);  // the common path stages 16-bit codes here
    uint32_t *s_val = W.s_val;
    uint2 *s_pack = W.s_pack;
    const uint32_t cnt = __popc(mm.x) + __popc(mm.y) + __popc(mm.z) + __popc(mm.w);
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (total == 0) return 0u;
    const uint32_t my_off = inc - cnt;
    const int64_t p0 = E.origin + (int64_t)row * kMaskRow + (int64_t)lane * 8;
    const int32_t e0 = (int32_t)(p0 + 1) + E.pos_base;  // end (exclusive) of a keyword whose last char is position p0
    const uint32_t words[4] = {mm.x, mm.y, mm.z, mm.w};

    Pack8 P0{0u, 0u}, P1{0u, 0u}, P2{0u, 0u};
    if (kIsMap) {
        uint32_t c4[8];
        const bool in0 = p0 >= 0 && p0 + 8 <= E.n;
        const uint4 v = ldcs_v4_if(E.hay + p0, in0);
        classify8x4(A, E.hay, E.n, p0, in0, v, W.s_cls4, c4);
        P0 = pack8(c4, sh);
        Pack8 h{0u, 0u};
        if (lane < 2) {
            const int64_t q0 = E.origin + (int64_t)row * kMaskRow - 16 + (int64_t)lane * 8;
            const bool inq = q0 >= 0 && q0 + 8 <= E.n;
            const uint4 hv = ldcs_v4_if(E.hay + q0, inq);
            uint32_t h4[8];
            classify8x4(A, E.hay, E.n, q0, inq, hv, W.s_cls4, h4);
            h = pack8(h4, sh);
        }
        Pack8 car0, car1;
        car0.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 0); car0.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 0);
        car1.hi = __shfl_sync(0xFFFFFFFFu, h.hi, 1); car1.lo = __shfl_sync(0xFFFFFFFFu, h.lo, 1);
        P1.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 1); P1.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 1);
        P2.hi = __shfl_up_sync(0xFFFFFFFFu, P0.hi, 2); P2.lo = __shfl_up_sync(0xFFFFFFFFu, P0.lo, 2);
        if (lane == 0) { P1 = car1; P2 = car0; }
        if (lane == 1) P2 = car1;
        __syncwarp();  // the previous row's readers are done with s_pack
        s_pack[lane + 2] = make_uint2(P0.hi, P0.lo);
        if (lane < 2) s_pack[lane] = make_uint2(h.hi, h.lo);
        __syncwarp();
    }
    if (total <= (uint32_t)kEmitStage && base + total <= (unsigned long long)E.cap) {
        // ---- common case: the whole row fits the staging window and the caller's buffer.  Phase 1: every lane
        //      expands its own bits into 16-bit CODES (code = index of the bit in the row's 4096-bit mask =
        //      row position << 4 | 16 - length) - 2-byte shared stores instead of 8-byte records, a third of the
        //      shared-memory wavefronts.  Phase 2 is balanced: a lane decodes TWO consecutive codes into the two
        //      records of one 16-byte streaming store.  Codes are staged at the parity of their final address so that
        //      both sides of the flush are aligned.
        const uint32_t par = ((uint32_t)base + W.out_par) & 1u;
        uint32_t sa = W.stage_sa + (my_off + par) * 2u;
        // A word holds two positions, the first one in the low half and a position's longest keyword in its lowest
        // bit: reversed, the order of the records is "highest bit first", one FLO per record.  Three predicated,
        // branch-free steps cover almost every word (0.87 records per position on configs[4]); a rare fuller word
        // finishes in a loop.  Records of a word are stored at fixed offsets from the word's first slot.
#pragma unroll
        for (int wi = 0; wi < 4; wi++) {
            uint32_t w = __brev(words[wi]);
            const uint32_t cb31 = W.lane_code + 32u * wi + 31u;  // code of the word's bit 0 = reversed bit 31
            const uint32_t pcw = (uint32_t)__popc(w);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const uint32_t f = flo32(w);  // 0xFFFFFFFF for an empty word: the store is predicated off
                if (w) sts_u16(sa + 2u * i, cb31 - f);
                w ^= shl32(1u, f);  // the bit is set (empty word: the shift gives 0)
            }
            uint32_t s2 = sa + 6u;
            while (w) {
                const uint32_t f = flo32(w);
                sts_u16(s2, cb31 - f);
                s2 += 2u;
                w ^= shl32(1u, f);  // the bit is set (empty word: the shift gives 0)
            }
            sa += 2u * pcw;
        }
        __syncwarp();
        const int32_t e_row = (int32_t)(E.origin + (int64_t)row * kMaskRow) + 1 + E.pos_base;  // end of a keyword whose last char is row position 0
        if (kIsMap) {
            // ---- values, one RECORD per lane: the code says which position and length it is; the contexts come from
            //      the row's packed classes in shared memory
            for (uint32_t r = lane; r < total; r += 32) {
                const uint32_t code = s_code[r + par];
                const uint32_t pos = code >> 4;  // 0..255
                const uint32_t own = pos >> 3;
                const uint2 a0 = s_pack[own + 2], a1 = s_pack[own + 1], a2 = s_pack[own];
                const Pack8 Q0{a0.x, a0.y}, Q1{a1.x, a1.y}, Q2{a2.x, a2.y};
                __stcs(E.val_out + base + r, tier_value_rt(T, context_of(Q0, Q1, Q2, (int)(pos & 7u), b), cm, 16 - (int)(code & 15u)));
            }
        }
        // pairs [k_lo, k_hi) are whole; a lone head record (par == 1) and a lone tail record go out as 8-byte stores
        int2 *g = E.pos_out + (base - par);  // 16-byte aligned
        const uint32_t end = par + total, k_hi = end >> 1;
        if (lane == 0 && par) __stcs(g + 1, decode_rec(s_code[1], e_row));
        if (lane == 1 && (end & 1u)) __stcs(g + (end - 1u), decode_rec(s_code[end - 1u], e_row));
        int4 *gp = reinterpret_cast<int4 *>(g) + par + lane;
        const uint32_t *sp = reinterpret_cast<const uint32_t *>(s_code) + par + lane;
        const uint32_t k0 = par + lane;
        // at most kEmitStage / 2 + 1 pairs: a fixed trip count keeps every address an immediate offset
#pragma unroll
        for (int i = 0; i < (kEmitStage / 2 + 1 + 31) / 32; i++) {
            if (par + 32u * i >= k_hi) break;  // warp-uniform
            if (k0 + 32u * i < k_hi) {
                const uint32_t cc = sp[32 * i];
                const int2 r0 = decode_rec(cc & 0xFFFFu, e_row), r1 = decode_rec(cc >> 16, e_row);
                __stcs(gp + 32 * i, make_int4(r0.x, r0.y, r1.x, r1.y));
            }
        }
        __syncwarp();
        return total;
    }

    for (uint32_t win = 0; win < total; win += kEmitWin) {
        if (cnt && my_off < win + kEmitWin && my_off + cnt > win) {
            uint32_t o = my_off - win;  // wraps below zero for records of an earlier window
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                uint32_t w = words[wi];
                while (w) {
                    const int t = __ffs(w) - 1;
                    w &= w - 1u;
                    const int j = 2 * wi + (t >> 4), d = 16 - (t & 15);
                    if (o < (uint32_t)kEmitWin) {
                        const int32_t e = e0 + j;
                        s_stage[o] = make_int2(e - d, e);
                        if (kIsMap) s_val[o] = tier_value_rt(T, context_of(P0, P1, P2, j, b), cm, d);
                    }
                    ++o;
                }
            }
        }
        __syncwarp();
        const uint32_t n_win = min((uint32_t)kEmitWin, total - win);
        const unsigned long long g0 = base + win;
        const unsigned long long room = g0 < (unsigned long long)E.cap ? (unsigned long long)E.cap - g0 : 0ull;
        const uint32_t n_out = (uint32_t)min((unsigned long long)n_win, room);
        for (uint32_t rr = lane; rr < n_out; rr += 32) {
            __stcs(&E.pos_out[g0 + rr], s_stage[rr]);
            if (kIsMap) __stcs(&E.val_out[g0 + rr], s_val[rr]);
        }
        __syncwarp();
    }
    return total;
}

__device__ __forceinline__ EmitWarp make_emit_warp(const DevTier &T, const EmitArgs &E, int2 *s_stage, uint32_t *s_val, uint2 *s_pack, const uint8_t *s_cls4,
                                                   int lane) {
    EmitWarp W;
    W.s_stage = s_stage;
    W.s_val = s_val;
    W.s_pack = s_pack;
    W.s_cls4 = s_cls4;
    W.b = T.b;
    W.cm = (1u << T.b) - 1u;
    W.sh = 1u << T.b;
    W.stage_sa = (uint32_t)__cvta_generic_to_shared(s_stage);
    // parity of the output buffer in 8-byte units: record g is 16-byte aligned iff (g + out_par) is even
    W.out_par = (uint32_t)(reinterpret_cast<uintptr_t>(E.pos_out) >> 3) & 1u;
    W.lane_code = (uint32_t)lane * 128u;
    W.lane = lane;
    return W;
}

template <bool kIsMap>
__global__ void __launch_bounds__(kEmitWarps * 32, ACGPU_EMIT_MIN_CTAS) k_tier_emit(const DevAutomaton A, const DevTier T, const EmitArgs E) {
    __shared__ __align__(16) int2 s_stage_all[kEmitWarps][kEmitWin + 1];
    __shared__ uint32_t s_val_all[kIsMap ? kEmitWarps : 1][kIsMap ? kEmitWin : 1];
    __shared__ __align__(16) uint32_t s_cls[64];
    __shared__ uint2 s_pack_all[kIsMap ? kEmitWarps : 1][kIsMap ? 34 : 1];  // packed classes of the row: [0,1] = the 16 chars before it
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *s_cls4 = reinterpret_cast<uint8_t *>(s_cls);
    if (kIsMap) {
        for (uint32_t i = tid; i < 256; i += kEmitWarps * 32) s_cls4[i] = (uint8_t)(((__ldg(&T.cls8[i >> 2]) >> ((i & 3) * 8)) & 0xFFu) * 4u);
        __syncthreads();
    }
    const EmitWarp W = make_emit_warp(T, E, s_stage_all[warp], s_val_all[kIsMap ? warp : 0], s_pack_all[kIsMap ? warp : 0], s_cls4, lane);

    // rows are taken round-robin; 32-bit row arithmetic (n_rows < 2^31) and running pointers keep the prefetch cheap
    const int n_rows = (int)E.n_rows, stride = (int)gridDim.x * kEmitWarps;
    int row = (int)blockIdx.x * kEmitWarps + warp;
    const uint4 *mp = reinterpret_cast<const uint4 *>(E.masks) + ((size_t)row * 32 + lane);
    const uint32_t *rp = E.row_excl + row;
    uint4 mm_n = make_uint4(0u, 0u, 0u, 0u);
    unsigned long long base_n = 0;
    if (row < n_rows) {
        mm_n = __ldcs(mp);
        base_n = __ldg(E.block_excl + (row >> 12)) + __ldg(rp);
    }
    static_assert(kScanRows == 4096, "row >> 12 is the scan block of a row");
    for (; row < n_rows; row += stride) {
        const uint4 mm = mm_n;
        const unsigned long long base = base_n;
        mp += (size_t)stride * 32;
        rp += stride;
        if (row + stride < n_rows) {
            mm_n = __ldcs(mp);
            base_n = __ldg(E.block_excl + ((row + stride) >> 12)) + __ldg(rp);
        }
        emit_row<kIsMap>(A, T, E, W, row, mm, base);
    }
}

}  // namespace acgpu
