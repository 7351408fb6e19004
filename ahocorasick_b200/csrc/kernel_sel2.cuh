// Longest / Shortest on narrow alphabets (generation 3): START masks from the mirrored k_tier_mask + exact chain resolution.
//
//   k_tier_mask<K, LOW, MIR = true>   one 16-bit mask per haystack position s: bit 16 - d set = a keyword of length d STARTS
//                                     at s (forward-trie tier tables; the kernel reads the haystack right to left).
//   k_sel2_map                        per tile of 8 192 positions: for every possible chain entry offset, where the chain
//                                     leaves the tile and how many matches it emits on the way (an "exit map").
//   k_sel2_scan                       composes the tile maps left to right: every tile learns its true entry offset and the
//                                     index of its first record; also the total.
//   k_sel2_emit<isMap>                rebuilds the tile's maps, walks the true chain and writes the records at their final
//                                     offsets (ascending start = the reference's listener order).
//
// The sequential selection of the reference (LongestMatchSet.java:192-265 + SetMatchQueue.java:45-95;
// ShortestMatchSet.java:182-260) is a chain  pos -> J(pos):
//   Longest   s = first start >= pos with a keyword, e = s + longest keyword at s;        emit (s, e), pos = e
//   Shortest  (e, s) = min over starts s >= pos of (s + shortest keyword at s, s);        emit (s, e), pos = e
// J(pos) only looks at most 16 + 15 positions to the right of pos, so a lane resolves 32 consecutive positions right to
// left in one pass (dynamic programming: exit(p) = exit(J(p))), lanes / warps / tiles compose their 16-entry maps, and no
// step of the resolution is speculative: the record stream is bit-identical to the sequential loop.
//
// All positions here are INDEX-space positions i = s + moff (the mask array of the mirrored kernel starts moff entries
// before haystack position 0; those entries and the ones past the end of the haystack are zero).
#pragma once
#include "kernel_tier.cuh"
#include "kernel_emit.cuh"
#include "kernels.cuh"

namespace acgpu {

constexpr int kS2Threads = 256;
constexpr int kS2Sub = 32;                       // positions per lane
constexpr int kS2Tile = kS2Threads * kS2Sub;     // positions per CTA tile
constexpr int kS2Ent = 16;                       // entry offsets of a map (exit offsets are < max_len <= 16)
constexpr int kS2LmapStride = 17;                // words per lane in the entry-window array (conflict-free both ways)
constexpr int kS2ScanThreads = 512;
// shared words: [lane entry windows: positions 0..15 of every lane][positions 16..31, position-major]
constexpr int kS2LmapWords = kS2Threads * kS2LmapStride;
constexpr int kS2SmemWords = kS2LmapWords + 16 * kS2Threads;

struct Sel2Args {
    const uint32_t *masks;          // start masks, two positions per word, index space [0, n_idx)
    int64_t n_idx;                  // multiple of 256
    int64_t moff;                   // haystack position s sits at index s + moff
    int64_t n_tiles;
    uint32_t *tile_map;             // [n_tiles][16]  exit offset | matches << 8, per entry offset
    uint8_t *tile_entry;            // [n_tiles]
    unsigned long long *tile_base;  // [n_tiles] index of the tile's first record
    unsigned long long *total_out;
    const uint16_t *hay;
    int64_t n;
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
};

// word of one position p of a lane's sub-tile (all relative to the sub-tile start a):
//   bits 0..3 exit offset of the chain standing at p | 4..9 matches emitted from p to the exit | 10..15 J(p) - a
//   | 16..21 start of the match emitted at p - a | 22 a match is emitted at p
__device__ __forceinline__ uint32_t s2_idx(int k, int tid) {
    return k < 16 ? (uint32_t)(tid * kS2LmapStride + k) : (uint32_t)(kS2LmapWords + (k - 16) * kS2Threads + tid);
}

// masks of the lane's 32 positions (mw) and of the 16 positions right of them (hw)
__device__ __forceinline__ void s2_load(const Sel2Args &P, int64_t tile, uint32_t (&mw)[16], uint32_t (&hw)[8]) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t w0 = (tile * kS2Tile + (int64_t)tid * kS2Sub) >> 1;  // first word of the lane
    const int64_t n_words = P.n_idx >> 1;
    const uint4 *src = reinterpret_cast<const uint4 *>(P.masks + w0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (w0 + 4 * k < n_words) v = __ldg(src + k);
        mw[4 * k] = v.x; mw[4 * k + 1] = v.y; mw[4 * k + 2] = v.z; mw[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) hw[k] = __shfl_down_sync(0xFFFFFFFFu, mw[k], 1);
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (w0 + 16 + 4 * k < n_words) v = __ldg(src + 4 + k);
            hw[4 * k] = v.x; hw[4 * k + 1] = v.y; hw[4 * k + 2] = v.z; hw[4 * k + 3] = v.w;
        }
    }
}

// Right-to-left resolution of the lane's 32 positions into s_w (see the word layout above).
template <int MODE>
__device__ __forceinline__ void s2_resolve(const uint32_t (&mw)[16], const uint32_t (&hw)[8], uint32_t *s_w) {
    const int tid = threadIdx.x;
    uint32_t h_end = 0xFFFFu, h_start = 0;  // Shortest: the best candidate right of the sub-tile
    if (MODE == kModeShortest) {
#pragma unroll
        for (int k = 15; k >= 0; k--) {
            const uint32_t m = (hw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
            if (m) {
                const uint32_t e = (uint32_t)(32 + k) + (uint32_t)__clz((int)m) - 15u;  // + 16 - (highest set bit)
                if (e <= h_end) { h_end = e; h_start = 32 + k; }                        // ties: leftmost start
            }
        }
    }
    uint32_t word = 32u << 10;  // nothing to the right: skip to the end of the sub-tile, no match
    uint32_t b_end = 0xFFFFu, b_start = 0;
#pragma unroll
    for (int k = 31; k >= 0; k--) {
        const uint32_t m = (mw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        if (m) {
            uint32_t nx, st;
            if (MODE == kModeLongest) {
                nx = (uint32_t)k + 17u - (uint32_t)__ffs((int)m);  // + 16 - (lowest set bit)
                st = k;
            } else {
                const uint32_t e = (uint32_t)k + (uint32_t)__clz((int)m) - 15u;
                if (e <= b_end) { b_end = e; b_start = k; }
                const bool halo = h_end < b_end;
                nx = halo ? h_end : b_end;
                st = halo ? h_start : b_start;
            }
            uint32_t x, c;
            if (nx >= 32u) {
                x = nx - 32u;
                c = 1u;
            } else {
                const uint32_t t = s_w[s2_idx((int)nx, tid)];
                x = t & 15u;
                c = ((t >> 4) & 63u) + 1u;
            }
            word = x | (c << 4) | (nx << 10) | (st << 16) | (1u << 22);
        }
        s_w[s2_idx(k, tid)] = word;
    }
}

// s2_idx with a run-time position
__device__ __forceinline__ uint32_t s2_idx_rt(uint32_t k, int tid) {
    return k < 16u ? (uint32_t)tid * kS2LmapStride + k : (uint32_t)kS2LmapWords + (k - 16u) * kS2Threads + (uint32_t)tid;
}

// Warp maps: s_wmap[warp * 16 + o] = exit offset | matches << 8 of the warp's 1 024 positions entered at offset o.
__device__ __forceinline__ void s2_warp_maps(const uint32_t *s_w, uint32_t *s_wmap) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncwarp();
    if (lane < kS2Ent) {
        uint32_t cur = lane, cnt = 0;
#pragma unroll 4
        for (int l = 0; l < 32; l++) {
            const uint32_t t = s_w[(warp * 32 + l) * kS2LmapStride + cur];
            cur = t & 15u;
            cnt += (t >> 4) & 63u;
        }
        s_wmap[warp * kS2Ent + lane] = cur | (cnt << 8);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kS2Threads, 4) k_sel2_map(const Sel2Args P) {
    extern __shared__ __align__(16) uint32_t s_w[];
    __shared__ uint32_t s_wmap[(kS2Threads / 32) * kS2Ent];
    const int tid = threadIdx.x;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        uint32_t mw[16], hw[8];
        s2_load(P, tile, mw, hw);
        __syncthreads();  // the previous tile's readers are done
        s2_resolve<MODE>(mw, hw, s_w);
        s2_warp_maps(s_w, s_wmap);
        __syncthreads();
        if (tid < kS2Ent) {
            uint32_t cur = tid, cnt = 0;
#pragma unroll
            for (int w = 0; w < kS2Threads / 32; w++) {
                const uint32_t t = s_wmap[w * kS2Ent + cur];
                cur = t & 0xFFu;
                cnt += t >> 8;
            }
            P.tile_map[tile * kS2Ent + tid] = cur | (cnt << 8);
        }
    }
}

// One block: thread j composes the maps of a contiguous slice of tiles for all 16 entry offsets, thread 0 then follows
// the one true chain over the 512 slice maps, and every thread hands its tiles their entry offset and record base.
__global__ void __launch_bounds__(kS2ScanThreads, 1) k_sel2_scan(const Sel2Args P) {
    __shared__ uint8_t s_exit[kS2ScanThreads][kS2Ent];
    __shared__ uint32_t s_cnt[kS2ScanThreads][kS2Ent + 1];
    __shared__ uint8_t s_in[kS2ScanThreads];
    __shared__ unsigned long long s_base[kS2ScanThreads];
    const int tid = threadIdx.x;
    const int64_t per = (P.n_tiles + kS2ScanThreads - 1) / kS2ScanThreads;
    const int64_t lo = min(P.n_tiles, (int64_t)tid * per), hi = min(P.n_tiles, lo + per);
    {
        uint32_t cur[kS2Ent], cnt[kS2Ent];
#pragma unroll
        for (int o = 0; o < kS2Ent; o++) { cur[o] = o; cnt[o] = 0; }
        for (int64_t t = lo; t < hi; t++) {
            const uint32_t *row = P.tile_map + t * kS2Ent;
#pragma unroll
            for (int o = 0; o < kS2Ent; o++) {
                const uint32_t w = __ldg(row + cur[o]);
                cur[o] = w & 0xFFu;
                cnt[o] += w >> 8;
            }
        }
#pragma unroll
        for (int o = 0; o < kS2Ent; o++) { s_exit[tid][o] = (uint8_t)cur[o]; s_cnt[tid][o] = cnt[o]; }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t cur = 0;  // the chain starts at index 0 (entries before haystack position 0 are empty)
        unsigned long long acc = 0;
        for (int j = 0; j < kS2ScanThreads; j++) {
            s_in[j] = (uint8_t)cur;
            s_base[j] = acc;
            acc += s_cnt[j][cur];
            cur = s_exit[j][cur];
        }
        *P.total_out = acc;
    }
    __syncthreads();
    uint32_t cur = s_in[tid];
    unsigned long long acc = s_base[tid];
    for (int64_t t = lo; t < hi; t++) {
        P.tile_entry[t] = (uint8_t)cur;
        P.tile_base[t] = acc;
        const uint32_t w = __ldg(P.tile_map + t * kS2Ent + cur);
        cur = w & 0xFFu;
        acc += w >> 8;
    }
}

template <int MODE, bool kIsMap>
__global__ void __launch_bounds__(kS2Threads, 4) k_sel2_emit(const DevAutomaton A, const DevTier T, const Sel2Args P) {
    extern __shared__ __align__(16) uint32_t s_w[];
    __shared__ uint32_t s_wmap[(kS2Threads / 32) * kS2Ent];
    __shared__ uint32_t s_wentry[kS2Threads / 32];
    __shared__ unsigned long long s_wbase[kS2Threads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = T.b;
    const uint32_t cm = (1u << b) - 1u;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        uint32_t mw[16], hw[8];
        s2_load(P, tile, mw, hw);
        __syncthreads();  // the previous tile's readers are done
        s2_resolve<MODE>(mw, hw, s_w);
        s2_warp_maps(s_w, s_wmap);
        __syncthreads();
        if (tid == 0) {
            uint32_t cur = P.tile_entry[tile];
            unsigned long long acc = P.tile_base[tile];
#pragma unroll
            for (int w = 0; w < kS2Threads / 32; w++) {
                s_wentry[w] = cur;
                s_wbase[w] = acc;
                const uint32_t t = s_wmap[w * kS2Ent + cur];
                cur = t & 0xFFu;
                acc += t >> 8;
            }
        }
        __syncthreads();
        // lane entries: lane 0 follows the warp's chain through the 32 lane maps
        uint32_t my_entry = 0, my_off = 0;
        {
            uint32_t cur = s_wentry[warp], acc = 0;
            if (lane == 0) {
                for (int l = 0; l < 32; l++) {
                    const uint32_t t = s_w[(warp * 32 + l) * kS2LmapStride + cur];
                    // park (entry, offset) of lane l in the unused 17th word of its entry window
                    s_w[(warp * 32 + l) * kS2LmapStride + 16] = cur | (acc << 8);
                    cur = t & 15u;
                    acc += (t >> 4) & 63u;
                }
            }
            __syncwarp();
            const uint32_t eo = s_w[tid * kS2LmapStride + 16];
            my_entry = eo & 0xFFu;
            my_off = eo >> 8;
        }
        unsigned long long idx = s_wbase[warp] + my_off;
        const int64_t a = tile * kS2Tile + (int64_t)tid * kS2Sub - P.moff;  // haystack position of the lane's first index
        uint32_t p = my_entry;
        while (p < 32u) {
            const uint32_t w = s_w[s2_idx_rt(p, tid)];
            if ((w >> 22) & 1u) {
                if (idx < (unsigned long long)P.cap) {
                    const int64_t st = a + ((w >> 16) & 63u), en = a + ((w >> 10) & 63u);
                    __stcs(&P.pos_out[idx], make_int2((int32_t)st + P.pos_base, (int32_t)en + P.pos_base));
                    if (kIsMap) {
                        unsigned long long ctx = 0;
                        const int d = (int)(en - st);
                        for (int i = 0; i < d; i++)
                            ctx |= (unsigned long long)__ldg(&A.cls[__ldg(&P.hay[st + i])]) << (b * i);
                        __stcs(&P.val_out[idx], tier_value_rt(T, ctx, cm, d));
                    }
                }
                ++idx;
            }
            p = (w >> 10) & 63u;
        }
    }
}

}  // namespace acgpu
