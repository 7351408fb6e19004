// Longest / Shortest on narrow alphabets (generation 3): START masks from the mirrored k_tier_mask + exact chain resolution.
//
//   k_tier_mask<K, LOW, MIR = true>   one 16-bit mask per haystack position s: bit 16 - d set = a keyword of length d STARTS
//                                     at s (forward-trie tier tables; the kernel reads the haystack right to left).
//   k_sel2_map                        per tile of 8 192 positions: for every possible chain entry offset, where the chain
//                                     leaves the tile and how many matches it emits on the way (an "exit map").
//   k_sel2_group / _top / _tiles      compose the tile maps left to right (groups of 256 tiles, one thread over the group
//                                     maps, back down): every tile learns its true entry offset and the index of its
//                                     first record; also the total.
//   k_sel2_emit                       rebuilds the tile's maps, walks the true chain and writes the records at their final
//                                     offsets (ascending start = the reference's listener order).
//   k_sel2_values (Maps)              one record per thread: the value index of every record.
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
constexpr int kS2LmapStride = 17;                // halfwords per lane in the entry-window array (odd: conflict-light both ways)
constexpr int kS2Group = 256;                    // tiles per composition group
// shared words: [32 positions x 256 threads, position-major (bank = lane: conflict-free for the resolution)] followed by
// the lane entry windows: (exit, matches) of positions 0..15 of every lane as halfwords, lane-major, for the map walks
constexpr int kS2LmapWords = (kS2Threads * kS2LmapStride + 1) / 2;
constexpr int kS2SmemWords = kS2Sub * kS2Threads + kS2LmapWords;

struct Sel2Args {
    const uint32_t *masks;          // start masks, two positions per word, index space [0, n_idx)
    int64_t n_idx;                  // multiple of 256
    int64_t moff;                   // haystack position s sits at index s + moff
    int64_t n_tiles;
    uint32_t *tile_map;             // [n_tiles][16]  exit offset | matches << 8, per entry offset
    uint8_t *tile_entry;            // [n_tiles]
    unsigned long long *tile_base;  // [n_tiles] index of the tile's first record
    struct S2Status *status;        // [n_tiles] look-back records of k_sel2_fused (zeroed)
    unsigned int *tile_counter;     // ticket counter (zeroed)
    unsigned int *err;              // set when a look-back wait gave up
    int64_t n_groups;
    uint32_t *group_map;            // [n_groups][16]
    uint8_t *group_entry;           // [n_groups]
    unsigned long long *group_base; // [n_groups]
    unsigned long long *total_out;
    const uint16_t *hay;
    int64_t n;
    int32_t pos_base;
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
    // range shards of ONE haystack: the chain enters the shard's first tile at offset entry0 (the exit offset
    // of the shard before it); shard_map, when given, receives the shard's own composed map - for every entry offset
    // 0..15: exit offset (relative to the end of the last tile) | matches << 8
    uint32_t entry0;
    unsigned long long *shard_map;
};

// word of one position p of a lane's sub-tile (all relative to the sub-tile start a):
//   bits 0..3 exit offset of the chain standing at p | 4..9 matches emitted from p to the exit | 10..15 J(p) - a
//   | 16..21 start of the match emitted at p - a | 22 a match is emitted at p
__device__ __forceinline__ uint16_t *s2_lmap(uint32_t *s_w) { return reinterpret_cast<uint16_t *>(s_w + kS2Sub * kS2Threads); }

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

// Right-to-left resolution of the lane's 32 positions into s_w (see the word layout above).  Branch-free; the fields
// are kept in place (no re-packing), addresses are position * 256 + thread.
template <int MODE>
__device__ __forceinline__ void s2_resolve(const uint32_t (&mw)[16], const uint32_t (&hw)[8], uint32_t *s_w) {
    const int tid = threadIdx.x;
    uint32_t h_end = 0xFFFFu, h_start = 0;  // Shortest: the best candidate right of the sub-tile
    if (MODE == kModeShortest) {
#pragma unroll
        for (int k = 15; k >= 0; k--) {
            const uint32_t m = (hw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
            const uint32_t e = (uint32_t)(32 + k) + (uint32_t)__clz((int)m) - 15u;  // + 16 - (highest set bit)
            const bool better = m != 0u && e <= h_end;                              // ties: leftmost start
            h_end = better ? e : h_end;
            h_start = better ? (uint32_t)(32 + k) : h_start;
        }
    }
    // state: the chain standing here jumps to nx (32 = leaves the sub-tile without a match); hi = nx << 10 | start << 16 |
    // emits << 22 (the upper fields of the word), emf = 16 once a match is emitted (the "1 match" of the count field)
    uint32_t nx = 32u, hi = 32u << 10, emf = 0u;
    uint32_t b_end = 0xFFFFu, b_start = 0u;
    uint32_t *col = s_w + tid;
    uint16_t *lm = s2_lmap(s_w) + tid * kS2LmapStride;
#pragma unroll
    for (int k = 31; k >= 0; k--) {
        const uint32_t m = (mw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        const bool has = m != 0u;
        if (MODE == kModeLongest) {
            const uint32_t e = (uint32_t)k + 17u - (uint32_t)__ffs((int)m);  // + 16 - (lowest set bit)
            nx = has ? e : nx;
            hi = has ? (e << 10) + (((uint32_t)k << 16) | (1u << 22)) : hi;
        } else {
            const uint32_t e = (uint32_t)k + (uint32_t)__clz((int)m) - 15u;
            const bool better = has && e <= b_end;
            b_end = better ? e : b_end;
            b_start = better ? (uint32_t)k : b_start;
            const bool halo = h_end < b_end;
            const uint32_t cand = halo ? h_end : b_end, st = halo ? h_start : b_start;
            const bool any = b_end != 0xFFFFu;  // no candidate inside the sub-tile: skip to its end (the halo is the next lane's)
            nx = any ? cand : 32u;
            hi = any ? (cand << 10) | (st << 16) | (1u << 22) : (32u << 10);
        }
        emf = has ? 16u : emf;
        const uint32_t t = col[min(nx, 31u) * kS2Threads];
        const bool out = nx >= 32u;
        const uint32_t lo = out ? (nx - 32u) | emf : (t & 0x3FFu) + 16u;  // exit offset | matches << 4
        const uint32_t word = hi | lo;
        col[k * kS2Threads] = word;
        if (k < 16) lm[k] = (uint16_t)lo;
    }
}

// Warp maps: s_wmap[warp * 16 + o] = exit offset | matches << 8 of the warp's 1 024 positions entered at offset o.
// TRAJ: also records, for every entry offset o, where the chain enters each lane and how many matches it has emitted by
// then (s_traj[(warp * 32 + l) * 16 + o] = entry | matches << 4), so the emit kernel needs no second walk.
template <bool TRAJ>
__device__ __forceinline__ void s2_warp_maps(const uint32_t *s_w, uint32_t *s_wmap, uint16_t *s_traj) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t *lmap = s2_lmap(const_cast<uint32_t *>(s_w));
    __syncwarp();
    if (lane < kS2Ent) {
        uint32_t cur = lane, cnt = 0;
#pragma unroll 4
        for (int l = 0; l < 32; l++) {
            if (TRAJ) s_traj[(warp * 32 + l) * kS2Ent + lane] = (uint16_t)(cur | (cnt << 4));
            const uint32_t t = lmap[(warp * 32 + l) * kS2LmapStride + cur];
            cur = t & 15u;
            cnt += t >> 4;
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
        s2_warp_maps<false>(s_w, s_wmap, nullptr);
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

// ---- composing the tile maps: groups of 256 tiles (one warp each), then the group maps (one thread), then back down.
// A map row is 16 words: exit offset | matches << 8 for entry offsets 0..15.

// one warp: the rows of 32 consecutive maps staged in shared memory
__device__ __forceinline__ void s2_stage_rows(const uint32_t *rows, int64_t first, int64_t n_rows, uint32_t *s_rows) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int64_t w = first * kS2Ent + k * 32 + lane;
        s_rows[k * 32 + lane] = w < n_rows * kS2Ent ? __ldg(rows + w) : ((uint32_t)lane & 15u);  // past the end: identity, no matches
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) k_sel2_group(const Sel2Args P) {
    __shared__ uint32_t s_rows[32 * kS2Ent];
    const int lane = threadIdx.x;
    const int64_t g = blockIdx.x;
    uint32_t cur = lane & 15u, cnt = 0;
    for (int c = 0; c < kS2Group / 32; c++) {
        const int64_t first = g * kS2Group + c * 32;
        if (first >= P.n_tiles) break;
        __syncwarp();
        s2_stage_rows(P.tile_map, first, P.n_tiles, s_rows);
#pragma unroll 8
        for (int l = 0; l < 32; l++) {
            const uint32_t t = s_rows[l * kS2Ent + cur];
            cur = t & 0xFFu;
            cnt += t >> 8;
        }
    }
    if (lane < kS2Ent) P.group_map[g * kS2Ent + lane] = cur | (cnt << 8);
}

// one thread follows the true chain over the group maps (at most 1 024 groups for a 2^31-char haystack)
__global__ void __launch_bounds__(1024, 1) k_sel2_top(const Sel2Args P) {
    extern __shared__ __align__(16) uint32_t s_gm[];
    for (int64_t i = threadIdx.x; i < P.n_groups * kS2Ent; i += blockDim.x) s_gm[i] = P.group_map[i];
    __syncthreads();
    if (P.shard_map && threadIdx.x >= 32 && threadIdx.x < 32 + kS2Ent) {
        uint32_t cur = threadIdx.x - 32u;
        unsigned long long acc = 0;
        for (int64_t g = 0; g < P.n_groups; g++) {
            const uint32_t t = s_gm[g * kS2Ent + cur];
            cur = t & 0xFFu;
            acc += t >> 8;
        }
        P.shard_map[threadIdx.x - 32] = (unsigned long long)cur | (acc << 8);
    }
    if (threadIdx.x == 0) {
        uint32_t cur = P.entry0;  // one-shot matches: index 0 (entries before haystack position 0 are empty)
        unsigned long long acc = 0;
        for (int64_t g = 0; g < P.n_groups; g++) {
            P.group_entry[g] = (uint8_t)cur;
            P.group_base[g] = acc;
            const uint32_t t = s_gm[g * kS2Ent + cur];
            cur = t & 0xFFu;
            acc += t >> 8;
        }
        *P.total_out = acc;
    }
}

// zero the start masks of index positions [from, to): positions a chain entering further right must not see
__global__ void k_sel2_zero_prefix(uint32_t *masks, int64_t from, int64_t to) {
    unsigned short *m16 = reinterpret_cast<unsigned short *>(masks);
    for (int64_t i = from + (int64_t)(blockIdx.x * blockDim.x + threadIdx.x); i < to; i += (int64_t)gridDim.x * blockDim.x) m16[i] = 0;
}

// one warp per group: every tile learns its entry offset and the index of its first record
__global__ void __launch_bounds__(32) k_sel2_tiles(const Sel2Args P) {
    __shared__ uint32_t s_rows[32 * kS2Ent];
    __shared__ uint8_t s_ent[32];
    __shared__ unsigned long long s_bas[32];
    const int lane = threadIdx.x;
    const int64_t g = blockIdx.x;
    uint32_t cur = P.group_entry[g];
    unsigned long long acc = P.group_base[g];
    for (int c = 0; c < kS2Group / 32; c++) {
        const int64_t first = g * kS2Group + c * 32;
        if (first >= P.n_tiles) break;
        __syncwarp();
        s2_stage_rows(P.tile_map, first, P.n_tiles, s_rows);
        if (lane == 0) {
            for (int l = 0; l < 32; l++) {
                s_ent[l] = (uint8_t)cur;
                s_bas[l] = acc;
                const uint32_t t = s_rows[l * kS2Ent + cur];
                cur = t & 0xFFu;
                acc += t >> 8;
            }
        }
        cur = __shfl_sync(0xFFFFFFFFu, cur, 0);
        acc = __shfl_sync(0xFFFFFFFFu, acc, 0);
        __syncwarp();
        if (first + lane < P.n_tiles) {
            P.tile_entry[first + lane] = s_ent[lane];
            P.tile_base[first + lane] = s_bas[lane];
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kS2Threads, 4) k_sel2_emit(const Sel2Args P) {
    extern __shared__ __align__(16) uint32_t s_w[];
    __shared__ uint32_t s_wmap[(kS2Threads / 32) * kS2Ent];
    __shared__ uint32_t s_wentry[kS2Threads / 32];
    __shared__ unsigned long long s_wbase[kS2Threads / 32];
    __shared__ uint16_t s_traj[kS2Threads * kS2Ent];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        uint32_t mw[16], hw[8];
        s2_load(P, tile, mw, hw);
        const int64_t a = tile * kS2Tile + (int64_t)tid * kS2Sub - P.moff;  // haystack position of the lane's first index
        __syncthreads();  // the previous tile's readers are done
        s2_resolve<MODE>(mw, hw, s_w);
        s2_warp_maps<true>(s_w, s_wmap, s_traj);
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
        const uint32_t eo = s_traj[tid * kS2Ent + s_wentry[warp]];
        unsigned long long idx = s_wbase[warp] + (eo >> 4);
        uint32_t p = eo & 15u;
        const int32_t a32 = (int32_t)a + P.pos_base;
        while (p < 32u) {
            const uint32_t w = s_w[p * kS2Threads + tid];
            if ((w >> 22) & 1u) {
                if (idx < (unsigned long long)P.cap)
                    __stcs(&P.pos_out[idx], make_int2(a32 + (int32_t)((w >> 16) & 63u), a32 + (int32_t)((w >> 10) & 63u)));
                ++idx;
            }
            p = (w >> 10) & 63u;
        }
    }
}

// ---- single pass: exit maps, decoupled look-back over the MAPS of the preceding tiles, emission.
//
// Tiles are taken in ticket order.  A tile publishes its 16-entry map (flag 1) as soon as it has it, then looks back for the
// nearest predecessor whose chain state is resolved (flag 2 = "the chain enters the NEXT tile at offset e with r records
// before it"), follows the chain through the maps of the tiles in between, publishes its own resolved state and emits.
// Nothing is speculative; a tile only ever waits for tiles that were started before it.
struct __align__(16) S2Status {   // 64 bytes per tile
    uint32_t flag;                // 0 = nothing yet, 1 = map published, 2 = map and resolved state published
    uint32_t next_entry;          // chain offset on entry to the next tile
    unsigned long long next_base; // records emitted before the next tile
    unsigned long long exits;     // the map: exit offset (4 bits) per entry offset
    uint32_t pad[2];
    uint16_t counts[16];          //          matches per entry offset (16-byte aligned: read as two uint4)
};
static_assert(sizeof(S2Status) == 64, "one look-back record per 64-byte line");

__device__ __forceinline__ uint32_t s2_ld_flag(const S2Status *st) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->flag) : "memory");
    return v;
}
__device__ __forceinline__ void s2_st_flag(S2Status *st, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&st->flag), "r"(v) : "memory");
}

constexpr uint32_t kS2SpinLimit = 1u << 26;  // a tile that waits this long reports an error instead of hanging the device

// Called by warp 0 after the tile map is in s_tmap[16] (exit | matches << 8).  Returns the tile's entry offset and record
// base to every lane of the warp.
__device__ __forceinline__ void s2_lookback(S2Status *status, int64_t tile, const uint32_t *s_tmap, unsigned long long *s_ex,
                                            uint16_t *s_cn, unsigned int *err, uint32_t &entry_out, unsigned long long &base_out) {
    const int lane = threadIdx.x & 31;
    S2Status *mine = status + tile;
    // ---- publish the map
    {
        const uint32_t t = lane < kS2Ent ? s_tmap[lane] : 0u;
        const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 8 ? (t & 15u) << (4 * lane) : 0u);
        const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, (lane >= 8 && lane < 16) ? (t & 15u) << (4 * (lane - 8)) : 0u);
        if (lane < kS2Ent) mine->counts[lane] = (uint16_t)(t >> 8);
        if (lane == 0) mine->exits = ((unsigned long long)hi << 32) | lo;
        __threadfence();
        __syncwarp();
        if (lane == 0 && tile != 0) s2_st_flag(mine, 1u);
    }
    uint32_t cur = 0;
    unsigned long long base = 0;
    if (tile != 0) {
        // ---- find the nearest resolved predecessor J (every tile between J and this one has published its map by then)
        int64_t J = -1;
        for (int64_t look = tile - 1; look >= 0 && J < 0; look -= 32) {
            const int64_t idx = look - lane;
            uint32_t f = 2u;  // lanes before tile 0 never stop the search on their own (masked below)
            if (idx >= 0) {
                uint32_t spins = 0;
                while ((f = s2_ld_flag(status + idx)) == 0u) {
                    __nanosleep(40);
                    if (++spins > kS2SpinLimit) { atomicExch(err, 1u); f = 2u; break; }
                }
            }
            const uint32_t res = __ballot_sync(0xFFFFFFFFu, idx >= 0 && f == 2u);
            if (res) J = look - (__ffs(res) - 1);
        }
        __threadfence();
        // ---- the chain state after tile J (tile -1 = the start of the haystack: offset 0, no records)
        if (J >= 0) {
            cur = __ldcg(&status[J].next_entry);
            base = __ldcg(&status[J].next_base);
        }
        // ---- follow the chain through the maps of tiles J+1 .. tile-1, 32 at a time
        for (int64_t first = J + 1; first < tile; first += 32) {
            const int64_t idx = first + lane;
            __syncwarp();
            if (idx < tile) {
                s_ex[lane] = __ldcg(&status[idx].exits);
                const uint4 c0 = __ldcg(reinterpret_cast<const uint4 *>(status[idx].counts));
                const uint4 c1 = __ldcg(reinterpret_cast<const uint4 *>(status[idx].counts) + 1);
                reinterpret_cast<uint4 *>(s_cn + lane * 16)[0] = c0;
                reinterpret_cast<uint4 *>(s_cn + lane * 16)[1] = c1;
            }
            __syncwarp();
            const int cnt = (int)min((int64_t)32, tile - first);
            for (int i = 0; i < cnt; i++) {
                base += s_cn[i * 16 + cur];
                cur = (uint32_t)(s_ex[i] >> (4 * cur)) & 15u;
            }
        }
    }
    entry_out = cur;
    base_out = base;
    // ---- publish the resolved state
    if (lane == 0) {
        const uint32_t t = s_tmap[cur];
        mine->next_entry = t & 0xFFu;
        mine->next_base = base + (t >> 8);
        __threadfence();
        s2_st_flag(mine, 2u);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kS2Threads, 4) k_sel2_fused(const Sel2Args P) {
    extern __shared__ __align__(16) uint32_t s_w[];
    __shared__ uint32_t s_wmap[(kS2Threads / 32) * kS2Ent];
    __shared__ uint32_t s_tmap[kS2Ent];
    __shared__ uint32_t s_wentry[kS2Threads / 32];
    __shared__ unsigned long long s_wbase[kS2Threads / 32];
    __shared__ uint16_t s_traj[kS2Threads * kS2Ent];
    __shared__ __align__(16) unsigned long long s_ex[32];
    __shared__ __align__(16) uint16_t s_cn[32 * 16];
    __shared__ long long s_tile;
    const int tid = threadIdx.x, warp = tid >> 5;
    while (true) {
        __syncthreads();  // the previous tile's readers are done
        if (tid == 0) s_tile = (long long)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        uint32_t mw[16], hw[8];
        s2_load(P, tile, mw, hw);
        const int64_t a = tile * kS2Tile + (int64_t)tid * kS2Sub - P.moff;  // haystack position of the lane's first index
        s2_resolve<MODE>(mw, hw, s_w);
        s2_warp_maps<true>(s_w, s_wmap, s_traj);
        __syncthreads();
        if (tid < kS2Ent) {
            uint32_t cur = tid, cnt = 0;
#pragma unroll
            for (int w = 0; w < kS2Threads / 32; w++) {
                const uint32_t t = s_wmap[w * kS2Ent + cur];
                cur = t & 0xFFu;
                cnt += t >> 8;
            }
            s_tmap[tid] = cur | (cnt << 8);
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t cur;
            unsigned long long acc;
            s2_lookback(P.status, tile, s_tmap, s_ex, s_cn, P.err, cur, acc);
            if (tid == 0) {
                if (tile == P.n_tiles - 1) *P.total_out = acc + (s_tmap[cur] >> 8);
#pragma unroll
                for (int w = 0; w < kS2Threads / 32; w++) {
                    s_wentry[w] = cur;
                    s_wbase[w] = acc;
                    const uint32_t t = s_wmap[w * kS2Ent + cur];
                    cur = t & 0xFFu;
                    acc += t >> 8;
                }
            }
        }
        __syncthreads();
        const uint32_t eo = s_traj[tid * kS2Ent + s_wentry[warp]];
        unsigned long long idx = s_wbase[warp] + (eo >> 4);
        uint32_t p = eo & 15u;
        const int32_t a32 = (int32_t)a + P.pos_base;
        while (p < 32u) {
            const uint32_t w = s_w[p * kS2Threads + tid];
            if ((w >> 22) & 1u) {
                if (idx < (unsigned long long)P.cap)
                    __stcs(&P.pos_out[idx], make_int2(a32 + (int32_t)((w >> 16) & 63u), a32 + (int32_t)((w >> 10) & 63u)));
                ++idx;
            }
            p = (w >> 10) & 63u;
        }
    }
}

// ---- Maps.  The records of a tile are consecutive (tile_base), so one CTA per tile packs the classes of the tile's
// 8 192 + 32 positions into a bit stream in shared memory (b bits per position, every position classified once) and then
// takes one record per thread: the keyword's context (its packed classes, first char lowest - the order the forward-trie
// tables use) is three word loads and two funnel shifts, the value one probe of the tier tables.
template <int B>
__device__ __forceinline__ void s2_pack_classes(const DevAutomaton &A, const uint16_t *hay, int64_t n, int64_t a,
                                                const uint8_t *s_cls, uint32_t *dst) {
    // 32 positions starting at haystack position a (a 16-byte aligned address when inside the haystack) -> B words at dst;
    // B is a compile-time constant so every class lands with one shift-or at a fixed place
    uint32_t out[B];
#pragma unroll
    for (int k = 0; k < B; k++) out[k] = 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int64_t p0 = a + q * 8;
        uint32_t ch[8];
        if (p0 >= 0 && p0 + 8 <= n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(hay + p0));
            ch[0] = v.x & 0xFFFFu; ch[1] = v.x >> 16; ch[2] = v.y & 0xFFFFu; ch[3] = v.y >> 16;
            ch[4] = v.z & 0xFFFFu; ch[5] = v.z >> 16; ch[6] = v.w & 0xFFFFu; ch[7] = v.w >> 16;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) ch[j] = (p0 + j >= 0 && p0 + j < n) ? (uint32_t)__ldg(&hay[p0 + j]) : 0x10000u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t c = ch[j] < 256u ? (uint32_t)s_cls[ch[j]] : (ch[j] < 0x10000u ? (uint32_t)__ldg(&A.cls[ch[j]]) : 0u);
            const int bit = (q * 8 + j) * B;
            out[bit >> 5] |= c << (bit & 31);
            if ((bit & 31) + B > 32) out[(bit >> 5) + 1] |= c >> (32 - (bit & 31));
        }
    }
#pragma unroll
    for (int k = 0; k < B; k++) dst[k] = out[k];
}

template <int B>
__global__ void __launch_bounds__(kS2Threads) k_sel2_values(const DevAutomaton A, const DevTier T, const Sel2Args P) {
    __shared__ uint32_t s_bits[(kS2Threads + 1) * B + 3];
    __shared__ uint8_t s_cls[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += kS2Threads) s_cls[i] = (uint8_t)((__ldg(&T.cls8[i >> 2]) >> ((i & 3) * 8)) & 0xFFu);
    constexpr int b = B;
    const uint32_t cm = (1u << b) - 1u;
    const unsigned long long n_rec = min(*P.total_out, (unsigned long long)P.cap);
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        // records of tile t: [base after tile t-1, base after tile t)
        unsigned long long r0, r1;
        if (P.status) {
            r0 = tile ? min(P.status[tile - 1].next_base, n_rec) : 0ull;
            r1 = min(P.status[tile].next_base, n_rec);
        } else {
            r0 = min(P.tile_base[tile], n_rec);
            r1 = tile + 1 < P.n_tiles ? min(P.tile_base[tile + 1], n_rec) : n_rec;
        }
        if (r0 >= r1) continue;  // uniform over the block
        const int64_t a0 = tile * kS2Tile - P.moff;  // haystack position of the tile's first index
        __syncthreads();  // the previous tile's readers are done (and s_cls is written)
        s2_pack_classes<B>(A, P.hay, P.n, a0 + (int64_t)tid * kS2Sub, s_cls, s_bits + tid * b);
        if (tid == 0) s2_pack_classes<B>(A, P.hay, P.n, a0 + kS2Tile, s_cls, s_bits + kS2Threads * b);
        __syncthreads();
        const int32_t a32 = (int32_t)a0 + P.pos_base;
        for (unsigned long long r = r0 + tid; r < r1; r += kS2Threads) {
            const int2 rec = __ldcs(&P.pos_out[r]);
            const uint32_t d = (uint32_t)(rec.y - rec.x);
            const uint32_t bit = (uint32_t)(rec.x - a32) * (uint32_t)b;
            const uint32_t *q = s_bits + (bit >> 5);
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
            const uint32_t lo = __funnelshift_r(w0, w1, bit & 31u), hi = __funnelshift_r(w1, w2, bit & 31u);
            const unsigned long long ctx = (((unsigned long long)hi << 32) | lo) & ((1ull << (d * (uint32_t)b)) - 1ull);
            __stcs(&P.val_out[r], tier_value_rt(T, ctx, cm, (int)d));
        }
    }
}

}  // namespace acgpu
