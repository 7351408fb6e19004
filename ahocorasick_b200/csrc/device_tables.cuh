// Device-side view of a flattened dictionary + small shared helpers for the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace acgpu {

constexpr uint32_t kNoneD = 0xFFFFFFFFu;
constexpr uint32_t kTerm = 1u;
constexpr uint32_t kKids = 2u;

struct DevAutomaton {
    const uint16_t *cls;       // [65536] code unit -> class (case folding folded in)
    const uint32_t *wordbits;  // [2048] raw word-char bitmap (WholeWord), else nullptr
    const uint32_t *wordbits_fold;  // [2048] bit c = wordChars[toLowerCase(c)] (literal WholeWord matchers, quirk Q7), else nullptr
    const uint2 *root;         // [n_classes] {child, info}
    const uint4 *edges;        // open addressing, {parent, cls, child, info}; empty: parent == kNoneD
    const uint32_t *node_value;  // [n_nodes]
    uint32_t edge_mask;
    int32_t max_len;
    int32_t n_classes;
    int32_t has_other;
    int32_t family;
    int32_t is_map;
};

// arguments of the AhoCorasick-family scan kernels (k_ac_scan, k_ac_tier)
struct AcArgs {
    const uint16_t *hay;   // haystack window (device)
    int64_t n;             // chars in the window
    int64_t emit_from;     // report matches whose last char index q is in [emit_from, emit_to)
    int64_t emit_to;
    int64_t origin;        // k_ac_tier: position of row 0 (<= emit_from; makes every lane's 128-bit load aligned)
    int32_t pos_base;      // added to reported positions (stream offset; wraps like a Java int)
    int2 *pos_out;
    uint32_t *val_out;
    int64_t cap;
    unsigned long long *total_out;
    unsigned int *tile_counter;
    unsigned long long *status;
    int64_t n_tiles;
};

__host__ __device__ __forceinline__ uint32_t edge_hash_d(uint32_t parent, uint32_t c) {
    uint32_t h = parent * 0x9E3779B1u ^ (c * 0x85EBCA6Bu + 0x7F4A7C15u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

// One trie step: (node, class) -> child.  Root uses the direct table, deeper nodes the hashed edge table.
__device__ __forceinline__ bool trie_step(const DevAutomaton &A, uint32_t &node, uint32_t c, uint32_t &info) {
    if (node == 0) {
        uint2 r = __ldg(&A.root[c]);
        if (r.x == kNoneD) return false;
        node = r.x;
        info = r.y;
        return true;
    }
    uint32_t i = edge_hash_d(node, c) & A.edge_mask;
    while (true) {
        uint4 e = __ldg(&A.edges[i]);
        if (e.x == node && e.y == c) {
            node = e.z;
            info = e.w;
            return true;
        }
        if (e.x == kNoneD) return false;
        i = (i + 1) & A.edge_mask;
    }
}

// The same step for walks that carry `info` along (it holds the current node's info word on entry; 0 at the root): bits 2 .. 31
// of an info word say which classes (mod 30) have a child, so a step that cannot succeed costs no probe of the edge table -
// an unsuccessful open-addressing look-up walks ~2.5 slots, and in running text most walks END with one.
__device__ __forceinline__ bool trie_step_sig(const DevAutomaton &A, uint32_t &node, uint32_t c, uint32_t &info) {
    if (node != 0u && !((info >> (2u + c % 30u)) & 1u)) return false;
    return trie_step(A, node, c, info);
}

__device__ __forceinline__ bool is_word_char(const DevAutomaton &A, uint32_t raw) {
    return (__ldg(&A.wordbits[raw >> 5]) >> (raw & 31)) & 1u;
}

// ---- decoupled look-back (single-pass ordered compaction across tiles) ------------------------------
// status word: bits 63..62 = flag (0 invalid, 1 tile aggregate, 2 inclusive prefix), bits 61..0 = value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagInc = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Step 1 (one lane): make this tile's aggregate visible.  Tile 0 has no predecessors: its aggregate is its
// inclusive prefix.
__device__ __forceinline__ void lookback_publish(unsigned long long *status, int64_t tile, unsigned long long aggregate) {
    st_status(&status[tile], (tile == 0 ? kFlagInc : kFlagAgg) | aggregate);
}

// Step 2, called by ONE full warp (all 32 lanes): looks back over predecessor tiles 32 at a time, returns the
// exclusive prefix (sum of aggregates of tiles < tile) to every lane and publishes the inclusive prefix.
__device__ __forceinline__ unsigned long long lookback_resolve(unsigned long long *status, int64_t tile,
                                                               unsigned long long aggregate) {
    const int lane = threadIdx.x & 31;
    unsigned long long exclusive = 0;
    int64_t base = tile - 1;
    while (base >= 0) {
        int64_t idx = base - lane;
        unsigned long long w = kFlagInc;  // lanes before tile 0 behave like "inclusive prefix 0"
        if (idx >= 0) {
            w = ld_status(&status[idx]);
            while ((w >> 62) == 0) {
                __nanosleep(64);
                w = ld_status(&status[idx]);
            }
        }
        unsigned inc_mask = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
        // nearest predecessor holding an inclusive prefix = lowest lane with the flag
        int stop = inc_mask ? (__ffs(inc_mask) - 1) : 31;
        unsigned long long contrib = (lane <= stop) ? (w & kValMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, o);
        exclusive += contrib;
        if (inc_mask) break;
        base -= 32;
    }
    if (lane == 0 && tile != 0) st_status(&status[tile], kFlagInc | (exclusive + aggregate));
    return exclusive;
}

// publish + resolve back to back (block-level tiles)
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, int64_t tile,
                                                                 unsigned long long aggregate) {
    if ((threadIdx.x & 31) == 0) lookback_publish(status, tile, aggregate);
    return lookback_resolve(status, tile, aggregate);
}

}  // namespace acgpu
