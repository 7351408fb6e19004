// Host-side dictionary flattening for libacgpu.so.
//
// Replaces the reference's constructors (object-graph tries of HashmapNode/RangeNode with fail
// links, e.g. AhoCorasickSet.java:16-191) by flat, upload-once tables:
//   * cls[65536]   UTF-16 code unit -> character class, with Character.toLowerCase folded in
//                  when case-insensitive (AhoCorasickSet.java:33,229); class 0 = "in no keyword"
//   * an anchored trie over class strings, stored as a direct root table plus an open-addressing
//     edge table  (parent, class) -> (child, flags)   in 16-byte slots
//   * per-node value index (Q6 rules: last duplicate wins; first wins for ShortestMatchMap)
//   * dense DFA / shared-memory tiers are derived from this trie by later stages.
// The trie is built over REVERSED keywords for the AhoCorasick family (every end position walks
// backwards and meets all keywords ending there, longest last) and over forward keywords for
// Longest / Shortest / WholeWord (every start position walks forwards).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace acgpu {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kInfoTerminal = 1u;
constexpr uint32_t kInfoHasChildren = 2u;

struct Edge {  // one 16-byte slot of the device edge table
    uint32_t parent;
    uint32_t cls;
    uint32_t child;
    uint32_t info;
};

struct RootEdge {
    uint32_t child;
    uint32_t info;
};

inline uint32_t edge_hash(uint32_t parent, uint32_t c) {
    uint32_t h = parent * 0x9E3779B1u ^ (c * 0x85EBCA6Bu + 0x7F4A7C15u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

struct IllegalArgument : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Generation-2 tables for narrow alphabets (at most 32 classes, every context of max_len classes packs into
// 60 bits):
//   * levels 1..K of the anchored trie are DIRECT-INDEXED by the mixed-radix number of the last j classes
//     (most recent class = lowest digit): terminal bitmaps for j < K, 2 bits (terminal, has-children) for j = K;
//     small enough to live in shared memory
//   * kidmask[level-K index] = set of classes that continue the context (exact), so level K+1 is only probed
//     for contexts that exist
//   * deeper nodes are PATH-COMPRESSED: a "head" is a level-(K+1) node, a child of a branching deep node, or the
//     continuation of a chain that outgrew one entry.  One 16-byte entry describes the head and the unbranched chain
//     hanging below it:
//         x = 28-bit tag << 4 | 8 (occupied)
//         y = child mask of the chain's LAST node (exact)
//         z,w (64 bits) = chain classes (b bits each, first step lowest) | chain length L << 40 | terminal flags of
//                         the head and the L chain nodes << 44        (L <= 8, b <= 5)
//     so a context that reaches a head resolves all of its deeper levels with ONE probe in the common case (keyword
//     tails rarely branch).  Entries live in a bucketed open-addressing table keyed by the packed context of the
//     head (two entries per 32-byte sector, bucket = fastrange(hash)).  Child masks are exact, so every probe is for
//     a key that exists, and the builder guarantees that no bucket on a key's probe path holds another entry with
//     the same tag - a tag match is exact.
constexpr int kTierChainMax = 8;
constexpr int kWideChainMax = 8;
struct TierTables {
    bool ok = false;
    int32_t C = 0;        // radix = number of classes (including class 0 = other)
    int32_t b = 0;        // bits per class in the packed context
    int32_t K = 0;        // direct-indexed levels
    uint32_t term_levels = 0;        // bit j: some keyword has length j (j <= K)
    uint32_t pow_c[10] = {0};        // C^(j-1)
    // levels 1..K laid out by ROW for k_tier_mask: the row of level j is the mixed-radix number of the j-1 classes
    // BEFORE the current one, the current class selects a bit.  Level j < K: one word per row, the terminal bit of
    // class c sits at bit (c + 16 - j) & 31, so rotating the word right by c drops it on bit 16 - j = where a keyword
    // of length j lives in a hit mask.  Level K: two words per row, terminal bits (same rotation) and has-children bits
    // (bit c).
    uint32_t row_off[10] = {0};      // word offset of level j's rows inside row_words (j = 1..K)
    std::vector<uint32_t> row_words;
    std::vector<uint32_t> kidmask;   // two words per level-K context G = (c[q], .., c[q+1-K]): [2G] bit c = G continues backwards with class c (child mask of the node); [2G+1] bit e = the context of the NEXT position, (e, c[q], .., c[q+2-K]), continues with c[q+1-K]
    // PAIR rows for k_tier_pair (kernel_pair.cuh), at most 31 classes: one 8-byte row {fwd, back} answers TWO positions.
    // The row of level j for the pair (q, q + 1) is the mixed-radix number of the j-1 classes c[q], .., c[q-j+2] (c[q]
    // lowest digit).  fwd: the terminal bit of node (a, row) - a keyword of length j ending at q + 1 with class a there -
    // at bit (a + 16 - j) & 31; back: the terminal bit of node (row, y) - a keyword of length j ending at q whose first
    // class is y - at bit (y + 16 - j) & 31.  Level K's fwd word also carries the GATE bit (pair_gate_bit, a position no
    // class uses): some level-K context (row, y) has a continuation mask that is not empty, i.e. the pair's gather of
    // kidmask can find something.
    // With at most 30 classes the level-K fwd word has a second spare position, the LOW bit (pair_low_bit): the K-1
    // classes that name the row are a keyword (it ends at q) - and prow_off[0] is a compact copy of level K-1's fwd words
    // (one word per row) for the odd position, so dictionaries whose only keywords below level K are of length K-1 pay
    // one 8-byte and one 4-byte shared load per pair.
    uint32_t prow_off[10] = {0};     // word offset of level j's pair rows inside prow_words (j = 1..K), even; [0]: compact level K-1
    std::vector<uint32_t> prow_words;
    uint32_t pair_gate_bit = 0;
    uint32_t pair_low_bit = 0;
    std::vector<uint32_t> buckets;   // 8 words per bucket: 2 entries x {x, y, z, w}
    uint32_t n_buckets = 0;
    uint64_t hash_seed = 0;
    uint64_t n_deep = 0;             // trie nodes deeper than K
    uint64_t n_heads = 0;            // entries of the compressed table
    // Map values: a keyword is looked up by its whole packed context in a bucketed hash table (two 16-byte entries
    // {key lo, key hi, value, 0} per 32-byte bucket; key = packed classes | (length - 1) << 60, compared exactly) - one
    // gather per record
    std::vector<uint32_t> vbuckets;  // 8 words per bucket
    uint32_t n_vbuckets = 0;
    uint64_t vseed = 0;
};

// WholeWord (generation 2): a word is a keyword iff its class string is in a hash table - no trie walk.
//   wcls[65536]  code unit -> class (15 bits, case folding folded in) | word-char flag << 15
//   buckets      two 16-byte entries {h1, length, pool offset, value} per 32-byte bucket, empty: pool offset = 0xFFFFFFFF
//   pool         the keywords' classes (u16 each) - compared exactly after a hash and length match
constexpr int kWwMaxLen = 255;
struct WwTables {
    bool ok = false;
    std::vector<uint16_t> wcls;
    std::vector<uint32_t> buckets;
    uint32_t n_buckets = 0;
    std::vector<uint16_t> pool;
    // generation 3 (kernel_ww3.cuh): the table is keyed by ww_poly_key, and a one-hash Bloom filter over the keys (a power of
    // two of bits, at most 64 KB, 0 = none: it lives in shared memory and only pays for dictionaries it can separate)
    bool poly = false;
    std::vector<uint32_t> bloom;
    uint32_t bloom_bits = 0;
};

// Generation-3 hash of a word (kernel_ww3.cuh): a polynomial in B over x = (class + 1) | 1 << 31 of its chars, mod 2^32 -
//     poly(c_0 .. c_{L-1}) = sum x_j * B^(L-1-j)
// so that with running prefix values G over the haystack (G[i] = G[i-1] * B + x[i], x = 0 for non-word chars) the hash of
// the run [s, t) is G[t-1] - G[s-1] * B^(t-s): no per-word loop on the device.  ww_poly_key mixes the length in; the
// Bloom filter takes the key's HIGH bits (the best mixed ones), the bucket another multiplicative mix.
constexpr uint32_t kWwPolyB = 0x9E3779B1u;
inline constexpr uint32_t ww_poly_digit(uint32_t cls) { return (cls + 1u) | 0x80000000u; }
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t ww_poly_key(uint32_t poly, uint32_t len) {
    uint32_t k = (poly ^ (len * 0x27D4EB2Fu)) * 0x85EBCA6Bu;
    k ^= k >> 15;
    k *= 0xC2B2AE35u;
    return k;
}
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t ww_poly_spread(uint32_t key) { return (key ^ (key >> 16)) * 0x2C1B3C6Du; }

// hash of a class string, shared by the builder and k_ww_scan: FNV-1a over PAIRS of classes (c[2i] | c[2i+1] << 16, a
// lone last class with a zero upper half), murmur-style finish
struct WwHash {
    uint32_t h1 = 0x811C9DC5u;
#ifdef __CUDACC__
    __host__ __device__
#endif
    inline void add_pair(uint32_t pair) { h1 = (h1 ^ pair) * 0x01000193u; }
#ifdef __CUDACC__
    __host__ __device__
#endif
    inline void finish(uint32_t len) {
        h1 ^= len * 0x27D4EB2Fu;
        h1 ^= h1 >> 16;
        h1 *= 0x85EBCA6Bu;
        h1 ^= h1 >> 13;
    }
#ifdef __CUDACC__
    __host__ __device__
#endif
    inline uint32_t spread() const {
        uint32_t x = h1 * 0xC2B2AE35u;
        return x ^ (x >> 16);
    }
};

// bucket hash of the keyword -> value table (32-bit arithmetic only: it runs once per Map record on the device)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t value_hash32(uint32_t lo, uint32_t hi) {
    uint32_t h = (lo * 0x9E3779B1u) ^ (hi * 0x85EBCA6Bu);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// hash of a keyword's class string for HostAutomaton::wide_vals: feed the classes in the order a match meets them walking
// back from its last char, then finish with the length
struct WideValHash {
    uint64_t h = 0xCBF29CE484222325ull;
#ifdef __CUDACC__
    __host__ __device__
#endif
    inline void add(uint32_t c) { h = (h ^ c) * 0x100000001B3ull; }
#ifdef __CUDACC__
    __host__ __device__
#endif
    inline uint64_t finish(uint32_t len) {
        uint64_t x = h ^ (static_cast<uint64_t>(len) * 0x9E3779B97F4A7C15ull);
        x ^= x >> 29;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 32;
        x *= 0x94D049BB133111EBull;
        x ^= x >> 29;
        return x;
    }
};

inline uint64_t deep_hash64(uint64_t key, uint64_t seed) {
    uint64_t h = key ^ seed;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    return h;
}

struct HostAutomaton {
    int family = 0;
    bool is_map = false;
    bool case_sensitive = true;
    bool reversed = false;   // trie over reversed keywords (AhoCorasick family)
    bool has_other = true;   // class 0 is "char occurs in no keyword"
    int32_t max_len = 0;     // longest effective keyword (chars)
    int32_t n_classes = 1;
    int32_t char_buffer_size = 4096;  // AhoCorasickMap.java:53
    int64_t n_nodes = 1;
    int64_t n_keywords_effective = 0;
    std::vector<uint16_t> cls;      // 65536
    std::vector<uint32_t> wordbits; // 2048 words: raw word-char bitmap (WholeWord)
    std::vector<RootEdge> root;     // n_classes
    std::vector<Edge> edges;        // pow2 slots, empty = parent kNone
    uint32_t edge_mask = 0;
    std::vector<uint32_t> node_value;  // n_nodes, kNone when not terminal / Set
    std::vector<uint8_t> node_info;    // n_nodes
    std::vector<uint32_t> depth_count; // nodes per depth (diagnostics / tiering)
    TierTables tier;                   // generation-2 tables (tier.ok == false: not applicable)
    // AhoCorasick family outside the tier envelope (kernel_wide.cuh): 32-bit hit masks by anchored walks over the edge
    // table; levels 1 and 2 come from a class-pair table when it fits shared memory.  wide_pair[2 * (c0 * C + c1)] =
    // {level-2 node of the reversed context (c0, c1) or kNone, info2 | info1 << 8 | (level-1 node exists) << 16}
    bool wide_ok = false;
    std::vector<uint32_t> wide_pair;
    // Path-compressed edges below level 2 for k_wide_tile (with the pair table, i.e. at most 64 classes).  A JUNCTION is
    // a level-2 node, a node with several children, or the end of a chain that outgrew one entry; every edge that leaves a
    // junction has one 32-byte ENTRY describing the child and the unbranched chain hanging below it, and the entries of
    // one junction's children are CONTIGUOUS, in class order.  With the junction's exact child mask a walk finds its
    // next entry by a population count - no hashing, no probing, never a gather for an edge that does not exist:
    //     [0,1] 64-bit child mask of the node at the END of the chain (bit c: it has a child of class c)
    //     [2]   index of the first entry of the end node's children
    //     [3]   chain length L (bits 0..3) | terminal flags << 4 (bit i: the i-th node of the run is a keyword end, 0 = the child)
    //     [4,5] the classes of the L chain steps, 8 bits each (L <= 8), step k in byte 7 - k of the 64-bit word: the same
    //           order as the haystack classes they are compared with (the walk runs right to left)
    //     [6]   node id of the end node   [7] 0
    // wide_pair16[4 * (c0 * C + c1)] = {first entry of the level-2 node's children, info2 | info1 << 8 | (level-1 node
    // exists) << 16 | (level-2 node exists) << 17, child mask lo, hi} is the shared-memory table of levels 1 and 2.
    std::vector<uint32_t> wide_chain;
    std::vector<uint32_t> wide_pair16;
    // Map values of the wide path: keyword -> value by a 64-bit hash of its class string (wide_value_hash: the classes as
    // a match meets them walking back from its last char), two 16-byte entries {hash lo, hash hi, length | 1 << 31, value}
    // per 32-byte bucket, linear probing over buckets.  A record is a real match, so its key is in the table: the builder
    // checks that no two keywords share (hash, length), and a hit on those IS the keyword - one gather per record.
    // Empty: no table (Set, or a hash collision) - k_wide_emit walks the trie again instead.
    std::vector<uint32_t> wide_vals;
    uint32_t wide_n_vbuckets = 0;
    WwTables ww;                       // WholeWord hash tables (ww.ok == false: not applicable)
    bool ww_plain = true;              // WholeWordLongest: no keyword holds a non-word char (then it equals WholeWord)
    std::vector<uint16_t> wwl_wcls;    // WholeWordLongest with phrases (k_wwl_starts): code unit -> class | word-char flag << 15; empty: not applicable
    // Quirk Q7: a case-insensitive WholeWord matcher whose word-char table is not closed under Character.toLowerCase.  The
    // reference tests the LOWER-CASED char where a walk fails (WholeWordMatchSet.java:96-101) and the RAW char in its two
    // scroll loops (:113,118; the Readable overload lower-cases there too, WholeWordMatchMap.java:328), and its trie may
    // hold chars that are no word chars - "maximal runs of word chars" no longer describes it.  Such matchers follow the
    // loop literally (kernel_wwlit.cuh) and need the table in both views.
    bool ww_literal = false;
    std::vector<uint32_t> wordbits_fold;  // 2048 words: bit c = wordChars[toLowerCase(c)] (ww_literal only)
};

// Throws IllegalArgument with the reference's message for WholeWord keywords holding non-word chars.
HostAutomaton build_automaton(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                              int64_t n_keywords, int64_t n_values, bool case_sensitive,
                              const uint8_t *word_chars);

// 64-bit FNV-1a over every table of the flattened automaton (tests: the serial and the sharded builder must agree, and
// a build must be deterministic)
uint64_t automaton_fingerprint(const HostAutomaton &a);

// WordCharacters.generateWordCharsFlags (WordCharacters.java:6-39)
void make_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n, uint8_t *out65536);

const uint16_t *java_lower_table();

}  // namespace acgpu
