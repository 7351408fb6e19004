/*
 * acgpu.h — C ABI of libacgpu.so, the B200-native (sm_100a) drop-in for the matching hot
 * path of RokLenarcic/AhoCorasick.
 *
 * The reference is a pure-Java library with no FFI of its own; its boundary is the public
 * Java API.  Every entry point below names the reference interface it stands in for
 * (paths relative to src/main/java/com/roklenarcic/util/strings/ of the reference); the JNI
 * veneer a maintainer adds on the Java side is shown in INTEGRATION.md and java/.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on
 * success and a negative ACGPU_E* code on failure (text via acgpu_last_error()); handles are
 * opaque; host buffers are caller-owned and only read during the call; haystacks are UTF-16
 * code units (Java char), positions are int32 like Java's int; `end` is exclusive.
 * There is no CPU fallback: every match entry point needs a CUDA device and fails loudly
 * (ACGPU_ENODEVICE) without one.
 */
#ifndef ACGPU_H
#define ACGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACGPU_OK 0
#define ACGPU_EINVAL (-1)      /* bad argument / bad handle */
#define ACGPU_EILLEGALARG (-2) /* the reference would throw IllegalArgumentException */
#define ACGPU_ENODEVICE (-3)   /* no usable CUDA device */
#define ACGPU_ECUDA (-4)       /* a CUDA call failed */
#define ACGPU_ENOMEM (-5)
#define ACGPU_EUNSUPPORTED (-6) /* input outside what the GPU path implements (documented in DESIGN.md) */

#define ACGPU_NO_VALUE 0xFFFFFFFFu

/* Matcher families: the four public class pairs named by BASELINE.json north_star, plus the fifth public pair of the
 * reference ("next" row 1). */
enum acgpu_family {
    ACGPU_AHOCORASICK = 0, /* AhoCorasickSet.java / AhoCorasickMap.java : all overlapping matches */
    ACGPU_LONGEST = 1,     /* LongestMatchSet.java / LongestMatchMap.java : leftmost-longest, non-overlapping */
    ACGPU_SHORTEST = 2,    /* ShortestMatchSet.java / ShortestMatchMap.java : earliest-end, non-overlapping */
    ACGPU_WHOLEWORD = 3,   /* WholeWordMatchSet.java / WholeWordMatchMap.java : whole word-character runs */
    ACGPU_WHOLEWORDLONGEST = 4 /* WholeWordLongestMatchSet.java / WholeWordLongestMatchMap.java : longest whole-word keyword
                                * from every walk start; keywords may hold non-word chars ("as if") */
};

/* A match stream in the reference's listener order.
 *   pos : 2*n int32, (start,end) pairs — the arguments of SetMatchListener.match(haystack,start,end)
 *         (SetMatchListener.java:6) / MapMatchListener.match(haystack,start,end,value) (MapMatchListener.java:6)
 *   val : n uint32 value indices (index into the values Iterable given to the Map constructor), NULL for Sets;
 *         for the Readable overloads (ReadableMatchListener.java:7) only val is meaningful to the listener.
 * Owned by the library; release with acgpu_free_result(). */
typedef struct {
    int64_t n;
    const int32_t *pos;
    const uint32_t *val;
} acgpu_result;

/*
 * Constructors.  Replaces:
 *   AhoCorasickSet(Iterable<String>, boolean[, Thresholder])            AhoCorasickSet.java:16-191
 *   AhoCorasickMap(Iterable<String>, Iterable<T>, boolean[, Thresholder]) AhoCorasickMap.java:20-206
 *   LongestMatchSet/Map   LongestMatchSet.java:15-140 / LongestMatchMap.java:19-201
 *   ShortestMatchSet/Map  ShortestMatchSet.java:14-135 / ShortestMatchMap.java:19-197
 *   WholeWordMatchSet/Map (all six overloads) WholeWordMatchSet.java:16-45,138-205 / WholeWordMatchMap.java:21-53,246-323
 *
 *  chars/offsets : keyword i = chars[offsets[i] .. offsets[i+1])   (UTF-16 code units)
 *  is_null       : optional n_keywords flags; nonzero = Java null (skipped, but consumes a value)
 *  n_values      : -1 => Set; >= 0 => Map whose values Iterable has n_values entries (zip stops at the shorter,
 *                  AhoCorasickMap.java:32).  The value reported for a match is the index of the winning entry.
 *  case_sensitive: as in the reference; folding is java.lang.Character.toLowerCase per UTF-16 unit (Unicode 15).
 *  word_chars    : WholeWord only: 65536 flags = the boolean[] WordCharacters.generateWordCharsFlags(...)
 *                  builds (WordCharacters.java:6-39); NULL = the default table.  acgpu_word_chars() builds all
 *                  three variants.
 *  device        : CUDA device ordinal the tables are uploaded to.
 * The Thresholder argument of the reference only shapes its node objects (never results) and has no counterpart.
 * ACGPU_EILLEGALARG + "<keyword> contains non-word characters." mirrors WholeWordMatchSet.java:149-153.
 */
int acgpu_create_from_keywords(int family, const uint16_t *chars, const int64_t *offsets,
                               const uint8_t *is_null, int64_t n_keywords, int64_t n_values,
                               int case_sensitive, const uint8_t *word_chars, int device,
                               uint64_t *handle);
int acgpu_destroy(uint64_t handle);

/*
 * The same constructors for a caller that has ALREADY built the dictionary trie - the Java-side builder of BASELINE.json
 * north_star (1): the reference's own constructor loops (null skip, keyword/value zip, "last duplicate wins", lower-casing of
 * keyword chars, WordCharacters.trim - AhoCorasickSet.java:20-66, AhoCorasickMap.java:24-77, WholeWordMatchSet.java:138-205)
 * run unchanged in Java and hand over the flattened goto trie instead of the keyword strings; libacgpu derives its device
 * tables (class remap, anchored tiers, hashes) from it and uploads them once.
 *   states are numbered so that parent[s] < s, state 0 is the root (parent[0] = -1);
 *   edge_char[s] = the UTF-16 unit on the edge parent[s] -> s (already lower-cased when case_sensitive == 0);
 *   terminal[s] != 0 : a keyword ends in s;  value[s] = its value index (Maps; < n_values; no index twice);
 *   fail (optional)  : the failure link of every state as the Java builder computed it (AhoCorasickSet.java:68-191 BFS);
 *                      when given it is checked against the links the trie implies (ACGPU_EINVAL on a mismatch) - the
 *                      anchored kernels themselves never follow failure links.
 * struct_size = sizeof(acgpu_automaton_desc) of the caller, so the struct can grow.
 */
typedef struct {
    int32_t struct_size;
    int32_t family;          /* enum acgpu_family */
    int32_t is_map;
    int32_t case_sensitive;
    int32_t device;
    int32_t reserved;
    int64_t n_states;
    const int32_t *parent;     /* [n_states] */
    const uint16_t *edge_char; /* [n_states] */
    const uint8_t *terminal;   /* [n_states] */
    const uint32_t *value;     /* [n_states] or NULL (Sets) */
    int64_t n_values;          /* Maps: entries of the Java-side values array */
    const uint8_t *word_chars; /* WholeWord families: 65536 flags, NULL = default table */
    const int32_t *fail;       /* [n_states] or NULL */
} acgpu_automaton_desc;
int acgpu_create(const acgpu_automaton_desc *desc, uint64_t *handle);
/* host only: the fingerprint acgpu_build_fingerprint would give for the dictionary the descriptor spells */
int acgpu_desc_fingerprint(const acgpu_automaton_desc *desc, uint64_t *fingerprint);

/* Diagnostics, host only (no device needed): runs the same dictionary flattening as acgpu_create_from_keywords and
 * returns a 64-bit fingerprint of every table it would upload.  Flattening is deterministic; large dictionaries are
 * inserted concurrently (one shard per first character class) and must give the fingerprint of the serial insert
 * (environment ACGPU_BUILDER=serial|sharded forces either; ACGPU_BUILD_TIMING=1 prints phase times). */
int acgpu_build_fingerprint(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                            int64_t n_keywords, int64_t n_values, int case_sensitive, const uint8_t *word_chars,
                            uint64_t *fingerprint, double *build_seconds /* optional: time of the flattening alone */);

/* WordCharacters.generateWordCharsFlags (WordCharacters.java:6-16 mode 0, :18-24 mode 1, :26-39 mode 2). */
int acgpu_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n, uint8_t *out65536);

/* Introspection (sizes in the flattened automaton; charBufferSize = AhoCorasickMap.java:53). */
int acgpu_info(uint64_t handle, int64_t *n_nodes, int32_t *n_classes, int32_t *max_len,
               int32_t *char_buffer_size, int64_t *table_bytes);

/* The character-class table of the flattened dictionary: out65536[c] = class of UTF-16 code unit c with the matcher's
 * case folding (Character.toLowerCase, AhoCorasickSet.java:33,229) folded in.  When *has_other is 1, class 0 means
 * "c occurs in no keyword": right after such a char every reference automaton is back in its root state, so the
 * haystack can be cut there into independent pieces (ahocorasick_b200/sharding.py::plan_sync_shards). */
int acgpu_char_classes(uint64_t handle, uint16_t *out65536, int32_t *has_other);

/*
 * match(String haystack, listener) — StringSet.java:4, StringMap.java:8; the scan loops
 * AhoCorasickSet.java:193-252, LongestMatchSet.java:192-265, ShortestMatchSet.java:182-260,
 * WholeWordMatchSet.java:47-132 and their Map twins.  Host buffer in, ordered records out
 * (H2D copy, kernels, D2H copy all inside the call).  The caller replays the records to the listener and
 * stops at the first `false` (family quirks: INTEGRATION.md).
 */
int acgpu_match_utf16(uint64_t handle, const uint16_t *haystack, int32_t n, acgpu_result *out);
void acgpu_free_result(acgpu_result *r);

/*
 * The same call with a COMPACT wire format for dense match streams (at 0.87 matches per char the 8-byte
 * records are 3.5 x the haystack and the call is bound by their D2H copy).  For an AhoCorasickSet whose first chunk holds
 * more than 0.25 matches per char the result is the per-char HIT MASKS the scan kernel produces - 2 bytes per char
 * whatever the density - and the record expansion moves into the caller's replay loop:
 *     for q in [0, n_chars): for every set bit t of masks[q], ascending:   match(haystack, q + 1 - (16 - t), q + 1)
 * which is exactly the reference's order (end ascending, longest first - AhoCorasickSet.java:522-535; a keyword of
 * length d that ends with char q sets bit 16 - d).  Everything else (Maps, sparse streams, the other families, wide
 * alphabets) comes back as records, kind ACGPU_MATCHES_RECORDS.  `n` is the number of matches in both kinds.
 */
#define ACGPU_MATCHES_RECORDS 0
#define ACGPU_MATCHES_MASKS 1
typedef struct {
    int64_t n;             /* matches in the stream */
    int32_t kind;          /* ACGPU_MATCHES_RECORDS: pos/val as in acgpu_result; ACGPU_MATCHES_MASKS: masks */
    int32_t reserved;
    const int32_t *pos;
    const uint32_t *val;
    const uint16_t *masks; /* n_chars hit masks */
    int64_t n_chars;
} acgpu_matches;
int acgpu_match_utf16_compact(uint64_t handle, const uint16_t *haystack, int32_t n, acgpu_matches *out);
void acgpu_free_matches(acgpu_matches *r);
/* Host-only helper: expands masks[first_char, n_chars) into (start,end) pairs, at most `cap` of them; returns the number
 * of matches from first_char on (which may exceed cap), -1 on bad arguments. */
int64_t acgpu_masks_to_records(const uint16_t *masks, int64_t n_chars, int64_t first_char, int32_t *pos_out, int64_t cap);

/*
 * Same scan with the haystack already resident in device memory (kernel-only measurements, multi-GPU
 * shards).  d_pos (capacity cap records, 8 B each) and d_val (4 B each; may be NULL for Sets) are device
 * buffers; *n_out receives the TOTAL number of matches, which may exceed cap (then only the first cap
 * records were written — enlarge and rerun).  cuda_stream: a cudaStream_t, or NULL for the default stream.
 * The call synchronises the stream before returning (it reads back *n_out).
 *
 * Sharding: emit_from/emit_to restrict reported matches; the scan itself may look outside:
 *   AhoCorasick: matches with  emit_from <  end   <= emit_to   (reads max_len - 1 chars before emit_from)
 *   WholeWord:   matches with  emit_from <= start <  emit_to   (reads 1 char before emit_from and needs
 *                n >= min(length of the haystack, emit_to + max_len + 1): a word is a keyword or not by itself)
 *   Longest / Shortest / WholeWordLongest carry a selection chain across positions and need the whole
 *   haystack — pass emit_from = 0, emit_to = n (corpora of independent haystacks shard by haystack).
 */
int acgpu_match_device(uint64_t handle, const void *d_haystack, int64_t n, int64_t emit_from, int64_t emit_to,
                       void *d_pos, void *d_val, int64_t cap, int64_t *n_out, void *cuda_stream);

/*
 * Longest / Shortest: range shards of ONE resident haystack.  The reference's selection state crosses any
 * cut (LongestMatchSet.java:192-265 + SetMatchQueue.java:45-95, ShortestMatchSet.java:182-260): it is the position the
 * scan has consumed up to ("chain position").  A shard's effect on it is a small map - for each of the 16 possible entry
 * offsets: where the chain leaves the shard and how many matches it emits on the way - so shards scan in parallel,
 * exchange their maps (they ride in the all-gather of the counts) and then emit from their true entry:
 *
 *   rank r:  acgpu_chain_shard_begin(h, d_window_r, n_r, n_domain_r, d_map_r, &s, stream)      (entry-independent work)
 *            all-gather d_map (16 x uint64 per rank: exit offset | matches << 8)
 *            entry_0 = 0;  entry_{r+1} = exit offset of map_r[entry_r];  first record of rank r = sum of the matches before
 *            acgpu_chain_shard_finish(s, entry_r, pos_base_r, d_pos, d_val, cap, d_total, stream)
 *
 * Geometry (acgpu_chain_shard_layout): shard r owns the chain positions [lo_r, hi_r) of the haystack; its window is
 * hay[lo_r, min(N, hi_r + lookahead)) in a 16-byte aligned device buffer of its own; lo_0 = 0 and every inner boundary
 * is a multiple of tile_chars (ahocorasick_b200/sharding.py::plan_chain_shards).  n_domain = hi_r - lo_r, or n for the
 * last shard (whose map nobody consults: when its window is not tile-aligned only row 0 is filled, the others are ~0).
 * Entry / exit offsets are relative to lo_r / hi_r and smaller than map_entries (a match is at most 16 chars).
 * Records are positions in the window plus pos_base (pass lo_r for haystack positions).  finish() releases the shard.
 * Needs the start-mask path (at most 31 keyword symbols, keywords of at most 16 chars): ACGPU_EUNSUPPORTED otherwise.
 */
int acgpu_chain_shard_layout(uint64_t handle, int64_t *tile_chars, int64_t *lookahead_chars, int32_t *map_entries);
int acgpu_chain_shard_begin(uint64_t handle, const void *d_window, int64_t n, int64_t n_domain, void *d_map16, uint64_t *shard,
                            void *cuda_stream);
int acgpu_chain_shard_finish(uint64_t shard, int32_t entry, int32_t pos_base, void *d_pos, void *d_val, int64_t cap, void *d_total,
                             void *cuda_stream);

/* Async flavour for benchmarking: enqueues the kernels only; the total lands in *d_total (device int64). */
int acgpu_match_device_async(uint64_t handle, const void *d_haystack, int64_t n, int64_t emit_from,
                             int64_t emit_to, void *d_pos, void *d_val, int64_t cap, void *d_total,
                             void *cuda_stream);
/* Number of kernel launches one acgpu_match_device_async enqueues for this matcher (for bench accounting). */
int acgpu_launches_per_match(uint64_t handle);

/*
 * match(Readable haystack, ReadableMatchListener) — StringMap.java:6; loops AhoCorasickMap.java:208-275,
 * LongestMatchMap.java:203-286, ShortestMatchMap.java:199-291, WholeWordMatchMap.java:55-153.
 * begin -> feed* -> end.  Each feed copies the block through pinned double buffers (cudaMemcpyAsync),
 * scans it with the automaton context carried over from previous blocks, and returns the records that
 * are final so far, in order; positions are offsets in the whole stream.  end() flushes the rest.
 * AhoCorasick / WholeWord streams are PIPELINED: the block of feed k goes up while the records of block k-1 come down,
 * and feed k returns the records of block k-1 (one block later than they could be known; no feed waits for its own scan).
 * Longest / Shortest streams on the start-mask path scan every feed as a chain shard (whole 8 192-position tiles; the rest of
 * the block waits for the next feed).  Matchers that follow the reference loop literally (quirk Q7 tables; keywords beyond what
 * the selection kernels hold: Longest / Shortest > 2 047 chars, WholeWordLongest > 254) collect the feeds and deliver every
 * record with end() - a segment of that loop can be as long as the input.
 */
int acgpu_stream_begin(uint64_t handle, uint64_t *stream_handle);
/* ReadableMatchListener.match(T value) sees VALUES only (ReadableMatchListener.java:7): with values_only on, the feeds of a
 * Map stream return val[] alone (pos == NULL) and the positions never cross PCIe - 4 instead of 12 bytes per match.  Leave
 * it off for ShortestMatchMap: its replay needs the match ends (quirk Q4, ShortestMatchMap.java:241-249). */
int acgpu_stream_set_values_only(uint64_t stream_handle, int on);
int acgpu_stream_feed(uint64_t stream_handle, const uint16_t *chars, int32_t n, acgpu_result *out);
int acgpu_stream_end(uint64_t stream_handle, acgpu_result *out);

const char *acgpu_last_error(void);
const char *acgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
