/*
 * ac_oracle.h — CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A literal, loop-by-loop C restatement of the matching hot path of
 * RokLenarcic/AhoCorasick (pure Java; no JVM exists in this image, so the
 * reference itself cannot be run here).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (libacgpu.so) never links, imports or calls it.
 *
 * Parity pin: the restatement is checked against every ordered known-answer
 * vector the reference's own tests hold for this path (MatchQueueTest, the
 * literal SetTest/MapTest cases with their brute-force counting oracles, the
 * README worked examples) — see tests/test_oracle_golden.py.  Ordered
 * (start,end,value) streams, case-insensitive mode, early stop and custom
 * word characters are NOT asserted by any reference test; for those the
 * oracle is pinned only by being a literal restatement ("parity unpinned by
 * the reference's tests" for those aspects; stated in DESIGN.md too).
 *
 * All file:line citations are relative to
 * /root/reference/src/main/java/com/roklenarcic/util/strings/.
 */
#ifndef AC_ORACLE_H
#define AC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORA_AHOCORASICK = 0, /* AhoCorasickSet.java / AhoCorasickMap.java */
    ORA_LONGEST = 1,     /* LongestMatchSet.java / LongestMatchMap.java */
    ORA_SHORTEST = 2,    /* ShortestMatchSet.java / ShortestMatchMap.java */
    ORA_WHOLEWORD = 3,   /* WholeWordMatchSet.java / WholeWordMatchMap.java */
    ORA_WHOLEWORDLONGEST = 4 /* WholeWordLongestMatchSet.java / WholeWordLongestMatchMap.java */
};

typedef struct ora_matcher ora_matcher;

/* One listener call. For the Readable overloads start = end = -1 (values only,
 * ReadableMatchListener.java:7). value = index of the dictionary entry whose
 * value object the reference would deliver, or -1 for a Set. */
typedef struct {
    int32_t start;
    int32_t end;
    int32_t value;
} ora_match;

/* Listener: return nonzero to continue, 0 to stop (SetMatchListener.java:6). */
typedef int (*ora_listener)(void *ctx, int32_t start, int32_t end, int32_t value);

/*
 * Build a matcher.
 *  chars/offsets : keyword i is chars[offsets[i] .. offsets[i+1]) (UTF-16 code units)
 *  is_null       : optional; is_null[i] != 0 means keyword i is a Java null
 *  n_keywords    : entries in the keyword iterable
 *  n_values      : -1 => Set.  >=0 => Map whose values iterable has n_values
 *                  entries; keywords and values are zipped and stop at the shorter
 *                  (AhoCorasickMap.java:32).  The value of entry i is the index i.
 *  word_chars    : WholeWord only: 65536 flags (the boolean[] of WordCharacters.java),
 *                  NULL = default table (WordCharacters.java:6-16).
 *  err/errlen    : receives "<keyword> contains non-word characters." style text
 *                  when the reference would throw IllegalArgumentException
 *                  (WholeWordMatchSet.java:149-153); then NULL is returned.
 */
ora_matcher *ora_create(int family, const uint16_t *chars, const int64_t *offsets,
                        const uint8_t *is_null, int64_t n_keywords, int64_t n_values,
                        int case_sensitive, const uint8_t *word_chars,
                        char *err, int errlen);
void ora_destroy(ora_matcher *m);

/* charBufferSize of the Map (AhoCorasickMap.java:53). */
int32_t ora_char_buffer_size(const ora_matcher *m);
int64_t ora_node_count(const ora_matcher *m);

/* match(String, listener): literal loops. Returns number of listener calls. */
int64_t ora_match_string(const ora_matcher *m, const uint16_t *hay, int32_t n,
                         ora_listener cb, void *ctx);

/* match(Readable, listener) (Maps only in the reference; allowed for any matcher here).
 * The Readable is emulated over hay[0..n): every read() transfers
 * min(buffer.remaining(), chars left, schedule[k]) chars (schedule NULL => fill the
 * CharBuffer, like java.io.StringReader) and returns -1 at end of input. */
int64_t ora_match_readable(const ora_matcher *m, const uint16_t *hay, int64_t n,
                           const int32_t *schedule, int64_t n_schedule,
                           ora_listener cb, void *ctx);

/* Convenience: run with a recording listener that answers `false` on its
 * stop_after-th call (stop_after <= 0: never). Writes up to cap records,
 * returns the total number of listener calls. readable != 0 selects the
 * Readable overload (default fill schedule). */
int64_t ora_match_collect(const ora_matcher *m, const uint16_t *hay, int64_t n,
                          int readable, int64_t stop_after, ora_match *out, int64_t cap);

/* Count-only run with an empty listener (the reference's "performanceListener",
 * SetTest.java:168-173) — used as the timed CPU baseline. */
int64_t ora_match_count(const ora_matcher *m, const uint16_t *hay, int32_t n);

/* SetMatchQueue (SetMatchQueue.java) exposed for the MatchQueueTest known answers. */
typedef struct ora_queue ora_queue;
ora_queue *ora_queue_new(void);
void ora_queue_free(ora_queue *q);
int ora_queue_push(ora_queue *q, int32_t length, int32_t idx);
/* returns number of records written to out (start,end pairs), flushed up to purge_to */
int64_t ora_queue_match_and_clear(ora_queue *q, int32_t purge_to, ora_match *out, int64_t cap);

/* JDK character helpers rebuilt from Unicode 15.0 (tools/gen_unicode_tables.py). */
uint16_t ora_to_lower(uint16_t c);
int ora_is_letter_or_digit(uint16_t c);
/* WordCharacters.generateWordCharsFlags variants (WordCharacters.java:6-39):
 * mode 0 default, 1 custom-only, 2 default+toggles. out = 65536 flags. */
void ora_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n,
                    uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
