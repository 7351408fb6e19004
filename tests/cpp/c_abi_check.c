/* include/acgpu.h must be a plain C header (C99, no C++/torch types) and every declared entry point must link from C.
 * Host-only calls are exercised; the matching entry points are only referenced (they need a CUDA device).
 * Built and run by tests/test_host_cpu.py::test_header_is_plain_c_and_links_from_c. */
#include <stdio.h>
#include <string.h>

#include "acgpu.h"

int main(void) {
    /* every entry point, by address: a missing export fails at link time */
    typedef void (*fn)(void);
    const fn syms[] = {(fn)acgpu_create_from_keywords, (fn)acgpu_build_fingerprint, (fn)acgpu_destroy, (fn)acgpu_word_chars,
                       (fn)acgpu_info, (fn)acgpu_char_classes, (fn)acgpu_match_utf16, (fn)acgpu_free_result, (fn)acgpu_match_device,
                       (fn)acgpu_match_device_async, (fn)acgpu_launches_per_match, (fn)acgpu_stream_begin, (fn)acgpu_stream_feed,
                       (fn)acgpu_stream_end, (fn)acgpu_last_error, (fn)acgpu_version, (fn)acgpu_match_utf16_compact, (fn)acgpu_free_matches,
                       (fn)acgpu_masks_to_records, (fn)acgpu_chain_shard_layout, (fn)acgpu_chain_shard_begin, (fn)acgpu_chain_shard_finish, (fn)acgpu_stream_set_values_only,
                       (fn)acgpu_create, (fn)acgpu_desc_fingerprint};
    unsigned i, n_syms = (unsigned)(sizeof syms / sizeof syms[0]);
    for (i = 0; i < n_syms; i++)
        if (!syms[i]) return 10;

    static uint8_t flags[65536];
    const uint16_t toggled[2] = {'_', '='};
    const uint8_t toggles[2] = {0, 1};
    if (acgpu_word_chars(2, toggled, toggles, 2, flags) != ACGPU_OK) return 11;
    if (flags['_'] || !flags['='] || !flags['a'] || flags[' ']) return 12;

    /* "he", "she", "hers" + a null entry; WholeWord refuses "as if" with the reference's message */
    const uint16_t chars[] = {'h', 'e', 's', 'h', 'e', 'h', 'e', 'r', 's', 'a', 's', ' ', 'i', 'f'};
    const int64_t offsets[] = {0, 2, 5, 9, 9, 14};
    const uint8_t is_null[] = {0, 0, 0, 1, 0};
    uint64_t fp1 = 0, fp2 = 0, fp3 = 0;
    double secs = -1.0;
    if (acgpu_build_fingerprint(ACGPU_AHOCORASICK, chars, offsets, is_null, 4, 4, 0, NULL, &fp1, &secs) != ACGPU_OK) return 13;
    if (acgpu_build_fingerprint(ACGPU_AHOCORASICK, chars, offsets, is_null, 4, 4, 0, NULL, &fp2, NULL) != ACGPU_OK) return 14;
    if (acgpu_build_fingerprint(ACGPU_LONGEST, chars, offsets, is_null, 4, 4, 0, NULL, &fp3, NULL) != ACGPU_OK) return 15;
    if (fp1 != fp2 || fp1 == fp3 || secs < 0.0) return 16;
    if (acgpu_build_fingerprint(ACGPU_WHOLEWORD, chars, offsets, is_null, 5, -1, 1, NULL, &fp3, NULL) != ACGPU_EILLEGALARG) return 17;
    if (strcmp(acgpu_last_error(), "as if contains non-word characters.") != 0) return 18;
    if (acgpu_build_fingerprint(ACGPU_WHOLEWORDLONGEST, chars, offsets, is_null, 5, -1, 1, NULL, &fp3, NULL) != ACGPU_OK) return 19;
    if (acgpu_destroy(0) != ACGPU_EINVAL) return 20;
    {
        /* the same dictionary as a flattened goto trie (acgpu_automaton_desc): h e | s h e | r s below "he" */
        const int32_t parent[] = {-1, 0, 1, 0, 3, 4, 2, 6};
        const uint16_t edge[] = {0, 'h', 'e', 's', 'h', 'e', 'r', 's'};
        const uint8_t term[] = {0, 0, 1, 0, 0, 1, 0, 1};
        const uint32_t value[] = {0, 0, 0, 0, 0, 1, 0, 2};
        const int32_t fail_links[] = {0, 0, 0, 0, 1, 2, 0, 3};
        acgpu_automaton_desc d;
        uint64_t fpd = 0;
        memset(&d, 0, sizeof d);
        d.struct_size = (int32_t)sizeof d;
        d.family = ACGPU_AHOCORASICK;
        d.is_map = 1;
        d.n_states = 8;
        d.parent = parent;
        d.edge_char = edge;
        d.terminal = term;
        d.value = value;
        d.n_values = 4;
        d.fail = fail_links;
        if (acgpu_desc_fingerprint(&d, &fpd) != ACGPU_OK) return 40;
        if (fpd != fp1) return 41;
        d.struct_size = 8;
        if (acgpu_desc_fingerprint(&d, &fpd) != ACGPU_EINVAL) return 42;
    }
    if (!acgpu_version() || !*acgpu_version()) return 21;
    acgpu_result empty = {0, NULL, NULL};
    acgpu_free_result(&empty); /* releasing an empty result is a no-op */
    {
        /* compact wire format, host-side expansion: "abcd" with keywords ending at char 1 (length 2) and char 3 (lengths 4, 1) */
        const uint16_t masks[4] = {0, 1u << 14, 0, (1u << 12) | (1u << 15)};
        int32_t pos[6];
        if (acgpu_masks_to_records(masks, 4, 0, pos, 3) != 3) return 22;
        if (pos[0] != 0 || pos[1] != 2 || pos[2] != 0 || pos[3] != 4 || pos[4] != 3 || pos[5] != 4) return 23;
        if (acgpu_masks_to_records(masks, 4, 2, pos, 0) != 2) return 24;
        acgpu_matches none;
        memset(&none, 0, sizeof none);
        acgpu_free_matches(&none);
    }
    printf("c abi ok: %u entry points, version %s\n", n_syms, acgpu_version());
    return 0;
}
