// TEST INFRASTRUCTURE ONLY — never shipped, never loaded by the product or by any GPU test.
//
// libacgpu_mock_oracle.so: the C ABI of include/acgpu.h answered by the CPU oracle (oracle/ac_oracle.c), so that the
// host-side logic of the C++ mirror (include/acgpu.hpp: keyword packing, value zipping, listener replay, early-stop
// quirks, Readable fill batching) and the expectations of tests/cpp/reference_style_test.cpp can be exercised in the
// CPU-only test run (`pytest -m "not gpu"`), where libacgpu.so itself refuses to match (ACGPU_ENODEVICE).
// The device entry points (acgpu_match_device*) are not mocked: they fail.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/acgpu.h"
#include "../../oracle/ac_oracle.h"

namespace {
thread_local std::string g_err;
int fail(int rc, const std::string &m) {
    g_err = m;
    return rc;
}
struct Mock {
    ora_matcher *m;
    bool is_map;
    int family;
};
struct MockStream {
    Mock *owner;
    std::vector<uint16_t> chars;
};
int collect_cb(void *ctx, int32_t s, int32_t e, int32_t v) {
    auto *out = static_cast<std::vector<ora_match> *>(ctx);
    out->push_back(ora_match{s, e, v});
    return 1;  // never stop: the ABI returns the whole ordered stream, the caller replays it
}
int fill(const Mock *mk, const uint16_t *hay, int32_t n, acgpu_result *out) {
    std::vector<ora_match> rec;
    static const uint16_t none = 0;
    ora_match_string(mk->m, n ? hay : &none, n, collect_cb, &rec);
    out->n = (int64_t)rec.size();
    int32_t *pos = new int32_t[2 * rec.size() + 1];
    uint32_t *val = mk->is_map ? new uint32_t[rec.size() + 1] : nullptr;
    for (size_t i = 0; i < rec.size(); i++) {
        pos[2 * i] = rec[i].start;
        pos[2 * i + 1] = rec[i].end;
        if (val) val[i] = (uint32_t)rec[i].value;
    }
    out->pos = pos;
    out->val = val;
    return ACGPU_OK;
}
}  // namespace

extern "C" {

int acgpu_create_from_keywords(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                               int64_t n_keywords, int64_t n_values, int case_sensitive, const uint8_t *word_chars, int,
                               uint64_t *handle) {
    char err[4096] = {0};
    ora_matcher *m = ora_create(family, chars, offsets, is_null, n_keywords, n_values, case_sensitive, word_chars, err, sizeof err);
    if (!m) return fail(ACGPU_EILLEGALARG, err);
    *handle = (uint64_t)(uintptr_t) new Mock{m, n_values >= 0, family};
    return ACGPU_OK;
}
// the trie descriptor: the dictionary it spells (Maps: entry v = the keyword of the state with value index v) through the oracle
int acgpu_create(const acgpu_automaton_desc *d, uint64_t *handle) {
    if (!d || !handle || d->struct_size < (int32_t)sizeof(acgpu_automaton_desc) || d->n_states < 1) return fail(ACGPU_EINVAL, "mock: bad descriptor");
    std::vector<int64_t> state_of;
    if (d->is_map) {
        state_of.assign((size_t)d->n_values, -1);
        for (int64_t s = 1; s < d->n_states; s++)
            if (d->terminal[s]) state_of[d->value[s]] = s;
    } else {
        for (int64_t s = 1; s < d->n_states; s++)
            if (d->terminal[s]) state_of.push_back(s);
    }
    std::vector<uint16_t> chars;
    std::vector<int64_t> offsets(1, 0);
    std::vector<uint8_t> is_null;
    for (int64_t s : state_of) {
        std::vector<uint16_t> rev;
        for (int64_t t = s; t > 0; t = d->parent[t]) rev.push_back(d->edge_char[t]);
        chars.insert(chars.end(), rev.rbegin(), rev.rend());
        offsets.push_back((int64_t)chars.size());
        is_null.push_back(s < 0 ? 1 : 0);
    }
    chars.push_back(0);
    is_null.push_back(0);
    return acgpu_create_from_keywords(d->family, chars.data(), offsets.data(), is_null.data(), (int64_t)state_of.size(),
                                      d->is_map ? d->n_values : -1, d->case_sensitive, d->word_chars, d->device, handle);
}
int acgpu_desc_fingerprint(const acgpu_automaton_desc *, uint64_t *) { return fail(ACGPU_EUNSUPPORTED, "mock: no builder"); }
int acgpu_build_fingerprint(int, const uint16_t *, const int64_t *, const uint8_t *, int64_t, int64_t, int, const uint8_t *, uint64_t *,
                            double *) {
    return fail(ACGPU_EUNSUPPORTED, "mock: no builder");
}
int acgpu_destroy(uint64_t h) {
    Mock *mk = (Mock *)(uintptr_t)h;
    if (!mk) return fail(ACGPU_EINVAL, "bad handle");
    ora_destroy(mk->m);
    delete mk;
    return ACGPU_OK;
}
int acgpu_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n, uint8_t *out) {
    ora_word_chars(mode, chars, toggles, n, out);
    return ACGPU_OK;
}
int acgpu_info(uint64_t h, int64_t *n_nodes, int32_t *n_classes, int32_t *max_len, int32_t *cbs, int64_t *bytes) {
    Mock *mk = (Mock *)(uintptr_t)h;
    if (n_nodes) *n_nodes = ora_node_count(mk->m);
    if (n_classes) *n_classes = 0;
    if (max_len) *max_len = 0;
    if (cbs) *cbs = ora_char_buffer_size(mk->m);
    if (bytes) *bytes = 0;
    return ACGPU_OK;
}
int acgpu_char_classes(uint64_t, uint16_t *, int32_t *) { return fail(ACGPU_EUNSUPPORTED, "mock: no class table"); }
int acgpu_match_utf16(uint64_t h, const uint16_t *hay, int32_t n, acgpu_result *out) { return fill((Mock *)(uintptr_t)h, hay, n, out); }
void acgpu_free_result(acgpu_result *r) {
    if (!r) return;
    delete[] r->pos;
    delete[] r->val;
    r->pos = nullptr;
    r->val = nullptr;
    r->n = 0;
}
// compact call: an AhoCorasickSet whose matches all fit 16-bit hit masks answers with MASKS (the wire format libacgpu.so
// picks for dense streams), so the CPU suite drives the mirrors' lazy mask replay; everything else answers with records
int acgpu_match_utf16_compact(uint64_t h, const uint16_t *hay, int32_t n, acgpu_matches *out) {
    Mock *mk = (Mock *)(uintptr_t)h;
    std::memset(out, 0, sizeof *out);
    acgpu_result r{0, nullptr, nullptr};
    int rc = fill(mk, hay, n, &r);
    if (rc != ACGPU_OK) return rc;
    bool fits = mk->family == ACGPU_AHOCORASICK && !mk->is_map && r.n > 0;
    for (int64_t i = 0; fits && i < r.n; i++) fits = r.pos[2 * i + 1] - r.pos[2 * i] <= 16;
    out->n = r.n;
    if (!fits) {
        out->kind = ACGPU_MATCHES_RECORDS;
        out->pos = r.pos;
        out->val = r.val;
        return ACGPU_OK;
    }
    uint16_t *masks = new uint16_t[n]();
    for (int64_t i = 0; i < r.n; i++) masks[r.pos[2 * i + 1] - 1] |= (uint16_t)(1u << (16 - (r.pos[2 * i + 1] - r.pos[2 * i])));
    acgpu_free_result(&r);
    out->kind = ACGPU_MATCHES_MASKS;
    out->masks = masks;
    out->n_chars = n;
    return ACGPU_OK;
}
void acgpu_free_matches(acgpu_matches *r) {
    if (!r) return;
    delete[] r->pos;
    delete[] r->val;
    delete[] r->masks;
    std::memset(r, 0, sizeof *r);
}
int64_t acgpu_masks_to_records(const uint16_t *masks, int64_t n_chars, int64_t first_char, int32_t *pos_out, int64_t cap) {
    int64_t k = 0;
    for (int64_t q = first_char; q < n_chars; q++)
        for (uint32_t w = masks[q]; w; w &= w - 1u, k++)
            if (k < cap) {
                pos_out[2 * k] = (int32_t)(q + 1 - (16 - __builtin_ctz(w)));
                pos_out[2 * k + 1] = (int32_t)(q + 1);
            }
    return k;
}
int acgpu_match_device(uint64_t, const void *, int64_t, int64_t, int64_t, void *, void *, int64_t, int64_t *, void *) {
    return fail(ACGPU_ENODEVICE, "mock: no device entry points");
}
int acgpu_match_device_async(uint64_t, const void *, int64_t, int64_t, int64_t, void *, void *, int64_t, void *, void *) {
    return fail(ACGPU_ENODEVICE, "mock: no device entry points");
}
int acgpu_launches_per_match(uint64_t) { return 0; }
int acgpu_chain_shard_layout(uint64_t, int64_t *, int64_t *, int32_t *) { return fail(ACGPU_ENODEVICE, "mock: no device entry points"); }
int acgpu_chain_shard_begin(uint64_t, const void *, int64_t, int64_t, void *, uint64_t *, void *) {
    return fail(ACGPU_ENODEVICE, "mock: no device entry points");
}
int acgpu_chain_shard_finish(uint64_t, int32_t, int32_t, void *, void *, int64_t, void *, void *) {
    return fail(ACGPU_ENODEVICE, "mock: no device entry points");
}
// Readable: every feed is buffered and reports nothing; end() reports the whole stream (a legal split of "the records
// that are final so far"), positions as stream offsets.
int acgpu_stream_begin(uint64_t h, uint64_t *s) {
    *s = (uint64_t)(uintptr_t) new MockStream{(Mock *)(uintptr_t)h, {}};
    return ACGPU_OK;
}
int acgpu_stream_set_values_only(uint64_t, int) { return ACGPU_OK; }  // the mock keeps returning positions too (legal: a superset)
int acgpu_stream_feed(uint64_t s, const uint16_t *chars, int32_t n, acgpu_result *out) {
    MockStream *st = (MockStream *)(uintptr_t)s;
    st->chars.insert(st->chars.end(), chars, chars + n);
    out->n = 0;
    out->pos = nullptr;
    out->val = nullptr;
    return ACGPU_OK;
}
int acgpu_stream_end(uint64_t s, acgpu_result *out) {
    MockStream *st = (MockStream *)(uintptr_t)s;
    int rc = ACGPU_OK;
    if (out) rc = fill(st->owner, st->chars.data(), (int32_t)st->chars.size(), out);
    delete st;
    return rc;
}
const char *acgpu_last_error(void) { return g_err.c_str(); }
const char *acgpu_version(void) { return "mock-oracle (tests only)"; }
}
