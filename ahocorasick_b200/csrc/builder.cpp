// Host-side dictionary flattening (see builder.hpp).
#include "builder.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "java_char_tables.h"
#include "trie_insert.hpp"

namespace acgpu {

namespace {

uint16_t g_lower[65536];
std::once_flag g_lower_once;

// ACGPU_BUILD_TIMING=1: phase times of build_automaton on stderr (dictionary-construction throughput)
struct PhaseTimer {
    const bool on = std::getenv("ACGPU_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[acgpu build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

std::string utf8_of(const uint16_t *s, int32_t len) {
    std::string out;
    for (int32_t i = 0; i < len; i++) {
        uint32_t c = s[i];
        if (c < 0x80 && c != 0) {
            out.push_back(static_cast<char>(c));
        } else if (c < 0x800) {
            out.push_back(static_cast<char>(0xC0 | (c >> 6)));
            out.push_back(static_cast<char>(0x80 | (c & 0x3F)));
        } else {
            out.push_back(static_cast<char>(0xE0 | (c >> 12)));
            out.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | (c & 0x3F)));
        }
    }
    return out;
}

}  // namespace

const uint16_t *java_lower_table() {
    std::call_once(g_lower_once, [] { java_fill_lower_table(g_lower); });
    return g_lower;
}

void make_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n, uint8_t *out) {
    std::memset(out, 0, 65536);
    if (mode == 0 || mode == 2) {
        out['-'] = 1;
        out['_'] = 1;
        for (uint32_t c = 0; c < 65536; c++) {
            if (java_is_letter_or_digit(static_cast<uint16_t>(c))) out[c] = 1;
        }
    }
    if (mode == 1) {
        for (int32_t i = 0; i < n; i++) out[chars[i]] = 1;
    } else if (mode == 2) {
        for (int32_t i = 0; i < n; i++) out[chars[i]] = toggles[i] ? 1 : 0;
    }
}

namespace {

constexpr uint64_t kTierRowBytes = 176ull * 1024;  // shared-memory budget for ALL direct-indexed level tables (row layout)

void build_tiers(HostAutomaton &a, const std::vector<uint32_t> &node_parent, const std::vector<uint16_t> &node_cls) {
    TierTables &t = a.tier;
    t.ok = false;
    const int64_t C = a.n_classes;
    if (!a.has_other || a.max_len < 1 || C < 2 || C > 32) return;  // child masks are 32-bit words
    int b = 1;
    while ((1 << b) < C) b++;
    if (static_cast<int64_t>(b) * a.max_len > 60 || a.max_len > 16) return;  // contexts pack into 60 bits, hit masks into 16
    // K = deepest level whose rows (two words per row) fit the shared-memory budget next to the one-word rows of the
    // levels below it
    int K = 0;
    uint64_t entries = 1, lower_bytes = 0;  // entries = C^K = rows of level K+1
    while (K < a.max_len && K < 8) {
        const uint64_t lower_next = lower_bytes + (K >= 1 ? entries / C * 8 : 0);  // level K becomes a lower level (8 bytes per row in the pair layout, 4 in the single one)
        if (lower_next + entries * 8 + 64 > kTierRowBytes) break;
        lower_bytes = lower_next;
        entries *= C;
        K++;
    }
    if (K < 1) return;
    t.C = static_cast<int32_t>(C);
    t.b = b;
    t.K = K;
    uint64_t pw = 1;
    for (int j = 1; j <= K; j++) {
        t.pow_c[j] = static_cast<uint32_t>(pw);  // C^(j-1)
        pw *= C;
    }
    {
        uint32_t roff = 0;
        uint64_t rows = 1;  // C^(j-1)
        for (int j = 1; j <= K; j++) {
            if (j == K) roff = (roff + 1u) & ~1u;  // level-K rows are read as 8-byte pairs
            t.row_off[j] = roff;
            roff += static_cast<uint32_t>(rows * (j == K ? 2 : 1));
            rows *= C;
        }
        t.row_words.assign(roff, 0);
    }
    const int64_t n = a.n_nodes;
    std::vector<uint8_t> depth(n, 0);
    std::vector<uint64_t> packed(n, 0);
    std::vector<uint32_t> radix(n, 0);
    std::vector<uint32_t> kids(n, 0);
    uint64_t n_deep = 0;
    for (int64_t id = 1; id < n; id++) {
        const uint32_t p = node_parent[id];
        const int d = depth[p] + 1;
        depth[id] = static_cast<uint8_t>(d);
        packed[id] = packed[p] | (static_cast<uint64_t>(node_cls[id]) << (b * (d - 1)));
        kids[p] |= 1u << node_cls[id];
        if (d <= K) {
            radix[id] = radix[p] + node_cls[id] * t.pow_c[d];
        } else {
            n_deep++;
        }
    }
    t.n_deep = n_deep;
    if (a.max_len > K) t.kidmask.assign(entries * 2, 0);  // {backward, forward} word per level-K context
    for (int64_t id = 1; id < n; id++) {
        const int d = depth[id];
        const uint32_t inf = a.node_info[id];
        if (d < K) {
            if (inf & kInfoTerminal) {
                t.row_words[t.row_off[d] + radix[id] / C] |= 1u << ((radix[id] % C + 16 - d) & 31);
                // a keyword of length K-1 is also flagged in the level-K row it names (spare bit 31 of the kids word)
                if (d == K - 1 && C <= 31) t.row_words[t.row_off[K] + 2 * radix[id] + 1] |= 1u << 31;
                t.term_levels |= 1u << d;
            }
        } else if (d == K) {
            uint32_t *rw = &t.row_words[t.row_off[K] + 2 * (radix[id] / C)];
            if (inf & kInfoTerminal) rw[0] |= 1u << ((radix[id] % C + 16 - K) & 31);
            if (inf & kInfoHasChildren) rw[1] |= 1u << (radix[id] % C);
            if (inf & kInfoTerminal) t.term_levels |= 1u << d;
            if (!t.kidmask.empty()) t.kidmask[2 * static_cast<size_t>(radix[id])] = kids[id];
        }
    }
    // forward words: position q + 1 asks "does the level-(K+1) node (c[q+1], c[q], ..., c[q+1-K]) exist" = bit c[q+1-K] of
    // the backward word of context (c[q+1], ..., c[q+2-K]); filed under the context of position q, G = (c[q], ..., c[q+1-K]),
    // as bit c[q+1] - so ONE gather at G answers positions q (backward word) and q + 1 (forward word)
    if (!t.kidmask.empty()) {
        const uint64_t top = entries / C;  // C^(K-1)
        for (uint64_t g = 0; g < entries; g++) {
            uint32_t back = t.kidmask[2 * g];
            const uint64_t e = g % C, rest = g / C;  // e = the most recent class of g, rest = the K-1 classes before it
            while (back) {
                const uint32_t a_cls = static_cast<uint32_t>(__builtin_ctz(back));
                back &= back - 1;
                t.kidmask[2 * (rest + a_cls * top) + 1] |= 1u << e;
            }
        }
    }
    // ---- pair rows (k_tier_pair): {fwd, back} per row of every level, see TierTables
    if (C <= 31) {
        uint32_t off = 0;
        uint64_t rows = 1;  // C^(j-1)
        for (int j = 1; j <= K; j++) {
            t.prow_off[j] = off;
            off += static_cast<uint32_t>(rows * 2);
            rows *= C;
        }
        if (C <= 30 && K >= 2) {
            t.prow_off[0] = off;  // compact fwd words of level K-1
            off += static_cast<uint32_t>(t.pow_c[K - 1]);  // C^(K-2) rows
            off = (off + 1u) & ~1u;
            t.pair_low_bit = 1u << ((C + 17 - K) & 31);
        }
        t.prow_words.assign(off, 0);
        t.pair_gate_bit = 1u << ((C + 16 - K) & 31);
        for (int64_t id = 1; id < n; id++) {
            const int d = depth[id];
            if (d > K || !(a.node_info[id] & kInfoTerminal)) continue;
            const uint32_t top = t.pow_c[d];  // C^(d-1)
            const uint32_t bit_f = 1u << ((radix[id] % C + 16 - d) & 31), bit_b = 1u << ((radix[id] / top + 16 - d) & 31);
            t.prow_words[t.prow_off[d] + 2 * static_cast<size_t>(radix[id] / C)] |= bit_f;
            t.prow_words[t.prow_off[d] + 2 * static_cast<size_t>(radix[id] % top) + 1] |= bit_b;
            if (d == K - 1 && t.pair_low_bit) {
                t.prow_words[t.prow_off[K] + 2 * static_cast<size_t>(radix[id])] |= t.pair_low_bit;
                t.prow_words[t.prow_off[0] + radix[id] / C] |= bit_f;
            }
        }
        if (!t.kidmask.empty()) {
            const uint64_t top = entries / C;  // C^(K-1)
            for (uint64_t g = 0; g < entries; g++) {
                if (t.kidmask[2 * g] | t.kidmask[2 * g + 1]) t.prow_words[t.prow_off[K] + 2 * static_cast<size_t>(g % top)] |= t.pair_gate_bit;
            }
        }
    }
    // ---- path-compressed deep table
    // ---- Map values: every keyword, keyed by its packed classes and length
    t.n_vbuckets = 1;
    t.vbuckets.assign(8, 0xFFFFFFFFu);
    if (a.is_map) {
        uint64_t n_keys = 0;
        for (int64_t id = 1; id < n; id++) n_keys += (a.node_info[id] & kInfoTerminal) ? 1 : 0;
        const uint64_t nvb = std::max<uint64_t>(4, (n_keys * 10 + 10) / 11);
        if (nvb > 0x7FFFFFFFull) return;
        t.n_vbuckets = static_cast<uint32_t>(nvb);
        t.vseed = 0xD6E8FEB86659FD93ull;
        t.vbuckets.assign(nvb * 8, 0xFFFFFFFFu);  // empty: key = all ones (no key has length bits 1111 and all classes set: b * max_len <= 60 leaves them apart)
        for (int64_t id = 1; id < n; id++) {
            if (!(a.node_info[id] & kInfoTerminal)) continue;
            const uint64_t key = packed[id] | (static_cast<uint64_t>(depth[id] - 1) << 60);
            const uint32_t hs = value_hash32(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
            uint32_t bk = static_cast<uint32_t>((static_cast<uint64_t>(hs) * nvb) >> 32);
            while (true) {
                uint32_t *e = &t.vbuckets[static_cast<size_t>(bk) * 8];
                int k = (e[0] == 0xFFFFFFFFu && e[1] == 0xFFFFFFFFu) ? 0 : ((e[4] == 0xFFFFFFFFu && e[5] == 0xFFFFFFFFu) ? 1 : -1);
                if (k >= 0) {
                    e[4 * k] = static_cast<uint32_t>(key);
                    e[4 * k + 1] = static_cast<uint32_t>(key >> 32);
                    e[4 * k + 2] = a.node_value[id];
                    e[4 * k + 3] = 0;
                    break;
                }
                bk = bk + 1 == t.n_vbuckets ? 0 : bk + 1;
            }
        }
    }
    if (n_deep == 0) {
        t.n_buckets = 1;
        t.buckets.assign(8, 0);
        t.ok = true;
        return;
    }
    std::vector<uint32_t> child_count(n, 0), only_child(n, 0);
    for (int64_t id = 1; id < n; id++) {
        child_count[node_parent[id]]++;
        only_child[node_parent[id]] = static_cast<uint32_t>(id);
    }
    std::vector<uint8_t> is_head(n, 0);
    for (int64_t id = 1; id < n; id++) {
        if (depth[id] == K + 1 || (depth[id] > K + 1 && child_count[node_parent[id]] >= 2)) is_head[id] = 1;
    }
    struct Head {
        uint64_t key;   // packed context of the head
        uint64_t zw;    // chain | L << 40 | terminal flags << 44
        uint32_t kids;  // child mask of the chain's last node
    };
    std::vector<Head> heads;
    for (int64_t id = 1; id < n; id++) {  // ids ascend from parent to child, so continuation heads are marked in time
        if (depth[id] <= K || !is_head[id]) continue;
        Head h;
        h.key = packed[id];
        uint64_t chain = 0, term = 0;
        uint32_t cur = static_cast<uint32_t>(id);
        int L = 0;
        while (true) {
            if (a.node_info[cur] & kInfoTerminal) term |= 1ull << L;
            if (child_count[cur] != 1 || L == kTierChainMax) break;
            const uint32_t nxt = only_child[cur];
            chain |= static_cast<uint64_t>(node_cls[nxt]) << (b * L);
            L++;
            cur = nxt;
        }
        if (child_count[cur] == 1) is_head[only_child[cur]] = 1;  // chain outgrew the entry: continue in a new head
        h.zw = chain | (static_cast<uint64_t>(L) << 40) | (term << 44);
        h.kids = kids[cur];
        heads.push_back(h);
    }
    t.n_heads = heads.size();
    // two entries per bucket, load factor about 0.55; retry with another seed if two entries on one probe path would
    // share a tag
    uint64_t n_buckets = std::max<uint64_t>(4, (heads.size() * 10 + 10) / 11);
    if (n_buckets > 0x7FFFFFFFull) return;
    t.n_buckets = static_cast<uint32_t>(n_buckets);
    for (int attempt = 0; attempt < 16; attempt++) {
        t.hash_seed = 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(attempt + 1);
        t.buckets.assign(n_buckets * 8, 0);
        bool clash = false;
        for (size_t i = 0; i < heads.size() && !clash; i++) {
            const Head &h = heads[i];
            const uint64_t hs = deep_hash64(h.key, t.hash_seed);
            const uint32_t want = (static_cast<uint32_t>(hs >> 36) << 4) | 8u;
            uint32_t bk = static_cast<uint32_t>(((hs & 0xFFFFFFFFull) * n_buckets) >> 32);
            while (true) {
                uint32_t *e = &t.buckets[static_cast<size_t>(bk) * 8];
                int free_slot = -1;
                for (int k = 0; k < 2; k++) {
                    if (e[4 * k] == 0) {
                        if (free_slot < 0) free_slot = k;
                    } else if ((e[4 * k] & ~7u) == want) {
                        clash = true;
                    }
                }
                if (clash) break;
                if (free_slot >= 0) {
                    e[4 * free_slot] = want;
                    e[4 * free_slot + 1] = h.kids;
                    e[4 * free_slot + 2] = static_cast<uint32_t>(h.zw);
                    e[4 * free_slot + 3] = static_cast<uint32_t>(h.zw >> 32);
                    break;
                }
                bk = bk + 1 == t.n_buckets ? 0 : bk + 1;
            }
        }
        // a later insertion may put an equal tag into a bucket that an earlier key's probe path crosses only if
        // that bucket was full when the earlier key passed it - and full buckets never change, so checking at
        // insertion time against every bucket passed (done above) is sufficient.
        if (!clash) {
            t.ok = true;
            return;
        }
    }
}

// Map values of the wide path (HostAutomaton::wide_vals)
void build_wide_values(HostAutomaton &a, const std::vector<uint32_t> &node_parent, const std::vector<uint16_t> &node_cls) {
    a.wide_vals.clear();
    a.wide_n_vbuckets = 0;
    if (!a.is_map) return;
    struct Key { uint64_t h; uint32_t len, value; };
    std::vector<Key> keys;
    std::vector<uint16_t> tmp;
    for (int64_t id = 1; id < a.n_nodes; id++) {
        if (!(a.node_info[id] & kInfoTerminal)) continue;
        tmp.clear();
        for (uint32_t cur = static_cast<uint32_t>(id); cur != 0; cur = node_parent[cur]) tmp.push_back(node_cls[cur]);
        WideValHash h;  // the trie is over reversed keywords: root -> node spells the keyword from its last char back
        for (size_t i = tmp.size(); i-- > 0;) h.add(tmp[i]);
        keys.push_back(Key{h.finish(static_cast<uint32_t>(tmp.size())), static_cast<uint32_t>(tmp.size()), a.node_value[id]});
    }
    {
        std::vector<std::pair<uint64_t, uint32_t>> seen;
        seen.reserve(keys.size());
        for (const Key &k : keys) seen.emplace_back(k.h, k.len);
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) return;  // (hash, length) collision: no table, the emit kernel walks
    }
    const uint64_t nb = std::max<uint64_t>(4, keys.size());  // two entries per bucket: load factor 0.5
    if (nb > 0x7FFFFFFFull) return;
    a.wide_n_vbuckets = static_cast<uint32_t>(nb);
    a.wide_vals.assign(static_cast<size_t>(nb) * 8, 0u);
    for (const Key &k : keys) {
        uint32_t bk = static_cast<uint32_t>((static_cast<uint64_t>(static_cast<uint32_t>(k.h >> 32)) * nb) >> 32);
        while (true) {
            uint32_t *e = &a.wide_vals[static_cast<size_t>(bk) * 8];
            const int slot = e[2] == 0u ? 0 : (e[6] == 0u ? 1 : -1);
            if (slot >= 0) {
                e[slot * 4 + 0] = static_cast<uint32_t>(k.h);
                e[slot * 4 + 1] = static_cast<uint32_t>(k.h >> 32);
                e[slot * 4 + 2] = k.len | 0x80000000u;
                e[slot * 4 + 3] = k.value;
                break;
            }
            bk = bk + 1u == a.wide_n_vbuckets ? 0u : bk + 1u;
        }
    }
}

// AhoCorasick family outside the tier envelope: class-pair table for levels 1 and 2 of the anchored walk (kernel_wide.cuh)
constexpr int kWideMaxLenHost = 32, kWidePairMaxHost = 64;
void build_wide(HostAutomaton &a, const std::vector<uint32_t> &node_parent, const std::vector<uint16_t> &node_cls) {
    a.wide_ok = false;
    a.wide_pair.clear();
    if (!a.has_other || a.max_len < 1 || a.max_len > kWideMaxLenHost) return;
    a.wide_ok = true;
    const int64_t C = a.n_classes;
    if (C > kWidePairMaxHost) {  // the walk starts at the root table instead
        build_wide_values(a, node_parent, node_cls);
        return;
    }
    a.wide_pair.assign(static_cast<size_t>(C * C * 2), 0);
    for (int64_t c0 = 0; c0 < C; c0++) {
        const RootEdge &r = a.root[c0];
        for (int64_t c1 = 0; c1 < C; c1++) {
            uint32_t *e = &a.wide_pair[static_cast<size_t>(c0 * C + c1) * 2];
            e[0] = kNone;
            e[1] = r.child == kNone ? 0u : ((r.info & 0xFFu) << 8) | (1u << 16);
        }
    }
    for (int64_t id = 1; id < a.n_nodes; id++) {
        const uint32_t p = node_parent[id];
        if (p == 0 || node_parent[p] != 0) continue;  // level-2 nodes only
        uint32_t *e = &a.wide_pair[static_cast<size_t>(static_cast<int64_t>(node_cls[p]) * C + node_cls[id]) * 2];
        e[0] = static_cast<uint32_t>(id);
        e[1] |= a.node_info[id] & 0xFFu;
    }
    // ---- path-compressed edges below level 2 (HostAutomaton::wide_chain, wide_pair16)
    const int64_t N = a.n_nodes;
    std::vector<uint32_t> first(static_cast<size_t>(N) + 1, 0), kids(static_cast<size_t>(N > 1 ? N - 1 : 0));
    for (int64_t id = 1; id < N; id++) ++first[node_parent[id] + 1];
    for (int64_t i = 0; i < N; i++) first[i + 1] += first[i];
    {
        std::vector<uint32_t> at(first.begin(), first.end() - 1);
        for (int64_t id = 1; id < N; id++) kids[at[node_parent[id]]++] = static_cast<uint32_t>(id);
        for (int64_t n = 0; n < N; n++)  // children in class order
            std::sort(kids.begin() + first[n], kids.begin() + first[n + 1], [&](uint32_t x, uint32_t y) { return node_cls[x] < node_cls[y]; });
    }
    auto n_kids = [&](uint32_t n) { return first[n + 1] - first[n]; };
    auto kid_mask = [&](uint32_t n) {
        uint64_t m = 0;
        for (uint32_t k = first[n]; k < first[n + 1]; k++) m |= 1ull << node_cls[kids[k]];
        return m;
    };
    a.wide_chain.clear();
    // junctions in breadth-first order; a junction's children take the next n_kids entries.  entry_of[i] = the child node
    // that entry i describes (filled in below).
    std::vector<uint32_t> entry_child;
    a.wide_pair16.assign(static_cast<size_t>(C * C * 4), 0);
    for (int64_t c0 = 0; c0 < C; c0++)
        for (int64_t c1 = 0; c1 < C; c1++) a.wide_pair16[static_cast<size_t>(c0 * C + c1) * 4 + 1] = a.wide_pair[static_cast<size_t>(c0 * C + c1) * 2 + 1];
    std::vector<std::pair<uint32_t, uint32_t>> todo;  // (junction, first entry of its children)
    auto reserve_children = [&](uint32_t j) {
        const uint32_t at = static_cast<uint32_t>(entry_child.size());
        for (uint32_t k = first[j]; k < first[j + 1]; k++) entry_child.push_back(kids[k]);
        todo.emplace_back(j, at);
        return at;
    };
    for (int64_t id = 1; id < N; id++) {
        const uint32_t p2 = node_parent[id];
        if (p2 == 0 || node_parent[p2] != 0) continue;  // level-2 nodes only
        uint32_t *e = &a.wide_pair16[static_cast<size_t>(static_cast<int64_t>(node_cls[p2]) * C + node_cls[id]) * 4];
        e[1] |= (a.node_info[id] & 0xFFu) | 1u << 17;
        const uint64_t m = kid_mask(static_cast<uint32_t>(id));
        e[2] = static_cast<uint32_t>(m);
        e[3] = static_cast<uint32_t>(m >> 32);
        e[0] = m ? reserve_children(static_cast<uint32_t>(id)) : 0u;
    }
    for (size_t t = 0; t < todo.size(); t++) {
        const uint32_t j = todo[t].first, at = todo[t].second;
        for (uint32_t k = 0; k < n_kids(j); k++) {
            const uint32_t child = entry_child[at + k];
            uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t n = child, L = 0, term = (a.node_info[child] & kInfoTerminal) ? 1u : 0u;
            while (n_kids(n) == 1 && L < static_cast<uint32_t>(kWideChainMax)) {
                n = kids[first[n]];
                w[4 + ((7u - L) >> 2)] |= static_cast<uint32_t>(node_cls[n]) << (((7u - L) & 3u) * 8u);  // step k in byte 7 - k
                ++L;
                if (a.node_info[n] & kInfoTerminal) term |= 1u << L;
            }
            const uint64_t m = kid_mask(n);
            w[0] = static_cast<uint32_t>(m);
            w[1] = static_cast<uint32_t>(m >> 32);
            w[2] = m ? reserve_children(n) : 0u;
            w[3] = L | term << 4;
            w[6] = n;
            const size_t o = static_cast<size_t>(at + k) * 8;
            if (a.wide_chain.size() < o + 8) a.wide_chain.resize(std::max(o + 8, a.wide_chain.size() * 2), 0);
            std::memcpy(&a.wide_chain[o], w, sizeof w);
        }
    }
    a.wide_chain.resize(entry_child.size() * 8);
    build_wide_values(a, node_parent, node_cls);
}

// WholeWord hash tables: one entry per distinct (trimmed, folded) keyword = per terminal node of the forward trie
void build_ww(HostAutomaton &a, const std::vector<uint8_t> &wc, const std::vector<uint32_t> &node_parent,
              const std::vector<uint16_t> &node_cls) {
    WwTables &w = a.ww;
    w.ok = false;
    if (!a.has_other || a.n_classes >= 32768 || a.max_len > kWwMaxLen) return;
    w.wcls.resize(65536);
    for (uint32_t c = 0; c < 65536; c++) w.wcls[c] = static_cast<uint16_t>(a.cls[c] | (wc[c] ? 0x8000u : 0u));
    uint64_t n_keys = 0;
    for (int64_t id = 1; id < a.n_nodes; id++) n_keys += a.node_info[id] & kInfoTerminal;
    {
        const char *gen = std::getenv("ACGPU_WW_GEN");  // ACGPU_WW_GEN=2: the generation-2 kernel and its hash (A/B runs)
        w.poly = !(gen && gen[0] == '2');
    }
    // two entries per bucket: load factor 0.25 (most probes are misses); generation 3 probes 32 keys in lockstep, so one
    // long probe path stalls a whole batch: 0.125
    const uint64_t nb = std::max<uint64_t>(4, n_keys * (w.poly ? 4 : 2));
    if (nb > 0x7FFFFFFFull) return;
    w.n_buckets = static_cast<uint32_t>(nb);
    w.buckets.assign(nb * 8, 0xFFFFFFFFu);
    if (w.poly && n_keys > 0 && n_keys <= 400000) {
        uint64_t bits = 4096;
        // 64 KB at most; 32 KB when keywords of 32 chars or more need the two-row ring (kernel_ww3.cuh: the CTA stays under the 164 KB carve-out)
        while (bits < n_keys * 8 && bits < (a.max_len < 32 ? 512 : 256) * 1024) bits <<= 1;
        w.bloom_bits = static_cast<uint32_t>(bits);
        w.bloom.assign(bits / 32, 0u);
    }
    std::vector<uint16_t> tmp;
    for (int64_t id = 1; id < a.n_nodes; id++) {
        if (!(a.node_info[id] & kInfoTerminal)) continue;
        tmp.clear();
        for (uint32_t cur = static_cast<uint32_t>(id); cur != 0; cur = node_parent[cur]) tmp.push_back(node_cls[cur]);
        std::reverse(tmp.begin(), tmp.end());
        uint32_t tag, spread;
        if (w.poly) {
            uint32_t poly = 0;
            for (uint16_t c : tmp) poly = poly * kWwPolyB + ww_poly_digit(c);
            tag = ww_poly_key(poly, static_cast<uint32_t>(tmp.size()));
            spread = ww_poly_spread(tag);
            if (w.bloom_bits) {
                int lb = 0;
                while ((1u << lb) < w.bloom_bits) lb++;
                const uint32_t bit = tag >> (32 - lb);
                w.bloom[bit >> 5] |= 1u << (bit & 31);
            }
        } else {
            WwHash h;
            for (size_t i = 0; i < tmp.size(); i += 2)
                h.add_pair(static_cast<uint32_t>(tmp[i]) | (i + 1 < tmp.size() ? static_cast<uint32_t>(tmp[i + 1]) << 16 : 0u));
            h.finish(static_cast<uint32_t>(tmp.size()));
            tag = h.h1;
            spread = h.spread();
        }
        if (w.pool.size() + tmp.size() > 0xFFFFFFF0ull) return;
        const uint32_t off = static_cast<uint32_t>(w.pool.size());
        w.pool.insert(w.pool.end(), tmp.begin(), tmp.end());
        uint32_t bk = static_cast<uint32_t>((static_cast<uint64_t>(spread) * nb) >> 32);
        while (true) {
            uint32_t *e = &w.buckets[static_cast<size_t>(bk) * 8];
            const int k = e[2] == 0xFFFFFFFFu ? 0 : (e[6] == 0xFFFFFFFFu ? 1 : -1);
            if (k >= 0) {
                e[4 * k] = tag;
                e[4 * k + 1] = static_cast<uint32_t>(tmp.size());
                e[4 * k + 2] = off;
                e[4 * k + 3] = a.node_value[id];
                break;
            }
            bk = bk + 1 == w.n_buckets ? 0 : bk + 1;
        }
    }
    if (w.pool.empty()) w.pool.push_back(0);
    w.ok = true;
}

}  // namespace

HostAutomaton build_automaton(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                              int64_t n_keywords, int64_t n_values, bool case_sensitive,
                              const uint8_t *word_chars) {
    if (family < 0 || family > 4) throw std::invalid_argument("unknown matcher family");
    const bool word_family = family == 3 || family == 4;  // WholeWord, WholeWordLongest
    const uint16_t *lower = java_lower_table();
    HostAutomaton a;
    PhaseTimer timer;
    a.family = family;
    a.is_map = n_values >= 0;
    a.case_sensitive = case_sensitive;
    a.reversed = (family == 0);

    // Maps zip keywords with values and stop at the shorter (AhoCorasickMap.java:32).
    int64_t n = n_keywords;
    if (a.is_map && n_values < n) n = n_values;

    std::vector<uint8_t> wc;
    if (word_family) {
        wc.resize(65536);
        if (word_chars) {
            std::memcpy(wc.data(), word_chars, 65536);
        } else {
            make_word_chars(0, nullptr, nullptr, 0, wc.data());
        }
        a.wordbits.assign(2048, 0);
        for (uint32_t c = 0; c < 65536; c++) {
            if (wc[c]) a.wordbits[c >> 5] |= 1u << (c & 31);
        }
        if (!case_sensitive) {
            // Quirk Q7 (WholeWordMatchSet.java:96-101 vs :113,118): the reference tests the lower-cased
            // char in the first word-char check but the raw char when scrolling.  The two views agree
            // (and the "maximal run of word chars" formulation is exact) iff the table is closed under
            // toLowerCase, which holds for the default table and every toggle of caseless chars.
            bool closed = true;
            for (uint32_t c = 0; c < 65536 && closed; c++) closed = wc[c] == wc[lower[c]];
            if (!closed) {
                // the reference's loop is followed literally (kernel_wwlit.cuh); it needs both views of the table
                a.ww_literal = true;
                a.wordbits_fold.assign(2048, 0);
                for (uint32_t c = 0; c < 65536; c++) {
                    if (wc[lower[c]]) a.wordbits_fold[c >> 5] |= 1u << (c & 31);
                }
            }
        }
    }

    // ---- effective keywords: (begin, len, entry index) into `chars`, after trim / skip rules
    std::vector<KwRef> kws;
    kws.reserve(static_cast<size_t>(n));
    int32_t longest = 0;
    std::vector<uint8_t> used(65536, 0);
    for (int64_t k = 0; k < n; k++) {
        if (is_null && is_null[k]) continue;
        int64_t b = offsets[k];
        int32_t len = static_cast<int32_t>(offsets[k + 1] - offsets[k]);
        if (word_family) {
            // WordCharacters.trim (WordCharacters.java:41-62) then validation (WholeWordMatchSet.java:147-153);
            // WholeWordLongest trims but accepts inner non-word chars (WholeWordLongestMatchSet.java:190-206)
            int32_t ws = 0, we = len;
            for (int32_t i = 0; i < len; i++) {
                if (wc[chars[b + i]]) {
                    ws = i;
                    break;
                }
            }
            for (int32_t i = len - 1; i >= 0; i--) {
                if (wc[chars[b + i]]) {
                    we = i + 1;
                    break;
                }
            }
            b += ws;
            len = we - ws;
            for (int32_t i = 0; i < len; i++) {
                if (family == 4 && !wc[chars[b + i]]) a.ww_plain = false;
                if (family == 3 && !wc[chars[b + i]]) {
                    throw IllegalArgument(utf8_of(chars + b, len) + " contains non-word characters.");
                }
            }
        }
        if (len > longest) longest = len;
        if (len <= 0) continue;
        kws.push_back(KwRef{b, len, k});
        for (int32_t i = 0; i < len; i++) {
            uint16_t c = chars[b + i];
            used[case_sensitive ? c : lower[c]] = 1;
        }
    }
    a.max_len = longest;
    a.char_buffer_size = longest > 2048 ? longest * 2 : 4096;
    a.n_keywords_effective = static_cast<int64_t>(kws.size());

    timer.lap("keywords: trim, validate");
    // ---- character classes
    uint32_t distinct = 0;
    for (uint32_t c = 0; c < 65536; c++) distinct += used[c];
    a.has_other = distinct < 65536;
    std::vector<uint16_t> class_of(65536, 0);
    uint32_t next_class = a.has_other ? 1 : 0;
    for (uint32_t c = 0; c < 65536; c++) {
        if (used[c]) class_of[c] = static_cast<uint16_t>(next_class++);
    }
    a.n_classes = static_cast<int32_t>(next_class < 1 ? 1 : next_class);
    a.cls.resize(65536);
    for (uint32_t c = 0; c < 65536; c++) {
        uint16_t f = case_sensitive ? static_cast<uint16_t>(c) : lower[c];
        a.cls[c] = class_of[f];  // 0 ("other") when f is in no keyword and has_other
    }

    // ---- trie over class strings (trie_insert.hpp).  Dictionaries of 50 000+ keywords are inserted concurrently, one
    // shard per first class; both ways give the same arrays.  ACGPU_BUILDER=serial|sharded forces one (tests).
    const bool first_wins = (family == 2);  // ShortestMatchMap.java:44-54
    const TrieInsertParams tp{chars, a.cls.data(), a.reversed, a.is_map, first_wins, longest};
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const char *mode = std::getenv("ACGPU_BUILDER");
    const bool sharded = mode ? std::strcmp(mode, "sharded") == 0 : (kws.size() >= 50000 && hw > 1);
    TrieArrays trie = sharded ? insert_sharded(kws, tp, a.n_classes, mode ? std::max(hw, 4u) : hw) : insert_serial(kws, tp);
    timer.lap(sharded ? "trie insert (sharded)" : "trie insert (serial)");
    a.n_nodes = static_cast<int64_t>(trie.info.size());
    a.node_info.swap(trie.info);
    a.node_value.swap(trie.value);
    a.depth_count.swap(trie.depth_count);
    const std::vector<uint32_t> &node_parent = trie.parent;
    const std::vector<uint16_t> &node_cls = trie.cls;

    // ---- device tables: direct root table + hashed deeper edges, inserted in child-id order (canonical layout)
    // info word of an edge = the child's flags | its child signature << 2 (bit c % 30: some child has class c; trie_step_sig)
    std::vector<uint32_t> info32(static_cast<size_t>(a.n_nodes), 0u);
    for (int64_t id = 1; id < a.n_nodes; id++) {
        info32[id] |= a.node_info[id];
        info32[node_parent[id]] |= 1u << (2u + node_cls[id] % 30u);
    }
    a.root.assign(static_cast<size_t>(a.n_classes), RootEdge{kNone, 0});
    uint64_t deep_edges = 0;
    for (int64_t id = 1; id < a.n_nodes; id++) {
        if (node_parent[id] == 0) {
            a.root[node_cls[id]] = RootEdge{static_cast<uint32_t>(id), info32[id]};
        } else {
            ++deep_edges;
        }
    }
    uint64_t cap = 16;
    while (cap < deep_edges * 2) cap <<= 1;
    if (cap > (1ull << 31)) throw std::length_error("dictionary too large for the edge table");
    a.edges.assign(static_cast<size_t>(cap), Edge{kNone, 0, 0, 0});
    a.edge_mask = static_cast<uint32_t>(cap - 1);
    for (int64_t id = 1; id < a.n_nodes; id++) {
        const uint32_t parent = node_parent[id];
        if (parent == 0) continue;
        uint32_t i = edge_hash(parent, node_cls[id]) & a.edge_mask;
        while (a.edges[i].parent != kNone) i = (i + 1) & a.edge_mask;
        a.edges[i] = Edge{parent, node_cls[id], static_cast<uint32_t>(id), info32[id]};
    }
    timer.lap("edge table");
    if (family != 4) build_tiers(a, node_parent, node_cls);
    if (family == 0 && !a.tier.ok) build_wide(a, node_parent, node_cls);
    timer.lap("tier tables");
    // WholeWordLongest with a dictionary whose (trimmed) keywords hold no non-word char: a walk can never leave its word
    // (there is no transition on a non-word char) and a keyword followed by a non-word char is the whole word, so the
    // family coincides with WholeWord and takes its hash path.
    if (!a.ww_literal && (family == 3 || (family == 4 && a.ww_plain))) build_ww(a, wc, node_parent, node_cls);
    if (family == 4 && !a.ww_plain && !a.ww_literal && a.has_other && a.n_classes < 32768 && a.max_len <= kWwMaxLen - 1) {
        a.wwl_wcls.resize(65536);
        for (uint32_t c = 0; c < 65536; c++) a.wwl_wcls[c] = static_cast<uint16_t>(a.cls[c] | (wc[c] ? 0x8000u : 0u));
    }
    timer.lap("whole-word hash");
    return a;
}

uint64_t automaton_fingerprint(const HostAutomaton &a) {
    uint64_t h = 0xCBF29CE484222325ull;
    auto bytes = [&](const void *p, size_t n) {  // 8 bytes per step (FNV-1a style over words), then the tail
        const unsigned char *b = static_cast<const unsigned char *>(p);
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            std::memcpy(&w, b + i, 8);
            h = (h ^ w) * 0x100000001B3ull;
            h ^= h >> 29;
        }
        for (; i < n; i++) h = (h ^ b[i]) * 0x100000001B3ull;
    };
    auto num = [&](uint64_t v) { bytes(&v, sizeof v); };
    auto vec = [&](const auto &v) {
        num(v.size());
        if (!v.empty()) bytes(v.data(), v.size() * sizeof(v[0]));
    };
    num((uint64_t)a.family); num(a.is_map); num(a.case_sensitive); num(a.reversed); num(a.has_other);
    num((uint64_t)a.max_len); num((uint64_t)a.n_classes); num((uint64_t)a.char_buffer_size); num((uint64_t)a.n_nodes);
    num((uint64_t)a.n_keywords_effective); num(a.edge_mask); num(a.ww_plain);
    vec(a.cls); vec(a.wordbits); vec(a.node_value); vec(a.node_info); vec(a.depth_count);
    num(a.root.size());
    for (const RootEdge &e : a.root) { num(e.child); num(e.info); }
    num(a.edges.size());
    for (const Edge &e : a.edges) { num(e.parent); num(e.cls); num(e.child); num(e.info); }
    const TierTables &t = a.tier;
    num(t.ok); num((uint64_t)t.C); num((uint64_t)t.b); num((uint64_t)t.K); num(t.term_levels);
    bytes(t.pow_c, sizeof t.pow_c); bytes(t.row_off, sizeof t.row_off);
    vec(t.row_words); vec(t.prow_words); bytes(t.prow_off, sizeof t.prow_off); num(t.pair_gate_bit); num(t.pair_low_bit); vec(t.kidmask); vec(t.buckets); num(t.n_buckets); num(t.hash_seed); num(t.n_deep); num(t.n_heads);
    vec(t.vbuckets); num(t.n_vbuckets); num(t.vseed);
    if (a.wide_ok) { num(a.wide_ok); vec(a.wide_pair); vec(a.wide_chain); vec(a.wide_pair16); vec(a.wide_vals); num(a.wide_n_vbuckets); }
    if (a.ww_literal) { num(a.ww_literal); vec(a.wordbits_fold); }
    vec(a.wwl_wcls);
    num(a.ww.ok); vec(a.ww.wcls); vec(a.ww.buckets); num(a.ww.n_buckets); vec(a.ww.pool); num(a.ww.poly); vec(a.ww.bloom);
    return h;
}

}  // namespace acgpu
