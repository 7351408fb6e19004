// Keyword insertion into the anchored trie (host side, part of the dictionary flattening).
//
// The reference inserts keywords one by one into an object graph (AhoCorasickSet.java:24-60).  Here the trie is a set
// of flat per-node arrays (parent, class, info, value) whose node ids are the creation order of a sequential insert.
// Two implementations produce EXACTLY the same arrays:
//   insert_serial   one growing (parent, class) -> child hash map, keywords in order;
//   insert_sharded  keywords are dealt into shards by the class of their first trie step (sub-tries of different
//                   first classes share no node but the root), the shards are built concurrently with private,
//                   pre-sized hash maps, and the sequential numbering is reconstructed afterwards: a node's id is
//                   1 + (nodes created by earlier keywords) + (its rank among the nodes its own keyword created),
//                   which only needs a prefix sum over per-keyword creation counts.
// tests/test_host_cpu.py compares the fingerprints of both (acgpu_build_fingerprint).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <exception>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

#include "builder.hpp"

namespace acgpu {

struct KwRef {  // one effective keyword: chars[begin, begin + len), entry = index in the caller's Iterable
    int64_t begin;
    int32_t len;
    int64_t entry;
};

struct TrieArrays {
    std::vector<uint8_t> info;        // per node: kInfoTerminal | kInfoHasChildren
    std::vector<uint32_t> value;      // per node: entry index of the winning keyword (Maps), else kNone
    std::vector<uint32_t> parent;     // per node (root: 0)
    std::vector<uint16_t> cls;        // per node: class of the edge from its parent (root: 0)
    std::vector<uint32_t> depth_count;  // nodes per depth, [0 .. longest]
};

struct TrieInsertParams {
    const uint16_t *chars;
    const uint16_t *cls_of;   // 65536: code unit -> class
    bool reversed;            // AhoCorasick family: the trie spells keywords right to left
    bool is_map;
    bool first_wins;          // ShortestMatchMap.java:44-54: the first duplicate keeps its value
    int32_t longest;
};

namespace trie_detail {

// open addressing (parent, class) -> child, 12-byte slots, never grows (callers size it for the worst case)
struct FixedEdgeMap {
    struct Slot {
        uint32_t parent, cls, child;
    };
    std::vector<Slot> slots;
    uint32_t mask = 0;
    static uint64_t capacity_for(uint64_t max_edges) {
        uint64_t cap = 4;
        while (cap < max_edges * 2) cap <<= 1;
        if (cap > (1ull << 31)) throw std::length_error("dictionary too large for the edge table");
        return cap;
    }
    FixedEdgeMap() = default;
    explicit FixedEdgeMap(uint64_t max_edges) { reset(max_edges); }
    // empty table for up to max_edges edges; keeps (and re-uses) the memory of earlier, larger tables - first-touch page
    // faults, not hashing, dominate a cold build
    void reset(uint64_t max_edges) {
        const uint64_t cap = capacity_for(max_edges);
        if (slots.size() < cap) {
            slots.assign((size_t)cap, Slot{kNone, 0, 0});
        } else {
            std::fill(slots.begin(), slots.begin() + (size_t)cap, Slot{kNone, 0, 0});
        }
        mask = (uint32_t)(cap - 1);
    }
    uint32_t get_or_add(uint32_t parent, uint32_t c, uint32_t next, bool &created) {
        uint32_t i = edge_hash(parent, c) & mask;
        while (slots[i].parent != kNone) {
            if (slots[i].parent == parent && slots[i].cls == c) {
                created = false;
                return slots[i].child;
            }
            i = (i + 1) & mask;
        }
        slots[i] = Slot{parent, c, next};
        created = true;
        return next;
    }
};

}  // namespace trie_detail

inline TrieArrays insert_serial(const std::vector<KwRef> &kws, const TrieInsertParams &P) {
    TrieArrays t;
    uint64_t total = 0;
    for (const KwRef &k : kws) total += (uint64_t)k.len;
    trie_detail::FixedEdgeMap map(total);
    t.info.assign(1, 0);
    t.value.assign(1, kNone);
    t.parent.assign(1, 0);
    t.cls.assign(1, 0);
    t.depth_count.assign((size_t)P.longest + 1, 0);
    t.depth_count[0] = 1;
    uint32_t next_node = 1;
    for (const KwRef &kw : kws) {
        uint32_t node = 0;
        for (int32_t i = 0; i < kw.len; i++) {
            const uint32_t c = P.cls_of[P.chars[kw.begin + (P.reversed ? (kw.len - 1 - i) : i)]];
            bool created = false;
            const uint32_t child = map.get_or_add(node, c, next_node, created);
            if (created) {
                if (next_node == kNone - 1) throw std::length_error("dictionary too large (node ids exceed 32 bits)");
                ++next_node;
                t.info.push_back(0);
                t.value.push_back(kNone);
                t.parent.push_back(node);
                t.cls.push_back((uint16_t)c);
                t.info[node] |= kInfoHasChildren;
                t.depth_count[(size_t)i + 1]++;
            }
            node = child;
        }
        if (!(P.first_wins && (t.info[node] & kInfoTerminal))) t.value[node] = P.is_map ? (uint32_t)kw.entry : kNone;
        t.info[node] |= kInfoTerminal;
    }
    return t;
}

inline TrieArrays insert_sharded(const std::vector<KwRef> &kws, const TrieInsertParams &P, int n_classes, unsigned n_threads) {
    const size_t n = kws.size();
    auto first_class = [&](const KwRef &k) { return (uint32_t)P.cls_of[P.chars[k.begin + (P.reversed ? k.len - 1 : 0)]]; };

    // ---- deal keywords into shards (counting sort by first class keeps the caller's order inside a shard)
    std::vector<uint64_t> shard_begin((size_t)n_classes + 1, 0);
    std::vector<uint64_t> shard_chars((size_t)n_classes, 0);  // upper bound of the shard's edges below its depth-1 node
    for (const KwRef &k : kws) {
        const uint32_t c = first_class(k);
        shard_begin[c + 1]++;
        shard_chars[c] += (uint64_t)k.len - 1;
    }
    for (int c = 0; c < n_classes; c++) shard_begin[c + 1] += shard_begin[c];
    std::vector<uint32_t> order(n);  // indices into kws, grouped by shard
    {
        std::vector<uint64_t> at(shard_begin.begin(), shard_begin.end() - 1);
        for (size_t g = 0; g < n; g++) order[at[first_class(kws[g])]++] = (uint32_t)g;
    }
    std::vector<uint32_t> work;  // non-empty shards, largest first
    for (int c = 0; c < n_classes; c++)
        if (shard_begin[c + 1] > shard_begin[c]) work.push_back((uint32_t)c);
    std::sort(work.begin(), work.end(), [&](uint32_t a, uint32_t b) { return shard_chars[a] > shard_chars[b]; });

    // ---- per-shard build.  Local node 0 = the shard's depth-1 node (child of the root on the shard's class).
    struct Shard {
        std::vector<uint32_t> parent;    // local parent; kNone = the root
        std::vector<uint16_t> cls;
        std::vector<uint8_t> info;
        std::vector<uint32_t> value;
        std::vector<uint32_t> creator;   // index into kws of the keyword that created the node
        std::vector<uint32_t> rank;      // its rank among the nodes that keyword created
        std::vector<uint32_t> depth_count;
    };
    std::vector<Shard> shards(work.size());
    std::vector<uint32_t> created_by(n, 0);  // nodes created per keyword (each keyword belongs to one shard)
    std::atomic<size_t> next_work{0};
    std::exception_ptr error;  // first exception of any worker, rethrown on the calling thread
    std::mutex error_mutex;
    auto build_shard = [&](size_t w, trie_detail::FixedEdgeMap &map) {
        const uint32_t c0 = work[w];
        Shard &S = shards[w];
        map.reset(shard_chars[c0]);
        const size_t guess = (size_t)(shard_chars[c0] / 2 + 16);  // nodes: usually about half the shard's chars
        S.parent.reserve(guess); S.cls.reserve(guess); S.info.reserve(guess);
        S.value.reserve(guess); S.creator.reserve(guess); S.rank.reserve(guess);
        S.depth_count.assign((size_t)P.longest + 1, 0);
        S.parent.push_back(kNone);
        S.cls.push_back((uint16_t)c0);
        S.info.push_back(0);
        S.value.push_back(kNone);
        S.creator.push_back(order[shard_begin[c0]]);
        S.rank.push_back(0);
        S.depth_count[1] = 1;
        created_by[order[shard_begin[c0]]] = 1;
        uint32_t next_local = 1;
        for (uint64_t o = shard_begin[c0]; o < shard_begin[c0 + 1]; o++) {
            const uint32_t g = order[o];
            const KwRef &kw = kws[g];
            uint32_t node = 0;
            for (int32_t i = 1; i < kw.len; i++) {
                const uint32_t c = P.cls_of[P.chars[kw.begin + (P.reversed ? (kw.len - 1 - i) : i)]];
                bool created = false;
                const uint32_t child = map.get_or_add(node, c, next_local, created);
                if (created) {
                    ++next_local;
                    S.parent.push_back(node);
                    S.cls.push_back((uint16_t)c);
                    S.info.push_back(0);
                    S.value.push_back(kNone);
                    S.creator.push_back(g);
                    S.rank.push_back(created_by[g]++);
                    S.info[node] |= kInfoHasChildren;
                    S.depth_count[(size_t)i + 1]++;
                }
                node = child;
            }
            if (!(P.first_wins && (S.info[node] & kInfoTerminal))) S.value[node] = P.is_map ? (uint32_t)kw.entry : kNone;
            S.info[node] |= kInfoTerminal;
        }
    };
    auto worker = [&](auto &&job) {
        try {
            trie_detail::FixedEdgeMap map;  // one per thread, re-used from shard to shard (largest shard first)
            for (size_t w; (w = next_work.fetch_add(1)) < work.size();) job(w, map);
        } catch (...) {
            std::lock_guard<std::mutex> lock(error_mutex);
            if (!error) error = std::current_exception();
            next_work = work.size();
        }
    };
    auto run_parallel = [&](auto &&job) {
        next_work = 0;
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(n_threads, work.size()));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; t++) pool.emplace_back([&] { worker(job); });
        worker(job);
        for (std::thread &th : pool) th.join();
        if (error) std::rethrow_exception(error);
    };
    run_parallel(build_shard);

    // ---- sequential numbering: id = 1 + nodes created by earlier keywords + rank inside the own keyword
    std::vector<uint64_t> base(n + 1, 0);
    for (size_t g = 0; g < n; g++) base[g + 1] = base[g] + created_by[g];
    const uint64_t n_nodes = 1 + base[n];
    if (n_nodes >= (uint64_t)kNone) throw std::length_error("dictionary too large (node ids exceed 32 bits)");

    TrieArrays t;
    t.info.assign((size_t)n_nodes, 0);
    t.value.assign((size_t)n_nodes, kNone);
    t.parent.assign((size_t)n_nodes, 0);
    t.cls.assign((size_t)n_nodes, 0);
    t.depth_count.assign((size_t)P.longest + 1, 0);
    t.depth_count[0] = 1;
    if (n) t.info[0] = kInfoHasChildren;
    auto scatter = [&](size_t w, trie_detail::FixedEdgeMap &) {
        const Shard &S = shards[w];
        const size_t m = S.parent.size();
        std::vector<uint32_t> gid(m);
        for (size_t l = 0; l < m; l++) gid[l] = (uint32_t)(1 + base[S.creator[l]] + S.rank[l]);
        for (size_t l = 0; l < m; l++) {
            const uint32_t id = gid[l];
            t.info[id] = S.info[l];
            t.value[id] = S.value[l];
            t.parent[id] = S.parent[l] == kNone ? 0u : gid[S.parent[l]];
            t.cls[id] = S.cls[l];
        }
    };
    run_parallel(scatter);
    for (const Shard &S : shards)
        for (size_t d = 1; d < S.depth_count.size(); d++) t.depth_count[d] += S.depth_count[d];
    return t;
}

}  // namespace acgpu
