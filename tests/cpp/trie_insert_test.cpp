// insert_sharded (concurrent, per-first-class shards) must produce exactly the arrays of insert_serial
// (ahocorasick_b200/csrc/trie_insert.hpp): same node numbering, parents, classes, flags, values, depth histogram.
// Host-only; run by tests/test_host_cpu.py.  usage: trie_insert_test [n_threads] [--time N_KEYWORDS]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../ahocorasick_b200/csrc/trie_insert.hpp"

using namespace acgpu;

static bool same(const TrieArrays &a, const TrieArrays &b) {
    return a.info == b.info && a.value == b.value && a.parent == b.parent && a.cls == b.cls && a.depth_count == b.depth_count;
}

int main(int argc, char **argv) {
    unsigned threads = argc > 1 ? (unsigned)std::atoi(argv[1]) : 4;
    std::mt19937_64 rng(12345);
    std::vector<uint16_t> cls_of(65536);
    int failures = 0, cases = 0;
    if (argc > 3 && !std::strcmp(argv[2], "--time")) {
        const size_t n = (size_t)std::atoll(argv[3]);
        for (uint32_t c = 0; c < 65536; c++) cls_of[c] = (c >= 'a' && c <= 'z') ? (uint16_t)(c - 'a' + 1) : 0;
        std::vector<uint16_t> chars;
        std::vector<KwRef> kws;
        for (size_t k = 0; k < n; k++) {
            int len = 3 + (int)(rng() % 10);
            kws.push_back(KwRef{(int64_t)chars.size(), len, (int64_t)k});
            for (int i = 0; i < len; i++) chars.push_back((uint16_t)('a' + rng() % 26));
        }
        TrieInsertParams P{chars.data(), cls_of.data(), true, false, false, 12};
        auto t0 = std::chrono::steady_clock::now();
        TrieArrays a = insert_serial(kws, P);
        auto t1 = std::chrono::steady_clock::now();
        TrieArrays b = insert_sharded(kws, P, 27, threads);
        auto t2 = std::chrono::steady_clock::now();
        std::printf("%zu keywords, %zu nodes: serial %.1f ms, sharded(%u threads) %.1f ms, equal=%d\n", n, a.info.size(),
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), threads,
                    std::chrono::duration<double, std::milli>(t2 - t1).count(), (int)same(a, b));
        return same(a, b) ? 0 : 1;
    }
    for (int round = 0; round < 60; round++) {
        const int n_classes = round % 6 == 5 ? 2000 : 2 + (int)(rng() % 40);  // class 0 = "other" never occurs in a keyword
        for (uint32_t c = 0; c < 65536; c++) cls_of[c] = (uint16_t)(c < (uint32_t)n_classes ? c : 0);
        const size_t n = round < 4 ? (size_t)round : 1 + (size_t)(rng() % 3000);
        const int max_len = 1 + (int)(rng() % 14);
        std::vector<uint16_t> chars;
        std::vector<KwRef> kws;
        int longest = 0;
        for (size_t k = 0; k < n; k++) {
            const int len = 1 + (int)(rng() % max_len);
            longest = std::max(longest, len);
            if (k > 0 && rng() % 5 == 0) {  // duplicate or prefix of an earlier keyword: value rules and shared paths
                const KwRef e = kws[rng() % kws.size()];
                const int l2 = 1 + (int)(rng() % e.len);
                kws.push_back(KwRef{(int64_t)chars.size(), l2, (int64_t)(k * 3)});
                for (int i = 0; i < l2; i++) chars.push_back(uint16_t(chars[(size_t)e.begin + i]));
                continue;
            }
            kws.push_back(KwRef{(int64_t)chars.size(), len, (int64_t)(k * 3)});
            for (int i = 0; i < len; i++) chars.push_back((uint16_t)(1 + rng() % (n_classes - 1)));
        }
        if (chars.empty()) chars.push_back(0);
        for (int variant = 0; variant < 4; variant++) {
            TrieInsertParams P{chars.data(), cls_of.data(), (variant & 1) != 0, true, (variant & 2) != 0, longest};
            TrieArrays a = insert_serial(kws, P);
            for (unsigned t : {1u, threads, 13u}) {
                TrieArrays b = insert_sharded(kws, P, n_classes, t);
                cases++;
                if (!same(a, b)) {
                    failures++;
                    std::fprintf(stderr, "FAIL round %d variant %d threads %u (n=%zu classes=%d nodes %zu vs %zu)\n", round, variant, t, n,
                                 n_classes, a.info.size(), b.info.size());
                }
            }
        }
    }
    std::printf("%d cases, %d failures\n", cases, failures);
    return failures ? 1 : 0;
}
