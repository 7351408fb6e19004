// bench.py's "e2e_with_replay" leg: one AhoCorasickSet.match(String, SetMatchListener) through the C++ mirror of the
// reference API (include/acgpu.hpp) with a COUNTING listener - host buffer in, H2D, kernels, D2H and the listener
// replay (one call per match, like the reference's loop AhoCorasickSet.java:193-252) all inside the timed region.
// Built by bench.py: g++ -std=c++17 -O2 -shared -fPIC tools/e2e_replay.cpp -Iinclude -Lahocorasick_b200 -lacgpu
#include <chrono>
#include <memory>
#include <vector>

#include "acgpu.hpp"

extern "C" {

// returns seconds per match() call (mean of `steps` after `warmup`), < 0 on error; *matches = listener calls of one call
double e2e_replay_count(const uint16_t *kw_chars, const int64_t *kw_offsets, int64_t n_keywords, const uint16_t *hay, int64_t n,
                        int device, int warmup, int steps, int64_t *matches) {
    try {
        std::vector<acgpu::String> kws;
        kws.reserve((size_t)n_keywords);
        for (int64_t i = 0; i < n_keywords; i++)
            kws.emplace_back(reinterpret_cast<const char16_t *>(kw_chars) + kw_offsets[i], (size_t)(kw_offsets[i + 1] - kw_offsets[i]));
        acgpu::AhoCorasickSet set(kws, true, device);
        const acgpu::String haystack(reinterpret_cast<const char16_t *>(hay), (size_t)n);
        int64_t count = 0;
        auto listener = [&count](const acgpu::String &, int, int) {
            ++count;
            return true;
        };
        for (int i = 0; i < warmup; i++) set.match(haystack, listener);
        count = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < steps; i++) set.match(haystack, listener);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *matches = count / (steps > 0 ? steps : 1);
        return dt / (steps > 0 ? steps : 1);
    } catch (const std::exception &) {
        return -1.0;
    }
}
}
