// libacgpu.so — C ABI (include/acgpu.h) over the sm_100a kernels.  No CPU matching path exists here:
// every match entry point launches CUDA kernels or fails.
#include <chrono>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <initializer_list>
#include <vector>
#include <unordered_map>
#include <stdexcept>

#include "../../include/acgpu.h"
#include "builder.hpp"
#include "kernels.cuh"
#include "kernel_tier.cuh"
#include "kernel_mask.cuh"
#include "kernel_emit.cuh"
#include "kernel_fuse.cuh"
#include "kernel_ww3.cuh"
#include "kernel_sel2.cuh"
#include "kernel_ww.cuh"
#include "kernel_wide.cuh"
#include "kernel_wwlit.cuh"
#include "tier_launch.hpp"

using namespace acgpu;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? ACGPU_ENODEVICE \
                                                                                       : ACGPU_ECUDA,  \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
        }                                                                                              \
    } while (0)

constexpr uint64_t kMagic = 0xAC69B200AC69B200ull;

// Streams, events and the pinned/device counters of one host-buffer call; parked in the matcher between calls
// (creating them costs far more than a small scan).
struct CallCtx {
    cudaStream_t s_up = nullptr, s_k = nullptr, s_dn = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr};
    unsigned long long *d_total = nullptr;  // [2] device
    unsigned long long *h_total = nullptr;  // [2] pinned
    unsigned long long *d_map = nullptr;    // [2][16] device: composed maps of two chain shards in flight (run_chain), made on first use
    unsigned long long *h_map = nullptr;    // [2][16] pinned
    void destroy() {
        for (int i = 0; i < 2; i++) {
            if (ev_up[i]) cudaEventDestroy(ev_up[i]);
            if (ev_k[i]) cudaEventDestroy(ev_k[i]);
            if (ev_dn[i]) cudaEventDestroy(ev_dn[i]);
        }
        if (d_total) cudaFree(d_total);
        if (h_total) cudaFreeHost(h_total);
        if (d_map) cudaFree(d_map);
        if (h_map) cudaFreeHost(h_map);
        if (s_up) cudaStreamDestroy(s_up);
        if (s_k) cudaStreamDestroy(s_k);
        if (s_dn) cudaStreamDestroy(s_dn);
    }
};

struct Matcher {
    uint64_t magic = kMagic;
    std::mutex ctx_mu;
    std::vector<CallCtx *> ctx_free;
    int device = 0;
    int sm_count = 148;
    HostAutomaton host;  // tables kept for introspection; device copies below
    DevAutomaton dev{};
    void *d_blob = nullptr;  // one allocation holding every table
    int64_t table_bytes = 0;
    bool sel_attr_set = false;
    // generation-2 (tiered) tables, AhoCorasick family on narrow alphabets
    bool use_tier = false;
    DevTier tier{};
    void *d_tier_blob = nullptr;
    L2Window l2win;         // child masks + deep table, kept L2-resident across the streaming traffic
    size_t mask_smem = 0;
    bool mask_pair = false;   // k_tier_pair (generation 4, pair rows) instead of k_tier_mask
    size_t fuse_smem = 0;     // k_tier_fused (generation 5, one launch): shared memory it needs; 0 = not available
    // AhoCorasick family outside the tier envelope (kernel_wide.cuh)
    bool use_wide = false;
    DevWide wide{};
    void *d_wide_blob = nullptr;
    void *d_wide_vals = nullptr;   // Map values by keyword hash (HostAutomaton::wide_vals)
    size_t wide_smem = 0;
    bool wide_tile = false;   // k_wide_tile (pair table + path-compressed edges) instead of k_wide_mask
    // WholeWordLongest with phrase keywords: walk starts compacted by k_wwl_starts (kernel_ww.cuh)
    bool use_wwl2 = false;
    uint16_t *d_wwl_wcls = nullptr;
    // WholeWord hash tables (kernel_ww.cuh)
    bool use_ww = false;
    bool use_ww3 = false;     // generation 3 (kernel_ww3.cuh: hits, row scan, emit) instead of k_ww_scan
    DevWw ww{};
    void *d_ww_blob = nullptr;
    // >= 0: the reference's loop is followed literally, one thread per synchronisation point (kernel_wwlit.cuh::k_segments):
    // 0 = WholeWord with quirk Q7, 1 / 2 = Longest / Shortest with keywords the selection kernels cannot hold (> 2 047 chars),
    // 4 = WholeWordLongest with keywords > 254 chars or quirk Q7
    int literal_family = -1;
};

Matcher *as_matcher(uint64_t h) {
    Matcher *m = reinterpret_cast<Matcher *>(static_cast<uintptr_t>(h));
    if (!m || m->magic != kMagic) return nullptr;
    return m;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int upload(Matcher *m) {
    const HostAutomaton &a = m->host;
    size_t off = 0;
    auto reserve = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    size_t o_cls = reserve(65536 * sizeof(uint16_t));
    size_t o_word = reserve(a.wordbits.size() * sizeof(uint32_t));
    size_t o_wfold = reserve(a.wordbits_fold.size() * sizeof(uint32_t));
    size_t o_root = reserve(a.root.size() * sizeof(RootEdge));
    size_t o_edges = reserve(a.edges.size() * sizeof(Edge));
    size_t o_val = reserve(a.node_value.size() * sizeof(uint32_t));
    CU_TRY(cudaMalloc(&m->d_blob, off));
    m->table_bytes = static_cast<int64_t>(off);
    char *b = static_cast<char *>(m->d_blob);
    CU_TRY(cudaMemcpy(b + o_cls, a.cls.data(), 65536 * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (!a.wordbits.empty())
        CU_TRY(cudaMemcpy(b + o_word, a.wordbits.data(), a.wordbits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!a.wordbits_fold.empty())
        CU_TRY(cudaMemcpy(b + o_wfold, a.wordbits_fold.data(), a.wordbits_fold.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_root, a.root.data(), a.root.size() * sizeof(RootEdge), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_edges, a.edges.data(), a.edges.size() * sizeof(Edge), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_val, a.node_value.data(), a.node_value.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    DevAutomaton &d = m->dev;
    d.cls = reinterpret_cast<const uint16_t *>(b + o_cls);
    d.wordbits = a.wordbits.empty() ? nullptr : reinterpret_cast<const uint32_t *>(b + o_word);
    d.wordbits_fold = a.wordbits_fold.empty() ? nullptr : reinterpret_cast<const uint32_t *>(b + o_wfold);
    d.root = reinterpret_cast<const uint2 *>(b + o_root);
    d.edges = reinterpret_cast<const uint4 *>(b + o_edges);
    d.node_value = reinterpret_cast<const uint32_t *>(b + o_val);
    d.edge_mask = a.edge_mask;
    d.max_len = a.max_len;
    d.n_classes = a.n_classes;
    d.has_other = a.has_other ? 1 : 0;
    d.family = a.family;
    d.is_map = a.is_map ? 1 : 0;
    return ACGPU_OK;
}

int upload_tier(Matcher *m) {
    const TierTables &t = m->host.tier;
    m->use_tier = false;
    // AhoCorasick: end masks over the reversed-keyword trie; Longest / Shortest: start masks over the forward trie
    if (!t.ok || m->host.family == ACGPU_WHOLEWORD || m->host.family == ACGPU_WHOLEWORDLONGEST) return ACGPU_OK;
    const char *force = getenv("ACGPU_FORCE_GEN1");
    if (force && force[0] == '1') return ACGPU_OK;
    // generation 4 (k_tier_pair, pair rows) or generation 3 (k_tier_mask).  The pair kernel's gate bit skips the gather of
    // pairs whose level-K row has no continuation at all: on configs[1] (100 k keywords) it wins 2 %; on a saturated
    // dictionary (configs[4]: most level-K contexts continue) next to nothing is skipped and k_tier_mask is 1.4 % faster
    // (profiles/r02_summary.md).  ACGPU_MASK_GEN=3 / 4 forces one (A/B runs).
    const char *gen = getenv("ACGPU_MASK_GEN");
    bool saturated = false;
    if (!t.kidmask.empty()) {
        size_t live = 0;
        for (size_t i = 0; i < t.kidmask.size(); i += 2) live += (t.kidmask[i] | t.kidmask[i + 1]) != 0u;
        saturated = live * 4 > (t.kidmask.size() / 2) * 3;  // more than three quarters of the contexts continue
    }
    m->mask_pair = !t.prow_words.empty() && (gen ? gen[0] != '3' : !saturated);
    m->mask_smem = mask_smem_bytes(m->mask_pair ? t.prow_words.size() : t.row_words.size());
    {
        // generation 5 (k_tier_fused: masks and records in one persistent launch) - opt-in (ACGPU_FUSE=1): parity-green, but
        // measured 7.1 ms against 3.8 ms per 10^9 chars on configs[4] (profiles/r02_fused_summary.md)
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
        const size_t need = fuse_smem_bytes(t.row_words.size(), m->dev.is_map != 0);
        const char *fz = getenv("ACGPU_FUSE");
        m->fuse_smem = (need <= static_cast<size_t>(max_smem) && fz && fz[0] == '1') ? need : 0;
    }
    if (m->mask_smem > 227 * 1024) return ACGPU_OK;
    size_t off = 0;
    auto reserve = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + std::max<size_t>(bytes, 16), 512);  // linear textures want 512-byte aligned bases
        return o;
    };
    size_t o_rows = reserve(t.row_words.size() * 4);
    size_t o_prows = reserve(t.prow_words.size() * 4);
    size_t o_cls8 = reserve(256);
    size_t o_kid = reserve(t.kidmask.size() * 4);
    size_t o_deep = reserve(t.buckets.size() * 4);
    size_t o_vb = reserve(t.vbuckets.size() * 4);
    CU_TRY(cudaMalloc(&m->d_tier_blob, off));
    m->table_bytes += static_cast<int64_t>(off);
    char *b = static_cast<char *>(m->d_tier_blob);
    uint8_t cls8[256];
    for (int c = 0; c < 256; c++) cls8[c] = static_cast<uint8_t>(m->host.cls[c]);
    if (!t.row_words.empty()) CU_TRY(cudaMemcpy(b + o_rows, t.row_words.data(), t.row_words.size() * 4, cudaMemcpyHostToDevice));
    if (!t.prow_words.empty()) CU_TRY(cudaMemcpy(b + o_prows, t.prow_words.data(), t.prow_words.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_cls8, cls8, 256, cudaMemcpyHostToDevice));
    if (!t.kidmask.empty()) CU_TRY(cudaMemcpy(b + o_kid, t.kidmask.data(), t.kidmask.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_deep, t.buckets.data(), t.buckets.size() * 4, cudaMemcpyHostToDevice));
    if (!t.vbuckets.empty()) CU_TRY(cudaMemcpy(b + o_vb, t.vbuckets.data(), t.vbuckets.size() * 4, cudaMemcpyHostToDevice));
    DevTier &d = m->tier;
    d.cls8 = reinterpret_cast<const uint32_t *>(b + o_cls8);
    d.kidmask = t.kidmask.empty() ? nullptr : reinterpret_cast<const uint32_t *>(b + o_kid);
    d.kid_tex = 0;
    if (!t.kidmask.empty()) {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = b + o_kid;
        rd.res.linear.desc = cudaCreateChannelDesc<uint2>();  // {backward, forward} continuation masks of one context
        rd.res.linear.sizeInBytes = t.kidmask.size() * 4;
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        td.filterMode = cudaFilterModePoint;
        td.addressMode[0] = cudaAddressModeClamp;
        td.normalizedCoords = 0;
        CU_TRY(cudaCreateTextureObject(&d.kid_tex, &rd, &td, nullptr));
    }
    d.buckets = reinterpret_cast<const uint4 *>(b + o_deep);
    d.hash_seed = t.hash_seed;
    d.vbuckets = reinterpret_cast<const uint4 *>(b + o_vb);
    d.vseed = t.vseed;
    d.n_vbuckets = t.n_vbuckets;
    d.row_words = reinterpret_cast<const uint32_t *>(b + o_rows);
    d.n_row_words = static_cast<uint32_t>(t.row_words.size());
    for (int j = 0; j < 10; j++) d.row_off[j] = t.row_off[j];
    d.prow_words = t.prow_words.empty() ? nullptr : reinterpret_cast<const uint32_t *>(b + o_prows);
    d.n_prow_words = static_cast<uint32_t>(t.prow_words.size());
    for (int j = 0; j < 10; j++) d.prow_off[j] = t.prow_off[j];
    d.pair_gate_bit = t.pair_gate_bit;
    d.pair_low_bit = t.pair_low_bit;
    {
        // L2 residency window over [child masks | deep table] (adjacent in the blob)
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, m->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, m->device);
        const size_t want = (o_deep + t.buckets.size() * 4) - o_kid;
        const char *win = getenv("ACGPU_L2_WINDOW");  // measured slower than plain L2 on B200 (profiles/): opt-in only
        if (max_win > 0 && max_persist > 0 && !t.kidmask.empty() && (win && win[0] == '1')) {
            const size_t persist = std::min<size_t>(static_cast<size_t>(max_persist), want);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
                m->l2win.base = b + o_kid;
                m->l2win.bytes = std::min<size_t>(want, static_cast<size_t>(max_win));
                m->l2win.hit_ratio = want <= persist ? 1.0f : static_cast<float>(persist) / static_cast<float>(want);
            } else {
                cudaGetLastError();
            }
        }
    }
    d.n_buckets = t.n_buckets;
    d.inv_b = (65536u + static_cast<uint32_t>(t.b) - 1u) / static_cast<uint32_t>(t.b);
    d.term_levels = t.term_levels;
    d.b = t.b;
    d.C = t.C;
    d.K = t.K;
    for (int j = 0; j < 10; j++) {
        d.pow_c[j] = t.pow_c[j];
    }
    m->use_tier = true;
    return ACGPU_OK;
}

int upload_wide(Matcher *m) {
    m->use_wide = false;
    if (!m->host.wide_ok || m->use_tier) return ACGPU_OK;
    const char *force = getenv("ACGPU_FORCE_GEN1");
    if (force && force[0] == '1') return ACGPU_OK;
    const bool pair = !m->host.wide_pair.empty();
    m->wide.C = m->host.n_classes;
    m->wide.pair = nullptr;
    m->wide.chain = nullptr;
    m->wide.pair16 = nullptr;
    m->wide.vals = nullptr;
    m->wide.n_vbuckets = 0;
    m->wide_tile = false;
    {
        const char *vw = getenv("ACGPU_WIDE_VALUES");  // ACGPU_WIDE_VALUES=walk: the second trie walk per record (A/B runs)
        if (!m->host.wide_vals.empty() && !(vw && vw[0] == 'w')) {
            const size_t bytes = m->host.wide_vals.size() * 4;
            CU_TRY(cudaMalloc(&m->d_wide_vals, bytes));
            m->table_bytes += static_cast<int64_t>(bytes);
            CU_TRY(cudaMemcpy(m->d_wide_vals, m->host.wide_vals.data(), bytes, cudaMemcpyHostToDevice));
            m->wide.vals = static_cast<const uint4 *>(m->d_wide_vals);
            m->wide.n_vbuckets = m->host.wide_n_vbuckets;
        }
    }
    if (pair) {
        const size_t pair_bytes = align_up(m->host.wide_pair.size() * 4, 256), pair16_bytes = align_up(m->host.wide_pair16.size() * 4, 256);
        const size_t chain_bytes = m->host.wide_chain.size() * 4;
        CU_TRY(cudaMalloc(&m->d_wide_blob, pair_bytes + pair16_bytes + chain_bytes + 256));
        m->table_bytes += static_cast<int64_t>(pair_bytes + pair16_bytes + chain_bytes);
        char *b = static_cast<char *>(m->d_wide_blob);
        CU_TRY(cudaMemcpy(b, m->host.wide_pair.data(), m->host.wide_pair.size() * 4, cudaMemcpyHostToDevice));
        m->wide.pair = reinterpret_cast<const uint2 *>(b);
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
        if (!m->host.wide_pair16.empty() && wide_tile_smem_bytes(m->host.n_classes) <= static_cast<size_t>(max_smem)) {
            CU_TRY(cudaMemcpy(b + pair_bytes, m->host.wide_pair16.data(), m->host.wide_pair16.size() * 4, cudaMemcpyHostToDevice));
            if (chain_bytes) CU_TRY(cudaMemcpy(b + pair_bytes + pair16_bytes, m->host.wide_chain.data(), chain_bytes, cudaMemcpyHostToDevice));
            m->wide.pair16 = reinterpret_cast<const uint4 *>(b + pair_bytes);
            m->wide.chain = reinterpret_cast<const uint4 *>(b + pair_bytes + pair16_bytes);
            // generation 2 of the wide path (k_wide_tile: CTA-wide walk queue over path-compressed edges); ACGPU_WIDE_GEN=1 keeps k_wide_mask (A/B runs)
            const char *gen = getenv("ACGPU_WIDE_GEN");
            m->wide_tile = !(gen && gen[0] == '1');
            CU_TRY(cudaFuncSetAttribute(k_wide_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_tile_smem_bytes(kWidePairMax)));
        }
    }
    // at most 512 + 64 * 64 * 8 + 8 * 2 880 = 56 320 bytes: one attribute for every matcher (a per-matcher value would
    // shrink the limit under a live matcher with a larger table)
    m->wide_smem = wide_smem_bytes(m->host.n_classes, pair);
    static_assert(wide_smem_bytes(kWidePairMax, true) <= 64 * 1024, "k_wide_mask shared memory");
    CU_TRY(cudaFuncSetAttribute(k_wide_mask<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(kWidePairMax, true)));
    m->use_wide = true;
    return ACGPU_OK;
}

int upload_ww(Matcher *m) {
    const WwTables &t = m->host.ww;
    m->use_ww = false;
    m->use_wwl2 = false;
    {
        const char *force = getenv("ACGPU_FORCE_GEN1"), *gen = getenv("ACGPU_WWL_GEN");  // ACGPU_WWL_GEN=1: the generation-1 chain over every position (A/B runs)
        if (m->host.family == ACGPU_WHOLEWORDLONGEST && !m->host.wwl_wcls.empty() && !(force && force[0] == '1') && !(gen && gen[0] == '1')) {
            CU_TRY(cudaMalloc(reinterpret_cast<void **>(&m->d_wwl_wcls), 65536 * 2));
            m->table_bytes += 65536 * 2;
            CU_TRY(cudaMemcpy(m->d_wwl_wcls, m->host.wwl_wcls.data(), 65536 * 2, cudaMemcpyHostToDevice));
            CU_TRY(cudaFuncSetAttribute(k_wwl_starts, cudaFuncAttributeMaxDynamicSharedMemorySize, (kWwTile + 16 * 16 + 2) * 2));
            m->use_wwl2 = true;
        }
    }
    if (!t.ok || (m->host.family != ACGPU_WHOLEWORD && m->host.family != ACGPU_WHOLEWORDLONGEST)) return ACGPU_OK;
    const char *force = getenv("ACGPU_FORCE_GEN1");
    if (force && force[0] == '1') return ACGPU_OK;
    size_t off = 0;
    auto reserve = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + std::max<size_t>(bytes, 16), 256);
        return o;
    };
    const size_t o_wcls = reserve(65536 * 2);
    const size_t o_bk = reserve(t.buckets.size() * 4);
    const size_t o_pool = reserve(t.pool.size() * 2);
    const size_t o_bloom = reserve(t.bloom.size() * 4);
    CU_TRY(cudaMalloc(&m->d_ww_blob, off));
    m->table_bytes += static_cast<int64_t>(off);
    char *b = static_cast<char *>(m->d_ww_blob);
    CU_TRY(cudaMemcpy(b + o_wcls, t.wcls.data(), 65536 * 2, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_bk, t.buckets.data(), t.buckets.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(b + o_pool, t.pool.data(), t.pool.size() * 2, cudaMemcpyHostToDevice));
    m->ww.wcls = reinterpret_cast<const uint16_t *>(b + o_wcls);
    m->ww.buckets = reinterpret_cast<const uint4 *>(b + o_bk);
    m->ww.pool = reinterpret_cast<const uint16_t *>(b + o_pool);
    m->ww.n_buckets = t.n_buckets;
    m->ww.max_len = m->host.max_len;
    m->ww.bloom = nullptr;
    m->ww.bloom_bits = 0;
    m->use_ww3 = t.poly;
    if (t.poly) {
        if (!t.bloom.empty()) {
            CU_TRY(cudaMemcpy(b + o_bloom, t.bloom.data(), t.bloom.size() * 4, cudaMemcpyHostToDevice));
            m->ww.bloom = reinterpret_cast<const uint32_t *>(b + o_bloom);
            m->ww.bloom_bits = t.bloom_bits;
        }
        CU_TRY(cudaFuncSetAttribute(k_ww3_hits<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ww3_smem_bytes(512 * 1024, true)));
        CU_TRY(cudaFuncSetAttribute(k_ww3_hits<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ww3_smem_bytes(512 * 1024, false)));
        CU_TRY(cudaFuncSetAttribute(k_ww3_hits<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ww3_smem_bytes(0, true)));
        CU_TRY(cudaFuncSetAttribute(k_ww3_hits<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ww3_smem_bytes(0, false)));
    }
    m->use_ww = true;
    return ACGPU_OK;
}

// k_tier_mask: variant 1 reads level K-1 from the spare bit of the level-K rows, which exists for at most 31 classes
// (k_tier_pair: its LOW bit, which exists for at most 30 classes)
int mask_low_variant(const DevTier &t, bool pair) {
    const uint32_t below = t.term_levels & ((1u << t.K) - 1u);  // bits 1..K-1 (bit 0 is never set)
    if (below == 0) return 2;
    if (below == (1u << (t.K - 1)) && (pair ? t.pair_low_bit != 0u : t.C <= 31)) return 1;
    return 0;
}

// Scratch for one match call, carved from one stream-ordered allocation.
struct Scratch {
    void *base = nullptr;
    cudaStream_t st = nullptr;
    size_t off = 0;
    size_t reserve(size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

struct RunOpts {
    int32_t pos_base = 0;  // stream offset added to positions
    int64_t ctx = 0;       // leading window chars that are left context only (streaming)
    int64_t entry0 = 0;    // chain position on entry (selection families)
    int64_t chain_n = -1;  // chain domain [0, chain_n); -1 => n
    int64_t *d_carry = nullptr;  // [2] int64 (selection families)
    int64_t abs0 = 0;      // window position of the first char of the input, -1 = before this window (WholeWordLongest)
    bool folded_scroll = false;  // literal WholeWord matchers (quirk Q7): the Readable overloads scroll on the lower-cased char
};

// rows per ticket of the mask kernels: 32 for long haystacks; short ones take smaller tickets so that every warp of the grid
// (sm_count x 32) gets work - 8 M chars are 977 tickets of 32 rows: 31 of 148 SMs busy.  ACGPU_CHUNK_ROWS forces a value.
int mask_chunk_rows(const Matcher *m, int64_t n_rows) {
    static const char *env = getenv("ACGPU_CHUNK_ROWS");
    if (env && atoi(env) > 0) return std::min(kMaskChunkRows, atoi(env));
    const int64_t warps = static_cast<int64_t>(m->sm_count) * kMaskWarps;
    int rows = kMaskChunkRows;
    while (rows > 2 && (n_rows + rows - 1) / rows < 2 * warps) rows >>= 1;
    return rows;
}

int launch_mask(Matcher *m, const MaskArgs &P, int grid, cudaStream_t st, bool mir = false) {
    const int low = mask_low_variant(m->tier, m->mask_pair);
    cudaError_t e;
    switch (m->tier.K) {
    case 1: e = mask_launch_1(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 2: e = mask_launch_2(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 3: e = mask_launch_3(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 4: e = mask_launch_4(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 5: e = mask_launch_5(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 6: e = mask_launch_6(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    case 7: e = mask_launch_7(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    default: e = mask_launch_8(low, mir, m->mask_pair, m->dev, m->tier, P, grid, m->mask_smem, m->l2win, st); break;
    }
    CU_TRY(e);
    return ACGPU_OK;
}

// Longest / Shortest, narrow alphabets: start masks (mirrored k_tier_mask) -> exit maps -> scan -> records
// (kernel_sel2.cuh).  Two phases so that range shards of one haystack can exchange their maps in between:
//   A  (entry-independent) masks of the window, the exit map of every domain tile, the group maps, and - on request -
//      the shard's composed map for all 16 entry offsets
//   B  the true chain from a given entry offset: group / tile entries, records, values
struct Sel2Run {
    Matcher *m = nullptr;
    void *ws = nullptr;
    cudaStream_t st = nullptr;
    MaskArgs P{};
    Sel2Args Q{};
    size_t o_status = 0, o_ctr = 0;
    int64_t n = 0, moff = 0, n_rows = 0;
    bool longest = true;
};

// n_dom_tiles < 0: the whole window is the chain domain
int sel2_setup(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t n_dom_tiles, cudaStream_t st, const RunOpts &opt, Sel2Run &R) {
    // the mirrored kernel loads hay[n - 8 - p0, n - p0): rows start at origin <= 0 with origin = mis + n (mod 8)
    const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(d_hay) >> 1) & 7);
    const int64_t r = (mis + n) & 7;
    const int64_t origin = r ? r - 8 : 0;
    const int64_t n_rows = (n - origin + kMaskRow - 1) / kMaskRow;
    const int64_t n_idx = n_rows * kMaskRow;
    const int64_t moff = n_idx - n + origin;
    const int64_t all_tiles = (n_idx + kS2Tile - 1) / kS2Tile;
    const int64_t n_tiles = n_dom_tiles < 0 ? all_tiles : std::min(n_dom_tiles, all_tiles);
    Scratch S;
    const size_t o_ctr = S.reserve(256);
    const size_t o_cnt = S.reserve(static_cast<size_t>(n_rows) * 4);
    const size_t o_mask = S.reserve(static_cast<size_t>(n_idx) * 2);
    const size_t o_map = S.reserve(static_cast<size_t>(n_tiles) * kS2Ent * 4);
    const size_t o_ent = S.reserve(static_cast<size_t>(n_tiles));
    const size_t o_base = S.reserve(static_cast<size_t>(n_tiles) * 8);
    const int64_t n_groups = (n_tiles + kS2Group - 1) / kS2Group;
    const size_t o_gmap = S.reserve(static_cast<size_t>(n_groups) * kS2Ent * 4);
    const size_t o_gent = S.reserve(static_cast<size_t>(n_groups));
    const size_t o_gbase = S.reserve(static_cast<size_t>(n_groups) * 8);
    const size_t o_status = S.reserve(static_cast<size_t>(n_tiles) * sizeof(S2Status));
    CU_TRY(cudaMallocAsync(&R.ws, S.off, st));
    char *w = static_cast<char *>(R.ws);
    R.m = m;
    R.st = st;
    R.n = n;
    R.moff = moff;
    R.n_rows = n_rows;
    R.o_status = o_status;
    R.o_ctr = o_ctr;
    R.longest = m->dev.family == ACGPU_LONGEST;
    MaskArgs &P = R.P;
    P.hay = d_hay;
    P.n = n;
    P.emit_from = 0;
    P.emit_to = n;
    P.origin = origin;
    P.masks = reinterpret_cast<uint32_t *>(w + o_mask);
    P.row_count = reinterpret_cast<uint32_t *>(w + o_cnt);
    P.ticket = reinterpret_cast<unsigned int *>(w + o_ctr);
    P.n_rows = n_rows;
    Sel2Args &Q = R.Q;
    Q.masks = P.masks;
    Q.n_idx = n_idx;
    Q.moff = moff;
    Q.n_tiles = n_tiles;
    Q.tile_map = reinterpret_cast<uint32_t *>(w + o_map);
    Q.tile_entry = reinterpret_cast<uint8_t *>(w + o_ent);
    Q.tile_base = reinterpret_cast<unsigned long long *>(w + o_base);
    Q.n_groups = n_groups;
    Q.group_map = reinterpret_cast<uint32_t *>(w + o_gmap);
    Q.group_entry = reinterpret_cast<uint8_t *>(w + o_gent);
    Q.group_base = reinterpret_cast<unsigned long long *>(w + o_gbase);
    Q.hay = d_hay;
    Q.n = n;
    Q.pos_base = opt.pos_base;
    return ACGPU_OK;
}

void sel2_free(Sel2Run &R) {
    if (R.ws) cudaFreeAsync(R.ws, R.st);
    R.ws = nullptr;
}

int sel2_masks(Sel2Run &R) {
    CU_TRY(cudaMemsetAsync(static_cast<char *>(R.ws) + R.o_ctr, 0, 256, R.st));
    R.P.chunk_rows = mask_chunk_rows(R.m, R.n_rows);
    const int64_t n_chunks = (R.n_rows + R.P.chunk_rows - 1) / R.P.chunk_rows;
    const int grid = static_cast<int>(std::min<int64_t>((n_chunks + kMaskWarps - 1) / kMaskWarps, R.m->sm_count));
    return launch_mask(R.m, R.P, grid, R.st, true);
}

int sel2_maps(Sel2Run &R, int64_t only_first_tiles = -1) {
    Sel2Args Q = R.Q;
    if (only_first_tiles >= 0) Q.n_tiles = std::min(Q.n_tiles, only_first_tiles);
    if (Q.n_tiles <= 0) return ACGPU_OK;
    const size_t smem = static_cast<size_t>(kS2SmemWords) * 4;
    const int sgrid = static_cast<int>(std::min<int64_t>(Q.n_tiles, static_cast<int64_t>(R.m->sm_count) * 16));
    if (R.longest)
        k_sel2_map<kModeLongest><<<sgrid, kS2Threads, smem, R.st>>>(Q);
    else
        k_sel2_map<kModeShortest><<<sgrid, kS2Threads, smem, R.st>>>(Q);
    CU_TRY(cudaGetLastError());
    // the group maps read every tile map of their group: all groups (one-shot) or group 0 only (a re-run of tile 0)
    k_sel2_group<<<static_cast<unsigned>(only_first_tiles >= 0 ? 1 : R.Q.n_groups), 32, 0, R.st>>>(R.Q);
    CU_TRY(cudaGetLastError());
    return ACGPU_OK;
}

// the composed map of the shard for all 16 entry offsets -> d_map[16] (exit offset | matches << 8)
int sel2_shard_map(Sel2Run &R, unsigned long long *d_map, unsigned long long *d_scratch_total) {
    Sel2Args Q = R.Q;
    Q.entry0 = 0;
    Q.shard_map = d_map;
    Q.total_out = d_scratch_total;
    k_sel2_top<<<1, 1024, static_cast<size_t>(Q.n_groups) * kS2Ent * 4, R.st>>>(Q);
    CU_TRY(cudaGetLastError());
    return ACGPU_OK;
}

int sel2_records(Sel2Run &R, uint32_t entry0, int2 *d_pos, uint32_t *d_val, int64_t cap, unsigned long long *d_total) {
    Sel2Args Q = R.Q;
    Q.entry0 = entry0;
    Q.shard_map = nullptr;
    Q.total_out = d_total;
    Q.pos_out = d_pos;
    Q.val_out = d_val;
    Q.cap = cap;
    if (Q.n_tiles <= 0) {
        CU_TRY(cudaMemsetAsync(d_total, 0, 8, R.st));
        return ACGPU_OK;
    }
    const size_t smem = static_cast<size_t>(kS2SmemWords) * 4;
    const int sgrid = static_cast<int>(std::min<int64_t>(Q.n_tiles, static_cast<int64_t>(R.m->sm_count) * 16));
    k_sel2_top<<<1, 1024, static_cast<size_t>(Q.n_groups) * kS2Ent * 4, R.st>>>(Q);
    CU_TRY(cudaGetLastError());
    k_sel2_tiles<<<static_cast<unsigned>(Q.n_groups), 32, 0, R.st>>>(Q);
    CU_TRY(cudaGetLastError());
    if (cap > 0) {
        if (R.longest)
            k_sel2_emit<kModeLongest><<<sgrid, kS2Threads, smem, R.st>>>(Q);
        else
            k_sel2_emit<kModeShortest><<<sgrid, kS2Threads, smem, R.st>>>(Q);
        CU_TRY(cudaGetLastError());
    }
    if (cap > 0 && R.m->dev.is_map) {
        const int vgrid = static_cast<int>(std::min<int64_t>(Q.n_tiles, static_cast<int64_t>(R.m->sm_count) * 16));
        switch (R.m->tier.b) {
        case 1: k_sel2_values<1><<<vgrid, kS2Threads, 0, R.st>>>(R.m->dev, R.m->tier, Q); break;
        case 2: k_sel2_values<2><<<vgrid, kS2Threads, 0, R.st>>>(R.m->dev, R.m->tier, Q); break;
        case 3: k_sel2_values<3><<<vgrid, kS2Threads, 0, R.st>>>(R.m->dev, R.m->tier, Q); break;
        case 4: k_sel2_values<4><<<vgrid, kS2Threads, 0, R.st>>>(R.m->dev, R.m->tier, Q); break;
        default: k_sel2_values<5><<<vgrid, kS2Threads, 0, R.st>>>(R.m->dev, R.m->tier, Q); break;
        }
        CU_TRY(cudaGetLastError());
    }
    return ACGPU_OK;
}

int enqueue_sel2(Matcher *m, const uint16_t *d_hay, int64_t n, int2 *d_pos, uint32_t *d_val, int64_t cap,
                 unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    Sel2Run R;
    int rc = sel2_setup(m, d_hay, n, -1, st, opt, R);
    if (rc != ACGPU_OK) return rc;
    rc = sel2_masks(R);
    // Measured (profiles/r01_s5_summary.md): the look-back over MAPS costs ~12 us per tile (following the chain through
    // the published maps is latency-bound), more than the second resolution pass it saves - opt-in only.
    const char *fused = getenv("ACGPU_SEL2_FUSED");
    if (rc == ACGPU_OK && fused && fused[0] == '1') {
        // single pass: maps, look-back over the maps, emission (k_sel2_fused)
        char *w = static_cast<char *>(R.ws);
        Sel2Args Q = R.Q;
        Q.total_out = d_total;
        Q.pos_out = d_pos;
        Q.val_out = d_val;
        Q.cap = cap;
        const size_t smem = static_cast<size_t>(kS2SmemWords) * 4;
        cudaError_t e = cudaMemsetAsync(w + R.o_status, 0, static_cast<size_t>(Q.n_tiles) * sizeof(S2Status), st);
        Q.status = reinterpret_cast<S2Status *>(w + R.o_status);
        Q.tile_counter = reinterpret_cast<unsigned int *>(w + R.o_ctr + 64);
        Q.err = reinterpret_cast<unsigned int *>(w + R.o_ctr + 128);
        const int fgrid = static_cast<int>(std::min<int64_t>(Q.n_tiles, static_cast<int64_t>(m->sm_count) * 4));
        if (e == cudaSuccess) {
            if (R.longest)
                k_sel2_fused<kModeLongest><<<fgrid, kS2Threads, smem, st>>>(Q);
            else
                k_sel2_fused<kModeShortest><<<fgrid, kS2Threads, smem, st>>>(Q);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && cap > 0 && m->dev.is_map) {
            const int vgrid = static_cast<int>(std::min<int64_t>(Q.n_tiles, static_cast<int64_t>(m->sm_count) * 16));
            switch (m->tier.b) {
            case 1: k_sel2_values<1><<<vgrid, kS2Threads, 0, st>>>(m->dev, m->tier, Q); break;
            case 2: k_sel2_values<2><<<vgrid, kS2Threads, 0, st>>>(m->dev, m->tier, Q); break;
            case 3: k_sel2_values<3><<<vgrid, kS2Threads, 0, st>>>(m->dev, m->tier, Q); break;
            case 4: k_sel2_values<4><<<vgrid, kS2Threads, 0, st>>>(m->dev, m->tier, Q); break;
            default: k_sel2_values<5><<<vgrid, kS2Threads, 0, st>>>(m->dev, m->tier, Q); break;
            }
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("k_sel2_fused: ") + cudaGetErrorString(e));
    } else {
        if (rc == ACGPU_OK) rc = sel2_maps(R);
        if (rc == ACGPU_OK) rc = sel2_records(R, 0, d_pos, d_val, cap, d_total);
    }
    sel2_free(R);
    return rc;
}

// AhoCorasick family, narrow alphabets: hit masks -> row-count scan -> records (kernel_mask.cuh)
struct MaskWs {  // one scratch block of a mask / scan (/ emit) run over n_rows rows
    int64_t n_rows = 0, n_blocks = 0;
    size_t o_ctr = 0, o_cnt = 0, o_blk = 0, o_mask = 0, bytes = 0;
};

MaskWs mask_ws_layout(int64_t emit_to, int64_t origin) {
    MaskWs L;
    L.n_rows = (emit_to - origin + kMaskRow - 1) / kMaskRow;
    L.n_blocks = (L.n_rows + kScanRows - 1) / kScanRows;
    Scratch S;
    L.o_ctr = S.reserve(256);
    L.o_cnt = S.reserve(static_cast<size_t>(L.n_rows) * 4);
    L.o_blk = S.reserve(static_cast<size_t>(L.n_blocks) * 8);
    L.o_mask = S.reserve(static_cast<size_t>(L.n_rows) * kMaskRow * 2);
    L.bytes = S.off;
    return L;
}

// k_tier_mask + k_row_scan into the scratch block w: hit masks of positions [origin, origin + n_rows * 256) at
// w + o_mask (16 bits each, zero outside [emit_from, emit_to)), the match count in *d_total
int enqueue_mask_scan(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int64_t origin, char *w,
                      const MaskWs &L, unsigned long long *d_total, cudaStream_t st) {
    CU_TRY(cudaMemsetAsync(w + L.o_ctr, 0, 256, st));
    MaskArgs P{};
    P.hay = d_hay;
    P.n = n;
    P.emit_from = emit_from;
    P.emit_to = emit_to;
    P.origin = origin;
    P.masks = reinterpret_cast<uint32_t *>(w + L.o_mask);
    P.row_count = reinterpret_cast<uint32_t *>(w + L.o_cnt);
    P.ticket = reinterpret_cast<unsigned int *>(w + L.o_ctr);
    P.n_rows = L.n_rows;
    P.chunk_rows = mask_chunk_rows(m, L.n_rows);
    const int64_t n_chunks = (L.n_rows + P.chunk_rows - 1) / P.chunk_rows;
    const int grid = static_cast<int>(std::min<int64_t>((n_chunks + kMaskWarps - 1) / kMaskWarps, m->sm_count));
    int rc = launch_mask(m, P, grid, st);
    if (rc != ACGPU_OK) return rc;
    ScanArgs SA{};
    SA.row_count = P.row_count;
    SA.block_excl = reinterpret_cast<unsigned long long *>(w + L.o_blk);
    SA.done = reinterpret_cast<unsigned int *>(w + L.o_ctr + 64);
    SA.total_out = d_total;
    SA.n_rows = L.n_rows;
    k_row_scan<<<static_cast<unsigned>(L.n_blocks), 1024, 0, st>>>(SA);
    CU_TRY(cudaGetLastError());
    return ACGPU_OK;
}

void fill_emit_args(EmitArgs &E, const uint16_t *d_hay, int64_t n, char *w, const MaskWs &L, int64_t origin, int2 *d_pos, uint32_t *d_val,
                    int64_t cap, const RunOpts &opt) {
    E.hay = d_hay;
    E.n = n;
    E.masks = reinterpret_cast<uint32_t *>(w + L.o_mask);
    E.row_excl = reinterpret_cast<uint32_t *>(w + L.o_cnt);
    E.block_excl = reinterpret_cast<unsigned long long *>(w + L.o_blk);
    E.n_rows = L.n_rows;
    E.origin = origin;
    E.pos_base = opt.pos_base;
    E.pos_out = d_pos;
    E.val_out = d_val;
    E.cap = cap;
}

int launch_emit(Matcher *m, const EmitArgs &E, cudaStream_t st) {
    // many small CTAs: the hardware scheduler evens out SMs that run at different speeds (measured: 128 per SM beats 8 by 12%)
    static const char *gm = getenv("ACGPU_EMIT_GRID");
    const int per_sm = gm ? std::max(1, atoi(gm)) : 128;
    const int egrid = static_cast<int>(std::min<int64_t>((E.n_rows + kEmitWarps - 1) / kEmitWarps, static_cast<int64_t>(m->sm_count) * per_sm));
    if (m->dev.is_map)
        k_tier_emit<true><<<egrid, kEmitWarps * 32, 0, st>>>(m->dev, m->tier, E);
    else
        k_tier_emit<false><<<egrid, kEmitWarps * 32, 0, st>>>(m->dev, m->tier, E);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ACGPU_ECUDA, std::string("k_tier_emit: ") + cudaGetErrorString(e));
    return ACGPU_OK;
}

// Tunables of the fused run (env: experiments, tools/gpu_r2o.sh)
struct FuseTuning {
    int ticket_rows = 16;   // rows of 256 positions per ticket
    int ring_mb = 64;       // ring of hit masks (stays in the 126 MB L2)
    int high_mb = 32;       // makers run at most this many MB of masks ahead of the expanders
    int64_t min_rows = 8192;
};
const FuseTuning &fuse_tuning() {
    static const FuseTuning t = [] {
        FuseTuning v;
        if (const char *e = getenv("ACGPU_FUSE_ROWS")) v.ticket_rows = std::min(32, std::max(1, atoi(e)));
        if (const char *e = getenv("ACGPU_FUSE_RING_MB")) v.ring_mb = std::max(1, atoi(e));
        if (const char *e = getenv("ACGPU_FUSE_HIGH_MB")) v.high_mb = std::max(1, atoi(e));
        if (const char *e = getenv("ACGPU_FUSE_MIN_ROWS")) v.min_rows = std::max<int64_t>(1, atoll(e));
        return v;
    }();
    return t;
}

cudaError_t launch_fuse(Matcher *m, const MaskArgs &P, const FuseArgs &F, int grid, cudaStream_t st) {
    const int low = mask_low_variant(m->tier, false);
    const bool is_map = m->dev.is_map != 0;
    switch (m->tier.K) {
    case 1: return fuse_launch_1(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 2: return fuse_launch_2(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 3: return fuse_launch_3(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 4: return fuse_launch_4(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 5: return fuse_launch_5(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 6: return fuse_launch_6(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    case 7: return fuse_launch_7(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    default: return fuse_launch_8(low, is_map, m->dev, m->tier, P, F, grid, m->fuse_smem, st);
    }
}

// k_tier_fused: one persistent launch makes the hit masks and expands them into records (kernel_fuse.cuh)
int enqueue_fused(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int64_t origin, int2 *d_pos,
                  uint32_t *d_val, int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    const FuseTuning &tune = fuse_tuning();
    const int64_t n_rows = (emit_to - origin + kMaskRow - 1) / kMaskRow;
    const int64_t tr = tune.ticket_rows;
    const int64_t n_tickets = (n_rows + tr - 1) / tr;
    const int64_t ticket_bytes = tr * kMaskRow * 2;
    int64_t ring = 1;
    while (ring * 2 * ticket_bytes <= static_cast<int64_t>(tune.ring_mb) << 20) ring *= 2;
    while (ring / 2 >= n_tickets && ring > 1) ring /= 2;  // short runs: no more slots than tickets (rounded up to a power of two)
    const int64_t high = std::max<int64_t>(1, std::min<int64_t>(ring, (static_cast<int64_t>(tune.high_mb) << 20) / ticket_bytes));
    const int64_t n_pad = (n_tickets + kFuseScanStep - 1) / kFuseScanStep * kFuseScanStep;
    Scratch S;
    const size_t o_ab = S.reserve(256);
    const size_t o_agg = S.reserve(static_cast<size_t>(n_pad) * 4);
    const size_t o_excl = S.reserve(static_cast<size_t>(n_pad) * 8);
    const size_t o_exp = S.reserve(static_cast<size_t>(n_tickets) * 4);
    const size_t o_zero_end = S.off;
    const size_t o_ring = S.reserve(static_cast<size_t>(ring * ticket_bytes));
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *w = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    if (cudaMemsetAsync(w, 0, o_zero_end, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    if (rc == ACGPU_OK) {
        MaskArgs P{};
        P.hay = d_hay;
        P.n = n;
        P.emit_from = emit_from;
        P.emit_to = emit_to;
        P.origin = origin;
        P.masks = reinterpret_cast<uint32_t *>(w + o_ring);
        P.n_rows = n_rows;
        FuseArgs F{};
        F.E.hay = d_hay;
        F.E.n = n;
        F.E.n_rows = n_rows;
        F.E.origin = origin;
        F.E.pos_base = opt.pos_base;
        F.E.pos_out = d_pos;
        F.E.val_out = d_val;
        F.E.cap = cap;
        F.agg = reinterpret_cast<uint32_t *>(w + o_agg);
        F.excl = reinterpret_cast<unsigned long long *>(w + o_excl);
        F.expanded = reinterpret_cast<uint32_t *>(w + o_exp);
        F.ab = reinterpret_cast<unsigned long long *>(w + o_ab);
        F.total_out = d_total;
        F.n_tickets = static_cast<uint32_t>(n_tickets);
        F.ticket_rows = static_cast<uint32_t>(tr);
        F.ring_tickets = static_cast<uint32_t>(ring);
        F.high = static_cast<uint32_t>(high);
        const int grid = static_cast<int>(std::min<int64_t>((n_tickets + kMaskWarps - 1) / kMaskWarps + 1, m->sm_count));
        const cudaError_t e = launch_fuse(m, P, F, grid, st);
        if (e != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("k_tier_fused: ") + cudaGetErrorString(e));
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// Tunables of the slab run (k_tier_duo, kernel_fuse.cuh).  ACGPU_DUO=1 turns it on.
struct DuoTuning {
    bool on = false;
    int slabs = 4;
    int emit_warps = 8;
    int64_t min_rows = 8 * kScanRows;
};
const DuoTuning &duo_tuning() {
    static const DuoTuning t = [] {
        DuoTuning v;
        if (const char *e = getenv("ACGPU_DUO")) v.on = e[0] == '1';
        if (const char *e = getenv("ACGPU_DUO_SLABS")) v.slabs = std::min(64, std::max(2, atoi(e)));
        if (const char *e = getenv("ACGPU_DUO_EMIT_WARPS")) v.emit_warps = std::min(kMaskWarps, std::max(1, atoi(e)));
        if (const char *e = getenv("ACGPU_DUO_MIN_ROWS")) v.min_rows = std::max<int64_t>(2 * kScanRows, atoll(e));
        return v;
    }();
    return t;
}

cudaError_t launch_duo(Matcher *m, const MaskArgs &P, const DuoArgs &D, int grid, size_t smem, cudaStream_t st) {
    const int low = mask_low_variant(m->tier, false);
    const bool is_map = m->dev.is_map != 0;
    switch (m->tier.K) {
    case 1: return duo_launch_1(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 2: return duo_launch_2(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 3: return duo_launch_3(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 4: return duo_launch_4(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 5: return duo_launch_5(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 6: return duo_launch_6(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    case 7: return duo_launch_7(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    default: return duo_launch_8(low, is_map, m->dev, m->tier, P, D, grid, smem, st);
    }
}

// Slab run: launch i = k_tier_duo(masks of slab i || records of slab i - 1), k_row_scan(slab i, running total); the last
// slab's records by k_tier_emit.  Slabs are multiples of kScanRows rows, so every slab has its own scan blocks.
int enqueue_duo(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int64_t origin, int2 *d_pos,
                uint32_t *d_val, int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt, size_t smem) {
    const DuoTuning &tune = duo_tuning();
    const MaskWs L = mask_ws_layout(emit_to, origin);
    const int64_t slab_rows = (((L.n_rows + tune.slabs - 1) / tune.slabs) + kScanRows - 1) / kScanRows * kScanRows;
    const int n_slabs = static_cast<int>((L.n_rows + slab_rows - 1) / slab_rows);
    const size_t ctr_bytes = static_cast<size_t>(n_slabs) * 64;   // per slab: mask ticket, emit ticket, scan "done", total (u64 at +16)
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, L.bytes + ctr_bytes, st));
    char *w = static_cast<char *>(ws);
    char *ctr = w + L.bytes;
    int rc = ACGPU_OK;
    if (cudaMemsetAsync(ctr, 0, ctr_bytes, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    uint32_t *masks = reinterpret_cast<uint32_t *>(w + L.o_mask);
    uint32_t *row_count = reinterpret_cast<uint32_t *>(w + L.o_cnt);
    unsigned long long *block_excl = reinterpret_cast<unsigned long long *>(w + L.o_blk);
    auto emit_args = [&](int slab) {
        EmitArgs E{};
        const int64_t row_lo = static_cast<int64_t>(slab) * slab_rows;
        E.hay = d_hay;
        E.n = n;
        E.masks = masks + row_lo * (kMaskRow / 2);
        E.row_excl = row_count + row_lo;
        E.block_excl = block_excl + row_lo / kScanRows;
        E.n_rows = std::min(slab_rows, L.n_rows - row_lo);
        E.origin = origin + row_lo * kMaskRow;
        E.pos_base = opt.pos_base;
        E.pos_out = d_pos;
        E.val_out = d_val;
        E.cap = cap;
        return E;
    };
    for (int i = 0; i < n_slabs && rc == ACGPU_OK; i++) {
        const int64_t row_lo = static_cast<int64_t>(i) * slab_rows;
        const int64_t rows = std::min(slab_rows, L.n_rows - row_lo);
        MaskArgs P{};
        P.hay = d_hay;
        P.n = n;
        P.emit_from = emit_from;
        P.emit_to = emit_to;
        P.origin = origin + row_lo * kMaskRow;
        P.masks = masks + row_lo * (kMaskRow / 2);
        P.row_count = row_count + row_lo;
        P.ticket = reinterpret_cast<unsigned int *>(ctr + i * 64);
        P.n_rows = rows;
        DuoArgs D{};
        if (i > 0) D.E = emit_args(i - 1);
        D.emit_ticket = reinterpret_cast<unsigned int *>(ctr + (i > 0 ? i - 1 : 0) * 64 + 4);
        D.emit_warps = tune.emit_warps;
        const cudaError_t e = launch_duo(m, P, D, m->sm_count, smem, st);
        if (e != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("k_tier_duo: ") + cudaGetErrorString(e));
        if (rc != ACGPU_OK) break;
        ScanArgs SA{};
        SA.row_count = P.row_count;
        SA.block_excl = block_excl + row_lo / kScanRows;
        SA.done = reinterpret_cast<unsigned int *>(ctr + i * 64 + 8);
        SA.total_out = reinterpret_cast<unsigned long long *>(ctr + i * 64 + 16);
        SA.n_rows = rows;
        SA.base_in = i > 0 ? reinterpret_cast<const unsigned long long *>(ctr + (i - 1) * 64 + 16) : nullptr;
        k_row_scan<<<static_cast<unsigned>((rows + kScanRows - 1) / kScanRows), 1024, 0, st>>>(SA);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ACGPU_ECUDA, "k_row_scan failed");
    }
    if (rc == ACGPU_OK) rc = launch_emit(m, emit_args(n_slabs - 1), st);
    if (rc == ACGPU_OK && cudaMemcpyAsync(d_total, ctr + (n_slabs - 1) * 64 + 16, 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = fail(ACGPU_ECUDA, "total copy failed");
    cudaFreeAsync(ws, st);
    return rc;
}

int enqueue_mask(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int64_t origin, int2 *d_pos,
                 uint32_t *d_val, int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    const MaskWs whole = mask_ws_layout(emit_to, origin);
    if (m->fuse_smem && cap > 0 && whole.n_rows >= fuse_tuning().min_rows)
        return enqueue_fused(m, d_hay, n, emit_from, emit_to, origin, d_pos, d_val, cap, d_total, st, opt);
    if (duo_tuning().on && !m->mask_pair && cap > 0 && whole.n_rows >= duo_tuning().min_rows) {
        const size_t smem = duo_smem_bytes(m->host.tier.row_words.size(), m->dev.is_map != 0, duo_tuning().emit_warps);
        if (smem <= 227 * 1024) return enqueue_duo(m, d_hay, n, emit_from, emit_to, origin, d_pos, d_val, cap, d_total, st, opt, smem);
    }
    const MaskWs &L = whole;
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, L.bytes, st));
    char *w = static_cast<char *>(ws);
    int rc = enqueue_mask_scan(m, d_hay, n, emit_from, emit_to, origin, w, L, d_total, st);
    if (rc == ACGPU_OK && cap > 0) {
        EmitArgs E{};
        fill_emit_args(E, d_hay, n, w, L, origin, d_pos, d_val, cap, opt);
        rc = launch_emit(m, E, st);
    }
    cudaFreeAsync(ws, st);  // on every path (no scratch leak when a launch fails)
    return rc;
}

// AhoCorasick family outside the tier envelope: 32-bit hit masks by anchored walks -> row-count scan -> records (kernel_wide.cuh)
int enqueue_wide(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int64_t origin, int2 *d_pos,
                 uint32_t *d_val, int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    const int64_t n_rows = (emit_to - origin + kMaskRow - 1) / kMaskRow;
    const int64_t n_blocks = (n_rows + kScanRows - 1) / kScanRows;
    Scratch S;
    const size_t o_ctr = S.reserve(256);
    const size_t o_cnt = S.reserve(static_cast<size_t>(n_rows) * 4);
    const size_t o_blk = S.reserve(static_cast<size_t>(n_blocks) * 8);
    const size_t o_mask = S.reserve(static_cast<size_t>(n_rows) * kMaskRow * 4);
    // k_wide_tile hands the walks of its thin late rounds to k_wide_tail: room for one walk per four positions (a tile
    // that finds the list full finishes its walks itself)
    static const char *tail_env = getenv("ACGPU_WIDE_TAIL");
    int64_t tail_cap = (m->wide_tile && !(tail_env && tail_env[0] == '0')) ? std::max<int64_t>(4096, n_rows * kMaskRow / 4) : 0;
    if (const char *tc = getenv("ACGPU_WIDE_TAIL_CAP")) tail_cap = tail_cap ? std::max<int64_t>(1, atoll(tc)) : 0;  // tests: force the "list full" path
    const size_t o_tail = S.reserve(static_cast<size_t>(tail_cap) * 16);
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *w = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    auto launch_ok = [&](const char *what) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (cudaMemsetAsync(w + o_ctr, 0, 256, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    if (rc == ACGPU_OK) {
        WideArgs P{};
        P.hay = d_hay;
        P.n = n;
        P.emit_from = emit_from;
        P.emit_to = emit_to;
        P.origin = origin;
        P.masks = reinterpret_cast<uint32_t *>(w + o_mask);
        P.row_count = reinterpret_cast<uint32_t *>(w + o_cnt);
        P.ticket = reinterpret_cast<unsigned int *>(w + o_ctr);
        P.n_rows = n_rows;
        P.tail = reinterpret_cast<uint4 *>(w + o_tail);
        P.tail_count = reinterpret_cast<unsigned int *>(w + o_ctr + 128);
        P.tail_cap = static_cast<uint32_t>(std::min<int64_t>(tail_cap, 0x7FFFFFFF));
        const char *mr = getenv("ACGPU_WT_MIN_ROUNDS"), *ho = getenv("ACGPU_WT_HANDOVER");  // tuning runs and tests
        P.min_rounds = mr ? static_cast<uint32_t>(atoi(mr)) : kWtMinRounds;
        P.hand_over = ho ? static_cast<uint32_t>(atoi(ho)) : kWtHandOver;
        const int64_t n_chunks = (n_rows + kMaskChunkRows - 1) / kMaskChunkRows;
        const int per_sm = std::max<int>(1, std::min<int>(6, static_cast<int>((200 * 1024) / std::max<size_t>(m->wide_smem, 1))));
        const int grid = static_cast<int>(std::min<int64_t>((n_chunks + kWideWarps - 1) / kWideWarps, static_cast<int64_t>(m->sm_count) * per_sm));
        if (m->wide_tile) {
            const int64_t n_tiles = (n_rows + kWtRows - 1) / kWtRows;
            const int tgrid = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(m->sm_count) * 2));
            k_wide_tile<<<tgrid, kWtThreads, wide_tile_smem_bytes(m->wide.C), st>>>(m->dev, m->wide, P);
            if (P.tail_cap) {
                launch_ok("k_wide_tile");
                k_wide_tail<<<m->sm_count * 8, 256, 0, st>>>(m->dev, m->wide, P);
            }
        } else if (m->wide.pair)
            k_wide_mask<true><<<grid, kWideThreads, m->wide_smem, st>>>(m->dev, m->wide, P);
        else
            k_wide_mask<false><<<grid, kWideThreads, m->wide_smem, st>>>(m->dev, m->wide, P);
        launch_ok("k_wide_mask");
    }
    if (rc == ACGPU_OK) {
        ScanArgs SA{};
        SA.row_count = reinterpret_cast<uint32_t *>(w + o_cnt);
        SA.block_excl = reinterpret_cast<unsigned long long *>(w + o_blk);
        SA.done = reinterpret_cast<unsigned int *>(w + o_ctr + 64);
        SA.total_out = d_total;
        SA.n_rows = n_rows;
        k_row_scan<<<static_cast<unsigned>(n_blocks), 1024, 0, st>>>(SA);
        launch_ok("k_row_scan");
    }
    if (rc == ACGPU_OK && cap > 0) {
        EmitArgs E{};
        E.hay = d_hay;
        E.n = n;
        E.masks = reinterpret_cast<uint32_t *>(w + o_mask);
        E.row_excl = reinterpret_cast<uint32_t *>(w + o_cnt);
        E.block_excl = reinterpret_cast<unsigned long long *>(w + o_blk);
        E.n_rows = n_rows;
        E.origin = origin;
        E.pos_base = opt.pos_base;
        E.pos_out = d_pos;
        E.val_out = d_val;
        E.cap = cap;
        const int egrid = static_cast<int>(std::min<int64_t>((n_rows + kEmitWarps - 1) / kEmitWarps, static_cast<int64_t>(m->sm_count) * 128));
        if (m->dev.is_map)
            k_wide_emit<true><<<egrid, kEmitWarps * 32, 0, st>>>(m->dev, m->wide, E);
        else
            k_wide_emit<false><<<egrid, kEmitWarps * 32, 0, st>>>(m->dev, m->wide, E);
        launch_ok("k_wide_emit");
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// WholeWord, generation 3 (kernel_ww3.cuh): hit bits + row counts by atomics on zeroed scratch, row scan, records
int enqueue_ww3(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t dom_lo, int64_t dom_hi, int64_t origin, int2 *d_pos, uint32_t *d_val,
                int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    const int64_t last = std::min<int64_t>(n, dom_hi + m->ww.max_len);   // the last position a reported run can end at (exclusive end)
    const int64_t n_rows = (last - origin) / kW3Row + 1;
    const int64_t n_blocks = (n_rows + kScanRows - 1) / kScanRows;
    Scratch S;
    const size_t o_ctr = S.reserve(256);
    const size_t o_bits = S.reserve(static_cast<size_t>(n_rows) * 32);
    const size_t o_start = S.reserve(static_cast<size_t>(n_rows) * 32);
    const size_t zeroed = S.off;
    const size_t o_cnt = S.reserve(static_cast<size_t>(n_rows) * 4);
    const size_t o_val = S.reserve(m->dev.is_map && cap > 0 ? static_cast<size_t>(n_rows) * (kW3Row / 2) * 4 : 0);   // written at hits only
    const size_t o_blk = S.reserve(static_cast<size_t>(n_blocks) * 8);
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *w = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    auto launch_ok = [&](const char *what) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (cudaMemsetAsync(w, 0, zeroed, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    if (rc == ACGPU_OK) {
        Ww3Args P{};
        P.hay = d_hay;
        P.n = n;
        P.dom_lo = dom_lo;
        P.dom_hi = dom_hi;
        P.origin = origin;
        P.n_rows = n_rows;
        P.hitbits = reinterpret_cast<uint32_t *>(w + o_bits);
        P.startbits = reinterpret_cast<uint32_t *>(w + o_start);
        P.val_scratch = (m->dev.is_map && cap > 0) ? reinterpret_cast<uint32_t *>(w + o_val) : nullptr;
        P.ticket = reinterpret_cast<unsigned int *>(w + o_ctr);
        P.chunk_rows = kW3ChunkRows;
        while (P.chunk_rows > 4 && (n_rows + P.chunk_rows - 1) / P.chunk_rows < 2 * static_cast<int64_t>(m->sm_count) * kW3Warps) P.chunk_rows >>= 1;
        const int64_t n_chunks = (n_rows + P.chunk_rows - 1) / P.chunk_rows;
        const int grid = static_cast<int>(std::min<int64_t>((n_chunks + kW3Warps - 1) / kW3Warps, m->sm_count));
        const bool shortk = m->ww.max_len < 32;
        const size_t smem = ww3_smem_bytes(m->ww.bloom_bits, shortk);
        if (m->ww.bloom_bits)
            shortk ? k_ww3_hits<true, true><<<grid, kW3Warps * 32, smem, st>>>(m->ww, P) : k_ww3_hits<true, false><<<grid, kW3Warps * 32, smem, st>>>(m->ww, P);
        else
            shortk ? k_ww3_hits<false, true><<<grid, kW3Warps * 32, smem, st>>>(m->ww, P) : k_ww3_hits<false, false><<<grid, kW3Warps * 32, smem, st>>>(m->ww, P);
        launch_ok("k_ww3_hits");
    }
    if (rc == ACGPU_OK) {
        const int cgrid = static_cast<int>(std::min<int64_t>((n_rows + 255) / 256, static_cast<int64_t>(m->sm_count) * 8));
        k_ww3_count<<<cgrid, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(w + o_bits), reinterpret_cast<uint32_t *>(w + o_cnt), n_rows);
        launch_ok("k_ww3_count");
    }
    if (rc == ACGPU_OK) {
        ScanArgs SA{};
        SA.row_count = reinterpret_cast<uint32_t *>(w + o_cnt);
        SA.block_excl = reinterpret_cast<unsigned long long *>(w + o_blk);
        SA.done = reinterpret_cast<unsigned int *>(w + o_ctr + 64);
        SA.total_out = d_total;
        SA.n_rows = n_rows;
        k_row_scan<<<static_cast<unsigned>(n_blocks), 1024, 0, st>>>(SA);
        launch_ok("k_row_scan");
    }
    if (rc == ACGPU_OK && cap > 0) {
        Ww3EmitArgs E{};
        E.hay = d_hay;
        E.n = n;
        E.origin = origin;
        E.n_rows = n_rows;
        E.hitbits = reinterpret_cast<const uint32_t *>(w + o_bits);
        E.startbits = reinterpret_cast<const uint32_t *>(w + o_start);
        E.val_scratch = m->dev.is_map ? reinterpret_cast<const uint32_t *>(w + o_val) : nullptr;
        E.row_excl = reinterpret_cast<const uint32_t *>(w + o_cnt);
        E.block_excl = reinterpret_cast<const unsigned long long *>(w + o_blk);
        E.pos_base = opt.pos_base;
        E.pos_out = d_pos;
        E.val_out = d_val;
        E.cap = cap;
        const int grid = static_cast<int>(std::min<int64_t>((n_rows + 255) / 256, static_cast<int64_t>(m->sm_count) * 8));
        if (m->dev.is_map)
            k_ww3_emit<true><<<grid, 256, 0, st>>>(m->ww, E);
        else
            k_ww3_emit<false><<<grid, 256, 0, st>>>(m->ww, E);
        launch_ok("k_ww3_emit");
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// WholeWord, case-insensitive with a word-character table that is not closed under toLowerCase (quirk Q7): the reference's
// loop, literally, one thread per synchronisation point (kernel_wwlit.cuh): count, scan, write
int enqueue_ww_literal(Matcher *m, const uint16_t *d_hay, int64_t n, int2 *d_pos, uint32_t *d_val, int64_t cap, unsigned long long *d_total,
                       cudaStream_t st, const RunOpts &opt) {
    const int64_t n_rows = (n + kMaskRow - 1) / kMaskRow;
    const int64_t n_blocks = (n_rows + kScanRows - 1) / kScanRows;
    Scratch S;
    const size_t o_ctr = S.reserve(256);
    const size_t o_cnt = S.reserve(static_cast<size_t>(n_rows) * 4);
    const size_t o_blk = S.reserve(static_cast<size_t>(n_blocks) * 8);
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *w = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    auto launch_ok = [&](const char *what) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (cudaMemsetAsync(w + o_ctr, 0, 256, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    WwLitArgs P{};
    P.hay = d_hay;
    P.n = n;
    P.n_rows = n_rows;
    P.row_count = reinterpret_cast<uint32_t *>(w + o_cnt);
    P.row_excl = P.row_count;
    P.block_excl = reinterpret_cast<unsigned long long *>(w + o_blk);
    P.pos_base = opt.pos_base;
    P.scroll_folded = opt.folded_scroll ? 1 : 0;
    P.pos_out = d_pos;
    P.val_out = d_val;
    P.cap = cap;
    const int grid = static_cast<int>(std::min<int64_t>(n_rows, static_cast<int64_t>(m->sm_count) * 8));
    auto launch = [&](auto write) {
        constexpr bool kWrite = decltype(write)::value;
        const bool is_map = m->dev.is_map != 0;
        switch (m->literal_family) {
        case 0: is_map ? k_segments<0, kWrite, true><<<grid, kMaskRow, 0, st>>>(m->dev, P) : k_segments<0, kWrite, false><<<grid, kMaskRow, 0, st>>>(m->dev, P); break;
        case 1: is_map ? k_segments<1, kWrite, true><<<grid, kMaskRow, 0, st>>>(m->dev, P) : k_segments<1, kWrite, false><<<grid, kMaskRow, 0, st>>>(m->dev, P); break;
        case 2: is_map ? k_segments<2, kWrite, true><<<grid, kMaskRow, 0, st>>>(m->dev, P) : k_segments<2, kWrite, false><<<grid, kMaskRow, 0, st>>>(m->dev, P); break;
        default: is_map ? k_segments<4, kWrite, true><<<grid, kMaskRow, 0, st>>>(m->dev, P) : k_segments<4, kWrite, false><<<grid, kMaskRow, 0, st>>>(m->dev, P); break;
        }
    };
    if (rc == ACGPU_OK) {
        launch(std::false_type{});
        launch_ok("k_segments (count)");
    }
    if (rc == ACGPU_OK) {
        ScanArgs SA{};
        SA.row_count = P.row_count;
        SA.block_excl = reinterpret_cast<unsigned long long *>(w + o_blk);
        SA.done = reinterpret_cast<unsigned int *>(w + o_ctr + 64);
        SA.total_out = d_total;
        SA.n_rows = n_rows;
        k_row_scan<<<static_cast<unsigned>(n_blocks), 1024, 0, st>>>(SA);
        launch_ok("k_row_scan");
    }
    if (rc == ACGPU_OK && cap > 0) {
        launch(std::true_type{});
        launch_ok("k_segments (write)");
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// The generation-1 selection passes over per-start values v (k_sel_map / group / top / entries / emit): P comes with the
// values, the chain domain, the mode and the outputs filled in; the tile scratch is allocated here.
int enqueue_selection(Matcher *m, SelArgs P, bool chain, cudaStream_t st) {
    const DevAutomaton &A = m->dev;
    const int64_t n_tiles = (P.n + kSelTile - 1) / kSelTile;
    const int64_t n_groups = (n_tiles + kSelGroup - 1) / kSelGroup;
    if (n_tiles == 0) {
        CU_TRY(cudaMemsetAsync(P.total_out, 0, sizeof(unsigned long long), st));
        return ACGPU_OK;
    }
    Scratch S;
    size_t o_ctr = S.reserve(256);
    size_t o_status = S.reserve(static_cast<size_t>(n_tiles) * 8);
    size_t o_zero_end = S.off;  // everything up to here is zero-initialised
    size_t o_exit1 = 0, o_exit2 = 0, o_entry2 = 0, o_entry1 = 0;
    if (chain) {
        o_exit1 = S.reserve(static_cast<size_t>(n_tiles) * P.M * sizeof(uint16_t));
        o_exit2 = S.reserve(static_cast<size_t>(n_groups) * P.M * sizeof(uint16_t));
        o_entry2 = S.reserve(static_cast<size_t>(n_groups) * sizeof(int64_t));
        o_entry1 = S.reserve(static_cast<size_t>(n_tiles) * sizeof(int32_t));
    }
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *b = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    auto launch_ok = [&](const char *what) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    if (cudaMemsetAsync(ws, 0, o_zero_end, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    P.n_tiles = n_tiles;
    P.n_groups = n_groups;
    if (P.carry_out && rc == ACGPU_OK && cudaMemsetAsync(P.carry_out, 0xFF, 8, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");  // -1 = chain did not cross chain_n
    P.tile_counter = reinterpret_cast<unsigned int *>(b + o_ctr);
    P.status = reinterpret_cast<unsigned long long *>(b + o_status);
    if (chain && rc == ACGPU_OK) {
        P.exit1 = reinterpret_cast<uint16_t *>(b + o_exit1);
        P.exit2 = reinterpret_cast<uint16_t *>(b + o_exit2);
        P.entry2 = reinterpret_cast<int64_t *>(b + o_entry2);
        P.entry1 = reinterpret_cast<int32_t *>(b + o_entry1);
        const int grid_map = static_cast<int>(std::min<int64_t>(n_tiles, m->sm_count * 4));
        k_sel_map<<<grid_map, kThreads, 0, st>>>(P);
        launch_ok("k_sel_map");
        k_sel_group<<<static_cast<unsigned>(n_groups), kThreads, 0, st>>>(P);
        launch_ok("k_sel_group");
        k_sel_top<<<1, 32, 0, st>>>(P);
        launch_ok("k_sel_top");
        k_sel_entries<<<static_cast<unsigned>((n_groups + 127) / 128), 128, 0, st>>>(P);
        launch_ok("k_sel_entries");
    }
    if (rc == ACGPU_OK) {
        const int grid = static_cast<int>(std::min<int64_t>(n_tiles, m->sm_count * 4));
        if (A.is_map)
            k_sel_emit<true><<<grid, kThreads, kSelEmitSmem, st>>>(A, P);
        else
            k_sel_emit<false><<<grid, kThreads, kSelEmitSmem, st>>>(A, P);
        launch_ok("k_sel_emit");
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// WholeWordLongest with phrase keywords, one-shot: walk starts compacted by k_wwl_starts (kernel_ww.cuh), then the
// selection passes over the compacted starts (one position in ~six) instead of every haystack position
int enqueue_wwl2(Matcher *m, const uint16_t *d_hay, int64_t n, int2 *d_pos, uint32_t *d_val, int64_t cap, unsigned long long *d_total,
                 cudaStream_t st, const RunOpts &opt) {
    const DevAutomaton &A = m->dev;
    const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(d_hay) >> 1) & 7);
    const int64_t origin = -mis;  // <= 0, hay + origin 16-byte aligned
    const int64_t n_tiles = (n - origin + kWwTile - 1) / kWwTile;
    const int64_t cap_m = n / 2 + 2;  // walk starts are at least two positions apart (+ the first char of the input)
    Scratch S;
    const size_t o_ctr = S.reserve(256);
    const size_t o_status = S.reserve(static_cast<size_t>(n_tiles) * 8);
    const size_t o_zero_end = S.off;
    const size_t o_wpos = S.reserve(static_cast<size_t>(cap_m) * 4);
    const size_t o_v = S.reserve(static_cast<size_t>(cap_m) * 2);
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *b = static_cast<char *>(ws);
    int rc = ACGPU_OK;
    if (cudaMemsetAsync(ws, 0, o_zero_end, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    unsigned long long n_starts = 0;
    if (rc == ACGPU_OK) {
        WwlArgs W{};
        W.hay = d_hay;
        W.n = n;
        W.origin = origin;
        W.n_tiles = n_tiles;
        W.wpos = reinterpret_cast<int32_t *>(b + o_wpos);
        W.v = reinterpret_cast<uint16_t *>(b + o_v);
        W.cap = cap_m;
        W.total_out = reinterpret_cast<unsigned long long *>(b + o_ctr + 64);
        W.tile_counter = reinterpret_cast<unsigned int *>(b + o_ctr);
        W.status = reinterpret_cast<unsigned long long *>(b + o_status);
        DevWw T{};
        T.wcls = m->d_wwl_wcls;
        T.max_len = A.max_len;
        const size_t smem = static_cast<size_t>(kWwTile + 16 * ((A.max_len + 1 + 15) / 16) + 2) * 2;
        const int grid = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(m->sm_count) * 6));
        k_wwl_starts<<<grid, kWwThreads, smem, st>>>(A, T, W);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ACGPU_ECUDA, "k_wwl_starts failed to launch");
        // the selection grid depends on the number of walk starts: one small read-back in the middle of the call
        if (rc == ACGPU_OK && (cudaMemcpyAsync(&n_starts, W.total_out, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                               cudaStreamSynchronize(st) != cudaSuccess))
            rc = fail(ACGPU_ECUDA, "k_wwl_starts failed");
    }
    if (rc == ACGPU_OK) {
        SelArgs P{};
        P.v = reinterpret_cast<const uint16_t *>(b + o_v);
        P.n_v = static_cast<int64_t>(n_starts);
        P.n = static_cast<int64_t>(n_starts);
        P.M = A.max_len + 2;
        P.halo = std::max(0, A.max_len - 1);
        P.dom_lo = 0;
        P.mode = kModeWholeWordLongest;
        P.entry0 = 0;
        P.carry_out = nullptr;
        P.wpos = reinterpret_cast<const int32_t *>(b + o_wpos);
        P.hay = d_hay;
        P.n_hay = n;
        P.pos_base = opt.pos_base;
        P.pos_out = d_pos;
        P.val_out = d_val;
        P.cap = cap;
        P.total_out = d_total;
        rc = enqueue_selection(m, P, true, st);
    }
    cudaFreeAsync(ws, st);
    return rc;
}

// Enqueue every kernel of one match on `st`.  d_total receives the total number of matches.
int enqueue_match(Matcher *m, const uint16_t *d_hay, int64_t n, int64_t emit_from, int64_t emit_to, int2 *d_pos,
                  uint32_t *d_val, int64_t cap, unsigned long long *d_total, cudaStream_t st, const RunOpts &opt) {
    const DevAutomaton &A = m->dev;
    if (n < 0 || n > 0x7FFFFFFFll) return fail(ACGPU_EINVAL, "haystack length must fit a Java int");
    if (A.is_map && !d_val && cap > 0) return fail(ACGPU_EINVAL, "Map matcher needs a value buffer");
    const int persistent = m->sm_count * 8;

    if (A.family == ACGPU_AHOCORASICK) {
        emit_from = std::max<int64_t>(0, emit_from);
        emit_to = std::min<int64_t>(n, emit_to);
        // k_ac_tier rows start at `origin` <= emit_from, placed so that every lane's 8-char load is 16-byte aligned
        const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(d_hay) >> 1) & 7);
        const int64_t origin = (m->use_tier || m->use_wide) ? ((emit_from + mis) & ~int64_t(7)) - mis : emit_from;
        const int64_t span = std::max<int64_t>(0, emit_to - origin);
        const int64_t tile_sz = kAcTile;
        const int64_t n_tiles = emit_to > emit_from ? (span + tile_sz - 1) / tile_sz : 0;
        if (n_tiles == 0) {
            CU_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
            return ACGPU_OK;
        }
        if (m->use_tier) return enqueue_mask(m, d_hay, n, emit_from, emit_to, origin, d_pos, d_val, cap, d_total, st, opt);
        if (m->use_wide) return enqueue_wide(m, d_hay, n, emit_from, emit_to, origin, d_pos, d_val, cap, d_total, st, opt);
        size_t bytes = 256 + static_cast<size_t>(n_tiles) * 8;
        void *ws = nullptr;
        CU_TRY(cudaMallocAsync(&ws, bytes, st));
        CU_TRY(cudaMemsetAsync(ws, 0, bytes, st));
        AcArgs P{};
        P.hay = d_hay;
        P.n = n;
        P.emit_from = emit_from;
        P.emit_to = emit_to;
        P.origin = origin;
        P.pos_base = opt.pos_base;
        P.pos_out = d_pos;
        P.val_out = d_val;
        P.cap = cap;
        P.total_out = d_total;
        P.tile_counter = static_cast<unsigned int *>(ws);
        P.status = reinterpret_cast<unsigned long long *>(static_cast<char *>(ws) + 256);
        P.n_tiles = n_tiles;
        const int grid = static_cast<int>(std::min<int64_t>(n_tiles, persistent));
        if (A.is_map)
            k_ac_scan<true><<<grid, kThreads, 0, st>>>(A, P);
        else
            k_ac_scan<false><<<grid, kThreads, 0, st>>>(A, P);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaFreeAsync(ws, st));
        return ACGPU_OK;
    }

    // ---- start-anchored families
    const int64_t chain_n = opt.chain_n < 0 ? n : opt.chain_n;
    if (n == 0 || chain_n == 0) {
        CU_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
        return ACGPU_OK;
    }
    if (m->literal_family >= 0) {
        if (opt.ctx != 0 || chain_n != n) return fail(ACGPU_EINVAL, "literal matchers scan whole haystacks");
        return enqueue_ww_literal(m, d_hay, n, d_pos, d_val, cap, d_total, st, opt);
    }
    const bool chain = A.family != ACGPU_WHOLEWORD && !(A.family == ACGPU_WHOLEWORDLONGEST && m->use_ww);
    if (m->use_ww) {
        // WholeWord: one launch (kernel_ww.cuh); words starting in [ctx, chain_n) are reported
        const int64_t dom_lo = opt.ctx, dom_hi = chain_n;
        const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(d_hay) >> 1) & 7);
        const int64_t origin = ((dom_lo + mis) & ~int64_t(7)) - mis;  // <= dom_lo, hay + origin 16-byte aligned
        const int64_t n_tiles = dom_hi > dom_lo ? (dom_hi - origin + kWwTile - 1) / kWwTile : 0;
        if (n_tiles == 0) {
            CU_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
            return ACGPU_OK;
        }
        if (m->use_ww3) return enqueue_ww3(m, d_hay, n, dom_lo, dom_hi, origin, d_pos, d_val, cap, d_total, st, opt);
        const size_t bytes = 256 + static_cast<size_t>(n_tiles) * 8;
        void *ws = nullptr;
        CU_TRY(cudaMallocAsync(&ws, bytes, st));
        CU_TRY(cudaMemsetAsync(ws, 0, bytes, st));
        WwArgs W{};
        W.hay = d_hay;
        W.n = n;
        W.dom_lo = dom_lo;
        W.dom_hi = dom_hi;
        W.origin = origin;
        W.n_tiles = n_tiles;
        W.pos_base = opt.pos_base;
        W.pos_out = d_pos;
        W.val_out = d_val;
        W.cap = cap;
        W.total_out = d_total;
        W.tile_counter = static_cast<unsigned int *>(ws);
        W.status = reinterpret_cast<unsigned long long *>(static_cast<char *>(ws) + 256);
        const size_t smem = static_cast<size_t>(kWwTile + 16 * ((m->ww.max_len + 1 + 15) / 16) + 2) * 2;  // + one word read ahead by the pair hash
        const int grid = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(m->sm_count) * 6));
        if (A.is_map)
            k_ww_scan<true><<<grid, kWwThreads, smem, st>>>(m->ww, W);
        else
            k_ww_scan<false><<<grid, kWwThreads, smem, st>>>(m->ww, W);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaFreeAsync(ws, st));
        return ACGPU_OK;
    }
    if (chain && m->use_tier && opt.ctx == 0 && chain_n == n && opt.entry0 == 0 && !opt.d_carry)
        return enqueue_sel2(m, d_hay, n, d_pos, d_val, cap, d_total, st, opt);
    if (m->use_wwl2 && opt.ctx == 0 && chain_n == n && opt.entry0 == 0 && !opt.d_carry)
        return enqueue_wwl2(m, d_hay, n, d_pos, d_val, cap, d_total, st, opt);
    const int64_t n_tiles = (chain_n + kSelTile - 1) / kSelTile;
    Scratch S;
    size_t o_v = S.reserve(static_cast<size_t>(n) * sizeof(uint16_t));
    void *ws = nullptr;
    CU_TRY(cudaMallocAsync(&ws, S.off, st));
    char *b = static_cast<char *>(ws);

    FwArgs F{};
    F.hay = d_hay;
    F.n = n;
    F.p_lo = opt.ctx;
    F.p_hi = n;
    F.v = reinterpret_cast<uint16_t *>(b + o_v);
    F.abs0 = opt.abs0;
    int rc = ACGPU_OK;
    if (opt.ctx > 0 && cudaMemsetAsync(F.v, 0, static_cast<size_t>(opt.ctx) * sizeof(uint16_t), st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "memset failed");
    if (rc == ACGPU_OK) {
        const int64_t ft = (n + kFwTile - 1) / kFwTile;
        const int grid = static_cast<int>(std::min<int64_t>(ft, persistent));
        if (A.family == ACGPU_LONGEST)
            k_fwd_v<1><<<grid, kThreads, 0, st>>>(A, F);
        else if (A.family == ACGPU_SHORTEST)
            k_fwd_v<2><<<grid, kThreads, 0, st>>>(A, F);
        else if (A.family == ACGPU_WHOLEWORDLONGEST)
            k_fwd_v<4><<<grid, kThreads, 0, st>>>(A, F);
        else
            k_fwd_v<3><<<grid, kThreads, 0, st>>>(A, F);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ACGPU_ECUDA, "k_fwd_v failed to launch");
    }
    if (rc == ACGPU_OK) {
        SelArgs P{};
        P.v = F.v;
        P.n_v = n;
        P.n = chain_n;
        P.M = A.max_len + (A.family == ACGPU_WHOLEWORDLONGEST ? 2 : 1);  // exit offsets are < M
        P.halo = std::max(0, A.max_len - 1);
        P.dom_lo = opt.ctx;
        P.mode = A.family == ACGPU_LONGEST ? kModeLongest
                 : (A.family == ACGPU_SHORTEST ? kModeShortest
                                               : (A.family == ACGPU_WHOLEWORDLONGEST ? kModeWholeWordLongest : kModeWholeWord));
        P.entry0 = opt.entry0;
        P.carry_out = reinterpret_cast<long long *>(opt.d_carry);
        P.hay = d_hay;
        P.n_hay = n;
        P.pos_base = opt.pos_base;
        P.pos_out = d_pos;
        P.val_out = d_val;
        P.cap = cap;
        P.total_out = d_total;
        (void)n_tiles;
        rc = enqueue_selection(m, P, chain, st);
    }
    cudaFreeAsync(ws, st);
    return rc;
}

int ensure_device(Matcher *m) {
    CU_TRY(cudaSetDevice(m->device));
    if (!m->sel_attr_set) {
        // keep stream-ordered scratch cached between calls instead of returning it to the driver at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, m->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        CU_TRY(cudaFuncSetAttribute(k_sel_emit<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelEmitSmem));
        CU_TRY(cudaFuncSetAttribute(k_sel_emit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelEmitSmem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_top, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        // static + dynamic shared memory of the selection kernels passes 48 KB
        const int s2_smem = kS2SmemWords * 4;
        CU_TRY(cudaFuncSetAttribute(k_sel2_map<kModeLongest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_map<kModeShortest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_emit<kModeLongest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_emit<kModeShortest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_fused<kModeLongest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        CU_TRY(cudaFuncSetAttribute(k_sel2_fused<kModeShortest>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2_smem));
        m->sel_attr_set = true;
    }
    return ACGPU_OK;
}

struct HostResult {  // owner block placed right before the arrays
    int32_t *pos;
    uint32_t *val;
};

void fill_empty(acgpu_result *out) {
    out->n = 0;
    out->pos = nullptr;
    out->val = nullptr;
}


// ---------------------------------------------------------------------------------------------------
// Host-buffer calls: pinned result blocks + a three-stream chunk pipeline.
//
// Match records leave the device through page-locked host memory (a pageable D2H copy runs at a fraction of the
// PCIe rate).  cudaHostAlloc is slow (hundreds of ms per GB), so released result blocks are parked in a small
// process-wide cache and reused by later calls; acgpu_free_result() hands a block back.
struct PinnedBlock {
    void *p = nullptr;
    size_t bytes = 0;
};
std::mutex g_pin_mu;
std::vector<PinnedBlock> g_pin_free;
std::vector<PinnedBlock> g_pin_live;
constexpr size_t kPinCacheBlocks = 4;

// sizes are rounded up to {1, 1.25, 1.5, 1.75} x 2^k so that the blocks of consecutive calls / feeds (whose record counts
// differ by a few percent) are interchangeable in the cache - cudaHostAlloc costs ~0.3 s per GB
size_t pin_bucket(size_t bytes) {
    bytes = std::max<size_t>(bytes, 1 << 20);
    size_t p2 = size_t(1) << 20;
    while (p2 * 2 <= bytes) p2 *= 2;
    for (int q = 4; q <= 8; q++) {
        const size_t b = p2 / 4 * q;
        if (b >= bytes) return b;
    }
    return p2 * 2;
}

bool pin_take(size_t bytes, PinnedBlock *out) {
    bytes = pin_bucket(bytes);
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        int best = -1;
        for (size_t i = 0; i < g_pin_free.size(); i++) {
            if (g_pin_free[i].bytes >= bytes && (best < 0 || g_pin_free[i].bytes < g_pin_free[best].bytes)) best = (int)i;
        }
        if (best >= 0) {
            *out = g_pin_free[best];
            g_pin_free.erase(g_pin_free.begin() + best);
            g_pin_live.push_back(*out);
            return true;
        }
    }
    PinnedBlock b;
    b.bytes = bytes;
    if (cudaHostAlloc(&b.p, b.bytes, cudaHostAllocDefault) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_live.push_back(b);
    *out = b;
    return true;
}

// returns false when p is not a live pinned block
bool pin_release(const void *p) {
    PinnedBlock victim;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        size_t i = 0;
        for (; i < g_pin_live.size(); i++) {
            if (g_pin_live[i].p == p) break;
        }
        if (i == g_pin_live.size()) return false;
        g_pin_free.push_back(g_pin_live[i]);
        g_pin_live.erase(g_pin_live.begin() + i);
        if (g_pin_free.size() <= kPinCacheBlocks) return true;
        size_t small = 0;  // keep the larger blocks
        for (size_t k = 1; k < g_pin_free.size(); k++) {
            if (g_pin_free[k].bytes < g_pin_free[small].bytes) small = k;
        }
        victim = g_pin_free[small];
        g_pin_free.erase(g_pin_free.begin() + small);
    }
    cudaFreeHost(victim.p);
    return true;
}

// One acgpu_match_utf16 call.
//   run_chunked (AhoCorasick family: every chunk is independent given max_len-1 chars of left context): the
//     haystack goes up chunk by chunk on an upload stream, the kernels of chunk k run on a compute stream as soon
//     as chunk k has landed, and the records of chunk k-1 come down on a third stream at the same time, so H2D,
//     scan and D2H overlap and the call is bound by the slower PCIe direction.
//   run_whole (Longest / Shortest / WholeWord: the selection chain runs over the whole haystack): upload, scan,
//     download.
struct HostCall {
    static constexpr int64_t kChunk = int64_t(1) << 23;  // chars per pipeline chunk (16 MiB of UTF-16)
    Matcher *m;
    bool is_map;
    CallCtx *cx = nullptr;
    cudaStream_t s_up = nullptr, s_k = nullptr, s_dn = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr};
    uint16_t *d_hay = nullptr;
    int2 *d_pos[2] = {nullptr, nullptr};
    uint32_t *d_val[2] = {nullptr, nullptr};
    char *d_ws[2] = {nullptr, nullptr};      // mask workspaces of the compact path
    unsigned long long *d_total = nullptr;   // [2]
    unsigned long long *h_total = nullptr;   // [2] pinned
    PinnedBlock blk;                          // result block: pos[cap] then val[cap]
    int64_t cap = 0, count = 0;               // records the block can hold / holds
    bool readable_view = false;               // literal WholeWord matchers: the semantics of the Readable overloads (RunOpts::folded_scroll)

    explicit HostCall(Matcher *mm) : m(mm), is_map(mm->host.is_map) {}

    int init() {
        {
            std::lock_guard<std::mutex> lk(m->ctx_mu);
            if (!m->ctx_free.empty()) {
                cx = m->ctx_free.back();
                m->ctx_free.pop_back();
            }
        }
        if (!cx) {
            cx = new (std::nothrow) CallCtx();
            if (!cx) return fail(ACGPU_ENOMEM, "out of memory");
            CU_TRY(cudaStreamCreateWithFlags(&cx->s_up, cudaStreamNonBlocking));
            CU_TRY(cudaStreamCreateWithFlags(&cx->s_k, cudaStreamNonBlocking));
            CU_TRY(cudaStreamCreateWithFlags(&cx->s_dn, cudaStreamNonBlocking));
            for (int i = 0; i < 2; i++) {
                CU_TRY(cudaEventCreateWithFlags(&cx->ev_up[i], cudaEventDisableTiming));
                CU_TRY(cudaEventCreateWithFlags(&cx->ev_k[i], cudaEventDisableTiming));
                CU_TRY(cudaEventCreateWithFlags(&cx->ev_dn[i], cudaEventDisableTiming));
            }
            CU_TRY(cudaMalloc(reinterpret_cast<void **>(&cx->d_total), 16));
            CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&cx->h_total), 16, cudaHostAllocDefault));
        }
        s_up = cx->s_up;
        s_k = cx->s_k;
        s_dn = cx->s_dn;
        for (int i = 0; i < 2; i++) {
            ev_up[i] = cx->ev_up[i];
            ev_k[i] = cx->ev_k[i];
            ev_dn[i] = cx->ev_dn[i];
        }
        d_total = cx->d_total;
        h_total = cx->h_total;
        return ACGPU_OK;
    }

    int32_t *h_pos() const { return static_cast<int32_t *>(blk.p); }
    uint32_t *h_val() const { return reinterpret_cast<uint32_t *>(static_cast<char *>(blk.p) + static_cast<size_t>(cap) * 8); }

    // make room for `need` records in the pinned block, keeping the `count` records already there
    int reserve(int64_t need) {
        if (need <= cap) return ACGPU_OK;
        const int64_t ncap = std::max<int64_t>(need, 1 << 16);
        PinnedBlock nb;
        if (!pin_take(static_cast<size_t>(ncap) * (is_map ? 12 : 8), &nb)) return fail(ACGPU_ENOMEM, "out of pinned host memory for the match records");
        const int64_t real_cap = static_cast<int64_t>(nb.bytes / (is_map ? 12 : 8));
        if (blk.p) {
            CU_TRY(cudaStreamSynchronize(s_dn));  // copies into the old block have landed
            std::memcpy(nb.p, blk.p, static_cast<size_t>(count) * 8);
            if (is_map) std::memcpy(static_cast<char *>(nb.p) + static_cast<size_t>(real_cap) * 8, h_val(), static_cast<size_t>(count) * 4);
            pin_release(blk.p);
        }
        blk = nb;
        cap = real_cap;
        return ACGPU_OK;
    }

    int download(const int2 *dp, const uint32_t *dv, int64_t k_records) {
        if (k_records == 0) return ACGPU_OK;
        CU_TRY(cudaMemcpyAsync(h_pos() + 2 * count, dp, static_cast<size_t>(k_records) * 8, cudaMemcpyDeviceToHost, s_dn));
        if (is_map) CU_TRY(cudaMemcpyAsync(h_val() + count, dv, static_cast<size_t>(k_records) * 4, cudaMemcpyDeviceToHost, s_dn));
        count += k_records;
        return ACGPU_OK;
    }

    int run_chunked(const uint16_t *hay, int64_t n) {
        const int64_t n_chunks = (n + kChunk - 1) / kChunk;
        const int64_t span0 = std::min<int64_t>(n, kChunk);
        const int64_t chunk_cap = std::max<int64_t>(1 << 16, span0 + span0 / 2);
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hay), static_cast<size_t>(n) * 2, s_k));
        for (int i = 0; i < (n_chunks > 1 ? 2 : 1); i++) {
            CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos[i]), static_cast<size_t>(chunk_cap) * 8, s_k));
            if (is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_val[i]), static_cast<size_t>(chunk_cap) * 4, s_k));
        }
        CU_TRY(cudaEventRecord(ev_dn[0], s_k));  // the upload stream may touch d_hay once it exists
        CU_TRY(cudaStreamWaitEvent(s_up, ev_dn[0], 0));
        auto issue = [&](int64_t k) -> int {
            const int bf = static_cast<int>(k & 1);
            const int64_t lo = k * kChunk, hi = std::min<int64_t>(n, lo + kChunk);
            CU_TRY(cudaMemcpyAsync(d_hay + lo, hay + lo, static_cast<size_t>(hi - lo) * 2, cudaMemcpyHostToDevice, s_up));
            CU_TRY(cudaEventRecord(ev_up[bf], s_up));
            CU_TRY(cudaStreamWaitEvent(s_k, ev_up[bf], 0));
            if (k >= 2) CU_TRY(cudaStreamWaitEvent(s_k, ev_dn[bf], 0));  // records of chunk k-2 have left the buffer
            RunOpts opt;
            // chars beyond `hi` are not on the device yet: the window ends at hi (matches end inside [lo, hi))
            int rc = enqueue_match(m, d_hay, hi, lo, hi, d_pos[bf], d_val[bf], chunk_cap, d_total + bf, s_k, opt);
            if (rc != ACGPU_OK) return rc;
            CU_TRY(cudaMemcpyAsync(h_total + bf, d_total + bf, 8, cudaMemcpyDeviceToHost, s_k));
            CU_TRY(cudaEventRecord(ev_k[bf], s_k));
            return ACGPU_OK;
        };
        int rc = issue(0);
        if (rc != ACGPU_OK) return rc;
        for (int64_t k = 0; k < n_chunks; k++) {
            const int bf = static_cast<int>(k & 1);
            if (k + 1 < n_chunks) {
                rc = issue(k + 1);
                if (rc != ACGPU_OK) return rc;
            }
            CU_TRY(cudaEventSynchronize(ev_k[bf]));
            const int64_t got = static_cast<int64_t>(h_total[bf]);
            if (count + got > cap) {
                // size the block from the density seen so far (plus head-room); grows again if the text gets denser
                const int64_t done_chars = std::min<int64_t>(n, (k + 1) * kChunk);
                const double density = static_cast<double>(count + got) / static_cast<double>(done_chars);
                const int64_t est = static_cast<int64_t>(density * 1.15 * static_cast<double>(n)) + (1 << 16);
                rc = reserve(std::max<int64_t>(count + got, k + 1 == n_chunks ? count + got : est));
                if (rc != ACGPU_OK) return rc;
            }
            if (got <= chunk_cap) {
                CU_TRY(cudaStreamWaitEvent(s_dn, ev_k[bf], 0));
                rc = download(d_pos[bf], d_val[bf], got);
                if (rc != ACGPU_OK) return rc;
            } else {
                // denser than the chunk buffers: scan this chunk again into an exact-size buffer (the first run
                // counted everything), after the pipeline has drained
                const int64_t lo = k * kChunk, hi = std::min<int64_t>(n, lo + kChunk);
                int2 *big_pos = nullptr;
                uint32_t *big_val = nullptr;
                CU_TRY(cudaStreamSynchronize(s_k));
                CU_TRY(cudaStreamSynchronize(s_dn));
                CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&big_pos), static_cast<size_t>(got) * 8, s_dn));
                if (is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&big_val), static_cast<size_t>(got) * 4, s_dn));
                RunOpts opt;
                rc = enqueue_match(m, d_hay, hi, lo, hi, big_pos, big_val, got, d_total + bf, s_dn, opt);
                if (rc == ACGPU_OK) rc = download(big_pos, big_val, got);
                cudaFreeAsync(big_pos, s_dn);
                if (big_val) cudaFreeAsync(big_val, s_dn);
                if (rc != ACGPU_OK) return rc;

            }
            CU_TRY(cudaEventRecord(ev_dn[bf], s_dn));
        }
        CU_TRY(cudaStreamSynchronize(s_dn));
        return ACGPU_OK;
    }

    // Dense AhoCorasickSet streams in the compact wire format (acgpu_match_utf16_compact): the 16-bit hit masks of
    // k_tier_mask ARE the ordered stream (position-major, ascending bit = longest first), 2 bytes per char whatever the
    // density, so the call moves 2 B/char up and 2 B/char down and k_tier_emit never runs.  Chunk 0 decides: when it
    // holds fewer than kMaskModeDensity records per char the caller falls back to records (*use_records = true).
    static constexpr double kMaskModeDensity = 0.25;  // 8-byte records cost as much as 2-byte masks at 0.25 records per char
    int64_t mask_total = 0;
    uint16_t *h_masks() const { return static_cast<uint16_t *>(blk.p); }

    int run_masks(const uint16_t *hay, int64_t n, bool *use_records) {
        *use_records = false;
        const int64_t n_chunks = (n + kChunk - 1) / kChunk;
        const int64_t mis_max = 8;
        const MaskWs LW = mask_ws_layout(std::min<int64_t>(n, kChunk) + mis_max, 0);  // workspace of the largest chunk
        char **ws = d_ws;  // freed by finish(), after every stream has drained
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hay), static_cast<size_t>(n) * 2, s_k));
        for (int i = 0; i < (n_chunks > 1 ? 2 : 1); i++) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&ws[i]), LW.bytes, s_k));
        CU_TRY(cudaEventRecord(ev_dn[0], s_k));
        CU_TRY(cudaStreamWaitEvent(s_up, ev_dn[0], 0));
        const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(d_hay) >> 1) & 7);
        unsigned long long *h_tot = nullptr;  // per-chunk totals, pinned, behind the masks in the result block
        auto origin_of = [&](int64_t lo) { return ((lo + mis) & ~int64_t(7)) - mis; };
        auto issue = [&](int64_t k) -> int {
            const int bf = static_cast<int>(k & 1);
            const int64_t lo = k * kChunk, hi = std::min<int64_t>(n, lo + kChunk);
            CU_TRY(cudaMemcpyAsync(d_hay + lo, hay + lo, static_cast<size_t>(hi - lo) * 2, cudaMemcpyHostToDevice, s_up));
            CU_TRY(cudaEventRecord(ev_up[bf], s_up));
            CU_TRY(cudaStreamWaitEvent(s_k, ev_up[bf], 0));
            if (k >= 2) CU_TRY(cudaStreamWaitEvent(s_k, ev_dn[bf], 0));  // the masks of chunk k-2 have left the workspace
            const int64_t origin = origin_of(lo);
            const MaskWs L = mask_ws_layout(hi, origin);
            int rc = enqueue_mask_scan(m, d_hay, hi, lo, hi, origin, ws[bf], L, d_total + bf, s_k);
            if (rc != ACGPU_OK) return rc;
            if (k == 0)
                CU_TRY(cudaMemcpyAsync(h_total, d_total, 8, cudaMemcpyDeviceToHost, s_k));
            else
                CU_TRY(cudaMemcpyAsync(h_tot + k, d_total + bf, 8, cudaMemcpyDeviceToHost, s_k));
            CU_TRY(cudaEventRecord(ev_k[bf], s_k));
            return ACGPU_OK;
        };
        auto download_masks = [&](int64_t k) -> int {
            const int bf = static_cast<int>(k & 1);
            const int64_t lo = k * kChunk, hi = std::min<int64_t>(n, lo + kChunk);
            const MaskWs L = mask_ws_layout(hi, origin_of(lo));
            CU_TRY(cudaStreamWaitEvent(s_dn, ev_k[bf], 0));
            CU_TRY(cudaMemcpyAsync(h_masks() + lo, ws[bf] + L.o_mask + static_cast<size_t>(lo - origin_of(lo)) * 2,
                                   static_cast<size_t>(hi - lo) * 2, cudaMemcpyDeviceToHost, s_dn));
            CU_TRY(cudaEventRecord(ev_dn[bf], s_dn));
            return ACGPU_OK;
        };
        int rc = issue(0);
        if (rc == ACGPU_OK) {
            const cudaError_t e = cudaEventSynchronize(ev_k[0]);
            if (e != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("match: ") + cudaGetErrorString(e));
        }
        if (rc == ACGPU_OK && static_cast<double>(h_total[0]) < kMaskModeDensity * static_cast<double>(std::min<int64_t>(n, kChunk))) {
            *use_records = true;
            return ACGPU_OK;
        }
        if (rc == ACGPU_OK) {
            const size_t mask_bytes = align_up(static_cast<size_t>(n) * 2, 64);
            if (!pin_take(mask_bytes + static_cast<size_t>(n_chunks) * 8, &blk)) rc = fail(ACGPU_ENOMEM, "out of pinned host memory for the hit masks");
            if (rc == ACGPU_OK) {
                h_tot = reinterpret_cast<unsigned long long *>(static_cast<char *>(blk.p) + mask_bytes);
                h_tot[0] = h_total[0];
            }
        }
        for (int64_t k = 0; rc == ACGPU_OK && k < n_chunks; k++) {
            if (k + 1 < n_chunks) rc = issue(k + 1);
            if (rc == ACGPU_OK) rc = download_masks(k);
        }
        if (rc == ACGPU_OK) {
            cudaError_t e = cudaStreamSynchronize(s_k);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s_dn);
            if (e != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("match: ") + cudaGetErrorString(e));
        }
        if (rc == ACGPU_OK) {
            mask_total = 0;
            for (int64_t k = 0; k < n_chunks; k++) mask_total += static_cast<int64_t>(h_tot[k]);
        }
        return rc;
    }

    // Longest / Shortest on the start-mask path, long haystacks: the haystack goes up chunk by chunk and every chunk is a CHAIN
    // SHARD (acgpu_chain_shard_*: the kernels of the multi-GPU shards and of the Readable feeds).  begin(k) - masks, maps, the
    // composed map - is entry-independent and runs as soon as chunk k and its look-ahead have landed; the map gives the record
    // count and the offset at which the chain enters chunk k + 1, so the records of chunk k are cut and come down while chunk
    // k + 2 goes up: H2D, kernels and D2H overlap as in run_chunked.  The last chunk takes what is left.
    int run_chain(const uint16_t *hay, int64_t n) {
        constexpr int64_t C = kChunk;   // a multiple of kS2Tile
        static_assert(kChunk % kS2Tile == 0, "inner shards own whole tiles");
        int64_t n_chunks = 1;
        while (n_chunks * C + kMaskRow < n) n_chunks++;   // every chunk but the last has kMaskRow chars of look-ahead behind it
        if (!cx->d_map) {
            CU_TRY(cudaMalloc(reinterpret_cast<void **>(&cx->d_map), 2 * kS2Ent * 8));
            CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&cx->h_map), 2 * kS2Ent * 8, cudaHostAllocDefault));
        }
        const int64_t last_lo = (n_chunks - 1) * C;
        const int64_t chunk_cap = std::max<int64_t>(C, n - last_lo) + 64;   // non-overlapping matches: at most one per char
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hay), static_cast<size_t>(n) * 2, s_k));
        for (int i = 0; i < (n_chunks > 1 ? 2 : 1); i++) {
            CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos[i]), static_cast<size_t>(chunk_cap) * 8, s_k));
            if (is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_val[i]), static_cast<size_t>(chunk_cap) * 4, s_k));
        }
        CU_TRY(cudaEventRecord(ev_dn[0], s_k));  // the upload stream may touch d_hay once it exists
        CU_TRY(cudaStreamWaitEvent(s_up, ev_dn[0], 0));
        Sel2Run R[2];
        int64_t up_to = 0;   // chars uploaded (enqueued) so far
        auto upload_to = [&](int64_t hi) -> int {
            hi = std::min<int64_t>(n, hi);
            while (up_to < hi) {
                const int64_t c = std::min<int64_t>(C, hi - up_to);
                CU_TRY(cudaMemcpyAsync(d_hay + up_to, hay + up_to, static_cast<size_t>(c) * 2, cudaMemcpyHostToDevice, s_up));
                up_to += c;
            }
            return ACGPU_OK;
        };
        auto begin = [&](int64_t k) -> int {
            const int bf = static_cast<int>(k & 1);
            const bool last = k + 1 == n_chunks;
            const int64_t lo = k * C, n_win = last ? n - lo : C + kMaskRow;
            int rc = upload_to(lo + n_win);
            if (rc != ACGPU_OK) return rc;
            CU_TRY(cudaEventRecord(ev_up[bf], s_up));
            CU_TRY(cudaStreamWaitEvent(s_k, ev_up[bf], 0));
            RunOpts opt;
            opt.pos_base = static_cast<int32_t>(lo);
            R[bf] = Sel2Run();
            rc = sel2_setup(m, d_hay + lo, n_win, last ? -1 : C / kS2Tile, s_k, opt, R[bf]);
            if (rc == ACGPU_OK) rc = sel2_masks(R[bf]);
            if (rc == ACGPU_OK) rc = sel2_maps(R[bf]);
            if (rc == ACGPU_OK && !last) {
                rc = sel2_shard_map(R[bf], cx->d_map + bf * kS2Ent, d_total + bf);
                if (rc == ACGPU_OK) CU_TRY(cudaMemcpyAsync(cx->h_map + bf * kS2Ent, cx->d_map + bf * kS2Ent, kS2Ent * 8, cudaMemcpyDeviceToHost, s_k));
            }
            if (rc == ACGPU_OK) CU_TRY(cudaEventRecord(ev_k[bf], s_k));
            return rc;
        };
        int rc = begin(0);
        int32_t entry = 0;
        for (int64_t k = 0; k < n_chunks && rc == ACGPU_OK; k++) {
            const int bf = static_cast<int>(k & 1);
            const bool last = k + 1 == n_chunks;
            if (!last) rc = begin(k + 1);
            if (rc != ACGPU_OK) break;
            if (k >= 2) CU_TRY(cudaStreamWaitEvent(s_k, ev_dn[bf], 0));   // the records of chunk k - 2 have left the buffer
            int64_t got = 0;
            uint32_t entry0 = static_cast<uint32_t>(entry);
            if (!last) {
                CU_TRY(cudaEventSynchronize(ev_k[bf]));   // the chunk's map is in pinned memory
                const unsigned long long row = cx->h_map[bf * kS2Ent + entry];
                got = static_cast<int64_t>(row >> 8);
                entry = static_cast<int32_t>(row & 0xFFu);
                if (got > 0) rc = sel2_records(R[bf], entry0, d_pos[bf], d_val[bf], chunk_cap, d_total + bf);
            } else {
                const int64_t idx = R[bf].moff + entry;   // index-space position the chain enters at
                if (idx < kS2Ent) {
                    entry0 = static_cast<uint32_t>(idx);
                } else {
                    entry0 = 0;
                    k_sel2_zero_prefix<<<1, 256, 0, s_k>>>(R[bf].P.masks, R[bf].moff, idx);
                    CU_TRY(cudaGetLastError());
                    rc = sel2_maps(R[bf], 1);
                }
                if (rc == ACGPU_OK) rc = sel2_records(R[bf], entry0, d_pos[bf], d_val[bf], chunk_cap, d_total + bf);
                if (rc == ACGPU_OK) {
                    CU_TRY(cudaMemcpyAsync(h_total + bf, d_total + bf, 8, cudaMemcpyDeviceToHost, s_k));
                    CU_TRY(cudaStreamSynchronize(s_k));
                    got = static_cast<int64_t>(h_total[bf]);
                }
            }
            sel2_free(R[bf]);
            if (rc != ACGPU_OK) break;
            CU_TRY(cudaEventRecord(ev_k[bf], s_k));
            if (count + got > cap) {
                const int64_t done_chars = std::min<int64_t>(n, (k + 1) * C);
                const double density = static_cast<double>(count + got) / static_cast<double>(done_chars);
                const int64_t est = static_cast<int64_t>(density * 1.15 * static_cast<double>(n)) + (1 << 16);
                rc = reserve(std::max<int64_t>(count + got, last ? count + got : est));
                if (rc != ACGPU_OK) break;
            }
            CU_TRY(cudaStreamWaitEvent(s_dn, ev_k[bf], 0));
            rc = download(d_pos[bf], d_val[bf], got);
            if (rc != ACGPU_OK) break;
            CU_TRY(cudaEventRecord(ev_dn[bf], s_dn));
        }
        if (rc != ACGPU_OK) {
            sel2_free(R[0]);
            sel2_free(R[1]);
            return rc;
        }
        CU_TRY(cudaStreamSynchronize(s_dn));
        return ACGPU_OK;
    }

    int run_whole(const uint16_t *hay, int64_t n) {
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hay), static_cast<size_t>(n) * 2, s_k));
        CU_TRY(cudaMemcpyAsync(d_hay, hay, static_cast<size_t>(n) * 2, cudaMemcpyHostToDevice, s_k));
        int64_t dcap = std::max<int64_t>(1 << 16, n / 4);
        for (int attempt = 0; attempt < 2; attempt++) {
            CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos[0]), static_cast<size_t>(dcap) * 8, s_k));
            if (is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_val[0]), static_cast<size_t>(dcap) * 4, s_k));
            RunOpts opt;
            opt.folded_scroll = readable_view;
            int rc = enqueue_match(m, d_hay, n, 0, n, d_pos[0], d_val[0], dcap, d_total, s_k, opt);
            if (rc != ACGPU_OK) return rc;
            CU_TRY(cudaMemcpyAsync(h_total, d_total, 8, cudaMemcpyDeviceToHost, s_k));
            CU_TRY(cudaStreamSynchronize(s_k));
            if (static_cast<int64_t>(h_total[0]) <= dcap) break;
            // first guess too small: the kernels still counted everything; rerun with the exact size
            cudaFreeAsync(d_pos[0], s_k);
            d_pos[0] = nullptr;
            if (d_val[0]) cudaFreeAsync(d_val[0], s_k);
            d_val[0] = nullptr;
            dcap = static_cast<int64_t>(h_total[0]);
        }
        const int64_t got = static_cast<int64_t>(h_total[0]);
        int rc = reserve(got);
        if (rc != ACGPU_OK) return rc;
        rc = download(d_pos[0], d_val[0], got);  // on s_dn; the scan on s_k has been synchronised above
        if (rc != ACGPU_OK) return rc;
        CU_TRY(cudaStreamSynchronize(s_dn));
        return ACGPU_OK;
    }

    // release everything; on success hand the pinned block to the caller
    int finish_masks(int rc, acgpu_matches *out, int64_t n_chars) {
        acgpu_result none;
        fill_empty(&none);
        PinnedBlock keep = blk;
        blk = PinnedBlock();
        count = 0;
        rc = finish(rc, &none);
        if (rc == ACGPU_OK) {
            out->kind = ACGPU_MATCHES_MASKS;
            out->n = mask_total;
            out->masks = static_cast<const uint16_t *>(keep.p);
            out->n_chars = n_chars;
        } else if (keep.p) {
            pin_release(keep.p);
        }
        return rc;
    }

    int finish(int rc, acgpu_result *out) {
        cudaError_t pending = cudaSuccess;
        for (cudaStream_t st : {s_up, s_dn, s_k}) {
            if (st) {
                cudaError_t e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) pending = e;
            }
        }
        if (rc == ACGPU_OK && pending != cudaSuccess) rc = fail(ACGPU_ECUDA, std::string("match: ") + cudaGetErrorString(pending));
        if (s_k) {
            if (d_hay) cudaFreeAsync(d_hay, s_k);
            for (int i = 0; i < 2; i++) {
                if (d_pos[i]) cudaFreeAsync(d_pos[i], s_k);
                if (d_val[i]) cudaFreeAsync(d_val[i], s_k);
                if (d_ws[i]) cudaFreeAsync(d_ws[i], s_k);
            }
        }
        if (cx) {
            if (pending == cudaSuccess && cx->h_total) {
                std::lock_guard<std::mutex> lk(m->ctx_mu);
                m->ctx_free.push_back(cx);
            } else {
                cx->destroy();
                delete cx;
            }
            cx = nullptr;
        }
        if (rc == ACGPU_OK && count > 0) {
            out->n = count;
            out->pos = h_pos();
            out->val = is_map ? h_val() : nullptr;
        } else if (blk.p) {
            pin_release(blk.p);
        }
        return rc;
    }
};

}  // namespace

extern "C" {

const char *acgpu_last_error(void) { return g_err.c_str(); }
const char *acgpu_version(void) { return "acgpu 0.2 (sm_100a; tiered hit-mask kernels, generation 4: pair rows)"; }

int acgpu_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n, uint8_t *out65536) {
    if (!out65536 || mode < 0 || mode > 2 || n < 0) return fail(ACGPU_EINVAL, "bad arguments");
    make_word_chars(mode, chars, toggles, n, out65536);
    return ACGPU_OK;
}

namespace {

// device half of a constructor: m->host is flattened, bring the tables to `device`
int finish_create(Matcher *m, int family, int device, uint64_t *handle) {
    // keywords the selection kernels cannot hold (one exit entry per possible keyword length in shared memory; 16-bit walk
    // records for WholeWordLongest) and quirk Q7: the literal loops of kernel_wwlit.cuh
    if (m->host.ww_literal)
        m->literal_family = family == ACGPU_WHOLEWORD ? 0 : 4;
    else if ((family == ACGPU_LONGEST || family == ACGPU_SHORTEST) && m->host.max_len + 1 > kSelMaxLen)
        m->literal_family = family == ACGPU_LONGEST ? 1 : 2;
    else if (family == ACGPU_WHOLEWORDLONGEST && m->host.max_len > 254)
        m->literal_family = 4;
    if (m->host.max_len > 65535) {
        delete m;
        return fail(ACGPU_EUNSUPPORTED, "keywords longer than 65535 chars are not supported");
    }
    m->device = device;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        delete m;
        return fail(ACGPU_ENODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (device < 0 || device >= n_dev) {
        delete m;
        return fail(ACGPU_EINVAL, "device ordinal out of range");
    }
    int rc = ensure_device(m);
    if (rc == ACGPU_OK) {
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;
        rc = upload(m);
        if (rc == ACGPU_OK) rc = upload_tier(m);
        if (rc == ACGPU_OK) rc = upload_wide(m);
        if (rc == ACGPU_OK) rc = upload_ww(m);
    }
    if (rc != ACGPU_OK) {
        if (m->tier.kid_tex) cudaDestroyTextureObject(m->tier.kid_tex);
        if (m->d_tier_blob) cudaFree(m->d_tier_blob);
        if (m->d_ww_blob) cudaFree(m->d_ww_blob);
        if (m->d_wide_blob) cudaFree(m->d_wide_blob);
        if (m->d_wide_vals) cudaFree(m->d_wide_vals);
        if (m->d_wwl_wcls) cudaFree(m->d_wwl_wcls);
        if (m->d_blob) cudaFree(m->d_blob);
        delete m;
        return rc;
    }
    *handle = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(m));
    return ACGPU_OK;
}

// flatten with the C++ builder, mapping its exceptions onto the ABI's codes; 0 and `out` filled on success
int flatten(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null, int64_t n_keywords, int64_t n_values,
            bool case_sensitive, const uint8_t *word_chars, HostAutomaton &out) {
    try {
        out = build_automaton(family, chars, offsets, is_null, n_keywords, n_values, case_sensitive, word_chars);
    } catch (const IllegalArgument &e) {
        return fail(ACGPU_EILLEGALARG, e.what());
    } catch (const std::domain_error &e) {
        return fail(ACGPU_EUNSUPPORTED, e.what());
    } catch (const std::bad_alloc &) {
        return fail(ACGPU_ENOMEM, "out of memory while flattening the dictionary");
    } catch (const std::exception &e) {
        return fail(ACGPU_EINVAL, e.what());
    }
    return ACGPU_OK;
}

// The dictionary a trie descriptor spells, in the arrays acgpu_create_from_keywords takes.  Sets: one keyword per terminal
// state, in state order.  Maps: entry v = the keyword whose state carries value index v, every other entry null - so the
// builder's "value = index of the entry" reproduces the descriptor's indices.
struct DescKeywords {
    std::vector<uint16_t> chars;
    std::vector<int64_t> offsets;
    std::vector<uint8_t> is_null;
    int64_t n_keywords = 0, n_values = -1;
};

int keywords_of_desc(const acgpu_automaton_desc *d, DescKeywords &K) {
    if (!d || d->struct_size < static_cast<int32_t>(sizeof(acgpu_automaton_desc))) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: bad struct_size");
    if (d->family < 0 || d->family > 4 || d->n_states < 1 || !d->parent || !d->edge_char || !d->terminal)
        return fail(ACGPU_EINVAL, "acgpu_automaton_desc: family / n_states / parent / edge_char / terminal");
    if (d->is_map && (!d->value || d->n_values < 0)) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: a Map needs value[] and n_values");
    const int64_t n = d->n_states;
    if (d->parent[0] != -1) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: state 0 must be the root (parent -1)");
    std::vector<int32_t> depth(static_cast<size_t>(n), 0);
    for (int64_t s = 1; s < n; s++) {
        const int32_t p = d->parent[s];
        if (p < 0 || p >= s) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: parent[s] must be in [0, s)");
        depth[s] = depth[p] + 1;
    }
    std::vector<int64_t> state_of;  // entry -> terminal state (-1: null entry)
    if (d->is_map) {
        state_of.assign(static_cast<size_t>(d->n_values), -1);
        for (int64_t s = 1; s < n; s++) {
            if (!d->terminal[s]) continue;
            const uint32_t v = d->value[s];
            if (static_cast<int64_t>(v) >= d->n_values) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: value index out of range");
            if (state_of[v] >= 0) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: a value index is carried by two states");
            state_of[v] = s;
        }
        K.n_values = d->n_values;
    } else {
        for (int64_t s = 1; s < n; s++)
            if (d->terminal[s]) state_of.push_back(s);
    }
    K.n_keywords = static_cast<int64_t>(state_of.size());
    K.offsets.assign(state_of.size() + 1, 0);
    K.is_null.assign(state_of.size(), 0);
    int64_t total = 0;
    for (size_t k = 0; k < state_of.size(); k++) {
        K.offsets[k] = total;
        if (state_of[k] < 0) K.is_null[k] = 1; else total += depth[state_of[k]];
    }
    K.offsets[state_of.size()] = total;
    K.chars.assign(static_cast<size_t>(std::max<int64_t>(total, 1)), 0);
    for (size_t k = 0; k < state_of.size(); k++) {
        int64_t s = state_of[k];
        if (s < 0) continue;
        int64_t at = K.offsets[k] + depth[s];
        for (; s > 0; s = d->parent[s]) K.chars[static_cast<size_t>(--at)] = d->edge_char[s];
    }
    if (d->fail) {
        // the links the trie implies (BFS order = state order is not required: process by depth)
        std::unordered_map<uint64_t, int32_t> child;
        child.reserve(static_cast<size_t>(n) * 2);
        for (int64_t s = 1; s < n; s++) child[(static_cast<uint64_t>(d->parent[s]) << 16) | d->edge_char[s]] = static_cast<int32_t>(s);
        std::vector<int32_t> order(static_cast<size_t>(n));
        for (int64_t s = 0; s < n; s++) order[s] = static_cast<int32_t>(s);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return depth[x] < depth[y]; });
        std::vector<int32_t> f(static_cast<size_t>(n), 0);
        for (int32_t s : order) {
            if (depth[s] <= 1) continue;
            int32_t t = f[d->parent[s]];
            while (true) {
                auto it = child.find((static_cast<uint64_t>(t) << 16) | d->edge_char[s]);
                if (it != child.end()) { f[s] = it->second; break; }
                if (t == 0) break;
                t = f[t];
            }
        }
        for (int64_t s = 0; s < n; s++)
            if (d->fail[s] != f[s]) return fail(ACGPU_EINVAL, "acgpu_automaton_desc: fail[] does not match the failure links of the trie");
    }
    return ACGPU_OK;
}

}  // namespace

int acgpu_create_from_keywords(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                               int64_t n_keywords, int64_t n_values, int case_sensitive, const uint8_t *word_chars,
                               int device, uint64_t *handle) {
    if (!handle || n_keywords < 0 || (n_keywords > 0 && (!chars || !offsets))) return fail(ACGPU_EINVAL, "bad arguments");
    *handle = 0;
    Matcher *m = new (std::nothrow) Matcher();
    if (!m) return fail(ACGPU_ENOMEM, "out of memory");
    const int rc = flatten(family, chars, offsets, is_null, n_keywords, n_values, case_sensitive != 0, word_chars, m->host);
    if (rc != ACGPU_OK) {
        delete m;
        return rc;
    }
    return finish_create(m, family, device, handle);
}

int acgpu_create(const acgpu_automaton_desc *desc, uint64_t *handle) {
    if (!handle) return fail(ACGPU_EINVAL, "bad arguments");
    *handle = 0;
    DescKeywords K;
    int rc = keywords_of_desc(desc, K);
    if (rc != ACGPU_OK) return rc;
    Matcher *m = new (std::nothrow) Matcher();
    if (!m) return fail(ACGPU_ENOMEM, "out of memory");
    rc = flatten(desc->family, K.chars.data(), K.offsets.data(), K.is_null.data(), K.n_keywords, K.n_values, desc->case_sensitive != 0,
                 desc->word_chars, m->host);
    if (rc != ACGPU_OK) {
        delete m;
        return rc;
    }
    return finish_create(m, desc->family, desc->device, handle);
}

int acgpu_desc_fingerprint(const acgpu_automaton_desc *desc, uint64_t *fingerprint) {
    if (!fingerprint) return fail(ACGPU_EINVAL, "bad arguments");
    DescKeywords K;
    int rc = keywords_of_desc(desc, K);
    if (rc != ACGPU_OK) return rc;
    HostAutomaton a;
    rc = flatten(desc->family, K.chars.data(), K.offsets.data(), K.is_null.data(), K.n_keywords, K.n_values, desc->case_sensitive != 0,
                 desc->word_chars, a);
    if (rc == ACGPU_OK) *fingerprint = automaton_fingerprint(a);
    return rc;
}

int acgpu_build_fingerprint(int family, const uint16_t *chars, const int64_t *offsets, const uint8_t *is_null,
                            int64_t n_keywords, int64_t n_values, int case_sensitive, const uint8_t *word_chars,
                            uint64_t *fingerprint, double *build_seconds) {
    if (!fingerprint || n_keywords < 0 || (n_keywords > 0 && (!chars || !offsets))) return fail(ACGPU_EINVAL, "bad arguments");
    try {
        const auto t0 = std::chrono::steady_clock::now();
        const HostAutomaton a = build_automaton(family, chars, offsets, is_null, n_keywords, n_values, case_sensitive != 0, word_chars);
        if (build_seconds) *build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *fingerprint = automaton_fingerprint(a);
    } catch (const IllegalArgument &e) {
        return fail(ACGPU_EILLEGALARG, e.what());
    } catch (const std::domain_error &e) {
        return fail(ACGPU_EUNSUPPORTED, e.what());
    } catch (const std::bad_alloc &) {
        return fail(ACGPU_ENOMEM, "out of memory while flattening the dictionary");
    } catch (const std::exception &e) {
        return fail(ACGPU_EINVAL, e.what());
    }
    return ACGPU_OK;
}

int acgpu_destroy(uint64_t handle) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    cudaSetDevice(m->device);
    for (CallCtx *cx : m->ctx_free) {
        cx->destroy();
        delete cx;
    }
    m->ctx_free.clear();
    if (m->d_blob) cudaFree(m->d_blob);
    if (m->tier.kid_tex) cudaDestroyTextureObject(m->tier.kid_tex);
    if (m->d_tier_blob) cudaFree(m->d_tier_blob);
    if (m->d_ww_blob) cudaFree(m->d_ww_blob);
    if (m->d_wide_blob) cudaFree(m->d_wide_blob);
    if (m->d_wide_vals) cudaFree(m->d_wide_vals);
    if (m->d_wwl_wcls) cudaFree(m->d_wwl_wcls);
    m->magic = 0;
    delete m;
    return ACGPU_OK;
}

int acgpu_info(uint64_t handle, int64_t *n_nodes, int32_t *n_classes, int32_t *max_len, int32_t *char_buffer_size,
               int64_t *table_bytes) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (n_nodes) *n_nodes = m->host.n_nodes;
    if (n_classes) *n_classes = m->host.n_classes;
    if (max_len) *max_len = m->host.max_len;
    if (char_buffer_size) *char_buffer_size = m->host.char_buffer_size;
    if (table_bytes) *table_bytes = m->table_bytes;
    return ACGPU_OK;
}

int acgpu_char_classes(uint64_t handle, uint16_t *out65536, int32_t *has_other) {
    Matcher *m = as_matcher(handle);
    if (!m || !out65536) return fail(ACGPU_EINVAL, "bad arguments");
    std::memcpy(out65536, m->host.cls.data(), 65536 * sizeof(uint16_t));
    if (has_other) *has_other = m->host.has_other ? 1 : 0;
    return ACGPU_OK;
}

int acgpu_launches_per_match(uint64_t handle) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (m->literal_family >= 0) return 3;  // count, scan, write
    switch (m->host.family) {
    case ACGPU_AHOCORASICK: return m->use_tier ? 3 : (m->use_wide ? (m->wide_tile ? 4 : 3) : 1);  // wide, generation 2: tile, tail, scan, emit
    case ACGPU_WHOLEWORD: return m->use_ww ? (m->use_ww3 ? 4 : 1) : 2;  // generation 3: hits, count, scan, emit
    default: return m->use_tier && m->host.is_map ? 7 : 6;  // one-shot tier path: mask, map, group, top, tiles, emit (+ values)  // one-shot matches; the streaming path always takes the 6-launch route
    }
}

int acgpu_match_device_async(uint64_t handle, const void *d_haystack, int64_t n, int64_t emit_from, int64_t emit_to,
                             void *d_pos, void *d_val, int64_t cap, void *d_total, void *cuda_stream) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (!d_total || (n > 0 && !d_haystack) || (cap > 0 && !d_pos) || cap < 0) return fail(ACGPU_EINVAL, "bad arguments");
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    RunOpts opt;
    if (m->host.family == ACGPU_WHOLEWORD || (m->host.family == ACGPU_WHOLEWORDLONGEST && m->use_ww)) {
        // word-start range shards: a word is matched from its own start, so a shard reports the words that
        // START in [emit_from, emit_to); it reads one char of look-behind and up to max_len + 1 chars past emit_to
        if (emit_from < 0 || emit_to > n || emit_from > emit_to) return fail(ACGPU_EINVAL, "emit range outside the haystack");
        opt.ctx = emit_from;
        opt.chain_n = emit_to;
    }
    return enqueue_match(m, static_cast<const uint16_t *>(d_haystack), n, emit_from, emit_to, static_cast<int2 *>(d_pos),
                         static_cast<uint32_t *>(d_val), cap, static_cast<unsigned long long *>(d_total),
                         static_cast<cudaStream_t>(cuda_stream), opt);
}

int acgpu_match_device(uint64_t handle, const void *d_haystack, int64_t n, int64_t emit_from, int64_t emit_to,
                       void *d_pos, void *d_val, int64_t cap, int64_t *n_out, void *cuda_stream) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (!n_out) return fail(ACGPU_EINVAL, "bad arguments");
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    unsigned long long *d_total = nullptr;
    CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_total), 8, st));
    rc = acgpu_match_device_async(handle, d_haystack, n, emit_from, emit_to, d_pos, d_val, cap, d_total, cuda_stream);
    unsigned long long total = 0;
    if (rc == ACGPU_OK) {
        CU_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    cudaFreeAsync(d_total, st);
    if (rc == ACGPU_OK) *n_out = static_cast<int64_t>(total);
    return rc;
}

// ---- Longest / Shortest range shards of ONE haystack (LongestMatchSet.java:192-265, SetMatchQueue.java:45-95,
//      ShortestMatchSet.java:182-260).  The selection chain pos -> J(pos) crosses shard boundaries; a shard's effect on it
//      is a map "entry offset -> (exit offset, matches)" over at most 16 entry offsets, so the shards scan in parallel,
//      exchange their maps (16 words each, they ride in the count all-gather) and then emit from their true entry.
namespace {
constexpr uint64_t kChainMagic = 0xC4A1135AAD0B200ull;
struct ChainShard {
    uint64_t magic = kChainMagic;
    Sel2Run R;
    int64_t n_domain = 0;
    unsigned long long *d_tmp = nullptr;  // scratch total of the map pass
};
}  // namespace

int acgpu_chain_shard_begin(uint64_t handle, const void *d_window, int64_t n, int64_t n_domain, void *d_map16, uint64_t *shard,
                            void *cuda_stream) {
    Matcher *m = as_matcher(handle);
    if (!m || !shard) return fail(ACGPU_EINVAL, "bad arguments");
    *shard = 0;
    if (m->host.family != ACGPU_LONGEST && m->host.family != ACGPU_SHORTEST)
        return fail(ACGPU_EINVAL, "chain shards are for the Longest / Shortest families (the others shard by range or word start)");
    if (n <= 0 || n > 0x7FFFFFFFll || !d_window || n_domain <= 0 || n_domain > n) return fail(ACGPU_EINVAL, "bad window");
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    if (!m->use_tier)
        return fail(ACGPU_EUNSUPPORTED, "chain shards need the start-mask path (at most 31 keyword symbols, keywords of at most 16 chars)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    ChainShard *c = new (std::nothrow) ChainShard();
    if (!c) return fail(ACGPU_ENOMEM, "out of memory");
    RunOpts opt;
    // a shard that ends before its window does owns whole tiles: its last chain position must be the last index of a tile
    const uint16_t *hay = static_cast<const uint16_t *>(d_window);
    {
        const int64_t mis = static_cast<int64_t>((reinterpret_cast<uintptr_t>(hay) >> 1) & 7);
        const int64_t r = (mis + n) & 7;
        const int64_t origin = r ? r - 8 : 0;
        const int64_t n_idx = (n - origin + kMaskRow - 1) / kMaskRow * kMaskRow;
        const int64_t moff = n_idx - n + origin;
        if (n_domain < n && ((n_domain + moff) % kS2Tile != 0 || n - n_domain < 2 * m->host.max_len + 2)) {
            delete c;
            return fail(ACGPU_EINVAL, "a chain shard inside the haystack must end on a tile boundary (acgpu_chain_shard_layout) and "
                                      "needs 2 * max_len + 2 chars of look-ahead");
        }
        c->n_domain = n_domain;
        rc = sel2_setup(m, hay, n, n_domain < n ? (n_domain + moff) / kS2Tile : -1, st, opt, c->R);
    }
    if (rc == ACGPU_OK) rc = sel2_masks(c->R);
    if (rc == ACGPU_OK) rc = sel2_maps(c->R);
    if (rc == ACGPU_OK && d_map16) {
        if (cudaMallocAsync(reinterpret_cast<void **>(&c->d_tmp), 8, st) != cudaSuccess) rc = fail(ACGPU_ECUDA, "cudaMallocAsync failed");
        if (rc == ACGPU_OK) rc = sel2_shard_map(c->R, static_cast<unsigned long long *>(d_map16), c->d_tmp);
        // a window that does not start on a tile boundary of its own index space (only the LAST shard may) has no rows
        // for the entry offsets > 0: they would index its first tile at moff + entry >= 16
        if (rc == ACGPU_OK && c->R.moff != 0 &&
            cudaMemsetAsync(static_cast<unsigned long long *>(d_map16) + 1, 0xFF, (kS2Ent - 1) * 8, st) != cudaSuccess)
            rc = fail(ACGPU_ECUDA, "cudaMemsetAsync failed");
    }
    if (rc != ACGPU_OK) {
        if (c->d_tmp) cudaFreeAsync(c->d_tmp, st);
        sel2_free(c->R);
        delete c;
        return rc;
    }
    *shard = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(c));
    return ACGPU_OK;
}

int acgpu_chain_shard_finish(uint64_t shard, int32_t entry, int32_t pos_base, void *d_pos, void *d_val, int64_t cap, void *d_total,
                             void *cuda_stream) {
    ChainShard *c = reinterpret_cast<ChainShard *>(static_cast<uintptr_t>(shard));
    if (!c || c->magic != kChainMagic) return fail(ACGPU_EINVAL, "bad shard handle");
    Sel2Run &R = c->R;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = ACGPU_OK;
    if (st != R.st) rc = fail(ACGPU_EINVAL, "finish must use the stream of begin");
    if (rc == ACGPU_OK && (entry < 0 || entry >= kS2Ent || !d_total || cap < 0 || (cap > 0 && !d_pos) || (cap > 0 && R.m->dev.is_map && !d_val)))
        rc = fail(ACGPU_EINVAL, "bad arguments");
    uint32_t entry0 = 0;
    if (rc == ACGPU_OK && entry > 0) {
        const int64_t idx = R.moff + entry;  // index-space position the chain enters at
        if (idx < kS2Ent) {
            entry0 = static_cast<uint32_t>(idx);
        } else {
            // the window is not tile-aligned at its start: hide the starts left of the entry and resolve tile 0 again
            k_sel2_zero_prefix<<<1, 256, 0, st>>>(R.P.masks, R.moff, idx);
            if (cudaGetLastError() != cudaSuccess) rc = fail(ACGPU_ECUDA, "k_sel2_zero_prefix failed");
            if (rc == ACGPU_OK) rc = sel2_maps(R, 1);
        }
    }
    if (rc == ACGPU_OK) {
        R.Q.pos_base = pos_base;
        rc = sel2_records(R, entry0, static_cast<int2 *>(d_pos), static_cast<uint32_t *>(d_val), cap, static_cast<unsigned long long *>(d_total));
    }
    if (c->d_tmp) cudaFreeAsync(c->d_tmp, R.st);
    sel2_free(R);
    c->magic = 0;
    delete c;
    return rc;
}

int acgpu_chain_shard_layout(uint64_t handle, int64_t *tile_chars, int64_t *lookahead_chars, int32_t *map_entries) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (tile_chars) *tile_chars = kS2Tile;
    if (lookahead_chars) *lookahead_chars = kMaskRow;  // a multiple of 256 keeps every shard's window aligned; >= 2 * max_len + 2
    if (map_entries) *map_entries = kS2Ent;
    return ACGPU_OK;
}

int acgpu_match_utf16(uint64_t handle, const uint16_t *haystack, int32_t n, acgpu_result *out) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (!out || n < 0 || (n > 0 && !haystack)) return fail(ACGPU_EINVAL, "bad arguments");
    fill_empty(out);
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    if (n == 0) return ACGPU_OK;
    HostCall hc(m);
    rc = hc.init();
    {
        static const char *chain_env = getenv("ACGPU_HOST_CHAIN");   // ACGPU_HOST_CHAIN=0: upload, scan, download in turn (A/B runs)
        const bool chain = (m->host.family == ACGPU_LONGEST || m->host.family == ACGPU_SHORTEST) && m->use_tier && m->literal_family < 0 &&
                           n >= 2 * HostCall::kChunk && !(chain_env && chain_env[0] == '0');
        if (rc == ACGPU_OK)
            rc = m->host.family == ACGPU_AHOCORASICK ? hc.run_chunked(haystack, n) : (chain ? hc.run_chain(haystack, n) : hc.run_whole(haystack, n));
    }
    return hc.finish(rc, out);
}

int acgpu_match_utf16_compact(uint64_t handle, const uint16_t *haystack, int32_t n, acgpu_matches *out) {
    Matcher *m = as_matcher(handle);
    if (!m) return fail(ACGPU_EINVAL, "bad handle");
    if (!out || n < 0 || (n > 0 && !haystack)) return fail(ACGPU_EINVAL, "bad arguments");
    std::memset(out, 0, sizeof(*out));
    out->kind = ACGPU_MATCHES_RECORDS;
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    if (n == 0) return ACGPU_OK;
    const char *off = getenv("ACGPU_NO_MASK_RESULTS");
    if (m->host.family == ACGPU_AHOCORASICK && m->use_tier && !m->host.is_map && !(off && off[0] == '1')) {
        HostCall hc(m);
        bool use_records = false;
        rc = hc.init();
        if (rc == ACGPU_OK) rc = hc.run_masks(haystack, n, &use_records);
        if (rc != ACGPU_OK || !use_records) return hc.finish_masks(rc, out, n);
        acgpu_result none;
        fill_empty(&none);
        hc.finish(ACGPU_OK, &none);
    }
    acgpu_result r;
    rc = acgpu_match_utf16(handle, haystack, n, &r);
    if (rc != ACGPU_OK) return rc;
    out->n = r.n;
    out->pos = r.pos;
    out->val = r.val;
    return ACGPU_OK;
}

void acgpu_free_matches(acgpu_matches *r) {
    if (!r) return;
    if (r->kind == ACGPU_MATCHES_MASKS) {
        if (r->masks) pin_release(r->masks);
    } else {
        acgpu_result q;
        q.n = r->n;
        q.pos = r->pos;
        q.val = r->val;
        acgpu_free_result(&q);
    }
    std::memset(r, 0, sizeof(*r));
}

int64_t acgpu_masks_to_records(const uint16_t *masks, int64_t n_chars, int64_t first_char, int32_t *pos_out, int64_t cap) {
    // host-side expansion of the compact stream, the loop every mirror's lazy replay runs: position ascending, then
    // bit index ascending = longest keyword first (AhoCorasickSet.java:522-535)
    if (!masks || n_chars < 0 || first_char < 0 || cap < 0 || (cap > 0 && !pos_out)) return -1;
    int64_t k = 0;
    for (int64_t q = first_char; q < n_chars; q++) {
        uint32_t mk = masks[q];
        while (mk) {
            const int t = __builtin_ctz(mk);
            mk &= mk - 1u;
            if (k < cap) {
                pos_out[2 * k] = static_cast<int32_t>(q + 1 - (16 - t));
                pos_out[2 * k + 1] = static_cast<int32_t>(q + 1);
            }
            k++;
        }
    }
    return k;
}

void acgpu_free_result(acgpu_result *r) {
    if (!r) return;
    if (r->pos) {
        if (!pin_release(r->pos)) {  // not from the pinned cache: plain heap blocks
            free(const_cast<int32_t *>(r->pos));
            free(const_cast<uint32_t *>(r->val));
        }
    } else if (r->val) {
        pin_release(r->val);  // values-only stream results: the block starts at the values
    }
    fill_empty(r);
}

// ---------------------------------------------------------------------------------------------------
// match(Readable, listener): block-wise scan with carried context.
//
// The device window holds [left context | new chars].  AhoCorasick (end-anchored) finalises every new
// position at once and keeps the last max_len-1 chars as context.  The start-anchored families finalise the
// chain domain [ctx, avail - D) where D = 2*max_len + 2 chars of look-ahead make v[] and the Shortest halo
// exact; the unfinalised tail (plus one char of left context) is carried to the next window together with
// the absolute chain position.  Host chars travel through two pinned staging buffers with cudaMemcpyAsync so
// the CPU copy of chunk k+1 overlaps the DMA of chunk k.
namespace {

constexpr uint64_t kStreamMagic = 0x5712EA30AC69B200ull;
constexpr size_t kPinChars = 1u << 21;  // 4 MiB per pinned staging buffer

struct StreamCtx {
    uint64_t magic = kStreamMagic;
    Matcher *m = nullptr;
    cudaStream_t st = nullptr;
    uint16_t *h_pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    uint16_t *d_win[2] = {nullptr, nullptr};
    int64_t win_cap[2] = {0, 0};
    int cur = 0;          // active window
    int64_t len = 0;      // chars in the active window (context + unprocessed tail)
    int64_t ctx = 0;      // leading context chars of the active window
    int64_t base = 0;     // absolute stream offset of window[0]
    int64_t chain = 0;    // absolute chain position (selection families)
    // Longest / Shortest on the start-mask path (stream_process_chain): the window starts at the chain domain, the chain enters
    // it at window position `entry` (< 16: an exit offset of the previous block's map)
    bool chain_blocks = false;
    int32_t entry = 0;
    unsigned long long *d_map = nullptr;   // [16] composed map of the block (exit offset | matches << 8)
    // pipelined chain-shard feeds: the block whose masks and maps are enqueued (entry-independent work) while the caller prepares
    // the next feed; its records are cut when the next feed (or end) arrives
    bool chain_pipe = false;
    struct ChainPending {
        bool live = false;
        Sel2Run R;
        int64_t n_domain = 0;
    } cpend;
    unsigned long long *h_map = nullptr;   // pinned [16]: the pending block's map
    cudaEvent_t ev_begin = nullptr;
    int64_t *d_carry = nullptr;
    unsigned long long *d_total = nullptr;
    double density = 0.0;  // records per finalised char seen so far (max over feeds): sizes the next feed's record buffer
    bool dma_pending = false;  // a DMA straight from the caller's page-locked buffer is in flight (acgpu.h: host buffers are only read during the call)
    // ---- pipelined feeds (AhoCorasick / WholeWord): the block of feed k is uploaded on st_up while the records of block
    //      k-1 come down on st, and feed k returns the records of block k-1 ("the records that are final so far" - one
    //      block later); end() returns the rest.  PCIe runs in both directions at once and no feed waits for its own scan.
    bool pipelined = false;
    bool values_only = false;  // acgpu_stream_set_values_only: ReadableMatchListener sees values only - leave the positions on the device
    cudaStream_t st_up = nullptr, st_dn = nullptr;
    cudaEvent_t ev_dn = nullptr;
    cudaEvent_t ev_up = nullptr, ev_scan = nullptr;
    struct Pending {
        bool live = false;
        int2 *d_pos = nullptr;
        uint32_t *d_val = nullptr;
        int64_t cap = 0, span = 0;
        // what was scanned (a block whose records overflowed cap is scanned again into an exact buffer)
        const uint16_t *win = nullptr;
        int64_t avail = 0, ctx = 0, limit = 0, base = 0;
    } pend;
    unsigned long long *h_total = nullptr;  // pinned: the match count of the pending block
    // literal WholeWord matchers (quirk Q7): a segment of the reference's loop can be as long as the input, so the feeds
    // are only collected and acgpu_stream_end scans the whole input in one call (every record arrives with the end)
    bool literal = false;
    std::vector<uint16_t> collected;
};

StreamCtx *as_stream(uint64_t h) {
    StreamCtx *s = reinterpret_cast<StreamCtx *>(static_cast<uintptr_t>(h));
    if (!s || s->magic != kStreamMagic) return nullptr;
    return s;
}

void stream_free(StreamCtx *s) {
    cudaSetDevice(s->m->device);
    if (s->st) cudaStreamSynchronize(s->st);
    for (int i = 0; i < 2; i++) {
        if (s->h_pin[i]) cudaFreeHost(s->h_pin[i]);
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
        if (s->d_win[i]) cudaFree(s->d_win[i]);
    }
    if (s->pend.d_pos) cudaFree(s->pend.d_pos);
    if (s->pend.d_val) cudaFree(s->pend.d_val);
    if (s->d_carry) cudaFree(s->d_carry);
    if (s->cpend.live) sel2_free(s->cpend.R);
    if (s->d_map) cudaFree(s->d_map);
    if (s->h_map) cudaFreeHost(s->h_map);
    if (s->ev_begin) cudaEventDestroy(s->ev_begin);
    if (s->d_total) cudaFree(s->d_total);
    if (s->h_total) cudaFreeHost(s->h_total);
    if (s->ev_up) cudaEventDestroy(s->ev_up);
    if (s->ev_scan) cudaEventDestroy(s->ev_scan);
    if (s->st_up) {
        cudaStreamSynchronize(s->st_up);
        cudaStreamDestroy(s->st_up);
    }
    if (s->st_dn) {
        cudaStreamSynchronize(s->st_dn);
        cudaStreamDestroy(s->st_dn);
    }
    if (s->ev_dn) cudaEventDestroy(s->ev_dn);
    if (s->st) cudaStreamDestroy(s->st);
    s->magic = 0;
    delete s;
}

int stream_reserve(StreamCtx *s, int which, int64_t chars) {
    if (s->win_cap[which] >= chars) return ACGPU_OK;
    int64_t cap = std::max<int64_t>(chars + chars / 4, 1 << 16);
    uint16_t *nw = nullptr;
    CU_TRY(cudaMalloc(reinterpret_cast<void **>(&nw), static_cast<size_t>(cap) * 2));
    if (which == s->cur && s->len > 0)
        CU_TRY(cudaMemcpyAsync(nw, s->d_win[which], static_cast<size_t>(s->len) * 2, cudaMemcpyDeviceToDevice, s->st));
    CU_TRY(cudaStreamSynchronize(s->st));
    if (s->d_win[which]) cudaFree(s->d_win[which]);
    s->d_win[which] = nw;
    s->win_cap[which] = cap;
    return ACGPU_OK;
}

// append host chars to the active window: straight DMA when the caller's buffer is page-locked (a pinned direct buffer
// on the Java side), else through the pinned double buffer
int stream_append(StreamCtx *s, const uint16_t *chars, int64_t n) {
    int rc = stream_reserve(s, s->cur, s->len + n);
    if (rc != ACGPU_OK) return rc;
    cudaStream_t up = (s->pipelined || s->chain_pipe) ? s->st_up : s->st;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, chars) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        CU_TRY(cudaMemcpyAsync(s->d_win[s->cur] + s->len, chars, static_cast<size_t>(n) * 2, cudaMemcpyHostToDevice, up));
        s->dma_pending = true;
        s->len += n;
        return ACGPU_OK;
    }
    cudaGetLastError();  // unregistered host memory reports an error on some drivers
    int64_t done = 0;
    int k = 0;
    while (done < n) {
        const int64_t c = std::min<int64_t>(static_cast<int64_t>(kPinChars), n - done);
        CU_TRY(cudaEventSynchronize(s->ev[k]));  // staging buffer k free again
        std::memcpy(s->h_pin[k], chars + done, static_cast<size_t>(c) * 2);
        CU_TRY(cudaMemcpyAsync(s->d_win[s->cur] + s->len + done, s->h_pin[k], static_cast<size_t>(c) * 2,
                               cudaMemcpyHostToDevice, up));
        CU_TRY(cudaEventRecord(s->ev[k], up));
        done += c;
        k ^= 1;
    }
    s->len += n;
    return ACGPU_OK;
}

// Longest / Shortest feeds on the start-mask path (kernel_sel2.cuh): every feed is a CHAIN SHARD (acgpu_chain_shard_*:
// the same kernels the multi-GPU shards use).  The window starts at the chain domain; a feed finalises whole tiles of it
// (8 192 positions, with 2 * max_len + 2 chars of look-ahead behind them), its composed map tells the number of records and
// the offset at which the chain enters the next block before a single record is written, the unfinalised tail moves to the
// front of the other window.  The window is cut to a multiple of 256 chars so that its index space starts at a tile boundary
// (every entry offset has a map row); the last block (end of the Readable) takes whatever is left.
int stream_process_chain(StreamCtx *s, bool final, acgpu_result *out) {
    Matcher *m = s->m;
    fill_empty(out);
    const int64_t avail = s->len;
    if (avail == 0) return ACGPU_OK;
    const int64_t D = 2 * static_cast<int64_t>(m->host.max_len) + 2;
    int64_t n = avail, n_domain = avail;
    if (!final) {
        n = avail & ~static_cast<int64_t>(kMaskRow - 1);
        n_domain = n > D ? (n - D) / kS2Tile * kS2Tile : 0;
        if (n_domain <= 0) return ACGPU_OK;   // not a whole tile yet: wait for the next feed
    }
    const uint16_t *win = s->d_win[s->cur];
    RunOpts opt;
    opt.pos_base = static_cast<int32_t>(static_cast<uint32_t>(s->base));
    Sel2Run R;
    int rc = sel2_setup(m, win, n, final ? -1 : n_domain / kS2Tile, s->st, opt, R);
    if (rc != ACGPU_OK) return rc;
    rc = sel2_masks(R);
    if (rc == ACGPU_OK) rc = sel2_maps(R);
    unsigned long long total = 0;
    int32_t exit_off = 0;
    uint32_t entry0 = static_cast<uint32_t>(s->entry);
    int64_t cap = 0;
    cudaError_t e = cudaSuccess;
    if (rc == ACGPU_OK && !final) {
        // R.moff == 0 (n is a multiple of 256 and the window is 256-byte aligned): the map has a row for every entry offset
        unsigned long long row = 0;
        rc = sel2_shard_map(R, s->d_map, s->d_total);
        if (rc == ACGPU_OK) e = cudaMemcpyAsync(&row, s->d_map + s->entry, 8, cudaMemcpyDeviceToHost, s->st);
        if (rc == ACGPU_OK && e == cudaSuccess) e = cudaStreamSynchronize(s->st);
        s->dma_pending = false;
        total = row >> 8;
        exit_off = static_cast<int32_t>(row & 0xFFu);
        cap = static_cast<int64_t>(total);
    } else if (rc == ACGPU_OK) {
        const int64_t idx = R.moff + s->entry;  // index-space position the chain enters at
        if (idx < kS2Ent) {
            entry0 = static_cast<uint32_t>(idx);
        } else {
            entry0 = 0;
            k_sel2_zero_prefix<<<1, 256, 0, s->st>>>(R.P.masks, R.moff, idx);
            e = cudaGetLastError();
            if (e == cudaSuccess) rc = sel2_maps(R, 1);
        }
        cap = n;   // non-overlapping matches: at most one per char
    }
    int2 *d_pos = nullptr;
    uint32_t *d_val = nullptr;
    if (rc == ACGPU_OK && e == cudaSuccess && cap > 0) {
        e = cudaMallocAsync(reinterpret_cast<void **>(&d_pos), static_cast<size_t>(cap) * 8, s->st);
        if (e == cudaSuccess && m->host.is_map) e = cudaMallocAsync(reinterpret_cast<void **>(&d_val), static_cast<size_t>(cap) * 4, s->st);
        if (e == cudaSuccess) rc = sel2_records(R, entry0, d_pos, d_val, cap, s->d_total);
        if (rc == ACGPU_OK && e == cudaSuccess && final) {
            e = cudaMemcpyAsync(&total, s->d_total, 8, cudaMemcpyDeviceToHost, s->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
            s->dma_pending = false;
        }
    }
    if (rc == ACGPU_OK && e == cudaSuccess && total > 0) {
        PinnedBlock blk;
        const bool want_pos = !(s->values_only && m->host.is_map);
        const size_t pos_bytes = want_pos ? align_up(static_cast<size_t>(total) * 8, 16) : 0;
        if (!pin_take(pos_bytes + (m->host.is_map ? static_cast<size_t>(total) * 4 : 0), &blk)) {
            rc = fail(ACGPU_ENOMEM, "out of pinned host memory for the match records");
        } else {
            int32_t *h_pos = want_pos ? static_cast<int32_t *>(blk.p) : nullptr;
            uint32_t *h_val = m->host.is_map ? reinterpret_cast<uint32_t *>(static_cast<char *>(blk.p) + pos_bytes) : nullptr;
            if (want_pos) e = cudaMemcpyAsync(h_pos, d_pos, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, s->st);
            if (e == cudaSuccess && h_val) e = cudaMemcpyAsync(h_val, d_val, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, s->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
            if (e != cudaSuccess) {
                pin_release(blk.p);
            } else {
                out->n = static_cast<int64_t>(total);
                out->pos = h_pos;
                out->val = h_val;
            }
        }
    }
    if (d_pos) cudaFreeAsync(d_pos, s->st);
    if (d_val) cudaFreeAsync(d_val, s->st);
    sel2_free(R);
    if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream block: ") + cudaGetErrorString(e));
    if (rc != ACGPU_OK) return rc;
    // slide: the unfinalised tail moves to the front of the other window (256-byte aligned: the next block's index space)
    const int64_t keep = avail - n_domain;
    const int other = s->cur ^ 1;
    rc = stream_reserve(s, other, std::max<int64_t>(keep, 1));
    if (rc != ACGPU_OK) return rc;
    if (keep > 0)
        CU_TRY(cudaMemcpyAsync(s->d_win[other], s->d_win[s->cur] + n_domain, static_cast<size_t>(keep) * 2, cudaMemcpyDeviceToDevice, s->st));
    s->cur = other;
    s->base += n_domain;
    s->len = keep;
    s->ctx = 0;
    s->entry = exit_off;
    s->chain = s->base + exit_off;
    return ACGPU_OK;
}

// scan what can be finalised, copy the records out, slide the window
int stream_process(StreamCtx *s, bool final, acgpu_result *out) {
    if (s->chain_blocks) return stream_process_chain(s, final, out);
    Matcher *m = s->m;
    const int family = m->host.family;
    const int64_t L = m->host.max_len;
    const int64_t avail = s->len;
    int64_t limit;  // window positions [ctx, limit) are finalised by this call
    if (family == ACGPU_AHOCORASICK) {
        limit = avail;
    } else {
        const int64_t D = 2 * L + 2;
        limit = final ? avail : std::max<int64_t>(s->ctx, avail - D);
    }
    const bool plain_words = family == ACGPU_WHOLEWORD || (family == ACGPU_WHOLEWORDLONGEST && m->use_ww);
    fill_empty(out);
    if (limit <= s->ctx) return ACGPU_OK;

    // record buffer: a quarter record per char, or 1.25 x the densest feed so far - a feed that overflows is scanned twice
    const int64_t span = limit - s->ctx;
    int64_t cap = std::max<int64_t>(1 << 16, std::max<int64_t>(span / 4, static_cast<int64_t>(1.25 * s->density * static_cast<double>(span)) + 1024));
    unsigned long long total = 0;
    int2 *d_pos = nullptr;
    uint32_t *d_val = nullptr;
    for (int attempt = 0; attempt < 2; attempt++) {
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos), static_cast<size_t>(cap) * 8, s->st));
        if (m->host.is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_val), static_cast<size_t>(cap) * 4, s->st));
        RunOpts opt;
        opt.pos_base = static_cast<int32_t>(static_cast<uint32_t>(s->base));
        opt.ctx = s->ctx;
        int rc;
        if (family == ACGPU_AHOCORASICK) {
            rc = enqueue_match(m, s->d_win[s->cur], avail, s->ctx, avail, d_pos, d_val, cap, s->d_total, s->st, opt);
        } else {
            opt.entry0 = s->chain - s->base;
            opt.chain_n = limit;
            opt.d_carry = s->d_carry;
            opt.abs0 = s->base == 0 ? 0 : -1;
            rc = enqueue_match(m, s->d_win[s->cur], avail, 0, avail, d_pos, d_val, cap, s->d_total, s->st, opt);
        }
        if (rc != ACGPU_OK) return rc;
        CU_TRY(cudaMemcpyAsync(&total, s->d_total, 8, cudaMemcpyDeviceToHost, s->st));
        CU_TRY(cudaStreamSynchronize(s->st));
        s->dma_pending = false;
        if (static_cast<int64_t>(total) <= cap) break;
        cudaFreeAsync(d_pos, s->st);
        if (d_val) cudaFreeAsync(d_val, s->st);
        d_pos = nullptr;
        d_val = nullptr;
        cap = static_cast<int64_t>(total);
    }
    s->density = std::max(s->density, static_cast<double>(total) / static_cast<double>(std::max<int64_t>(span, 1)));
    if (total > 0) {
        // one page-locked block from the process-wide cache: positions, then values (acgpu_free_result hands it back)
        PinnedBlock blk;
        const bool want_pos = !(s->values_only && m->host.is_map);
        const size_t pos_bytes = want_pos ? align_up(static_cast<size_t>(total) * 8, 16) : 0;
        if (!pin_take(pos_bytes + (m->host.is_map ? static_cast<size_t>(total) * 4 : 0), &blk))
            return fail(ACGPU_ENOMEM, "out of pinned host memory for the match records");
        int32_t *h_pos = want_pos ? static_cast<int32_t *>(blk.p) : nullptr;
        uint32_t *h_val = m->host.is_map ? reinterpret_cast<uint32_t *>(static_cast<char *>(blk.p) + pos_bytes) : nullptr;
        cudaError_t e = want_pos ? cudaMemcpyAsync(h_pos, d_pos, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, s->st) : cudaSuccess;
        if (e == cudaSuccess && h_val) e = cudaMemcpyAsync(h_val, d_val, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, s->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
        if (e != cudaSuccess) {
            pin_release(blk.p);
            return fail(ACGPU_ECUDA, std::string("stream download: ") + cudaGetErrorString(e));
        }
        out->n = static_cast<int64_t>(total);
        out->pos = h_pos;
        out->val = h_val;
    }
    cudaFreeAsync(d_pos, s->st);
    if (d_val) cudaFreeAsync(d_val, s->st);

    // chain position for the next block
    if (!plain_words && family != ACGPU_AHOCORASICK) {
        int64_t carry = -1;
        CU_TRY(cudaMemcpyAsync(&carry, s->d_carry, sizeof(carry), cudaMemcpyDeviceToHost, s->st));
        CU_TRY(cudaStreamSynchronize(s->st));
        if (carry >= 0) s->chain = s->base + carry;  // else: the chain already stood beyond this block
    } else {
        s->chain = s->base + limit;
    }

    // slide: keep the context the next block needs at the front of the other window
    int64_t keep_from;
    if (family == ACGPU_AHOCORASICK) {
        keep_from = std::max<int64_t>(0, avail - std::max<int64_t>(0, L - 1));
    } else {
        keep_from = std::max<int64_t>(0, limit - 1);
    }
    const int64_t keep = avail - keep_from;
    const int other = s->cur ^ 1;
    int rc = stream_reserve(s, other, std::max<int64_t>(keep, 1));
    if (rc != ACGPU_OK) return rc;
    if (keep > 0)
        CU_TRY(cudaMemcpyAsync(s->d_win[other], s->d_win[s->cur] + keep_from, static_cast<size_t>(keep) * 2,
                               cudaMemcpyDeviceToDevice, s->st));
    s->cur = other;
    s->base += keep_from;
    s->len = keep;
    s->ctx = (family == ACGPU_AHOCORASICK) ? keep : std::min<int64_t>(keep, limit - keep_from);
    return ACGPU_OK;
}

// ---- pipelined feeds -----------------------------------------------------------------------------------------

int stream_scan_block(StreamCtx *s, const StreamCtx::Pending &b, int2 *d_pos, uint32_t *d_val, int64_t cap) {
    Matcher *m = s->m;
    RunOpts opt;
    opt.pos_base = static_cast<int32_t>(static_cast<uint32_t>(b.base));
    opt.ctx = b.ctx;
    if (m->host.family == ACGPU_AHOCORASICK)
        return enqueue_match(m, b.win, b.avail, b.ctx, b.avail, d_pos, d_val, cap, s->d_total, s->st, opt);
    opt.entry0 = 0;
    opt.chain_n = b.limit;
    opt.abs0 = b.base == 0 ? 0 : -1;
    return enqueue_match(m, b.win, b.avail, 0, b.avail, d_pos, d_val, cap, s->d_total, s->st, opt);
}

// The records of the pending block, in two steps so that the scan of the NEXT block can be enqueued in between:
//   begin  waits for the block's scan (its count is in pinned memory), takes a pinned block and starts the download on st_dn
//   end    waits for the download and hands the block out
struct Collected {
    bool live = false;
    int64_t total = 0;
    PinnedBlock blk;
    int32_t *h_pos = nullptr;
    uint32_t *h_val = nullptr;
    int2 *d_pos = nullptr;
    uint32_t *d_val = nullptr;
};

int stream_collect_begin(StreamCtx *s, Collected &c) {
    StreamCtx::Pending &b = s->pend;
    if (!b.live) return ACGPU_OK;
    b.live = false;
    Matcher *m = s->m;
    CU_TRY(cudaEventSynchronize(s->ev_scan));
    const int64_t total = static_cast<int64_t>(*s->h_total);
    if (total > b.cap) {
        // denser than the buffer: the kernels counted everything; scan the block again (its window is still intact)
        cudaFreeAsync(b.d_pos, s->st);
        if (b.d_val) cudaFreeAsync(b.d_val, s->st);
        b.d_pos = nullptr;
        b.d_val = nullptr;
        b.cap = total;
        CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&b.d_pos), static_cast<size_t>(total) * 8, s->st));
        if (m->host.is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&b.d_val), static_cast<size_t>(total) * 4, s->st));
        int rc = stream_scan_block(s, b, b.d_pos, b.d_val, b.cap);
        if (rc != ACGPU_OK) return rc;
        CU_TRY(cudaEventRecord(s->ev_scan, s->st));
    }
    s->density = std::max(s->density, static_cast<double>(total) / static_cast<double>(std::max<int64_t>(b.span, 1)));
    c.d_pos = b.d_pos;
    c.d_val = b.d_val;
    b.d_pos = nullptr;
    b.d_val = nullptr;
    c.total = total;
    c.live = true;
    if (total > 0) {
        const bool want_pos = !(s->values_only && m->host.is_map);
        const size_t pos_bytes = want_pos ? align_up(static_cast<size_t>(total) * 8, 16) : 0;
        if (!pin_take(pos_bytes + (m->host.is_map ? static_cast<size_t>(total) * 4 : 0), &c.blk))
            return fail(ACGPU_ENOMEM, "out of pinned host memory for the match records");
        c.h_pos = want_pos ? static_cast<int32_t *>(c.blk.p) : nullptr;
        c.h_val = m->host.is_map ? reinterpret_cast<uint32_t *>(static_cast<char *>(c.blk.p) + pos_bytes) : nullptr;
        CU_TRY(cudaStreamWaitEvent(s->st_dn, s->ev_scan, 0));
        if (c.h_pos) CU_TRY(cudaMemcpyAsync(c.h_pos, c.d_pos, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, s->st_dn));
        if (c.h_val) CU_TRY(cudaMemcpyAsync(c.h_val, c.d_val, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, s->st_dn));
    }
    CU_TRY(cudaEventRecord(s->ev_dn, s->st_dn));
    return ACGPU_OK;
}

int stream_collect_end(StreamCtx *s, Collected &c, int rc, acgpu_result *out) {
    fill_empty(out);
    if (!c.live) return rc;
    c.live = false;
    const cudaError_t e = cudaEventSynchronize(s->ev_dn);
    if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream download: ") + cudaGetErrorString(e));
    // the download has finished (host wait above): free on the stream the NEXT block allocates on, so that the pool hands the same
    // memory out again (a free on st_dn is only reusable on st when the allocator happens to see it completed - feeds then
    // grew the pool by a record buffer each, at random)
    if (c.d_pos) cudaFreeAsync(c.d_pos, s->st);
    if (c.d_val) cudaFreeAsync(c.d_val, s->st);
    if (rc == ACGPU_OK && c.total > 0) {
        out->n = c.total;
        out->pos = c.h_pos;
        out->val = c.h_val;
    } else if (c.blk.p) {
        pin_release(c.blk.p);
    }
    return rc;
}

int stream_collect(StreamCtx *s, acgpu_result *out) {
    Collected c;
    const int rc = stream_collect_begin(s, c);
    return stream_collect_end(s, c, rc, out);
}

// enqueue the scan of what the window can finalise (no host wait), remember it as the pending block, slide the window
int stream_enqueue(StreamCtx *s, bool final) {
    Matcher *m = s->m;
    const int family = m->host.family;
    const int64_t L = m->host.max_len;
    const int64_t avail = s->len;
    const int64_t limit = family == ACGPU_AHOCORASICK ? avail : (final ? avail : std::max<int64_t>(s->ctx, avail - (2 * L + 2)));
    CU_TRY(cudaEventRecord(s->ev_up, s->st_up));
    CU_TRY(cudaStreamWaitEvent(s->st, s->ev_up, 0));  // the block has landed before anything on st touches the window
    if (limit <= s->ctx) return ACGPU_OK;
    StreamCtx::Pending &b = s->pend;
    b.span = limit - s->ctx;
    b.cap = std::max<int64_t>(1 << 16, std::max<int64_t>(b.span / 4, static_cast<int64_t>(1.25 * s->density * static_cast<double>(b.span)) + 1024));
    b.win = s->d_win[s->cur];
    b.avail = avail;
    b.ctx = s->ctx;
    b.limit = limit;
    b.base = s->base;
    CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&b.d_pos), static_cast<size_t>(b.cap) * 8, s->st));
    if (m->host.is_map) CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&b.d_val), static_cast<size_t>(b.cap) * 4, s->st));
    int rc = stream_scan_block(s, b, b.d_pos, b.d_val, b.cap);
    if (rc != ACGPU_OK) return rc;
    CU_TRY(cudaMemcpyAsync(s->h_total, s->d_total, 8, cudaMemcpyDeviceToHost, s->st));
    CU_TRY(cudaEventRecord(s->ev_scan, s->st));
    b.live = true;
    s->chain = s->base + limit;
    // slide: keep the context the next block needs at the front of the other window (the scanned window stays intact
    // until the slide after the NEXT scan: an overflowing block can be scanned again)
    const int64_t keep_from = family == ACGPU_AHOCORASICK ? std::max<int64_t>(0, avail - std::max<int64_t>(0, L - 1))
                                                            : std::max<int64_t>(0, limit - 1);
    const int64_t keep = avail - keep_from;
    const int other = s->cur ^ 1;
    rc = stream_reserve(s, other, std::max<int64_t>(keep, 1));
    if (rc != ACGPU_OK) return rc;
    if (keep > 0)
        CU_TRY(cudaMemcpyAsync(s->d_win[other], s->d_win[s->cur] + keep_from, static_cast<size_t>(keep) * 2,
                               cudaMemcpyDeviceToDevice, s->st));
    s->cur = other;
    s->base += keep_from;
    s->len = keep;
    s->ctx = (family == ACGPU_AHOCORASICK) ? keep : std::min<int64_t>(keep, limit - keep_from);
    return ACGPU_OK;
}

// ---- pipelined chain-shard feeds: begin(block k) = masks, maps and the composed map, enqueued without a host wait;
//      finish(block k) = records for the entry offset block k-1 handed over, cut when feed k+1 arrives - so the upload of block
//      k+1 (st_up) runs next to the record kernels and the download (st_dn) of block k.

// the records of the pending block: kernels on st, download on st_dn; the caller waits for ev_dn (stream_chain_wait)
int stream_chain_finish(StreamCtx *s, Collected &c) {
    StreamCtx::ChainPending &b = s->cpend;
    if (!b.live) return ACGPU_OK;
    b.live = false;
    Matcher *m = s->m;
    int rc = ACGPU_OK;
    cudaError_t e = cudaEventSynchronize(s->ev_begin);   // the map is in pinned memory
    const unsigned long long row = s->h_map[s->entry];
    const int64_t total = static_cast<int64_t>(row >> 8);
    const int32_t exit_off = static_cast<int32_t>(row & 0xFFu);
    if (e == cudaSuccess && total > 0) {
        e = cudaMallocAsync(reinterpret_cast<void **>(&c.d_pos), static_cast<size_t>(total) * 8, s->st);
        if (e == cudaSuccess && m->host.is_map) e = cudaMallocAsync(reinterpret_cast<void **>(&c.d_val), static_cast<size_t>(total) * 4, s->st);
        if (e == cudaSuccess) rc = sel2_records(b.R, static_cast<uint32_t>(s->entry), c.d_pos, c.d_val, total, s->d_total);
        if (e == cudaSuccess) e = cudaEventRecord(s->ev_scan, s->st);
        const bool want_pos = !(s->values_only && m->host.is_map);
        const size_t pos_bytes = want_pos ? align_up(static_cast<size_t>(total) * 8, 16) : 0;
        if (rc == ACGPU_OK && e == cudaSuccess && !pin_take(pos_bytes + (m->host.is_map ? static_cast<size_t>(total) * 4 : 0), &c.blk))
            rc = fail(ACGPU_ENOMEM, "out of pinned host memory for the match records");
        if (rc == ACGPU_OK && e == cudaSuccess) {
            c.h_pos = want_pos ? static_cast<int32_t *>(c.blk.p) : nullptr;
            c.h_val = m->host.is_map ? reinterpret_cast<uint32_t *>(static_cast<char *>(c.blk.p) + pos_bytes) : nullptr;
            e = cudaStreamWaitEvent(s->st_dn, s->ev_scan, 0);
            if (e == cudaSuccess && c.h_pos) e = cudaMemcpyAsync(c.h_pos, c.d_pos, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, s->st_dn);
            if (e == cudaSuccess && c.h_val) e = cudaMemcpyAsync(c.h_val, c.d_val, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, s->st_dn);
        }
    }
    sel2_free(b.R);
    if (e == cudaSuccess) e = cudaEventRecord(s->ev_dn, s->st_dn);
    c.total = total;
    c.live = true;
    s->entry = exit_off;
    s->chain = s->base + exit_off;   // s->base already stands at the end of the block's domain
    if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream block: ") + cudaGetErrorString(e));
    return rc;
}

// masks, maps and the composed map of what the window can finalise (no host wait); slide the window
int stream_chain_begin(StreamCtx *s) {
    Matcher *m = s->m;
    CU_TRY(cudaEventRecord(s->ev_up, s->st_up));
    CU_TRY(cudaStreamWaitEvent(s->st, s->ev_up, 0));   // the block has landed before anything on st touches the window
    const int64_t avail = s->len;
    const int64_t D = 2 * static_cast<int64_t>(m->host.max_len) + 2;
    const int64_t n = avail & ~static_cast<int64_t>(kMaskRow - 1);
    const int64_t n_domain = n > D ? (n - D) / kS2Tile * kS2Tile : 0;
    if (n_domain <= 0) return ACGPU_OK;   // not a whole tile yet
    StreamCtx::ChainPending &b = s->cpend;
    RunOpts opt;
    opt.pos_base = static_cast<int32_t>(static_cast<uint32_t>(s->base));
    b.R = Sel2Run();
    int rc = sel2_setup(m, s->d_win[s->cur], n, n_domain / kS2Tile, s->st, opt, b.R);
    if (rc != ACGPU_OK) return rc;
    rc = sel2_masks(b.R);
    if (rc == ACGPU_OK) rc = sel2_maps(b.R);
    if (rc == ACGPU_OK) rc = sel2_shard_map(b.R, s->d_map, s->d_total);
    cudaError_t e = cudaSuccess;
    if (rc == ACGPU_OK) e = cudaMemcpyAsync(s->h_map, s->d_map, kS2Ent * 8, cudaMemcpyDeviceToHost, s->st);
    if (rc == ACGPU_OK && e == cudaSuccess) e = cudaEventRecord(s->ev_begin, s->st);
    if (rc != ACGPU_OK || e != cudaSuccess) {
        sel2_free(b.R);
        return rc != ACGPU_OK ? rc : fail(ACGPU_ECUDA, std::string("stream block: ") + cudaGetErrorString(e));
    }
    b.n_domain = n_domain;
    b.live = true;
    // slide: the unfinalised tail moves to the front of the other window; the scanned window stays intact until the block's
    // records are cut (the next feed), and is written again only by the feed after that
    const int64_t keep = avail - n_domain;
    const int other = s->cur ^ 1;
    rc = stream_reserve(s, other, std::max<int64_t>(keep, 1));
    if (rc != ACGPU_OK) return rc;
    if (keep > 0)
        CU_TRY(cudaMemcpyAsync(s->d_win[other], s->d_win[s->cur] + n_domain, static_cast<size_t>(keep) * 2, cudaMemcpyDeviceToDevice, s->st));
    s->cur = other;
    s->base += n_domain;
    s->len = keep;
    s->ctx = 0;
    return ACGPU_OK;
}

// two results -> one (end(): the pending block, then the final flush)
int merge_results(acgpu_result *a, acgpu_result *b, bool is_map, acgpu_result *out) {
    if (a->n == 0) {
        *out = *b;
        return ACGPU_OK;
    }
    if (b->n == 0) {
        *out = *a;
        return ACGPU_OK;
    }
    const int64_t n = a->n + b->n;
    PinnedBlock blk;
    const bool want_pos = a->pos && b->pos;
    const size_t pos_bytes = want_pos ? align_up(static_cast<size_t>(n) * 8, 16) : 0;
    if (!pin_take(pos_bytes + (is_map ? static_cast<size_t>(n) * 4 : 0), &blk)) return fail(ACGPU_ENOMEM, "out of pinned host memory");
    int32_t *pos = want_pos ? static_cast<int32_t *>(blk.p) : nullptr;
    if (want_pos) {
        std::memcpy(pos, a->pos, static_cast<size_t>(a->n) * 8);
        std::memcpy(pos + 2 * a->n, b->pos, static_cast<size_t>(b->n) * 8);
    }
    uint32_t *val = nullptr;
    if (is_map) {
        val = reinterpret_cast<uint32_t *>(static_cast<char *>(blk.p) + pos_bytes);
        std::memcpy(val, a->val, static_cast<size_t>(a->n) * 4);
        std::memcpy(val + a->n, b->val, static_cast<size_t>(b->n) * 4);
    }
    acgpu_free_result(a);
    acgpu_free_result(b);
    out->n = n;
    out->pos = pos;
    out->val = val;
    return ACGPU_OK;
}

}  // namespace

int acgpu_stream_begin(uint64_t handle, uint64_t *stream_handle) {
    Matcher *m = as_matcher(handle);
    if (!m || !stream_handle) return fail(ACGPU_EINVAL, "bad arguments");
    *stream_handle = 0;
    int rc = ensure_device(m);
    if (rc != ACGPU_OK) return rc;
    StreamCtx *s = new (std::nothrow) StreamCtx();
    if (!s) return fail(ACGPU_ENOMEM, "out of memory");
    s->m = m;
    if (m->literal_family >= 0) {
        s->literal = true;
        *stream_handle = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(s));
        return ACGPU_OK;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMallocHost(reinterpret_cast<void **>(&s->h_pin[i]), kPinChars * 2);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s->d_carry), 16);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s->d_total), 8);
    {
        const char *gen = getenv("ACGPU_STREAM_CHAIN_GEN");  // ACGPU_STREAM_CHAIN_GEN=1: the generation-1 kernels per feed (A/B runs)
        s->chain_blocks = (m->host.family == ACGPU_LONGEST || m->host.family == ACGPU_SHORTEST) && m->use_tier && !(gen && gen[0] == '1');
        if (s->chain_blocks && e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&s->d_map), kS2Ent * 8);
        const char *sync = getenv("ACGPU_STREAM_SYNC");
        s->chain_pipe = s->chain_blocks && !(sync && sync[0] == '1');
        if (s->chain_pipe) {
            if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&s->h_map), kS2Ent * 8, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_begin, cudaEventDisableTiming);
        }
    }
    {
        const int fam = m->host.family;
        const char *sync = getenv("ACGPU_STREAM_SYNC");
        s->pipelined = (fam == ACGPU_AHOCORASICK || fam == ACGPU_WHOLEWORD || (fam == ACGPU_WHOLEWORDLONGEST && m->use_ww)) &&
                       !(sync && sync[0] == '1');
    }
    if (s->pipelined || s->chain_pipe) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st_up, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->st_dn, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_dn, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_up, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_scan, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&s->h_total), 8, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        stream_free(s);
        return fail(ACGPU_ECUDA, std::string("stream setup: ") + cudaGetErrorString(e));
    }
    *stream_handle = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(s));
    return ACGPU_OK;
}

int acgpu_stream_set_values_only(uint64_t stream_handle, int on) {
    StreamCtx *s = as_stream(stream_handle);
    if (!s) return fail(ACGPU_EINVAL, "bad stream handle");
    if (!s->m->host.is_map && on) return fail(ACGPU_EINVAL, "a Set stream has no values");
    s->values_only = on != 0;
    return ACGPU_OK;
}

int acgpu_stream_feed(uint64_t stream_handle, const uint16_t *chars, int32_t n, acgpu_result *out) {
    StreamCtx *s = as_stream(stream_handle);
    if (!s || !out || n < 0 || (n > 0 && !chars)) return fail(ACGPU_EINVAL, "bad arguments");
    fill_empty(out);
    CU_TRY(cudaSetDevice(s->m->device));
    if (n == 0) return ACGPU_OK;
    if (s->literal) {
        if (s->collected.size() + static_cast<size_t>(n) > 0x7FFFFFFFull) return fail(ACGPU_EINVAL, "stream longer than a Java int");
        try {
            s->collected.insert(s->collected.end(), chars, chars + n);
        } catch (const std::bad_alloc &) {
            return fail(ACGPU_ENOMEM, "out of memory collecting the stream");
        }
        return ACGPU_OK;
    }
    int rc = stream_append(s, chars, n);
    if (s->chain_pipe) {
        // upload of this block (st_up) || record kernels and download of the previous block (st, st_dn); then this block's
        // masks and maps are enqueued and the call returns without waiting for them
        Collected col;
        if (rc == ACGPU_OK) rc = stream_chain_finish(s, col);
        if (rc == ACGPU_OK) rc = stream_chain_begin(s);
        rc = stream_collect_end(s, col, rc, out);
        if (s->dma_pending || rc != ACGPU_OK) {
            s->dma_pending = false;
            const cudaError_t e = cudaStreamSynchronize(s->st_up);  // the caller's buffer is free again
            if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream upload: ") + cudaGetErrorString(e));
        }
        return rc;
    }
    if (s->pipelined) {
        // upload of this block (st_up) || download of the previous block's records (st); then this block's scan is
        // enqueued and the call returns without waiting for it
        Collected col;
        if (rc == ACGPU_OK) rc = stream_collect_begin(s, col);   // download of block k-1 on st_dn ...
        if (rc == ACGPU_OK) rc = stream_enqueue(s, false);       // ... while block k is scanned on st
        rc = stream_collect_end(s, col, rc, out);
        if (s->dma_pending || rc != ACGPU_OK) {
            s->dma_pending = false;
            const cudaError_t e = cudaStreamSynchronize(s->st_up);  // the caller's buffer is free again
            if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream upload: ") + cudaGetErrorString(e));
        }
        return rc;
    }
    if (rc == ACGPU_OK) rc = stream_process(s, false, out);
    if (s->dma_pending) {
        // nothing was finalised (a feed shorter than the look-ahead) or an error cut the call short: the upload still
        // reads the caller's buffer - wait for it, the caller may refill the buffer as soon as we return
        s->dma_pending = false;
        const cudaError_t e = cudaStreamSynchronize(s->st);
        if (e != cudaSuccess && rc == ACGPU_OK) rc = fail(ACGPU_ECUDA, std::string("stream upload: ") + cudaGetErrorString(e));
    }
    return rc;
}

int acgpu_stream_end(uint64_t stream_handle, acgpu_result *out) {
    StreamCtx *s = as_stream(stream_handle);
    if (!s) return fail(ACGPU_EINVAL, "bad stream handle");
    int rc = ACGPU_OK;
    if (out) {
        fill_empty(out);
        cudaSetDevice(s->m->device);
        if (s->literal) {
            if (!s->collected.empty()) {
                HostCall hc(s->m);
                hc.readable_view = true;
                rc = hc.init();
                if (rc == ACGPU_OK) rc = hc.run_whole(s->collected.data(), static_cast<int64_t>(s->collected.size()));
                rc = hc.finish(rc, out);
            }
        } else if (s->chain_pipe) {
            acgpu_result a, b;
            fill_empty(&b);
            Collected col;
            rc = stream_chain_finish(s, col);
            rc = stream_collect_end(s, col, rc, &a);
            if (rc == ACGPU_OK) {
                // whatever is left (less than a tile + look-ahead, or everything of a short Readable), synchronously
                cudaEventRecord(s->ev_up, s->st_up);
                cudaStreamWaitEvent(s->st, s->ev_up, 0);
                rc = stream_process_chain(s, true, &b);
            }
            if (rc == ACGPU_OK) rc = merge_results(&a, &b, s->m->host.is_map, out);
            if (rc != ACGPU_OK) {
                acgpu_free_result(&a);
                acgpu_free_result(&b);
            }
        } else if (s->pipelined) {
            acgpu_result a, b;
            fill_empty(&b);
            rc = stream_collect(s, &a);
            if (rc == ACGPU_OK) rc = stream_enqueue(s, true);
            if (rc == ACGPU_OK) rc = stream_collect(s, &b);
            if (rc == ACGPU_OK) rc = merge_results(&a, &b, s->m->host.is_map, out);
            if (rc != ACGPU_OK) {
                acgpu_free_result(&a);
                acgpu_free_result(&b);
            }
        } else {
            rc = stream_process(s, true, out);
        }
    }
    stream_free(s);
    return rc;
}

}  // extern "C"
