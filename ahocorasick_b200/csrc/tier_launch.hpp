// Launch entry points of the k_tier_mask instantiations.  Every K lives in its own translation unit (tier_inst.cu
// compiled with -DTIER_K=k) so the kernels build in parallel.
#pragma once
#include <cuda_runtime.h>

#include "device_tables.cuh"

namespace acgpu {

struct DevTier;
struct MaskArgs;
struct FuseArgs;
struct DuoArgs;

// Tables that every probe gathers from (child masks + deep table): the launches ask L2 to keep them resident while the
// haystack, mask and record streams pass through (cudaLaunchAttributeAccessPolicyWindow; bytes == 0: no window).
struct L2Window {
    const void *base = nullptr;
    size_t bytes = 0;
    float hit_ratio = 1.0f;
};

// mir: the kernel reads the haystack right to left (start anchors over the forward trie: Longest / Shortest).
// low: 0 = every level below K may hold keywords, 1 = only level K-1 does (and rides in the level-K rows), 2 = none does.
// pair: k_tier_pair (kernel_pair.cuh, pair rows) instead of k_tier_mask.
// k_tier_mask<K, LOW> (kernel_mask.cuh): persistent, one CTA per SM; no inter-CTA waiting, plain launch.
// fuse_launch_k: k_tier_fused<K, LOW, isMap> (kernel_fuse.cuh) - masks and records in one persistent launch.
// duo_launch_k: k_tier_duo<K, LOW, isMap> (kernel_fuse.cuh) - the masks of one slab and the records of the slab before it.
#define ACGPU_DECLARE_MASK(k) \
    cudaError_t mask_launch_##k(int low, bool mir, bool pair, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, int grid, size_t smem, const L2Window &W, cudaStream_t st); \
    cudaError_t fuse_launch_##k(int low, bool is_map, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const FuseArgs &F, int grid, size_t smem, cudaStream_t st); \
    cudaError_t duo_launch_##k(int low, bool is_map, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const DuoArgs &D, int grid, size_t smem, cudaStream_t st);
ACGPU_DECLARE_MASK(1)
ACGPU_DECLARE_MASK(2)
ACGPU_DECLARE_MASK(3)
ACGPU_DECLARE_MASK(4)
ACGPU_DECLARE_MASK(5)
ACGPU_DECLARE_MASK(6)
ACGPU_DECLARE_MASK(7)
ACGPU_DECLARE_MASK(8)
#undef ACGPU_DECLARE_MASK

}  // namespace acgpu
