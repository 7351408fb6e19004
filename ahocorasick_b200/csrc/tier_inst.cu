// One translation unit per K: nvcc ... -DTIER_K=k tier_inst.cu
#include "kernel_tier.cuh"
#include "kernel_mask.cuh"
#include "kernel_pair.cuh"
#include "kernel_fuse.cuh"
#include "tier_launch.hpp"

#ifndef TIER_K
#error "compile with -DTIER_K=1..8"
#endif

namespace acgpu {

namespace {

template <typename Kern>
cudaError_t launch_windowed(Kern kern, int grid, int block, size_t smem, const L2Window &W, cudaStream_t st, const DevAutomaton &A,
                            const DevTier &T, const MaskArgs &P) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (W.bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void *>(W.base);
        attr[0].val.accessPolicyWindow.num_bytes = W.bytes;
        attr[0].val.accessPolicyWindow.hitRatio = W.hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, A, T, P);
}

template <int LOW, bool MIR, bool PAIR>
cudaError_t launch_mask_variant(const DevAutomaton &A, const DevTier &T, const MaskArgs &P, int grid, size_t smem, const L2Window &W,
                                cudaStream_t st) {
    static size_t attr_smem[64] = {0};
    auto kern = PAIR ? k_tier_pair<TIER_K, LOW, MIR> : k_tier_mask<TIER_K, LOW, MIR>;
    const void *fn = reinterpret_cast<const void *>(kern);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || attr_smem[dev] < smem) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_smem[dev] = smem;
    }
    return launch_windowed(kern, grid, kMaskThreads, smem, W, st, A, T, P);
}

template <int LOW, bool MAP>
cudaError_t launch_fuse_variant(const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const FuseArgs &F, int grid, size_t smem, cudaStream_t st) {
    static size_t attr_smem[64] = {0};
    auto kern = k_tier_fused<TIER_K, LOW, MAP>;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || attr_smem[dev] < smem) {
        e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_smem[dev] = smem;
    }
    kern<<<grid, kMaskThreads, smem, st>>>(A, T, P, F);
    return cudaGetLastError();
}

template <int LOW, bool MAP>
cudaError_t launch_duo_variant(const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const DuoArgs &D, int grid, size_t smem, cudaStream_t st) {
    static size_t attr_smem[64] = {0};
    auto kern = k_tier_duo<TIER_K, LOW, MAP>;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || attr_smem[dev] < smem) {
        e = cudaFuncSetAttribute(reinterpret_cast<const void *>(kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_smem[dev] = smem;
    }
    kern<<<grid, kMaskThreads, smem, st>>>(A, T, P, D);
    return cudaGetLastError();
}

}  // namespace

#define ACGPU_CAT2(a, b) a##b
#define ACGPU_CAT(a, b) ACGPU_CAT2(a, b)

cudaError_t ACGPU_CAT(mask_launch_, TIER_K)(int low, bool mir, bool pair, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, int grid,
                                            size_t smem, const L2Window &W, cudaStream_t st) {
    if (TIER_K == 1) low = 2;
#define ACGPU_MASK_CASE(L, M, Q) \
    if (low == L && mir == M && pair == Q) return launch_mask_variant<L, M, Q>(A, T, P, grid, smem, W, st);
    if (low < 0 || low > 2) low = 2;
    ACGPU_MASK_CASE(0, false, false) ACGPU_MASK_CASE(1, false, false) ACGPU_MASK_CASE(2, false, false)
    ACGPU_MASK_CASE(0, true, false) ACGPU_MASK_CASE(1, true, false) ACGPU_MASK_CASE(2, true, false)
    ACGPU_MASK_CASE(0, false, true) ACGPU_MASK_CASE(1, false, true) ACGPU_MASK_CASE(2, false, true)
    ACGPU_MASK_CASE(0, true, true) ACGPU_MASK_CASE(1, true, true) ACGPU_MASK_CASE(2, true, true)
#undef ACGPU_MASK_CASE
    return cudaErrorInvalidValue;
}

cudaError_t ACGPU_CAT(fuse_launch_, TIER_K)(int low, bool is_map, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const FuseArgs &F, int grid,
                                            size_t smem, cudaStream_t st) {
    if (TIER_K == 1 || low < 0 || low > 2) low = 2;
#define ACGPU_FUSE_CASE(L, M) \
    if (low == L && is_map == M) return launch_fuse_variant<L, M>(A, T, P, F, grid, smem, st);
    ACGPU_FUSE_CASE(0, false) ACGPU_FUSE_CASE(1, false) ACGPU_FUSE_CASE(2, false)
    ACGPU_FUSE_CASE(0, true) ACGPU_FUSE_CASE(1, true) ACGPU_FUSE_CASE(2, true)
#undef ACGPU_FUSE_CASE
    return cudaErrorInvalidValue;
}

cudaError_t ACGPU_CAT(duo_launch_, TIER_K)(int low, bool is_map, const DevAutomaton &A, const DevTier &T, const MaskArgs &P, const DuoArgs &D, int grid,
                                           size_t smem, cudaStream_t st) {
    if (TIER_K == 1 || low < 0 || low > 2) low = 2;
#define ACGPU_DUO_CASE(L, M) \
    if (low == L && is_map == M) return launch_duo_variant<L, M>(A, T, P, D, grid, smem, st);
    ACGPU_DUO_CASE(0, false) ACGPU_DUO_CASE(1, false) ACGPU_DUO_CASE(2, false)
    ACGPU_DUO_CASE(0, true) ACGPU_DUO_CASE(1, true) ACGPU_DUO_CASE(2, true)
#undef ACGPU_DUO_CASE
    return cudaErrorInvalidValue;
}

}  // namespace acgpu
