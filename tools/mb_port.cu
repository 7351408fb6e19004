// Microbenchmark behind DESIGN.md §4.1 (round 2): what bounds k_tier_mask is the SM's L1TEX -> XBAR request port
// (ncu: l1tex__m_l1tex2xbar_req_cycles_active 82 %, one 32-byte sector request per cycle per SM).  Questions:
//   1. random 4-byte gathers from an L2-resident table: sectors per cycle per SM through TEX and through LSU;
//   2. how much L1 (228 KB - shared-memory carve-out) the gathers need;
//   3. does streaming the haystack / masks with cp.async.bulk (TMA unit) instead of LDG.128 / STG.128 relieve the port?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_port tools/mb_port.cu ; run on a B200: tools/mb_port
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

constexpr int kWarps = 32;
constexpr int kRowBytes = 512;   // one warp row: 32 lanes x 16 bytes (256 UTF-16 chars)
constexpr int kChunkRows = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// STREAM: 0 none, 1 LDG.128 + STG.128, 2 cp.async.bulk load + store.  GATHER: 0 none, 1 TEX, 2 LDG.  G gathers per lane per row.
template <int STREAM, int GATHER, int G>
__global__ void __launch_bounds__(kWarps * 32, 1)
k_port(const uint32_t *table, cudaTextureObject_t tex, uint32_t tmask, const uint4 *in, uint4 *out, int64_t n_rows, unsigned *ticket,
       uint32_t *sink) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp: in[2][512], out[2][512], mbar[2]
    unsigned char *base = s_raw + (size_t)warp * (4 * kRowBytes + 64);
    const uint32_t s_in = smem_u32(base), s_out = smem_u32(base + 2 * kRowBytes), s_bar = smem_u32(base + 4 * kRowBytes);
    if (STREAM == 2 && lane == 0) {
        mbar_init(s_bar, 1);
        mbar_init(s_bar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t acc = 0, phase0 = 0, phase1 = 0;
    while (true) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xFFFFFFFFu, c, 0);
        const int64_t row0 = (int64_t)c * kChunkRows;
        if (row0 >= n_rows) break;
        const int n_cr = (int)min((int64_t)kChunkRows, n_rows - row0);
        const uint4 *ip = in + row0 * 32 + lane;
        uint4 *op = out + row0 * 32 + lane;
        uint4 v = make_uint4(0, 0, 0, 0), vn = v;
        if (STREAM == 1) {
            v = __ldcs(ip);
            if (n_cr > 1) vn = __ldcs(ip + 32);
        }
        if (STREAM == 2 && lane == 0) {
            mbar_expect(s_bar, kRowBytes);
            bulk_g2s(s_in, in + row0 * 32, kRowBytes, s_bar);
            if (n_cr > 1) {
                mbar_expect(s_bar + 8, kRowBytes);
                bulk_g2s(s_in + kRowBytes, in + (row0 + 1) * 32, kRowBytes, s_bar + 8);
            }
        }
        for (int r = 0; r < n_cr; ++r) {
            uint4 vnn = make_uint4(0, 0, 0, 0);
            if (STREAM == 1 && r + 2 < n_cr) vnn = __ldcs(ip + (r + 2) * 32);
            if (STREAM == 2) {
                const int bf = r & 1;
                mbar_wait(s_bar + 8 * bf, bf ? phase1 : phase0);
                if (bf) phase1 ^= 1u; else phase0 ^= 1u;
                v = *reinterpret_cast<const uint4 *>(base + bf * kRowBytes + lane * 16);
                __syncwarp();
                if (lane == 0 && r + 2 < n_cr) {
                    mbar_expect(s_bar + 8 * bf, kRowBytes);
                    bulk_g2s(s_in + bf * kRowBytes, in + (row0 + r + 2) * 32, kRowBytes, s_bar + 8 * bf);
                }
            }
            if (STREAM == 0) v = make_uint4(c * 977u + r, lane, r * 31u, c);
            uint32_t h = mix(v.x ^ (v.y * 0x9E3779B1u) ^ v.z ^ (v.w << 7) ^ (uint32_t)(row0 + r) ^ ((uint32_t)lane * 0x85EBCA6Bu));
            uint32_t g[G > 0 ? G : 1];
#pragma unroll
            for (int j = 0; j < G; j++) {
                const uint32_t idx = mix(h + j * 0x632BE5ABu) & tmask;
                if (GATHER == 1) g[j] = tex1Dfetch<unsigned int>(tex, (int)idx);
                if (GATHER == 2) g[j] = __ldg(table + idx);
            }
#pragma unroll
            for (int j = 0; j < G; j++) acc ^= g[j] + j;
            const uint4 w = make_uint4(v.x ^ acc, v.y, v.z + r, v.w);
            if (STREAM == 1) {
                op[r * 32] = w;
                v = vn;
                vn = vnn;
            }
            if (STREAM == 2) {
                const int bf = r & 1;
                if (r >= 2) {  // the bulk store that read this buffer two rows ago must have finished reading
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                *reinterpret_cast<uint4 *>(base + 2 * kRowBytes + bf * kRowBytes + lane * 16) = w;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) bulk_s2g(out + (row0 + r) * 32, s_out + bf * kRowBytes, kRowBytes);
            }
        }
        if (STREAM == 2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
    }
    if (STREAM == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int STREAM, int GATHER, int G>
void run(const char *name, const uint32_t *table, cudaTextureObject_t tex, uint32_t tmask, const uint4 *in, uint4 *out, int64_t n_rows,
         unsigned *ticket, uint32_t *sink, size_t smem, int sms) {
    CK(cudaFuncSetAttribute(k_port<STREAM, GATHER, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(ticket, 0, 4));
        CK(cudaEventRecord(e0));
        k_port<STREAM, GATHER, G><<<sms, kWarps * 32, smem>>>(table, tex, tmask, in, out, n_rows, ticket, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    const double gathers = (double)n_rows * 32 * G, stream_sectors = STREAM ? (double)n_rows * 32 : 0.0;  // 16 in + 16 out per row
    const double cyc = best * 1e-3 * 1.965e9;
    std::printf("%-34s smem %3zu KB  %7.3f ms  gathers %6.1f M  stream sectors %6.1f M  gather sectors/cycle/SM %.3f  all sectors/cycle/SM %.3f\n", name,
                smem / 1024, best, gathers / 1e6, stream_sectors / 1e6, gathers / sms / cyc, (gathers + stream_sectors) / sms / cyc);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t n_tab = 1u << 19;  // 2 MB of words, like kidmask[27^4]
    uint32_t *table;
    CK(cudaMalloc(&table, n_tab * 4));
    CK(cudaMemset(table, 0x5A, n_tab * 4));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>();
    rd.res.linear.sizeInBytes = n_tab * 4;
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const int64_t n_rows = 1 << 21;  // 2^21 rows x 512 B = 1 GB in, 1 GB out (5.4e8 "chars")
    uint4 *in, *out;
    CK(cudaMalloc(&in, n_rows * kRowBytes));
    CK(cudaMalloc(&out, n_rows * kRowBytes));
    CK(cudaMemset(in, 0x11, n_rows * kRowBytes));
    unsigned *ticket;
    uint32_t *sink;
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMalloc(&sink, 4));
    const uint32_t tmask = n_tab - 1;
    const size_t need = (size_t)kWarps * (4 * kRowBytes + 64);
    std::printf("SMs %d, table 2 MB, %ld rows of 512 B\n", sms, (long)n_rows);
    for (size_t kb : {68, 160, 170, 178, 185, 192, 200}) {
        const size_t smem = kb * 1024;
        if (smem < need) continue;
        run<0, 1, 4>("gather TEX x4", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<0, 2, 4>("gather LDG x4", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
    }
    for (size_t kb : {185, 217}) {
        const size_t smem = kb * 1024;
        run<1, 0, 0>("stream LDG/STG only", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<2, 0, 0>("stream bulk only", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<1, 1, 4>("TEX x4 + stream LDG/STG", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<2, 1, 4>("TEX x4 + stream bulk", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<1, 1, 8>("TEX x8 + stream LDG/STG", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
        run<2, 1, 8>("TEX x8 + stream bulk", table, tex, tmask, in, out, n_rows, ticket, sink, smem, sms);
    }
    return 0;
}
