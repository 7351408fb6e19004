// k_tier_fused<K, LOW, isMap>: the AhoCorasick path for narrow alphabets in ONE persistent launch (generation 5).
//
// The three-launch path (kernel_mask.cuh) writes 2 B of hit masks per position to HBM and reads them back: 4 of the 13 GB
// of DRAM traffic per 10^9 chars of configs[4], and k_tier_emit is DRAM-bound (94 % of the measured copy rate) while the
// mask kernel is bound by the L1 miss path (one gathered sector per pair of positions).  Here both run in the same
// kernel and the masks only ever live in L2:
//
//   * a TICKET is a run of `ticket_rows` consecutive rows.  Every warp alternates between two roles, decided per ticket
//     from two counters (next ticket to make, next ticket to expand): MAKE the hit masks of the
//     next ticket (the body of k_tier_mask, into a ring of `ring_tickets` slots), or - when the makers are `high` tickets
//     ahead, or have nothing left - EXPAND the oldest ticket into records (emit_row of kernel_emit.cuh);
//   * the reference's output order is position-major (AhoCorasickSet.java:522-535), so a ticket's records start at the
//     sum of the counts of all earlier tickets: a maker publishes its ticket's count, ONE warp of the grid (CTA 0, warp 0)
//     scans the counts in order, 512 at a time, and publishes exclusive prefixes; expanders run `high` tickets behind
//     the makers, so their prefix is (almost) always there already - no ordered waiting on the hot path;
//   * ring slots are reused: a maker waits until the ticket that used the slot `ring_tickets` earlier has been expanded.
//
// Progress: the grid is persistent (one CTA per SM, all resident).  A ticket is only handed to an expander after it was
// handed to a maker, makers wait only for expanders of far older tickets, the scanner only for makers.
#pragma once
#include "kernel_emit.cuh"

namespace acgpu {

constexpr uint32_t kFuseReady = 0x80000000u;
constexpr unsigned long long kFuseKnown = 1ull << 63;
constexpr uint32_t kFuseScanStep = 512;  // tickets per step of the scanning warp

struct FuseArgs {
    EmitArgs E;                  // hay, n, origin, pos_base, pos_out, val_out, cap; n_rows = rows of the run (masks / row_excl / block_excl unused)
    uint32_t *agg;               // [n_tickets rounded up to kFuseScanStep] records of a ticket | kFuseReady once its masks are final
    unsigned long long *excl;    // [same] records of all earlier tickets | kFuseKnown
    uint32_t *expanded;          // [n_tickets] 1 once the ticket's masks have been read (its ring slot is free)
    unsigned long long *ab;      // hi: next ticket to make, lo: next ticket to expand
    unsigned long long *total_out;
    uint32_t n_tickets, ticket_rows, ring_tickets, high;  // ring_tickets: a power of two
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ uint4 ld_volatile_v4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <bool kIsMap>
struct FuseSched {
    static constexpr bool kFused = true;
    const DevAutomaton &A;
    const DevTier &T;
    const FuseArgs &F;
    const MaskArgs &P;
    EmitWarp W;
    bool scanner;
    uint32_t pending;  // an expand ticket claimed before it was handed to a maker

    __device__ __forceinline__ int rows() const { return (int)F.ticket_rows; }
    __device__ __forceinline__ int64_t mask_row0(uint32_t chunk, int64_t) const { return (int64_t)(chunk & (F.ring_tickets - 1u)) * F.ticket_rows; }

    // a maker is done with ticket t: publish its count
    __device__ __forceinline__ void finish(uint32_t t, uint32_t total, int lane) {
        __syncwarp();
        if (lane == 0) {
            __threadfence();  // the warp's mask stores and atomics come before the count
            *reinterpret_cast<volatile uint32_t *>(F.agg + t) = total | kFuseReady;
        }
    }

    // counts -> exclusive prefixes, in ticket order, kFuseScanStep tickets per step (one warp of the whole grid): a lane
    // owns 16 consecutive tickets (four 16-byte loads in flight; the next step's are issued before this step's stores)
    __device__ __noinline__ void scan_all(int lane) {
        unsigned long long running = 0;
        uint4 v[4];
        auto load = [&](uint32_t t) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = ld_volatile_v4(F.agg + t + 4 * k);
        };
        load((uint32_t)lane * 16u);
        for (uint32_t t0 = 0; t0 < F.n_tickets; t0 += kFuseScanStep) {
            const uint32_t t = t0 + (uint32_t)lane * 16u;
            uint32_t c[16];
            while (true) {
#pragma unroll
                for (int k = 0; k < 4; k++) { c[4 * k] = v[k].x; c[4 * k + 1] = v[k].y; c[4 * k + 2] = v[k].z; c[4 * k + 3] = v[k].w; }
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 16; k++) ok = ok && (t + k >= F.n_tickets || (c[k] & kFuseReady));
                if (__all_sync(0xFFFFFFFFu, ok)) break;
                __nanosleep(100);
                load(t);
            }
            if (t0 + kFuseScanStep < F.n_tickets) load(t + kFuseScanStep);  // the array is padded to a whole step
            uint32_t mine = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                c[k] = t + k < F.n_tickets ? c[k] & ~kFuseReady : 0u;
                mine += c[k];
            }
            uint32_t inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            unsigned long long e = running + (inc - mine);
            __threadfence();
            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(F.excl + t);
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + (k >> 1)), "l"(e | kFuseKnown), "l"((e + c[k]) | kFuseKnown) : "memory");
                e += c[k] + c[k + 1];
            }
            running += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) *F.total_out = running;
    }

    // expand ticket t into records
    __device__ __forceinline__ void expand(uint32_t t, int lane) {
        unsigned long long base = 0;
        if (lane == 0) {
            while (!((base = ld_volatile_u64(F.excl + t)) & kFuseKnown)) __nanosleep(200);
            base &= ~kFuseKnown;
            __threadfence();
        }
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const int row0 = (int)(t * F.ticket_rows);
        const int n_cr = min((int)F.ticket_rows, (int)F.E.n_rows - row0);
        const uint4 *mp = reinterpret_cast<const uint4 *>(P.masks) + ((size_t)(t & (F.ring_tickets - 1u)) * F.ticket_rows * 32 + lane);
        uint4 mm_n = __ldcg(mp);
        for (int r = 0; r < n_cr; ++r) {
            const uint4 mm = mm_n;
            if (r + 1 < n_cr) mm_n = __ldcg(mp + (size_t)(r + 1) * 32);
            base += emit_row<kIsMap>(A, T, F.E, W, row0 + r, mm, base);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            *reinterpret_cast<volatile uint32_t *>(F.expanded + t) = 1u;
        }
    }

    // the warp's next ticket to MAKE (0xFFFFFFFF: nothing left); tickets to expand are served on the way.  Both counters
    // move by fetch-and-add (a compare-and-swap on the pair convoys: one success per L2 round trip), so a burst of
    // claims can run past the other counter by at most the number of warps: an expander whose ticket has not been handed
    // to a maker yet keeps it PENDING and makes tickets until it has been; a maker's ring slot wait covers the rest.
    __device__ __forceinline__ uint32_t next(int lane) {
        if (scanner) {
            scan_all(lane);
            return 0xFFFFFFFFu;
        }
        uint32_t *pa = reinterpret_cast<uint32_t *>(F.ab) + 1, *pb = reinterpret_cast<uint32_t *>(F.ab);
        const uint32_t n = F.n_tickets;
        while (true) {
            uint32_t role = 4u, t = 0u;  // 0 make t, 1 expand t, 2 nothing left, 3 keep t pending, 4 look again
            if (lane == 0) {
                const uint32_t a = ld_volatile_u32(pa), b = ld_volatile_u32(pb);
                bool make = false;
                if (pending != 0xFFFFFFFFu) {
                    if (a > pending) {
                        role = 1u;
                        t = pending;
                    } else {
                        make = true;
                    }
                } else if (b < n && b < a && (a >= n || a - b >= F.high)) {
                    t = atomicAdd(pb, 1u);
                    if (t < n) role = ld_volatile_u32(pa) > t ? 1u : 3u;
                } else if (a < n) {
                    make = true;
                } else if (b >= n) {
                    role = 2u;
                }
                if (make) {
                    const uint32_t mt = atomicAdd(pa, 1u);
                    if (mt < n) {
                        if (mt >= F.ring_tickets) {  // the ring slot must have been read by its previous ticket's expander
                            while (!ld_volatile_u32(F.expanded + (mt - F.ring_tickets))) __nanosleep(200);
                            __threadfence();
                        }
                        role = 0u;
                        t = mt;
                    }
                }
            }
            role = __shfl_sync(0xFFFFFFFFu, role, 0);
            t = __shfl_sync(0xFFFFFFFFu, t, 0);
            if (role == 0u) return t;
            if (role == 1u) {
                expand(t, lane);
                pending = 0xFFFFFFFFu;
            } else if (role == 3u) {
                pending = t;
            } else if (role == 2u) {
                return 0xFFFFFFFFu;
            } else {
                __nanosleep(100);
            }
        }
    }
};

// shared memory: [k_tier_mask's block: classes, level tables, probe queues][per warp: the emit role's windows]
__host__ __device__ constexpr size_t fuse_smem_bytes(size_t n_row_words, bool is_map) {
    return mask_smem_bytes(n_row_words) + (size_t)kMaskWarps * (is_map ? kEmitWarpBytesMap : kEmitWarpBytesSet);
}

template <int K, int LOW, bool kIsMap>
__global__ void __launch_bounds__(kMaskThreads, 1) k_tier_fused(const DevAutomaton A, const DevTier T, const MaskArgs P, const FuseArgs F) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *s_emit = reinterpret_cast<unsigned char *>(s_mem) + mask_smem_bytes(T.n_row_words) +
                            (size_t)warp * (kIsMap ? kEmitWarpBytesMap : kEmitWarpBytesSet);
    int2 *s_stage = reinterpret_cast<int2 *>(s_emit);
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_emit + kEmitWarpBytesSet);
    uint2 *s_pack = reinterpret_cast<uint2 *>(s_emit + kEmitWarpBytesSet + 4 * ((kEmitWin + 3) & ~3));
    FuseSched<kIsMap> S{A, T, F, P, make_emit_warp(T, F.E, s_stage, s_val, s_pack, reinterpret_cast<const uint8_t *>(s_mem), lane),
                        blockIdx.x == 0 && warp == 0, 0xFFFFFFFFu};
    tier_mask_body<K, LOW, false>(A, T, P, S, s_mem);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_tier_duo<K, LOW, isMap> (generation 6): the masks of one SLAB of the haystack and the records of the slab BEFORE it in
// the same launch.
//
// k_tier_mask is bound by the L1 miss path (gathers, DRAM at 22 %), k_tier_emit by DRAM (94 % of the copy rate): run one
// after the other each leaves the other's unit idle.  Running them as two kernels on two streams does not overlap them (the
// persistent mask CTAs fill the register file), generation 5 interleaved them per ticket inside one kernel and lost to its own
// ordering machinery and to 214 KB of shared memory.  Here the haystack is cut into a few slabs (multiples of 4 096 rows) and
// launch i makes the masks of slab i while `emit_warps` warps of every CTA expand slab i - 1, whose masks and row offsets
// (k_row_scan with the running total of the slabs before it) are final since the launches before: no ordering inside the
// kernel, masks through HBM as in generation 3.  The emit warps take tickets of 8 rows until slab i - 1 is done, then mask
// tickets like everyone else; only they have staging windows (8 x 0.9 KB: the CTA stays on the 196 KB carve-out).  The last
// slab is expanded by k_tier_emit.
constexpr int kDuoEmitRows = 8;

struct DuoArgs {
    EmitArgs E;                 // the slab to expand (n_rows = 0: none)
    unsigned int *emit_ticket;
    int emit_warps;             // warps 0 .. emit_warps - 1 of every CTA expand first
};

template <bool kIsMap>
struct DuoSched {
    static constexpr bool kFused = false;
    const DevAutomaton &A;
    const DevTier &T;
    const DuoArgs &D;
    unsigned int *mask_ticket;
    EmitWarp W;
    bool emit_pending;

    __device__ __forceinline__ int rows() const { return kMaskChunkRows; }
    __device__ __forceinline__ int64_t mask_row0(uint32_t, int64_t row0) const { return row0; }
    __device__ __forceinline__ void finish(uint32_t, uint32_t, int) {}

    __device__ __noinline__ void expand_all(int lane) {
        const EmitArgs &E = D.E;
        const int n_rows = (int)E.n_rows;
        while (true) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(D.emit_ticket, 1u);
            t = __shfl_sync(0xFFFFFFFFu, t, 0);
            const int row0 = (int)t * kDuoEmitRows;
            if (row0 >= n_rows) break;
            const int n_r = min(kDuoEmitRows, n_rows - row0);
            const uint4 *mp = reinterpret_cast<const uint4 *>(E.masks) + ((size_t)row0 * 32 + lane);
            const unsigned long long block_base = __ldg(E.block_excl + (row0 >> 12));   // kDuoEmitRows divides 4 096: one scan block per ticket
            uint4 mm_n = __ldcs(mp);
            uint32_t off_n = __ldg(E.row_excl + row0);
            for (int r = 0; r < n_r; ++r) {
                const uint4 mm = mm_n;
                const unsigned long long base = block_base + off_n;
                if (r + 1 < n_r) {
                    mm_n = __ldcs(mp + (size_t)(r + 1) * 32);
                    off_n = __ldg(E.row_excl + row0 + r + 1);
                }
                emit_row<kIsMap>(A, T, E, W, row0 + r, mm, base);
            }
        }
    }

    __device__ __forceinline__ uint32_t next(int lane) {
        if (emit_pending) {
            expand_all(lane);
            emit_pending = false;
        }
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(mask_ticket, 1u);
        return __shfl_sync(0xFFFFFFFFu, chunk, 0);
    }
};

__host__ __device__ constexpr size_t duo_smem_bytes(size_t n_row_words, bool is_map, int emit_warps) {
    return mask_smem_bytes(n_row_words) + (size_t)emit_warps * (is_map ? kEmitWarpBytesMap : kEmitWarpBytesSet);
}

template <int K, int LOW, bool kIsMap>
__global__ void __launch_bounds__(kMaskThreads, 1) k_tier_duo(const DevAutomaton A, const DevTier T, const MaskArgs P, const DuoArgs D) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool emitter = warp < D.emit_warps;
    unsigned char *s_emit = reinterpret_cast<unsigned char *>(s_mem) + mask_smem_bytes(T.n_row_words) +
                            (size_t)(emitter ? warp : 0) * (kIsMap ? kEmitWarpBytesMap : kEmitWarpBytesSet);
    int2 *s_stage = reinterpret_cast<int2 *>(s_emit);
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_emit + kEmitWarpBytesSet);
    uint2 *s_pack = reinterpret_cast<uint2 *>(s_emit + kEmitWarpBytesSet + 4 * ((kEmitWin + 3) & ~3));
    DuoSched<kIsMap> S{A, T, D, P.ticket, make_emit_warp(T, D.E, s_stage, s_val, s_pack, reinterpret_cast<const uint8_t *>(s_mem), lane),
                       emitter && D.E.n_rows > 0};
    tier_mask_body<K, LOW, false>(A, T, P, S, s_mem);
    static_assert(kScanRows % kDuoEmitRows == 0, "an emit ticket lies inside one scan block");
}

}  // namespace acgpu
