// K2: kNN + ball-query grouping with the relative-xyz gather, one warp per (cloud, centre).
//
// Replaces group_points_3DV and its copies (reference training_code/utils_my.py:255-291, :7-42, :217-253,
// :293-328).  The reference materialises (M,S,3,N) difference tensors, runs a full-row torch.topk, then a
// Python loop of S masked index writes.  Here the cloud tile is staged once into shared memory by a 1-D bulk
// TMA copy, distances live in registers only, and the K nearest are selected exactly:
//
//   pass 1  every lane tracks its T = ceil(K/32) + 1 smallest distances; tau = the K-th smallest of those 32*T values
//           (bitwise binary search with warp-wide counts) bounds the K-th smallest distance from above, tightly enough
//           that only ~1.2 K points pass it;
//   pass 2  points with d <= tau are compacted (ballot) into a per-warp candidate list of 64-bit keys
//           (distance bits << 32 | index) -- distances are >= 0, so key order == (distance, index) order;
//   rank    a candidate's output slot is the number of smaller keys; slots < K are written.
//
// Keys are unique, so the chosen SET is exactly "the K smallest by (distance, index)" -- the set
// torch.topk(largest=False) returns whenever the K-th distance is not tied.  Slots whose
// squared distance is > r2 (strict, utils_my.py:272) are redirected to the centre itself (:274-275).
// Distances use ((dx*dx + dy*dy) + dz*dz) without FMA contraction -> bit-identical to the fp32 reference.
#include <type_traits>

#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"

namespace facl {

namespace {

constexpr int GW = 8;             // warps (= centres) per block
constexpr int CAP = 256;          // candidate slots per warp
constexpr int TILE_PTS = 4096;    // points per shared-memory tile (at most 64 KB as float4)
constexpr int TMAX = 5;           // supports K <= 128 (T = ceil(K/32) + 1 tracked per lane, a template parameter)

// keep t[0] <= t[1] <= ... <= t[T-1] = the T smallest seen: a branch-free insertion network (2 min/max per level)
template <int T>
__device__ __forceinline__ void insert_smallest(float (&t)[T], float d) {
    float carry = d;
#pragma unroll
    for (int q = 0; q < T; ++q) {
        const float lo = fminf(t[q], carry);
        carry = fmaxf(t[q], carry);
        t[q] = lo;
    }
}

// NPL > 0: the cloud is exactly NPL * 32 points (one resident tile); every lane keeps its NPL distances in registers, so
// the second pass is a compare + ballot per point instead of a reload and a recomputation.  NPL == 0: any N.
template <bool D4, int T, int NPL>
__global__ void __launch_bounds__(GW * 32) group_kernel(const float* __restrict__ points, int N, int D, int S, int K, float r2,
                                                        float* __restrict__ xt, int* __restrict__ idx_out) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tile = reinterpret_cast<float4*>(smem_raw);                                  // TILE_PTS x 16 B
    unsigned long long* cand_all = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)(N < TILE_PTS ? N : TILE_PTS) * 16);
    __shared__ __align__(8) uint64_t bar;

    const int m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * GW + warp;
    const bool active = s < S;
    const float* cloud = points + (long long)m * N * D;
    unsigned long long* cand = cand_all + warp * CAP;
    const int ntiles = (N + TILE_PTS - 1) / TILE_PTS;
    const int tile_pts = N < TILE_PTS ? N : TILE_PTS;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t tma_phase = 0;

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        cx = __ldg(cloud + (long long)s * D + 0);
        cy = __ldg(cloud + (long long)s * D + 1);
        cz = __ldg(cloud + (long long)s * D + 2);
    }

    auto load_tile = [&](int t) {
        const int n0 = t * TILE_PTS;
        const int cnt = min(TILE_PTS, N - n0);
        if (D4) {
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(&bar, (uint32_t)cnt * 16u);
                tma_bulk_g2s(tile, cloud + (long long)n0 * 4, (uint32_t)cnt * 16u, &bar);
            }
            mbar_wait(&bar, tma_phase);
            tma_phase ^= 1;
        } else {
            for (int i = threadIdx.x; i < cnt; i += GW * 32) {
                const float* p = cloud + (long long)(n0 + i) * D;
                tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            }
            __syncthreads();
        }
        return cnt;
    };

    // ---------------- pass 1: per-lane T smallest -> tau ----------------
    float tsm[T];
#pragma unroll
    for (int q = 0; q < T; ++q) tsm[q] = INFINITY;
    float dreg[NPL > 0 ? NPL : 1];
    if constexpr (NPL > 0) {
        load_tile(0);
        if (active) {
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                float4 p = tile[q * 32 + lane];
                dreg[q] = sqdist_ref(p.x, p.y, p.z, cx, cy, cz);
                insert_smallest<T>(tsm, dreg[q]);
            }
        }
    } else {
        for (int t = 0; t < ntiles; ++t) {
            if (t > 0) __syncthreads();           // everyone is done with the previous tile
            const int cnt = load_tile(t);
            if (active) {
                for (int i = lane; i < cnt; i += 32) {
                    float4 p = tile[i];
                    insert_smallest<T>(tsm, sqdist_ref(p.x, p.y, p.z, cx, cy, cz));
                }
            }
        }
    }
    // tau = K-th smallest of the 32*T tracked values (+inf where a lane saw fewer than T points): distances are >= 0, so
    // their bit patterns order like unsigned integers; build the answer bit by bit from warp-wide counts.
    unsigned tb[T];
#pragma unroll
    for (int q = 0; q < T; ++q) tb[q] = __float_as_uint(tsm[q]);
    unsigned prefix = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned trial = prefix | (1u << bit);
        int c = 0;
#pragma unroll
        for (int q = 0; q < T; ++q) c += (tb[q] < trial) ? 1 : 0;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (c < K) prefix = trial;            // fewer than K values below `trial`: the K-th smallest is >= trial
    }
    const float tau = __uint_as_float(prefix);
    (void)tile_pts;

    // ---------------- pass 2: compact candidates with d <= tau ----------------
    int count = 0;
    if constexpr (NPL > 0) {
        if (active) {
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const float d = dreg[q];
                const bool take = d <= tau;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, take);
                if (bal == 0u) continue;
                if (take) {
                    int pos = count + __popc(bal & ((1u << lane) - 1u));
                    if (pos < CAP) cand[pos] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(q * 32 + lane);
                }
                count += __popc(bal);
            }
        }
    } else {
        for (int t = 0; t < ntiles; ++t) {
            int cnt;
            if (ntiles > 1) {
                __syncthreads();
                cnt = load_tile(t);
            } else {
                cnt = N;                           // single tile: still resident
            }
            if (active) {
                const int n0 = t * TILE_PTS;
                for (int i0 = 0; i0 < cnt; i0 += 32) {
                    int i = i0 + lane;
                    bool take = false;
                    float d = 0.f;
                    if (i < cnt) {
                        float4 p = tile[i];
                        d = sqdist_ref(p.x, p.y, p.z, cx, cy, cz);
                        take = d <= tau;
                    }
                    unsigned bal = __ballot_sync(0xFFFFFFFFu, take);
                    if (bal == 0u) continue;               // ~96 % of the 32-point rows hold no candidate
                    if (take) {
                        int pos = count + __popc(bal & ((1u << lane) - 1u));
                        if (pos < CAP) cand[pos] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(n0 + i);
                    }
                    count += __popc(bal);
                }
            }
        }
    }
    if (!active) return;
    __syncwarp();

    float* xrow = xt ? xt + ((long long)m * S + s) * K * D : nullptr;      // nullptr: indices only (level-2 grouping)
    int* irow = idx_out ? idx_out + ((long long)m * S + s) * K : nullptr;
    const unsigned r2bits = __float_as_uint(r2);

    auto emit = [&](int slot, unsigned long long key) {
        unsigned dbits = (unsigned)(key >> 32);
        int n = (int)(unsigned)(key & 0xFFFFFFFFu);
        if (dbits > r2bits) n = s;             // d > r2 (both >= 0): redirect to the centre itself
        if (irow) irow[slot] = n;
        if (!xrow) return;
        const float* p = cloud + (long long)n * D;
        float* o = xrow + (long long)slot * D;
        if (D4) {
            float4 q = __ldg(reinterpret_cast<const float4*>(p));
            *reinterpret_cast<float4*>(o) = make_float4(__fsub_rn(q.x, cx), __fsub_rn(q.y, cy), __fsub_rn(q.z, cz), q.w);
        } else {
            o[0] = __fsub_rn(__ldg(p), cx);
            o[1] = __fsub_rn(__ldg(p + 1), cy);
            o[2] = __fsub_rn(__ldg(p + 2), cz);
            for (int d = 3; d < D; ++d) o[d] = __ldg(p + d);
        }
    };

    if (count <= CAP) {
        // rank = number of smaller keys; four of this lane's keys stay in registers while one broadcast read per
        // candidate serves all four comparisons
        // (the usual ~1.2 K = 77 candidates need three keys per lane, not four: the loop body is instantiated per key count)
        auto rank_block = [&](auto nk_tag, int base) {
            constexpr int NK = decltype(nk_tag)::value;
            unsigned long long k[NK];
            int r[NK];
#pragma unroll
            for (int q = 0; q < NK; ++q) {
                const int i = base + lane + 32 * q;
                k[q] = (i < count) ? cand[i] : ~0ull;
                r[q] = 0;
            }
#pragma unroll 4
            for (int j = 0; j < count; ++j) {
                const unsigned long long c = cand[j];
#pragma unroll
                for (int q = 0; q < NK; ++q) r[q] += (c < k[q]) ? 1 : 0;
            }
#pragma unroll
            for (int q = 0; q < NK; ++q)
                if (base + lane + 32 * q < count && r[q] < K) emit(r[q], k[q]);
        };
        for (int base = 0; base < count; base += 128) {
            const int left = count - base;
            if (left <= 64) rank_block(std::integral_constant<int, 2>{}, base);
            else if (left <= 96) rank_block(std::integral_constant<int, 3>{}, base);
            else rank_block(std::integral_constant<int, 4>{}, base);
        }
    } else {
        // Rare (heavy duplication around the centre): K rounds of "smallest key greater than the last one",
        // straight from global memory.
        unsigned long long last = 0ull;
        bool first = true;
        for (int slot = 0; slot < K; ++slot) {
            unsigned long long best = ~0ull;
            for (int i = lane; i < N; i += 32) {
                const float* p = cloud + (long long)i * D;
                float d = sqdist_ref(__ldg(p), __ldg(p + 1), __ldg(p + 2), cx, cy, cz);
                unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
                if ((first || key > last) && key < best) best = key;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
                best = other < best ? other : best;
            }
            if (lane == 0) emit(slot, best);
            last = best;
            first = false;
        }
    }
}

}  // namespace

int group_launch(const float* points, int M, int N, int D, int S, int K, float r2, float* xt, int* idx_out, cudaStream_t st) {
    if (M <= 0 || N <= 0 || D < 3 || S <= 0 || S > N || K <= 0 || K > N || K > 32 * TMAX) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)(N < TILE_PTS ? N : TILE_PTS) * 16 + (size_t)GW * CAP * 8;
    static DeviceOnce configured;
    if (configured.need()) {
        const int max_smem = TILE_PTS * 16 + GW * CAP * 8;
#define FACL_GROUP_ATTR(TT)                                                                                              \
        FACL_CHECK(cudaFuncSetAttribute(group_kernel<true, TT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));   \
        FACL_CHECK(cudaFuncSetAttribute(group_kernel<false, TT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FACL_GROUP_ATTR(2) FACL_GROUP_ATTR(3) FACL_GROUP_ATTR(4) FACL_GROUP_ATTR(5)
#undef FACL_GROUP_ATTR
        FACL_CHECK(cudaFuncSetAttribute(group_kernel<true, 3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FACL_CHECK(cudaFuncSetAttribute(group_kernel<true, 3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured.done();
    }
    dim3 grid((S + GW - 1) / GW, M);
    ScopedTimer timer(TAG_GROUP, st);
    count_launch();
    const bool d4 = (D == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const int T = (K + 31) / 32 + 1;           // smallest distances tracked per lane in pass 1
#define FACL_GROUP_LAUNCH(TT)                                                                                     \
    do {                                                                                                          \
        if (d4) FACL_LAUNCH_OK(launch_pdl(group_kernel<true, TT, 0>, dim3(grid), dim3(GW * 32), smem, st, points, N, D, S, K, r2, xt, idx_out));      \
        else FACL_LAUNCH_OK(launch_pdl(group_kernel<false, TT, 0>, dim3(grid), dim3(GW * 32), smem, st, points, N, D, S, K, r2, xt, idx_out));        \
    } while (0)
    // the training shapes (K = 64; N = 2048 or 1024): distances stay in registers between the two passes
    if (d4 && T == 3 && (N == 2048 || N == 1024)) {
        if (N == 2048) FACL_LAUNCH_OK(launch_pdl(group_kernel<true, 3, 64>, dim3(grid), dim3(GW * 32), smem, st, points, N, D, S, K, r2, xt, idx_out));
        else FACL_LAUNCH_OK(launch_pdl(group_kernel<true, 3, 32>, dim3(grid), dim3(GW * 32), smem, st, points, N, D, S, K, r2, xt, idx_out));
        return (int)cudaGetLastError();
    }
    switch (T) {
        case 2: FACL_GROUP_LAUNCH(2); break;
        case 3: FACL_GROUP_LAUNCH(3); break;
        case 4: FACL_GROUP_LAUNCH(4); break;
        default: FACL_GROUP_LAUNCH(5); break;
    }
#undef FACL_GROUP_LAUNCH
    return (int)cudaGetLastError();
}

}  // namespace facl
