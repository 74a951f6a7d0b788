// K3: the first shared-MLP stack (net3DV_1: 4 -> 64 -> 64 -> 256, BN + ReLU after each, max over the K neighbours)
// as fused tcgen05 kernels that never write a per-row activation to HBM.
//
// Replaces nn.Sequential net3DV_1 of reference training_code/cn3d_model_conbag.py:162-177 (= :43-58).
//
// Everything is laid out "channel on the TMEM lane / batch row on the TMEM column": accumulators are D[c][r], so an
// epilogue thread owns ONE channel and sees its rows in registers -- BatchNorm statistics and the neighbourhood
// max-pool are in-thread reductions, BN+ReLU constants are per-thread scalars, and a thread that has 8 consecutive rows
// of its channel packs them into one 16-byte shared-memory store.  Activation tiles are therefore images of
// [channel][64 consecutive rows] (128-byte rows, 128B swizzle), fed to the next tcgen05.mma as an MN-major B operand.
//
// Train-mode BatchNorm needs full-batch statistics before it can be applied, so the forward is three launches:
//   l1_moments : sum x, sum x x^T over all rows  -> BN1 statistics in closed form (z1 is affine in the 4 inputs)
//   pass A     : x -> h1 -> z2 (tensor core)     -> BN2 statistics
//   pass B     : x -> h1 -> z2 -> h2 -> z3 (tensor core) -> BN3 statistics + max over K of the pre-BN value
// (max_k relu(a z_k + b) = relu(a * (a >= 0 ? max z : min z) + b), so pooling can precede BN3.)
#include <stdio.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"

namespace facl {

namespace {

// -DFACL_PROFILE_ROLES: clock64 accounting of where every warp role of the backward passes spends its tile period (waits on
// each mbarrier vs work), printed by one warp per role of CTA 1.  Diagnostic builds only (tools/role_profile.sh).
#ifdef FACL_PROFILE_ROLES
#define PROF_DECL(n) long long prof_[n]; for (int pi_ = 0; pi_ < n; ++pi_) prof_[pi_] = 0; long long prof_t_ = clock64();
#define PROF_MARK(i) { const long long pt_ = clock64(); prof_[i] += pt_ - prof_t_; prof_t_ = pt_; }
#else
#define PROF_DECL(n)
#define PROF_MARK(i)
#endif

constexpr int TILE = 128;                 // batch rows per tile
constexpr uint32_t ACT_LBO = 8192;        // MN-major activation image: 64-row blocks 8 KB apart,
constexpr uint32_t ACT_SBO = 1024;        //                            8-channel groups 1 KB apart (16 KB per half)
constexpr int ACT_BYTES = 16384;
constexpr int NTHREADS_A = 25 * 32;        // pass A: warps 0-7 (idle z3 roles) and 17-24 are the h1 producers
constexpr int NTHREADS_B = 17 * 32;        // pass B: warps 10,11,14,15 are the h1 producers (the z3 consumers own the issue slots)

struct L1Params {
    const float* xt;          // [R][4]
    long long R;              // rows, multiple of TILE
    int K;                    // neighbours per group (power of two, divides TILE)
    int nhl;                  // 1: bf16, 2: bf16 hi+lo (fp32 mode)
    const float* w1;          // [64][4]
    const float* b1;          // [64]
    const float* scale1;      // BN1 scale / shift (slot 0)
    const float* shift1;
    const uint8_t* w2_img;    // packed image of W2 (1 m-tile, 1 k-block)
    const float* b2;
    const float* scale2;      // BN2 (pass B only)
    const float* shift2;
    const uint8_t* w3_img;    // packed image of W3 (2 m-tiles, 1 k-block)
    const float* b3;
    const float* gamma3;      // sign decides max vs min
    float* stats;             // pass A: [2*grid][64][2], pass B: [grid][256][2]
    float* gram;              // pass B with GRAM: H2 += sum_r h2 h2^T [64][64], s2 += sum_r h2 [64] (atomics onto zeroed buffers)
    float* hsum;
    float* pooled;            // pass B: [256][ldp]  selected pre-BN value per (channel, group)
    unsigned char* pool_arg;  // pass B: [256][ldp]  position of the winner inside its group (first hit), or null
    long long ldp;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// write 8 consecutive rows of one channel into an MN-major activation image (hi / hi+lo)
// (lbo = distance between the 64-row blocks, lo_off = distance from a hi block to its lo block)
__device__ __forceinline__ void store_act8(uint8_t* img, int nhl, int channel, int chunk, const float (&v)[8], uint32_t lbo = ACT_LBO,
                                           uint32_t lo_off = ACT_BYTES) {
    uint32_t off = mn_sw128_offset((uint32_t)channel, (uint32_t)chunk, lbo, ACT_SBO);
    if (nhl == 2) {
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(img + off) = h;
        *reinterpret_cast<uint4*>(img + lo_off + off) = l;
    } else {
        *reinterpret_cast<uint4*>(img + off) = pack_bf16x8(v);
    }
}

// A tiles in tensor memory.  tcgen05.mma on a 64-wide tile spends its time fetching the 4 KB A tile from shared memory (70
// cycles per instruction against a 32-cycle math floor); the 64 x 64 weight matrices of the backward are constant for the whole
// kernel, so they are copied ONCE into TMEM (32 columns per 128 x 64 bf16 tile; rows 64..127 zero) and the instructions that use
// them read only their B tile from shared memory.
// This warp's 32 lanes of one tile: `img` = K-major 128B-swizzled image (64 rows x 64 k) in shared memory, or nullptr for zeros.
__device__ __forceinline__ void tmem_put_a_tile(uint32_t taddr, const uint8_t* img, int row) {
    uint32_t u[32];
    if (img) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 q = *reinterpret_cast<const uint4*>(img + sw128_offset((uint32_t)row, (uint32_t)c));
            u[4 * c] = q.x; u[4 * c + 1] = q.y; u[4 * c + 2] = q.z; u[4 * c + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) u[c] = 0u;
    }
    tmem_st32(taddr, u);
}
// D[tmem] (+)= A (TMEM tile, 64-wide K) * B (MN-major activation image), all bf16 hi/lo combinations.
// b = descriptor of the stage-0 hi image; boff = stage offset, b_lo = distance to the lo image (16-byte units)
__device__ __forceinline__ void mma_tmem_weight_act(uint32_t d_tmem, uint32_t a_hi_t, uint32_t a_lo_t, UDesc b, uint32_t boff, uint32_t b_lo,
                                                    bool split, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        umma_ts(d_tmem, a_hi_t + ks * 8, b, boff + ks * 128, idesc, ks > 0 ? 1u : 0u);
        if (split) {
            umma_ts(d_tmem, a_hi_t + ks * 8, b, boff + b_lo + ks * 128, idesc, 1u);
            umma_ts(d_tmem, a_lo_t + ks * 8, b, boff + ks * 128, idesc, 1u);
        }
    }
}
// D[tmem] (+)= A (K-major weight image, 64-wide K) * B (MN-major activation image), all bf16 hi/lo combinations
__device__ __forceinline__ void mma_weight_act(uint32_t d_tmem, UDesc a, uint32_t aoff, uint32_t a_lo, UDesc b, uint32_t boff, uint32_t b_lo,
                                               bool split, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        umma_ss(d_tmem, a, aoff + ks * 2, b, boff + ks * 128, idesc, ks > 0 ? 1u : 0u);
        if (split) {
            umma_ss(d_tmem, a, aoff + ks * 2, b, boff + b_lo + ks * 128, idesc, 1u);
            umma_ss(d_tmem, a, aoff + a_lo + ks * 2, b, boff + ks * 128, idesc, 1u);
        }
    }
}

// STAT: 0 no statistics (eval), 1 per-channel sum / sum of squares of the layer output in the consumer threads.
// GRAM (pass B, training with a backward to follow): H2 = sum h2 h2^T also accumulates on the tensor core and s2 = sum h2 in the
// h2 producers -- the dense part of dW3 needs exactly these (see the backward below), and here the h2 tiles are already in
// shared memory, so pass C does not have to spend tensor-pipe time on them.  (BN3 statistics are NOT derived from H2: the
// closed form w^T (H2/n - m m^T) w amplifies the rounding of the long fp32 accumulation ~100x, measured 1e-3 on the gradients.)
// pass B stage of h2: [block 0: hi 8 KB | lo 8 KB][block 1: hi | lo], so that a 64-row block is also a 128-row K-major A tile
// [h2_hi ; h2_lo] (the stacked form of the Gram product).
constexpr uint32_t H2_LBO = 16384, H2_LO = 8192;
template <bool PASS_B, int STAT, bool GRAM>
__global__ void __launch_bounds__(PASS_B ? NTHREADS_B : NTHREADS_A, 1) l1_fwd_kernel(const L1Params p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS, not generic LD/ST)
    const int nhl = p.nhl;
    constexpr int NPW = PASS_B ? 4 : 16;       // producer warps: 64 channels x (2 | 8) row parts
    constexpr int PROWS = TILE / (NPW / 2);    // rows per producer thread
    // ---- shared-memory carve-up (all operand images 1 KB aligned) ----
    uint8_t* w2s = smem;                                   // hi 8 KB | lo 8 KB (64 valid rows each)
    uint8_t* w3s = w2s + 16384;                            // [half][hi 16 KB | lo 16 KB]      (pass B)
    uint8_t* h1s = w3s + (PASS_B ? 65536 : 16384);         // pass A: leave 16 KB readable behind W2 (M=128 reads 128 rows)
    uint8_t* h2s = h1s + 2 * 2 * ACT_BYTES;                // 2 stages x (hi|lo)
    uint8_t* xs = h2s + (PASS_B ? 2 * 2 * ACT_BYTES : 0);  // 2 stages x 128 rows x 16 B
    // pass A runs layer 0 on the tensor pipe (TCZ1, see pass D): operand rows [Wh | Wh | Wl | bh bl 0 0] x [xh | xl | xh | 1 1 0 0], K = 16
    constexpr bool TCZ1 = !PASS_B;
    uint8_t* w1a = xs + 2 * TILE * 16;                     // 128 rows (channel, twice) x 128 B            (pass A)
    uint8_t* x16s = w1a + (TCZ1 ? 16384 : 0);              // 2 stages x 128 rows x 128 B                    (pass A)
    uint64_t* bars = reinterpret_cast<uint64_t*>(x16s + (TCZ1 ? 2 * 16384 : 0));
    uint64_t *h1_full = bars, *h1_empty = bars + 2, *d2_full = bars + 4, *d2_empty = bars + 6, *h2_full = bars + 8,
             *h2_empty = bars + 10, *d3_full = bars + 12, *d3_empty = bars + 14, *w_bar = bars + 16, *fin_bar = bars + 17,
             *a_ready = bars + 18, *x16_full = bars + 19, *d1_full = bars + 21, *d1_empty = bars + 22;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / TILE;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h1_full[i], NPW);
            mbar_init(&h1_empty[i], 1);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 4);
            mbar_init(&h2_full[i], 4);
            mbar_init(&h2_empty[i], 1);
            mbar_init(&d3_full[i], 1);
            mbar_init(&d3_empty[i], 4);
        }
        mbar_init(w_bar, 1);
        mbar_init(fin_bar, 1);
        mbar_init(a_ready, 4);
        mbar_init(&x16_full[0], 4);
        mbar_init(&x16_full[1], 4);
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, NPW);
        mbar_fence_init();
    }
    if (warp == 16) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    if (TCZ1 && warp < 4) {
        const int row = warp * 32 + lane, ch = row & 63;
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float v[8] = {s1 * w.x, s1 * w.y, s1 * w.z, s1 * w.w, fmaf(s1, __ldg(p.b1 + ch), t1), 0.f, 0.f, 0.f};
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(w1a + sw128_offset((uint32_t)row, 0)) = make_uint4(h.x, h.y, h.x, h.y);
        *reinterpret_cast<uint4*>(w1a + sw128_offset((uint32_t)row, 1)) = make_uint4(l.x, l.y, (h.z & 0xFFFFu) | (l.z << 16), 0u);
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns.  pass A: D2[b] at 128*b (b = 0,1), W2 A tiles (hi, lo) at 256 / 288, z1' (layer 0, 128 rows) at 320..447.  pass B: D2 at 0 (single: its consumers
    // copy it out at once), H2 Gram accumulator at 128..255 (lanes 0..63 h2_hi rows, 64..127 h2_lo rows; columns 0..63 x h2_hi,
    // 64..127 x h2_lo), D3[half] at 256 + 128*half
    const uint32_t idesc = umma_idesc_bf16(128, TILE) | UMMA_B_MN_MAJOR;
    const uint32_t idesc_kk = umma_idesc_bf16(128, 64), idesc_kk2 = umma_idesc_bf16(128, 128);
    const uint32_t w2_t = tmem_base + 256;
    // pass A: W2 is constant and becomes an A operand in tensor memory (see tmem_put_a_tile); warps 8, 9 own lanes 0..63, warps 10,
    // 11 zero the unused lanes 64..127.  (Not in pass B: there the z3 consumers keep the TMEM read port busy, and A tiles read from
    // TMEM made the pass 8 % slower, measured.)
    constexpr bool W2_TMEM = !PASS_B;
    if (W2_TMEM && warp >= 8 && warp < 12) {
        if (warp < 10) mbar_wait(w_bar, 0);
#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
            if (nhl == 1 && tile == 1) continue;
            tmem_put_a_tile(w2_t + ((uint32_t)((warp & 3) * 32) << 16) + 32 * tile, warp < 10 ? w2s + tile * 8192 : nullptr, (warp & 1) * 32 + lane);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
    }

    if (warp == 16) {
        // ======================= MMA issuer (one elected lane: umma.cuh, lean issue path) =======================
        if (elect_one_sync()) {
            // stage the weight images once (bulk TMA)
            const uint32_t w2_bytes = 8192u * nhl, w3_bytes = PASS_B ? 32768u * nhl : 0u;
            mbar_arrive_expect_tx(w_bar, w2_bytes + w3_bytes);
            tma_bulk_g2s(w2s, p.w2_img, 8192, w_bar);
            if (nhl == 2) tma_bulk_g2s(w2s + 8192, p.w2_img + 16384, 8192, w_bar);
            if (PASS_B) {
                for (int h = 0; h < 2; ++h) {
                    tma_bulk_g2s(w3s + h * 32768, p.w3_img + h * 32768, 16384, w_bar);
                    if (nhl == 2) tma_bulk_g2s(w3s + h * 32768 + 16384, p.w3_img + h * 32768 + 16384, 16384, w_bar);
                }
            }
            mbar_wait(w_bar, 0);
            if (W2_TMEM) {
                mbar_wait(a_ready, 0);
                tc_fence_after_sync();
            }
            const bool split = nhl == 2;
            const UDesc w2_k = udesc_k(smem_u32(w2s));                               // lo image 8 KB further
            const UDesc w3_k = udesc_k(smem_u32(w3s));                               // half h 32 KB further, lo 16 KB further
            const UDesc h1_mn = udesc_mn(smem_u32(h1s), ACT_LBO, ACT_SBO);           // stage b 32 KB further, lo 16 KB further
            const UDesc h2_mn = udesc_mn(smem_u32(h2s), H2_LBO, ACT_SBO);            // stage b 32 KB further, lo 8 KB further
            const UDesc h2_k = udesc_k(smem_u32(h2s));                               // a 64-row block [hi ; lo] as a K-major tile
            constexpr uint32_t STAGE = 2 * ACT_BYTES / 16;
            PROF_DECL(8)
            auto issue_mma2 = [&](int it) {
                const int b = it & 1, u = (it >> 1) & 1;
                PROF_MARK(7)
                mbar_wait(&h1_full[b], u);
                PROF_MARK(0)
                if (PASS_B) {
                    if (it > 0) mbar_wait(&d2_empty[(it - 1) & 1], ((it - 1) >> 1) & 1);      // the one D2 has been copied out
                } else {
                    mbar_wait(&d2_empty[b], u ^ 1);
                }
                PROF_MARK(1)
                tc_fence_after_sync();
                if (W2_TMEM)
                    mma_tmem_weight_act(tmem_base + 128 * b, w2_t, w2_t + 32, h1_mn, b * STAGE, ACT_BYTES / 16, split, idesc);
                else
                    mma_weight_act(tmem_base, w2_k, 0, 8192 / 16, h1_mn, b * STAGE, ACT_BYTES / 16, split, idesc);
                umma_commit(&h1_empty[b]);
                umma_commit(&d2_full[b]);
                PROF_MARK(2)
            };
            int it = 0;
            long long t = blockIdx.x;
            if (TCZ1) {
                // pass A: z1'(it + 1) is issued before z2(it) so that the producers turn it into h1(it + 1) while z2(it) runs
                const UDesc w1a_k = udesc_k(smem_u32(w1a));
                const UDesc x16_k = udesc_k(smem_u32(x16s));
                auto issue_z1 = [&](int j) {
                    mbar_wait(&x16_full[j & 1], (j >> 1) & 1);
                    mbar_wait(d1_empty, (j & 1) ^ 1);                     // the producers have copied z1'(j - 1) out
                    tc_fence_after_sync();
                    umma_ss(tmem_base + 320, w1a_k, 0, x16_k, (uint32_t)(j & 1) * (16384 / 16), idesc_kk2, 0u);
                    umma_commit(d1_full);
                };
                if (t < ntiles) issue_z1(0);
                for (; t < ntiles; t += gridDim.x, ++it) {
                    if (t + gridDim.x < ntiles) issue_z1(it + 1);
                    issue_mma2(it);
                }
            }
            if (!TCZ1 && t < ntiles) issue_mma2(0);
            for (; t < ntiles; t += gridDim.x, ++it) {
                if (t + gridDim.x < ntiles) issue_mma2(it + 1);
                if (PASS_B) {
                    const int b = it & 1, u = (it >> 1) & 1;
                    PROF_MARK(7)
                    mbar_wait(&h2_full[b], u);
                    PROF_MARK(3)
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&d3_empty[h], (it & 1) ^ 1);
                        PROF_MARK(4)
                        tc_fence_after_sync();
                        mma_weight_act(tmem_base + 256 + 128 * h, w3_k, h * (32768 / 16), 16384 / 16, h2_mn, b * STAGE, H2_LO / 16, split, idesc);
                        umma_commit(&d3_full[h]);
                        PROF_MARK(5)
                    }
                    if (GRAM) {
                        // H2 += h2 h2^T, reduction over the tile's rows.  A 64-row block [hi 64 ch ; lo 64 ch] is used as the 128-row A
                        // tile AND as the 128-row B tile of ONE instruction per k-step: the accumulator's four 64 x 64 quadrants
                        // are hi.hi, hi.lo, lo.hi, lo.lo, and H2 is their sum (taken when the accumulator is flushed)
#pragma unroll
                        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t off = b * STAGE + blk * (H2_LBO / 16) + ks * 2;
                                umma_ss(tmem_base + 128, h2_k, off, h2_k, off, split ? idesc_kk2 : idesc_kk, (it == 0 && blk == 0 && ks == 0) ? 0u : 1u);
                            }
                        }
                    }
                    umma_commit(&h2_empty[b]);
                    PROF_MARK(6)
                }
            }
#ifdef FACL_PROFILE_ROLES
            if (PASS_B && blockIdx.x == 1 && it > 0)
                printf("fwd pass B MMA thread, cycles/tile: wait h1_full %lld | wait d2_empty %lld | issue z2 %lld | wait h2_full %lld | wait d3_empty %lld | "
                       "issue z3 %lld | issue Gram %lld | other %lld\n", prof_[0] / it, prof_[1] / it, prof_[2] / it, prof_[3] / it, prof_[4] / it,
                       prof_[5] / it, prof_[6] / it, prof_[7] / it);
#endif
            if (PASS_B && GRAM) umma_commit(fin_bar);
        }
    } else if (PASS_B ? (warp == 10 || warp == 11 || warp == 14 || warp == 15) : (warp < 8 || warp >= 17)) {
        // ======================= producers: x -> h1 = relu(bn1(W1 x + b1)), thread = (channel, row quarter) ===============
        if constexpr (TCZ1) {
            // ---- pass A: h1 = relu(z1') with z1' read from tensor memory.  TMEM lane quarter = warp % 4: quarters 0, 1 hold channels
            //      0..63 and take rows 0..63 of the tile, quarters 2, 3 hold the second copy and take rows 64..127; the four warps of a
            //      quarter split its 64 rows.  Warps 0..3 also write the K = 16 operand rows of the NEXT tile. ----
            const int quarter = warp & 3;
            const int kq = warp < 8 ? (warp >> 2) : 2 + ((warp - 17) >> 2);
            const int ch = (quarter & 1) * 32 + lane;
            const int rowbase = (quarter >> 1) * 64 + kq * 16;
            const bool writer = warp < 4;
            const int wrow = warp * 32 + lane;
            float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);
            if (writer && (long long)blockIdx.x < ntiles) xnext = __ldg(reinterpret_cast<const float4*>(p.xt) + (long long)blockIdx.x * TILE + wrow);
            auto put_x = [&](int j, long long tj) {                 // rows of tile j (global tile index tj) -> slot j & 1
                const float xv[8] = {xnext.x, xnext.y, xnext.z, xnext.w, 1.f, 0.f, 0.f, 0.f};
                uint4 h, l;
                split_bf16x8(xv, h, l);
                uint8_t* x16 = x16s + (j & 1) * 16384;
                *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)wrow, 0)) = make_uint4(h.x, h.y, l.x, l.y);
                *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)wrow, 1)) = make_uint4(h.x, h.y, h.z | (h.z << 16), 0u);
                if (tj + gridDim.x < ntiles) xnext = __ldg(reinterpret_cast<const float4*>(p.xt) + (tj + gridDim.x) * TILE + wrow);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&x16_full[j & 1]);
            };
            if (writer && (long long)blockIdx.x < ntiles) put_x(0, blockIdx.x);
            int it = 0;
#ifdef FACL_PROFILE_ROLES
            long long prof_wait = 0, prof_comp = 0, prof_fence = 0;
#endif
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int b = it & 1, u = (it >> 1) & 1;
#ifdef FACL_PROFILE_ROLES
                long long c0 = clock64();
#endif
                mbar_wait(d1_full, it & 1);
                tc_fence_after_sync();
                float z[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(320 + rowbase), z);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(d1_empty);
                // z1'(it) is complete, so the slot of tile it - 1 is free: the operand rows of tile it + 1 go there
                if (writer && t + gridDim.x < ntiles) put_x(it + 1, t + gridDim.x);
                mbar_wait(&h1_empty[b], u ^ 1);
#ifdef FACL_PROFILE_ROLES
                long long c1 = clock64();
#endif
                uint8_t* img = h1s + b * 2 * ACT_BYTES;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaxf(z[q * 8 + e], 0.f);
                    store_act8(img, nhl, ch, rowbase / 8 + q, v);
                }
#ifdef FACL_PROFILE_ROLES
                long long c2 = clock64();
#endif
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h1_full[b]);
#ifdef FACL_PROFILE_ROLES
                long long c3 = clock64();
                prof_wait += c1 - c0; prof_comp += c2 - c1; prof_fence += c3 - c2;
#endif
            }
#ifdef FACL_PROFILE_ROLES
            if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 24))
                printf("fwd pass_b=0 producer warp %d: tiles %d wait %lld compute %lld fence %lld (cycles per tile)\n", warp, it,
                       prof_wait / it, prof_comp / it, prof_fence / it);
#endif
        } else {
        const int pw = !PASS_B ? (warp < 8 ? warp : warp - 9) : (warp == 10) ? 0 : (warp == 11) ? 1 : (warp == 14) ? 2 : 3;
        const int ptid = pw * 32 + lane;          // 0 .. 32 NPW - 1
        const int ch = ptid & 63, part = ptid >> 6;
        // BN1 folded into the 4-wide layer: h1 = max(wf . x + bf, 0)
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float wx = s1 * w.x, wy = s1 * w.y, wz = s1 * w.z, ww = s1 * w.w;
        const float bf = fmaf(s1, __ldg(p.b1 + ch), t1);
        int it = 0;
#ifdef FACL_PROFILE_ROLES
        long long prof_wait = 0, prof_comp = 0, prof_fence = 0;
#endif
        float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ptid < TILE && (long long)blockIdx.x < ntiles) xnext = __ldg(reinterpret_cast<const float4*>(p.xt) + (long long)blockIdx.x * TILE + ptid);
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            float4* xtile = reinterpret_cast<float4*>(xs + b * TILE * 16);
            if (ptid < TILE) {
                xtile[ptid] = xnext;
                if (t + gridDim.x < ntiles) xnext = __ldg(reinterpret_cast<const float4*>(p.xt) + (t + gridDim.x) * TILE + ptid);
            }
#ifdef FACL_PROFILE_ROLES
            long long c0 = clock64();
#endif
            named_bar_sync(1, NPW * 32);
            mbar_wait(&h1_empty[b], u ^ 1);
#ifdef FACL_PROFILE_ROLES
            long long c1 = clock64();
#endif
            uint8_t* img = h1s + b * 2 * ACT_BYTES;
#pragma unroll 2
            for (int q = 0; q < PROWS / 8; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float4 x = xtile[part * PROWS + q * 8 + e];
                    v[e] = fmaxf(fmaf(wx, x.x, fmaf(wy, x.y, fmaf(wz, x.z, fmaf(ww, x.w, bf)))), 0.f);
                }
                store_act8(img, nhl, ch, part * (PROWS / 8) + q, v);
            }
#ifdef FACL_PROFILE_ROLES
            long long c2 = clock64();
#endif
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[b]);
#ifdef FACL_PROFILE_ROLES
            long long c3 = clock64();
            prof_wait += c1 - c0; prof_comp += c2 - c1; prof_fence += c3 - c2;
#endif
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 0 && lane == 0 && (pw == 0 || pw == NPW - 1))
            printf("fwd pass_b=%d producer warp %d: tiles %d wait %lld compute %lld fence %lld (cycles per tile)\n", (int)PASS_B, pw, it,
                   prof_wait / it, prof_comp / it, prof_fence / it);
#endif
        }
        if (PASS_B && GRAM && nhl == 2) {
            // these warps sit on TMEM lanes 64..127: the h2_lo rows of the stacked Gram accumulator
            mbar_wait(fin_bar, 0);
            tc_fence_after_sync();
            const int lg = warp & 1, colhalf = (warp >= 14) ? 1 : 0, j = lg * 32 + lane;
            float a[32];
            float a2[32];
            tmem_ld32(tmem_base + ((uint32_t)(64 + lg * 32) << 16) + (uint32_t)(128 + colhalf * 32), a);
            tmem_ld32(tmem_base + ((uint32_t)(64 + lg * 32) << 16) + (uint32_t)(192 + colhalf * 32), a2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(p.gram + j * 64 + colhalf * 32 + i, a[i] + a2[i]);
        }
    } else if (warp == 8 || warp == 9 || warp == 12 || warp == 13) {
        // ======================= z2 consumers, thread = channel j (lanes 0..63 of D2), two column halves =============
        const int lg = warp & 1;                  // warps 8,12 -> TMEM lanes 0..31; 9,13 -> 32..63   (warp % 4 == lg)
        const int colhalf = (warp >= 12) ? 1 : 0;
        const int j = lg * 32 + lane;
        const float b2 = __ldg(p.b2 + j);
        float a2 = 0.f, c2 = 0.f;
        if (PASS_B) {
            a2 = __ldg(p.scale2 + j);
            c2 = fmaf(a2, b2, __ldg(p.shift2 + j));
        }
        float s_acc = 0.f, q_acc = 0.f, hsum = 0.f;
        long long nrows = 0;
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            mbar_wait(&d2_full[b], u);
            tc_fence_after_sync();
            if (PASS_B) {
                // copy this thread's 64 rows out of the single D2 and hand it back before doing the arithmetic
                float v[64];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(colhalf * 64), *reinterpret_cast<float(*)[32]>(&v[0]));
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(colhalf * 64 + 32), *reinterpret_cast<float(*)[32]>(&v[32]));
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d2_empty[b]);
                mbar_wait(&h2_empty[b], u ^ 1);
                uint8_t* img = h2s + b * 2 * ACT_BYTES;
#pragma unroll
                for (int g8 = 0; g8 < 8; ++g8) {
                    float h[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        h[e] = fmaxf(fmaf(a2, v[g8 * 8 + e], c2), 0.f);
                        if (GRAM) hsum += h[e];
                    }
                    store_act8(img, nhl, j, colhalf * 8 + g8, h, H2_LBO, H2_LO);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h2_full[b]);
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(128 * b + colhalf * 64 + q * 32), v);
                    tmem_ld_wait();
                    if (STAT == 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            s_acc += v[i];
                            q_acc = fmaf(v[i], v[i], q_acc);
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d2_empty[b]);
            }
            nrows += 64;
        }
        if (PASS_B && GRAM) {
            atomicAdd(p.hsum + j, hsum);
            // lanes 0..63 of the Gram accumulator: the h2_hi rows
            mbar_wait(fin_bar, 0);
            tc_fence_after_sync();
            float a[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(128 + colhalf * 32), a);
            if (nhl == 2) {      // + the (.) h2_lo quadrant
                float a2[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(192 + colhalf * 32), a2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] += a2[i];
            } else {
                tmem_ld_wait();
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(p.gram + j * 64 + colhalf * 32 + i, a[i]);
        }
        if (!PASS_B && STAT == 1) {
            // statistics of z2 = acc + b2 from the sums of acc
            const float n = (float)nrows;
            float* st = p.stats + ((long long)(blockIdx.x * 2 + colhalf) * 64 + j) * 2;
            st[0] = fmaf(n, b2, s_acc);
            st[1] = q_acc + 2.f * b2 * s_acc + n * b2 * b2;
        }
    } else if (PASS_B && warp < 8) {
        // ======================= z3 consumers (pass B), thread = channel c =======================
        {
            const int h = warp >> 2, lq = warp & 3;
            const int c = h * 128 + lq * 32 + lane;
            const float b3 = __ldg(p.b3 + c);
            const float sgn = (__ldg(p.gamma3 + c) >= 0.f) ? 1.f : -1.f;
            const bool allpos = __all_sync(0xffffffffu, sgn > 0.f);
            const int K = p.K, groups = TILE / K;
            float s_acc = 0.f, q_acc = 0.f;
            long long nrows = 0;
            int it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                mbar_wait(&d3_full[h], it & 1);
                tc_fence_after_sync();
                float best = -INFINITY;
                int barg = 0;
                // POS: every channel of this warp takes the MAXIMUM (BatchNorm weight >= 0, the usual case): the per-element sign multiply
                // is skipped.  The z3 consumers share their issue slots with the h1 producers and are the period of this pass.
                auto chunk = [&](auto pos_c, const float (&v)[32], int q) {
                    constexpr bool POS = decltype(pos_c)::value;
                    if (K >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (STAT == 1) {
                                s_acc += v[i];
                                q_acc = fmaf(v[i], v[i], q_acc);
                            }
                            const float sv = POS ? v[i] : v[i] * sgn;
                            barg = (sv > best) ? (q * 32 + i) : barg;
                            best = fmaxf(best, sv);
                        }
                        if (((q + 1) * 32) % K == 0) {
                            const long long g = t * groups + (q * 32) / K;
                            p.pooled[(long long)c * p.ldp + g] = fmaf(best, sgn, b3);
                            if (p.pool_arg) p.pool_arg[(long long)c * p.ldp + g] = (unsigned char)(barg & (K - 1));
                            best = -INFINITY;
                            barg = (q + 1) * 32;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (STAT == 1) {
                                s_acc += v[i];
                                q_acc = fmaf(v[i], v[i], q_acc);
                            }
                            const float sv = POS ? v[i] : v[i] * sgn;
                            barg = (sv > best) ? i : barg;
                            best = fmaxf(best, sv);
                            if (((i + 1) & (K - 1)) == 0) {
                                const long long g = t * groups + (q * 32 + i) / K;
                                p.pooled[(long long)c * p.ldp + g] = fmaf(best, sgn, b3);
                                if (p.pool_arg) p.pool_arg[(long long)c * p.ldp + g] = (unsigned char)(barg & (K - 1));
                                best = -INFINITY;
                                barg = i + 1;
                            }
                        }
                    }
                };
#pragma unroll 1
                for (int q = 0; q < TILE / 32; ++q) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(256 + 128 * h + q * 32), v);
                    tmem_ld_wait();
                    if (allpos) chunk(std::true_type{}, v, q); else chunk(std::false_type{}, v, q);
                }
                nrows += TILE;
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d3_empty[h]);
            }
            if (STAT == 1) {
                const float n = (float)nrows;
                float* st = p.stats + ((long long)blockIdx.x * 256 + c) * 2;
                st[0] = fmaf(n, b3, s_acc);
                st[1] = q_acc + 2.f * b3 * s_acc + n * b3 * b3;
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================================
// Backward of net3DV_1.  Activations are RECOMPUTED from the 16-byte input rows (cheaper than storing 384 fp32 per row);
// tiles are 64 batch rows = one max-pool group (K = 64).  All images are [channel][64 rows] (see above): the same bytes
// serve as an MN-major B operand (reduction over channels) and as a K-major operand (reduction over rows).
//
// The BatchNorm-backward map dz = c0*dy + c1*z + c2 is affine in z, and z is linear in the layer input, so the DENSE
// part of every gradient collapses onto 64x64 matrices and never needs z3 (256 channels) again:
//
//   dz3 = k1 (.) z3 + k2 + Sp          Sp[c][r] = k0[c] dP[c][g] at the max-pool winner r = arg[c][g] (pass B records it)
//   dh2 = W3^T dz3 = P3 h2 + q3 + W3^T Sp            P3 = W3^T diag(k1) W3,   q3 = W3^T (k1 (.) b3 + k2)
//   dW3 = dz3 h2^T = diag(k1) W3 H2 + (k1 (.) b3 + k2) s2^T + Sp h2^T            H2 = sum h2 h2^T,  s2 = sum h2
//   dz2 = e0 (.) dh2' + e1 (.) z2 + e2               dh2' = dh2 masked by ReLU2
//   dh1 = W2^T dz2 = (diag(e0) W2)^T dh2' + P2 h1 + q2        P2 = W2^T diag(e1) W2,   q2 = W2^T (e1 (.) b2 + e2)
//   dW2 = diag(e0) (dh2' h1^T) + diag(e1) (W2 H1 + b2 s1^T) + e2 s1^T            H1 = sum h1 h1^T,  s1 = sum h1
//
//   pass C : x -> h1 -> z2 -> h2;  Sp image (one non-zero per channel);  TMEM: dh2 = W3^T Sp + P3 h2,  dW3s += Sp h2^T,
//            H2 += h2 h2^T;  dh2' -> HBM as a bf16 hi/lo image (+ the BN2-backward sums).
//   pass D : x -> h1;  dh2' image by bulk TMA;  TMEM: dh1 = (E0 W2)^T dh2' + P2 h1,  dW2s += dh2' h1^T,  H1 += h1 h1^T;
//            dh1 masked by ReLU1 -> BN1-backward sums and A = sum dh1' x^T (64x4); nothing per-row is written.
//   l1_prep / l1_fin : the 64x64 matrix algebra around the passes;  dW1 follows in closed form from A and the moments.
// =====================================================================================================================
constexpr int BT = 64;                    // batch rows per backward tile
constexpr int IMG64 = 8192;               // [64 ch][64 rows] bf16 image (one half)

struct L1BwdParams {
    const float* xt;
    long long R;
    int nhl;
    int tiles_per_cta;        // contiguous tile range per CTA (multiple of 4)
    int kshift;               // log2(K / 64): a max-pool group of K = 64 << kshift neighbours spans 1 << kshift consecutive tiles
    const float* w1;  const float* b1;  const float* scale1;  const float* shift1;
    // pass C
    const uint8_t* w2_img;  const float* b2;  const float* scale2;  const float* shift2;
    const uint8_t* w3_img;
    const uint8_t* p3_img;    // packed image of P3 (64x64)
    const float* q3;          // [64]
    const unsigned char* arg; // [256][ldp] max-pool winner inside its group (pass B)
    const float* dpooled;     // [256][ldp] gradient w.r.t. the pooled BN3 output, ReLU-masked
    long long ldp;
    const float* c3_0;        // BN3 backward coefficient c0 (slot 2)
    uint8_t* dh2;             // [R/64][hi 8 KB | lo 8 KB] image of the masked gradient w.r.t. relu(bn2(z2))  (C out, D in)
    float* dw3;               // [256][64]  += Sp h2^T (atomic)
    float* gram;              // [64][64]   += H2 (pass C) / H1 (pass D)
    float* hsum;              // [64]       += s2 / s1
    // pass D
    const uint8_t* e0w2_img;  // packed image of diag(e0) W2
    const uint8_t* p2_img;    // packed image of P2
    const float* q2;          // [64]
    float* dw2s;              // [64][64] += dh2' h1^T (atomic)
    float* amat;              // [4*grid][64][4] per-CTA (consumer warp group) sum dh1' x^T
    float* stats;             // [4*grid][64][2]: pass C (sum dh2', sum dh2' z2); pass D (sum dh1', sum dh1' z1)
    // optional test hooks (pass C): the ReLU decisions of the recomputed forward, so a checker can impose them
    unsigned char* dbg_mask1; // [R/64][64][64]  h1 > 0 as pass D sees it (the only place the ReLU1 mask enters the backward)
    unsigned char* dbg_mask2; // [R/64][64][64]  h2 > 0
};

__device__ __forceinline__ void store_img8(uint8_t* img, int nhl, int img_bytes, int row, int chunk, const float (&v)[8]) {
    uint32_t off = sw128_offset((uint32_t)row, (uint32_t)chunk);
    if (nhl == 2) {
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(img + off) = h;
        *reinterpret_cast<uint4*>(img + img_bytes + off) = l;
    } else {
        *reinterpret_cast<uint4*>(img + off) = pack_bf16x8(v);
    }
}

// the per-element part of the dh2 role for 16 rows of channel j (two 8-row chunks starting at `chunk0`): + q3, ReLU2 mask taken from
// the bf16 h2 image (hh / hl = its hi / lo words), BN2-backward sums, masked gradient -> bf16 hi / lo image in HBM
__device__ __forceinline__ void dh2_rows16(const float (&gv)[16], const uint4 (&hh)[2], const uint4 (&hl)[2], uint8_t* out, int nhl, int j,
                                           int chunk0, float q3, float c2, float& s_acc, float& q_acc) {
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
        const uint32_t hw[4] = {hh[g8].x, hh[g8].y, hh[g8].z, hh[g8].w}, lw[4] = {hl[g8].x, hl[g8].y, hl[g8].z, hl[g8].w};
        float v8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t hb = (e & 1) ? (hw[e >> 1] & 0xFFFF0000u) : (hw[e >> 1] << 16);     // bf16 -> fp32 bits
            const uint32_t lb = (e & 1) ? (lw[e >> 1] & 0xFFFF0000u) : (lw[e >> 1] << 16);
            const float h2v = __uint_as_float(hb) + __uint_as_float(lb);
            const float v = hb ? gv[g8 * 8 + e] + q3 : 0.f;                  // bf16(h2) != 0  <=>  ReLU2 active
            s_acc += v;
            q_acc = fmaf(v, h2v - c2, q_acc);
            v8[e] = v;
        }
        store_img8(out, nhl, IMG64, j, chunk0 + g8, v8);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// pass C   (22 warps)
// warps 0,1,4,5     : Sp producers (thread = channels c and c + 64): one bf16 hi/lo pair per channel and tile
// warps 2,3,6,7     : partners of the dh2 consumers: they own TMEM lanes 64..127, where the A_lo products of dh2 land (see below)
//                     (all of warps 0-7 flush dW3s at the end)
// warps 8,9,12,13   : z2 -> h2 image (thread = channel j, TMEM lanes 0..63)
// warps 10,11,14,15 : h1 image producers: relu(z1') read from tensor memory (layer 0 is one K = 16 instruction per tile); warps 10, 11
//                     also write the K = 16 operand rows of the tile two ahead
// warps 16,17,20,21 : dh2 consumers (thread = channel j): + q3, ReLU2 mask, BN2-backward sums, image -> HBM
// warp 18           : MMA issuer (warp 19 idles)
// TMEM columns: D2[b] 0/64 (z1'(t), then z2(t)), DH2 128..255 ([B_hi | B_lo] column halves), DW3s[h] 256/320, A tiles W2_hi 384, W2_lo 416,
//               [P3_hi ; P3_lo] 448, W1' (layer 0, K = 16) 480..487
// (H2 = sum h2 h2^T and s2, which dW3 also needs, were accumulated by pass B of the forward)
//
// tcgen05.mma time on these narrow tiles is set by the operand bytes it pulls from shared memory (the 4 KB A tile above
// all), not by the math.  For dh2 the bf16x3 scheme is ONE instruction per k-step: the B images are stored [hi | lo] and the A
// operand is stacked [A_hi 64 rows ; A_lo 64 rows], so an M = 128, N = 128 instruction yields A_hi B_hi | A_hi B_lo in lanes
// 0..63 and A_lo B_hi (| A_lo B_lo, unused) in lanes 64..127.  A consumer thread adds its two column halves in registers; the
// lane-64..127 half is read by the partner warp, and the two warps swap half of their rows through shared memory so that each
// finishes 16 complete rows.  (z2 and dW3s use the plain 3-term form: their accumulators have no columns to spare.)
// --------------------------------------------------------------------------------------------------------------------
constexpr int C_THREADS = 22 * 32;

__global__ void __launch_bounds__(C_THREADS, 1) l1_bwd_c_kernel(const L1BwdParams p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS, not generic LD/ST)
    const int nhl = p.nhl;
    uint8_t* w2s = smem;                       // hi 8 KB | lo 8 KB
    uint8_t* w3s = w2s + 16384;                // [half][hi 16 KB | lo 16 KB]
    uint8_t* p3s = w3s + 65536;                // hi 8 KB | lo 8 KB
    uint8_t* h1s = p3s + 16384;                // 2 stages x (hi 8 KB | lo 8 KB): with one stage, h1(t + 1) could not be produced before
                                               // z2(t) had read h1(t), and the 64-row recompute sat on the critical chain of the tile
    uint8_t* h2s = h1s + 2 * 2 * IMG64;        // 2 stages x (hi | lo)
    uint8_t* sps = h2s + 2 * 2 * IMG64;        // Sp [256 c][64 r]: hi 32 KB | lo 32 KB; in bf16 mode (no lo part) two STAGES of 32 KB, so that
                                               // Sp(t + 1) is written while W3^T Sp(t) still reads Sp(t)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sps + 65536);
    uint64_t *h1_full = bars + 18, *h1_empty = bars + 20, *x16_full = bars + 22, *d1_full = bars + 24, *d2_full = bars + 2, *d2_empty = bars + 4, *h2_full = bars + 6,
             *h2_empty = bars + 8, *sp_full = bars + 10, *sp_empty = bars + 26, *dh_full = bars + 12, *dh_empty = bars + 13,
             *a_ready = bars + 14, *w_bar = bars + 16, *fin_bar = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    // Layer 0 runs on the tensor pipe as in passes A and D: z1'(t) is ONE K = 16 instruction into the columns of D2[t & 1] (z2(t) overwrites
    // them only after the h1 producers have read z1'), its A tile (W1 folded with BN1, [hi | hi | lo | bias] K slots, rows 64..127 a copy
    // of rows 0..63) sits in tensor memory columns 480..487 and its B tiles (the x rows of a tile as K = 16 bf16) in the P3 staging area.
    uint8_t* x16s = p3s;                       // 2 x 64 rows x 128 B (after the A tiles are in tensor memory)
    // The lo rows of the stacked A operands ([hi 64 rows ; lo 64 rows]) leave their products in TMEM lanes 64..127, which the dh2
    // consumers (warps of lanes 0..63) cannot read.  Four Sp-producer warps own those lanes: each is paired with the consumer warp of
    // the same (channel group, column half), the two swap half of their 32 rows through shared memory (as in pass D) and each
    // does the per-element work and the global stores of 16 rows.  The W2 staging area is free once its A tiles are in tensor memory:
    // it is the buffer of the swap (4 pairs x [consumer -> partner 2 KB | partner -> consumer 2 KB], a second barrier after the reads).

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / BT;
    const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
    const int my_tiles = (int)((ntiles - t0 < p.tiles_per_cta) ? (ntiles - t0) : p.tiles_per_cta);   // > 0 by the launch

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h1_full[i], 4);
            mbar_init(&h1_empty[i], 1);
            mbar_init(&x16_full[i], 2);
            mbar_init(&d1_full[i], 1);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 4);        // the h2 producers read z2 (the dh2 consumers take mask and z2 from the h2 image)
        }
        mbar_init(dh_full, 1);
        mbar_init(dh_empty, nhl == 2 ? 8 : 4);     // the dh2 consumers + (bf16x3) the warps that forward the lo lanes

        for (int i = 0; i < 2; ++i) {
            mbar_init(&h2_full[i], 4);
            mbar_init(&h2_empty[i], 9);        // tcgen05.commit of the MMAs that read the stage + the eight warps of the dh2 role
        }
        for (int i = 0; i < 2; ++i) {          // bf16 mode: the unused lo half of the Sp image is a second stage (see `sps`)
            mbar_init(&sp_full[i], 4);         // warps 0, 1, 4, 5 write Sp (two channels per thread); 2, 3, 6, 7 are the dh2 partners
            mbar_init(&sp_empty[i], 1);
        }
        mbar_init(a_ready, 4);
        mbar_init(w_bar, 1);
        mbar_init(fin_bar, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 65536 / 16; i += C_THREADS) reinterpret_cast<uint4*>(sps)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 18) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc_mn = umma_idesc_bf16(128, BT) | UMMA_B_MN_MAJOR;                      // weights x activation image
    const uint32_t idesc_kk = umma_idesc_bf16(128, 64);                                        // reduction over rows
    const int nB = (nhl == 2) ? 2 * BT : BT;                                                   // [hi | lo] stacked B operand
    const uint32_t idesc_mn2 = umma_idesc_bf16(128, nB) | UMMA_B_MN_MAJOR;
    const uint32_t idesc_dg2 = umma_idesc_bf16(128, nB) | UMMA_A_MN_MAJOR | UMMA_B_MN_MAJOR;
    // this warp's 32 lanes of the layer-0 A tile (row = channel `ch`): K slots [Wh0..3 | Wh0..3 | Wl0..3 | bh bl 0 0], two bf16 per column
    auto put_w1a = [&](int ch) {
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float v[8] = {s1 * w.x, s1 * w.y, s1 * w.z, s1 * w.w, fmaf(s1, __ldg(p.b1 + ch), t1), 0.f, 0.f, 0.f};
        uint4 h, l;
        split_bf16x8(v, h, l);
        const uint32_t u[8] = {h.x, h.y, h.x, h.y, l.x, l.y, (h.z & 0xFFFFu) | (l.z << 16), 0u};
        tmem_st8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 480, u);
    };

    if (warp == 18) {
        // one elected thread issues every bulk copy, tcgen05.mma and tcgen05.commit of the CTA (elect.sync: see umma.cuh, lean issue path)
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(w_bar, (8192u + 32768u + 8192u) * nhl);
            tma_bulk_g2s(w2s, p.w2_img, 8192, w_bar);
            tma_bulk_g2s(p3s, p.p3_img, 8192, w_bar);
            if (nhl == 2) {
                tma_bulk_g2s(w2s + 8192, p.w2_img + 16384, 8192, w_bar);
                tma_bulk_g2s(p3s + 8192, p.p3_img + 16384, 8192, w_bar);
            }
            for (int h = 0; h < 2; ++h) {
                tma_bulk_g2s(w3s + h * 32768, p.w3_img + h * 32768, 16384, w_bar);
                if (nhl == 2) tma_bulk_g2s(w3s + h * 32768 + 16384, p.w3_img + h * 32768 + 16384, 16384, w_bar);
            }
            mbar_wait(w_bar, 0);
            mbar_wait(a_ready, 0);                  // W2 / P3 A tiles are in tensor memory
            tc_fence_after_sync();
            const uint32_t w2_hi = tmem_base + 384, w2_lo = tmem_base + 416;
            const uint32_t p3_hi = tmem_base + 448;      // lanes 0..63 P3_hi, lanes 64..127 P3_lo
            // operand descriptors, built once; all per-instruction offsets below are immediates in 16-byte units
            const UDesc h1_mn = udesc_mn(smem_u32(h1s), 8192, 1024);            // h1 [64 ch][64 rows] as B of W2 h1 (reduction over channels)
            const UDesc h2_k = udesc_k(smem_u32(h2s));                           // h2 stage 0, K-major: B of Sp h2^T (reduction over rows)
            const UDesc h2_mn = udesc_mn(smem_u32(h2s), 8192, 1024);             // h2 stage 0, MN-major [hi | lo]: B of P3 h2
            const UDesc sp_k = udesc_k(smem_u32(sps));                           // Sp K-major: A of Sp h2^T
            const UDesc sp_mn = udesc_mn(smem_u32(sps), 32768, 1024);            // Sp MN-major [hi | lo]: B of W3^T Sp
            const UDesc w3_mn = udesc_mn(smem_u32(w3s), 16384, 1024);            // W3 read transposed: A of W3^T Sp, [hi | lo] 16 KB apart
            constexpr uint32_t LO64 = IMG64 / 16;                                // hi -> lo half of a [64][64] image
            const bool split = nhl == 2;
            PROF_DECL(8)
            auto issue_z2 = [&](int it) {          // z2(it) = W2 h1(it): plain 3-term form, double-buffered accumulator
                const int b = it & 1, u = (it >> 1) & 1;
                mbar_wait(&h1_full[b], u);
                PROF_MARK(3)
                mbar_wait(&d2_empty[b], u ^ 1);
                PROF_MARK(4)
                tc_fence_after_sync();
                const uint32_t d = tmem_base + 64 * b;
                const uint32_t hs = (uint32_t)b * (2 * IMG64 / 16);                  // h1 stage offset
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_ts(d, w2_hi + ks * 8, h1_mn, hs + ks * 128, idesc_mn, ks > 0 ? 1u : 0u);
                    if (split) {
                        umma_ts(d, w2_hi + ks * 8, h1_mn, hs + LO64 + ks * 128, idesc_mn, 1u);
                        umma_ts(d, w2_lo + ks * 8, h1_mn, hs + ks * 128, idesc_mn, 1u);
                    }
                }
                umma_commit(&h1_empty[b]);
                umma_commit(&d2_full[b]);
                PROF_MARK(5)
            };
            // z1'(j) = W1' x(j) into D2[j & 1]: two tiles ahead of Sp h2^T, one ahead of z2 (the producers need it to make h1(j))
            const uint32_t w1a_t = tmem_base + 480;
            const UDesc x16_k = udesc_k(smem_u32(x16s));
            auto issue_z1 = [&](int j) {
                const int b = j & 1, u = (j >> 1) & 1;
                mbar_wait(&x16_full[b], u);
                mbar_wait(&d2_empty[b], u ^ 1);                    // z2(j - 2) has been copied out of these columns
                tc_fence_after_sync();
                umma_ts(tmem_base + 64 * b, w1a_t, x16_k, (uint32_t)b * (IMG64 / 16), idesc_kk, 0u);
                umma_commit(&d1_full[b]);
            };
            issue_z1(0);
            if (my_tiles > 1) issue_z1(1);
            issue_z2(0);
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1, u = (it >> 1) & 1;
                const uint32_t st = (uint32_t)b * (2 * IMG64 / 16);               // h2 stage offset
                const uint32_t acc0 = it == 0 ? 0u : 1u;
                const int ss = split ? 0 : b, sp_par = split ? (it & 1) : u;      // Sp stage / phase
                const uint32_t so = (uint32_t)ss * (32768 / 16);
                mbar_wait(&h2_full[b], u);
                PROF_MARK(0)
                mbar_wait(&sp_full[ss], sp_par);
                PROF_MARK(1)
                tc_fence_after_sync();
#pragma unroll
                for (int h = 0; h < 2; ++h) {        // dW3s += Sp h2^T   (A = Sp half h: 16 KB apart, lo 32 KB further)
                    const uint32_t d = tmem_base + 256 + 64 * h;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_ss(d, sp_k, so + h * 1024 + ks * 2, h2_k, st + ks * 2, idesc_kk, ks > 0 ? 1u : acc0);
                        if (split) {
                            umma_ss(d, sp_k, h * 1024 + ks * 2, h2_k, st + LO64 + ks * 2, idesc_kk, 1u);
                            umma_ss(d, sp_k, 2048 + h * 1024 + ks * 2, h2_k, st + ks * 2, idesc_kk, 1u);
                        }
                    }
                }
                PROF_MARK(2)
                if (it + 1 < my_tiles) issue_z2(it + 1);
                if (it + 2 < my_tiles) issue_z1(it + 2);
                mbar_wait(dh_empty, (it & 1) ^ 1);
                PROF_MARK(6)
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {   // dh2 = W3^T Sp: reduction over the 256 channels
                    const uint32_t wa = (ks >> 3) * 2048 + (ks & 7) * 128;
                    umma_ss(tmem_base + 128, w3_mn, wa, sp_mn, so + ks * 128, idesc_dg2, ks > 0 ? 1u : 0u);
                }
                umma_commit(&sp_empty[ss]);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)      // dh2 += P3 h2: [P3_hi ; P3_lo] x [h2_hi | h2_lo]
                    umma_ts(tmem_base + 128, p3_hi + ks * 8, h2_mn, st + ks * 128, idesc_mn2, 1u);
                umma_commit(dh_full);
                umma_commit(&h2_empty[b]);
                PROF_MARK(7)
            }
            umma_commit(fin_bar);
#ifdef FACL_PROFILE_ROLES
            if (blockIdx.x == 1)
                printf("pass C nhl=%d MMA thread, cycles/tile: wait h2_full %lld | wait sp_full %lld | issue SpH %lld | wait h1_full %lld | wait d2_empty %lld | "
                       "issue z2 %lld | wait dh_empty %lld | issue WSp+P3h2 %lld\n", nhl, prof_[0] / my_tiles, prof_[1] / my_tiles, prof_[2] / my_tiles,
                       prof_[3] / my_tiles, prof_[4] / my_tiles, prof_[5] / my_tiles, prof_[6] / my_tiles, prof_[7] / my_tiles);
#endif
        }
    } else if (warp == 10 || warp == 11 || warp == 14 || warp == 15) {
        // ---- x -> h1 producers ----
        if (warp >= 14) {      // warps 14, 15 own TMEM lanes 64..127: P3_lo below P3_hi (stacked A operand), zeros under the W2 tiles
            if (nhl == 2) mbar_wait(w_bar, 0);
#pragma unroll 1
            for (int tile = 0; tile < 4; ++tile)
                tmem_put_a_tile(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384 + 32 * tile, (tile == 2 && nhl == 2) ? p3s + 8192 : nullptr,
                                (warp & 1) * 32 + lane);
            put_w1a((warp & 1) * 32 + lane);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        const int pw = (warp == 10) ? 0 : (warp == 11) ? 1 : (warp == 14) ? 2 : 3;
        const int ptid = pw * 32 + lane, ch = ptid & 63, half = ptid >> 6;      // TMEM lane 64 + ch: the second copy of the channels
        const bool writer = ptid < BT;                                           // warps 10, 11: one x row each
        const float4* xg = reinterpret_cast<const float4*>(p.xt) + t0 * BT;
        auto put_x = [&](int j, float4 x) {                                      // rows of tile j -> K = 16 operand rows, slot j & 1
            const float xv[8] = {x.x, x.y, x.z, x.w, 1.f, 0.f, 0.f, 0.f};
            uint4 h, l;
            split_bf16x8(xv, h, l);
            uint8_t* x16 = x16s + (j & 1) * IMG64;
            *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)ptid, 0)) = make_uint4(h.x, h.y, l.x, l.y);
            *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)ptid, 1)) = make_uint4(h.x, h.y, h.z | (h.z << 16), 0u);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x16_full[j & 1]);
        };
        if (writer) {
            mbar_wait(a_ready, 0);                                               // the P3 staging area has been copied to tensor memory
            put_x(0, __ldg(xg + ptid));
            if (my_tiles > 1) put_x(1, __ldg(xg + BT + ptid));
        }
        PROF_DECL(4)
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (writer && it + 2 < my_tiles) xn = __ldg(xg + (long long)(it + 2) * BT + ptid);
            mbar_wait(&d1_full[b], u);
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(64 * b + half * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            PROF_MARK(0)
            if (writer && it + 2 < my_tiles) put_x(it + 2, xn);                  // z1'(it) is complete: its operand slot is free
            mbar_wait(&h1_empty[b], u ^ 1);
            PROF_MARK(1)
            uint8_t* img = h1s + b * 2 * IMG64;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(z[q * 8 + e], 0.f);
                store_img8(img, nhl, IMG64, ch, half * 4 + q, v);
            }
            PROF_MARK(2)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[b]);
            PROF_MARK(3)
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 1 && ptid == 0)
            printf("pass C h1 producer, cycles/tile: wait z1 + ld %lld | x16 + wait h1_empty %lld | relu + store %lld | fence+arrive %lld\n",
                   prof_[0] / my_tiles, prof_[1] / my_tiles, prof_[2] / my_tiles, prof_[3] / my_tiles);
#endif
    } else if (warp == 8 || warp == 9 || warp == 12 || warp == 13) {
        // ---- z2 -> h2 = relu(bn2(z2)) image (thread = channel j) ----
        const int lg = warp & 1, colhalf = (warp >= 12) ? 1 : 0;
        const int j = lg * 32 + lane;
        const float a2 = __ldg(p.scale2 + j);
        const float c2 = fmaf(a2, __ldg(p.b2 + j), __ldg(p.shift2 + j));
        PROF_DECL(4)
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            mbar_wait(&d2_full[b], u);
            PROF_MARK(0)
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(64 * b + colhalf * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[b]);
            PROF_MARK(1)
            mbar_wait(&h2_empty[b], u ^ 1);
            PROF_MARK(2)
            uint8_t* img = h2s + b * 2 * IMG64;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = fmaxf(fmaf(a2, z[g8 * 8 + e], c2), 0.f);
                store_img8(img, nhl, IMG64, j, colhalf * 4 + g8, h);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_full[b]);
            PROF_MARK(3)
            if (p.dbg_mask2) {
                for (int i = 0; i < 32; ++i)
                    p.dbg_mask2[((t0 + it) * 64 + j) * 64 + colhalf * 32 + i] = fmaf(a2, z[i], c2) > 0.f ? 1 : 0;
            }
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 1 && warp == 8 && lane == 0)
            printf("pass C h2 producer, cycles/tile: wait d2_full %lld | ld z2 + arrive %lld | wait h2_empty %lld | compute+store+fence+arrive %lld\n",
                   prof_[0] / my_tiles, prof_[1] / my_tiles, prof_[2] / my_tiles, prof_[3] / my_tiles);
#endif
    } else if (warp == 16 || warp == 17 || warp == 20 || warp == 21) {
        // ---- dh2 consumers (thread = channel j): + q3, ReLU2 mask, BN2 backward sums, masked gradient image -> HBM ----
        const int lg = warp & 1, colhalf = (warp >= 20) ? 1 : 0;
        const int j = lg * 32 + lane;
        const int pair = lg + 2 * colhalf;
        if (warp < 18) {       // warps 16, 17 own TMEM lanes 0..63: rows of W2 (hi, lo) and P3 (hi, lo) -> A tiles at columns 384..511
            mbar_wait(w_bar, 0);
#pragma unroll 1
            for (int tile = 0; tile < 4; ++tile) {
                if (nhl == 1 && (tile & 1)) continue;
                const uint8_t* img = (tile < 2 ? w2s : p3s) + (tile & 1) * 8192;
                tmem_put_a_tile(tmem_base + ((uint32_t)(lg * 32) << 16) + 384 + 32 * tile, img, j);
            }
            put_w1a(j);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        const float b2 = __ldg(p.b2 + j), a2 = __ldg(p.scale2 + j);
        const float c2 = fmaf(a2, b2, __ldg(p.shift2 + j));
        const float q3 = __ldg(p.q3 + j);
        // The ReLU2 mask and z2 come from the h2 image the producers left in shared memory, not from the z2 accumulator: under the
        // MMA stream's TMEM traffic a tcgen05.ld round trip costs ~500 cycles (role profile), and the SINGLE dh2 accumulator is held
        // for as long as this role's loads take.  With no z2 in registers both halves of dh2 are loaded under one wait (1500 -> 500
        // cycles of hold per tile).  h2 = a2 z + c2 > 0 on the active rows, so z2 + b2 = (h2 - c2) / a2 + b2 there, and only active
        // rows contribute to sum dh2' (z2 + b2):  q = sum dh2' (h2 - c2),  sum dh2' (z2 + b2) = q / a2 + b2 sum dh2'.
        // (a2 == 0, a BatchNorm weight that is exactly zero: l1_gamma0_fix_kernel recomputes that channel's sum from the inputs.)
        float s_acc = 0.f, q_acc = 0.f;
        PROF_DECL(5)
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1;
            mbar_wait(dh_full, it & 1);                 // implies h2(it) is complete in stage b (the MMAs that read it have run)
            PROF_MARK(2)
            tc_fence_after_sync();
            float g[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(128 + colhalf * 32), g);
            if (nhl == 2) {
                float gl[32];                          // the B_lo products (columns 64..127)
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(192 + colhalf * 32), gl);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) g[i] += gl[i];
            } else {
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(dh_empty);
            PROF_MARK(3)
            // this thread's 16 rows of h2 (2 chunks of 8, hi and lo) into registers at once, so that the stage can be handed back to
            // the h2 producers before the arithmetic and the global stores below (holding it through them stalled the producers of
            // tile it + 2, measured)
            const uint8_t* h2img = h2s + b * 2 * IMG64;
            uint4 hh[2], hl[2];
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
                const uint32_t hoff = sw128_offset((uint32_t)j, (uint32_t)(colhalf * 4 + g8));
                hh[g8] = *reinterpret_cast<const uint4*>(h2img + hoff);
                hl[g8] = (nhl == 2) ? *reinterpret_cast<const uint4*>(h2img + IMG64 + hoff) : make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_empty[b]);   // the h2 stage may be refilled once the MMA stream and the eight warps have read it
            // rows 16..31 go to the partner warp; its A_lo B_hi half of rows 0..15 (TMEM lanes 64..127) comes back
            uint8_t* ex = w2s + pair * 4096 + lane * 16;
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4)
                *reinterpret_cast<float4*>(ex + r4 * 512) = make_float4(g[16 + 4 * r4], g[17 + 4 * r4], g[18 + 4 * r4], g[19 + 4 * r4]);
            named_bar_sync(2 + pair, 64);
            float gv[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) gv[r] = g[r];
            if (nhl == 2) {
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const float4 f = *reinterpret_cast<const float4*>(ex + 2048 + r4 * 512);
                    gv[4 * r4] += f.x; gv[4 * r4 + 1] += f.y; gv[4 * r4 + 2] += f.z; gv[4 * r4 + 3] += f.w;
                }
            }
            named_bar_sync(2 + pair, 64);                 // both warps have read: the (single) swap buffer may be rewritten
            dh2_rows16(gv, hh, hl, p.dh2 + (t0 + it) * (2 * IMG64), nhl, j, colhalf * 4, q3, c2, s_acc, q_acc);
            PROF_MARK(4)
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 1 && warp == 16 && lane == 0)
            printf("pass C dh2 consumer, cycles/tile: wait dh_full %lld | ld dh2 + arrive %lld | mask+sums+global store %lld\n",
                   prof_[2] / my_tiles, prof_[3] / my_tiles, prof_[4] / my_tiles);
#endif
        q_acc = (a2 != 0.f) ? q_acc / a2 + b2 * s_acc : 0.f;
        float* st = p.stats + ((long long)(blockIdx.x * 4 + colhalf) * 64 + j) * 2;
        st[0] = s_acc;
        st[1] = q_acc;
    } else if (warp < 8) {
        // ---- warps 0, 1, 4, 5: Sp producers (thread = channels c and c + 64): k0 dP at the max-pool winner, everything else stays
        //      zero.  warps 2, 3, 6, 7: partners of the dh2 consumer warps.  All eight flush dW3s at the end. ----
        const int h = warp >> 2, lq = warp & 3;
        const int c = h * 128 + lq * 32 + lane;
        // K = 64: group == tile.  K = 128 / 256: the group of tile t is t >> kshift and its winner (position 0..K-1 inside the
        // group) lies in this tile iff (position >> 6) == (t & kmask); the other tiles of the group get a zero column
        const int ksh = p.kshift, kmask = (1 << ksh) - 1;
        // warps 2, 3, 6, 7 own TMEM lanes 64..127 and are the partners of the dh2 consumer warps (see the top of the kernel): after
        // Sp(it) is out they take rows 16..31 of dh2(it - 1) for channel j of their (channel group, column half)
        const bool partner = (warp & 2) != 0;
        const int pj = (warp & 1) * 32 + lane, pcol = warp >> 2, ppair = (warp & 1) + 2 * pcol;
        float pq3 = 0.f, pc2 = 0.f, ps_acc = 0.f, pq_acc = 0.f;
        if (partner) {
            const float a2 = __ldg(p.scale2 + pj);
            pc2 = fmaf(a2, __ldg(p.b2 + pj), __ldg(p.shift2 + pj));
            pq3 = __ldg(p.q3 + pj);
        }
        auto partner_rows = [&](int t) {
            uint8_t* ex = w2s + ppair * 4096 + lane * 16;
            float gv[16];
            if (nhl == 2) {
                mbar_wait(dh_full, t & 1);
                tc_fence_after_sync();
                float f[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(128 + pcol * 32), f);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(dh_empty);
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4)
                    *reinterpret_cast<float4*>(ex + 2048 + r4 * 512) = make_float4(f[4 * r4], f[4 * r4 + 1], f[4 * r4 + 2], f[4 * r4 + 3]);
#pragma unroll
                for (int r = 0; r < 16; ++r) gv[r] = f[16 + r];
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) gv[r] = 0.f;
            }
            named_bar_sync(2 + ppair, 64);                 // the consumer warp waited for dh_full(t): h2(t) is complete as well
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                const float4 f4 = *reinterpret_cast<const float4*>(ex + r4 * 512);
                gv[4 * r4] += f4.x; gv[4 * r4 + 1] += f4.y; gv[4 * r4 + 2] += f4.z; gv[4 * r4 + 3] += f4.w;
            }
            named_bar_sync(2 + ppair, 64);
            const uint8_t* h2img = h2s + (t & 1) * 2 * IMG64;
            uint4 hh[2], hl[2];
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
                const uint32_t hoff = sw128_offset((uint32_t)pj, (uint32_t)(pcol * 4 + 2 + g8));
                hh[g8] = *reinterpret_cast<const uint4*>(h2img + hoff);
                hl[g8] = (nhl == 2) ? *reinterpret_cast<const uint4*>(h2img + IMG64 + hoff) : make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h2_empty[t & 1]);
            dh2_rows16(gv, hh, hl, p.dh2 + (t0 + t) * (2 * IMG64), nhl, pj, pcol * 4 + 2, pq3, pc2, ps_acc, pq_acc);
        };
        if (partner) {
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) partner_rows(it);
            const float a2 = __ldg(p.scale2 + pj), b2 = __ldg(p.b2 + pj);
            float* st = p.stats + ((long long)(blockIdx.x * 4 + 2 + pcol) * 64 + pj) * 2;
            st[0] = ps_acc;
            st[1] = (a2 != 0.f) ? pq_acc / a2 + b2 * ps_acc : 0.f;
        } else {
            struct SpChannel {
                float k0;
                const float* dpp;
                const unsigned char* argp;
                uint8_t* row_hi;
                int prev[2];
                float4 d4;
                uchar4 a4;
                float dcur;
                int acur;
            } sc[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int cq = c + 64 * q;
                sc[q].k0 = __ldg(p.c3_0 + cq);
                sc[q].dpp = p.dpooled + (long long)cq * p.ldp + (t0 >> ksh);
                sc[q].argp = p.arg + (long long)cq * p.ldp + (t0 >> ksh);
                sc[q].row_hi = sps + cq * 128;
                sc[q].prev[0] = sc[q].prev[1] = -1;
                sc[q].d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                sc[q].a4 = make_uchar4(0, 0, 0, 0);
                sc[q].dcur = 0.f;
                sc[q].acur = 0;
            }
            PROF_DECL(3)
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                __nv_bfloat16 vh[2], vl[2];
                int off[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    SpChannel& S = sc[q];
                    float val;
                    int r;
                    if (ksh == 0) {
                        if ((it & 3) == 0) {
                            if (it + 4 <= my_tiles) {
                                S.d4 = __ldg(reinterpret_cast<const float4*>(S.dpp + it));
                                S.a4 = __ldg(reinterpret_cast<const uchar4*>(S.argp + it));
                            } else {
                                S.d4.x = __ldg(S.dpp + it);
                                S.a4.x = __ldg(S.argp + it);
                                if (it + 1 < my_tiles) { S.d4.y = __ldg(S.dpp + it + 1); S.a4.y = __ldg(S.argp + it + 1); }
                                if (it + 2 < my_tiles) { S.d4.z = __ldg(S.dpp + it + 2); S.a4.z = __ldg(S.argp + it + 2); }
                            }
                        }
                        const int sel = it & 3;
                        val = S.k0 * (sel == 0 ? S.d4.x : sel == 1 ? S.d4.y : sel == 2 ? S.d4.z : S.d4.w);
                        r = (sel == 0 ? S.a4.x : sel == 1 ? S.a4.y : sel == 2 ? S.a4.z : S.a4.w) & 63;
                    } else {
                        if ((it & kmask) == 0) {
                            S.dcur = __ldg(S.dpp + (it >> ksh));
                            S.acur = __ldg(S.argp + (it >> ksh));
                        }
                        r = S.acur & 63;
                        val = ((S.acur >> 6) == (it & kmask)) ? S.k0 * S.dcur : 0.f;
                    }
                    vh[q] = __float2bfloat16_rn(val);
                    vl[q] = __float2bfloat16_rn(val - __bfloat162float(vh[q]));
                    const int cq = c + 64 * q;
                    off[q] = ((((r >> 3) ^ (cq & 7)) << 4) | ((r & 7) << 1));
                }
                PROF_MARK(0)
                const int ss = (nhl == 2) ? 0 : (it & 1);                       // Sp stage (bf16 mode: two)
                mbar_wait(&sp_empty[ss], ((nhl == 2) ? (it & 1) : ((it >> 1) & 1)) ^ 1);
                PROF_MARK(1)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    SpChannel& S = sc[q];
                    uint8_t* row = S.row_hi + ss * 32768;
                    const int pv = ss ? S.prev[1] : S.prev[0];
                    if (pv >= 0) {
                        *reinterpret_cast<unsigned short*>(row + pv) = 0;
                        if (nhl == 2) *reinterpret_cast<unsigned short*>(row + 32768 + pv) = 0;
                    }
                    *reinterpret_cast<unsigned short*>(row + off[q]) = __bfloat16_as_ushort(vh[q]);
                    if (nhl == 2) *reinterpret_cast<unsigned short*>(row + 32768 + off[q]) = __bfloat16_as_ushort(vl[q]);
                    if (ss) S.prev[1] = off[q]; else S.prev[0] = off[q];
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sp_full[ss]);
                PROF_MARK(2)
            }
#ifdef FACL_PROFILE_ROLES
            if (blockIdx.x == 1 && warp == 0 && lane == 0)
                printf("pass C Sp producer, cycles/tile: loads+convert %lld | wait sp_empty %lld | write+fence+arrive %lld\n", prof_[0] / my_tiles,
                       prof_[1] / my_tiles, prof_[2] / my_tiles);
#endif
        }
        // dW3s of this CTA's rows sits in TMEM: add it to the global gradient
        mbar_wait(fin_bar, 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            float a[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(256 + 64 * h + q * 32), a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(p.dw3 + c * 64 + q * 32 + i, a[i]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 18) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// pass D   (14 warps)
// warps 0-7   : dh1 consumers: TMEM lane quarter = warp % 4, column half = warp / 4; the warps of quarters q and q + 2 (A_hi / A_lo products
//               of the same channels) swap half of their rows through shared memory and each finishes 16 complete rows
// warps 8-11  : h1 image producers: relu(z1') read from tensor memory (layer 0 is one K = 16 instruction per tile); warps 8, 9 also write
//               the fp32 x rows (for the consumers) and the K = 16 operand rows of the next tile
// warp 12     : MMA issuer          warp 13 : bulk-TMA loader of dh2' tiles
// TMEM columns: DH1[b] 0..127 / 128..255 (column halves = B_hi / B_lo products), [dW2s ; H1] 256, z1'[b] 320 / 384,
//               A tiles (E0 W2)^T 448, P2 480
//
// Every product here has only 64 real output rows, so the M = 128 instruction is fed STACKED operands instead of padding:
//  * a weight image is stored [hi 64 rows | lo 64 rows]; read as one 128-row A tile, lanes 0..63 receive A_hi * B and lanes
//    64..127 A_lo * B -- the bf16x3 scheme needs two instructions (B_hi, B_lo) instead of three, and because everything
//    the consumers derive from dh1 is a SUM over rows, the two lane halves could be reduced as separate partials; since round 2 the two
//    warps that hold the halves of the same channels swap half of their rows instead, and each finishes 16 complete rows;
//  * dh2' and h1 of a tile sit next to each other in a stage ([dh2'_hi | h1_hi | dh2'_lo | h1_lo]), so one 128-row A tile
//    against B = h1 yields dh2' h1^T (lanes 0..63) and h1 h1^T (lanes 64..127) at once.
// --------------------------------------------------------------------------------------------------------------------
constexpr int D_THREADS = 14 * 32;
constexpr int D_STAGES = 3;
constexpr int D_STAGE_BYTES = 4 * IMG64;
constexpr int D_XS = 4;                   // x ring (fp32 rows for the consumers + K = 16 rows for the z1' instruction): deeper than the
                                          // stage ring, because x of tile t+1 is written while tile t-4's consumers may still read theirs

__global__ void __launch_bounds__(D_THREADS, 1) l1_bwd_d_kernel(const L1BwdParams p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS, not generic LD/ST)
    const int nhl = p.nhl;
    uint8_t* ews = smem;                       // diag(e0) W2: hi 8 KB | lo 8 KB
    uint8_t* p2s = ews + 16384;                // P2:          hi 8 KB | lo 8 KB
    uint8_t* stg = p2s + 16384;                // D_STAGES x [dh2'_hi | h1_hi | dh2'_lo | h1_lo]
    uint8_t* w1a = stg + D_STAGES * D_STAGE_BYTES;  // layer-0 operand: 128 rows (channel, twice) x K = 16, see below
    uint8_t* x16s = w1a + 16384;               // D_XS x 64 rows x K = 16 (128-byte rows, two chunks used)
    uint8_t* xs = x16s + D_XS * IMG64;         // D_XS x 64 rows x 16 B
    uint8_t* exch = xs + D_XS * BT * 16;       // row swap between the hi / lo consumer warps: 4 pairs x 2 buffers x 2 directions x 2 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(exch + 32768);
    uint64_t *h1_full = bars, *in_full = bars + 3, *st_empty = bars + 6, *dh_full = bars + 9, *dh_empty = bars + 11,
             *w_bar = bars + 13, *fin_bar = bars + 14, *d1_full = bars + 15, *d1_empty = bars + 17, *x_free = bars + 19,
             *x16_full = bars + 19 + D_XS, *img_free = bars + 19 + 2 * D_XS, *a_ready = bars + 22 + 2 * D_XS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23 + 2 * D_XS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / BT;
    const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
    const int my_tiles = (int)((ntiles - t0 < p.tiles_per_cta) ? (ntiles - t0) : p.tiles_per_cta);

    if (threadIdx.x == 0) {
        for (int i = 0; i < D_STAGES; ++i) {
            mbar_init(&h1_full[i], 4);
            mbar_init(&in_full[i], 1);
            mbar_init(&st_empty[i], 1);
            mbar_init(&img_free[i], 8);
        }
        for (int i = 0; i < D_XS; ++i) {
            mbar_init(&x_free[i], 8);
            mbar_init(&x16_full[i], 2);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&dh_full[i], 1);
            mbar_init(&dh_empty[i], 8);
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 4);
        }
        mbar_init(w_bar, 1);
        mbar_init(fin_bar, 1);
        mbar_init(a_ready, 4);
        mbar_fence_init();
    }
    if (warp == 12) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    if (warp >= 8 && warp < 12) {
        // Layer 0 on the tensor pipe: z1' = s1 (W1 x + b1) + t1 is ONE K = 16 instruction per tile.  With W = s1 W1 and b = s1 b1 + t1
        // split into bf16 hi + lo, the operand rows are
        //     A[ch] = [ Wh0..3 | Wh0..3 | Wl0..3 | bh bl 0 0 ]        B[row] = [ xh0..3 | xl0..3 | xh0..3 | 1 1 0 0 ]
        // so that A.B = Wh.xh + Wh.xl + Wl.xh + bh + bl: the same three error-compensated products as every other layer, in both
        // precision modes.  Rows 64..127 of A repeat rows 0..63: TMEM lanes 64..127 then hold a second copy of z1', which the producer
        // warps of the upper lane quarters read.
        const int row = (warp - 8) * 32 + lane, ch = row & 63;
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float v[8] = {s1 * w.x, s1 * w.y, s1 * w.z, s1 * w.w, fmaf(s1, __ldg(p.b1 + ch), t1), 0.f, 0.f, 0.f};
        uint4 h, l;
        split_bf16x8(v, h, l);
        // h = [Wh0 Wh1 | Wh2 Wh3 | bh 0 | 0 0], l likewise
        *reinterpret_cast<uint4*>(w1a + sw128_offset((uint32_t)row, 0)) = make_uint4(h.x, h.y, h.x, h.y);
        *reinterpret_cast<uint4*>(w1a + sw128_offset((uint32_t)row, 1)) = make_uint4(l.x, l.y, (h.z & 0xFFFFu) | (l.z << 16), 0u);
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc_mn = umma_idesc_bf16(128, BT) | UMMA_B_MN_MAJOR;
    const uint32_t idesc_dg = umma_idesc_bf16(128, BT) | UMMA_A_MN_MAJOR | UMMA_B_MN_MAJOR;
    const uint32_t idesc_kk = umma_idesc_bf16(128, 64);
    // B images are [hi | lo] (2 * IMG64 apart): one N = 128 instruction gives the products with both halves (columns 0..63 / 64..127)
    const int nB = (nhl == 2) ? 2 * BT : BT;
    const uint32_t idesc_mn2 = umma_idesc_bf16(128, nB) | UMMA_B_MN_MAJOR;
    (void)idesc_mn; (void)idesc_dg;

    // The two constant A operands of the dh1 product, (E0 W2)^T and P2 (each [hi 64 rows ; lo 64 rows] x K = 64), live in tensor memory
    // (columns 448 / 480): this pass is bound by shared-memory bandwidth -- per 64-row tile the instructions fetched 136 KB of operands
    // next to ~90 KB of CUDA-core traffic at 128 B / cycle -- and an A tile in TMEM is not fetched at all.
    const uint32_t ew_t = tmem_base + 448, p2_t = tmem_base + 480;
    if (warp < 4) {
        mbar_wait(w_bar, 0);
        const int m = warp * 32 + lane, ci = m & 63, hl = m >> 6;
        const bool real = (hl == 0) || (nhl == 2);                 // bf16 mode: no lo rows, lanes 64..127 get zeros
        uint32_t u[32];
        const uint8_t* img = ews + hl * 8192;                      // stored [j][i]; the A tile wants lane = i, K = j
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) {
            uint32_t v = 0u;
            if (real) {
                const uint32_t e0 = *reinterpret_cast<const unsigned short*>(img + sw128_offset((uint32_t)(2 * j2), (uint32_t)(ci >> 3)) + (ci & 7) * 2);
                const uint32_t e1 = *reinterpret_cast<const unsigned short*>(img + sw128_offset((uint32_t)(2 * j2 + 1), (uint32_t)(ci >> 3)) + (ci & 7) * 2);
                v = e0 | (e1 << 16);
            }
            u[j2] = v;
        }
        tmem_st32(ew_t + ((uint32_t)(warp * 32) << 16), u);
        tmem_put_a_tile(p2_t + ((uint32_t)(warp * 32) << 16), real ? p2s + hl * 8192 : nullptr, ci);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
    }
    if (warp == 12) {
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(w_bar, 2 * 8192u * nhl);
            tma_bulk_g2s(ews, p.e0w2_img, 8192, w_bar);
            tma_bulk_g2s(p2s, p.p2_img, 8192, w_bar);
            if (nhl == 2) {
                tma_bulk_g2s(ews + 8192, p.e0w2_img + 16384, 8192, w_bar);
                tma_bulk_g2s(p2s + 8192, p.p2_img + 16384, 8192, w_bar);
            }
            mbar_wait(w_bar, 0);
            mbar_wait(a_ready, 0);
            tc_fence_after_sync();
            const bool split = nhl == 2;
            const UDesc stg_mn = udesc_mn(smem_u32(stg), 2 * IMG64, 1024);           // a stage image with its lo half as the second N block
            const UDesc stg_k = udesc_k(smem_u32(stg));
            constexpr uint32_t STG = D_STAGE_BYTES / 16, I64 = IMG64 / 16;
            const UDesc w1a_k = udesc_k(smem_u32(w1a));
            const UDesc x16_k = udesc_k(smem_u32(x16s));
            int s = 0, ph = 0;
            int sa = 0, pha = 0;                      // stage / phase of the tile whose z1' is issued (one tile ahead)
            // z1'(tile j) -> TMEM columns 320 + 64 (j & 1): lanes = channel (twice), columns = the 64 rows
            auto issue_z1 = [&](int j) {
                mbar_wait(&x16_full[sa], pha);
                mbar_wait(&d1_empty[j & 1], ((j >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                umma_ss(tmem_base + 320 + 64 * (j & 1), w1a_k, 0, x16_k, (uint32_t)sa * I64, idesc_kk, 0u);
                umma_commit(&d1_full[j & 1]);
                if (++sa == D_XS) { sa = 0; pha ^= 1; }
            };
            if (my_tiles > 0) issue_z1(0);
#ifdef FACL_PROFILE_ROLES
            long long pw_h1 = 0, pw_in = 0, pw_dh = 0, pw_issue = 0, pt0 = clock64();
#endif
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1, u = (it >> 1) & 1;
                if (it + 1 < my_tiles) issue_z1(it + 1);
#ifdef FACL_PROFILE_ROLES
                long long c0 = clock64();
#endif
                mbar_wait(&h1_full[s], ph);
#ifdef FACL_PROFILE_ROLES
                long long c1 = clock64();
#endif
                mbar_wait(&in_full[s], ph);
#ifdef FACL_PROFILE_ROLES
                long long c2 = clock64();
#endif
                mbar_wait(&dh_empty[b], u ^ 1);
#ifdef FACL_PROFILE_ROLES
                long long c3 = clock64();
                pw_h1 += c1 - c0; pw_in += c2 - c1; pw_dh += c3 - c2;
#endif
                tc_fence_after_sync();
                const uint32_t so = (uint32_t)s * STG;            // stage: [dh2'_hi | h1_hi | dh2'_lo | h1_lo], 8 KB each
                // dh1[i][r] = sum_j (e0 W2)[j][i] dh2'[j][r] + sum_i' P2[i][i'] h1[i'][r];  lanes 0..63 hi part, 64..127 lo part
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(tmem_base + 128 * b, ew_t + ks * 8, stg_mn, so + ks * 128, idesc_mn2, ks > 0 ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_ts(tmem_base + 128 * b, p2_t + ks * 8, stg_mn, so + I64 + ks * 128, idesc_mn2, 1u);
                umma_commit(&dh_full[b]);
                // [dW2s ; H1] += [dh2' ; h1] h1^T  (reduction over the 64 rows): A = the 128-row tile at the stage start, B = h1
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_ss(tmem_base + 256, stg_k, so + ks * 2, stg_k, so + I64 + ks * 2, idesc_kk, (it == 0 && ks == 0) ? 0u : 1u);
                    if (split) {
                        umma_ss(tmem_base + 256, stg_k, so + ks * 2, stg_k, so + 3 * I64 + ks * 2, idesc_kk, 1u);
                        umma_ss(tmem_base + 256, stg_k, so + 2 * I64 + ks * 2, stg_k, so + I64 + ks * 2, idesc_kk, 1u);
                    }
                }
                umma_commit(&st_empty[s]);
                if (++s == D_STAGES) { s = 0; ph ^= 1; }
#ifdef FACL_PROFILE_ROLES
                pw_issue += clock64() - c3;
#endif
            }
#ifdef FACL_PROFILE_ROLES
            if (blockIdx.x == 0)      // measured: waits 68 / 51 / 94, issue 2013 of 2396 cycles per tile -- the issuing thread is back-pressured by the tensor pipe
                printf("pass D mma warp: tiles %d, per tile: wait h1 %lld, wait dh2' TMA %lld, wait dh1 buffer %lld, issue %lld, total %lld\n",
                       my_tiles, pw_h1 / my_tiles, pw_in / my_tiles, pw_dh / my_tiles, pw_issue / my_tiles, (clock64() - pt0) / my_tiles);
#endif
            umma_commit(fin_bar);
        }
    } else if (warp == 13) {
        if (elect_one_sync()) {
            int s = 0, ph = 0;
            const uint8_t* src = p.dh2 + t0 * (2 * IMG64);
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&st_empty[s], ph ^ 1);
                uint8_t* dst = stg + s * D_STAGE_BYTES;
                mbar_arrive_expect_tx(&in_full[s], (uint32_t)(IMG64 * nhl));
                tma_bulk_g2s(dst, src + (long long)it * (2 * IMG64), IMG64, &in_full[s]);
                if (nhl == 2) tma_bulk_g2s(dst + 2 * IMG64, src + (long long)it * (2 * IMG64) + IMG64, IMG64, &in_full[s]);
                if (++s == D_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ---- h1 producers: x tile (fp32 for the consumers, K = 16 bf16 rows for the z1' instruction) one tile ahead; then
        //      h1 = relu(z1') from TMEM -> bf16 hi / lo image of the stage ----
        const int ptid = (warp - 8) * 32 + lane, ch = ptid & 63, half = ptid >> 6;
        const float4* xg = reinterpret_cast<const float4*>(p.xt) + t0 * BT;
        float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ptid < BT && my_tiles > 0) xnext = __ldg(xg + ptid);
        float hsum = 0.f;
        int s = 0, ph = 0;
        int sa = 0, pha = 0;
        auto put_x = [&](int j) {                                 // warps 8, 9: rows of tile j -> slot sa
            mbar_wait(&x_free[sa], pha ^ 1);                      // the dh1 consumers of the tile that used this x slot are done
            reinterpret_cast<float4*>(xs + sa * BT * 16)[ptid] = xnext;
            const float xv[8] = {xnext.x, xnext.y, xnext.z, xnext.w, 1.f, 0.f, 0.f, 0.f};
            uint4 h, l;
            split_bf16x8(xv, h, l);
            uint8_t* x16 = x16s + sa * IMG64;
            *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)ptid, 0)) = make_uint4(h.x, h.y, l.x, l.y);
            *reinterpret_cast<uint4*>(x16 + sw128_offset((uint32_t)ptid, 1)) = make_uint4(h.x, h.y, h.z | (h.z << 16), 0u);
            if (j + 1 < my_tiles) xnext = __ldg(xg + (long long)(j + 1) * BT + ptid);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x16_full[sa]);
            if (++sa == D_XS) { sa = 0; pha ^= 1; }
        };
        if (ptid < BT && my_tiles > 0) put_x(0);
        PROF_DECL(5)
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            if (ptid < BT && it + 1 < my_tiles) put_x(it + 1);
            PROF_MARK(0)
            mbar_wait(&d1_full[b], u);
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(320 + 64 * b + half * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[b]);
            PROF_MARK(1)
            mbar_wait(&st_empty[s], ph ^ 1);                      // the Gram instructions that read this stage are done
            mbar_wait(&img_free[s], ph ^ 1);                      // ... and so are the consumers that take the ReLU1 mask from its h1 image
            PROF_MARK(2)
            // h1 image of this stage: hi at +IMG64, lo at +3*IMG64
            float acc = 0.f;
            uint8_t* img = stg + s * D_STAGE_BYTES + IMG64;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[e] = fmaxf(z[q * 8 + e], 0.f);
                    acc += v[e];
                }
                store_img8(img, nhl, 2 * IMG64, ch, half * 4 + q, v);
            }
            hsum += acc;
            PROF_MARK(3)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[s]);
            PROF_MARK(4)
            if (p.dbg_mask1) {
#pragma unroll
                for (int r = 0; r < 32; ++r) p.dbg_mask1[((t0 + it) * 64 + ch) * 64 + half * 32 + r] = z[r] > 0.f ? 1 : 0;
            }
            if (++s == D_STAGES) { s = 0; ph ^= 1; }
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 1 && ptid == 0)
            printf("pass D h1 producer, cycles/tile: x tile ahead %lld | wait d1_full + ld %lld | wait st_empty %lld | produce %lld | fence+arrive %lld\n",
                   prof_[0] / my_tiles, prof_[1] / my_tiles, prof_[2] / my_tiles, prof_[3] / my_tiles, prof_[4] / my_tiles);
#endif
        atomicAdd(p.hsum + ch, hsum);
    } else {
        // ---- dh1 consumers (thread = channel i): + q2, ReLU1 mask from the h1 image, BN1 backward sums, A = sum dh1' x^T.
        //      The A_hi products of a channel sit in TMEM lanes 0..63 and the A_lo products in lanes 64..127, i.e. in DIFFERENT warps
        //      (quarter q and q + 2).  The two warps of such a pair swap half of their 32 rows through shared memory, so each one
        //      holds the COMPLETE dh1 of 16 rows and does the per-element work only for those (in bf16 mode, where there is no lo
        //      part, the upper warp simply takes over half of the lower warp's rows). ----
        const int quarter = warp & 3, colhalf = warp >> 2;
        const int part_rt = quarter >> 1;                        // 0: lanes 0..63 (A_hi products), 1: lanes 64..127 (A_lo)
        const int i = (quarter & 1) * 32 + lane;
        const int pair = (quarter & 1) + 2 * colhalf;
        const float q2 = __ldg(p.q2 + i);
        // (bf16 mode has no lo part: lanes 64..127 are meaningless and are not read)
        float s_acc = 0.f, ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        int s = 0, sx = 0;
        PROF_DECL(7)
        // the loop is instantiated once per lane half: with `part` a compile-time constant the row selection below is register naming
        auto run = [&](auto part_c) {
        constexpr int part = decltype(part_c)::value;
        const bool have = (part == 0) || (nhl == 2);
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            const float4* xtile = reinterpret_cast<const float4*>(xs + sx * BT * 16) + colhalf * 32 + part * 16;
            const uint8_t* h1img = stg + s * D_STAGE_BYTES + IMG64;      // bf16(h1) of this tile: non-zero <=> ReLU1 active
            uint8_t* ex = exch + pair * 8192 + (it & 1) * 4096;          // [direction][4 row groups][32 lanes] float4, double-buffered
            mbar_wait(&dh_full[b], u);
            PROF_MARK(0)
            tc_fence_after_sync();
            // Every shared-memory load below is issued well before its first use (the ReLU1 bits while the accumulator loads are in
            // flight, the first eight x rows before the row swap, the other eight before the first eight are consumed): with one exposed
            // LDS round trip per row the consumers spent 40 % of their time on the short scoreboard (ncu source page).
            float g[32], gl[32];
            if (have) {
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(128 * b + colhalf * 32), g);
                if (nhl == 2) tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(128 * b + 64 + colhalf * 32), gl);
            }
            const uint4 hb0 = *reinterpret_cast<const uint4*>(h1img + sw128_offset((uint32_t)i, (uint32_t)(colhalf * 4 + part * 2)));
            const uint4 hb1 = *reinterpret_cast<const uint4*>(h1img + sw128_offset((uint32_t)i, (uint32_t)(colhalf * 4 + part * 2 + 1)));
            float4 xq[4][4];                                              // x rows in groups of four, one group ahead of its use
            if (have) {
                tmem_ld_wait();
                if (nhl == 2) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) g[r] += gl[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) g[r] = 0.f;
            }
            tc_fence_before_sync();
            PROF_MARK(3)
#pragma unroll
            for (int e = 0; e < 4; ++e) xq[0][e] = xtile[e];
            __syncwarp();
            if (lane == 0) mbar_arrive(&dh_empty[b]);                     // the accumulator is in registers now
            // rows 0..15 of the column half belong to the lower warp, rows 16..31 to the upper one
            if (have) {
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const int k = r4 * 4;
                    const float4 snd = part ? make_float4(g[k], g[k + 1], g[k + 2], g[k + 3])
                                            : make_float4(g[16 + k], g[17 + k], g[18 + k], g[19 + k]);
                    *reinterpret_cast<float4*>(ex + part * 2048 + (r4 * 32 + lane) * 16) = snd;
                }
            }
            float gv[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) gv[r] = part ? g[16 + r] : g[r];
            PROF_MARK(4)
            named_bar_sync(2 + pair, 64);
            PROF_MARK(5)
            float4 rc[4];
            const bool recv = (part == 1) || (nhl == 2);
            if (recv) {
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) rc[r4] = *reinterpret_cast<const float4*>(ex + (1 - part) * 2048 + (r4 * 32 + lane) * 16);
            }
            if (recv) {
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    gv[r4 * 4] += rc[r4].x; gv[r4 * 4 + 1] += rc[r4].y; gv[r4 * 4 + 2] += rc[r4].z; gv[r4 * 4 + 3] += rc[r4].w;
                }
            }
            PROF_MARK(1)
            const uint32_t hw[8] = {hb0.x, hb0.y, hb0.z, hb0.w, hb1.x, hb1.y, hb1.z, hb1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < 3) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) xq[j + 1][e] = xtile[(j + 1) * 4 + e];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = j * 4 + e;
                    const uint32_t bits = (r & 1) ? (hw[r >> 1] >> 16) : (hw[r >> 1] & 0xFFFFu);
                    const float4 x = xq[j][e];
                    const float v = bits ? gv[r] + q2 : 0.f;
                    s_acc += v;
                    ax = fmaf(v, x.x, ax); ay = fmaf(v, x.y, ay); az = fmaf(v, x.z, az); aw = fmaf(v, x.w, aw);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&x_free[sx]);
                mbar_arrive(&img_free[s]);
            }
            PROF_MARK(2)
            if (++s == D_STAGES) s = 0;
            if (++sx == D_XS) sx = 0;
        }
        };
        if (part_rt == 0) run(std::integral_constant<int, 0>{}); else run(std::integral_constant<int, 1>{});
        const int part = part_rt;
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 1 && (warp == 0 || warp == 2) && lane == 0)
            printf("pass D dh1 consumer (quarter %d), cycles/tile: wait dh_full %lld | TMEM loads %lld | x prefetch + send %lld | pair barrier %lld | receive %lld | mask+sums+arrive %lld\n", quarter,
                   prof_[0] / my_tiles, prof_[3] / my_tiles, prof_[4] / my_tiles, prof_[5] / my_tiles, prof_[1] / my_tiles, prof_[2] / my_tiles);
#endif
        // z1 = w.x + b1 is affine in x, so sum dh1' z1 follows from A = sum dh1' x^T and sum dh1'
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + i);
        const float b1 = __ldg(p.b1 + i);
        const float q_acc = fmaf(w.x, ax, fmaf(w.y, ay, fmaf(w.z, az, fmaf(w.w, aw, b1 * s_acc))));
        const long long slot = (long long)(blockIdx.x * 4 + part * 2 + colhalf) * 64 + i;
        p.stats[slot * 2 + 0] = s_acc;
        p.stats[slot * 2 + 1] = q_acc;
        reinterpret_cast<float4*>(p.amat)[slot] = make_float4(ax, ay, az, aw);
        // [dW2s ; H1] of this CTA's rows: lanes 0..63 -> dW2s rows, lanes 64..127 -> H1 rows
        mbar_wait(fin_bar, 0);
        tc_fence_after_sync();
        float a[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + colhalf * 32), a);
        tmem_ld_wait();
        float* dst = (part == 0 ? p.dw2s : p.gram) + i * 64 + colhalf * 32;
#pragma unroll
        for (int q = 0; q < 32; ++q) atomicAdd(dst + q, a[q]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// The 64x64 algebra around the passes (one small block each).
// l1_prep: P = W^T diag(d) W (64x64) and, optionally, E = diag(e0) W (C == 64), both as packed bf16 hi/lo operand images;
//          q = W^T (d (.) bias + c2).   W is [C][64].
// l1_fin : dW[c][j] (+)= d[c] (sum_j' W[c][j'] H[j'][j] + bias[c] s[j]) + c2[c] s[j] (+ e0[c] sparse[c][j])
// --------------------------------------------------------------------------------------------------------------------
// grid = 32 blocks (2 rows j each) x 256 threads = (16 slices of the reduction over c) x (2 rows) x (8 chunks of 8 columns)
constexpr int PREP_SLICES = 16, PREP_ROWS = 2, PREP_BLOCKS = 64 / PREP_ROWS;
__global__ void __launch_bounds__(256) l1_prep_kernel(const float* __restrict__ W, int C, const float* __restrict__ d,
                                                      const float* __restrict__ bias, const float* __restrict__ c2,
                                                      const float* __restrict__ e0, uint8_t* __restrict__ p_img, float* __restrict__ q,
                                                      uint8_t* __restrict__ e_img) {
    pdl_prologue();
    __shared__ double red[PREP_SLICES - 1][PREP_ROWS * 8][9];
    const int slice = threadIdx.x / (PREP_ROWS * 8), t = threadIdx.x % (PREP_ROWS * 8);
    const int j = blockIdx.x * PREP_ROWS + (t >> 3), chunk = t & 7;
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    double qa = 0.0;
    for (int c = slice; c < C; c += PREP_SLICES) {
        const double wj = (double)__ldg(W + c * 64 + j), dc = (double)__ldg(d + c);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + c * 64 + chunk * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + c * 64 + chunk * 8) + 1);
        const double wd = wj * dc;
        acc[0] += wd * w0.x; acc[1] += wd * w0.y; acc[2] += wd * w0.z; acc[3] += wd * w0.w;
        acc[4] += wd * w1.x; acc[5] += wd * w1.y; acc[6] += wd * w1.z; acc[7] += wd * w1.w;
        if (chunk == 0) qa += wj * (dc * (double)__ldg(bias + c) + (double)__ldg(c2 + c));
    }
    if (slice > 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) red[slice - 1][t][e] = acc[e];
        red[slice - 1][t][8] = qa;
    }
    __syncthreads();
    if (slice > 0) return;
#pragma unroll 1
    for (int sl = 0; sl < PREP_SLICES - 1; ++sl) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += red[sl][t][e];
        qa += red[sl][t][8];
    }
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (float)acc[e];
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    const uint32_t off = sw128_offset((uint32_t)j, (uint32_t)chunk);
    *reinterpret_cast<uint4*>(p_img + off) = hi;
    *reinterpret_cast<uint4*>(p_img + 16384 + off) = lo;
    if (chunk == 0) q[j] = (float)qa;
    if (e_img) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e0[j] * W[j * 64 + chunk * 8 + e];
        split_bf16x8(v, hi, lo);
        *reinterpret_cast<uint4*>(e_img + off) = hi;
        *reinterpret_cast<uint4*>(e_img + 16384 + off) = lo;
    }
}

__global__ void __launch_bounds__(256) l1_fin_kernel(const float* __restrict__ W, int C, const float* __restrict__ d,
                                                     const float* __restrict__ bias, const float* __restrict__ c2,
                                                     const float* __restrict__ H, const float* __restrict__ s,
                                                     const float* __restrict__ e0, const float* __restrict__ sparse,
                                                     float* __restrict__ dW, int accumulate) {
    pdl_prologue();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= C * 64) return;
    const int c = idx >> 6, j = idx & 63;
    double acc = 0.0;
    for (int k = 0; k < 64; ++k) acc += (double)W[c * 64 + k] * (double)H[k * 64 + j];
    double out = (double)d[c] * (acc + (double)bias[c] * (double)s[j]) + (double)c2[c] * (double)s[j];
    if (sparse) out += (double)e0[c] * (double)sparse[idx];
    dW[idx] = accumulate ? dW[idx] + (float)out : (float)out;
}

// dW1[i][:] = c0_i * A_i + c1_i * (C w1_i + b1_i * sx) + c2_i * sx   with C = sum x x^T, sx = sum x (z1 is affine in x)
__global__ void __launch_bounds__(1024) l1_dw1_kernel(const float* __restrict__ amat, int P, const double* __restrict__ mom,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ c0, const float* __restrict__ c1,
                                                      const float* __restrict__ c2, float* __restrict__ dw1) {
    pdl_prologue();
    // 16 slices of the P partials x 64 channels, reduced through shared memory
    __shared__ double red[16][64][4];
    const int i = threadIdx.x & 63, slice = threadIdx.x >> 6;
    double a[4] = {0, 0, 0, 0};
    for (int q = slice; q < P; q += 16) {
        float4 v = __ldg(reinterpret_cast<const float4*>(amat) + (long long)q * 64 + i);
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) red[slice][i][d] = a[d];
    __syncthreads();
    if (slice > 0) return;
    for (int sl = 1; sl < 16; ++sl)
#pragma unroll
        for (int d = 0; d < 4; ++d) a[d] += red[sl][i][d];
    const int idx[4][4] = {{4, 5, 6, 7}, {5, 8, 9, 10}, {6, 9, 11, 12}, {7, 10, 12, 13}};
    for (int d = 0; d < 4; ++d) {
        double zx = (double)b1[i] * mom[d];
        for (int e = 0; e < 4; ++e) zx += (double)w1[i * 4 + e] * mom[idx[e][d]];
        dw1[i * 4 + d] = (float)((double)c0[i] * a[d] + (double)c1[i] * zx + (double)c2[i] * mom[d]);
    }
}

// Degenerate BatchNorm weight.  Pass C derives sum dh2' (z2 + b2) from the h2 image as q / a2 + b2 sum dh2' (a2 = gamma2 rstd2), which is
// undefined for a channel whose gamma2 is EXACTLY zero (h2 is then constant and carries no z2).  Such a channel gets 0 from pass C and
// its sum from here: z2 recomputed from the 16-byte input rows on CUDA cores, dh2' read back from the image pass C wrote.  One block; it
// returns after reading the 64 scales when no channel is degenerate (the normal case).
__global__ void __launch_bounds__(1024) l1_gamma0_fix_kernel(const float4* __restrict__ xt, long long R, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, const float* __restrict__ scale1,
                                                             const float* __restrict__ shift1, const float* __restrict__ w2,
                                                             const float* __restrict__ b2, const float* __restrict__ scale2,
                                                             const uint8_t* __restrict__ dh2, int nhl, float* __restrict__ stats) {
    pdl_prologue();
    __shared__ int any;
    __shared__ double red[32];
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    if (threadIdx.x < 64 && scale2[threadIdx.x] == 0.f) any = 1;
    __syncthreads();
    if (!any) return;
    for (int j = 0; j < 64; ++j) {
        if (scale2[j] != 0.f) continue;
        double acc = 0.0;
        for (long long r = threadIdx.x; r < R; r += 1024) {
            const float4 x = __ldg(xt + r);
            float z = 0.f;
            for (int i = 0; i < 64; ++i) {
                const float s1 = scale1[i];
                const float h = fmaxf(fmaf(s1 * w1[4 * i], x.x, fmaf(s1 * w1[4 * i + 1], x.y, fmaf(s1 * w1[4 * i + 2], x.z,
                                      fmaf(s1 * w1[4 * i + 3], x.w, fmaf(s1, b1[i], shift1[i]))))), 0.f);
                z = fmaf(w2[j * 64 + i], h, z);
            }
            const long long tile = r >> 6;
            const int rr = (int)(r & 63);
            const uint8_t* at = dh2 + tile * (2 * IMG64) + sw128_offset((uint32_t)j, (uint32_t)(rr >> 3)) + (rr & 7) * 2;
            float g = __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(at)) << 16);
            if (nhl == 2) g += __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(at + IMG64)) << 16);
            acc += (double)g * (double)(z + b2[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 32; ++w) t += red[w];
            stats[(0 * 64 + j) * 2 + 1] += (float)t;          // partial 0 of channel j (pass C left 0 in every partial)
        }
        __syncthreads();
    }
}

size_t l1_bwd_c_smem() { return 16384 + 65536 + 16384 + 4 * IMG64 + 4 * IMG64 + 65536 + 256 + 1024; }
size_t l1_bwd_d_smem() { return 16384 + 16384 + D_STAGES * D_STAGE_BYTES + 16384 + D_XS * IMG64 + D_XS * BT * 16 + 32768 + 512 + 1024; }

// sum x (4) and sum x x^T (10 unique) over all rows, in double
__global__ void __launch_bounds__(256) l1_moments_kernel(const float4* __restrict__ xt, long long R, double* __restrict__ out) {
    pdl_prologue();
    double s[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) s[i] = 0.0;
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < R; r += (long long)gridDim.x * 256) {
        float4 x = __ldg(xt + r);
        double a = x.x, b = x.y, c = x.z, d = x.w;
        s[0] += a; s[1] += b; s[2] += c; s[3] += d;
        s[4] += a * a; s[5] += a * b; s[6] += a * c; s[7] += a * d;
        s[8] += b * b; s[9] += b * c; s[10] += b * d;
        s[11] += c * c; s[12] += c * d; s[13] += d * d;
    }
    __shared__ double sh[14][8];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        double v = s[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) sh[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}

// BN1 in closed form: z1 = W1 x + b1 is affine in x, so mean = w.mx + b, var = w^T Cov(x) w
__global__ void l1_bn1_kernel(const double* __restrict__ mom, double n, const float* __restrict__ w1, const float* __restrict__ b1,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
                              float eps, float momentum, int training, float* __restrict__ mean, float* __restrict__ rstd,
                              float* __restrict__ scale, float* __restrict__ shift) {
    pdl_prologue();
    int c = threadIdx.x;
    if (c >= 64) return;
    double mu, var;
    if (training) {
        double mx[4], cov[4][4];
        for (int i = 0; i < 4; ++i) mx[i] = mom[i] / n;
        const int idx[4][4] = {{4, 5, 6, 7}, {5, 8, 9, 10}, {6, 9, 11, 12}, {7, 10, 12, 13}};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) cov[i][k] = mom[idx[i][k]] / n - mx[i] * mx[k];
        double w[4];
        for (int i = 0; i < 4; ++i) w[i] = w1[c * 4 + i];
        mu = b1[c];
        for (int i = 0; i < 4; ++i) mu += w[i] * mx[i];
        var = 0.0;
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) var += w[i] * cov[i][k] * w[k];
        if (var < 0.0) var = 0.0;
        double unbiased = (n > 1.0) ? var * n / (n - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    } else {
        mu = running_mean[c];
        var = running_var[c];
    }
    double r = 1.0 / sqrt(var + (double)eps), g = gamma[c];
    mean[c] = (float)mu;
    rstd[c] = (float)r;
    scale[c] = (float)(g * r);
    shift[c] = (float)((double)beta[c] - mu * g * r);
}

size_t l1_smem_bytes(bool pass_b) {
    size_t b = 16384 + (pass_b ? 65536 : 16384) + 2 * 2 * ACT_BYTES + (pass_b ? 2 * 2 * ACT_BYTES : 0) + 2 * TILE * 16 + 256 +
               (pass_b ? 0 : 3 * 16384);
    return b + 1024;
}

}  // namespace

int l1_fused_grid(long long R) {
    long long tiles = R / TILE;
    return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}

int l1_moments_launch(const float* xt, long long R, double* mom14, cudaStream_t st) {
    FACL_CHECK(cudaMemsetAsync(mom14, 0, 14 * sizeof(double), st));
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_moments_kernel, dim3(kNumSMs * 4), dim3(256), 0, st, reinterpret_cast<const float4*>(xt), R, mom14));
    return (int)cudaGetLastError();
}

int l1_bn1_launch(const double* mom14, double n, const float* w1, const float* b1, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float eps, float momentum, int training, float* mean, float* rstd,
                  float* scale, float* shift, cudaStream_t st) {
    ScopedTimer timer(TAG_BN, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_bn1_kernel, dim3(1), dim3(64), 0, st, mom14, n, w1, b1, gamma, beta, running_mean, running_var, eps, momentum, training, mean, rstd,
                                    scale, shift));
    return (int)cudaGetLastError();
}

// pass A: statistics of z2 -> stats [2*grid][64][2];  pass B: pooled [256][ldp] (+ winners); stat_mode 1: statistics of z3 ->
// stats [grid][256][2] (0: none, eval); gram / hsum non-null: H2 [64][64] and s2 [64] accumulated with atomics (zeroed by the caller)
int l1_fwd_launch(bool pass_b, const float* xt, long long R, int K, int nsplit, const float* w1, const float* b1, const float* scale1,
                  const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                  const void* w3_img, const float* b3, const float* gamma3, int stat_mode, float* stats, float* gram, float* hsum,
                  float* pooled, unsigned char* pool_arg, long long ldp, cudaStream_t st) {
    if (R <= 0 || R % TILE != 0 || K <= 0 || (K & (K - 1)) || TILE % K != 0) return (int)cudaErrorInvalidValue;
    if (stat_mode < 0 || stat_mode > 1 || (!pass_b && stat_mode != 1) || (stat_mode == 1 && !stats) || ((gram == nullptr) != (hsum == nullptr)) ||
        (gram && (!pass_b || stat_mode != 1)))
        return (int)cudaErrorInvalidValue;
    static DeviceOnce configured;
    if (configured.need()) {
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(false)));
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<true, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(true)));
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(true)));
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(true)));
        configured.done();
    }
    L1Params p;
    p.xt = xt; p.R = R; p.K = K; p.nhl = (nsplit == 3) ? 2 : 1;
    p.w1 = w1; p.b1 = b1; p.scale1 = scale1; p.shift1 = shift1;
    p.w2_img = reinterpret_cast<const uint8_t*>(w2_img); p.b2 = b2; p.scale2 = scale2; p.shift2 = shift2;
    p.w3_img = reinterpret_cast<const uint8_t*>(w3_img); p.b3 = b3; p.gamma3 = gamma3;
    p.stats = stats; p.gram = gram; p.hsum = hsum; p.pooled = pooled; p.pool_arg = pool_arg; p.ldp = ldp;
    const int grid = l1_fused_grid(R);
    ScopedTimer timer(pass_b ? TAG_L1_PASS_B : TAG_L1_PASS_A, st);
    count_launch();
    if (!pass_b)
        FACL_LAUNCH_OK(launch_pdl(l1_fwd_kernel<false, 1, false>, dim3(grid), dim3(NTHREADS_A), l1_smem_bytes(false), st, p));
    else if (stat_mode == 0)
        FACL_LAUNCH_OK(launch_pdl(l1_fwd_kernel<true, 0, false>, dim3(grid), dim3(NTHREADS_B), l1_smem_bytes(true), st, p));
    else if (!gram)
        FACL_LAUNCH_OK(launch_pdl(l1_fwd_kernel<true, 1, false>, dim3(grid), dim3(NTHREADS_B), l1_smem_bytes(true), st, p));
    else
        FACL_LAUNCH_OK(launch_pdl(l1_fwd_kernel<true, 1, true>, dim3(grid), dim3(NTHREADS_B), l1_smem_bytes(true), st, p));
    return (int)cudaGetLastError();
}

static unsigned char *g_dbg_mask1 = nullptr, *g_dbg_mask2 = nullptr;
void l1_set_debug_dump(unsigned char* mask1, unsigned char* mask2) {
    g_dbg_mask1 = mask1; g_dbg_mask2 = mask2;
}

// backward passes: every CTA owns a contiguous range of tiles (a multiple of 4, so the Sp producers can fetch four
// groups' winners with one aligned vector load)
static int bwd_tiles_per_cta(long long R) {
    const long long tiles = R / BT;
    long long per = (tiles + kNumSMs - 1) / kNumSMs;
    per = (per + 3) & ~3LL;
    return (int)per;
}
int l1_bwd_grid(long long R) {
    const long long tiles = R / BT, per = bwd_tiles_per_cta(R);
    return (int)((tiles + per - 1) / per);
}

static void fill_bwd_common(L1BwdParams& p, const float* xt, long long R, int nsplit, const float* w1, const float* b1,
                            const float* scale1, const float* shift1) {
    memset(&p, 0, sizeof(p));
    p.xt = xt; p.R = R; p.nhl = (nsplit == 3) ? 2 : 1; p.tiles_per_cta = bwd_tiles_per_cta(R);
    p.w1 = w1; p.b1 = b1; p.scale1 = scale1; p.shift1 = shift1;
}

int l1_prep_launch(const float* W, int C, const float* d, const float* bias, const float* c2, const float* e0, void* p_img, float* q,
                   void* e_img, cudaStream_t st) {
    if (C <= 0 || (e_img && C != 64)) return (int)cudaErrorInvalidValue;
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_prep_kernel, dim3(PREP_BLOCKS), dim3(256), 0, st, W, C, d, bias, c2, e0, reinterpret_cast<uint8_t*>(p_img), q, reinterpret_cast<uint8_t*>(e_img)));
    return (int)cudaGetLastError();
}

int l1_fin_launch(const float* W, int C, const float* d, const float* bias, const float* c2, const float* H, const float* s,
                  const float* e0, const float* sparse, float* dW, int accumulate, cudaStream_t st) {
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_fin_kernel, dim3((C * 64 + 255) / 256), dim3(256), 0, st, W, C, d, bias, c2, H, s, e0, sparse, dW, accumulate));
    return (int)cudaGetLastError();
}

// pass C: dw3 is accumulated with atomics (zero-initialised by the caller); stats [4*grid][64][2]
int l1_bwd_c_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                    const void* w3_img, const void* p3_img, const float* q3, const unsigned char* arg, const float* dpooled,
                    long long ldp, const float* c3_0, void* dh2, float* dw3, float* stats, int K,
                    cudaStream_t st) {
    if (R <= 0 || R % BT != 0 || ldp % 4 != 0 || (K != 64 && K != 128 && K != 256) || R % K != 0) return (int)cudaErrorInvalidValue;
    static DeviceOnce configured;
    if (configured.need()) {
        FACL_CHECK(cudaFuncSetAttribute(l1_bwd_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_bwd_c_smem()));
        configured.done();
    }
    L1BwdParams p;
    fill_bwd_common(p, xt, R, nsplit, w1, b1, scale1, shift1);
    p.w2_img = reinterpret_cast<const uint8_t*>(w2_img); p.b2 = b2; p.scale2 = scale2; p.shift2 = shift2;
    p.w3_img = reinterpret_cast<const uint8_t*>(w3_img); p.p3_img = reinterpret_cast<const uint8_t*>(p3_img); p.q3 = q3;
    p.arg = arg; p.dpooled = dpooled; p.ldp = ldp; p.c3_0 = c3_0;
    p.kshift = K == 64 ? 0 : (K == 128 ? 1 : 2);
    p.dh2 = reinterpret_cast<uint8_t*>(dh2); p.dw3 = dw3; p.gram = nullptr; p.hsum = nullptr; p.stats = stats;
    p.dbg_mask2 = g_dbg_mask2;
    ScopedTimer timer(TAG_L1_PASS_C, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_bwd_c_kernel, dim3(l1_bwd_grid(R)), dim3(C_THREADS), l1_bwd_c_smem(), st, p));
    return (int)cudaGetLastError();
}

int l1_gamma0_fix_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1, const float* shift1,
                         const float* w2, const float* b2, const float* scale2, const void* dh2, float* stats, cudaStream_t st) {
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_gamma0_fix_kernel, dim3(1), dim3(1024), 0, st, reinterpret_cast<const float4*>(xt), R, w1, b1, scale1, shift1, w2,
                              b2, scale2, reinterpret_cast<const uint8_t*>(dh2), (nsplit == 3) ? 2 : 1, stats));
    return (int)cudaGetLastError();
}

// pass D: dw2s / gram / hsum accumulated with atomics (zero-initialised by the caller); stats [4*grid][64][2]; amat [4*grid][64][4]
int l1_bwd_d_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* e0w2_img, const void* p2_img, const float* q2, const void* dh2, float* dw2s,
                    float* gram, float* hsum, float* amat, float* stats, cudaStream_t st) {
    if (R <= 0 || R % BT != 0) return (int)cudaErrorInvalidValue;
    static DeviceOnce configured;
    if (configured.need()) {
        FACL_CHECK(cudaFuncSetAttribute(l1_bwd_d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_bwd_d_smem()));
        configured.done();
    }
    L1BwdParams p;
    fill_bwd_common(p, xt, R, nsplit, w1, b1, scale1, shift1);
    p.e0w2_img = reinterpret_cast<const uint8_t*>(e0w2_img); p.p2_img = reinterpret_cast<const uint8_t*>(p2_img); p.q2 = q2;
    p.dh2 = reinterpret_cast<uint8_t*>(const_cast<void*>(dh2)); p.dw2s = dw2s; p.gram = gram; p.hsum = hsum; p.amat = amat;
    p.stats = stats;
    p.dbg_mask1 = g_dbg_mask1;
    ScopedTimer timer(TAG_L1_PASS_D, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_bwd_d_kernel, dim3(l1_bwd_grid(R)), dim3(D_THREADS), l1_bwd_d_smem(), st, p));
    return (int)cudaGetLastError();
}

int l1_dw1_launch(const float* amat, int P, const double* mom14, const float* w1, const float* b1, const float* c0, const float* c1,
                  const float* c2, float* dw1, cudaStream_t st) {
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l1_dw1_kernel, dim3(1), dim3(1024), 0, st, amat, P, mom14, w1, b1, c0, c1, c2, dw1));
    return (int)cudaGetLastError();
}

}  // namespace facl
