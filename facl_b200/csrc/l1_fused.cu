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
#include <string.h>

#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"

namespace facl {

namespace {

constexpr int TILE = 128;                 // batch rows per tile
constexpr uint32_t ACT_LBO = 8192;        // MN-major activation image: 64-row blocks 8 KB apart,
constexpr uint32_t ACT_SBO = 1024;        //                            8-channel groups 1 KB apart (16 KB per half)
constexpr int ACT_BYTES = 16384;
constexpr int NTHREADS = 17 * 32;

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
    float* pooled;            // pass B: [256][ldp]  selected pre-BN value per (channel, group)
    long long ldp;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// write 8 consecutive rows of one channel into an MN-major activation image (hi / hi+lo)
__device__ __forceinline__ void store_act8(uint8_t* img, int nhl, int channel, int chunk, const float (&v)[8]) {
    uint32_t off = mn_sw128_offset((uint32_t)channel, (uint32_t)chunk, ACT_LBO, ACT_SBO);
    if (nhl == 2) {
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(img + off) = h;
        *reinterpret_cast<uint4*>(img + ACT_BYTES + off) = l;
    } else {
        *reinterpret_cast<uint4*>(img + off) = pack_bf16x8(v);
    }
}

// D[tmem] (+)= A (K-major weight image, 64-wide K) * B (MN-major activation image), all bf16 hi/lo combinations
__device__ __forceinline__ void mma_weight_act(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_img, int nhl, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = umma_desc_sw128(a_hi + ks * 32);
        const uint64_t bd = umma_desc_mn_sw128(b_img + ks * 2 * ACT_SBO, ACT_LBO, ACT_SBO);
        umma_bf16_ss(d_tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        if (nhl == 2) {
            umma_bf16_ss(d_tmem, ad, umma_desc_mn_sw128(b_img + ACT_BYTES + ks * 2 * ACT_SBO, ACT_LBO, ACT_SBO), idesc, 1u);
            umma_bf16_ss(d_tmem, umma_desc_sw128(a_lo + ks * 32), bd, idesc, 1u);
        }
    }
}

template <bool PASS_B>
__global__ void __launch_bounds__(NTHREADS, 1) l1_fwd_kernel(const L1Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nhl = p.nhl;
    // ---- shared-memory carve-up (all operand images 1 KB aligned) ----
    uint8_t* w2s = smem;                                   // hi 8 KB | lo 8 KB (64 valid rows each)
    uint8_t* w3s = w2s + 16384;                            // [half][hi 16 KB | lo 16 KB]      (pass B)
    uint8_t* h1s = w3s + (PASS_B ? 65536 : 16384);         // pass A: leave 16 KB readable behind W2 (M=128 reads 128 rows)
    uint8_t* h2s = h1s + 2 * 2 * ACT_BYTES;                // 2 stages x (hi|lo)
    uint8_t* xs = h2s + (PASS_B ? 2 * 2 * ACT_BYTES : 0);  // 2 stages x 128 rows x 16 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + 2 * TILE * 16);
    uint64_t *h1_full = bars, *h1_empty = bars + 2, *d2_full = bars + 4, *d2_empty = bars + 6, *h2_full = bars + 8,
             *h2_empty = bars + 10, *d3_full = bars + 12, *d3_empty = bars + 14, *w_bar = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / TILE;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h1_full[i], 4);
            mbar_init(&h1_empty[i], 1);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 4);
            mbar_init(&h2_full[i], 4);
            mbar_init(&h2_empty[i], 1);
            mbar_init(&d3_full[i], 1);
            mbar_init(&d3_empty[i], 4);
        }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 16) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: D2[b] at 128*b (b = 0,1); D3[half] at 256 + 128*half
    const uint32_t idesc = umma_idesc_bf16(128, TILE) | UMMA_B_MN_MAJOR;

    if (warp == 16) {
        // ======================= MMA issuer (one lane) =======================
        if (lane == 0) {
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
            const uint32_t w2_hi = smem_u32(w2s), w2_lo = w2_hi + 8192;
            auto issue_mma2 = [&](int it) {
                const int b = it & 1, u = (it >> 1) & 1;
                mbar_wait(&h1_full[b], u);
                mbar_wait(&d2_empty[b], u ^ 1);
                tc_fence_after_sync();
                mma_weight_act(tmem_base + 128 * b, w2_hi, w2_lo, smem_u32(h1s + b * 2 * ACT_BYTES), nhl, idesc);
                umma_commit(&h1_empty[b]);
                umma_commit(&d2_full[b]);
            };
            int it = 0;
            long long t = blockIdx.x;
            if (t < ntiles) issue_mma2(0);
            for (; t < ntiles; t += gridDim.x, ++it) {
                if (t + gridDim.x < ntiles) issue_mma2(it + 1);
                if (PASS_B) {
                    const int b = it & 1, u = (it >> 1) & 1;
                    mbar_wait(&h2_full[b], u);
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&d3_empty[h], (it & 1) ^ 1);
                        tc_fence_after_sync();
                        const uint32_t w3_hi = smem_u32(w3s + h * 32768);
                        mma_weight_act(tmem_base + 256 + 128 * h, w3_hi, w3_hi + 16384, smem_u32(h2s + b * 2 * ACT_BYTES), nhl, idesc);
                        umma_commit(&d3_full[h]);
                    }
                    umma_commit(&h2_empty[b]);
                }
            }
        }
    } else if (warp == 10 || warp == 11 || warp == 14 || warp == 15) {
        // ======================= producers: x -> h1 = relu(bn1(W1 x + b1)), thread = channel =======================
        const int pw = (warp == 10) ? 0 : (warp == 11) ? 1 : (warp == 14) ? 2 : 3;
        const int ptid = pw * 32 + lane;          // 0..127
        const int ch = ptid & 63, half = ptid >> 6;
        // BN1 folded into the 4-wide layer: h1 = max(wf . x + bf, 0)
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float wx = s1 * w.x, wy = s1 * w.y, wz = s1 * w.z, ww = s1 * w.w;
        const float bf = fmaf(s1, __ldg(p.b1 + ch), t1);
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            float4* xtile = reinterpret_cast<float4*>(xs + b * TILE * 16);
            xtile[ptid] = __ldg(reinterpret_cast<const float4*>(p.xt) + t * TILE + ptid);
            named_bar_sync(1, 128);
            mbar_wait(&h1_empty[b], u ^ 1);
            uint8_t* img = h1s + b * 2 * ACT_BYTES;
#pragma unroll 2
            for (int q = 0; q < 8; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float4 x = xtile[half * 64 + q * 8 + e];
                    v[e] = fmaxf(fmaf(wx, x.x, fmaf(wy, x.y, fmaf(wz, x.z, fmaf(ww, x.w, bf)))), 0.f);
                }
                store_act8(img, nhl, ch, half * 8 + q, v);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[b]);
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
        float s_acc = 0.f, q_acc = 0.f;
        long long nrows = 0;
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            mbar_wait(&d2_full[b], u);
            tc_fence_after_sync();
            if (PASS_B) mbar_wait(&h2_empty[b], u ^ 1);
            uint8_t* img = h2s + b * 2 * ACT_BYTES;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(128 * b + colhalf * 64 + q * 32), v);
                tmem_ld_wait();
                if (PASS_B) {
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        float h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) h[e] = fmaxf(fmaf(a2, v[g8 * 8 + e], c2), 0.f);
                        store_act8(img, nhl, j, colhalf * 8 + q * 4 + g8, h);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        s_acc += v[i];
                        q_acc = fmaf(v[i], v[i], q_acc);
                    }
                }
            }
            nrows += 64;
            tc_fence_before_sync();
            if (PASS_B) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (PASS_B) mbar_arrive(&h2_full[b]);
                mbar_arrive(&d2_empty[b]);
            }
        }
        if (!PASS_B) {
            // statistics of z2 = acc + b2 from the sums of acc
            const float n = (float)nrows;
            float* st = p.stats + ((long long)(blockIdx.x * 2 + colhalf) * 64 + j) * 2;
            st[0] = fmaf(n, b2, s_acc);
            st[1] = q_acc + 2.f * b2 * s_acc + n * b2 * b2;
        }
    } else if (warp < 8) {
        // ======================= z3 consumers (pass B), thread = channel c =======================
        if (PASS_B) {
            const int h = warp >> 2, lq = warp & 3;
            const int c = h * 128 + lq * 32 + lane;
            const float b3 = __ldg(p.b3 + c);
            const float sgn = (__ldg(p.gamma3 + c) >= 0.f) ? 1.f : -1.f;
            const int K = p.K, groups = TILE / K;
            float s_acc = 0.f, q_acc = 0.f;
            long long nrows = 0;
            int it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                mbar_wait(&d3_full[h], it & 1);
                tc_fence_after_sync();
                float best = -INFINITY;
#pragma unroll 1
                for (int q = 0; q < TILE / 32; ++q) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(256 + 128 * h + q * 32), v);
                    tmem_ld_wait();
                    if (K >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            s_acc += v[i];
                            q_acc = fmaf(v[i], v[i], q_acc);
                            best = fmaxf(best, v[i] * sgn);
                        }
                        if (((q + 1) * 32) % K == 0) {
                            const long long g = t * groups + (q * 32) / K;
                            p.pooled[(long long)c * p.ldp + g] = fmaf(best, sgn, b3);
                            best = -INFINITY;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            s_acc += v[i];
                            q_acc = fmaf(v[i], v[i], q_acc);
                            best = fmaxf(best, v[i] * sgn);
                            if (((i + 1) & (K - 1)) == 0) {
                                const long long g = t * groups + (q * 32 + i) / K;
                                p.pooled[(long long)c * p.ldp + g] = fmaf(best, sgn, b3);
                                best = -INFINITY;
                            }
                        }
                    }
                }
                nrows += TILE;
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d3_empty[h]);
            }
            const float n = (float)nrows;
            float* st = p.stats + ((long long)blockIdx.x * 256 + c) * 2;
            st[0] = fmaf(n, b3, s_acc);
            st[1] = q_acc + 2.f * b3 * s_acc + n * b3 * b3;
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
// tiles are 64 batch rows.  All images are [channel][64 rows] (see above): the same bytes serve as an MN-major B
// operand (reduction over channels: forward and data-gradient GEMMs) and as a K-major operand (reduction over rows:
// weight-gradient GEMMs).
//
//   pass C : x -> h1 -> z2 -> h2 -> z3;  dz3 = c0*dy3 + c1*z3 + c2 with dy3 = the pooled gradient placed at the max-pool
//            winner (first row whose z3 equals the pooled value);  dW3 += dz3 h2^T (TMEM-resident over the whole
//            kernel);  dh2 = W3^T dz3, masked by ReLU2 -> written to HBM (64 x R fp32, 64-row blocks) with the BN2
//            backward sums.
//   pass D : x -> h1 -> z2;  dz2 = c0*dh2' + c1*z2 + c2;  dW2 += dz2 h1^T (TMEM-resident);  dh1 = W2^T dz2 masked by
//            ReLU1 -> BN1 backward sums and A = sum dh1' x^T (64x4); nothing per-row is written.
//   dW1 follows in closed form from A and the input moments (z1 is affine in x).
// =====================================================================================================================
constexpr int BT = 64;                    // batch rows per backward tile
constexpr int IMG64 = 8192;               // [64 ch][64 rows] bf16 image (one half)
constexpr int BWD_THREADS = 17 * 32;

struct L1BwdParams {
    const float* xt;
    long long R;
    int K, nhl;
    const float* w1;  const float* b1;  const float* scale1;  const float* shift1;
    const uint8_t* w2_img;  const float* b2;  const float* scale2;  const float* shift2;
    const uint8_t* w3_img;  const float* b3;
    // pass C
    const float* pooled;      // [256][ldp] forward max-pool values (selected z3)
    const float* dpooled;     // [256][ldp] gradient w.r.t. the pooled BN3 output, ReLU-masked
    long long ldp;
    const float* c3_0; const float* c3_1; const float* c3_2;   // BN3 backward coefficients (slot 2)
    const float* gamma3;
    float* dh2;               // [R/64][64][64] masked gradient w.r.t. relu(bn2(z2))    (pass C out, pass D in)
    float* dw3;               // [256][64]  += (atomic)
    // pass D
    const float* c2_0; const float* c2_1; const float* c2_2;   // BN2 backward coefficients (slot 1)
    float* dw2;               // [64][64] += (atomic)
    float* amat;              // [grid][64][4] per-CTA sum dh1' x^T
    float* stats;             // pass C: [2*grid][64][2] (sum dh2', sum dh2' z2); pass D: [2*grid][64][2] (sum dh1', sum dh1' z1)
    // optional test hooks (pass C): the discrete decisions of the recomputed forward, so a checker can impose them
    unsigned char* dbg_mask1; // [R/64][64][64]  h1 > 0
    unsigned char* dbg_mask2; // [R/64][64][64]  h2 > 0
    unsigned char* dbg_arg;   // [256][ldp]      max-pool winner inside its group
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

// D (+)= A(K-major weight image, 64-wide K) * B(MN-major [64 ch][64 rows] image): reduction over the 64 channels
__device__ __forceinline__ void mma_w_act64(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int nhl,
                                            uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = umma_desc_sw128(a_hi + ks * 32);
        const uint64_t bd = umma_desc_mn_sw128(b_hi + ks * 2048, 8192, 1024);
        umma_bf16_ss(d, ad, bd, idesc, ks > 0 ? 1u : 0u);
        if (nhl == 2) {
            umma_bf16_ss(d, ad, umma_desc_mn_sw128(b_lo + ks * 2048, 8192, 1024), idesc, 1u);
            umma_bf16_ss(d, umma_desc_sw128(a_lo + ks * 32), bd, idesc, 1u);
        }
    }
}
// D (+)= A(K-major image, rows = channels, K = 64 batch rows) * B(K-major image, K = 64 batch rows): reduction over rows
__device__ __forceinline__ void mma_rows64(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int nhl,
                                           uint32_t idesc, bool first) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = umma_desc_sw128(a_hi + ks * 32), bd = umma_desc_sw128(b_hi + ks * 32);
        umma_bf16_ss(d, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
        if (nhl == 2) {
            umma_bf16_ss(d, ad, umma_desc_sw128(b_lo + ks * 32), idesc, 1u);
            umma_bf16_ss(d, umma_desc_sw128(a_lo + ks * 32), bd, idesc, 1u);
        }
    }
}

// producer shared by both backward passes: thread = channel (2 threads per channel, 32 rows each)
__device__ __forceinline__ void produce_h1_tile64(const float4* xtile, uint8_t* img, int nhl, int ch, int half, float wx, float wy,
                                                  float wz, float ww, float bf) {
#pragma unroll 2
    for (int q = 0; q < 4; ++q) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float4 x = xtile[half * 32 + q * 8 + e];
            v[e] = fmaxf(fmaf(wx, x.x, fmaf(wy, x.y, fmaf(wz, x.z, fmaf(ww, x.w, bf)))), 0.f);
        }
        store_img8(img, nhl, IMG64, ch, half * 4 + q, v);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// pass C   (22 warps)
// warps 0-7         : z3 consumers (thread = channel c)
// warps 8,9,12,13   : z2 -> h2 image            warps 16,17,20,21 : dh2 consumers      (thread = channel j, TMEM lanes 0..63)
// warps 10,11,14,15 : x -> h1 image producers   warp 18 : MMA issuer (warp 19 idles)
// TMEM columns: D2[b] 0/64, D3[h] 128/192, DH2 256, DW3[h] 320/384
// The MMA issue order is software-pipelined -- z3(it+1) is issued before the gradient GEMMs of tile it -- so the z3
// consumers (the longest role) always have the next accumulator waiting.
// --------------------------------------------------------------------------------------------------------------------
constexpr int C_THREADS = 22 * 32;

__global__ void __launch_bounds__(C_THREADS, 1) l1_bwd_c_kernel(const L1BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nhl = p.nhl;
    uint8_t* w2s = smem;                       // hi 8 KB | lo 8 KB
    uint8_t* w3s = w2s + 16384;                // [half][hi 16 KB | lo 16 KB]
    uint8_t* h1s = w3s + 65536;                // 2 stages x (hi 8 KB | lo 8 KB)
    uint8_t* h2s = h1s + 2 * 2 * IMG64;        // 2 stages x (hi | lo)
    uint8_t* dzs = h2s + 2 * 2 * IMG64;        // [256 c][64 r]: hi 32 KB | lo 32 KB
    uint8_t* xs = dzs + 65536;                 // 2 stages x 64 rows x 16 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + 2 * BT * 16);
    uint64_t *h1_full = bars, *h1_empty = bars + 2, *d2_full = bars + 4, *d2_empty = bars + 6, *h2_full = bars + 8,
             *h2_empty = bars + 10, *d3_full = bars + 12, *d3_empty = bars + 14, *dz_full = bars + 16, *dz_empty = bars + 17,
             *dh_full = bars + 18, *dh_empty = bars + 19, *w_bar = bars + 20, *fin_bar = bars + 21;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / BT;
    const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (grid <= ntiles)

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h1_full[i], 4);
            mbar_init(&h1_empty[i], 1);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 8);        // h2 producers + dh2 consumers both read z2
            mbar_init(&h2_full[i], 4);
            mbar_init(&h2_empty[i], 1);
            mbar_init(&d3_full[i], 1);
            mbar_init(&d3_empty[i], 4);
        }
        mbar_init(dz_full, 8);
        mbar_init(dz_empty, 1);
        mbar_init(dh_full, 1);
        mbar_init(dh_empty, 4);
        mbar_init(w_bar, 1);
        mbar_init(fin_bar, 1);
        mbar_fence_init();
    }
    if (warp == 18) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc_mn = umma_idesc_bf16(128, BT) | UMMA_B_MN_MAJOR;                      // weights x activation image
    const uint32_t idesc_dg = umma_idesc_bf16(128, BT) | UMMA_A_MN_MAJOR | UMMA_B_MN_MAJOR;    // W3^T x dz3
    const uint32_t idesc_kk = umma_idesc_bf16(128, 64);                                        // reduction over rows

    if (warp == 18) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, 8192u * nhl + 32768u * nhl);
            tma_bulk_g2s(w2s, p.w2_img, 8192, w_bar);
            if (nhl == 2) tma_bulk_g2s(w2s + 8192, p.w2_img + 16384, 8192, w_bar);
            for (int h = 0; h < 2; ++h) {
                tma_bulk_g2s(w3s + h * 32768, p.w3_img + h * 32768, 16384, w_bar);
                if (nhl == 2) tma_bulk_g2s(w3s + h * 32768 + 16384, p.w3_img + h * 32768 + 16384, 16384, w_bar);
            }
            mbar_wait(w_bar, 0);
            const uint32_t w2_hi = smem_u32(w2s), w2_lo = w2_hi + 8192;
            const uint32_t dz_hi = smem_u32(dzs), dz_lo = dz_hi + 32768;
            auto issue_mma2 = [&](int it) {        // z2(it) = W2 h1(it)
                const int b = it & 1, u = (it >> 1) & 1;
                mbar_wait(&h1_full[b], u);
                mbar_wait(&d2_empty[b], u ^ 1);
                tc_fence_after_sync();
                const uint32_t h1 = smem_u32(h1s + b * 2 * IMG64);
                mma_w_act64(tmem_base + 64 * b, w2_hi, w2_lo, h1, h1 + IMG64, nhl, idesc_mn);
                umma_commit(&h1_empty[b]);
                umma_commit(&d2_full[b]);
            };
            auto issue_mma3 = [&](int it) {        // z3(it) = W3 h2(it)
                const int b = it & 1, u = (it >> 1) & 1;
                mbar_wait(&h2_full[b], u);
                const uint32_t h2 = smem_u32(h2s + b * 2 * IMG64);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&d3_empty[h], (it & 1) ^ 1);
                    tc_fence_after_sync();
                    const uint32_t w3_hi = smem_u32(w3s + h * 32768);
                    mma_w_act64(tmem_base + 128 + 64 * h, w3_hi, w3_hi + 16384, h2, h2 + IMG64, nhl, idesc_mn);
                    umma_commit(&d3_full[h]);
                }
            };
            if (my_tiles > 0) issue_mma2(0);
            if (my_tiles > 1) issue_mma2(1);
            if (my_tiles > 0) issue_mma3(0);
#pragma unroll 1
            for (int it = 0; it < my_tiles; ++it) {
                const int b = it & 1;
                if (it + 2 < my_tiles) issue_mma2(it + 2);
                if (it + 1 < my_tiles) issue_mma3(it + 1);
                // dh2(it) = W3^T dz3(it) (reduction over the 256 channels); dW3 += dz3(it) h2(it)^T (reduction over the rows)
                const uint32_t h2 = smem_u32(h2s + b * 2 * IMG64);
                mbar_wait(dz_full, it & 1);
                mbar_wait(dh_empty, (it & 1) ^ 1);
                tc_fence_after_sync();
#pragma unroll 1
                for (int ks = 0; ks < 16; ++ks) {
                    const uint32_t wa = smem_u32(w3s + (ks >> 3) * 32768) + (ks & 7) * 2048;
                    const uint64_t a_hi = umma_desc_mn_sw128(wa, 16384, 1024), a_lo = umma_desc_mn_sw128(wa + 16384, 16384, 1024);
                    const uint64_t b_hi = umma_desc_mn_sw128(dz_hi + ks * 2048, 8192, 1024);
                    umma_bf16_ss(tmem_base + 256, a_hi, b_hi, idesc_dg, ks > 0 ? 1u : 0u);
                    if (nhl == 2) {
                        umma_bf16_ss(tmem_base + 256, a_hi, umma_desc_mn_sw128(dz_lo + ks * 2048, 8192, 1024), idesc_dg, 1u);
                        umma_bf16_ss(tmem_base + 256, a_lo, b_hi, idesc_dg, 1u);
                    }
                }
                umma_commit(dh_full);
#pragma unroll 1
                for (int h = 0; h < 2; ++h)
                    mma_rows64(tmem_base + 320 + 64 * h, dz_hi + h * 16384, dz_lo + h * 16384, h2, h2 + IMG64, nhl, idesc_kk, it == 0);
                umma_commit(dz_empty);
                umma_commit(&h2_empty[b]);
            }
            umma_commit(fin_bar);
        }
    } else if (warp == 10 || warp == 11 || warp == 14 || warp == 15) {
        // ---- producers ----
        const int pw = (warp == 10) ? 0 : (warp == 11) ? 1 : (warp == 14) ? 2 : 3;
        const int ptid = pw * 32 + lane, ch = ptid & 63, half = ptid >> 6;
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float wx = s1 * w.x, wy = s1 * w.y, wz = s1 * w.z, ww = s1 * w.w, bf = fmaf(s1, __ldg(p.b1 + ch), t1);
        int it = 0;
#pragma unroll 1
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            float4* xtile = reinterpret_cast<float4*>(xs + b * BT * 16);
            if (ptid < BT) xtile[ptid] = __ldg(reinterpret_cast<const float4*>(p.xt) + t * BT + ptid);
            named_bar_sync(1, 128);
            mbar_wait(&h1_empty[b], u ^ 1);
            produce_h1_tile64(xtile, h1s + b * 2 * IMG64, nhl, ch, half, wx, wy, wz, ww, bf);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[b]);
            if (p.dbg_mask1) {
                for (int r = 0; r < 32; ++r) {
                    float4 x = xtile[half * 32 + r];
                    float v = fmaf(wx, x.x, fmaf(wy, x.y, fmaf(wz, x.z, fmaf(ww, x.w, bf))));
                    p.dbg_mask1[(t * 64 + ch) * 64 + half * 32 + r] = v > 0.f ? 1 : 0;
                }
            }
        }
    } else if (warp == 8 || warp == 9 || warp == 12 || warp == 13) {
        // ---- z2 -> h2 = relu(bn2(z2)) image (thread = channel j) ----
        const int lg = warp & 1, colhalf = (warp >= 12) ? 1 : 0;
        const int j = lg * 32 + lane;
        const float a2 = __ldg(p.scale2 + j);
        const float c2 = fmaf(a2, __ldg(p.b2 + j), __ldg(p.shift2 + j));
        int it = 0;
#pragma unroll 1
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            mbar_wait(&d2_full[b], u);
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(64 * b + colhalf * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[b]);
            mbar_wait(&h2_empty[b], u ^ 1);
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
        }
    } else if (warp == 16 || warp == 17 || warp == 20 || warp == 21) {
        // ---- dh2 consumers (thread = channel j): ReLU2 mask, BN2 backward sums, masked gradient -> HBM ----
        const int lg = warp & 1, colhalf = (warp >= 20) ? 1 : 0;
        const int j = lg * 32 + lane;
        const float b2 = __ldg(p.b2 + j), a2 = __ldg(p.scale2 + j);
        const float c2 = fmaf(a2, b2, __ldg(p.shift2 + j));
        float s_acc = 0.f, q_acc = 0.f;
        int it = 0;
#pragma unroll 1
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            mbar_wait(&d2_full[b], u);
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(64 * b + colhalf * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[b]);
            mbar_wait(dh_full, it & 1);
            tc_fence_after_sync();
            float g[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(256 + colhalf * 32), g);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(dh_empty);
            float* out = p.dh2 + (t * 64 + j) * 64 + colhalf * 32;
            if (p.dbg_mask2) {
                for (int i = 0; i < 32; ++i) p.dbg_mask2[(t * 64 + j) * 64 + colhalf * 32 + i] = fmaf(a2, z[i], c2) > 0.f ? 1 : 0;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = (fmaf(a2, z[i], c2) > 0.f) ? g[i] : 0.f;
                s_acc += v;
                q_acc = fmaf(v, z[i] + b2, q_acc);
                g[i] = v;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(out)[q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
        }
        float* st = p.stats + ((long long)(blockIdx.x * 2 + colhalf) * 64 + j) * 2;
        st[0] = s_acc;
        st[1] = q_acc;
    } else if (warp < 8) {
        // ---- z3 consumers (thread = channel c): locate the pool winner, build dz3, finally flush dW3 ----
        const int h = warp >> 2, lq = warp & 3;
        const int c = h * 128 + lq * 32 + lane;
        const float b3 = __ldg(p.b3 + c);
        const float sgn = (__ldg(p.gamma3 + c) >= 0.f) ? 1.f : -1.f;
        const float k0 = __ldg(p.c3_0 + c), k1 = __ldg(p.c3_1 + c);
        const float k2 = fmaf(k1, b3, __ldg(p.c3_2 + c));
        const int K = p.K;                         // 32 or 64: a 32-column chunk never straddles two groups
        const int groups = BT / K;
        int it = 0;
#pragma unroll 1
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            // the pooled value / its gradient of this tile's group(s): issue the loads before waiting for the accumulator
            const long long g0 = (long long)c * p.ldp + t * groups;
            const float pool0 = __ldg(p.pooled + g0), dp0 = k0 * __ldg(p.dpooled + g0);
            float pool1 = pool0, dp1 = dp0;
            if (groups == 2) {
                pool1 = __ldg(p.pooled + g0 + 1);
                dp1 = k0 * __ldg(p.dpooled + g0 + 1);
            }
            mbar_wait(&d3_full[h], it & 1);
            tc_fence_after_sync();
            bool found = false;
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(128 + 64 * h + q * 32), v);
                tmem_ld_wait();
                if (q == 1) {                      // both halves are in registers / consumed: release the accumulator
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d3_empty[h]);
                }
                const float pool = (q == 1 && groups == 2) ? pool1 : pool0;
                const float dpv = (q == 1 && groups == 2) ? dp1 : dp0;
                if (q == 1 && groups == 2) found = false;
                int hit_at = -1;
                // dz3 = k1*z + k2 everywhere, + k0*dP at the first row whose value equals the pooled one
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const bool hit = !found && (fmaf(v[i] * sgn, sgn, b3) == pool);
                    found = found || hit;
                    hit_at = hit ? i : hit_at;
                    v[i] = fmaf(k1, v[i], k2) + (hit ? dpv : 0.f);
                }
                if (p.dbg_arg && hit_at >= 0)
                    p.dbg_arg[(long long)c * p.ldp + t * groups + (q * 32) / K] = (unsigned char)((q * 32 + hit_at) & (K - 1));
                if (q == 0) mbar_wait(dz_empty, (it & 1) ^ 1);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float w8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) w8[e] = v[g8 * 8 + e];
                    store_img8(dzs, nhl, 32768, c, q * 4 + g8, w8);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(dz_full);
        }
        // dW3 of this CTA's rows sits in TMEM: add it to the global gradient
        mbar_wait(fin_bar, 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            float a[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(320 + 64 * h + q * 32), a);
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
// pass D
// warps 0,1,4,5 : dh1 consumers (thread = channel i)    warps 8,9,12,13 : z2 consumers -> dz2 image (thread = channel j)
// warps 10,11,14,15 : producers                          warp 16 : MMA issuer
// TMEM columns: D2[b] 0/64, DH1[b] 128/192, DW2 256
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BWD_THREADS, 1) l1_bwd_d_kernel(const L1BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nhl = p.nhl;
    uint8_t* w2s = smem;                       // hi 8 KB | lo 8 KB, then 16 KB of slack (M=128 reads 128 rows)
    uint8_t* h1s = w2s + 32768;                // 2 stages x (hi | lo)
    uint8_t* dzs = h1s + 2 * 2 * IMG64;        // 2 stages x (hi | lo)   dz2 image [64 j][64 r]
    uint8_t* xs = dzs + 2 * 2 * IMG64 + 16384; // slack behind the last image, then 2 stages x 64 rows x 16 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + 2 * BT * 16);
    uint64_t *h1_full = bars, *h1_empty = bars + 2, *d2_full = bars + 4, *d2_empty = bars + 6, *dz_full = bars + 8,
             *dz_empty = bars + 10, *dh_full = bars + 12, *dh_empty = bars + 14, *x_free = bars + 16, *w_bar = bars + 18,
             *fin_bar = bars + 19;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = p.R / BT;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h1_full[i], 4);
            mbar_init(&h1_empty[i], 1);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 4);
            mbar_init(&dz_full[i], 4);
            mbar_init(&dz_empty[i], 1);
            mbar_init(&dh_full[i], 1);
            mbar_init(&dh_empty[i], 4);
            mbar_init(&x_free[i], 4);
        }
        mbar_init(w_bar, 1);
        mbar_init(fin_bar, 1);
        mbar_fence_init();
    }
    if (warp == 16) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc_mn = umma_idesc_bf16(128, BT) | UMMA_B_MN_MAJOR;
    const uint32_t idesc_dg = umma_idesc_bf16(128, BT) | UMMA_A_MN_MAJOR | UMMA_B_MN_MAJOR;
    const uint32_t idesc_kk = umma_idesc_bf16(128, 64);

    if (warp == 16) {
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, 8192u * nhl);
            tma_bulk_g2s(w2s, p.w2_img, 8192, w_bar);
            if (nhl == 2) tma_bulk_g2s(w2s + 8192, p.w2_img + 16384, 8192, w_bar);
            mbar_wait(w_bar, 0);
            const uint32_t w2_hi = smem_u32(w2s), w2_lo = w2_hi + 8192;
            auto issue_mma2 = [&](int it) {
                const int b = it & 1, u = (it >> 1) & 1;
                mbar_wait(&h1_full[b], u);
                mbar_wait(&d2_empty[b], u ^ 1);
                tc_fence_after_sync();
                const uint32_t h1 = smem_u32(h1s + b * 2 * IMG64);
                mma_w_act64(tmem_base + 64 * b, w2_hi, w2_lo, h1, h1 + IMG64, nhl, idesc_mn);
                umma_commit(&d2_full[b]);
            };
            int it = 0;
            long long t = blockIdx.x;
            if (t < ntiles) issue_mma2(0);
            for (; t < ntiles; t += gridDim.x, ++it) {
                const int b = it & 1, u = (it >> 1) & 1;
                if (t + gridDim.x < ntiles) issue_mma2(it + 1);
                mbar_wait(&dz_full[b], u);
                mbar_wait(&dh_empty[b], u ^ 1);
                tc_fence_after_sync();
                const uint32_t dz = smem_u32(dzs + b * 2 * IMG64), h1 = smem_u32(h1s + b * 2 * IMG64);
                // dh1[i][r] = sum_j W2[j][i] dz2[j][r]: W2 image read as an MN-major A operand (M = i, K = j)
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t a_hi = umma_desc_mn_sw128(w2_hi + ks * 2048, 16384, 1024);
                    const uint64_t a_lo = umma_desc_mn_sw128(w2_lo + ks * 2048, 16384, 1024);
                    const uint64_t b_hi = umma_desc_mn_sw128(dz + ks * 2048, 8192, 1024);
                    umma_bf16_ss(tmem_base + 128 + 64 * b, a_hi, b_hi, idesc_dg, ks > 0 ? 1u : 0u);
                    if (nhl == 2) {
                        umma_bf16_ss(tmem_base + 128 + 64 * b, a_hi, umma_desc_mn_sw128(dz + IMG64 + ks * 2048, 8192, 1024), idesc_dg, 1u);
                        umma_bf16_ss(tmem_base + 128 + 64 * b, a_lo, b_hi, idesc_dg, 1u);
                    }
                }
                umma_commit(&dh_full[b]);
                // dW2[j][i] += sum_r dz2[j][r] h1[i][r]
                mma_rows64(tmem_base + 256, dz, dz + IMG64, h1, h1 + IMG64, nhl, idesc_kk, it == 0);
                umma_commit(&dz_empty[b]);
                umma_commit(&h1_empty[b]);
            }
            umma_commit(fin_bar);
        }
    } else if (warp == 10 || warp == 11 || warp == 14 || warp == 15) {
        const int pw = (warp == 10) ? 0 : (warp == 11) ? 1 : (warp == 14) ? 2 : 3;
        const int ptid = pw * 32 + lane, ch = ptid & 63, half = ptid >> 6;
        const float s1 = __ldg(p.scale1 + ch), t1 = __ldg(p.shift1 + ch);
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + ch);
        const float wx = s1 * w.x, wy = s1 * w.y, wz = s1 * w.z, ww = s1 * w.w, bf = fmaf(s1, __ldg(p.b1 + ch), t1);
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            float4* xtile = reinterpret_cast<float4*>(xs + b * BT * 16);
            mbar_wait(&x_free[b], u ^ 1);                         // the dh1 consumers of tile it-2 are done with this x tile
            if (ptid < BT) xtile[ptid] = __ldg(reinterpret_cast<const float4*>(p.xt) + t * BT + ptid);
            named_bar_sync(1, 128);
            mbar_wait(&h1_empty[b], u ^ 1);
            produce_h1_tile64(xtile, h1s + b * 2 * IMG64, nhl, ch, half, wx, wy, wz, ww, bf);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h1_full[b]);
        }
    } else if (warp == 8 || warp == 9 || warp == 12 || warp == 13) {
        // ---- z2 consumers: dz2 = c0*dh2' + c1*z2 + c2 -> image ----
        const int lg = warp & 1, colhalf = (warp >= 12) ? 1 : 0;
        const int j = lg * 32 + lane;
        const float b2 = __ldg(p.b2 + j);
        const float k0 = __ldg(p.c2_0 + j), k1 = __ldg(p.c2_1 + j);
        const float k2 = fmaf(k1, b2, __ldg(p.c2_2 + j));
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            const float4* src = reinterpret_cast<const float4*>(p.dh2 + (t * 64 + j) * 64 + colhalf * 32);
            float g[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 v4 = __ldg(src + q);
                g[4 * q] = v4.x; g[4 * q + 1] = v4.y; g[4 * q + 2] = v4.z; g[4 * q + 3] = v4.w;
            }
            mbar_wait(&d2_full[b], u);
            tc_fence_after_sync();
            float z[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(64 * b + colhalf * 32), z);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_wait(&dz_empty[b], u ^ 1);
            uint8_t* img = dzs + b * 2 * IMG64;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = fmaf(k0, g[g8 * 8 + e], fmaf(k1, z[g8 * 8 + e], k2));
                store_img8(img, nhl, IMG64, j, colhalf * 4 + g8, h);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&dz_full[b]);
                mbar_arrive(&d2_empty[b]);
            }
        }
    } else if (warp == 0 || warp == 1 || warp == 4 || warp == 5) {
        // ---- dh1 consumers (thread = channel i): ReLU1 mask by recomputation, BN1 backward sums, A = sum dh1' x^T ----
        const int lg = warp & 1, colhalf = (warp >= 4) ? 1 : 0;
        const int i = lg * 32 + lane;
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w1) + i);
        const float b1 = __ldg(p.b1 + i), s1 = __ldg(p.scale1 + i), t1 = __ldg(p.shift1 + i);
        const float wx = s1 * w.x, wy = s1 * w.y, wz = s1 * w.z, ww = s1 * w.w, bf = fmaf(s1, b1, t1);   // as the producer folds BN1
        float s_acc = 0.f, q_acc = 0.f, ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int b = it & 1, u = (it >> 1) & 1;
            const float4* xtile = reinterpret_cast<const float4*>(xs + b * BT * 16) + colhalf * 32;
            mbar_wait(&dh_full[b], u);
            tc_fence_after_sync();
            float g[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(128 + 64 * b + colhalf * 32), g);
            tmem_ld_wait();
            tc_fence_before_sync();
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float4 x = xtile[r];
                const float z1 = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, b1))));
                const float zf = fmaf(wx, x.x, fmaf(wy, x.y, fmaf(wz, x.z, fmaf(ww, x.w, bf))));
                const float v = (zf > 0.f) ? g[r] : 0.f;
                s_acc += v;
                q_acc = fmaf(v, z1, q_acc);
                ax = fmaf(v, x.x, ax); ay = fmaf(v, x.y, ay); az = fmaf(v, x.z, az); aw = fmaf(v, x.w, aw);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&dh_empty[b]);
                mbar_arrive(&x_free[b]);
            }
        }
        const long long slot = (long long)(blockIdx.x * 2 + colhalf) * 64 + i;
        p.stats[slot * 2 + 0] = s_acc;
        p.stats[slot * 2 + 1] = q_acc;
        reinterpret_cast<float4*>(p.amat)[slot] = make_float4(ax, ay, az, aw);
        // dW2 of this CTA's rows: TMEM -> global (lanes 0..63 = channel j, 64 columns = channel i)
        mbar_wait(fin_bar, 0);
        tc_fence_after_sync();
        float a[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(256 + colhalf * 32), a);
        tmem_ld_wait();
        if (blockIdx.x < ntiles) {
#pragma unroll
            for (int q = 0; q < 32; ++q) atomicAdd(p.dw2 + i * 64 + colhalf * 32 + q, a[q]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// dW1[i][:] = c0_i * A_i + c1_i * (C w1_i + b1_i * sx) + c2_i * sx   with C = sum x x^T, sx = sum x (z1 is affine in x)
__global__ void l1_dw1_kernel(const float* __restrict__ amat, int P, const double* __restrict__ mom, const float* __restrict__ w1,
                              const float* __restrict__ b1, const float* __restrict__ c0, const float* __restrict__ c1,
                              const float* __restrict__ c2, float* __restrict__ dw1) {
    int i = threadIdx.x;
    if (i >= 64) return;
    double a[4] = {0, 0, 0, 0};
    for (int q = 0; q < P; ++q) {
        float4 v = reinterpret_cast<const float4*>(amat)[(long long)q * 64 + i];
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
    const int idx[4][4] = {{4, 5, 6, 7}, {5, 8, 9, 10}, {6, 9, 11, 12}, {7, 10, 12, 13}};
    for (int d = 0; d < 4; ++d) {
        double zx = (double)b1[i] * mom[d];
        for (int e = 0; e < 4; ++e) zx += (double)w1[i * 4 + e] * mom[idx[e][d]];
        dw1[i * 4 + d] = (float)((double)c0[i] * a[d] + (double)c1[i] * zx + (double)c2[i] * mom[d]);
    }
}

size_t l1_bwd_c_smem() { return 16384 + 65536 + 4 * IMG64 + 4 * IMG64 + 65536 + 2 * BT * 16 + 256 + 1024; }
size_t l1_bwd_d_smem() { return 32768 + 4 * IMG64 + 4 * IMG64 + 16384 + 2 * BT * 16 + 256 + 1024; }

// sum x (4) and sum x x^T (10 unique) over all rows, in double
__global__ void __launch_bounds__(256) l1_moments_kernel(const float4* __restrict__ xt, long long R, double* __restrict__ out) {
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
    size_t b = 16384 + (pass_b ? 65536 : 16384) + 2 * 2 * ACT_BYTES + (pass_b ? 2 * 2 * ACT_BYTES : 0) + 2 * TILE * 16 + 256;
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
    l1_moments_kernel<<<kNumSMs * 4, 256, 0, st>>>(reinterpret_cast<const float4*>(xt), R, mom14);
    return (int)cudaGetLastError();
}

int l1_bn1_launch(const double* mom14, double n, const float* w1, const float* b1, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float eps, float momentum, int training, float* mean, float* rstd,
                  float* scale, float* shift, cudaStream_t st) {
    ScopedTimer timer(TAG_BN, st);
    count_launch();
    l1_bn1_kernel<<<1, 64, 0, st>>>(mom14, n, w1, b1, gamma, beta, running_mean, running_var, eps, momentum, training, mean, rstd,
                                    scale, shift);
    return (int)cudaGetLastError();
}

// pass A: statistics of z2 -> stats [2*grid][64][2];  pass B: pooled [256][ldp] + statistics of z3 -> stats [grid][256][2]
int l1_fwd_launch(bool pass_b, const float* xt, long long R, int K, int nsplit, const float* w1, const float* b1, const float* scale1,
                  const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                  const void* w3_img, const float* b3, const float* gamma3, float* stats, float* pooled, long long ldp,
                  cudaStream_t st) {
    if (R <= 0 || R % TILE != 0 || K <= 0 || (K & (K - 1)) || TILE % K != 0) return (int)cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(false)));
        FACL_CHECK(cudaFuncSetAttribute(l1_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_smem_bytes(true)));
        configured = true;
    }
    L1Params p;
    p.xt = xt; p.R = R; p.K = K; p.nhl = (nsplit == 3) ? 2 : 1;
    p.w1 = w1; p.b1 = b1; p.scale1 = scale1; p.shift1 = shift1;
    p.w2_img = reinterpret_cast<const uint8_t*>(w2_img); p.b2 = b2; p.scale2 = scale2; p.shift2 = shift2;
    p.w3_img = reinterpret_cast<const uint8_t*>(w3_img); p.b3 = b3; p.gamma3 = gamma3;
    p.stats = stats; p.pooled = pooled; p.ldp = ldp;
    const int grid = l1_fused_grid(R);
    ScopedTimer timer(pass_b ? TAG_L1_PASS_B : TAG_L1_PASS_A, st);
    count_launch();
    if (pass_b)
        l1_fwd_kernel<true><<<grid, NTHREADS, l1_smem_bytes(true), st>>>(p);
    else
        l1_fwd_kernel<false><<<grid, NTHREADS, l1_smem_bytes(false), st>>>(p);
    return (int)cudaGetLastError();
}

static unsigned char *g_dbg_mask1 = nullptr, *g_dbg_mask2 = nullptr, *g_dbg_arg = nullptr;
void l1_set_debug_dump(unsigned char* mask1, unsigned char* mask2, unsigned char* arg) {
    g_dbg_mask1 = mask1; g_dbg_mask2 = mask2; g_dbg_arg = arg;
}

int l1_bwd_grid(long long R) {
    long long tiles = R / BT;
    return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}

static void fill_bwd_common(L1BwdParams& p, const float* xt, long long R, int K, int nsplit, const float* w1, const float* b1,
                            const float* scale1, const float* shift1, const void* w2_img, const float* b2, const float* scale2,
                            const float* shift2) {
    p.xt = xt; p.R = R; p.K = K; p.nhl = (nsplit == 3) ? 2 : 1;
    p.w1 = w1; p.b1 = b1; p.scale1 = scale1; p.shift1 = shift1;
    p.w2_img = reinterpret_cast<const uint8_t*>(w2_img); p.b2 = b2; p.scale2 = scale2; p.shift2 = shift2;
}

// pass C: needs dw3 zero-initialised by the caller (accumulated with atomics); stats [2*grid][64][2]
int l1_bwd_c_launch(const float* xt, long long R, int K, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                    const void* w3_img, const float* b3, const float* gamma3, const float* pooled, const float* dpooled,
                    long long ldp, const float* c3_0, const float* c3_1, const float* c3_2, float* dh2, float* dw3, float* stats,
                    cudaStream_t st) {
    if (R <= 0 || R % BT != 0 || (K != 32 && K != 64)) return (int)cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        FACL_CHECK(cudaFuncSetAttribute(l1_bwd_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_bwd_c_smem()));
        configured = true;
    }
    L1BwdParams p;
    memset(&p, 0, sizeof(p));
    fill_bwd_common(p, xt, R, K, nsplit, w1, b1, scale1, shift1, w2_img, b2, scale2, shift2);
    p.w3_img = reinterpret_cast<const uint8_t*>(w3_img); p.b3 = b3; p.gamma3 = gamma3;
    p.pooled = pooled; p.dpooled = dpooled; p.ldp = ldp; p.c3_0 = c3_0; p.c3_1 = c3_1; p.c3_2 = c3_2;
    p.dh2 = dh2; p.dw3 = dw3; p.stats = stats;
    p.dbg_mask1 = g_dbg_mask1; p.dbg_mask2 = g_dbg_mask2; p.dbg_arg = g_dbg_arg;
    ScopedTimer timer(TAG_L1_PASS_C, st);
    count_launch();
    l1_bwd_c_kernel<<<l1_bwd_grid(R), C_THREADS, l1_bwd_c_smem(), st>>>(p);
    return (int)cudaGetLastError();
}

// pass D: dw2 zero-initialised by the caller; stats [2*grid][64][2]; amat [2*grid][64][4]
int l1_bwd_d_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* w2_img, const float* b2, const float* c2_0, const float* c2_1,
                    const float* c2_2, const float* dh2, float* dw2, float* amat, float* stats, cudaStream_t st) {
    if (R <= 0 || R % BT != 0) return (int)cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        FACL_CHECK(cudaFuncSetAttribute(l1_bwd_d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l1_bwd_d_smem()));
        configured = true;
    }
    L1BwdParams p;
    memset(&p, 0, sizeof(p));
    fill_bwd_common(p, xt, R, 64, nsplit, w1, b1, scale1, shift1, w2_img, b2, nullptr, nullptr);
    p.c2_0 = c2_0; p.c2_1 = c2_1; p.c2_2 = c2_2;
    p.dh2 = const_cast<float*>(dh2); p.dw2 = dw2; p.amat = amat; p.stats = stats;
    ScopedTimer timer(TAG_L1_PASS_D, st);
    count_launch();
    l1_bwd_d_kernel<<<l1_bwd_grid(R), BWD_THREADS, l1_bwd_d_smem(), st>>>(p);
    return (int)cudaGetLastError();
}

int l1_dw1_launch(const float* amat, int P, const double* mom14, const float* w1, const float* b1, const float* c0, const float* c1,
                  const float* c2, float* dw1, cudaStream_t st) {
    ScopedTimer timer(TAG_L1_MISC, st);
    count_launch();
    l1_dw1_kernel<<<1, 64, 0, st>>>(amat, P, mom14, w1, b1, c0, c1, c2, dw1);
    return (int)cudaGetLastError();
}

}  // namespace facl
