// tcgen05 GEMM over PRE-CONVERTED activation images: every operand tile is staged by 1-D bulk TMA copies, no thread
// touches an operand element.  Used for the net3DV_3 stack (reference training_code/cn3d_model_conbag.py:180-196),
// whose three layers run on M*S rows -- few enough that converting each activation once (act_image_kernel: BN + ReLU
// or the BN-backward affine map, fp32 -> bf16 hi/lo, written straight in UMMA tile layout) costs far less HBM time
// than re-converting it inside every GEMM that reads it.
//
//   forward / data-gradient:  D[m][n] = sum_k W[m][k] * act[k][n]      A = packed weight image, B = image (MN-major)
//   weight-gradient        :  D[m][n] = sum_r dz[m][r] * act[n][r]     A, B = images (K-major over rows), split-K
//   loss similarity / back :  the same two forms with the (row-major) embedding matrices read as images whose
//                             "channels" are embedding rows, plus D[m][n] = sum_r W[m][r] * img[n][r] (packed A, image B)
//
// Warp roles: warps 0-3 epilogue (TMEM lane = output channel), warp 4 MMA issuer, warp 5 TMA issuer.
#include <stdio.h>
#include "gemm_tc.cuh"
#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"
#include "gemm_sched.cuh"

namespace facl {

namespace {

constexpr int IMG_THREADS = 192;
constexpr uint32_t IMG_LBO = 8192, IMG_SBO = 1024;   // B tile as staged: 64-row blocks 8 KB apart, 8-channel atoms 1 KB apart

template <int NHL, int EPI>
__global__ void __launch_bounds__(IMG_THREADS, 1) gemm_img_kernel(const GemmParams p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS, not generic LD/ST)
    // Two rings: A tiles (one per 64-wide k-block) and B HALF tiles (32 of the 64 k of a block when B is reduced over its
    // channels, so that three half-stages of the large operand are in flight instead of one full stage; a pair of
    // consecutive half slots forms one full K-major tile when B is reduced over its rows).
    constexpr int A_SLOT = A_TILE_BYTES * NHL, B_SLOT = (B_TILE_BYTES / 2) * NHL;
    constexpr int SA = (NHL == 2) ? 2 : 4, SB = (NHL == 2) ? 4 : 8;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + SA * A_SLOT;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + SB * B_SLOT);
    uint64_t* a_empty = a_full + SA;
    uint64_t* b_full = a_empty + SA;
    uint64_t* b_empty = b_full + SB;
    uint64_t* acc_full = b_empty + SB;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // per-epilogue-warp 32 x 36 float staging tiles: accumulator rows are (lane = channel, register = column), global rows
    // are channels, so a direct store scatters 16-byte pieces over 32 lines per instruction; through this tile every store /
    // load instruction moves four full 128-byte lines
    float* xpose_all = reinterpret_cast<float*>(b_ring + SB * B_SLOT + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mma_n = (p.Nd >= N_TILE) ? N_TILE : ((p.Nd + 15) & ~15);
    const bool a_img = (p.a_mode == A_IMAGE);        // A = activation image, K = its rows (else: packed weight tiles)
    const bool b_k = (p.b_mode == B_IMAGE_K);        // B image reduced over its rows (else over its channels, MN-major)

    if (threadIdx.x == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    Schedule sched(p);
    Work w;

    if (warp < 4) {
        // =============================== epilogue ===============================
        // EPI selects a compile-time specialisation (the fully general body is ~150 KB of unrolled SASS and ran out of the
        // instruction cache: "no_inst" stalls, 2-3 k cycles per 32-column chunk):
        //   1  forward      : bias, statistics, optional max-pool, channel-major store          (full warps, aligned)
        //   2  data-gradient: ReLU mask from zin, optional statistics, channel-major store        (full warps, aligned)
        //   3  weight-grad  : atomic accumulate
        //   0  everything else (row-major outputs, ragged shapes): runtime branches
        constexpr bool GEN = (EPI == 0);
        const int c = sched.mt * M_TILE + warp * 32 + lane;
        const bool cvalid = c < p.Md;
        const float bias = (cvalid && p.bias) ? __ldg(p.bias + c) : 0.f;
        const float zs0 = (cvalid && p.zs0) ? __ldg(p.zs0 + c) : 1.f;
        const float zs2 = (cvalid && p.zs2) ? __ldg(p.zs2 + c) : 0.f;
        const float psign = (cvalid && p.pool_sign) ? __ldg(p.pool_sign + c) : 1.f;
        const bool keep_max = psign >= 0.f;
        float stat0 = 0.f, stat1 = 0.f;
        float* tb = xpose_all + warp * (32 * 36);
        const int ch0 = sched.mt * M_TILE + warp * 32;                      // first channel of this warp
        const int tch = lane >> 3, tcol = (lane & 7) * 4;                   // transposed access: 4 channels x 8 float4 per instruction
        const bool has_zin = GEN ? (p.zin != nullptr) : (EPI == 2);
        const bool has_pool = (GEN || EPI == 1) && p.pool != 0;
        const bool has_stats = (EPI != 3) && p.stats != nullptr;
#ifdef FACL_PROFILE_ROLES
        long long pr_wait = 0, pr_work = 0, pr_t = clock64();
        int pr_n = 0;
#endif
        // EPI == 2: the zin rows of a chunk are fetched one chunk ahead (and the first chunk of a tile before the accumulator is
        // waited for): with one epilogue warp per scheduler nothing else hides the ~1 k-cycle latency of these loads, which made this
        // epilogue 21 k cycles per tile (3x the others) and the data-gradient GEMMs epilogue-bound
        float4 zn[8];
        auto zin_prefetch = [&](int nt, int cc) {
            const int n0p = nt * N_TILE + cc * 32;
            if (!cvalid || n0p >= p.Nd) return;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                zn[t] = __ldg(reinterpret_cast<const float4*>(p.zin + (long long)(ch0 + 4 * t + tch) * p.ldz + n0p + tcol));
        };
        if (EPI == 4 || EPI == 5) {
            // ===== contrastive-loss epilogues (see LossEpi): thread = anchor row, S never leaves the chip =====
            const LossEpi& L = p.loss;
            const int a = L.row0 + c;                                   // anchor index in [0, Ml + Bl)
            const int bloc = (a < L.Ml) ? (a % L.Bl) : (a - L.Ml);      // local sample of the anchor
            const int nsmp = L.n0 + bloc;                               // its global sample: keys of that sample are masked
            const int nr = nsmp / L.Bl, nb = nsmp - nr * L.Bl;
            float run_m = -INFINITY, run_e = 0.f;                       // EPI 4: online log-sum-exp state
            // EPI 5: the row's softmax constants
            float lc = 0.f, pgv = 0.f;
            const float* pgrow = nullptr;
            int nextv = -1;
            bool dead = !cvalid;
            const float invB = 1.f / (float)L.B;
            if (EPI == 5 && cvalid) {
                lc = __ldg(L.lc + bloc);
                if (a >= L.Ml) {
                    pgrow = L.pg + (long long)bloc * L.G;
                } else {
                    const int io = __ldg(L.inv_order + a / L.Bl);
                    if (io >= L.G - 1) {
                        dead = true;                                    // the last view of the chain is never an anchor
                    } else {
                        pgv = __ldg(L.pg + (long long)bloc * L.G + io);
                        nextv = __ldg(L.order + io + 1);
                    }
                }
            }
            // EPI 5 writes whole 64-row blocks of the image (its padding columns must be zero): an even number of 32-column chunks
            const int nchunks = (EPI == 5) ? ((mma_n + 63) / 64) * 2 : (mma_n + 31) / 32;
            // the masked columns of this row are the G keys of its own sample: j = nr*Ml + g*Bl + nb, g = 0..G-1
            const int jbase = nr * L.Ml + nb;
            for (int it = 0; sched.get(it, p, w); ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_full[buf], (it >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N_TILE);
#pragma unroll 1
                for (int cc = 0; cc < nchunks; ++cc) {
                    float v[32];
                    tmem_ld32(trow + cc * 32, v);
                    tmem_ld_wait();
                    const int n0 = w.nt * N_TILE + cc * 32;
                    int nvalid = p.Nd - n0;
                    nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
                    // bit i of `masked`: column n0 + i is a key of the anchor's own sample; mview = view of the lowest such column
                    unsigned masked = 0u;
                    int g0 = 0;
                    if (cvalid && nvalid > 0) {
                        int g = n0 > jbase ? (n0 - jbase + L.Bl - 1) / L.Bl : 0;          // first view whose key is >= n0
                        g0 = g;
                        for (int j = jbase + g * L.Bl; g < L.G && j < n0 + nvalid; ++g, j += L.Bl) masked |= 1u << (j - n0);
                    }
                    if (EPI == 4) {
                        if (!cvalid || nvalid <= 0) continue;
                        float cm = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid && !((masked >> i) & 1u)) cm = fmaxf(cm, v[i]);
                        if (masked) {                               // G masked columns per row: the positives are among them
                            int g = g0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if ((masked >> i) & 1u) L.pos[(long long)a * L.G + g++] = v[i];
                        }
                        if (cm > -INFINITY) {
                            const float nm = fmaxf(run_m, cm);
                            float ssum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                ssum[i & 3] += (i < nvalid && !((masked >> i) & 1u)) ? __expf(v[i] - nm) : 0.f;
                            run_e = run_e * __expf(run_m - nm) + ((ssum[0] + ssum[1]) + (ssum[2] + ssum[3]));   // exp(-inf) = 0 on the first chunk
                            run_m = nm;
                        }
                    } else {
                        const int rb = n0 >> 6;
                        if (c >= L.ds_cgs * 8 || rb >= L.ds_rbs) continue;       // outside the (padded) image
                        uint8_t* atom = reinterpret_cast<uint8_t*>(L.ds_hi) + ((long long)rb * L.ds_cgs + (c >> 3)) * 1024;
                        uint8_t* atom_lo = reinterpret_cast<uint8_t*>(L.ds_lo) + ((long long)rb * L.ds_cgs + (c >> 3)) * 1024;
                        const int chunk0 = (n0 & 63) >> 3;
                        int g = g0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int i = q * 8 + e;
                                float out = 0.f;
                                if (!dead && i < nvalid) {
                                    if ((masked >> i) & 1u) {
                                        out = pgrow ? __ldg(pgrow + g) : (g == nextv ? pgv : 0.f);
                                        ++g;
                                    } else {
                                        out = __expf(v[i] + lc) * invB;
                                    }
                                }
                                o[e] = out;
                            }
                            const uint32_t off = sw128_offset((uint32_t)(c & 7), (uint32_t)(chunk0 + q));
                            if (NHL == 2) {
                                uint4 h, l;
                                split_bf16x8(o, h, l);
                                *reinterpret_cast<uint4*>(atom + off) = h;
                                *reinterpret_cast<uint4*>(atom_lo + off) = l;
                            } else {
                                *reinterpret_cast<uint4*>(atom + off) = pack_bf16x8(o);
                            }
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            if (EPI == 4 && cvalid) {
                const int pidx = blockIdx.x / sched.numMT;
                float* st = L.part + ((long long)pidx * p.Md + c) * 2;
                st[0] = run_m;
                st[1] = run_e;
            }
        } else
        for (int it = 0; sched.get(it, p, w); ++it) {
            const int buf = it & 1;
            const int nchunks = (mma_n + 31) / 32;
            if (!GEN && has_zin) zin_prefetch(w.nt, 0);
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
#ifdef FACL_PROFILE_ROLES
            { long long t = clock64(); pr_wait += t - pr_t; pr_t = t; ++pr_n; }
#endif
            float best = 0.f;
            int barg = 0;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N_TILE);
#pragma unroll 1
            for (int cc = 0; cc < nchunks; ++cc) {
                float v[32];
                tmem_ld32(trow + cc * 32, v);
                tmem_ld_wait();
                const int n0 = w.nt * N_TILE + cc * 32;
                int nvalid = p.Nd - n0;
                nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
                if (EPI == 3 || (GEN && p.out_mode == OUT_ATOMIC_CHMAJOR)) {
                    if (cvalid) {
                        float* o = p.out + (long long)c * p.ldo + n0;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) atomicAdd(o + i, v[i]);
                    }
                    continue;
                }
                // nvalid is warp-uniform; in the specialised paths a warp is either wholly inside Md or wholly outside (Md % 32 == 0)
                if (nvalid <= 0 || !cvalid) continue;
                float z[32];
                if (has_zin) {
                    if (!GEN) {
                        // coalesced: each (prefetched) load covered 4 channel rows x 128 bytes; every thread now picks up its own row
#pragma unroll
                        for (int t = 0; t < 8; ++t) *reinterpret_cast<float4*>(tb + (4 * t + tch) * 36 + tcol) = zn[t];
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t4 = *reinterpret_cast<const float4*>(tb + lane * 36 + q * 4);
                            z[4 * q] = t4.x; z[4 * q + 1] = t4.y; z[4 * q + 2] = t4.z; z[4 * q + 3] = t4.w;
                        }
                        __syncwarp();
                        if (cc + 1 < nchunks) zin_prefetch(w.nt, cc + 1);
                    } else {
                        const float* zr = p.zin + (long long)c * p.ldz + n0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) z[i] = (i < nvalid) ? __ldg(zr + i) : 0.f;
                    }
                }
                // four independent partial sums: one warp per scheduler has to hide its own FADD / FFMA latency
                float s0p[4] = {0.f, 0.f, 0.f, 0.f}, s1p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float val = v[i] + bias;
                    if (has_zin) val = (fmaf(zs0, z[i], zs2) > 0.f) ? val : 0.f;
                    if (has_stats && (!GEN || i < nvalid)) {
                        s0p[i & 3] += val;
                        s1p[i & 3] = fmaf(val, has_zin ? z[i] : val, s1p[i & 3]);
                    }
                    v[i] = val;
                }
                stat0 += (s0p[0] + s0p[1]) + (s0p[2] + s0p[3]);
                stat1 += (s1p[0] + s1p[1]) + (s1p[2] + s1p[3]);
                if (has_pool && p.pool >= 32 && nvalid == 32) {
                    // the whole 32-column chunk lies inside one pooling group: branch-free scan, one merge per chunk
                    const float sg = keep_max ? 1.f : -1.f;
                    // four interleaved scans (i = k mod 4), merged with the lowest index winning ties (first-hit rule)
                    float cbk[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    int cik[4] = {0, 1, 2, 3};
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float sv = v[i] * sg;
                        cik[i & 3] = (sv > cbk[i & 3]) ? i : cik[i & 3];
                        cbk[i & 3] = fmaxf(cbk[i & 3], sv);
                    }
                    float cb = cbk[0];
                    int ci = cik[0];
#pragma unroll
                    for (int k = 1; k < 4; ++k) {
                        const bool take = (cbk[k] > cb) || (cbk[k] == cb && cik[k] < ci);
                        ci = take ? cik[k] : ci;
                        cb = take ? cbk[k] : cb;
                    }
                    const int pos0 = (cc * 32) & (p.pool - 1);
                    if (pos0 == 0 || cb > best) {
                        best = cb;
                        barg = pos0 + ci;
                    }
                    if (pos0 + 32 == p.pool) {
                        const long long gi = (long long)c * p.ldp + n0 / p.pool;
                        p.pool_out[gi] = best * sg;
                        if (p.pool_arg) p.pool_arg[gi] = (unsigned char)barg;
                    }
                } else if (has_pool) {
                    const int pm = p.pool - 1;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i < nvalid) {
                            int pos = (cc * 32 + i) & pm;
                            float sv = keep_max ? v[i] : -v[i];
                            if (pos == 0 || sv > best) {
                                best = sv;
                                barg = pos;
                            }
                            if (pos == pm) {
                                long long gi = (long long)c * p.ldp + (n0 + i) / p.pool;
                                p.pool_out[gi] = keep_max ? best : -best;
                                if (p.pool_arg) p.pool_arg[gi] = (unsigned char)barg;
                            }
                        }
                    }
                }
                if (!GEN) {
                    if (p.out_mode == OUT_CHMAJOR) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(tb + lane * 36 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int ch = 4 * t + tch;
                            *reinterpret_cast<float4*>(p.out + (long long)(ch0 + ch) * p.ldo + n0 + tcol) =
                                *reinterpret_cast<const float4*>(tb + ch * 36 + tcol);
                        }
                        __syncwarp();
                    }
                } else if (p.out_mode == OUT_CHMAJOR) {
                    float* o = p.out + (long long)c * p.ldo + n0;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) o[i] = v[i];
                } else if (p.out_mode == OUT_ROWMAJOR) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) p.out[(long long)(n0 + i) * p.ldo + c] = v[i];
                } else if (p.out_mode == OUT_ROWMAJOR_ACC) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) p.out[(long long)(n0 + i) * p.ldo + c] += v[i];
                } else if (p.out_mode == OUT_ATOMIC_ROWMAJOR) {   // split-K partial: lanes are consecutive m -> coalesced reductions
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) atomicAdd(p.out + (long long)(n0 + i) * p.ldo + c, v[i]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
#ifdef FACL_PROFILE_ROLES
            { long long t = clock64(); pr_work += t - pr_t; pr_t = t; }
#endif
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 0 && threadIdx.x == 0 && pr_n > 4)
            printf("gemm_img<%d> Md %d Nd %d Kd %d: epilogue tiles %d wait_acc %lld work %lld (cycles/tile)\n", EPI, p.Md, p.Nd, p.Kd,
                   pr_n, pr_wait / pr_n, pr_work / pr_n);
#endif
        if (has_stats && cvalid) {
            int pidx = sched.split ? 0 : (blockIdx.x / sched.numMT);
            float* st = p.stats + ((long long)pidx * p.Md + c) * 2;
            st[0] = stat0;
            st[1] = stat1;
        }
    } else if (warp == 4) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_bf16(M_TILE, mma_n) | (b_k ? 0u : UMMA_B_MN_MAJOR);
        const UDesc a_kd = udesc_k(smem_u32(a_ring)), b_kd = udesc_k(smem_u32(b_ring));
        const UDesc b_mnd = udesc_mn(smem_u32(b_ring), IMG_LBO / 2, IMG_SBO);
        int sa = 0, pa = 0, sb = 0, pb = 0;
#ifdef FACL_PROFILE_ROLES
        long long pm_acc = 0, pm_ops = 0, pm_t = clock64(), pm_t0 = pm_t;
        int pm_n = 0;
#endif
        for (int it = 0; sched.get(it, p, w); ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after_sync();
#ifdef FACL_PROFILE_ROLES
            { long long t = clock64(); pm_acc += t - pm_t; pm_t = t; ++pm_n; }
#endif
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
#ifdef FACL_PROFILE_ROLES
                pm_t = clock64();
#endif
                mbar_wait(&a_full[sa], pa);
                mbar_wait(&b_full[sb], pb);
#ifdef FACL_PROFILE_ROLES
                { long long t = clock64(); pm_ops += t - pm_t; pm_t = t; }
#endif
                // descriptors: ring bases built once (above); slot, hi/lo half and k-step are offsets in 16-byte units (umma.cuh, lean issue path)
                const uint32_t ao = (uint32_t)sa * (A_SLOT / 16);
                if (b_k) {
                    // one full K-major B tile in a pair of half slots: [hi | lo]
                    mbar_wait(&b_full[sb], pb);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        const uint32_t bo = (uint32_t)sb * (B_SLOT / 16);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t acc = (kb > w.kb0 || ks > 0) ? 1u : 0u;
                            umma_ss(d_tmem, a_kd, ao + ks * 2, b_kd, bo + ks * 2, idesc, acc);
                            if (NHL == 2) {
                                umma_ss(d_tmem, a_kd, ao + ks * 2, b_kd, bo + B_TILE_BYTES / 16 + ks * 2, idesc, 1u);
                                umma_ss(d_tmem, a_kd, ao + A_TILE_BYTES / 16 + ks * 2, b_kd, bo + ks * 2, idesc, 1u);
                            }
                        }
                        umma_commit(&b_empty[sb]);
                        umma_commit(&a_empty[sa]);
                        if (kb == w.kb1 - 1) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    sb += 2;
                    if (sb == SB) { sb = 0; pb ^= 1; }
                } else {
                    // two MN-major half tiles (32 channels each): slot = [hi 16 KB | lo 16 KB], 64-row blocks 4 KB apart
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after_sync();
                        if (elect_one_sync()) {
                            const uint32_t bo = (uint32_t)sb * (B_SLOT / 16);
                            const uint32_t ah = ao + h * 4;                      // k-steps 2h, 2h + 1 of the A tile
#pragma unroll
                            for (int k2 = 0; k2 < 2; ++k2) {
                                const uint32_t acc = (kb > w.kb0 || h > 0 || k2 > 0) ? 1u : 0u;
                                umma_ss(d_tmem, a_kd, ah + k2 * 2, b_mnd, bo + k2 * (2 * IMG_SBO / 16), idesc, acc);
                                if (NHL == 2) {
                                    umma_ss(d_tmem, a_kd, ah + k2 * 2, b_mnd, bo + (B_TILE_BYTES / 2) / 16 + k2 * (2 * IMG_SBO / 16), idesc, 1u);
                                    umma_ss(d_tmem, a_kd, ah + A_TILE_BYTES / 16 + k2 * 2, b_mnd, bo + k2 * (2 * IMG_SBO / 16), idesc, 1u);
                                }
                            }
                            umma_commit(&b_empty[sb]);
                            if (h == 1) {
                                umma_commit(&a_empty[sa]);
                                if (kb == w.kb1 - 1) umma_commit(&acc_full[buf]);
                            }
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
                if (++sa == SA) { sa = 0; pa ^= 1; }
            }
        }
#ifdef FACL_PROFILE_ROLES
        if (blockIdx.x == 0 && lane == 0 && pm_n > 4)
            printf("gemm_img Md %d Nd %d Kd %d: mma tiles %d wait_acc_empty %lld wait_operands %lld total %lld (cycles/tile)\n", p.Md, p.Nd,
                   p.Kd, pm_n, pm_acc / pm_n, pm_ops / pm_n, (clock64() - pm_t0) / pm_n);
#endif
    } else if (elect_one_sync()) {
        // =============================== TMA issuer (warp 5, one elected lane) ===============================
        int sa = 0, pa = 0, sb = 0, pb = 0;
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.a_packed);
        const uint8_t* ai_hi = reinterpret_cast<const uint8_t*>(p.a_img.hi);
        const uint8_t* ai_lo = reinterpret_cast<const uint8_t*>(p.a_img.lo);
        const uint8_t* bi_hi = reinterpret_cast<const uint8_t*>(p.b_img.hi);
        const uint8_t* bi_lo = reinterpret_cast<const uint8_t*>(p.b_img.lo);
        // The activation images do not fit in L2 (168 MB for 512 channels x 81 920 rows), so a stage's bulk copies see HBM latency,
        // and two 96 KB stages cannot cover it.  In the weight-gradient form (both operands are row blocks of images: 16-32 KB
        // contiguous pieces that no other CTA reads) the pieces of the k-block PF_AHEAD steps ahead are requested into L2 while the
        // ring is still busy (-3..-8 %).  Not in the weights x image form: there every piece is shared by the CTAs of all m-tiles
        // and comes in 8 KB fragments; prefetching them made those GEMMs 30 % slower, also when only one CTA per column did it (measured).
        constexpr int PF_AHEAD = 2;
        auto prefetch_kblock = [&](const Work& wk, int kb) {
            if (a_img) {
                int na = p.a_img.cgs - wk.mt * 16;
                na = na > 16 ? 16 : na;
                const long long aoff = ((long long)kb * p.a_img.cgs + wk.mt * 16) * 1024;
                tma_prefetch_l2(ai_hi + aoff, na * 1024);
                if (NHL == 2) tma_prefetch_l2(ai_lo + aoff, na * 1024);
            }
            if (b_k) {
                int nb = p.b_img.cgs - wk.nt * 32;
                nb = nb > 32 ? 32 : nb;
                const long long boff = ((long long)kb * p.b_img.cgs + wk.nt * 32) * 1024;
                tma_prefetch_l2(bi_hi + boff, nb * 1024);
                if (NHL == 2) tma_prefetch_l2(bi_lo + boff, nb * 1024);
            }
        };
        for (int it = 0; sched.get(it, p, w); ++it) {
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                if (b_k) {
                    int pk = kb + PF_AHEAD;
                    if (pk < w.kb1) {
                        prefetch_kblock(w, pk);
                    } else {
                        Work wn;
                        if (sched.get(it + 1, p, wn) && wn.kb0 + (pk - w.kb1) < wn.kb1) prefetch_kblock(wn, wn.kb0 + (pk - w.kb1));
                    }
                }
                // ---- A: one 128 x 64 tile ----
                mbar_wait(&a_empty[sa], pa ^ 1);
                uint8_t* a_hi = a_ring + sa * A_SLOT;
                if (a_img) {                         // 16 channel atoms of row block kb, contiguous
                    int na = p.a_img.cgs - w.mt * 16;
                    na = na > 16 ? 16 : na;
                    const long long aoff = ((long long)kb * p.a_img.cgs + w.mt * 16) * 1024;
                    mbar_arrive_expect_tx(&a_full[sa], (uint32_t)(na * 1024 * NHL));
                    tma_bulk_g2s(a_hi, ai_hi + aoff, na * 1024, &a_full[sa]);
                    if (NHL == 2) tma_bulk_g2s(a_hi + A_TILE_BYTES, ai_lo + aoff, na * 1024, &a_full[sa]);
                } else {                             // packed weight tile, hi | lo adjacent
                    mbar_arrive_expect_tx(&a_full[sa], (uint32_t)(A_TILE_BYTES * NHL));
                    tma_bulk_g2s(a_hi, wimg + ((long long)w.mt * p.a_packed_kblocks + kb) * (2ll * A_TILE_BYTES), A_TILE_BYTES * NHL,
                                 &a_full[sa]);
                }
                if (++sa == SA) { sa = 0; pa ^= 1; }
                // ---- B ----
                if (b_k) {                           // 32 channel atoms of row block kb, contiguous, into a pair of half slots
                    int nb = p.b_img.cgs - w.nt * 32;
                    nb = nb > 32 ? 32 : nb;
                    const long long boff = ((long long)kb * p.b_img.cgs + w.nt * 32) * 1024;
                    mbar_wait(&b_empty[sb], pb ^ 1);
                    uint8_t* b_hi = b_ring + sb * B_SLOT;
                    mbar_arrive_expect_tx(&b_full[sb], (uint32_t)(nb * 1024 * NHL));
                    tma_bulk_g2s(b_hi, bi_hi + boff, nb * 1024, &b_full[sb]);
                    if (NHL == 2) tma_bulk_g2s(b_hi + B_TILE_BYTES, bi_lo + boff, nb * 1024, &b_full[sb]);
                    sb += 2;
                    if (sb == SB) { sb = 0; pb ^= 1; }
                } else {                             // per half: up to four 64-row blocks x 4 channel atoms (4 KB pieces)
                    const int rb0 = w.nt * (N_TILE / 64);
                    int nrb = p.b_img.rbs - rb0;
                    nrb = nrb > N_TILE / 64 ? N_TILE / 64 : nrb;
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        uint8_t* b_hi = b_ring + sb * B_SLOT;
                        mbar_arrive_expect_tx(&b_full[sb], (uint32_t)(nrb * 4096 * NHL));
                        for (int r = 0; r < nrb; ++r) {
                            const long long off = ((long long)(rb0 + r) * p.b_img.cgs + kb * 8 + h * 4) * 1024;
                            tma_bulk_g2s(b_hi + r * 4096, bi_hi + off, 4096, &b_full[sb]);
                            if (NHL == 2) tma_bulk_g2s(b_hi + B_TILE_BYTES / 2 + r * 4096, bi_lo + off, 4096, &b_full[sb]);
                        }
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32 channel-major activation -> bf16 hi/lo image.  Block = one 8-channel group x 256 rows; thread = (channel, 8 rows):
// a warp reads 1 KB of one channel row per source and writes four full 128-byte swizzled rows.
// ---------------------------------------------------------------------------------------------------------------------
struct ImgParams {
    OperandSrc s;
    int C;
    long long R;
    const unsigned char* pool_arg;
    int pool;
    int nhl;
    uint8_t* hi;
    uint8_t* lo;
    int cgs, rbs;
    long long ld1;            // leading dimension of src1
};

__global__ void __launch_bounds__(256) act_image_kernel(const ImgParams q) {
    pdl_prologue();
    const int cg = blockIdx.y;
    const int c8 = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = cg * 8 + c8;
    const long long r0 = ((long long)blockIdx.x * 32 + lane) * 8;           // first of this thread's 8 rows
    const long long rb = r0 >> 6;
    const int chunk = (int)((r0 >> 3) & 7);
    if (rb >= q.rbs) return;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (c < q.C && r0 < q.R) {
        const bool full8 = r0 + 8 <= q.R;
        float a[8], b[8];
        if (q.pool) {
            // rows r0..r0+7 lie in one pooling group (pool is a multiple of 8)
            const long long g = r0 / q.pool;
            const int pos0 = (int)(r0 - g * q.pool);
            const float dv = __ldg(q.s.src0 + (long long)c * q.s.ld + g);
            const int arg = q.pool_arg[(long long)c * q.s.ld + g];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = (arg == pos0 + e) ? dv : 0.f;
        } else {
            const float* p0 = q.s.src0 + (long long)c * q.s.ld + r0;
            if (full8 && ((reinterpret_cast<uintptr_t>(p0) & 15) == 0)) {
                float4 t0 = __ldg(reinterpret_cast<const float4*>(p0)), t1 = __ldg(reinterpret_cast<const float4*>(p0) + 1);
                a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = (r0 + e < q.R) ? __ldg(p0 + e) : 0.f;
            }
        }
        if (q.s.src1) {
            const float* p1 = q.s.src1 + (long long)c * q.ld1 + r0;
            if (full8 && ((reinterpret_cast<uintptr_t>(p1) & 15) == 0)) {
                float4 t0 = __ldg(reinterpret_cast<const float4*>(p1)), t1 = __ldg(reinterpret_cast<const float4*>(p1) + 1);
                b[0] = t0.x; b[1] = t0.y; b[2] = t0.z; b[3] = t0.w; b[4] = t1.x; b[5] = t1.y; b[6] = t1.z; b[7] = t1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) b[e] = (r0 + e < q.R) ? __ldg(p1 + e) : 0.f;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) b[e] = 0.f;
        }
        const float s0 = q.s.s0 ? __ldg(q.s.s0 + c) : 1.f, s1 = q.s.s1 ? __ldg(q.s.s1 + c) : 0.f;
        const float s2 = q.s.s2 ? __ldg(q.s.s2 + c) : 0.f, lo = q.s.lo ? __ldg(q.s.lo + c) : -INFINITY;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (r0 + e < q.R) ? fmaxf(fmaf(s0, a[e], fmaf(s1, b[e], s2)), lo) : 0.f;
    }
    const long long off = (rb * q.cgs + cg) * 1024 + sw128_offset((uint32_t)c8, (uint32_t)chunk);
    if (q.nhl == 2) {
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(q.hi + off) = h;
        *reinterpret_cast<uint4*>(q.lo + off) = l;
    } else {
        *reinterpret_cast<uint4*>(q.hi + off) = pack_bf16x8(v);
    }
}

}  // namespace

size_t act_image_half_bytes(int C, long long R) {
    const size_t cgs = (size_t)((C + 63) / 64) * 8, rbs = (size_t)((R + 63) / 64);
    return cgs * rbs * 1024;
}

int launch_gemm_img(const GemmParams& p, cudaStream_t stream) {
    static DeviceOnce configured;
    const int smem_bytes = 4 * (A_TILE_BYTES + B_TILE_BYTES) + 1024 + 256 + 4 * 32 * 36 * 4;
    if (configured.need()) {
#define FACL_IMG_ATTR(N, E) \
        FACL_CHECK(cudaFuncSetAttribute(gemm_img_kernel<N, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        FACL_IMG_ATTR(1, 0) FACL_IMG_ATTR(1, 1) FACL_IMG_ATTR(1, 2) FACL_IMG_ATTR(1, 3) FACL_IMG_ATTR(1, 4) FACL_IMG_ATTR(1, 5)
        FACL_IMG_ATTR(2, 0) FACL_IMG_ATTR(2, 1) FACL_IMG_ATTR(2, 2) FACL_IMG_ATTR(2, 3) FACL_IMG_ATTR(2, 4) FACL_IMG_ATTR(2, 5)
#undef FACL_IMG_ATTR
        configured.done();
    }
    const bool a_img = p.a_mode == A_IMAGE, b_k = p.b_mode == B_IMAGE_K;
    if ((!a_img && p.a_mode != A_PACKED) || (!b_k && p.b_mode != B_IMAGE_MN) || (a_img && !b_k)) return (int)cudaErrorInvalidValue;
    if (p.Md <= 0 || p.Nd <= 0 || p.Kd <= 0 || (p.nsplit != 1 && p.nsplit != 3)) return (int)cudaErrorInvalidValue;
    if (p.pool && ((p.pool & (p.pool - 1)) != 0 || p.pool > N_TILE)) return (int)cudaErrorInvalidValue;
    if (!p.b_img.hi || (p.nsplit == 3 && !p.b_img.lo)) return (int)cudaErrorInvalidValue;
    const int numMT = (p.Md + M_TILE - 1) / M_TILE, numNT = (p.Nd + N_TILE - 1) / N_TILE, KB = (p.Kd + K_BLK - 1) / K_BLK;
    if (a_img) {
        // Kd = rows (a multiple of 64 after padding); every row block of the image is read
        if (!p.a_img.hi || (p.nsplit == 3 && !p.a_img.lo) || p.a_img.rbs != KB || p.a_img.cgs * 8 < p.Md) return (int)cudaErrorInvalidValue;
    } else if (!p.a_packed || p.a_packed_kblocks < KB) {
        return (int)cudaErrorInvalidValue;
    }
    if (b_k) {
        if (p.b_img.rbs != KB || p.b_img.cgs * 8 < p.Nd) return (int)cudaErrorInvalidValue;
    } else if (p.b_img.cgs < KB * 8 || (long long)p.b_img.rbs * 64 < p.Nd) {
        return (int)cudaErrorInvalidValue;
    }
    if (p.ksplit > 1 && ((p.out_mode != OUT_ATOMIC_CHMAJOR && p.out_mode != OUT_ATOMIC_ROWMAJOR) || p.stats || p.pool))
        return (int)cudaErrorInvalidValue;
    int grid;
    if (p.ksplit > 1) {
        if (p.ksplit > KB) return (int)cudaErrorInvalidValue;
        grid = numMT * numNT * p.ksplit;
    } else {
        grid = numMT * gemm_tc_ctas_per_mtile(p.Md, p.Nd);
    }
    // epilogue specialisation: the fast ones need full 32-channel warps, whole 32-column chunks and 16-byte aligned rows
    const bool tidy = (p.Md % 32 == 0) && (p.Nd % 32 == 0) && p.out && (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int epi = 0;
    if (p.loss.mode) {
        const LossEpi& L = p.loss;
        if (p.ksplit > 1 || p.stats || p.pool || L.G <= 0 || L.Bl <= 0 || L.Ml != L.G * L.Bl || L.B % L.Bl != 0 || p.Nd % L.Ml != 0)
            return (int)cudaErrorInvalidValue;
        if (L.mode == 1 && (!L.part || !L.pos)) return (int)cudaErrorInvalidValue;
        if (L.mode == 2 && (!L.lc || !L.pg || !L.ds_hi || (p.nsplit == 3 && !L.ds_lo) || (L.row0 == 0 && (!L.order || !L.inv_order))))
            return (int)cudaErrorInvalidValue;
        epi = L.mode == 1 ? 4 : 5;
    } else if (p.out_mode == OUT_ATOMIC_CHMAJOR) epi = 3;
    else if (p.out_mode == OUT_CHMAJOR && tidy && !p.zin) epi = 1;
    else if (p.out_mode == OUT_CHMAJOR && tidy && p.zin && !p.pool && (p.ldz % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.zin) & 15) == 0)) epi = 2;
    ScopedTimer timer(p.tag, stream);
    count_launch();
#define FACL_IMG_LAUNCH(N, E) FACL_LAUNCH_OK(launch_pdl(gemm_img_kernel<N, E>, dim3(grid), dim3(IMG_THREADS), smem_bytes, stream, p))
    if (p.nsplit == 3) {
        if (epi == 1) FACL_IMG_LAUNCH(2, 1); else if (epi == 2) FACL_IMG_LAUNCH(2, 2); else if (epi == 3) FACL_IMG_LAUNCH(2, 3);
        else if (epi == 4) FACL_IMG_LAUNCH(2, 4); else if (epi == 5) FACL_IMG_LAUNCH(2, 5); else FACL_IMG_LAUNCH(2, 0);
    } else {
        if (epi == 1) FACL_IMG_LAUNCH(1, 1); else if (epi == 2) FACL_IMG_LAUNCH(1, 2); else if (epi == 3) FACL_IMG_LAUNCH(1, 3);
        else if (epi == 4) FACL_IMG_LAUNCH(1, 4); else if (epi == 5) FACL_IMG_LAUNCH(1, 5); else FACL_IMG_LAUNCH(1, 0);
    }
#undef FACL_IMG_LAUNCH
    return (int)cudaGetLastError();
}

int act_image_launch(const OperandSrc& src, long long ld1, int C, long long R, const unsigned char* pool_arg, int pool, int nhl,
                     const ActImage& img, int tag, cudaStream_t st) {
    if (C <= 0 || R <= 0 || !src.src0 || !img.hi || (nhl == 2 && !img.lo)) return (int)cudaErrorInvalidValue;
    if (pool && (pool % 8 != 0 || !pool_arg)) return (int)cudaErrorInvalidValue;
    if (img.cgs * 8 < C || (long long)img.rbs * 64 < R || img.cgs % 8 != 0) return (int)cudaErrorInvalidValue;
    ImgParams q;
    q.s = src; q.C = C; q.R = R; q.pool_arg = pool_arg; q.pool = pool; q.nhl = nhl;
    q.hi = reinterpret_cast<uint8_t*>(const_cast<void*>(img.hi));
    q.lo = reinterpret_cast<uint8_t*>(const_cast<void*>(img.lo));
    q.cgs = img.cgs; q.rbs = img.rbs; q.ld1 = ld1 ? ld1 : src.ld;
    dim3 grid((unsigned)(img.rbs + 3) / 4, (unsigned)img.cgs);
    ScopedTimer timer(tag, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(act_image_kernel, grid, dim3(256), 0, st, q));
    return (int)cudaGetLastError();
}

}  // namespace facl
