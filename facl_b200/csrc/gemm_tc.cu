// Generic tcgen05 GEMM (see gemm_tc.cuh).  Warp-specialised, persistent over output tiles:
//   warps 0-3  epilogue   (TMEM lanes 32w..32w+31 -> one output channel per thread)
//   warps 4-9  producers  (fp32 global -> transform -> bf16 hi/lo -> swizzled smem tiles; loads batched 4 tasks deep)
//   warp  10   MMA issuer (one lane issues tcgen05.mma; owns the TMEM allocation)
//   warp  11   TMA        (one lane bulk-copies pre-packed weight tiles, cp.async.bulk + mbarrier tx)
#include "gemm_tc.cuh"
#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"
#include "gemm_sched.cuh"

namespace facl {

namespace {

constexpr int NUM_PROD_WARPS = 6;
constexpr int PROD_THREADS = NUM_PROD_WARPS * 32;
constexpr int THREADS = 384;           // 4 epilogue + 6 producer + MMA + TMA warps

__device__ __forceinline__ float xform(float a, float b, float s0, float s1, float s2, float lo) {
    return fmaxf(fmaf(s0, a, fmaf(s1, b, s2)), lo);
}

__device__ __forceinline__ void store_chunk(uint8_t* hi, uint8_t* lo, int nhl, uint32_t row, uint32_t chunk, const float (&v)[8]) {
    uint32_t off = sw128_offset(row, chunk);
    if (nhl == 2) {
        uint4 h, l;
        split_bf16x8(v, h, l);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
    } else {
        *reinterpret_cast<uint4*>(hi + off) = pack_bf16x8(v);
    }
}

// One producer task = 8 consecutive source floats (two 16-byte loads per source) -> one 16-byte chunk of an operand tile.
// Loads of PBATCH tasks are issued back to back before any of them is consumed, so a thread keeps PBATCH * (2..4)
// independent 16-byte requests in flight instead of paying the memory latency once per task.
constexpr int PBATCH = 2;

struct TaskData {
    float a[8], b[8];
    float c0, c1, c2, cl;
    int nvalid;       // valid leading elements (0 = whole chunk is zero)
};

__device__ __forceinline__ void load_task(const OperandSrc& s, long long off, int nvalid, int ch, TaskData& t) {
    t.nvalid = nvalid;
    if (nvalid <= 0) return;
    const float* p0 = s.src0 + off;
    const bool vec = (nvalid == 8) && ((reinterpret_cast<uintptr_t>(p0) & 15) == 0);
    if (vec) {
        float4 t0 = __ldg(reinterpret_cast<const float4*>(p0));
        float4 t1 = __ldg(reinterpret_cast<const float4*>(p0) + 1);
        t.a[0] = t0.x; t.a[1] = t0.y; t.a[2] = t0.z; t.a[3] = t0.w; t.a[4] = t1.x; t.a[5] = t1.y; t.a[6] = t1.z; t.a[7] = t1.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) t.a[e] = (e < nvalid) ? __ldg(p0 + e) : 0.f;
    }
    if (s.src1) {
        const float* p1 = s.src1 + off;
        if (vec && ((reinterpret_cast<uintptr_t>(p1) & 15) == 0)) {
            float4 t0 = __ldg(reinterpret_cast<const float4*>(p1));
            float4 t1 = __ldg(reinterpret_cast<const float4*>(p1) + 1);
            t.b[0] = t0.x; t.b[1] = t0.y; t.b[2] = t0.z; t.b[3] = t0.w; t.b[4] = t1.x; t.b[5] = t1.y; t.b[6] = t1.z; t.b[7] = t1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) t.b[e] = (e < nvalid) ? __ldg(p1 + e) : 0.f;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) t.b[e] = 0.f;
    }
    t.c0 = s.s0 ? __ldg(s.s0 + ch) : 1.f;
    t.c1 = s.s1 ? __ldg(s.s1 + ch) : 0.f;
    t.c2 = s.s2 ? __ldg(s.s2 + ch) : 0.f;
    t.cl = s.lo ? __ldg(s.lo + ch) : -INFINITY;
}

__device__ __forceinline__ void finish_task(const TaskData& t, float (&v)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (e < t.nvalid) ? xform(t.a[e], t.b[e], t.c0, t.c1, t.c2, t.cl) : 0.f;
}

// fp32 source with K contiguous: rows of the tile are rows of the source; transform constants are per ROW.
__device__ __forceinline__ void produce_rowmajor(const OperandSrc& s, uint8_t* hi, uint8_t* lo, int nhl, int nrows, int row0,
                                                 int row_limit, int k0, int k_limit, int ptid) {
    const int ntasks = nrows * 8;
    for (int base = ptid; base < ntasks; base += PROD_THREADS * PBATCH) {
        TaskData td[PBATCH];
#pragma unroll
        for (int u = 0; u < PBATCH; ++u) {
            const int task = base + u * PROD_THREADS;
            const int r = task >> 3, j = task & 7;
            const int grow = row0 + r, k = k0 + j * 8;
            int nv = (task < ntasks && grow < row_limit) ? (k_limit - k) : 0;
            nv = nv > 8 ? 8 : nv;
            load_task(s, (long long)grow * s.ld + k, nv, grow, td[u]);
        }
#pragma unroll
        for (int u = 0; u < PBATCH; ++u) {
            const int task = base + u * PROD_THREADS;
            if (task < ntasks) {
                float v[8];
                finish_task(td[u], v);
                store_chunk(hi, lo, nhl, task >> 3, task & 7, v);
            }
        }
    }
}

// fp32 source stored channel-major [K][ld], staged as an MN-MAJOR operand tile: a task is (k, 8 consecutive rows), i.e.
// two float4 loads and one 16-byte shared store -- no transposition anywhere.  Transform constants are per K.
constexpr uint32_t B_MN_LBO = 8192, B_MN_SBO = 1024;   // 64-row blocks 8 KB apart, 8-k groups 1 KB apart (32 KB tile)
__device__ __forceinline__ void produce_chmajor_mn(const OperandSrc& s, uint8_t* hi, uint8_t* lo, int nhl, int nrows, int n0,
                                                   int n_limit, int k0, int k_limit, int ptid) {
    const int chunks = (nrows + 7) >> 3;
    const int ntasks = K_BLK * chunks;
    for (int base = ptid; base < ntasks; base += PROD_THREADS * PBATCH) {
        TaskData td[PBATCH];
#pragma unroll
        for (int u = 0; u < PBATCH; ++u) {
            const int task = base + u * PROD_THREADS;
            const int kk = task / chunks, c = task - kk * chunks;
            const int k = k0 + kk, n = n0 + c * 8;
            int nv = (task < ntasks && k < k_limit) ? (n_limit - n) : 0;
            nv = nv > 8 ? 8 : nv;
            load_task(s, (long long)k * s.ld + n, nv, k, td[u]);
        }
#pragma unroll
        for (int u = 0; u < PBATCH; ++u) {
            const int task = base + u * PROD_THREADS;
            if (task < ntasks) {
                const int kk = task / chunks, c = task - kk * chunks;
                float v[8];
                finish_task(td[u], v);
                uint32_t off = mn_sw128_offset((uint32_t)kk, (uint32_t)c, B_MN_LBO, B_MN_SBO);
                if (nhl == 2) {
                    uint4 h, l;
                    split_bf16x8(v, h, l);
                    *reinterpret_cast<uint4*>(hi + off) = h;
                    *reinterpret_cast<uint4*>(lo + off) = l;
                } else {
                    *reinterpret_cast<uint4*>(hi + off) = pack_bf16x8(v);
                }
            }
        }
    }
}

// grouped rows [Nd][4] (x - cx, y - cy, z - cz, feature): K = 4, zero-padded to one 16-wide MMA step.
__device__ __forceinline__ void produce_xt4(const OperandSrc& s, uint8_t* hi, uint8_t* lo, int nhl, int nrows, int n0, int n_limit,
                                            int ptid) {
    for (int r = ptid; r < nrows; r += PROD_THREADS) {
        int n = n0 + r;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (n < n_limit) {
            float4 t = __ldg(reinterpret_cast<const float4*>(s.src0) + n);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        store_chunk(hi, lo, nhl, r, 0, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = 0.f;
        store_chunk(hi, lo, nhl, r, 1, v);
    }
}

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const GemmParams p) {
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer (LDS/STS, not generic LD/ST)
    const int nhl = (p.nsplit == 3) ? 2 : 1;
    const int stage_bytes = (A_TILE_BYTES + B_TILE_BYTES) * nhl;
    const int num_stages = (p.nsplit == 3) ? 2 : 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + num_stages * stage_bytes);
    uint64_t* empty = full + num_stages;
    uint64_t* acc_full = empty + num_stages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mma_n = (p.Nd >= N_TILE) ? N_TILE : ((p.Nd + 15) & ~15);

    if (threadIdx.x == 0) {
        for (int i = 0; i < num_stages; ++i) {
            mbar_init(&full[i], NUM_PROD_WARPS + (p.a_mode == A_PACKED ? 1 : 0));
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 10) {
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
        const int c = sched.mt * M_TILE + warp * 32 + lane;
        const bool cvalid = c < p.Md;
        const float bias = (cvalid && p.bias) ? __ldg(p.bias + c) : 0.f;
        const float zs0 = (cvalid && p.zs0) ? __ldg(p.zs0 + c) : 1.f;
        const float zs2 = (cvalid && p.zs2) ? __ldg(p.zs2 + c) : 0.f;
        const float psign = (cvalid && p.pool_sign) ? __ldg(p.pool_sign + c) : 1.f;
        const bool keep_max = psign >= 0.f;
        float stat0 = 0.f, stat1 = 0.f;
        for (int it = 0; sched.get(it, p, w); ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
            float best = 0.f;
            int barg = 0;
            const int nchunks = (mma_n + 31) / 32;
            for (int cc = 0; cc < nchunks; ++cc) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N_TILE + cc * 32), v);
                tmem_ld_wait();
                const int n0 = w.nt * N_TILE + cc * 32;
                int nvalid = p.Nd - n0;
                nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
                if (cvalid && nvalid > 0) {
                    float z[32];
                    if (p.zin) {
                        const float* zr = p.zin + (long long)c * p.ldz + n0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(zr) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                float4 t = __ldg(reinterpret_cast<const float4*>(zr) + q);
                                z[4 * q] = t.x; z[4 * q + 1] = t.y; z[4 * q + 2] = t.z; z[4 * q + 3] = t.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) z[i] = (i < nvalid) ? __ldg(zr + i) : 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float val = v[i] + bias;
                        if (p.zin) {
                            val = (fmaf(zs0, z[i], zs2) > 0.f) ? val : 0.f;
                        }
                        if (i < nvalid) {
                            stat0 += val;
                            stat1 = fmaf(val, p.zin ? z[i] : val, stat1);
                        }
                        v[i] = val;
                    }
                    if (p.pool) {
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
                    if (p.out_mode == OUT_CHMAJOR) {
                        float* o = p.out + (long long)c * p.ldo + n0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nvalid) o[i] = v[i];
                        }
                    } else if (p.out_mode == OUT_ROWMAJOR) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) p.out[(long long)(n0 + i) * p.ldo + c] = v[i];
                    } else if (p.out_mode == OUT_ROWMAJOR_ACC) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) p.out[(long long)(n0 + i) * p.ldo + c] += v[i];
                    } else if (p.out_mode == OUT_ATOMIC_CHMAJOR) {
                        float* o = p.out + (long long)c * p.ldo + n0;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) atomicAdd(o + i, v[i]);
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (p.stats && cvalid) {
            int pidx = sched.split ? 0 : (blockIdx.x / sched.numMT);
            float* st = p.stats + ((long long)pidx * p.Md + c) * 2;
            st[0] = stat0;
            st[1] = stat1;
        }
    } else if (warp < 10) {
        // =============================== producers ===============================
        const int ptid = threadIdx.x - 128;
        int stage = 0, phase = 0;
        for (int it = 0; sched.get(it, p, w); ++it) {
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * stage_bytes;
                uint8_t* a_hi = st;
                uint8_t* a_lo = st + A_TILE_BYTES;
                uint8_t* b_hi = st + A_TILE_BYTES * nhl;
                uint8_t* b_lo = b_hi + B_TILE_BYTES;
                if (p.a_mode == A_ROWMAJOR)
                    produce_rowmajor(p.a, a_hi, a_lo, nhl, M_TILE, w.mt * M_TILE, p.Md, kb * K_BLK, p.Kd, ptid);
                if (p.b_mode == B_ROWMAJOR)
                    produce_rowmajor(p.b, b_hi, b_lo, nhl, mma_n, w.nt * N_TILE, p.Nd, kb * K_BLK, p.Kd, ptid);
                else if (p.b_mode == B_CHMAJOR)
                    produce_chmajor_mn(p.b, b_hi, b_lo, nhl, mma_n, w.nt * N_TILE, p.Nd, kb * K_BLK, p.Kd, ptid);
                else
                    produce_xt4(p.b, b_hi, b_lo, nhl, mma_n, w.nt * N_TILE, p.Nd, ptid);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
                if (++stage == num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 10) {
        // =============================== MMA issuer ===============================
        const bool b_mn = (p.b_mode == B_CHMAJOR);
        const uint32_t idesc = umma_idesc_bf16(M_TILE, mma_n) | (b_mn ? UMMA_B_MN_MAJOR : 0u);
        int stage = 0, phase = 0;
        for (int it = 0; sched.get(it, p, w); ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after_sync();
                if (elect_one_sync()) {            // elect.sync: ptxas then emits the bare UTCHMMA / UTCBAR (umma.cuh, lean issue path)
                    uint8_t* st = smem + stage * stage_bytes;
                    const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + A_TILE_BYTES * nhl, b_lo = b_hi + B_TILE_BYTES;
                    int kleft = p.Kd - kb * K_BLK;
                    int ksteps = kleft >= K_BLK ? 4 : (kleft + 15) / 16;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t acc = (kb > w.kb0 || ks > 0) ? 1u : 0u;
                        const uint32_t ko = ks * 32;   // 16 bf16 = 32 bytes inside the 128-byte swizzled row
                        // MN-major B: one k-step = two 8-k groups = 2 * SBO bytes
                        const uint64_t bd_hi = b_mn ? umma_desc_mn_sw128(b_hi + ks * 2 * B_MN_SBO, B_MN_LBO, B_MN_SBO)
                                                    : umma_desc_sw128(b_hi + ko);
                        umma_bf16_ss(d_tmem, umma_desc_sw128(a_hi + ko), bd_hi, idesc, acc);
                        if (nhl == 2) {
                            const uint64_t bd_lo = b_mn ? umma_desc_mn_sw128(b_lo + ks * 2 * B_MN_SBO, B_MN_LBO, B_MN_SBO)
                                                        : umma_desc_sw128(b_lo + ko);
                            umma_bf16_ss(d_tmem, umma_desc_sw128(a_hi + ko), bd_lo, idesc, 1u);
                            umma_bf16_ss(d_tmem, umma_desc_sw128(a_lo + ko), bd_hi, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (kb == w.kb1 - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++stage == num_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // =============================== TMA (packed A) ===============================
        if (p.a_mode == A_PACKED && lane == 0) {
            int stage = 0, phase = 0;
            const uint8_t* img = reinterpret_cast<const uint8_t*>(p.a_packed);
            for (int it = 0; sched.get(it, p, w); ++it) {
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * stage_bytes;
                    const uint8_t* src = img + ((long long)w.mt * p.a_packed_kblocks + kb) * (2ll * A_TILE_BYTES);
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(A_TILE_BYTES * nhl));
                    tma_bulk_g2s(st, src, A_TILE_BYTES, &full[stage]);
                    if (nhl == 2) tma_bulk_g2s(st + A_TILE_BYTES, src + A_TILE_BYTES, A_TILE_BYTES, &full[stage]);
                    if (++stage == num_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int gemm_tc_ctas_per_mtile(int Md, int Nd) {
    int numMT = (Md + M_TILE - 1) / M_TILE;
    int numNT = (Nd + N_TILE - 1) / N_TILE;
    int P = NUM_SMS / numMT;
    if (P < 1) P = 1;
    if (P > numNT) P = numNT;
    return P;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t stream) {
    if (p.a_mode == A_IMAGE || p.b_mode == B_IMAGE_MN || p.b_mode == B_IMAGE_K) return launch_gemm_img(p, stream);
    static DeviceOnce configured;
    const int smem_bytes = 4 * (A_TILE_BYTES + B_TILE_BYTES) + 1024 + 256;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured.done();
    }
    if (p.Md <= 0 || p.Nd <= 0 || p.Kd <= 0) return (int)cudaErrorInvalidValue;
    if (p.nsplit != 1 && p.nsplit != 3) return (int)cudaErrorInvalidValue;
    if (p.pool && ((p.pool & (p.pool - 1)) != 0 || p.pool > N_TILE)) return (int)cudaErrorInvalidValue;
    int numMT = (p.Md + M_TILE - 1) / M_TILE;
    int numNT = (p.Nd + N_TILE - 1) / N_TILE;
    int KB = (p.Kd + K_BLK - 1) / K_BLK;
    int grid;
    if (p.ksplit > 1) {
        if (p.ksplit > KB || p.stats || p.pool) return (int)cudaErrorInvalidValue;
        grid = numMT * numNT * p.ksplit;
    } else {
        grid = numMT * gemm_tc_ctas_per_mtile(p.Md, p.Nd);
    }
    ScopedTimer timer(p.tag, stream);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(gemm_tc_kernel, dim3(grid), dim3(THREADS), smem_bytes, stream, p));
    return (int)cudaGetLastError();
}

}  // namespace facl
