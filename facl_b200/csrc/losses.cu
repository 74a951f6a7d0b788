// K6/K7: the "global" and "circle" contrastive losses, forward + gradient in one call.
//
// Replaces the inline losses of reference training_code/cn3d_train_motion_GL.py:265-287 / :290-316
// (= utils_my.py:53-83 global_contrast, :85-116 circle_contrast).  The reference builds G (resp. G-1) logits
// tensors with python loops, `repeat`s the shared negatives and calls CrossEntropyLoss per view.  Here every
// needed dot product is an entry of two similarity GEMMs (tcgen05):
//
//     S_x = x x^T   (M x M, M = G*B)          S_g = x_global x^T   (B x M)
//
// * negatives of anchor row a: all S[a][j] with (j mod B) != n(a); the G "masked" entries (j mod B == n) enter
//   the softmax denominator as exp(0) each (the reference multiplies them by 0, utils_my.py:72,106);
// * the positives are themselves masked entries: global  pos[g][n] = S_g[n][g*B+n],
//   circle pos[i][n] = S_x[o_i*B+n][o_{i+1}*B+n];
// * circle: the G-1 anchor rows of sample n share one negative set (their union), utils_my.py:105-109.
// S is never materialised.  Forward: the similarity GEMM's TMEM epilogue keeps an ONLINE log-sum-exp per anchor row
// (running max and sum of exponentials over the unmasked columns, merged across the CTAs that share a row) and stores only the
// G masked entries of each row (the positives are among them); a finalize pass forms the log-sum-exps, the two losses and the
// per-sample softmax coefficients.  Backward: the similarity GEMM runs again and its epilogue turns S into dL/dS in registers,
// written directly as a bf16 hi/lo operand image; two GEMMs per loss then give dX.  (gemm_img.cu, LossEpi)
#include <math.h>
#include <string.h>

#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"
#include "gemm_tc.cuh"

namespace facl {

namespace {

// Index conventions (single- and multi-GPU).  The batch of B sequences is sharded over R ranks, Bl = B / R each.
// Keys (the columns of S) are ALL ranks' view embeddings in rank-major order, which is what an all-gather of the
// per-rank x (G-major inside a rank) produces:   key j = r*Ml + g*Bl + b   (Ml = G*Bl)  <->  view g of sample
// n = r*Bl + b.  Anchors (the rows of S) are this rank's Ml view embeddings followed by its Bl sequence embeddings;
// local anchor row a < Ml is view a / Bl of sample n0 + a % Bl, row Ml + b is sample n0 + b.  With R = 1 this is the
// reference's own row order (g*B + n).
struct Idx {
    int G, B, Bl, Ml, Mk, n0;
    __host__ __device__ int key_sample(int j) const {
        int r = j / Ml, rem = j - r * Ml;
        return r * Bl + (rem % Bl);
    }
    __host__ __device__ int key_view(int j) const { return (j % Ml) / Bl; }
    __host__ __device__ int key_of(int g, int n) const { return (n / Bl) * Ml + g * Bl + (n % Bl); }
    __host__ __device__ int anchor_sample(int a) const { return n0 + (a < Ml ? a % Bl : a - Ml); }
};

__device__ __forceinline__ double lse3(double pos, double m, double e, double nzero) {
    // log(exp(pos) + e*exp(m) + nzero*exp(0)), with e possibly 0 (m = -inf)
    double mx = fmax(pos, 0.0);
    if (e > 0.0) mx = fmax(mx, m);
    double s = exp(pos - mx) + nzero * exp(-mx);
    if (e > 0.0) s += e * exp(m - mx);
    return mx + log(s);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// single block of 16 warps; warp per LOCAL sample b, lanes over the views (the f64 exp / log chain per (b, view) is the
// cost: spread over lanes it is ~10x shorter than a thread per sample).  Produces loss[0] = global, loss[1] = circle
// (this rank's anchors, already divided by the global B) and the softmax coefficients:
//   lcG[b] = log sum_g exp(-LSE_g,n)      pgG[b*G+g] = (exp(pos - LSE) - 1)/B       (global)
//   lcC[b] = log sum_i exp(-LSE_i,n)      pgC[b*G+i] = (exp(pos - LSE) - 1)/B       (circle, i < G-1)
constexpr int LF_THREADS = 512;
__global__ void __launch_bounds__(LF_THREADS) loss_finalize_kernel(const float* __restrict__ pos, Idx ix, const int* __restrict__ order,
                                                                   const float* __restrict__ partX, int PX, const float* __restrict__ partG,
                                                                   int PG, float* rmax, float* rsum,
                                                                   int want_global, int want_circle, float* __restrict__ loss,
                                                                   float* __restrict__ lcG, float* __restrict__ pgG,
                                                                   float* __restrict__ lcC, float* __restrict__ pgC) {
    pdl_prologue();
    const int G = ix.G, B = ix.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // merge the per-CTA online log-sum-exp partials of every anchor row: m = max_p m_p, e = sum_p e_p exp(m_p - m)
    for (int a = threadIdx.x; a < ix.Ml + ix.Bl; a += LF_THREADS) {
        const bool seq = a >= ix.Ml;
        if ((seq && !want_global) || (!seq && !want_circle)) continue;
        const float* part = seq ? partG : partX;
        const int P = seq ? PG : PX, Md = seq ? ix.Bl : ix.Ml, r = seq ? a - ix.Ml : a;
        float m = -INFINITY;
        for (int q = 0; q < P; ++q)
            if (part[((long long)q * Md + r) * 2 + 1] > 0.f) m = fmaxf(m, part[((long long)q * Md + r) * 2]);
        float e = 0.f;
        for (int q = 0; q < P; ++q) {
            const float eq = part[((long long)q * Md + r) * 2 + 1];
            if (eq > 0.f) e += eq * expf(part[((long long)q * Md + r) * 2] - m);
        }
        rmax[a] = m;
        rsum[a] = e;
    }
    __syncthreads();
    double accG = 0.0, accC = 0.0;
    for (int b = warp; b < ix.Bl; b += LF_THREADS / 32) {
        const int n = ix.n0 + b;
        if (want_global) {
            const int a = ix.Ml + b;
            const double m = rmax[a], e = rsum[a];
            double minL = 1e300;
            for (int g = lane; g < G; g += 32) {
                double pos_v = pos[(long long)a * G + g];
                double L = lse3(pos_v, m, e, (double)G);
                accG += L - pos_v;
                pgG[b * G + g] = (float)((exp(pos_v - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            minL = warp_min_d(minL);
            double s = 0.0;
            for (int g = lane; g < G; g += 32) {
                double pos_v = pos[(long long)a * G + g];
                s += exp(minL - lse3(pos_v, m, e, (double)G));
            }
            s = warp_sum_d(s);
            if (lane == 0) lcG[b] = (float)(-minL + log(s));
        }
        if (want_circle) {
            double mx = -INFINITY;
            for (int i = lane; i < G - 1; i += 32) {
                int a = order[i] * ix.Bl + b;
                if (rsum[a] > 0.f) mx = fmax(mx, (double)rmax[a]);
            }
            mx = warp_max_d(mx);
            double e = 0.0;
            for (int i = lane; i < G - 1; i += 32) {
                int a = order[i] * ix.Bl + b;
                if (rsum[a] > 0.f) e += (double)rsum[a] * exp((double)rmax[a] - mx);
            }
            e = warp_sum_d(e);
            const double nzero = (double)(G - 1) * G;
            double minL = 1e300;
            for (int i = lane; i < G - 1; i += 32) {
                int a = order[i] * ix.Bl + b;
                double pos_v = pos[(long long)a * G + order[i + 1]];
                double L = lse3(pos_v, mx, e, nzero);
                accC += L - pos_v;
                pgC[b * G + i] = (float)((exp(pos_v - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            minL = warp_min_d(minL);
            double s = 0.0;
            for (int i = lane; i < G - 1; i += 32) {
                int a = order[i] * ix.Bl + b;
                double pos_v = pos[(long long)a * G + order[i + 1]];
                s += exp(minL - lse3(pos_v, mx, e, nzero));
            }
            s = warp_sum_d(s);
            if (lane == 0) lcC[b] = (G > 1) ? (float)(-minL + log(s)) : -INFINITY;
        }
    }
    __shared__ double sh[2][LF_THREADS / 32];
    accG = warp_sum_d(accG);
    accC = warp_sum_d(accC);
    if (lane == 0) {
        sh[0][warp] = accG;
        sh[1][warp] = accC;
    }
    __syncthreads();
    if (warp == 0) {
        double g = warp_sum_d(lane < LF_THREADS / 32 ? sh[0][lane] : 0.0), c = warp_sum_d(lane < LF_THREADS / 32 ? sh[1][lane] : 0.0);
        if (lane == 0) {
            loss[0] = want_global ? (float)(g / B) : 0.f;
            loss[1] = want_circle ? (float)(c / B) : 0.f;
        }
    }
}

__global__ void invert_order_kernel(const int* __restrict__ order, int G, int* __restrict__ inv) {
    pdl_prologue();
    int i = threadIdx.x;
    if (i < G) inv[order[i]] = i;
}

struct LossWs {
    float *part, *pos, *rmax, *rsum, *lcG, *pgG, *lcC, *pgC;
    int* inv;
    uint8_t *img_keys, *img_x, *img_xg;
    uint8_t *im_x, *im_xg, *im_keys, *im_ds;      // activation images (gemm_img.cu): embeddings [rows as channels][C], dS [rows][Mk]
};
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
size_t loss_ws_layout(int G, int Bl, int R, int C, uint8_t* base, LossWs* w) {
    size_t Ml = (size_t)G * Bl, Mk = Ml * R, rows = Ml + Bl, off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    uint8_t* pS = take((size_t)kNumSMs * rows * 2 * 4);     // online log-sum-exp partials: at most one per (CTA, anchor row)
    uint8_t* pP = take(rows * G * 4);                       // the G masked entries of every anchor row
    uint8_t* p1 = take(rows * 4);
    uint8_t* p2 = take(rows * 4);
    uint8_t* p3 = take((size_t)Bl * 4);
    uint8_t* p4 = take((size_t)Bl * G * 4);
    uint8_t* p5 = take((size_t)Bl * 4);
    uint8_t* p6 = take((size_t)Bl * G * 4);
    uint8_t* p7 = take((size_t)G * 4);
    uint8_t* p8 = take(packed_weight_bytes(C, (int)Mk));
    uint8_t* p9 = take(packed_weight_bytes(C, (int)Ml));
    uint8_t* p10 = take(packed_weight_bytes(C, Bl));
    uint8_t* p11 = take(2 * act_image_half_bytes((int)Ml, C));
    uint8_t* p12 = take(2 * act_image_half_bytes(Bl, C));
    uint8_t* p13 = take(2 * act_image_half_bytes((int)Mk, C));
    uint8_t* p14 = take(2 * act_image_half_bytes((int)Ml, (long long)Mk));
    if (w) {
        w->im_x = p11; w->im_xg = p12; w->im_keys = p13; w->im_ds = p14;
        w->part = (float*)pS; w->pos = (float*)pP; w->rmax = (float*)p1; w->rsum = (float*)p2; w->lcG = (float*)p3; w->pgG = (float*)p4;
        w->lcC = (float*)p5; w->pgC = (float*)p6; w->inv = (int*)p7; w->img_keys = p8; w->img_x = p9; w->img_xg = p10;
    }
    return off;
}

#define RUN(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != 0) return _rc;      \
    } while (0)

}  // namespace

// x [G*Bl][C], xg [Bl][C]: this rank's embeddings; keys [R*G*Bl][C]: every rank's x, rank-major (== x when R == 1).
int contrast_losses(const float* x, const float* xg, const float* keys, int G, int B, int Bl, int n0, int C, const int* order,
                    int want_global, int want_circle, int nsplit, void* workspace, float* loss, float* dx_anchor, float* dxg,
                    float* dkeys, cudaStream_t st) {
    if (G <= 0 || B <= 0 || Bl <= 0 || B % Bl != 0 || n0 < 0 || n0 + Bl > B || C <= 0 || (C & 3)) return (int)cudaErrorInvalidValue;
    if (!x || !keys || !workspace || !loss || !dx_anchor || !dkeys) return (int)cudaErrorInvalidValue;
    if (want_global && (!xg || !dxg)) return (int)cudaErrorInvalidValue;
    if (want_circle && (!order || G < 2)) return (int)cudaErrorInvalidValue;
    const int R = B / Bl, Ml = G * Bl, Mk = Ml * R;
    Idx ix{G, B, Bl, Ml, Mk, n0};
    LossWs w;
    loss_ws_layout(G, Bl, R, C, reinterpret_cast<uint8_t*>(workspace), &w);
    count_launch(1 + (want_circle ? 1 : 0));                                   // the small loss kernels below

    const int nhl = nsplit == 3 ? 2 : 1;
    auto image = [&](uint8_t* buf, int ch, long long rows) {
        ActImage im;
        im.hi = buf;
        im.lo = buf + act_image_half_bytes(ch, rows);
        im.cgs = ((ch + 63) / 64) * 8;
        im.rbs = (int)((rows + 63) / 64);
        return im;
    };
    // a row-major matrix [rows][ld] is the "channel-major" source of an image whose channels are its rows
    auto make_image = [&](const float* src, int rows, long long cols, uint8_t* buf) {
        OperandSrc o;
        memset(&o, 0, sizeof(o));
        o.src0 = src; o.ld = cols;
        return act_image_launch(o, 0, rows, cols, nullptr, 0, nhl, image(buf, rows, cols), TAG_LOSS_MISC, st);
    };
    const bool keys_are_x = (keys == x) && (R == 1);     // single rank: one image serves as anchors and keys
    const ActImage im_keys = image(keys_are_x ? w.im_x : w.im_keys, Mk, C);
    // similarity GEMM S = a keys^T (reduction over the C "rows") with a contrastive-loss epilogue; S is never stored.
    //   mode 1: online log-sum-exp partials + the masked entries;  mode 2: dL/dS written as the operand image `ds`
    float* partX = w.part;
    float* partG = w.part + (size_t)kNumSMs * Ml * 2;
    auto sim = [&](const ActImage& a, int rows, int row0, int mode, const ActImage* ds) {
        GemmParams g;
        memset(&g, 0, sizeof(g));
        g.Md = rows; g.Nd = Mk; g.Kd = C; g.nsplit = nsplit; g.ksplit = 1;
        g.a_mode = A_IMAGE; g.a_img = a;
        g.b_mode = B_IMAGE_K; g.b_img = im_keys;
        g.out_mode = OUT_NONE;
        g.tag = TAG_LOSS_GEMM;
        LossEpi& L = g.loss;
        L.mode = mode; L.G = G; L.B = B; L.Bl = Bl; L.Ml = Ml; L.n0 = n0; L.row0 = row0;
        if (mode == 1) {
            L.part = row0 == 0 ? partX : partG;
            L.pos = w.pos;
        } else {
            L.lc = row0 == 0 ? w.lcC : w.lcG;
            L.pg = row0 == 0 ? w.pgC : w.pgG;
            L.order = order; L.inv_order = w.inv;
            L.ds_hi = const_cast<void*>(ds->hi); L.ds_lo = const_cast<void*>(ds->lo); L.ds_cgs = ds->cgs; L.ds_rbs = ds->rbs;
        }
        return launch_gemm_tc(g, st);
    };
    // out[Nd rows][C] (+)= dS-block * features, with the feature matrix (transposed) as the packed "A" operand and the
    // dS block [rows][Mk] as an image: reduced over its rows-as-channels (b_mode MN) or over its Mk columns (b_mode K)
    auto dgemm = [&](const uint8_t* img, int Kd, int b_mode, const ActImage& ds, int Nd, float* out) {
        GemmParams g;
        memset(&g, 0, sizeof(g));
        g.Md = C; g.Nd = Nd; g.Kd = Kd; g.nsplit = nsplit;
        // few output tiles, long reductions (Kd grows with the number of ranks): split K over the idle SMs, reduce with atomics
        const int tiles = ((C + 127) / 128) * ((Nd + 255) / 256), KB = (Kd + 63) / 64;
        int ks = kNumSMs / tiles;
        ks = ks > KB ? KB : ks;
        g.ksplit = ks < 1 ? 1 : ks;
        g.a_mode = A_PACKED; g.a_packed = img; g.a_packed_kblocks = KB;
        g.b_mode = b_mode; g.b_img = ds;
        g.out_mode = OUT_ATOMIC_ROWMAJOR; g.out = out; g.ldo = C;
        g.tag = TAG_LOSS_GEMM;
        return launch_gemm_tc(g, st);
    };

    if (want_circle || keys_are_x) RUN(make_image(x, Ml, C, w.im_x));
    if (!keys_are_x) RUN(make_image(keys, Mk, C, w.im_keys));
    if (want_circle) {
        RUN(sim(image(w.im_x, Ml, C), Ml, 0, 1, nullptr));
        FACL_LAUNCH_OK(launch_pdl(invert_order_kernel, dim3(1), dim3(256), 0, st, order, G, w.inv));
    }
    if (want_global) {
        RUN(make_image(xg, Bl, C, w.im_xg));
        RUN(sim(image(w.im_xg, Bl, C), Bl, Ml, 1, nullptr));
    }
    FACL_LAUNCH_OK(launch_pdl(loss_finalize_kernel, dim3(1), dim3(LF_THREADS), 0, st, w.pos, ix, order, partX, gemm_tc_ctas_per_mtile(Ml, Mk), partG,
                                                   gemm_tc_ctas_per_mtile(Bl, Mk), w.rmax, w.rsum, want_global, want_circle, loss, w.lcG,
                                                   w.pgG, w.lcC, w.pgC));
    FACL_CHECK_LAUNCH();
    // feature matrices as packed A operands: A[m = c][k = row] = feat[row][c]
    RUN(pack_weight_launch(keys, 1, C, C, Mk, w.img_keys, st));
    // every gradient GEMM accumulates into zeroed outputs (dx_anchor and dkeys may be the same buffer on a single GPU)
    FACL_CHECK(cudaMemsetAsync(dx_anchor, 0, sizeof(float) * (size_t)Ml * C, st));
    if (dkeys != dx_anchor) FACL_CHECK(cudaMemsetAsync(dkeys, 0, sizeof(float) * (size_t)Mk * C, st));
    if (want_global) {
        FACL_CHECK(cudaMemsetAsync(dxg, 0, sizeof(float) * (size_t)Bl * C, st));
        RUN(pack_weight_launch(xg, 1, C, C, Bl, w.img_xg, st));
        // the similarity GEMM again, its epilogue writing dL/dS_g straight into the operand image
        const ActImage ds = image(w.im_ds, Bl, Mk);
        RUN(sim(image(w.im_xg, Bl, C), Bl, Ml, 2, &ds));
        // dxg[b] = sum_j dS_g[b][j] keys[j] ;  dkeys[j] += sum_b dS_g[b][j] xg[b]
        RUN(dgemm(w.img_keys, Mk, B_IMAGE_K, ds, Bl, dxg));
        RUN(dgemm(w.img_xg, Bl, B_IMAGE_MN, ds, Mk, dkeys));
    }
    if (want_circle) {
        RUN(pack_weight_launch(x, 1, C, C, Ml, w.img_x, st));
        const ActImage ds = image(w.im_ds, Ml, Mk);
        RUN(sim(image(w.im_x, Ml, C), Ml, 0, 2, &ds));
        // dx_anchor[a] += sum_j dS[a][j] keys[j]   and   dkeys[j] += sum_a dS[a][j] x[a]
        RUN(dgemm(w.img_keys, Mk, B_IMAGE_K, ds, Ml, dx_anchor));
        RUN(dgemm(w.img_x, Ml, B_IMAGE_MN, ds, Mk, dkeys));
    }
    return 0;
}

}  // namespace facl

extern "C" {

size_t facl_contrast_workspace_bytes(int G, int B_local, int world, int C) {
    return facl::loss_ws_layout(G, B_local, world, C, nullptr, nullptr);
}

int facl_contrast_losses(const float* x, const float* x_global, const float* keys, int G, int B, int B_local, int sample_offset,
                         int C, const int* order, int want_global, int want_circle, int nsplit, void* workspace, float* loss,
                         float* dx_anchor, float* dx_global, float* dkeys, void* stream) {
    return facl::contrast_losses(x, x_global, keys ? keys : x, G, B, B_local, sample_offset, C, order, want_global, want_circle,
                                 nsplit, workspace, loss, dx_anchor, dx_global, dkeys, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
