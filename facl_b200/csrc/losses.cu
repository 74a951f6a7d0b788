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
// A row pass computes (max, sum exp) over the unmasked entries, a finalize pass forms the log-sum-exps, the two
// losses and the per-sample softmax coefficients; dS is then written in place and two GEMMs per loss give dX.
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

// one block per anchor row: m = max, e = sum exp(s - m) over the unmasked columns
__global__ void __launch_bounds__(128) loss_rowstats_kernel(const float* __restrict__ Smat, Idx ix, int row0, int nrows,
                                                            float* __restrict__ rmax, float* __restrict__ rsum) {
    int a = row0 + blockIdx.x;
    if (blockIdx.x >= nrows) return;
    int n = ix.anchor_sample(a);
    const float* row = Smat + (long long)a * ix.Mk;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < ix.Mk; j += 128)
        if (ix.key_sample(j) != n) m = fmaxf(m, row[j]);
    __shared__ float sh[128];
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    m = sh[0];
    __syncthreads();
    float e = 0.f;
    if (m > -INFINITY)
        for (int j = threadIdx.x; j < ix.Mk; j += 128)
            if (ix.key_sample(j) != n) e += expf(row[j] - m);
    sh[threadIdx.x] = e;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rmax[a] = m;
        rsum[a] = sh[0];
    }
}

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
__global__ void __launch_bounds__(LF_THREADS) loss_finalize_kernel(const float* __restrict__ Smat, Idx ix, const int* __restrict__ order,
                                                                   const float* __restrict__ rmax, const float* __restrict__ rsum,
                                                                   int want_global, int want_circle, float* __restrict__ loss,
                                                                   float* __restrict__ lcG, float* __restrict__ pgG,
                                                                   float* __restrict__ lcC, float* __restrict__ pgC) {
    const int G = ix.G, B = ix.B;
    const long long ld = ix.Mk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double accG = 0.0, accC = 0.0;
    for (int b = warp; b < ix.Bl; b += LF_THREADS / 32) {
        const int n = ix.n0 + b;
        if (want_global) {
            const int a = ix.Ml + b;
            const double m = rmax[a], e = rsum[a];
            double minL = 1e300;
            for (int g = lane; g < G; g += 32) {
                double pos = Smat[(long long)a * ld + ix.key_of(g, n)];
                double L = lse3(pos, m, e, (double)G);
                accG += L - pos;
                pgG[b * G + g] = (float)((exp(pos - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            minL = warp_min_d(minL);
            double s = 0.0;
            for (int g = lane; g < G; g += 32) {
                double pos = Smat[(long long)a * ld + ix.key_of(g, n)];
                s += exp(minL - lse3(pos, m, e, (double)G));
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
                double pos = Smat[(long long)a * ld + ix.key_of(order[i + 1], n)];
                double L = lse3(pos, mx, e, nzero);
                accC += L - pos;
                pgC[b * G + i] = (float)((exp(pos - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            minL = warp_min_d(minL);
            double s = 0.0;
            for (int i = lane; i < G - 1; i += 32) {
                int a = order[i] * ix.Bl + b;
                double pos = Smat[(long long)a * ld + ix.key_of(order[i + 1], n)];
                s += exp(minL - lse3(pos, mx, e, nzero));
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

// S -> dL/dS in place (unit upstream gradient for each loss)
__global__ void loss_ds_kernel(float* __restrict__ Smat, Idx ix, const int* __restrict__ order, const int* __restrict__ inv_order,
                               int row0, int nrows, const float* __restrict__ lcG, const float* __restrict__ pgG,
                               const float* __restrict__ lcC, const float* __restrict__ pgC) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nrows * ix.Mk) return;
    const int G = ix.G;
    int a = row0 + (int)(t / ix.Mk), j = (int)(t % ix.Mk);
    float* s = Smat + (long long)a * ix.Mk + j;
    const int n = ix.anchor_sample(a);
    const bool same = ix.key_sample(j) == n;
    float out;
    if (a >= ix.Ml) {
        int b = a - ix.Ml;
        if (same) out = pgG[b * G + ix.key_view(j)];
        else out = expf(*s + lcG[b]) / ix.B;
    } else {
        int b = a % ix.Bl, i = inv_order[a / ix.Bl];
        if (i >= G - 1) out = 0.f;                     // the last view in the chain is never an anchor
        else if (same) out = (ix.key_view(j) == order[i + 1]) ? pgC[b * G + i] : 0.f;
        else out = expf(*s + lcC[b]) / ix.B;
    }
    *s = out;
}

__global__ void invert_order_kernel(const int* __restrict__ order, int G, int* __restrict__ inv) {
    int i = threadIdx.x;
    if (i < G) inv[order[i]] = i;
}

struct LossWs {
    float *S, *rmax, *rsum, *lcG, *pgG, *lcC, *pgC;
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
    uint8_t* pS = take(rows * Mk * 4);
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
        w->S = (float*)pS; w->rmax = (float*)p1; w->rsum = (float*)p2; w->lcG = (float*)p3; w->pgG = (float*)p4;
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
    count_launch(2 + 2 * (want_circle ? 2 : 0) + 2 * (want_global ? 1 : 0));   // the small loss kernels below

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
    auto sim = [&](const ActImage& a, int rows, float* out) {   // out[rows][Mk] = a keys^T  (reduction over the C "rows")
        GemmParams g;
        memset(&g, 0, sizeof(g));
        g.Md = rows; g.Nd = Mk; g.Kd = C; g.nsplit = nsplit; g.ksplit = 1;
        g.a_mode = A_IMAGE; g.a_img = a;
        g.b_mode = B_IMAGE_K; g.b_img = im_keys;
        g.out_mode = OUT_CHMAJOR; g.out = out; g.ldo = Mk;
        g.tag = TAG_LOSS_GEMM;
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

    float* Sx = w.S;
    float* Sg = w.S + (size_t)Ml * Mk;
    if (want_circle || keys_are_x) RUN(make_image(x, Ml, C, w.im_x));
    if (!keys_are_x) RUN(make_image(keys, Mk, C, w.im_keys));
    if (want_circle) {
        RUN(sim(image(w.im_x, Ml, C), Ml, Sx));
        invert_order_kernel<<<1, 256, 0, st>>>(order, G, w.inv);
        loss_rowstats_kernel<<<Ml, 128, 0, st>>>(w.S, ix, 0, Ml, w.rmax, w.rsum);
    }
    if (want_global) {
        RUN(make_image(xg, Bl, C, w.im_xg));
        RUN(sim(image(w.im_xg, Bl, C), Bl, Sg));
        loss_rowstats_kernel<<<Bl, 128, 0, st>>>(w.S, ix, Ml, Bl, w.rmax, w.rsum);
    }
    loss_finalize_kernel<<<1, LF_THREADS, 0, st>>>(w.S, ix, order, w.rmax, w.rsum, want_global, want_circle, loss, w.lcG, w.pgG, w.lcC,
                                            w.pgC);
    FACL_CHECK_LAUNCH();
    // feature matrices as packed A operands: A[m = c][k = row] = feat[row][c]
    RUN(pack_weight_launch(keys, 1, C, C, Mk, w.img_keys, st));
    // every gradient GEMM accumulates into zeroed outputs (dx_anchor and dkeys may be the same buffer on a single GPU)
    FACL_CHECK(cudaMemsetAsync(dx_anchor, 0, sizeof(float) * (size_t)Ml * C, st));
    if (dkeys != dx_anchor) FACL_CHECK(cudaMemsetAsync(dkeys, 0, sizeof(float) * (size_t)Mk * C, st));
    if (want_global) {
        FACL_CHECK(cudaMemsetAsync(dxg, 0, sizeof(float) * (size_t)Bl * C, st));
        loss_ds_kernel<<<div_up((long long)Bl * Mk, 256), 256, 0, st>>>(w.S, ix, order, w.inv, Ml, Bl, w.lcG, w.pgG, w.lcC, w.pgC);
        FACL_CHECK_LAUNCH();
        RUN(pack_weight_launch(xg, 1, C, C, Bl, w.img_xg, st));
        // dxg[b] = sum_j dS_g[b][j] keys[j] ;  dkeys[j] += sum_b dS_g[b][j] xg[b]
        RUN(make_image(Sg, Bl, Mk, w.im_ds));
        RUN(dgemm(w.img_keys, Mk, B_IMAGE_K, image(w.im_ds, Bl, Mk), Bl, dxg));
        RUN(dgemm(w.img_xg, Bl, B_IMAGE_MN, image(w.im_ds, Bl, Mk), Mk, dkeys));
    }
    if (want_circle) {
        loss_ds_kernel<<<div_up((long long)Ml * Mk, 256), 256, 0, st>>>(w.S, ix, order, w.inv, 0, Ml, w.lcG, w.pgG, w.lcC, w.pgC);
        FACL_CHECK_LAUNCH();
        RUN(pack_weight_launch(x, 1, C, C, Ml, w.img_x, st));
        // dx_anchor[a] += sum_j dS[a][j] keys[j]   and   dkeys[j] += sum_a dS[a][j] x[a]
        RUN(make_image(Sx, Ml, Mk, w.im_ds));
        RUN(dgemm(w.img_keys, Mk, B_IMAGE_K, image(w.im_ds, Ml, Mk), Ml, dx_anchor));
        RUN(dgemm(w.img_x, Ml, B_IMAGE_MN, image(w.im_ds, Ml, Mk), Mk, dkeys));
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
