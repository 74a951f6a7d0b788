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

// one block per anchor row: m = max, e = sum exp(s - m) over the unmasked columns
__global__ void __launch_bounds__(128) loss_rowstats_kernel(const float* __restrict__ Smat, long long ld, int M, int B, int row0,
                                                            int nrows, float* __restrict__ rmax, float* __restrict__ rsum) {
    int a = row0 + blockIdx.x;
    if (blockIdx.x >= nrows) return;
    int n = (a < M) ? (a % B) : (a - M);
    const float* row = Smat + (long long)a * ld;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < M; j += 128)
        if (j % B != n) m = fmaxf(m, row[j]);
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
        for (int j = threadIdx.x; j < M; j += 128)
            if (j % B != n) e += expf(row[j] - m);
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

// single block; thread per sample n.  Produces loss[0] = global, loss[1] = circle and the softmax coefficients:
//   lcG[n] = log sum_g exp(-LSE_g,n)      pgG[n*G+g] = (exp(pos - LSE) - 1)/B       (global)
//   lcC[n] = log sum_i exp(-LSE_i,n)      pgC[n*G+i] = (exp(pos - LSE) - 1)/B       (circle, i < G-1)
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ Smat, long long ld, int M, int B, int G,
                                                            const int* __restrict__ order, const float* __restrict__ rmax,
                                                            const float* __restrict__ rsum, int want_global, int want_circle,
                                                            float* __restrict__ loss, float* __restrict__ lcG, float* __restrict__ pgG,
                                                            float* __restrict__ lcC, float* __restrict__ pgC) {
    double accG = 0.0, accC = 0.0;
    for (int n = threadIdx.x; n < B; n += 256) {
        if (want_global) {
            int a = M + n;
            double m = rmax[a], e = rsum[a];
            double minL = 1e300;
            for (int g = 0; g < G; ++g) {
                double pos = Smat[(long long)a * ld + (long long)g * B + n];
                double L = lse3(pos, m, e, (double)G);
                accG += L - pos;
                pgG[n * G + g] = (float)((exp(pos - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            double s = 0.0;
            for (int g = 0; g < G; ++g) {
                double pos = Smat[(long long)a * ld + (long long)g * B + n];
                s += exp(minL - lse3(pos, m, e, (double)G));
            }
            lcG[n] = (float)(-minL + log(s));
        }
        if (want_circle) {
            double mx = -INFINITY;
            for (int i = 0; i < G - 1; ++i) {
                int a = order[i] * B + n;
                if (rsum[a] > 0.f) mx = fmax(mx, (double)rmax[a]);
            }
            double e = 0.0;
            for (int i = 0; i < G - 1; ++i) {
                int a = order[i] * B + n;
                if (rsum[a] > 0.f) e += (double)rsum[a] * exp((double)rmax[a] - mx);
            }
            double nzero = (double)(G - 1) * G;
            double minL = 1e300;
            for (int i = 0; i < G - 1; ++i) {
                int a = order[i] * B + n;
                double pos = Smat[(long long)a * ld + (long long)order[i + 1] * B + n];
                double L = lse3(pos, mx, e, nzero);
                accC += L - pos;
                pgC[n * G + i] = (float)((exp(pos - L) - 1.0) / B);
                minL = fmin(minL, L);
            }
            double s = 0.0;
            for (int i = 0; i < G - 1; ++i) {
                int a = order[i] * B + n;
                double pos = Smat[(long long)a * ld + (long long)order[i + 1] * B + n];
                s += exp(minL - lse3(pos, mx, e, nzero));
            }
            lcC[n] = (G > 1) ? (float)(-minL + log(s)) : -INFINITY;
        }
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = accG;
    sh[1][threadIdx.x] = accC;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        loss[0] = want_global ? (float)(sh[0][0] / B) : 0.f;
        loss[1] = want_circle ? (float)(sh[1][0] / B) : 0.f;
    }
}

// S -> dL/dS in place (unit upstream gradient for each loss)
__global__ void loss_ds_kernel(float* __restrict__ Smat, long long ld, int M, int B, int G, const int* __restrict__ order,
                               const int* __restrict__ inv_order, int row0, int nrows, const float* __restrict__ lcG,
                               const float* __restrict__ pgG, const float* __restrict__ lcC, const float* __restrict__ pgC) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nrows * M) return;
    int a = row0 + (int)(t / M), j = (int)(t % M);
    float* s = Smat + (long long)a * ld + j;
    float out;
    if (a >= M) {
        int n = a - M;
        if (j % B == n) out = pgG[n * G + j / B];
        else out = expf(*s + lcG[n]) / B;
    } else {
        int n = a % B, i = inv_order[a / B];
        if (i >= G - 1) out = 0.f;                     // the last view in the chain is never an anchor
        else if (j % B == n) out = (j / B == order[i + 1]) ? pgC[n * G + i] : 0.f;
        else out = expf(*s + lcC[n]) / B;
    }
    *s = out;
}

__global__ void invert_order_kernel(const int* __restrict__ order, int G, int* __restrict__ inv) {
    int i = threadIdx.x;
    if (i < G) inv[order[i]] = i;
}

struct LossWs {
    float *S, *xall, *rmax, *rsum, *lcG, *pgG, *lcC, *pgC;
    int* inv;
    uint8_t *img_x, *img_xall;
};
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
size_t loss_ws_layout(int G, int B, int C, uint8_t* base, LossWs* w) {
    size_t M = (size_t)G * B, MB = M + B, off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    uint8_t* pS = take(MB * M * 4);
    uint8_t* pX = take(MB * C * 4);
    uint8_t* p1 = take(MB * 4);
    uint8_t* p2 = take(MB * 4);
    uint8_t* p3 = take((size_t)B * 4);
    uint8_t* p4 = take((size_t)B * G * 4);
    uint8_t* p5 = take((size_t)B * 4);
    uint8_t* p6 = take((size_t)B * G * 4);
    uint8_t* p7 = take((size_t)G * 4);
    uint8_t* p8 = take(packed_weight_bytes(C, (int)M));
    uint8_t* p9 = take(packed_weight_bytes(C, (int)MB));
    if (w) {
        w->S = (float*)pS; w->xall = (float*)pX; w->rmax = (float*)p1; w->rsum = (float*)p2; w->lcG = (float*)p3;
        w->pgG = (float*)p4; w->lcC = (float*)p5; w->pgC = (float*)p6; w->inv = (int*)p7; w->img_x = p8; w->img_xall = p9;
    }
    return off;
}

#define RUN(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != 0) return _rc;      \
    } while (0)

}  // namespace

int contrast_losses(const float* x, const float* xg, int G, int B, int C, const int* order, int want_global, int want_circle,
                    int nsplit, void* workspace, float* loss, float* dx_global_part, float* dxg, float* dx_circle_part,
                    cudaStream_t st) {
    if (G <= 0 || B <= 0 || C <= 0 || (C & 3) || !x || !workspace || !loss) return (int)cudaErrorInvalidValue;
    if (want_global && (!xg || !dx_global_part || !dxg)) return (int)cudaErrorInvalidValue;
    if (want_circle && (!order || !dx_circle_part || G < 2)) return (int)cudaErrorInvalidValue;
    const int M = G * B, MB = M + B;
    LossWs w;
    loss_ws_layout(G, B, C, reinterpret_cast<uint8_t*>(workspace), &w);
    count_launch(2 + 2 * (want_circle ? 2 : 0) + 2 * (want_global ? 1 : 0));   // the small loss kernels below

    auto sim = [&](const float* a, int rows, float* out) {   // out[rows][M] = a x^T
        GemmParams g;
        memset(&g, 0, sizeof(g));
        g.Md = rows; g.Nd = M; g.Kd = C; g.nsplit = nsplit; g.ksplit = 1;
        g.a_mode = A_ROWMAJOR; g.a.src0 = a; g.a.ld = C;
        g.b_mode = B_ROWMAJOR; g.b.src0 = x; g.b.ld = C;
        g.out_mode = OUT_CHMAJOR; g.out = out; g.ldo = M;
        g.tag = TAG_LOSS_GEMM;
        return launch_gemm_tc(g, st);
    };
    // out[rows of B-operand][C] (+)= dS-block * features, with the feature matrix as the packed "A" operand
    auto dgemm = [&](const uint8_t* img, int Kd, int b_mode, const float* ds, int Nd, float* out, int accumulate) {
        GemmParams g;
        memset(&g, 0, sizeof(g));
        g.Md = C; g.Nd = Nd; g.Kd = Kd; g.nsplit = nsplit; g.ksplit = 1;
        g.a_mode = A_PACKED; g.a_packed = img; g.a_packed_kblocks = (Kd + 63) / 64;
        g.b_mode = b_mode; g.b.src0 = ds; g.b.ld = M;
        g.out_mode = accumulate ? OUT_ROWMAJOR_ACC : OUT_ROWMAJOR; g.out = out; g.ldo = C;
        g.tag = TAG_LOSS_GEMM;
        return launch_gemm_tc(g, st);
    };

    if (want_circle) {
        RUN(sim(x, M, w.S));
        invert_order_kernel<<<1, 256, 0, st>>>(order, G, w.inv);
        loss_rowstats_kernel<<<M, 128, 0, st>>>(w.S, M, M, B, 0, M, w.rmax, w.rsum);
    }
    if (want_global) {
        RUN(sim(xg, B, w.S + (size_t)M * M));
        loss_rowstats_kernel<<<B, 128, 0, st>>>(w.S, M, M, B, M, B, w.rmax, w.rsum);
    }
    loss_finalize_kernel<<<1, 256, 0, st>>>(w.S, M, M, B, G, order, w.rmax, w.rsum, want_global, want_circle, loss, w.lcG, w.pgG,
                                            w.lcC, w.pgC);
    FACL_CHECK_LAUNCH();
    // feature matrices as packed A operands: A[m = c][k = row] = feat[row][c]
    RUN(pack_weight_launch(x, 1, C, C, M, w.img_x, st));
    // When both gradients go to the same buffer (dx_circle_part == dx_global_part) the second one accumulates.
    const bool one_buffer = want_global && want_circle && dx_circle_part == dx_global_part;
    if (want_global) {
        float* Sg = w.S + (size_t)M * M;
        loss_ds_kernel<<<div_up((long long)B * M, 256), 256, 0, st>>>(w.S, M, M, B, G, order, w.inv, M, B, w.lcG, w.pgG, w.lcC, w.pgC);
        FACL_CHECK_LAUNCH();
        RUN(pack_weight_launch(xg, 1, C, C, B, w.img_xall, st));
        // dxg[n] = sum_j dS_g[n][j] x[j] ;  dx[j] = sum_n dS_g[n][j] xg[n]
        RUN(dgemm(w.img_x, M, B_ROWMAJOR, Sg, B, dxg, 0));
        RUN(dgemm(w.img_xall, B, B_CHMAJOR, Sg, M, dx_global_part, 0));
    }
    if (want_circle) {
        loss_ds_kernel<<<div_up((long long)M * M, 256), 256, 0, st>>>(w.S, M, M, B, G, order, w.inv, 0, M, w.lcG, w.pgG, w.lcC, w.pgC);
        FACL_CHECK_LAUNCH();
        // dx[a] = sum_j dS[a][j] x[j]   and   dx[j] += sum_a dS[a][j] x[a]
        RUN(dgemm(w.img_x, M, B_ROWMAJOR, w.S, M, dx_circle_part, one_buffer ? 1 : 0));
        RUN(dgemm(w.img_x, M, B_CHMAJOR, w.S, M, dx_circle_part, 1));
    }
    return 0;
}

}  // namespace facl

extern "C" {

size_t facl_contrast_workspace_bytes(int G, int B, int C) { return facl::loss_ws_layout(G, B, C, nullptr, nullptr); }

int facl_contrast_losses(const float* x, const float* x_global, int G, int B, int C, const int* order, int want_global,
                         int want_circle, int nsplit, void* workspace, float* loss, float* dx_global_part, float* dx_global,
                         float* dx_circle_part, void* stream) {
    return facl::contrast_losses(x, x_global, G, B, C, order, want_global, want_circle, nsplit, workspace, loss, dx_global_part,
                                 dx_global, dx_circle_part, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
