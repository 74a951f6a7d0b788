// Small HBM-bound helper kernels around the tcgen05 GEMMs: BatchNorm statistic finalisation (forward and
// backward), pooling across the views of a sequence, max-pool gradient routing, transposes, Adam.
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {

// ---- BatchNorm forward: partial sums -> (mean, rstd, scale, shift), running-stat update -------------------
// reference: nn.BatchNorm2d / nn.BatchNorm1d instantiated at cn3d_model_conbag.py:165,169,173,183,187,191,203
// (eps 1e-5, momentum 0.1, biased variance for normalisation, unbiased for running_var).
// one warp per channel: lanes stride over the P partial sums, then a shuffle reduction (in double)
__device__ __forceinline__ void warp_sum_partials(const float* __restrict__ partials, int P, int C, int c, double& s, double& q) {
    const int lane = threadIdx.x & 31;
    s = 0.0;
    q = 0.0;
    for (int p = lane; p < P; p += 32) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(partials) + (long long)p * C + c);
        s += (double)v.x;
        q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ partials, int P, int C, double n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var, float eps,
                                   float momentum, int training, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
    pdl_prologue();
    int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double mu, var;
    double s = 0.0, q = 0.0;
    if (training) warp_sum_partials(partials, P, C, c, s, q);
    if ((threadIdx.x & 31) != 0) return;
    if (training) {
        mu = s / n;
        var = q / n - mu * mu;
        if (var < 0.0) var = 0.0;
        double unbiased = (n > 1.0) ? var * n / (n - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    } else {
        mu = running_mean[c];
        var = running_var[c];
    }
    double r = 1.0 / sqrt(var + (double)eps);
    double g = gamma[c];
    mean[c] = (float)mu;
    rstd[c] = (float)r;
    scale[c] = (float)(g * r);
    shift[c] = (float)((double)beta[c] - mu * g * r);
}

// ---- BatchNorm backward: (sum dy, sum dy*z) -> dgamma, dbeta and the coefficients of
//      dz = c0*dy + c1*z + c2   (dy already ReLU-masked) ------------------------------------------------------
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int P, int C, double n, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, float* dgamma, float* dbeta,
                                       int accumulate, float* __restrict__ c0, float* __restrict__ c1, float* __restrict__ c2) {
    pdl_prologue();
    int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s, q;
    warp_sum_partials(partials, P, C, c, s, q);
    if ((threadIdx.x & 31) != 0) return;
    double mu = mean[c], r = rstd[c], g = gamma[c];
    double dbe = s;
    double dga = r * (q - mu * s);
    double k0 = g * r;
    double k1 = -k0 * (dga / n) * r;
    double k2 = -k0 * (dbe / n) - k1 * mu;
    c0[c] = (float)k0;
    c1[c] = (float)k1;
    c2[c] = (float)k2;
    if (accumulate) {
        dgamma[c] += (float)dga;
        dbeta[c] += (float)dbe;
    } else {
        dgamma[c] = (float)dga;
        dbeta[c] = (float)dbe;
    }
}

// per-channel (sum v, sum v*z) over a channel-major matrix -> one "partial" (P = 1); block per channel
__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ v, const float* __restrict__ z, long long ld, int n,
                                                       int pairs, float* __restrict__ out) {
    pdl_prologue();
    int c = blockIdx.x;
    const float* vr = v + (long long)c * ld;
    const float* zr = z ? z + (long long)c * ld : nullptr;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        float a = vr[i];
        s += a;
        q += (double)a * (zr ? zr[i] : a);
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (pairs) {
            out[c * 2 + 0] = (float)sh[0][0];
            out[c * 2 + 1] = (float)sh[1][0];
        } else {
            out[c] = (float)sh[0][0];
        }
    }
}

// out[c][r] = in[r][c]   (rows R, cols C), 32x32 tiles through shared memory; blockIdx.x walks the rows (may be millions)
__global__ void transpose_kernel(const float* __restrict__ in, long long ldi, float* __restrict__ out, long long ldo, int R, int C) {
    pdl_prologue();
    __shared__ float tile[32][33];
    int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? in[(long long)r * ldi + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[(long long)c * ldo + r] = tile[threadIdx.x][i];
    }
}

// sequence aggregation (reference cn3d_model_conbag.py:225-226): clouds are G-major, cloud g*B+b is view g of
// sequence b; the max over all G*S positions = max over g of the per-cloud pooled value (min when the BN scale
// that follows is negative).  pooled [C][ldp] (first M = G*B columns) -> seq [C][lds] (B columns), argg [C][B].
__global__ void seq_pool_kernel(const float* __restrict__ pooled, long long ldp, const float* __restrict__ sign, int C, int G, int B,
                                float* __restrict__ seq, long long lds, unsigned char* __restrict__ argg) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)C * B) return;
    int c = (int)(t / B), b = (int)(t % B);
    bool keep_max = sign[c] >= 0.f;
    const float* row = pooled + (long long)c * ldp;
    float best = 0.f;
    int bg = 0;
    for (int g = 0; g < G; ++g) {
        float v = row[g * B + b];
        float sv = keep_max ? v : -v;
        if (g == 0 || sv > best) {
            best = sv;
            bg = g;
        }
    }
    seq[(long long)c * lds + b] = keep_max ? best : -best;
    argg[(long long)c * B + b] = (unsigned char)bg;
}

// gradient arriving at the per-cloud pooled feature = grad from the cloud's own embedding + (if this cloud won
// the sequence max) the grad from the sequence embedding
__global__ void combine_pool_grads_kernel(float* __restrict__ dcloud, long long ldc, const float* __restrict__ dseq, long long lds,
                                          const unsigned char* __restrict__ argg, int C, int G, int B) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int M = G * B;
    if (t >= (long long)C * M) return;
    int c = (int)(t / M), m = (int)(t % M);
    int g = m / B, b = m % B;
    if (argg[(long long)c * B + b] == g) dcloud[(long long)c * ldc + m] += dseq[(long long)c * lds + b];
}

// max-pool backward: dense[c][grp*pool + arg[c][grp]] = v[c][grp]; dense must be zero-filled beforehand
__global__ void pool_scatter_kernel(const float* __restrict__ v, long long ldv, const unsigned char* __restrict__ arg, long long lda,
                                    int C, int groups, int pool, float* __restrict__ dense, long long ldd) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)C * groups) return;
    int c = (int)(t / groups), g = (int)(t % groups);
    dense[(long long)c * ldd + (long long)g * pool + arg[(long long)c * lda + g]] = v[(long long)c * ldv + g];
}

// centres [R][3] -> channel-major [3][R]
__global__ void centres_to_chmajor_kernel(const float* __restrict__ c, int R, float* __restrict__ out, long long ldo) {
    pdl_prologue();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    out[r] = c[(long long)r * 3 + 0];
    out[ldo + r] = c[(long long)r * 3 + 1];
    out[2 * ldo + r] = c[(long long)r * 3 + 2];
}

// x_nor = x / max(||x||_2, 1e-12)  (F.normalize, cn3d_model_conbag.py:231); one warp per row
__global__ void l2_normalize_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ out) {
    pdl_prologue();
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * C;
    float s = 0.f;
    for (int i = lane; i < C; i += 32) s = fmaf(xr[i], xr[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
    for (int i = lane; i < C; i += 32) out[(long long)row * C + i] = xr[i] * inv;
}

// the constant vectors of the encoder's `vec` buffer in one launch: identity transform for the 3 centre-xyz channels of the
// net3DV_3 input (scale 1 at 0, shift 0 at 320, lower bound -inf at 640), zero lower bounds (ReLU) for 256 channels at 643 and
// 1024 zeros at 960
__global__ void encoder_const_vectors_kernel(float* __restrict__ vec) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3) {
        vec[i] = 1.f;
        vec[320 + i] = 0.f;
        vec[640 + i] = -INFINITY;
    }
    if (i < 256) vec[643 + i] = 0.f;
    if (i < 1024) vec[960 + i] = 0.f;
}

__global__ void fill_kernel(float* p, long long n, float v) {
    pdl_prologue();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Adam (torch.optim.Adam semantics, no weight decay / amsgrad; reference cn3d_train_motion_GL.py:180):
// one launch over a table of tensors.
struct AdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
};
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_c, float b1, float b2, float eps, float bc2_sqrt) {
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= lr_c * (m / denom);
}
__global__ void adam_kernel(const AdamTensor* __restrict__ tab, int ntensors, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt) {
    pdl_prologue();
    const float lr_c = lr / bc1;
    for (int t = blockIdx.y; t < ntensors; t += gridDim.y) {
        AdamTensor a = tab[t];
        const bool vec = ((a.n & 3) == 0) && (((reinterpret_cast<uintptr_t>(a.p) | reinterpret_cast<uintptr_t>(a.g) |
                                                reinterpret_cast<uintptr_t>(a.m) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0);
        if (vec) {
            const long long n4 = a.n >> 2;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
                float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
                const float4 g = reinterpret_cast<const float4*>(a.g)[i];
                adam_one(p.x, g.x, m.x, v.x, lr_c, b1, b2, eps, bc2_sqrt);
                adam_one(p.y, g.y, m.y, v.y, lr_c, b1, b2, eps, bc2_sqrt);
                adam_one(p.z, g.z, m.z, v.z, lr_c, b1, b2, eps, bc2_sqrt);
                adam_one(p.w, g.w, m.w, v.w, lr_c, b1, b2, eps, bc2_sqrt);
                reinterpret_cast<float4*>(a.p)[i] = p;
                reinterpret_cast<float4*>(a.m)[i] = m;
                reinterpret_cast<float4*>(a.v)[i] = v;
            }
        } else {
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
                float p = a.p[i], m = a.m[i], v = a.v[i];
                adam_one(p, a.g[i], m, v, lr_c, b1, b2, eps, bc2_sqrt);
                a.p[i] = p; a.m[i] = m; a.v[i] = v;
            }
        }
    }
}

}  // namespace

int bn_finalize_launch(const float* partials, int P, int C, double n, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float eps, float momentum, int training, float* mean, float* rstd, float* scale,
                       float* shift, cudaStream_t st) {
    ScopedTimer timer(TAG_BN, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(bn_finalize_kernel, dim3(div_up(C, 4)), dim3(128), 0, st, partials, P, C, n, gamma, beta, running_mean, running_var, eps, momentum, training,
                                                       mean, rstd, scale, shift));
    return (int)cudaGetLastError();
}

int bn_bwd_finalize_launch(const float* partials, int P, int C, double n, const float* gamma, const float* mean, const float* rstd,
                           float* dgamma, float* dbeta, int accumulate, float* c0, float* c1, float* c2, cudaStream_t st) {
    ScopedTimer timer(TAG_BN, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(bn_bwd_finalize_kernel, dim3(div_up(C, 4)), dim3(128), 0, st, partials, P, C, n, gamma, mean, rstd, dgamma, dbeta, accumulate, c0, c1, c2));
    return (int)cudaGetLastError();
}

int rowstats_launch(const float* v, const float* z, long long ld, int C, int n, int pairs, float* out, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(rowstats_kernel, dim3(C), dim3(256), 0, st, v, z, ld, n, pairs, out));
    return (int)cudaGetLastError();
}

int transpose_launch(const float* in, long long ldi, float* out, long long ldo, int R, int C, cudaStream_t st) {
    ScopedTimer timer(TAG_TRANSPOSE, st);
    count_launch();
    dim3 grid(div_up(R, 32), div_up(C, 32)), block(32, 8);
    FACL_LAUNCH_OK(launch_pdl(transpose_kernel, dim3(grid), dim3(block), 0, st, in, ldi, out, ldo, R, C));
    return (int)cudaGetLastError();
}

int seq_pool_launch(const float* pooled, long long ldp, const float* sign, int C, int G, int B, float* seq, long long lds,
                    unsigned char* argg, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(seq_pool_kernel, dim3(div_up((long long)C * B, 256)), dim3(256), 0, st, pooled, ldp, sign, C, G, B, seq, lds, argg));
    return (int)cudaGetLastError();
}

int combine_pool_grads_launch(float* dcloud, long long ldc, const float* dseq, long long lds, const unsigned char* argg, int C, int G,
                              int B, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(combine_pool_grads_kernel, dim3(div_up((long long)C * G * B, 256)), dim3(256), 0, st, dcloud, ldc, dseq, lds, argg, C, G, B));
    return (int)cudaGetLastError();
}

int pool_scatter_launch(const float* v, long long ldv, const unsigned char* arg, long long lda, int C, int groups, int pool,
                        float* dense, long long ldd, cudaStream_t st) {
    ScopedTimer timer(TAG_SCATTER, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(pool_scatter_kernel, dim3(div_up((long long)C * groups, 256)), dim3(256), 0, st, v, ldv, arg, lda, C, groups, pool, dense, ldd));
    return (int)cudaGetLastError();
}

int centres_to_chmajor_launch(const float* c, int R, float* out, long long ldo, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(centres_to_chmajor_kernel, dim3(div_up(R, 256)), dim3(256), 0, st, c, R, out, ldo));
    return (int)cudaGetLastError();
}

int l2_normalize_launch(const float* x, int rows, int C, float* out, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(l2_normalize_kernel, dim3(div_up((long long)rows * 32, 256)), dim3(256), 0, st, x, rows, C, out));
    return (int)cudaGetLastError();
}

int encoder_const_vectors_launch(float* vec, cudaStream_t st) {
    ScopedTimer timer(TAG_MEMSET, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(encoder_const_vectors_kernel, dim3(4), dim3(256), 0, st, vec));
    return (int)cudaGetLastError();
}

int fill_launch(float* p, long long n, float v, cudaStream_t st) {
    if (n <= 0) return 0;
    ScopedTimer timer(TAG_MEMSET, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(fill_kernel, dim3(div_up(n, 256)), dim3(256), 0, st, p, n, v));
    return (int)cudaGetLastError();
}

int adam_launch(const void* table_dev, int ntensors, float lr, float b1, float b2, float eps, int step, cudaStream_t st) {
    ScopedTimer timer(TAG_ADAM, st);
    count_launch();
    float bc1 = (float)(1.0 - pow((double)b1, (double)step));
    float bc2 = (float)sqrt(1.0 - pow((double)b2, (double)step));
    dim3 grid(128, ntensors < 64 ? ntensors : 64);
    FACL_LAUNCH_OK(launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const AdamTensor*>(table_dev), ntensors, lr, b1, b2, eps, bc1, bc2));
    return (int)cudaGetLastError();
}

}  // namespace facl
