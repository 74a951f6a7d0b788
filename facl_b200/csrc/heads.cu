// The loss heads the reference scripts carry but disable by constants: SwAV-style cluster
// assignment with Sinkhorn-Knopp normalisation and the CLD k-means grouping loss.
//
//   distributed_sinkhorn / shoot_infs   reference training_code/cn3d_model_conbag.py:391-425
//   the inline SwAV block               reference training_code/cn3d_train_motion_GL.py:236-262  (swa_if = 0)
//   KMeans / grouping / CLD_Loss        reference training_code/utils_my.py:152-198 (= cn3d_train_motion_GL.py:36-70, cld_if = 0)
//
// The matrices are tiny (64 prototypes x B samples; 3B points x 60 clusters x 512), so every routine is ONE thread block that
// keeps its working set in shared memory and runs all iterations in a single launch; the matrix products of the heads (mapping,
// affinity) go through the tensor-core GEMM of the library (facl_gemm_tc), see facl_b200/heads.py.
#include <math.h>

#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {

__device__ __forceinline__ float block_sum(float v, float* red, int nwarps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red, int nwarps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int w = 0; w < nwarps; ++w) t = fmaxf(t, red[w]);
    return t;
}

// Sinkhorn-Knopp on Q [K][B] (K prototypes, B samples), cn3d_model_conbag.py:391-406:
//   Q = shoot_infs(Q); Q /= sum(Q); repeat iters: { u = sum_b Q; u = shoot_infs((1/K) / u); Q *= u[:,None]; Q *= ((1/B) / sum_k Q)[None,:] }
//   return (Q / sum_k Q).t()   -> out [B][K]
// shoot_infs (:409-425): infinities are replaced by the maximum of the remaining entries.
constexpr int SK_THREADS = 256;
__global__ void __launch_bounds__(SK_THREADS) sinkhorn_kernel(const float* __restrict__ Qin, int K, int B, int iters, float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ float sk_smem[];
    float* Q = sk_smem;              // [K][B]
    float* u = Q + (size_t)K * B;    // [K]
    float* cs = u + K;               // [B]
    __shared__ float red[SK_THREADS / 32];
    const int n = K * B, tid = threadIdx.x;
    float mx = -INFINITY;
    int inf_seen = 0;
    for (int i = tid; i < n; i += SK_THREADS) {
        const float v = Qin[i];
        Q[i] = v;
        if (isinf(v)) inf_seen = 1; else mx = fmaxf(mx, v);
    }
    mx = fmaxf(block_max(mx, red, SK_THREADS / 32), 0.f);          // the reference takes the max AFTER writing 0 into the inf slots
    if (__syncthreads_or(inf_seen))
        for (int i = tid; i < n; i += SK_THREADS)
            if (isinf(Q[i])) Q[i] = mx;
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < n; i += SK_THREADS) s += Q[i];
    s = block_sum(s, red, SK_THREADS / 32);
    for (int i = tid; i < n; i += SK_THREADS) Q[i] /= s;
    __syncthreads();
    const float r = 1.f / (float)K, c = 1.f / (float)B;
    const int warp = tid >> 5, lane = tid & 31;
    for (int it = 0; it < iters; ++it) {
        // u[k] = r / sum_b Q[k][b], infinities shot to the maximum of the rest
        for (int k = warp; k < K; k += SK_THREADS / 32) {
            float t = 0.f;
            for (int b = lane; b < B; b += 32) t += Q[(size_t)k * B + b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            if (lane == 0) u[k] = r / t;
        }
        __syncthreads();
        float um = -INFINITY;
        int ui = 0;
        for (int k = tid; k < K; k += SK_THREADS) {
            if (isinf(u[k])) ui = 1; else um = fmaxf(um, u[k]);
        }
        um = fmaxf(block_max(um, red, SK_THREADS / 32), 0.f);
        if (__syncthreads_or(ui))
            for (int k = tid; k < K; k += SK_THREADS)
                if (isinf(u[k])) u[k] = um;
        __syncthreads();
        for (int i = tid; i < n; i += SK_THREADS) Q[i] *= u[i / B];
        __syncthreads();
        for (int b = tid; b < B; b += SK_THREADS) {
            float t = 0.f;
            for (int k = 0; k < K; ++k) t += Q[(size_t)k * B + b];
            cs[b] = c / t;
        }
        __syncthreads();
        for (int i = tid; i < n; i += SK_THREADS) Q[i] *= cs[i % B];
        __syncthreads();
    }
    for (int b = tid; b < B; b += SK_THREADS) {
        float t = 0.f;
        for (int k = 0; k < K; ++k) t += Q[(size_t)k * B + b];
        cs[b] = t;
    }
    __syncthreads();
    for (int i = tid; i < n; i += SK_THREADS) {
        const int k = i / B, b = i % B;
        out[(size_t)b * K + k] = Q[i] / cs[b];
    }
}

// loss = -(1/rows) sum_r sum_k q[r][k] log softmax(scale * logits[r])[k]   (cn3d_train_motion_GL.py:258-260, scale = 1 / 0.1)
// dlogits[r][k] = scale * (softmax[k] * sum_k' q[r][k'] - q[r][k]) / rows.   Warp per row; loss accumulated atomically (memset first).
__global__ void __launch_bounds__(256) soft_xent_kernel(const float* __restrict__ logits, const float* __restrict__ q, int rows, int K,
                                                        float scale, float* __restrict__ loss, float* __restrict__ dlogits) {
    pdl_prologue();
    const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* z = logits + (size_t)row * K;
    const float* qr = q + (size_t)row * K;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float se = 0.f, sq = 0.f, sqz = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float t = z[k] * scale;
        se += expf(t - mx);
        sq += qr[k];
        sqz += qr[k] * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xFFFFFFFFu, se, o);
        sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
        sqz += __shfl_xor_sync(0xFFFFFFFFu, sqz, o);
    }
    const float lse = mx + logf(se), inv = 1.f / (float)rows;
    if (lane == 0 && loss) atomicAdd(loss, (sq * lse - sqz) * inv);
    if (dlogits)
        for (int k = lane; k < K; k += 32) dlogits[(size_t)row * K + k] = scale * (expf(z[k] * scale - lse) * sq - qr[k]) * inv;
}

// k-means exactly as utils_my.py:180-198: centroids start as the first K points; every iteration assigns each point to its nearest
// centroid (squared distance, FIRST minimum) and replaces every centroid by the mean of its members (an empty cluster divides its
// zero sum by 1).  Returns the labels of the LAST assignment and the centroids after the last update.
// One block; centroids [K][D] live in shared memory.  labels: [N] int32, counts: [K] int32 (>= 1, the divisor used).
constexpr int KM_THREADS = 1024;
__global__ void __launch_bounds__(KM_THREADS) kmeans_kernel(const float* __restrict__ x, int N, int D, int K, int iters,
                                                            int* __restrict__ labels, float* __restrict__ centroids,
                                                            int* __restrict__ counts) {
    pdl_prologue();
    extern __shared__ float km_smem[];
    float* c = km_smem;                                  // [K][D]
    int* lab = reinterpret_cast<int*>(c + (size_t)K * D);   // [N]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < K * D; i += KM_THREADS) c[i] = x[i];       // c = x[:K]
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        for (int n = warp; n < N; n += KM_THREADS / 32) {
            const float* xn = x + (size_t)n * D;
            float best = INFINITY;
            int bk = 0;
            for (int k = 0; k < K; ++k) {
                const float* ck = c + (size_t)k * D;
                float s = 0.f;
                for (int d = lane; d < D; d += 32) {
                    const float df = xn[d] - ck[d];
                    s = fmaf(df, df, s);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                if (s < best) { best = s; bk = k; }                  // strict '<' while k ascends: first minimum (torch.argmin)
            }
            if (lane == 0) lab[n] = bk;
        }
        __syncthreads();
        for (int k = warp; k < K; k += KM_THREADS / 32) {
            int cnt = 0;
            for (int n = 0; n < N; ++n) cnt += (lab[n] == k);
            const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
            for (int d = lane; d < D; d += 32) {
                float s = 0.f;
                for (int n = 0; n < N; ++n)
                    if (lab[n] == k) s += x[(size_t)n * D + d];          // members in ascending row order, like scatter_add_
                c[(size_t)k * D + d] = s * inv;
            }
            if (lane == 0 && counts) counts[k] = cnt > 0 ? cnt : 1;
        }
        __syncthreads();
    }
    for (int i = tid; i < K * D; i += KM_THREADS) centroids[i] = c[i];
    for (int n = tid; n < N; n += KM_THREADS) labels[n] = lab[n];
}

}  // namespace

int sinkhorn_launch(const float* q, int K, int B, int iters, float* out, cudaStream_t st) {
    if (!q || !out || K <= 0 || B <= 0 || iters < 0) return (int)cudaErrorInvalidValue;
    const size_t smem = ((size_t)K * B + K + B) * sizeof(float);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    if (smem > 48 * 1024) FACL_CHECK(cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScopedTimer timer(TAG_LOSS_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(sinkhorn_kernel, dim3(1), dim3(SK_THREADS), smem, st, q, K, B, iters, out));
    return (int)cudaGetLastError();
}

int soft_xent_launch(const float* logits, const float* q, int rows, int K, float scale, float* loss, float* dlogits, cudaStream_t st) {
    if (!logits || !q || rows <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
    ScopedTimer timer(TAG_LOSS_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(soft_xent_kernel, dim3(div_up((long long)rows * 32, 256)), dim3(256), 0, st, logits, q, rows, K, scale, loss, dlogits));
    return (int)cudaGetLastError();
}

int kmeans_launch(const float* x, int N, int D, int K, int iters, int* labels, float* centroids, int* counts, cudaStream_t st) {
    if (!x || !labels || !centroids || N <= 0 || D <= 0 || K <= 0 || K > N || iters < 1) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)K * D * sizeof(float) + (size_t)N * sizeof(int);
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    if (smem > 48 * 1024) FACL_CHECK(cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScopedTimer timer(TAG_LOSS_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(kmeans_kernel, dim3(1), dim3(KM_THREADS), smem, st, x, N, D, K, iters, labels, centroids, counts));
    return (int)cudaGetLastError();
}

}  // namespace facl

extern "C" {

int facl_sinkhorn(const float* q, int K, int B, int iters, float* out, void* stream) {
    return facl::sinkhorn_launch(q, K, B, iters, out, reinterpret_cast<cudaStream_t>(stream));
}
int facl_soft_xent(const float* logits, const float* q, int rows, int K, float scale, float* loss, float* dlogits, void* stream) {
    return facl::soft_xent_launch(logits, q, rows, K, scale, loss, dlogits, reinterpret_cast<cudaStream_t>(stream));
}
int facl_kmeans(const float* x, int N, int D, int K, int iters, int* labels, float* centroids, int* counts, void* stream) {
    return facl::kmeans_launch(x, N, D, K, iters, labels, centroids, counts, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
