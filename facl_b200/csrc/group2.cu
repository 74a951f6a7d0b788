// Level-2 set-abstraction grouping over channel-first features (reference training_code/utils_my.py:332-381,
// group_points_2 / group_points_2_3DV): points (M, C, N1) with channels 0..2 = xyz, centres = the first S2 points;
// K nearest of the N1 points per centre, out-of-ball slots redirected to the centre, then a gather of ALL C channels
// into (M, C, S2, K) with the centre subtracted from xyz.
//
// The neighbour selection is the level-1 kernel (group.cu, indices only) run on an xyz row image of the cloud; the
// gather below is HBM-write bound: 4*C*S2*K bytes out per cloud against 4*C*N1 in.  One CTA stages CH channel rows
// (N1 floats each) in shared memory and streams the (S2*K) neighbour indices through registers, so every index is
// read once per CH channels and every store is a coalesced 128-byte line.
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {

__global__ void xyz_rows_kernel(const float* __restrict__ feats, int C, int N1, float4* __restrict__ rows, long long total) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    long long m = t / N1;
    int n = (int)(t - m * N1);
    const float* f = feats + m * C * N1 + n;
    rows[t] = make_float4(__ldg(f), __ldg(f + N1), __ldg(f + 2ll * N1), 0.f);
}

constexpr int CH = 8;          // channel rows per CTA
constexpr int GT = 256;

__global__ void __launch_bounds__(GT) gather_channels_kernel(const float* __restrict__ feats, const int* __restrict__ idx, int C, int N1,
                                                             int S2, int K, float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ float rows[];                 // [CH][N1]
    const int m = blockIdx.y, c0 = blockIdx.x * CH;
    const int nch = min(CH, C - c0);
    const float* __restrict__ src = feats + ((long long)m * C + c0) * N1;
    for (int i = threadIdx.x; i < nch * N1; i += GT) rows[i] = __ldg(src + i);
    __syncthreads();
    const long long J = (long long)S2 * K;
    const int* __restrict__ im = idx + (long long)m * J;
    float* __restrict__ om = out + ((long long)m * C + c0) * J;
    for (long long j = threadIdx.x; j < J; j += GT) {
        const int n = __ldg(im + j);
        const int s = (int)(j / K);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (c >= nch) break;
            float v = rows[c * N1 + n];
            if (c0 + c < 3) v = __fsub_rn(v, rows[c * N1 + s]);       // utils_my.py:354: xyz relative to the centre
            __stcs(om + c * J + j, v);
        }
    }
}

}  // namespace

size_t group_level2_scratch_bytes(int M, int N1, int S2, int K) {
    return (size_t)M * N1 * 16 + (size_t)M * S2 * K * 4;
}

int group_level2_launch(const float* feats, int M, int C, int N1, int S2, int K, float r2, float* out, int* idx, void* scratch,
                        cudaStream_t st) {
    if (!feats || !out || !scratch || M <= 0 || C < 3 || N1 <= 0 || S2 <= 0 || S2 > N1 || K <= 0 || K > N1) return (int)cudaErrorInvalidValue;
    if ((size_t)CH * N1 * 4 > 200 * 1024) return (int)cudaErrorInvalidValue;
    float4* xyz = reinterpret_cast<float4*>(scratch);
    int* nbr = idx ? idx : reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + (size_t)M * N1 * 16);
    long long total = (long long)M * N1;
    {
        ScopedTimer timer(TAG_GROUP2, st);
        count_launch();
        FACL_LAUNCH_OK(launch_pdl(xyz_rows_kernel, dim3(div_up(total, 256)), dim3(256), 0, st, feats, C, N1, xyz, total));
        FACL_CHECK(cudaGetLastError());
    }
    int e = group_launch(reinterpret_cast<const float*>(xyz), M, N1, 4, S2, K, r2, nullptr, nbr, st);
    if (e) return e;
    const size_t smem = (size_t)CH * N1 * 4;
    if (smem > 48 * 1024) FACL_CHECK(cudaFuncSetAttribute(gather_channels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScopedTimer timer(TAG_GROUP2, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(gather_channels_kernel, dim3(dim3((C + CH - 1) / CH, M)), dim3(GT), smem, st, feats, nbr, C, N1, S2, K, out));
    return (int)cudaGetLastError();
}

}  // namespace facl
