// Linear probe on the extracted features: the pieces of linear_classify/fc_model.py:12-25
// (`Final_FC`: F.normalize -> nn.Linear(22*512, 120)) and of the loop in linear_classify/linercls.py:106-124
// (CrossEntropyLoss with mean reduction, top-1 accuracy) that are not a GEMM.  The two GEMMs (logits, weight gradient)
// run on the tcgen05 kernel of gemm_tc.cu; the row normalisation is l2_normalize_launch (elementwise.cu).
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {

// one warp per sample: log-sum-exp, loss, d(loss)/d(logits) written TRANSPOSED ([C][rows], the A operand of the
// weight-gradient GEMM), bias gradient and top-1 hit count
__global__ void softmax_xent_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int rows, int C,
                                    float* __restrict__ loss, float* __restrict__ dlogits_t, float* __restrict__ dbias,
                                    int* __restrict__ hits) {
    pdl_prologue();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* z = logits + (long long)row * C;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        float v = z[c];
        if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        int oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xFFFFFFFFu, se, o);
    int y = labels[row];
    const bool bad_label = y < 0 || y >= C;              // torch's CrossEntropyLoss device-asserts here; a kernel cannot raise, so the
    y = bad_label ? 0 : y;                               // loss turns NaN (loud) instead of reading outside the logits row
    const float inv_rows = 1.f / (float)rows, lse = mx + __logf(se);
    if (lane == 0) {
        if (loss) atomicAdd(loss, bad_label ? __int_as_float(0x7fc00000) : (lse - z[y]) * inv_rows);
        if (hits && arg == y) atomicAdd(hits, 1);
    }
    if (dlogits_t || dbias) {
        for (int c = lane; c < C; c += 32) {
            float g = (__expf(z[c] - lse) - (c == y ? 1.f : 0.f)) * inv_rows;
            if (dlogits_t) dlogits_t[(long long)c * rows + row] = g;
            if (dbias) atomicAdd(dbias + c, g);
        }
    }
}

}  // namespace

int softmax_xent_launch(const float* logits, const int* labels, int rows, int C, float* loss, float* dlogits_t, float* dbias, int* hits,
                        cudaStream_t st) {
    if (!logits || !labels || rows <= 0 || C <= 0) return (int)cudaErrorInvalidValue;
    ScopedTimer timer(TAG_LOSS_MISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(softmax_xent_kernel, dim3(div_up((long long)rows * 32, 256)), dim3(256), 0, st, logits, labels, rows, C, loss, dlogits_t, dbias, hits));
    return (int)cudaGetLastError();
}

}  // namespace facl
