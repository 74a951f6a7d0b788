// Internal launcher prototypes shared between the .cu files and the C-ABI shim (facl_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace facl {

// fps.cu
int fps_launch(const float* pts, int V, int N, int D, const int* start, int m, int* out, cudaStream_t st);
int fps_reorder_launch(const float* pts, int V, int N, int D, const int* picks, int m, float* out, cudaStream_t st);
// group.cu
int group_launch(const float* points, int M, int N, int D, int S, int K, float r2, float* xt, int* idx_out, cudaStream_t st);
// pack.cu
size_t packed_weight_bytes(int Md, int Kd);
int pack_weight_launch(const float* src, long long stride_m, long long stride_k, int Md, int Kd, void* image, cudaStream_t st);

}  // namespace facl
