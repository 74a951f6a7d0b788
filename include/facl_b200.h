/* facl_b200 -- C ABI of the B200 (sm_100a) hot-path library, libfacl_b200.so.
 *
 * The reference (tangent-T/FACL) has no FFI layer: its hot path sits behind Python callables
 * (training_code/utils_my.py, training_code/cn3d_model_conbag.py, the loop body of
 * training_code/cn3d_train_motion_GL.py).  Each entry point below names the reference call it replaces;
 * facl_b200/*.py binds them with ctypes and re-exports the reference's Python names.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a cudaError_t value (facl_error_string() decodes it);
 *   - pointers are DEVICE pointers unless the name ends in _host; all buffers are owned by the caller
 *     (PyTorch's caching allocator on the Python side) -- the library never allocates or frees device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream and never synchronise;
 *   - not thread-safe per stream: call from one host thread per stream (the Python side holds the GIL).
 */
#ifndef FACL_B200_H
#define FACL_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define FACL_API __attribute__((visibility("default")))
#else
#define FACL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

FACL_API const char* facl_version(void);
FACL_API const char* facl_error_string(int code);

/* ---- K1: farthest-point sampling -------------------------------------------------------------------------
 * replaces NTU_RGBD_new.farthest_point_sampling_fast, reference training_code/cn3D_data_set.py:675-694
 * (copies: cn3d_data_load.py:301-320, generate_data/generate_NTU.py:299-318).
 * points (V,N,D) fp32, xyz = first three channels; start_idx (V) int32 = first pick of each cloud (the
 * reference draws it from numpy's global RNG, :679); out_idx (V,m) int32 in pick order.  N <= 32768. */
FACL_API int facl_fps(const float* points, int V, int N, int D, const int* start_idx, int m, int* out_idx, void* stream);

/* replaces NTU_RGBD_new.fps_sample_data, cn3D_data_set.py:665-672: rows [0,m) of `out` are the picks, the
 * remaining rows follow in ascending original index.  out (V,N,D) must not alias points. */
FACL_API int facl_fps_reorder(const float* points, int V, int N, int D, const int* picks, int m, float* out, void* stream);

/* ---- K2: kNN + ball query + relative-xyz gather ----------------------------------------------------------
 * replaces group_points_3DV (reference training_code/utils_my.py:255-291) and its copies
 * group_points_3DV_2048 (:7-42), group_points (:217-253), group_points_3DV_nums (:293-328).
 * points (M,N,D) fp32; centres are rows [0,S) of each cloud; r2 = squared ball radius.
 * xt  (M,S,K,D) fp32: the K nearest rows (ascending (distance, index)), xyz minus the centre, slots with
 *     distance > r2 replaced by the centre itself.  The reference's `inputs_level1` is the (M,D,S,K)
 *     permuted view of this buffer.
 * idx (M,S,K) int32 or NULL: the chosen point index per slot, after the ball redirect.  K <= 128. */
FACL_API int facl_group_points(const float* points, int M, int N, int D, int S, int K, float r2, float* xt, int* idx, void* stream);

/* ---- tcgen05 GEMM building block (operand packing + fused GEMM), exposed for tests and for the Python-side
 * orchestration of the encoder layers (replaces the cuDNN/cuBLAS calls behind nn.Conv2d(1x1)/nn.Linear in
 * reference training_code/cn3d_model_conbag.py:162-207) ------------------------------------------------------ */
FACL_API size_t facl_packed_weight_bytes(int rows, int cols);
/* A[m][k] = src[m*stride_m + k*stride_k]  ->  bf16 hi/lo tile image (see csrc/pack.cu) */
FACL_API int facl_pack_weight(const float* src, long long stride_m, long long stride_k, int rows, int cols, void* image, void* stream);

typedef struct facl_operand {
    const float* src0;   /* fp32 source */
    const float* src1;   /* optional second source, same layout */
    long long ld;
    const float* s0;     /* per-channel transform  v = max(s0*src0 + s1*src1 + s2, lo); NULL = (1, 0, 0, -inf) */
    const float* s1;
    const float* s2;
    const float* lo;
} facl_operand;

typedef struct facl_gemm {
    int Md, Nd, Kd;          /* D[m,n] = sum_k A[m,k] * B[n,k] */
    int nsplit;              /* 1 = bf16 operands, 3 = bf16x3 split ("fp32" mode); fp32 accumulation either way */
    int a_mode;              /* 0 packed image, 1 fp32 row-major [Md][ld] */
    int b_mode;              /* 1 fp32 row-major [Nd][ld], 2 fp32 channel-major [Kd][ld], 3 grouped rows [Nd][4] */
    const void* a_packed;
    int a_packed_kblocks;
    facl_operand a, b;
    int ksplit;              /* >1: split-K, use out_mode 3 (atomic accumulate) */
    const float* bias;       /* [Md] or NULL */
    int out_mode;            /* 0 none, 1 out[m*ldo+n], 2 out[n*ldo+m], 3 atomicAdd out[m*ldo+n] */
    float* out;
    long long ldo;
    const float* zin;        /* optional [Md][ldz]: v *= (zs0[m]*zin + zs2[m] > 0); 2nd statistic becomes sum v*zin */
    long long ldz;
    const float* zs0;
    const float* zs2;
    float* stats;            /* optional [facl_gemm_stat_partials(Md,Nd)][Md][2] */
    int pool;                /* 0 or power of two <= 256: max-pool over groups of `pool` consecutive n */
    const float* pool_sign;  /* [Md] >=0 keep max, <0 keep min */
    float* pool_out;         /* [Md][ldp] */
    unsigned char* pool_arg; /* [Md][ldp] or NULL */
    long long ldp;
} facl_gemm;

FACL_API int facl_gemm_stat_partials(int Md, int Nd);
FACL_API int facl_gemm_tc(const facl_gemm* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FACL_B200_H */
