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
struct PackJob {
    const float* src;
    long long sm, sk;
    int Md, Kd, KBp;
    long long task0;
    uint8_t* img;
};
struct PackTable {
    int n = 0;
    long long total = 0;
    PackJob job[20];
};
void pack_table_add(PackTable& tbl, const float* src, long long sm, long long sk, int Md, int Kd, void* image);
int pack_table_launch(const PackTable& tbl, cudaStream_t st);

}  // namespace facl

namespace facl {
// elementwise.cu
int bn_finalize_launch(const float* partials, int P, int C, double n, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float eps, float momentum, int training, float* mean, float* rstd, float* scale,
                       float* shift, cudaStream_t st);
int bn_bwd_finalize_launch(const float* partials, int P, int C, double n, const float* gamma, const float* mean, const float* rstd,
                           float* dgamma, float* dbeta, int accumulate, float* c0, float* c1, float* c2, cudaStream_t st);
int rowstats_launch(const float* v, const float* z, long long ld, int C, int n, int pairs, float* out, cudaStream_t st);
int transpose_launch(const float* in, long long ldi, float* out, long long ldo, int R, int C, cudaStream_t st);
int seq_pool_launch(const float* pooled, long long ldp, const float* sign, int C, int G, int B, float* seq, long long lds,
                    unsigned char* argg, cudaStream_t st);
int combine_pool_grads_launch(float* dcloud, long long ldc, const float* dseq, long long lds, const unsigned char* argg, int C, int G,
                              int B, cudaStream_t st);
int pool_scatter_launch(const float* v, long long ldv, const unsigned char* arg, long long lda, int C, int groups, int pool,
                        float* dense, long long ldd, cudaStream_t st);
int centres_to_chmajor_launch(const float* c, int R, float* out, long long ldo, cudaStream_t st);
int l2_normalize_launch(const float* x, int rows, int C, float* out, cudaStream_t st);
int fill_launch(float* p, long long n, float v, cudaStream_t st);
int encoder_const_vectors_launch(float* vec, cudaStream_t st);
int adam_launch(const void* table_dev, int ntensors, float lr, float b1, float b2, float eps, int step, cudaStream_t st);
}  // namespace facl

namespace facl {
// profiler.cu
void count_launch(int n = 1);
struct ScopedTimer {
    ScopedTimer(int tag, cudaStream_t st);
    ~ScopedTimer();
    int tag_;
    cudaStream_t st_;
    bool active_;
    cudaEvent_t a_, b_;
};
// timing tags (facl_timing_collect index).  GEMM tags: 3*layer + kind (0 fwd, 1 wgrad, 2 dgrad), layer 0..8
// (7 = netR_FC.3, 8 = mapping); then the non-GEMM kernels.
enum TimingTag {
    TAG_GEMM_BASE = 0,
    TAG_GROUP = 27, TAG_FPS = 28, TAG_PACK = 29, TAG_BN = 30, TAG_POOLMISC = 31, TAG_SCATTER = 32, TAG_LOSS_GEMM = 33,
    TAG_LOSS_MISC = 34, TAG_ADAM = 35, TAG_TRANSPOSE = 36, TAG_MEMSET = 37, TAG_L1_MISC = 38, TAG_L1_PASS_A = 39,
    TAG_L1_PASS_B = 40, TAG_L1_PASS_C = 41, TAG_L1_PASS_D = 42, TAG_IMAGE = 43, TAG_AUGMENT = 44, TAG_GROUP2 = 45, NUM_TIMING_TAGS = 46
};
}  // namespace facl

namespace facl {
// augment.cu
constexpr int AUGMENT_MAX_VIEWS = 32;
constexpr int AUGMENT_MAX_SOURCES = 8;
struct AugmentSource {
    const float* rows;       // (total rows, C) row-major
    const int* offsets;      // (B + 1) first row of each sequence
    int C;
};
struct AugmentRecipe {
    int source, channel, nonzero_only, jitter, mirror, rotate;
};
struct AugmentParams {
    int B, G, N, n_sources, g_major;
    double sigma, clip;
    AugmentSource source[AUGMENT_MAX_SOURCES];
    AugmentRecipe recipe[AUGMENT_MAX_VIEWS];
    const int* idx;          // explicit draws (all three or none) ...
    const double* noise;
    const double* angle_u;
    unsigned long long seed, step;   // ... else Philox4x32-10 keyed by (seed, step)
    float* out;
    int* out_rows;
};
int augment_launch(const AugmentParams& p, int max_rows, cudaStream_t st);
// probe.cu
int softmax_xent_launch(const float* logits, const int* labels, int rows, int C, float* loss, float* dlogits_t, float* dbias, int* hits,
                        cudaStream_t st);
// group2.cu
size_t group_level2_scratch_bytes(int M, int N1, int S2, int K);
int group_level2_launch(const float* feats, int M, int C, int N1, int S2, int K, float r2, float* out, int* idx, void* scratch,
                        cudaStream_t st);
}  // namespace facl

struct facl_encoder_dims;
struct facl_encoder_params;
struct facl_encoder_grads;
namespace facl {
// encoder.cu
int encoder_forward(const facl_encoder_dims* d, const facl_encoder_params* p, const float* xt, const float* centres, void* const* bufs,
                    float* x, float* xg, float* x_nor, float* code, int stages, cudaStream_t st);
int encoder_backward(const facl_encoder_dims* d, const facl_encoder_params* p, const float* xt, void* const* bufs, const float* dx,
                     const float* dxg, const facl_encoder_grads* gr, int stages, cudaStream_t st);
// train_step.cu
int gmajor_launch(const float* in, float* out, int B, int G, int N, cudaStream_t st);
int centres_launch(const float* clouds, int M, int N, int S, float* centres, cudaStream_t st);
}  // namespace facl

namespace facl {
// l1_fused.cu
int l1_fused_grid(long long R);
int l1_bwd_grid(long long R);
void l1_set_debug_dump(unsigned char* mask1, unsigned char* mask2);
int l1_prep_launch(const float* W, int C, const float* d, const float* bias, const float* c2, const float* e0, void* p_img, float* q,
                   void* e_img, cudaStream_t st);
int l1_fin_launch(const float* W, int C, const float* d, const float* bias, const float* c2, const float* H, const float* s,
                  const float* e0, const float* sparse, float* dW, int accumulate, cudaStream_t st);
int l1_bwd_c_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                    const void* w3_img, const void* p3_img, const float* q3, const unsigned char* arg, const float* dpooled,
                    long long ldp, const float* c3_0, void* dh2, float* dw3, float* stats, int K, cudaStream_t st);
int l1_gamma0_fix_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1, const float* shift1,
                         const float* w2, const float* b2, const float* scale2, const void* dh2, float* stats, cudaStream_t st);
int l1_bwd_d_launch(const float* xt, long long R, int nsplit, const float* w1, const float* b1, const float* scale1,
                    const float* shift1, const void* e0w2_img, const void* p2_img, const float* q2, const void* dh2, float* dw2s,
                    float* gram, float* hsum, float* amat, float* stats, cudaStream_t st);
int l1_dw1_launch(const float* amat, int P, const double* mom14, const float* w1, const float* b1, const float* c0, const float* c1,
                  const float* c2, float* dw1, cudaStream_t st);
int l1_moments_launch(const float* xt, long long R, double* mom14, cudaStream_t st);
int l1_bn1_launch(const double* mom14, double n, const float* w1, const float* b1, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float eps, float momentum, int training, float* mean, float* rstd,
                  float* scale, float* shift, cudaStream_t st);
int l1_fwd_launch(bool pass_b, const float* xt, long long R, int K, int nsplit, const float* w1, const float* b1, const float* scale1,
                  const float* shift1, const void* w2_img, const float* b2, const float* scale2, const float* shift2,
                  const void* w3_img, const float* b3, const float* gamma3, int stat_mode, float* stats, float* gram, float* hsum,
                  float* pooled, unsigned char* pool_arg, long long ldp, cudaStream_t st);
}  // namespace facl
