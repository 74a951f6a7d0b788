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

/* ---- level-2 set-abstraction grouping ------------------------------------------------------
 * replaces group_points_2 / group_points_2_3DV, reference training_code/utils_my.py:332-356 / :358-381.
 * feats (M,C,N1) fp32 CHANNEL-first, channels 0..2 = xyz; centres = the first S2 points of each cloud.
 * The K nearest of the N1 points per centre (same selection rule as facl_group_points), slots with squared
 * distance > r2 replaced by the centre; all C channels gathered:
 * out (M,C,S2,K) fp32 channel-first, xyz minus the centre; idx (M,S2,K) int32 or NULL.
 * scratch: facl_group_level2_scratch_bytes() bytes of device memory (xyz row image + neighbour table). */
FACL_API size_t facl_group_level2_scratch_bytes(int M, int N1, int S2, int K);
FACL_API int facl_group_level2(const float* feats, int M, int C, int N1, int S2, int K, float r2, float* out, int* idx,
                               void* scratch, void* stream);

/* ---- training-view augmentation on the device -------------------------------------------
 * replaces NTU_RGBD_new.__getitem__ / get_data_train / get_temporal_augment_data and the helpers they call
 * (reverse_transform, rotate_trans, jitter_point_cloud), reference training_code/cn3D_data_set.py:105-121,
 * :285-350, :654-663, :708-713, :734-748, :765-776, plus the permute + float cast of cn3d_train_motion_GL.py:226-228.
 * One launch turns the source clouds of B sequences into G views of N points each.
 * A source is a ragged batch: rows (sum of P_b, C) fp32 row-major (x,y,z,channels...), offsets (B+1) int32.
 * View g is described by recipes[g]:
 *   source        which source is resampled (with replacement);
 *   channel       the source column copied to output column 3;
 *   nonzero_only  draw only among the rows whose `channel` column is != 0, in row order (:657-658);
 *   jitter        xyz += clip(sigma * z, +-clip)                                  (:765-776)
 *   mirror        round to f32, x = -x, jitter with a second set of normals, round (:708-713)
 *   rotate        round to f32, (x,y,z) . Ry((u - 0.5) * 0.8 * pi), round        (:734-748)
 * applied in that order, in f64 like numpy.  Random draws: either all three arrays are given (parity with a
 * recorded numpy stream: idx (B,G,N) int32, noise (B,G,2,N,3) f64 standard normals, angle_u (B,G) f64 in [0,1)),
 * or all three are NULL and the kernel draws them itself (Philox4x32-10 keyed by seed, step).
 * out: (G,B,N,4) fp32 if g_major (what facl_group_points / facl_train_step consume) else (B,G,N,4);
 * out_rows (B,G,N) int32 or NULL: the source row each output point came from.
 * max_rows: the largest P_b of any source used with nonzero_only (sizes the compaction list; <= 51200).
 * A view with nothing to draw from (numpy raises) is filled with NaN. */
typedef struct {
    const float* rows;
    const int* offsets;
    int C;
} facl_point_source;
typedef struct {
    int source, channel, nonzero_only, jitter, mirror, rotate;
} facl_view_recipe;
typedef struct {
    int B, G, N;
    int n_sources;
    const facl_point_source* sources;   /* host array [n_sources], <= 8 */
    const facl_view_recipe* recipes;    /* host array [G], G <= 32 */
    double sigma, clip;                 /* reference: 0.01, 0.05 (f64 like numpy's arithmetic) */
    const int* idx;
    const double* noise;
    const double* angle_u;
    unsigned long long seed, step;
    int g_major;
    int max_rows;
    float* out;
    int* out_rows;
} facl_augment_args;
FACL_API int facl_augment_views(const facl_augment_args* args, void* stream);

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

/* A pre-converted bf16 activation image: atoms of [8 channels][64 rows] bf16 (1024 bytes, 128-byte swizzle), atom
 * (row block rb, channel group cg) at ((rb * cgs) + cg) * 1024; channels zero-padded to a multiple of 64, rows to a
 * multiple of 64.  The same bytes are an MN-major UMMA operand (reduction over channels) and a K-major one (reduction
 * over rows), so one conversion serves the forward, data-gradient and weight-gradient GEMMs (csrc/gemm_img.cu). */
typedef struct facl_image {
    void* hi;            /* bf16(v) */
    void* lo;            /* bf16(v - hi); used when nsplit == 3 */
    int cgs;             /* 8-channel groups per row block = padded channels / 8 */
    int rbs;             /* 64-row blocks */
} facl_image;

FACL_API size_t facl_act_image_half_bytes(int C, long long R);
/* image[c][r] = max(s0[c]*src0[c][r] + s1[c]*src1[c][r] + s2[c], lo[c]) from fp32 channel-major sources (src->ld, ld1 =
 * leading dimension of src1, 0 = same).  pool > 0 (multiple of 8): src0 and pool_arg are [C][src->ld] over groups of
 * `pool` rows and element (c, r) reads src0[c][r / pool] where pool_arg[c][r / pool] == r % pool, else 0 (the
 * max-pool scatter of a pooled gradient).  Declared after facl_operand below. */

typedef struct facl_gemm {
    int Md, Nd, Kd;          /* D[m,n] = sum_k A[m,k] * B[n,k] */
    int nsplit;              /* 1 = bf16 operands, 3 = bf16x3 split ("fp32" mode); fp32 accumulation either way */
    int a_mode;              /* 0 packed image, 1 fp32 row-major [Md][ld], 2 activation image a_img (K = its rows) */
    int b_mode;              /* 1 fp32 row-major [Nd][ld], 2 fp32 channel-major [Kd][ld], 3 grouped rows [Nd][4],
                                4 activation image b_img, channels = k, rows = n (with a_mode 0),
                                5 activation image b_img, channels = n, K = its rows (with a_mode 2, out_mode 3) */
    const void* a_packed;
    int a_packed_kblocks;
    facl_operand a, b;
    int ksplit;              /* >1: split-K, use out_mode 3 (atomic accumulate) */
    const float* bias;       /* [Md] or NULL */
    int out_mode;            /* 0 none, 1 out[m*ldo+n], 2 out[n*ldo+m], 3 atomicAdd out[m*ldo+n], 4 out[n*ldo+m] += */
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
    facl_image a_img;    /* a_mode 2 */
    facl_image b_img;    /* b_mode 4 / 5 */
} facl_gemm;

FACL_API int facl_act_image(const facl_operand* src, long long ld1, int C, long long R, const unsigned char* pool_arg, int pool,
                            int nsplit, const facl_image* img, void* stream);

FACL_API int facl_gemm_stat_partials(int Md, int Nd);
FACL_API int facl_gemm_tc(const facl_gemm* desc, void* stream);

/* ---- encoder: PointNet_Plus / PointNet_Plus_fine forward + backward ---------------------------------------
 * replaces PointNet_Plus_fine.forward (reference training_code/cn3d_model_conbag.py:213-234; layers :162-210,
 * identical in PointNet_Plus :43-91) and the autograd backward PyTorch derives for it.
 *
 * Layer table (index -> reference state-dict prefix): 0 net3DV_1.0/.1 (4->64)  1 net3DV_1.3/.4 (64->64)
 * 2 net3DV_1.6/.7 (64->256)  3 net3DV_3.0/.1 (259->256)  4 net3DV_3.3/.4 (256->512)  5 net3DV_3.6/.7 (512->1024)
 * 6 netR_FC.0/.1 (1024->1024).  Weights are the reference tensors themselves ([Cout][Cin] row-major fp32).    */
#define FACL_NUM_BN_LAYERS 7

typedef struct facl_layer {
    const float* w;        /* conv / linear weight [Cout][Cin] */
    const float* b;        /* bias [Cout] */
    const float* gamma;    /* BatchNorm weight */
    const float* beta;     /* BatchNorm bias */
    float* running_mean;   /* updated in place in training mode (momentum 0.1, unbiased variance) */
    float* running_var;
} facl_layer;

typedef struct facl_encoder_params {
    facl_layer layer[FACL_NUM_BN_LAYERS];
    const float* fc3_w;    /* netR_FC.3.weight [512][1024] */
    const float* fc3_b;    /* netR_FC.3.bias   [512] */
    const float* map_w;    /* mapping.weight   [64][512] */
} facl_encoder_params;

typedef struct facl_encoder_grads {   /* outputs of the backward; every pointer must be valid */
    float* dw[FACL_NUM_BN_LAYERS];
    float* db[FACL_NUM_BN_LAYERS];
    float* dgamma[FACL_NUM_BN_LAYERS];
    float* dbeta[FACL_NUM_BN_LAYERS];
    float* dfc3_w;
    float* dfc3_b;
} facl_encoder_grads;

typedef struct facl_encoder_dims {
    int M;          /* clouds = G * B, G-major (row g*B + b), cn3d_train_motion_GL.py:225-226 */
    int S;          /* centres per cloud (sample_num_level1) */
    int K;          /* neighbours per centre (knn_K); power of two <= 256 */
    int G;          /* views per sequence (the model's `gost`) */
    int nsplit;     /* 1 = bf16 tensor-core operands, 3 = bf16x3 split (fp32-class accuracy) */
    int training;   /* 1: batch statistics + running-stat update; 0: running statistics (extract_*_feature.py) */
    int flags;      /* FACL_ENC_* */
} facl_encoder_dims;

#define FACL_ENC_FUSED_L1 1   /* net3DV_1 as the fused tcgen05 kernels (csrc/l1_fused.cu) instead of per-layer GEMMs */
#define FACL_ENC_SPLIT_LAYER(l) (1 << (8 + (l)))   /* bf16 mode (nsplit 1): layer l (0..8) keeps the bf16x3 split products in the FORWARD */
#define FACL_ENC_SPLIT_BACKWARD 2                  /* bf16 mode: the backward honours the FACL_ENC_SPLIT_LAYER mask too (default: single products) */

/* Work buffers are owned by the caller: query the count / name / size, allocate each (256-byte aligned device
 * memory) and pass the pointer table.  Buffers flagged "backward only" may be NULL for forward-only use. */
FACL_API int facl_encoder_num_buffers(void);
FACL_API const char* facl_encoder_buffer_name(int i);
FACL_API size_t facl_encoder_buffer_bytes(int i, const facl_encoder_dims* dims);
FACL_API int facl_encoder_buffer_backward_only(int i);

/* xt [M*S*K][4] grouped rows (the physical layout of the reference's xt (M,4,S,K) view), centres [M*S][3]
 * (the physical layout of yt (M,3,S,1)).  Outputs: x [M][512], x_global [M/G][512], x_nor [M][512],
 * code [M][64] (x_nor / code may be NULL). */
FACL_API int facl_encoder_forward(const facl_encoder_dims* dims, const facl_encoder_params* params, const float* xt,
                                  const float* centres, void* const* buffers, float* x, float* x_global, float* x_nor,
                                  float* code, void* stream);
/* dx [M][512], dx_global [M/G][512]: gradients of the loss w.r.t. x / x_global (either may be NULL = zero).
 * Must follow a training-mode forward on the same buffers. */
FACL_API int facl_encoder_backward(const facl_encoder_dims* dims, const facl_encoder_params* params, const float* xt,
                                   void* const* buffers, const float* dx, const float* dx_global,
                                   const facl_encoder_grads* grads, void* stream);

/* ---- optimiser: torch.optim.Adam(lr, betas, eps) step over a table of tensors (reference
 * cn3d_train_motion_GL.py:180,332).  table_dev: device array of {float* p; const float* g; float* m; float* v;
 * long long n} records; step is the 1-based step count used for bias correction. */
FACL_API int facl_adam_step(const void* table_dev, int ntensors, float lr, float beta1, float beta2, float eps, int step,
                            void* stream);

/* ---- K6/K7: contrastive losses, forward + gradient -------------------------------------------------------
 * replaces the inline "global" loss (reference training_code/cn3d_train_motion_GL.py:265-287 = utils_my.py:53-83
 * global_contrast) and "circle" loss (:290-316 = utils_my.py:85-116 circle_contrast).
 *
 * Single GPU: B_local = B, sample_offset = 0, keys = NULL (= x).  x [G*B][C] (G-major rows), x_global [B][C];
 * order: device int32[G], the view permutation the reference draws with np.random.shuffle (:297-298).
 * loss[0] = global, loss[1] = circle (0 when not requested).  Gradients for unit upstream weight:
 *   dx_anchor [G*B_local][C]  from x as the anchor (row) side,      dx_global [B_local][C],
 *   dkeys     [world*G*B_local][C]  from x as the key (column) side; may alias dx_anchor when keys == x, the two
 *   contributions are then summed into it.
 * Multi GPU (batch of B sequences sharded, B_local per rank, this rank owns samples [sample_offset, +B_local)):
 * keys = the all-gather of every rank's x, rank-major; losses are this rank's anchors' share (already divided by B);
 * dkeys must be sum-reduce-scattered back to the ranks and added to dx_anchor.  C must be a multiple of 4. */
FACL_API size_t facl_contrast_workspace_bytes(int G, int B_local, int world, int C);
FACL_API int facl_contrast_losses(const float* x, const float* x_global, const float* keys, int G, int B, int B_local,
                                  int sample_offset, int C, const int* order, int want_global, int want_circle, int nsplit,
                                  void* workspace, float* loss, float* dx_anchor, float* dx_global, float* dkeys,
                                  void* stream);

/* ---- one training step in one call -------------------------------------------------------------------------
 * replaces the loop body of reference training_code/cn3d_train_motion_GL.py:224-335 (and the identical
 * cn3d_train_apperance_GL.py): G-major flatten (:225-226) -> H2D (:228) -> group_points_3DV (:230) -> netR (:234)
 * -> global loss (:265-287) -> circle loss (:290-316) -> backward + Adam step (:329-332) -> loss.item() (:335).
 * All scratch is caller-owned.  loss2: device float[3] = {global, circle, total}. */
#define FACL_MAX_VIEWS 256           /* G <= 255: the sequence max-pool records its winning view in one byte */
typedef struct facl_train_step_args {
    const facl_encoder_dims* dims;
    const facl_encoder_params* params;
    const facl_encoder_grads* grads;
    void* const* enc_buffers;        /* facl_encoder_* work buffers, backward set included */
    int N;                           /* points per cloud */
    float r2;                        /* squared ball radius */
    const float* points_bgnd;        /* device (B,G,N,4) batch, or NULL when points_host is given */
    const float* points_host;        /* optional PINNED host (B,G,N,4) batch: copied into `staging` first */
    float* staging;                  /* device (B,G,N,4), used with points_host */
    float* clouds;                   /* device (G*B,N,4) scratch */
    float* xt;                       /* device (G*B,S,K,4) scratch */
    float* centres;                  /* device (G*B*S,3) scratch */
    float* x;                        /* device (G*B,512) */
    float* x_global;                 /* device (B,512) */
    int* order;                      /* device int32[G]: view permutation of the circle loss (read; written first when order_by_value) */
    void* loss_ws;                   /* facl_contrast_workspace_bytes(G,B,512) */
    float* loss2;                    /* device float[3] */
    float* dx;                       /* device (G*B,512) scratch */
    float* dx_global;                /* device (B,512) scratch */
    const void* adam_table;          /* see facl_adam_step */
    int adam_ntensors;
    float lr, beta1, beta2, eps;
    int step;                        /* 1-based optimiser step */
    float* loss_host;                /* optional PINNED host float: receives the total loss (async D2H) */
    /* --- sharded (multi-GPU) operation: the step is issued in phases with the collectives in between --- */
    int phases;                      /* bit mask of FACL_PHASE_*; 0 = all (single GPU) */
    int B_global;                    /* sequences over all ranks (0 = this rank's batch) */
    int sample_offset;               /* first global sample index owned by this rank */
    const float* keys;               /* all-gathered x of every rank, rank-major (NULL = x) */
    float* dkeys;                    /* (world*G*B_local,512): key-side gradient, to be sum-reduce-scattered */
    const float* dx_extra;           /* (G*B_local,512): this rank's slice of the reduced dkeys, added to dx */
    /* --- the view permutation passed BY VALUE: the reference draws it on the host every step (np.random.shuffle,
     * cn3d_train_motion_GL.py:297-298); carried in the launch arguments it needs no host staging buffer whose reuse
     * could race with a pending asynchronous copy --- */
    int order_by_value;              /* non-zero: the LOSS phase first writes order_vals[0..G) into `order` */
    int order_vals[FACL_MAX_VIEWS];
} facl_train_step_args;

#define FACL_PHASE_FORWARD 1         /* H2D, G-major flatten, grouping, encoder forward  -> x, x_global */
#define FACL_PHASE_LOSS 2            /* losses + dL/dx (anchor side), dL/dx_global, dL/dkeys */
#define FACL_PHASE_BACKWARD 4        /* dx += dx_extra; encoder backward -> parameter gradients */
#define FACL_PHASE_UPDATE 8          /* Adam step, loss D2H */
#define FACL_PHASE_ALL 15
#define FACL_PHASE_BACKWARD_HEAD 16  /* first half of BACKWARD: dx += dx_extra; head + net3DV_3 -> all gradients but net3DV_1's */
#define FACL_PHASE_BACKWARD_L1 32    /* second half of BACKWARD: net3DV_1 (the all-reduce of the rest can run beside it) */
/* finer cuts for a sharded caller that hides its two embedding exchanges as well (facl_b200/dist.py):
 *   FORWARD = FORWARD_X then FORWARD_G;  BACKWARD_HEAD = BACKWARD_HEAD_G then BACKWARD_HEAD_X */
#define FACL_PHASE_FORWARD_X 64        /* FORWARD up to the cloud embeddings x (the all-gather of x can start) */
#define FACL_PHASE_FORWARD_G 128       /* the head on the sequence features -> x_global */
#define FACL_PHASE_BACKWARD_HEAD_G 256 /* sequence half of the head backward: needs dx_global only (runs beside the reduce-scatter of dkeys) */
#define FACL_PHASE_BACKWARD_HEAD_X 512 /* dx += dx_extra; cloud half of the head backward + net3DV_3 */

FACL_API int facl_gmajor(const float* points_bgnd, float* clouds, int B, int G, int N, void* stream);
FACL_API int facl_train_step(const facl_train_step_args* args, void* stream);

/* ---- linear probe on the extracted features -------------------------------------------
 * replaces the non-GEMM parts of Final_FC (reference linear_classify/fc_model.py:12-25: F.normalize(x, p=2, dim=1)
 * then nn.Linear) and of the loop in linear_classify/linercls.py:106-124 (CrossEntropyLoss, top-1 accuracy); the
 * logits and weight-gradient GEMMs go through facl_gemm_tc.
 * facl_l2_normalize: out[r] = x[r] / max(|x[r]|_2, 1e-12), x and out (rows, C) fp32.
 * facl_softmax_xent: logits (rows, C) fp32, labels (rows) int32.  Accumulates (memset first) the mean cross-entropy
 * into *loss, d loss / d bias into dbias [C] and the number of top-1 hits into *hits; writes d loss / d logits
 * TRANSPOSED into dlogits_t [C][rows] (the A operand of the weight-gradient GEMM).  Any output may be NULL. */
FACL_API int facl_l2_normalize(const float* x, int rows, int C, float* out, void* stream);
FACL_API int facl_softmax_xent(const float* logits, const int* labels, int rows, int C, float* loss, float* dlogits_t, float* dbias,
                               int* hits, void* stream);

/* ---- the loss heads the reference scripts disable by constants ------------------------------------------
 * facl_sinkhorn replaces distributed_sinkhorn + shoot_infs (reference training_code/cn3d_model_conbag.py:391-425): q [K][B] fp32
 * (K prototypes x B samples, already exponentiated as at cn3d_train_motion_GL.py:254-255) -> out [B][K], the rows of the
 * Sinkhorn-Knopp normalised assignment (`(Q / sum_k Q).t()`); infinities are replaced by the maximum of the finite entries.
 * facl_soft_xent replaces the inner SwAV term `-mean(sum(q * log(softmax(code / 0.1)), dim=1))` (cn3d_train_motion_GL.py:258-260):
 * logits, q [rows][K]; accumulates the loss into *loss (memset first) and writes d loss / d logits [rows][K]; either may be NULL.
 * facl_kmeans replaces KMeans (reference training_code/utils_my.py:180-198 = cn3d_train_motion_GL.py:52-70): x [N][D] fp32, centroids
 * initialised to the first K rows, `iters` rounds of (assign to the FIRST nearest centroid, replace centroids by member means, empty
 * clusters divide by 1) -> labels [N] int32 of the last assignment, centroids [K][D], counts [K] int32 (the divisors; may be NULL). */
FACL_API int facl_sinkhorn(const float* q, int K, int B, int iters, float* out, void* stream);
FACL_API int facl_soft_xent(const float* logits, const float* q, int rows, int K, float scale, float* loss, float* dlogits, void* stream);
FACL_API int facl_kmeans(const float* x, int N, int D, int K, int iters, int* labels, float* centroids, int* counts, void* stream);

/* ---- instrumentation: per-kernel device timing and launch counting (used by bench.py) ---------------------
 * facl_timing_enable(1): every tagged launch site is bracketed by CUDA events on its stream.
 * facl_timing_collect: synchronises on the recorded events, returns total ms / launches per tag and resets.
 * Tags: 3*layer + {0 forward, 1 weight-grad, 2 data-grad} for layer 0..8 (7 = netR_FC.3, 8 = mapping), then
 * 27 grouping, 28 fps, 29 weight packing, 30 BN finalize, 31 pooling misc, 32 max-pool scatter, 33 loss GEMMs,
 * 34 loss misc, 35 adam, 36 transposes, 37 memset/fill, 38 fused-L1 misc, 39..42 fused-L1 passes A, B, C, D,
 * 43 activation-image conversion, 44 view augmentation, 45 level-2 grouping (xyz image + channel gather).
 * facl_launch_count: kernels launched so far. */
#define FACL_NUM_TIMING_TAGS 46
FACL_API void facl_timing_enable(int on);
FACL_API int facl_timing_collect(float* ms_per_tag, int* count_per_tag, int ntags);
FACL_API long long facl_launch_count(void);
/* test hook: when set (device pointers), the next fused net3DV_1 backward also records the ReLU decisions of its
 * recomputed forward -- mask1 / mask2 [R/64][64][64] bytes (ReLU1 / ReLU2 active, 64-row blocks x channel x row).  The
 * max-pool winners are in the encoder work buffer "arg3" ([256][M*S] bytes, position inside the K-group). */
FACL_API void facl_debug_l1_dump(unsigned char* mask1, unsigned char* mask2);

#ifdef __cplusplus
}
#endif
#endif /* FACL_B200_H */
