// Generic tcgen05 GEMM with fused operand prologues and reduction epilogues.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]          m < Md ("channels", TMEM lanes), n < Nd ("rows", TMEM columns)
//
// Both operands are staged in shared memory as K-major bf16 tiles (128-byte rows, 128B swizzle) and
// multiplied by tcgen05.mma (kind::f16, fp32 accumulators in TMEM).  nsplit = 3 runs the error-compensated
// bf16x3 scheme (hi*hi + hi*lo + lo*hi), which is the "fp32" arithmetic mode of this library.
//
// Because the accumulator tile has output channels on TMEM lanes, an epilogue thread owns one channel and walks
// its rows in registers: per-channel batch statistics, neighbourhood max-pooling and ReLU masks are in-thread
// reductions -- no shuffles, no atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace facl {

// element transform applied while an fp32 source is converted into a bf16 operand tile:
//     v = max(s0[ch] * src0 + s1[ch] * src1 + s2[ch], lo[ch])
// null vectors mean s0 = 1, s1 = 0 (src1 unused), s2 = 0, lo = -inf.  `ch` is the K index for channel-major
// sources and the tile-row index (m or n) for row-major sources.
struct OperandSrc {
    const float* src0;
    const float* src1;
    long long ld;        // leading dimension in elements
    const float* s0;
    const float* s1;
    const float* s2;
    const float* lo;
};

// A pre-converted bf16 activation image (gemm_img.cu).  Atoms of [8 channels][64 rows] bf16 = 1024 bytes with the
// 128-byte swizzle (the 16-byte chunk holding rows 8j..8j+7 of channel c sits at chunk (j ^ (c & 7)) of the 128-byte row
// c & 7); atom (rb, cg) = rows 64 rb.., channels 8 cg.. lives at ((rb * cgs) + cg) * 1024.  The SAME bytes are an
// MN-major UMMA operand when the GEMM reduces over channels and a K-major operand when it reduces over rows.
// Channels are padded with zeros to a multiple of 64 (cgs multiple of 8), rows to a multiple of 64.
struct ActImage {
    const void* hi;             // bf16(v)
    const void* lo;             // bf16(v - hi), nsplit == 3 only
    int cgs;                    // 8-channel groups per row block (padded channel count / 8)
    int rbs;                    // 64-row blocks
};

enum : int { A_PACKED = 0, A_ROWMAJOR = 1, A_IMAGE = 2 };
enum : int { B_ROWMAJOR = 1, B_CHMAJOR = 2, B_XT4 = 3, B_IMAGE_MN = 4, B_IMAGE_K = 5 };
enum : int { OUT_NONE = 0, OUT_CHMAJOR = 1, OUT_ROWMAJOR = 2, OUT_ATOMIC_CHMAJOR = 3, OUT_ROWMAJOR_ACC = 4,
              OUT_ATOMIC_ROWMAJOR = 5 /* gemm_img.cu only: atomicAdd out[n*ldo + m], for split-K */ };

// Contrastive-loss epilogues of the similarity GEMM S = anchors x keys^T (gemm_img.cu, losses.cu).  The GEMM's rows (TMEM lanes)
// are anchors, its columns keys; S itself is never written.
//   mode 1 (forward) : per anchor row an ONLINE log-sum-exp over the unmasked columns -- running (max, sum exp) kept in the
//                      epilogue thread across all tiles of its CTA, one partial per (CTA, row) -> part[cta][Md][2]; the masked
//                      columns (keys of the anchor's own sample: the positives among them) are stored to pos[row][view]
//   mode 2 (backward): S is recomputed and turned into dL/dS in registers (softmax weight of the negatives, the stored
//                      coefficient at the positive, 0 at the other masked columns), written as a bf16 hi/lo operand image
// Index conventions as in losses.cu `Idx`: key j = r*Ml + g*Bl + b is view g of sample r*Bl + b.
struct LossEpi {
    int mode;                   // 0 none
    int G, B, Bl, Ml, n0;       // views, global batch, local batch, G*Bl, first global sample of this rank
    int row0;                   // anchor index of GEMM row 0: 0 for the view anchors (S_x), Ml for the sequence anchors (S_g)
    float* part;                // mode 1: [ctas per m-tile][Md][2]
    float* pos;                 // mode 1: [(Ml + Bl)][G]
    const float* lc;            // mode 2: lcC (row0 == 0) or lcG (row0 == Ml), per local sample
    const float* pg;            // mode 2: pgC / pgG  [Bl][G]
    const int* order;           // mode 2, circle rows
    const int* inv_order;
    void* ds_hi;                // mode 2: dS image [Md channels][Nd rows]
    void* ds_lo;
    int ds_cgs, ds_rbs;
};

struct GemmParams {
    int Md, Nd, Kd;
    int nsplit;                 // 1 (bf16) or 3 (bf16x3)
    int a_mode, b_mode;
    const void* a_packed;       // A_PACKED: [m_tile][k_blk][hi, lo][128 rows x 128 B], pre-swizzled
    int a_packed_kblocks;       // k-blocks per m_tile in the packed image
    OperandSrc a;               // A_ROWMAJOR: fp32 [Md][ld], K contiguous
    OperandSrc b;               // B_ROWMAJOR: fp32 [Nd][ld]; B_CHMAJOR: fp32 [Kd][ld]; B_XT4: fp32 [Nd][4]
    int ksplit;                 // 1: every CTA walks the whole K range; >1: split-K, one output tile per CTA
    // ---- epilogue ----
    const float* bias;          // [Md] or null
    int out_mode;
    float* out;
    long long ldo;
    const float* zin;           // optional fp32 [Md][ldz]: ReLU mask source and second statistic
    long long ldz;
    const float* zs0;           // mask = (zs0[c] * zin + zs2[c] > 0)
    const float* zs2;
    float* stats;               // optional [ctas_per_mtile][Md][2] partial sums: (sum v, sum v*v) or (sum v, sum v*zin)
    int pool;                   // 0, or the number of consecutive columns max-pooled (divides 256)
    const float* pool_sign;     // [Md]: >= 0 keep the max, < 0 keep the min (sign of the BN scale that follows)
    float* pool_out;            // [Md][ldp] selected pre-activation value
    unsigned char* pool_arg;    // [Md][ldp] position inside the group (first hit), or null
    long long ldp;
    int tag;                    // timing tag (profiler.cu), -1 = untimed
    // ---- pre-converted activation images (gemm_img.cu): operands staged by bulk TMA only, no conversion warps ----
    ActImage a_img;             // A_IMAGE: channels = m, reduction over the image's rows (weight-gradient GEMMs)
    ActImage b_img;             // B_IMAGE_MN: channels = k, rows = n (forward / data-gradient GEMMs);
                                // B_IMAGE_K: channels = n, reduction over rows
    LossEpi loss;               // loss.mode != 0: contrastive-loss epilogue instead of an output matrix
};

// bytes of one half (hi or lo) of an activation image of C channels x R rows
size_t act_image_half_bytes(int C, long long R);
// build an image from an fp32 channel-major source [C][ld] with the OperandSrc transform (per-channel constants).
// pooled source (optional): src0 is then [C][ld] over GROUPS of `pool` rows and pool_arg [C][ld] the winner position
// inside each group -- element (c, r) reads src0[c][r / pool] if pool_arg[c][r / pool] == r % pool, else 0.
// ld1 = leading dimension of src.src1 (0: same as src.ld).
int act_image_launch(const OperandSrc& src, long long ld1, int C, long long R, const unsigned char* pool_arg, int pool, int nhl,
                     const ActImage& img, int tag, cudaStream_t st);
// gemm_img.cu: the TMA-only kernel behind launch_gemm_tc for the *_IMAGE operand modes
int launch_gemm_img(const GemmParams& p, cudaStream_t stream);

// host launcher (gemm_tc.cu); returns cudaError_t as int
int launch_gemm_tc(const GemmParams& p, cudaStream_t stream);
// CTAs that share one m-tile in the non-split schedule == number of stats partials per channel
int gemm_tc_ctas_per_mtile(int Md, int Nd);

}  // namespace facl
