// Encoder forward / backward orchestration (generic-GEMM schedule).
//
// Replaces PointNet_Plus_fine.forward (reference training_code/cn3d_model_conbag.py:213-234) and its autograd
// backward.  Every Conv2d(1x1)/Linear is a tcgen05 GEMM whose PROLOGUE applies the previous layer's
// BatchNorm+ReLU while converting the fp32 activation to bf16 operand tiles, and whose EPILOGUE accumulates the
// batch statistics the next BatchNorm needs (and, for the last layer of each stack, the max-pool).  Activations
// are kept pre-BN ("z"), channel-major [C][rows], fp32.
//
// Train-mode BatchNorm + ReLU + max-pool identities used:
//   max_k relu(a*z_k + b) = relu(a * (a >= 0 ? max_k z_k : min_k z_k) + b)            -> pool before BN
//   dz = c0*dy + c1*z + c2   with c0 = g*rstd, c1 = -c0*rstd*dgamma/n, c2 = -c0*dbeta/n - c1*mean
//   a conv/linear bias in front of a train-mode BN has zero gradient.
#include <math.h>
#include <string.h>

#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"
#include "gemm_tc.cuh"

namespace facl {

namespace {

const int CIN[7] = {4, 64, 64, 259, 256, 512, 1024};
const int COUT[7] = {64, 64, 256, 256, 512, 1024, 1024};
const int C_EMB = 512, C_MAP = 64, C_FEAT = 1024;
const float BN_EPS = 1e-5f, BN_MOM = 0.1f;

enum Buf {
    B_Z1, B_Z2, B_Z3, B_ARG3, B_PCAT, B_Z4, B_Z5, B_Z6, B_ARG6, B_PALL, B_ARGG, B_Z7, B_BN, B_VEC, B_STATS, B_WPACK,
    B_IMG_H3, B_IMG_H4, B_IMG_H5, B_IMG_HD,
    B_DXT, B_DH7, B_DF, B_DY6, B_DH5, B_DH4, B_DP3, B_DY3, B_DH2, B_DH1, B_XTT, B_WPACKT, B_IMG_DZ, B_L1S, B_IMG_HDB,
    NUM_BUFS
};
const char* BUF_NAMES[NUM_BUFS] = {"z1", "z2", "z3", "arg3", "pcat", "z4", "z5", "z6", "arg6", "pall", "argg", "z7", "bn", "vec",
                                   "stats", "wpack", "img_h3", "img_h4", "img_h5", "img_hd", "dxt", "dh7", "df", "dy6", "dh5", "dh4", "dp3", "dy3", "dh2", "dh1", "xtt",
                                   "wpackt", "img_dz", "l1s", "img_hdb"};
const int FIRST_BWD_BUF = B_DXT;

// forward weight images: layers 0..6, then fc3 (512x1024), then mapping (64x512)
size_t wpack_offset(int idx) {
    size_t off = 0;
    for (int l = 0; l < idx; ++l) {
        if (l < 7) off += packed_weight_bytes(COUT[l], CIN[l]);
        else if (l == 7) off += packed_weight_bytes(C_EMB, C_FEAT);
        else off += packed_weight_bytes(C_MAP, C_EMB);
    }
    return off;
}
// transposed images for the data-gradient GEMMs: layers 1..6 ([Cin'][Cout], layer 3 without its 3 xyz inputs), then fc3^T
int tin(int l) { return l == 3 ? 256 : CIN[l]; }
size_t wpackt_offset(int idx) {   // idx 1..6 layers, 7 = fc3^T, 8 = end
    size_t off = 0;
    for (int l = 1; l < idx; ++l) {
        if (l < 7) off += packed_weight_bytes(tin(l), COUT[l]);
        else off += packed_weight_bytes(C_FEAT, C_EMB);
    }
    return off;
}

// scratch of the fused net3DV_1 backward: 64x64 accumulators (floats), then three packed 64x64 operand images
enum : int { L1S_H2 = 0, L1S_S2 = 4096, L1S_H1 = 4160, L1S_S1 = 8256, L1S_DW2S = 8320, L1S_ACC_END = 12416, L1S_Q3 = 12416,
             L1S_Q2 = 12480 };
const size_t L1S_IMG_OFF = 65536, L1S_BYTES = 65536 + 3 * 32768;

// net3DV_3 runs on pre-converted activation images (gemm_img.cu) whenever a pooling group is a whole number of chunks
bool use_images(const facl_encoder_dims* d) { return d->S % 8 == 0; }

size_t buffer_bytes(int i, const facl_encoder_dims* d) {
    const size_t M = d->M, R3 = (size_t)d->M * d->S, R1 = R3 * d->K, B = d->M / d->G, MB = M + B, f = sizeof(float);
    if (d->flags & FACL_ENC_FUSED_L1) {   // no per-row activation is stored: only dh2 (pass C -> pass D) and small scratch
        switch (i) {
            case B_Z1: case B_Z2: case B_Z3: case B_DY3: case B_XTT: return 256;
            case B_DH1: return (size_t)4 * kNumSMs * 64 * 4 * f;   // pass D: [4 * grid][64][4]
            default: break;
        }
    }
    switch (i) {
        case B_Z1: case B_Z2: case B_DH2: case B_DH1: return 64 * R1 * f;
        case B_Z3: case B_DY3: return 256 * R1 * f;
        case B_ARG3: return 256 * R3;
        case B_PCAT: return 259 * R3 * f;
        case B_Z4: case B_DH4: case B_DP3: return 256 * R3 * f;
        case B_Z5: case B_DH5: return 512 * R3 * f;
        case B_DY6: return use_images(d) ? 256 : 1024 * R3 * f;
        case B_Z6: return 1024 * R3 * f;
        case B_ARG6: return 1024 * MB;
        case B_PALL: case B_Z7: case B_DH7: case B_DF: return 1024 * MB * f;
        case B_ARGG: return 1024 * B;
        case B_BN: return 8 * 7 * 1024 * f;
        case B_VEC: return 2048 * f;
        case B_STATS: return 131072 * f;
        case B_WPACK: return wpack_offset(9);
        case B_DXT: return 512 * MB * f;
        case B_XTT: return 4 * R1 * f;
        case B_WPACKT: return wpackt_offset(8);
        case B_IMG_H3: return use_images(d) ? 2 * act_image_half_bytes(259, (long long)R3) : 256;
        case B_IMG_H4: return use_images(d) ? 2 * act_image_half_bytes(256, (long long)R3) : 256;
        case B_IMG_H5: return use_images(d) ? 2 * act_image_half_bytes(512, (long long)R3) : 256;
        case B_IMG_DZ: return use_images(d) ? 2 * act_image_half_bytes(1024, (long long)R3) : 256;
        case B_L1S: return L1S_BYTES;
        case B_IMG_HD: return 2 * (2 * act_image_half_bytes(1024, (long long)M) + 2 * act_image_half_bytes(1024, (long long)B));
        case B_IMG_HDB: return 2 * (act_image_half_bytes(512, (long long)M) + act_image_half_bytes(1024, (long long)M) +
                                    act_image_half_bytes(512, (long long)B) + act_image_half_bytes(1024, (long long)B));
    }
    return 0;
}

struct Slot {
    float *mean, *rstd, *scale, *shift, *c0, *c1, *c2;
};
Slot bn_slot(void* const* bufs, int i) {
    float* base = reinterpret_cast<float*>(bufs[B_BN]) + (size_t)i * 7 * 1024;
    float* vec = reinterpret_cast<float*>(bufs[B_VEC]);
    Slot s{base, base + 1024, base + 2048, base + 3072, base + 4096, base + 5120, base + 6144};
    if (i == 2) {   // BN of the last L1 layer feeds the 259-channel L3 input: its scale/shift live inside the input vectors
        s.scale = vec + 3;
        s.shift = vec + 320 + 3;
    }
    return s;
}

GemmParams gemm_base(int Md, int Nd, int Kd, int nsplit) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.Md = Md; p.Nd = Nd; p.Kd = Kd; p.nsplit = nsplit; p.ksplit = 1;
    p.a_mode = A_PACKED; p.b_mode = B_CHMAJOR;
    p.tag = -1;
    return p;
}
void set_packed_a(GemmParams& p, const void* img, int Kd) {
    p.a_mode = A_PACKED;
    p.a_packed = img;
    p.a_packed_kblocks = (Kd + 63) / 64;
}
OperandSrc src1(const float* s, long long ld, const float* s0, const float* s2, const float* lo) {
    OperandSrc o;
    memset(&o, 0, sizeof(o));
    o.src0 = s; o.ld = ld; o.s0 = s0; o.s2 = s2; o.lo = lo;
    return o;
}
OperandSrc src2(const float* a, const float* b, long long ld, const float* c0, const float* c1, const float* c2) {
    OperandSrc o;
    memset(&o, 0, sizeof(o));
    o.src0 = a; o.src1 = b; o.ld = ld; o.s0 = c0; o.s1 = c1; o.s2 = c2;
    return o;
}
// Arithmetic mode per layer.  nsplit 3 ("fp32" mode) everywhere, or in "bf16" mode: single-pass bf16 for the layers
// that carry the FLOPs (1..5) and the split scheme for the ones that are free -- the 4-wide first layer and the
// head, which runs on M + B rows only and sits behind a small-batch BatchNorm1d that amplifies rounding.
// `ns` carries d->nsplit in its low byte and, above it, the FACL_ENC_SPLIT_LAYER mask of layers that keep the split scheme in
// bf16 mode (mode_of() below).
int layer_nsplit(int ns, int layer) {
    if ((ns & 0xFF) == 3) return 3;
    if ((ns >> 8) & (1 << layer)) return 3;
    return (layer == 0 || layer >= 6) ? 3 : 1;
}
int mode_of(const facl_encoder_dims* d) { return d->nsplit | (((d->flags >> 8) & 0x1FF) << 8); }
// the fused net3DV_1 kernels take one arithmetic mode for layers 1 and 2 together
int l1_nsplit(int ns) { return (layer_nsplit(ns, 1) == 3 || layer_nsplit(ns, 2) == 3) ? 3 : 1; }

int layer_nhl(int ns, int layer) { return layer_nsplit(ns, layer) == 3 ? 2 : 1; }

ActImage image_of(void* const* bufs, int buf, int C, long long R) {
    ActImage im;
    im.hi = bufs[buf];
    im.lo = reinterpret_cast<const uint8_t*>(bufs[buf]) + act_image_half_bytes(C, R);
    im.cgs = ((C + 63) / 64) * 8;
    im.rbs = (int)((R + 63) / 64);
    return im;
}

// head images: img_hd holds, per half (cloud rows / sequence rows), the inputs of netR_FC.0 and netR_FC.3;
// img_hdb the two gradient images of the backward
ActImage image_at(uint8_t* base, size_t& off, int C, long long R) {
    ActImage im;
    const size_t hb = act_image_half_bytes(C, R);
    im.hi = base + off;
    im.lo = base + off + hb;
    im.cgs = ((C + 63) / 64) * 8;
    im.rbs = (int)((R + 63) / 64);
    off += 2 * hb;
    return im;
}
struct HeadImages { ActImage h6, h7, dx, dz; };
HeadImages head_images(void* const* bufs, int half, int M, int B) {
    HeadImages hi{};
    size_t off = 0, offb = 0;
    uint8_t* f = reinterpret_cast<uint8_t*>(bufs[B_IMG_HD]);
    uint8_t* b = reinterpret_cast<uint8_t*>(bufs[B_IMG_HDB]);
    for (int h = 0; h <= half; ++h) {
        const long long R = h == 0 ? M : B;
        hi.h6 = image_at(f, off, 1024, R);
        hi.h7 = image_at(f, off, 1024, R);
        if (b) {
            hi.dx = image_at(b, offb, 512, R);
            hi.dz = image_at(b, offb, 1024, R);
        }
    }
    return hi;
}

int pick_ksplit(int Md, int Nd, int Kd) {
    int tiles = ((Md + 127) / 128) * ((Nd + 255) / 256);
    int KB = (Kd + 63) / 64;
    int ks = kNumSMs / tiles;
    if (ks > KB) ks = KB;
    return ks < 1 ? 1 : ks;
}

#define RUN(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != 0) return _rc;      \
    } while (0)

int check_dims(const facl_encoder_dims* d) {
    if (!d || d->M <= 0 || d->G <= 0 || d->M % d->G != 0) return (int)cudaErrorInvalidValue;
    if (d->K <= 0 || d->K > 256 || (d->K & (d->K - 1))) return (int)cudaErrorInvalidValue;
    if (d->S <= 0 || d->S > 256 || (d->S & (d->S - 1))) return (int)cudaErrorInvalidValue;
    if (d->nsplit != 1 && d->nsplit != 3) return (int)cudaErrorInvalidValue;
    if (d->G > 255) return (int)cudaErrorInvalidValue;   // the sequence max-pool records its winning view in one byte (seq_pool_kernel)
    return 0;
}

}  // namespace

// stages: 1 = everything up to the cloud embeddings x (and x_nor / code), 2 = the head on the sequence features -> x_global, 3 = both.
// A sharded caller starts the all-gather of x after stage 1; the sequence half of the head then runs beside the collective.
int encoder_forward(const facl_encoder_dims* d, const facl_encoder_params* p, const float* xt, const float* centres,
                    void* const* bufs, float* x, float* xg, float* x_nor, float* code, int stages, cudaStream_t st) {
    RUN(check_dims(d));
    if ((stages & 3) == 0) return (int)cudaErrorInvalidValue;
    const bool stage_x = (stages & 1) != 0, stage_g = (stages & 2) != 0;
    const int M = d->M, S = d->S, K = d->K, G = d->G, B = M / G, ns = mode_of(d), tr = d->training;
    const long long R3 = (long long)M * S, R1 = R3 * K, MB = M + B;
    if (R1 > 2147483647LL) return (int)cudaErrorInvalidValue;
    auto F = [&](int i) { return reinterpret_cast<float*>(bufs[i]); };
    auto U = [&](int i) { return reinterpret_cast<unsigned char*>(bufs[i]); };
    float* vec = F(B_VEC);
    float* lo0 = vec + 960;
    float* stats = tr ? F(B_STATS) : nullptr;
    uint8_t* wp = reinterpret_cast<uint8_t*>(bufs[B_WPACK]);

    if (stage_x) {
    // constant vectors: L3 input transform for the 3 centre-xyz channels (identity, no ReLU), zero lower bounds
    RUN(encoder_const_vectors_launch(vec, st));
    // operand images of the current weights
    {
        PackTable tbl;
        for (int l = 0; l < 7; ++l) pack_table_add(tbl, p->layer[l].w, CIN[l], 1, COUT[l], CIN[l], wp + wpack_offset(l));
        pack_table_add(tbl, p->fc3_w, C_FEAT, 1, C_EMB, C_FEAT, wp + wpack_offset(7));
        if (code) pack_table_add(tbl, p->map_w, C_EMB, 1, C_MAP, C_EMB, wp + wpack_offset(8));
        if (tr && bufs[B_WPACKT]) {
            // transposed images for the data-gradient GEMMs of the backward that follows on these buffers
            uint8_t* wt = reinterpret_cast<uint8_t*>(bufs[B_WPACKT]);
            for (int l = 1; l < 7; ++l)
                pack_table_add(tbl, p->layer[l].w + (l == 3 ? 3 : 0), 1, CIN[l], tin(l), COUT[l], wt + wpackt_offset(l));
            pack_table_add(tbl, p->fc3_w, 1, C_FEAT, C_FEAT, C_EMB, wt + wpackt_offset(7));
        }
        RUN(pack_table_launch(tbl, st));
    }
    }

    auto finalize = [&](int layer, int slot_i, int Nd, double n) {
        Slot s = bn_slot(bufs, slot_i);
        const facl_layer& L = p->layer[layer];
        return bn_finalize_launch(stats, gemm_tc_ctas_per_mtile(COUT[layer], Nd), COUT[layer], n, L.gamma, L.beta, L.running_mean,
                                  L.running_var, BN_EPS, BN_MOM, tr, s.mean, s.rstd, s.scale, s.shift, st);
    };

    auto trunk = [&]() -> int {
    if (d->flags & FACL_ENC_FUSED_L1) {
        // ---- L1 fused: three launches, no per-row activation ever reaches HBM (l1_fused.cu) -------------------------
        if ((R1 % 128) != 0 || 128 % K != 0) return (int)cudaErrorInvalidValue;
        double* mom = reinterpret_cast<double*>(vec + 1984);
        Slot s0 = bn_slot(bufs, 0), s1 = bn_slot(bufs, 1);
        const facl_layer& L0 = p->layer[0];
        if (tr) RUN(l1_moments_launch(xt, R1, mom, st));
        RUN(l1_bn1_launch(mom, (double)R1, L0.w, L0.b, L0.gamma, L0.beta, L0.running_mean, L0.running_var, BN_EPS, BN_MOM, tr,
                          s0.mean, s0.rstd, s0.scale, s0.shift, st));
        const int grid = l1_fused_grid(R1);
        if (tr)
            RUN(l1_fwd_launch(false, xt, R1, K, l1_nsplit(ns), L0.w, L0.b, s0.scale, s0.shift, wp + wpack_offset(1), p->layer[1].b, nullptr,
                              nullptr, nullptr, nullptr, nullptr, 1, stats, nullptr, nullptr, nullptr, nullptr, 0, st));
        {
            const facl_layer& L = p->layer[1];
            RUN(bn_finalize_launch(stats, 2 * grid, 64, (double)R1, L.gamma, L.beta, L.running_mean, L.running_var, BN_EPS, BN_MOM,
                                   tr, s1.mean, s1.rstd, s1.scale, s1.shift, st));
        }
        // training with the backward buffers present: pass B also accumulates H2 = sum h2 h2^T and s2 = sum h2 (tensor core), which
        // the dense part of dW3 needs -- the backward then does not recompute them
        float* l1s = (tr && bufs[B_L1S]) ? F(B_L1S) : nullptr;
        if (l1s) FACL_CHECK(cudaMemsetAsync(l1s + L1S_H2, 0, sizeof(float) * (L1S_H1 - L1S_H2), st));
        RUN(l1_fwd_launch(true, xt, R1, K, l1_nsplit(ns), L0.w, L0.b, s0.scale, s0.shift, wp + wpack_offset(1), p->layer[1].b, s1.scale,
                          s1.shift, wp + wpack_offset(2), p->layer[2].b, p->layer[2].gamma, tr ? 1 : 0, F(B_STATS),
                          l1s ? l1s + L1S_H2 : nullptr, l1s ? l1s + L1S_S2 : nullptr, F(B_PCAT) + 3 * R3, tr ? U(B_ARG3) : nullptr, R3,
                          st));
        {
            Slot s2 = bn_slot(bufs, 2);
            const facl_layer& L = p->layer[2];
            RUN(bn_finalize_launch(F(B_STATS), grid, 256, (double)R1, L.gamma, L.beta, L.running_mean, L.running_var, BN_EPS, BN_MOM,
                                   tr, s2.mean, s2.rstd, s2.scale, s2.shift, st));
        }
    } else {
    // ---- L1: 4 -> 64 -> 64 -> 256 over all M*S*K grouped rows, max over the K neighbours -------------------
    {
        GemmParams g = gemm_base(64, (int)R1, 4, layer_nsplit(ns, 0));
        g.tag = 0;
        set_packed_a(g, wp + wpack_offset(0), 4);
        g.b_mode = B_XT4;
        g.b = src1(xt, 4, nullptr, nullptr, nullptr);
        g.bias = p->layer[0].b;
        g.out_mode = OUT_CHMAJOR; g.out = F(B_Z1); g.ldo = R1;
        g.stats = stats;
        RUN(launch_gemm_tc(g, st));
        RUN(finalize(0, 0, (int)R1, (double)R1));
    }
    {
        Slot s = bn_slot(bufs, 0);
        GemmParams g = gemm_base(64, (int)R1, 64, layer_nsplit(ns, 1));
        g.tag = 3;
        set_packed_a(g, wp + wpack_offset(1), 64);
        g.b = src1(F(B_Z1), R1, s.scale, s.shift, lo0);
        g.bias = p->layer[1].b;
        g.out_mode = OUT_CHMAJOR; g.out = F(B_Z2); g.ldo = R1;
        g.stats = stats;
        RUN(launch_gemm_tc(g, st));
        RUN(finalize(1, 1, (int)R1, (double)R1));
    }
    {
        Slot s = bn_slot(bufs, 1);
        GemmParams g = gemm_base(256, (int)R1, 64, layer_nsplit(ns, 2));
        g.tag = 6;
        set_packed_a(g, wp + wpack_offset(2), 64);
        g.b = src1(F(B_Z2), R1, s.scale, s.shift, lo0);
        g.bias = p->layer[2].b;
        if (tr) { g.out_mode = OUT_CHMAJOR; g.out = F(B_Z3); g.ldo = R1; }
        g.stats = stats;
        g.pool = K; g.pool_sign = p->layer[2].gamma; g.pool_out = F(B_PCAT) + 3 * R3; g.ldp = R3;
        g.pool_arg = tr ? U(B_ARG3) : nullptr;
        RUN(launch_gemm_tc(g, st));
        RUN(finalize(2, 2, (int)R1, (double)R1));
    }
    }
    // ---- L3: [centre xyz | pooled 256] -> 256 -> 512 -> 1024 over the M*S centres, max over the S centres ----
    RUN(centres_to_chmajor_launch(centres, (int)R3, F(B_PCAT), R3, st));
    const bool img = use_images(d);
    // one forward layer of net3DV_3: (image of the BN+ReLU'd input ->) GEMM with statistics (and max-pool) epilogue
    auto l3_layer = [&](int layer, int img_buf, const OperandSrc& in, float* zout, bool pool) {
        GemmParams g = gemm_base(COUT[layer], (int)R3, CIN[layer], layer_nsplit(ns, layer));
        g.tag = 3 * layer;
        set_packed_a(g, wp + wpack_offset(layer), CIN[layer]);
        if (img) {
            ActImage im = image_of(bufs, img_buf, CIN[layer], R3);
            int rc = act_image_launch(in, 0, CIN[layer], R3, nullptr, 0, layer_nhl(ns, layer), im, TAG_IMAGE, st);
            if (rc != 0) return rc;
            g.b_mode = B_IMAGE_MN;
            g.b_img = im;
        } else {
            g.b = in;
        }
        g.bias = p->layer[layer].b;
        if (!pool || tr) { g.out_mode = OUT_CHMAJOR; g.out = zout; g.ldo = R3; }
        g.stats = stats;
        if (pool) {
            g.pool = S; g.pool_sign = p->layer[layer].gamma; g.pool_out = F(B_PALL); g.ldp = MB;
            g.pool_arg = tr ? U(B_ARG6) : nullptr;
        }
        int rc = launch_gemm_tc(g, st);
        if (rc != 0) return rc;
        return finalize(layer, layer, (int)R3, (double)R3);
    };
    RUN(l3_layer(3, B_IMG_H3, src1(F(B_PCAT), R3, vec, vec + 320, vec + 640), F(B_Z4), false));
    {
        Slot s = bn_slot(bufs, 3);
        RUN(l3_layer(4, B_IMG_H4, src1(F(B_Z4), R3, s.scale, s.shift, lo0), F(B_Z5), false));
    }
    {
        Slot s = bn_slot(bufs, 4);
        RUN(l3_layer(5, B_IMG_H5, src1(F(B_Z5), R3, s.scale, s.shift, lo0), F(B_Z6), true));
    }
    // ---- sequence aggregation: max over the G views (cn3d_model_conbag.py:225-226) ---------------------------
    RUN(seq_pool_launch(F(B_PALL), MB, p->layer[5].gamma, C_FEAT, G, B, F(B_PALL) + M, MB, U(B_ARGG), st));
    return 0;
    };
    if (stage_x) RUN(trunk());
    // ---- head netR_FC on the M cloud features, then on the B sequence features (two BN batches) --------------
    for (int half = 0; half < 2; ++half) {
        if (half == 0 ? !stage_x : !stage_g) continue;
        const int Nd = half == 0 ? M : B;
        const long long off = half == 0 ? 0 : M;
        Slot s5 = bn_slot(bufs, 5);
        HeadImages hi = head_images(bufs, half, M, B);
        RUN(act_image_launch(src1(F(B_PALL) + off, MB, s5.scale, s5.shift, lo0), 0, 1024, Nd, nullptr, 0, layer_nhl(ns, 6), hi.h6,
                             TAG_IMAGE, st));
        GemmParams g = gemm_base(1024, Nd, 1024, layer_nsplit(ns, 6));
        g.tag = 18;
        set_packed_a(g, wp + wpack_offset(6), 1024);
        g.b_mode = B_IMAGE_MN; g.b_img = hi.h6;
        g.bias = p->layer[6].b;
        g.out_mode = OUT_CHMAJOR; g.out = F(B_Z7) + off; g.ldo = MB;
        g.stats = stats;
        RUN(launch_gemm_tc(g, st));
        RUN(finalize(6, 6 + half, Nd, (double)Nd));
        Slot s6 = bn_slot(bufs, 6 + half);
        RUN(act_image_launch(src1(F(B_Z7) + off, MB, s6.scale, s6.shift, lo0), 0, 1024, Nd, nullptr, 0, layer_nhl(ns, 7), hi.h7,
                             TAG_IMAGE, st));
        GemmParams h = gemm_base(C_EMB, Nd, 1024, layer_nsplit(ns, 7));
        h.tag = 21;
        set_packed_a(h, wp + wpack_offset(7), 1024);
        h.b_mode = B_IMAGE_MN; h.b_img = hi.h7;
        h.bias = p->fc3_b;
        h.out_mode = OUT_ROWMAJOR; h.out = half == 0 ? x : xg; h.ldo = C_EMB;
        RUN(launch_gemm_tc(h, st));
    }
    // ---- x_nor = normalize(x), code = mapping(x_nor) (cn3d_model_conbag.py:231-232) ---------------------------
    if (x_nor && stage_x) RUN(l2_normalize_launch(x, M, C_EMB, x_nor, st));
    if (code && stage_x) {
        if (!x_nor) return (int)cudaErrorInvalidValue;
        GemmParams g = gemm_base(C_MAP, M, C_EMB, layer_nsplit(ns, 8));
        g.tag = 24;
        set_packed_a(g, wp + wpack_offset(8), C_EMB);
        g.b_mode = B_ROWMAJOR;
        g.b = src1(x_nor, C_EMB, nullptr, nullptr, nullptr);
        g.out_mode = OUT_ROWMAJOR; g.out = code; g.ldo = C_MAP;
        RUN(launch_gemm_tc(g, st));
    }
    return 0;
}

// stages: 1 = head + net3DV_3 (every gradient except net3DV_1's is final afterwards), 2 = net3DV_1, 3 = both.  The split lets a
// data-parallel caller start the all-reduce of the large gradients while the net3DV_1 backward (43 % of the step) still runs.
// Stage 1 itself splits into 4 = the SEQUENCE half of the head (needs dxg only) and 8 = the cloud half of the head + net3DV_3
// (needs the final dx): the sharded caller runs 4 beside the reduce-scatter of the key-side gradients that completes dx.
int encoder_backward(const facl_encoder_dims* d, const facl_encoder_params* p, const float* xt, void* const* bufs, const float* dx,
                     const float* dxg, const facl_encoder_grads* gr, int stages, cudaStream_t st) {
    RUN(check_dims(d));
    if (!d->training || (stages & 15) == 0) return (int)cudaErrorInvalidValue;
    const bool head_g = (stages & (1 | 4)) != 0, head_x = (stages & (1 | 8)) != 0, stage_hi = head_x, stage_l1 = (stages & 2) != 0;
    // The FACL_ENC_SPLIT_LAYER mask of the bf16 mixed mode protects the FORWARD: an operand-rounding error made in net3DV_1 or the 259-wide
    // layer is amplified ~12x by the BatchNorms / max-pools downstream and lands in the embeddings.  The backward has no such
    // amplifier -- every weight gradient is a sum over 10^5..10^7 rows in which independent roundings average out -- so in bf16
    // mode net3DV_1 layers 1-2 (passes C / D, 64 % of the backward FLOPs) run single bf16 products regardless of the mask (measured
    // stage-wise against the fp64 oracle: every gradient within 1e-2, tests).
    // (The 259-wide layer keeps its split products in the backward as well when the mask names it: the BatchNorm-3 beta gradient is a
    // sum over all rows of a data gradient whose unmasked sum is exactly zero, and single products there left a 2e-5-of-the-largest-
    // gradient residue on it.)
    const int bwd_mask_drop = (d->flags & FACL_ENC_SPLIT_BACKWARD) ? 0 : (((1 << 1) | (1 << 2)) << 8);
    const int M = d->M, S = d->S, K = d->K, G = d->G, B = M / G, ns = mode_of(d) & ~bwd_mask_drop;
    const long long R3 = (long long)M * S, R1 = R3 * K, MB = M + B;
    auto F = [&](int i) { return reinterpret_cast<float*>(bufs[i]); };
    auto U = [&](int i) { return reinterpret_cast<unsigned char*>(bufs[i]); };
    float* vec = F(B_VEC);
    float* lo0 = vec + 960;
    float* stats = F(B_STATS);
    uint8_t* wt = reinterpret_cast<uint8_t*>(bufs[B_WPACKT]);

    if (head_g) {
    // (the transposed weight images for the data-gradient GEMMs were packed by the training-mode forward)
    // weight gradients are accumulated with atomics (split-K): start from zero; biases in front of a BN get 0
    for (int l = 0; l < 7; ++l) {
        FACL_CHECK(cudaMemsetAsync(gr->dw[l], 0, sizeof(float) * COUT[l] * CIN[l], st));
        FACL_CHECK(cudaMemsetAsync(gr->db[l], 0, sizeof(float) * COUT[l], st));
    }
    FACL_CHECK(cudaMemsetAsync(gr->dfc3_w, 0, sizeof(float) * C_EMB * C_FEAT, st));

    // upstream gradients -> channel-major [512][M | B]; the sequence columns first
    FACL_CHECK(cudaMemsetAsync(F(B_DXT), 0, sizeof(float) * C_EMB * MB, st));
    if (dxg) RUN(transpose_launch(dxg, C_EMB, F(B_DXT) + M, MB, B, C_EMB, st));
    }
    if (head_x) {
    if (dx) RUN(transpose_launch(dx, C_EMB, F(B_DXT), MB, M, C_EMB, st));
    RUN(rowstats_launch(F(B_DXT), nullptr, MB, C_EMB, (int)MB, 0, gr->dfc3_b, st));   // netR_FC.3.bias: sum over both batches
    }
    auto wgrad = [&](int layer, int Md, int Nd, int Kd, const OperandSrc& a, const OperandSrc& b, float* out) {
        GemmParams g = gemm_base(Md, Nd, Kd, layer_nsplit(ns, layer));
        g.tag = 3 * layer + 1;
        g.a_mode = A_ROWMAJOR; g.a = a;
        g.b_mode = B_ROWMAJOR; g.b = b;
        g.ksplit = pick_ksplit(Md, Nd, Kd);
        g.out_mode = OUT_ATOMIC_CHMAJOR; g.out = out; g.ldo = Nd;
        return launch_gemm_tc(g, st);
    };
    // dgrad: D[ci][r] = sum_co W^T[ci][co] * dz[co][r], masked by the ReLU of the layer below, + BN-backward sums
    auto dgrad = [&](int layer, int Md, int Nd, int Kd, const void* wimg, const OperandSrc& b, const float* zin, long long ldz,
                     const float* zs0, const float* zs2, float* out, long long ldo, bool want_stats) {
        GemmParams g = gemm_base(Md, Nd, Kd, layer_nsplit(ns, layer));
        g.tag = 3 * layer + 2;
        set_packed_a(g, wimg, Kd);
        g.b_mode = B_CHMAJOR; g.b = b;
        g.zin = zin; g.ldz = ldz; g.zs0 = zs0; g.zs2 = zs2;
        g.out_mode = OUT_CHMAJOR; g.out = out; g.ldo = ldo;
        g.stats = want_stats ? stats : nullptr;
        return launch_gemm_tc(g, st);
    };
    auto bwd_finalize = [&](int layer, int slot_i, int Md, int Nd, double n, int P, int accumulate) {
        Slot s = bn_slot(bufs, slot_i);
        (void)Md; (void)Nd;
        return bn_bwd_finalize_launch(stats, P, COUT[layer], n, p->layer[layer].gamma, s.mean, s.rstd, gr->dgamma[layer],
                                      gr->dbeta[layer], accumulate, s.c0, s.c1, s.c2, st);
    };

    // ---- head -------------------------------------------------------------------------------------------------
    Slot s5 = bn_slot(bufs, 5);
    // the sequence half (1) first: it does not need dx; the BatchNorm gradients of netR_FC.1 accumulate over the two halves
    for (int half = 1; half >= 0; --half) {
        if (half == 1 ? !head_g : !head_x) continue;
        const int Nd = half == 0 ? M : B;
        const long long off = half == 0 ? 0 : M;
        Slot s6 = bn_slot(bufs, 6 + half);
        HeadImages hi = head_images(bufs, half, M, B);
        auto img_wgrad = [&](int layer, int Md, int Nd_, const ActImage& a, const ActImage& b, float* out) {
            GemmParams g = gemm_base(Md, Nd_, Nd, layer_nsplit(ns, layer));
            g.tag = 3 * layer + 1;
            g.a_mode = A_IMAGE; g.a_img = a;
            g.b_mode = B_IMAGE_K; g.b_img = b;
            g.ksplit = pick_ksplit(Md, Nd_, Nd);
            g.out_mode = OUT_ATOMIC_CHMAJOR; g.out = out; g.ldo = Nd_;
            return launch_gemm_tc(g, st);
        };
        auto img_dgrad = [&](int layer, int Md, int Kd, const void* wimg, const ActImage& b, const float* zin, const float* zs0,
                             const float* zs2, float* out, bool want_stats) {
            GemmParams g = gemm_base(Md, Nd, Kd, layer_nsplit(ns, layer));
            g.tag = 3 * layer + 2;
            set_packed_a(g, wimg, Kd);
            g.b_mode = B_IMAGE_MN; g.b_img = b;
            g.zin = zin; g.ldz = MB; g.zs0 = zs0; g.zs2 = zs2;
            g.out_mode = OUT_CHMAJOR; g.out = out; g.ldo = MB;
            g.stats = want_stats ? stats : nullptr;
            return launch_gemm_tc(g, st);
        };
        // netR_FC.3: dW = dx^T-major * relu(bn(z7))  (its input image is still there from the forward)
        RUN(act_image_launch(src1(F(B_DXT) + off, MB, nullptr, nullptr, nullptr), 0, C_EMB, Nd, nullptr, 0, layer_nhl(ns, 7), hi.dx,
                             TAG_IMAGE, st));
        RUN(img_wgrad(7, C_EMB, C_FEAT, hi.dx, hi.h7, gr->dfc3_w));
        // grad wrt relu(bn(z7)), masked -> dh7, sums for BN(netR_FC.1)
        RUN(img_dgrad(7, C_FEAT, C_EMB, wt + wpackt_offset(7), hi.dx, F(B_Z7) + off, s6.scale, s6.shift, F(B_DH7) + off, true));
        RUN(bwd_finalize(6, 6 + half, C_FEAT, Nd, (double)Nd, gemm_tc_ctas_per_mtile(C_FEAT, Nd), half == 0 ? 1 : 0));
        // netR_FC.0: dW = dz7 * relu(bn6(pooled))^T ; data grad masked by the pooled feature's ReLU
        RUN(act_image_launch(src2(F(B_DH7) + off, F(B_Z7) + off, MB, s6.c0, s6.c1, s6.c2), 0, C_FEAT, Nd, nullptr, 0, layer_nhl(ns, 6),
                             hi.dz, TAG_IMAGE, st));
        RUN(img_wgrad(6, C_FEAT, C_FEAT, hi.dz, hi.h6, gr->dw[6]));
        RUN(img_dgrad(6, C_FEAT, C_FEAT, wt + wpackt_offset(6), hi.dz, F(B_PALL) + off, s5.scale, s5.shift, F(B_DF) + off, false));
    }
    // sequence max -> the winning view's cloud; then the BN6 sums live on the pooled positions only
    if (stage_hi) {
        RUN(combine_pool_grads_launch(F(B_DF), MB, F(B_DF) + M, MB, U(B_ARGG), C_FEAT, G, B, st));
        RUN(rowstats_launch(F(B_DF), F(B_PALL), MB, C_FEAT, M, 1, stats, st));
        RUN(bwd_finalize(5, 5, C_FEAT, M, (double)R3, 1, 0));
    }

    // ---- L3 ---------------------------------------------------------------------------------------------------
    Slot s4 = bn_slot(bufs, 4), s3 = bn_slot(bufs, 3), s2 = bn_slot(bufs, 2), s1 = bn_slot(bufs, 1), s0 = bn_slot(bufs, 0);
    if (!stage_hi) {
        // net3DV_3 and the head ran in an earlier call on these buffers
    } else if (use_images(d)) {
        // dz_l = c0*dy + c1*z + c2 is converted ONCE into an image and read by both the weight- and the data-gradient GEMM;
        // the forward input images (img_h*) are still there.  dy of layer 5 is the max-pool scatter of df, read through arg6.
        auto l3_bwd = [&](int layer, const OperandSrc& dzsrc, long long ld1, const unsigned char* parg, int h_buf, const float* zin,
                          const float* zs0, const float* zs2, float* dout, int slot_below, double n_below) {
            const int Co = COUT[layer], Ci = CIN[layer];
            ActImage dz = image_of(bufs, B_IMG_DZ, Co, R3);
            int rc = act_image_launch(dzsrc, ld1, Co, R3, parg, parg ? S : 0, layer_nhl(ns, layer), dz, TAG_IMAGE, st);
            if (rc != 0) return rc;
            GemmParams g = gemm_base(Co, Ci, (int)R3, layer_nsplit(ns, layer));
            g.tag = 3 * layer + 1;
            g.a_mode = A_IMAGE; g.a_img = dz;
            g.b_mode = B_IMAGE_K; g.b_img = image_of(bufs, h_buf, Ci, R3);
            g.ksplit = pick_ksplit(Co, Ci, (int)R3);
            g.out_mode = OUT_ATOMIC_CHMAJOR; g.out = gr->dw[layer]; g.ldo = Ci;
            rc = launch_gemm_tc(g, st);
            if (rc != 0) return rc;
            GemmParams h = gemm_base(tin(layer), (int)R3, Co, layer_nsplit(ns, layer));
            h.tag = 3 * layer + 2;
            set_packed_a(h, wt + wpackt_offset(layer), Co);
            h.b_mode = B_IMAGE_MN; h.b_img = dz;
            h.zin = zin; h.ldz = R3; h.zs0 = zs0; h.zs2 = zs2;
            h.out_mode = OUT_CHMAJOR; h.out = dout; h.ldo = R3;
            h.stats = stats;
            rc = launch_gemm_tc(h, st);
            if (rc != 0) return rc;
            return bwd_finalize(slot_below, slot_below, 0, 0, n_below, gemm_tc_ctas_per_mtile(tin(layer), (int)R3), 0);
        };
        RUN(l3_bwd(5, src2(F(B_DF), F(B_Z6), MB, s5.c0, s5.c1, s5.c2), R3, U(B_ARG6), B_IMG_H5, F(B_Z5), s4.scale, s4.shift, F(B_DH5),
                   4, (double)R3));
        RUN(l3_bwd(4, src2(F(B_DH5), F(B_Z5), R3, s4.c0, s4.c1, s4.c2), 0, nullptr, B_IMG_H4, F(B_Z4), s3.scale, s3.shift, F(B_DH4), 3,
                   (double)R3));
        RUN(l3_bwd(3, src2(F(B_DH4), F(B_Z4), R3, s3.c0, s3.c1, s3.c2), 0, nullptr, B_IMG_H3, F(B_PCAT) + 3 * R3, s2.scale, s2.shift,
                   F(B_DP3), 2, (double)R1));
    } else {
    {
        ScopedTimer timer(TAG_MEMSET, st);
        FACL_CHECK(cudaMemsetAsync(F(B_DY6), 0, sizeof(float) * 1024 * R3, st));
    }
    RUN(pool_scatter_launch(F(B_DF), MB, U(B_ARG6), MB, C_FEAT, M, S, F(B_DY6), R3, st));
    {   // layer 5 (512 -> 1024)
        OperandSrc dz = src2(F(B_DY6), F(B_Z6), R3, s5.c0, s5.c1, s5.c2);
        RUN(wgrad(5, 1024, 512, (int)R3, dz, src1(F(B_Z5), R3, s4.scale, s4.shift, lo0), gr->dw[5]));
        RUN(dgrad(5, 512, (int)R3, 1024, wt + wpackt_offset(5), dz, F(B_Z5), R3, s4.scale, s4.shift, F(B_DH5), R3, true));
        RUN(bwd_finalize(4, 4, 512, (int)R3, (double)R3, gemm_tc_ctas_per_mtile(512, (int)R3), 0));
    }
    {   // layer 4 (256 -> 512)
        OperandSrc dz = src2(F(B_DH5), F(B_Z5), R3, s4.c0, s4.c1, s4.c2);
        RUN(wgrad(4, 512, 256, (int)R3, dz, src1(F(B_Z4), R3, s3.scale, s3.shift, lo0), gr->dw[4]));
        RUN(dgrad(4, 256, (int)R3, 512, wt + wpackt_offset(4), dz, F(B_Z4), R3, s3.scale, s3.shift, F(B_DH4), R3, true));
        RUN(bwd_finalize(3, 3, 256, (int)R3, (double)R3, gemm_tc_ctas_per_mtile(256, (int)R3), 0));
    }
    {   // layer 3 (259 -> 256): input = [centre xyz | relu(bn3(pooled))]; no gradient flows to the xyz channels
        OperandSrc dz = src2(F(B_DH4), F(B_Z4), R3, s3.c0, s3.c1, s3.c2);
        RUN(wgrad(3, 256, 259, (int)R3, dz, src1(F(B_PCAT), R3, vec, vec + 320, vec + 640), gr->dw[3]));
        RUN(dgrad(3, 256, (int)R3, 256, wt + wpackt_offset(3), dz, F(B_PCAT) + 3 * R3, R3, s2.scale, s2.shift, F(B_DP3), R3, true));
        RUN(bwd_finalize(2, 2, 256, (int)R3, (double)R1, gemm_tc_ctas_per_mtile(256, (int)R3), 0));
    }
    }
    if (!stage_l1) return 0;
    if (d->flags & FACL_ENC_FUSED_L1) {
        // ---- L1 fused backward: activations recomputed from the 16-byte input rows (l1_fused.cu) ------------------------
        if ((R1 % 64) != 0 || (K != 64 && K != 128)) return (int)cudaErrorInvalidValue;
        const uint8_t* wp = reinterpret_cast<const uint8_t*>(bufs[B_WPACK]);   // forward images of this step's weights
        const double* mom = reinterpret_cast<const double*>(vec + 1984);
        const facl_layer &L0 = p->layer[0], &L1 = p->layer[1], &L2 = p->layer[2];
        const int P = 2 * l1_bwd_grid(R1);
        float* acc = F(B_L1S);
        uint8_t* imgs = reinterpret_cast<uint8_t*>(bufs[B_L1S]) + L1S_IMG_OFF;   // P3 | P2 | diag(e0) W2
        // H2 / s2 (acc + L1S_H2 .. L1S_H1) were accumulated by pass B of the forward on these buffers
        FACL_CHECK(cudaMemsetAsync(acc + L1S_H1, 0, sizeof(float) * (L1S_ACC_END - L1S_H1), st));
        RUN(l1_prep_launch(L2.w, 256, s2.c1, L2.b, s2.c2, nullptr, imgs, acc + L1S_Q3, nullptr, st));
        RUN(l1_bwd_c_launch(xt, R1, l1_nsplit(ns), L0.w, L0.b, s0.scale, s0.shift, wp + wpack_offset(1), L1.b, s1.scale, s1.shift,
                            wp + wpack_offset(2), imgs, acc + L1S_Q3, U(B_ARG3), F(B_DP3), R3, s2.c0, bufs[B_DH2], gr->dw[2], stats, K, st));
        RUN(l1_gamma0_fix_launch(xt, R1, l1_nsplit(ns), L0.w, L0.b, s0.scale, s0.shift, L1.w, L1.b, s1.scale, bufs[B_DH2], stats, st));
        RUN(bwd_finalize(1, 1, 64, (int)R1, (double)R1, 2 * P, 0));
        RUN(l1_fin_launch(L2.w, 256, s2.c1, L2.b, s2.c2, acc + L1S_H2, acc + L1S_S2, nullptr, nullptr, gr->dw[2], 1, st));
        RUN(l1_prep_launch(L1.w, 64, s1.c1, L1.b, s1.c2, s1.c0, imgs + 32768, acc + L1S_Q2, imgs + 65536, st));
        RUN(l1_bwd_d_launch(xt, R1, l1_nsplit(ns), L0.w, L0.b, s0.scale, s0.shift, imgs + 65536, imgs + 32768, acc + L1S_Q2, bufs[B_DH2],
                            acc + L1S_DW2S, acc + L1S_H1, acc + L1S_S1, F(B_DH1), stats, st));
        RUN(bwd_finalize(0, 0, 64, (int)R1, (double)R1, 2 * P, 0));
        RUN(l1_fin_launch(L1.w, 64, s1.c1, L1.b, s1.c2, acc + L1S_H1, acc + L1S_S1, s1.c0, acc + L1S_DW2S, gr->dw[1], 0, st));
        RUN(l1_dw1_launch(F(B_DH1), 2 * P, mom, L0.w, L0.b, s0.c0, s0.c1, s0.c2, gr->dw[0], st));
        return 0;
    }
    // ---- L1 ---------------------------------------------------------------------------------------------------
    {
        ScopedTimer timer(TAG_MEMSET, st);
        FACL_CHECK(cudaMemsetAsync(F(B_DY3), 0, sizeof(float) * 256 * R1, st));
    }
    RUN(pool_scatter_launch(F(B_DP3), R3, U(B_ARG3), R3, 256, (int)R3, K, F(B_DY3), R1, st));
    {   // layer 2 (64 -> 256)
        OperandSrc dz = src2(F(B_DY3), F(B_Z3), R1, s2.c0, s2.c1, s2.c2);
        RUN(wgrad(2, 256, 64, (int)R1, dz, src1(F(B_Z2), R1, s1.scale, s1.shift, lo0), gr->dw[2]));
        RUN(dgrad(2, 64, (int)R1, 256, wt + wpackt_offset(2), dz, F(B_Z2), R1, s1.scale, s1.shift, F(B_DH2), R1, true));
        RUN(bwd_finalize(1, 1, 64, (int)R1, (double)R1, gemm_tc_ctas_per_mtile(64, (int)R1), 0));
    }
    {   // layer 1 (64 -> 64)
        OperandSrc dz = src2(F(B_DH2), F(B_Z2), R1, s1.c0, s1.c1, s1.c2);
        RUN(wgrad(1, 64, 64, (int)R1, dz, src1(F(B_Z1), R1, s0.scale, s0.shift, lo0), gr->dw[1]));
        RUN(dgrad(1, 64, (int)R1, 64, wt + wpackt_offset(1), dz, F(B_Z1), R1, s0.scale, s0.shift, F(B_DH1), R1, true));
        RUN(bwd_finalize(0, 0, 64, (int)R1, (double)R1, gemm_tc_ctas_per_mtile(64, (int)R1), 0));
    }
    {   // layer 0 (4 -> 64): only the weight gradient (the input points need none)
        RUN(transpose_launch(xt, 4, F(B_XTT), R1, (int)R1, 4, st));
        OperandSrc dz = src2(F(B_DH1), F(B_Z1), R1, s0.c0, s0.c1, s0.c2);
        RUN(wgrad(0, 64, 4, (int)R1, dz, src1(F(B_XTT), R1, nullptr, nullptr, nullptr), gr->dw[0]));
    }
    return 0;
}

}  // namespace facl

using namespace facl;

extern "C" {

int facl_encoder_num_buffers(void) { return NUM_BUFS; }
const char* facl_encoder_buffer_name(int i) { return (i >= 0 && i < NUM_BUFS) ? BUF_NAMES[i] : ""; }
size_t facl_encoder_buffer_bytes(int i, const facl_encoder_dims* dims) {
    if (i < 0 || i >= NUM_BUFS || !dims || dims->G <= 0) return 0;
    size_t b = buffer_bytes(i, dims);
    return (b + 255) & ~(size_t)255;
}
int facl_encoder_buffer_backward_only(int i) { return i >= FIRST_BWD_BUF ? 1 : 0; }

int facl_encoder_forward(const facl_encoder_dims* dims, const facl_encoder_params* params, const float* xt, const float* centres,
                         void* const* buffers, float* x, float* x_global, float* x_nor, float* code, void* stream) {
    if (!dims || !params || !xt || !centres || !buffers || !x || !x_global) return (int)cudaErrorInvalidValue;
    return encoder_forward(dims, params, xt, centres, buffers, x, x_global, x_nor, code, 3, reinterpret_cast<cudaStream_t>(stream));
}

int facl_encoder_backward(const facl_encoder_dims* dims, const facl_encoder_params* params, const float* xt, void* const* buffers,
                          const float* dx, const float* dx_global, const facl_encoder_grads* grads, void* stream) {
    if (!dims || !params || !xt || !buffers || !grads) return (int)cudaErrorInvalidValue;
    return encoder_backward(dims, params, xt, buffers, dx, dx_global, grads, 3, reinterpret_cast<cudaStream_t>(stream));
}

int facl_adam_step(const void* table_dev, int ntensors, float lr, float beta1, float beta2, float eps, int step, void* stream) {
    if (!table_dev || ntensors <= 0 || step < 1) return (int)cudaErrorInvalidValue;
    return adam_launch(table_dev, ntensors, lr, beta1, beta2, eps, step, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
