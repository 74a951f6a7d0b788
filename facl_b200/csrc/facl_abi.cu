// C-ABI shim: the only translation unit that defines exported symbols (see include/facl_b200.h).
#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"
#include "gemm_tc.cuh"

using namespace facl;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* facl_version(void) { return "facl_b200 0.1 (sm_100a)"; }

const char* facl_error_string(int code) { return cudaGetErrorString(static_cast<cudaError_t>(code)); }

int facl_fps(const float* points, int V, int N, int D, const int* start_idx, int m, int* out_idx, void* stream) {
    return fps_launch(points, V, N, D, start_idx, m, out_idx, S(stream));
}

int facl_fps_reorder(const float* points, int V, int N, int D, const int* picks, int m, float* out, void* stream) {
    return fps_reorder_launch(points, V, N, D, picks, m, out, S(stream));
}

int facl_group_points(const float* points, int M, int N, int D, int Sc, int K, float r2, float* xt, int* idx, void* stream) {
    return group_launch(points, M, N, D, Sc, K, r2, xt, idx, S(stream));
}

int facl_l2_normalize(const float* x, int rows, int C, float* out, void* stream) {
    if (!x || !out || rows <= 0 || C <= 0) return (int)cudaErrorInvalidValue;
    return l2_normalize_launch(x, rows, C, out, S(stream));
}

int facl_softmax_xent(const float* logits, const int* labels, int rows, int C, float* loss, float* dlogits_t, float* dbias, int* hits,
                      void* stream) {
    return softmax_xent_launch(logits, labels, rows, C, loss, dlogits_t, dbias, hits, S(stream));
}

size_t facl_group_level2_scratch_bytes(int M, int N1, int S2, int K) { return group_level2_scratch_bytes(M, N1, S2, K); }

int facl_group_level2(const float* feats, int M, int C, int N1, int S2, int K, float r2, float* out, int* idx, void* scratch,
                      void* stream) {
    return group_level2_launch(feats, M, C, N1, S2, K, r2, out, idx, scratch, S(stream));
}

int facl_augment_views(const facl_augment_args* a, void* stream) {
    if (!a || !a->sources || !a->recipes || a->n_sources <= 0 || a->n_sources > AUGMENT_MAX_SOURCES || a->G <= 0 ||
        a->G > AUGMENT_MAX_VIEWS)
        return (int)cudaErrorInvalidValue;
    AugmentParams p{};
    p.B = a->B; p.G = a->G; p.N = a->N; p.n_sources = a->n_sources; p.g_major = a->g_major;
    p.sigma = a->sigma; p.clip = a->clip;
    for (int i = 0; i < a->n_sources; ++i) p.source[i] = AugmentSource{a->sources[i].rows, a->sources[i].offsets, a->sources[i].C};
    for (int g = 0; g < a->G; ++g) {
        const facl_view_recipe& r = a->recipes[g];
        p.recipe[g] = AugmentRecipe{r.source, r.channel, r.nonzero_only, r.jitter, r.mirror, r.rotate};
    }
    p.idx = a->idx; p.noise = a->noise; p.angle_u = a->angle_u;
    p.seed = a->seed; p.step = a->step;
    p.out = a->out; p.out_rows = a->out_rows;
    return augment_launch(p, a->max_rows, S(stream));
}

size_t facl_packed_weight_bytes(int rows, int cols) { return packed_weight_bytes(rows, cols); }

int facl_pack_weight(const float* src, long long stride_m, long long stride_k, int rows, int cols, void* image, void* stream) {
    return pack_weight_launch(src, stride_m, stride_k, rows, cols, image, S(stream));
}

static OperandSrc to_src(const facl_operand& o) {
    OperandSrc s;
    s.src0 = o.src0; s.src1 = o.src1; s.ld = o.ld; s.s0 = o.s0; s.s1 = o.s1; s.s2 = o.s2; s.lo = o.lo;
    return s;
}

static ActImage to_img(const facl_image& i) {
    ActImage a;
    a.hi = i.hi; a.lo = i.lo; a.cgs = i.cgs; a.rbs = i.rbs;
    return a;
}

int facl_gemm_stat_partials(int Md, int Nd) { return gemm_tc_ctas_per_mtile(Md, Nd); }

int facl_gemm_tc(const facl_gemm* d, void* stream) {
    if (!d) return (int)cudaErrorInvalidValue;
    GemmParams p{};
    p.Md = d->Md; p.Nd = d->Nd; p.Kd = d->Kd; p.nsplit = d->nsplit;
    p.a_mode = d->a_mode; p.b_mode = d->b_mode;
    p.a_packed = d->a_packed; p.a_packed_kblocks = d->a_packed_kblocks;
    p.a = to_src(d->a); p.b = to_src(d->b);
    p.ksplit = d->ksplit < 1 ? 1 : d->ksplit;
    p.bias = d->bias; p.out_mode = d->out_mode; p.out = d->out; p.ldo = d->ldo;
    p.zin = d->zin; p.ldz = d->ldz; p.zs0 = d->zs0; p.zs2 = d->zs2;
    p.stats = d->stats; p.pool = d->pool; p.pool_sign = d->pool_sign; p.pool_out = d->pool_out;
    p.pool_arg = d->pool_arg; p.ldp = d->ldp;
    p.tag = -1;
    p.a_img = to_img(d->a_img);
    p.b_img = to_img(d->b_img);
    return launch_gemm_tc(p, S(stream));
}

size_t facl_act_image_half_bytes(int C, long long R) { return act_image_half_bytes(C, R); }

int facl_act_image(const facl_operand* src, long long ld1, int C, long long R, const unsigned char* pool_arg, int pool, int nsplit,
                   const facl_image* img, void* stream) {
    if (!src || !img || (nsplit != 1 && nsplit != 3)) return (int)cudaErrorInvalidValue;
    return act_image_launch(to_src(*src), ld1, C, R, pool_arg, pool, nsplit == 3 ? 2 : 1, to_img(*img), -1, S(stream));
}

}  // extern "C"
