// Training-view augmentation on the device: one launch turns the B resident source clouds of a batch into the
// (G*B) augmented, resampled views the grouping kernel consumes -- the work of NTU_RGBD_new.__getitem__ /
// get_data_train / get_temporal_augment_data (training_code/cn3D_data_set.py:105-121, 285-350, 654-663) and of the
// permute + float cast at cn3d_train_motion_GL.py:226-228.
//
// One CTA per (sequence, view).  A view is described by a recipe: which source cloud to resample, which source
// column becomes the 4th output channel, whether only rows with a non-zero value in that column may be drawn
// (the temporal segments), and which of jitter / mirror+jitter / y-rotation to apply.  All arithmetic on xyz is
// done in f64 with the reference's intermediate roundings to f32 (reverse_transform / rotate_trans copy into f32
// arrays), without FMA contraction, so that with explicit random draws the result is bit-identical to numpy's.
// HBM-bound: per view 16 B * N gathered reads + 16 B * N writes (+ draws when they are explicit).
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {

struct Philox {
    // Philox4x32-10 (Salmon et al., SC'11): counter-based, no state to carry between launches
    static __device__ __forceinline__ uint4 run(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u;
            k.y += 0xBB67AE85u;
        }
        return c;
    }
};

__device__ __forceinline__ float u01(uint32_t r) { return ((float)r + 0.5f) * 2.3283064365386963e-10f; }   // (0,1)

__device__ __forceinline__ void normals4(uint4 r, float z[4]) {
    float a = sqrtf(-2.f * __logf(u01(r.x))), b = sqrtf(-2.f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    z[0] = a * c0; z[1] = a * s0; z[2] = b * c1; z[3] = b * s1;
}

__device__ __forceinline__ double jit(double v, double z, double sigma, double clip) {
    double j = __dmul_rn(sigma, z);
    j = fmin(fmax(j, -clip), clip);
    return __dadd_rn(j, v);
}

constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS) augment_kernel(const AugmentParams p) {
    pdl_prologue();
    extern __shared__ int nz_list[];
    __shared__ int warp_cnt[THREADS / 32];
    const int view = blockIdx.x;                 // b * G + g
    const int b = view / p.G, g = view - b * p.G;
    const AugmentRecipe rc = p.recipe[g];
    const AugmentSource sc = p.source[rc.source];
    const int row0 = sc.offsets[b], P = sc.offsets[b + 1] - row0;
    const float* __restrict__ rows = sc.rows + (long long)row0 * sc.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int count = P;
    if (rc.nonzero_only) {
        // ordered compaction of the rows whose channel is non-zero (np.where(point_temporal[:, 3] != 0), :657)
        int base = 0;
        for (int r0 = 0; r0 < P; r0 += THREADS) {
            int r = r0 + tid;
            bool keep = r < P && rows[(long long)r * sc.C + rc.channel] != 0.f;
            unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_cnt[warp] = __popc(m);
            __syncthreads();
            int before = base;
            for (int w = 0; w < warp; ++w) before += warp_cnt[w];
            if (keep) nz_list[before + __popc(m & ((1u << lane) - 1u))] = r;
            int tot = 0;
            for (int w = 0; w < THREADS / 32; ++w) tot += warp_cnt[w];
            base += tot;
            __syncthreads();
        }
        count = base;
    }

    const long long oview = p.g_major ? (long long)g * p.B + b : (long long)view;
    float4* __restrict__ out = reinterpret_cast<float4*>(p.out) + oview * p.N;
    int* __restrict__ out_rows = p.out_rows ? p.out_rows + (long long)view * p.N : nullptr;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint32_t step_lo = (uint32_t)p.step, step_hi = (uint32_t)(p.step >> 32);

    double cs = 1.0, sn = 0.0;
    if (rc.rotate) {
        double u = p.angle_u ? p.angle_u[view] : (double)u01(Philox::run(make_uint4(0u, (uint32_t)view, 3u ^ (step_hi << 8), step_lo), key).x);
        double a = __dmul_rn(__dmul_rn(__dadd_rn(u, -0.5), 3.141592653589793), 0.8);      // (rand - 0.5) * np.pi * 0.8, :737
        cs = cos(a);
        sn = sin(a);
    }

    for (int n = tid; n < p.N; n += THREADS) {
        if (count <= 0) {                         // numpy raises here (randint(0, 0)); make the failure visible
            float q = __int_as_float(0x7fc00000);
            out[n] = make_float4(q, q, q, q);
            if (out_rows) out_rows[n] = -1;
            continue;
        }
        uint4 r_idx = make_uint4(0, 0, 0, 0);
        int i;
        if (p.idx) {
            i = p.idx[(long long)view * p.N + n];
        } else {
            r_idx = Philox::run(make_uint4((uint32_t)n, (uint32_t)view, 0u ^ (step_hi << 8), step_lo), key);
            i = (int)(((unsigned long long)r_idx.x * (unsigned long long)count) >> 32);
        }
        i = min(max(i, 0), count - 1);
        int r = rc.nonzero_only ? nz_list[i] : i;
        const float* __restrict__ src = rows + (long long)r * sc.C;
        double x = (double)__ldg(src), y = (double)__ldg(src + 1), z = (double)__ldg(src + 2);
        float ch = __ldg(src + rc.channel);
        double z0[3] = {0, 0, 0}, z1[3] = {0, 0, 0};
        if (rc.jitter || rc.mirror) {
            if (p.noise) {
                const double* nz = p.noise + (((long long)view * 2) * p.N + n) * 3;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    z0[e] = nz[e];
                    z1[e] = nz[(long long)p.N * 3 + e];
                }
            } else {
                float f[4], h[4];
                normals4(Philox::run(make_uint4((uint32_t)n, (uint32_t)view, 1u ^ (step_hi << 8), step_lo), key), f);
                normals4(Philox::run(make_uint4((uint32_t)n, (uint32_t)view, 2u ^ (step_hi << 8), step_lo), key), h);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    z0[e] = (double)f[e];
                    z1[e] = (double)h[e];
                }
            }
        }
        const double sg = p.sigma, cl = p.clip;
        if (rc.jitter) {
            x = jit(x, z0[0], sg, cl);
            y = jit(y, z0[1], sg, cl);
            z = jit(z, z0[2], sg, cl);
        }
        if (rc.mirror) {                          // reverse_transform, :708-713
            x = (double)(-__double2float_rn(x));
            y = (double)__double2float_rn(y);
            z = (double)__double2float_rn(z);
            x = jit(x, z1[0], sg, cl);
            y = jit(y, z1[1], sg, cl);
            z = jit(z, z1[2], sg, cl);
        }
        if (rc.rotate) {                          // rotate_trans, :734-748: row vector times Ry
            double xf = (double)__double2float_rn(x), yf = (double)__double2float_rn(y), zf = (double)__double2float_rn(z);
            x = __dadd_rn(__dadd_rn(__dmul_rn(xf, cs), __dmul_rn(yf, 0.0)), __dmul_rn(zf, -sn));
            y = __dadd_rn(__dadd_rn(__dmul_rn(xf, 0.0), __dmul_rn(yf, 1.0)), __dmul_rn(zf, 0.0));
            z = __dadd_rn(__dadd_rn(__dmul_rn(xf, sn), __dmul_rn(yf, 0.0)), __dmul_rn(zf, cs));
        }
        out[n] = make_float4(__double2float_rn(x), __double2float_rn(y), __double2float_rn(z), ch);
        if (out_rows) out_rows[n] = r;
    }
}

}  // namespace

int augment_launch(const AugmentParams& p, int max_rows, cudaStream_t st) {
    if (p.B <= 0 || p.G <= 0 || p.G > AUGMENT_MAX_VIEWS || p.N <= 0 || !p.out) return (int)cudaErrorInvalidValue;
    if ((p.idx == nullptr) != (p.noise == nullptr) || (p.idx == nullptr) != (p.angle_u == nullptr)) return (int)cudaErrorInvalidValue;
    bool any_nz = false;
    for (int g = 0; g < p.G; ++g) {
        const AugmentRecipe& rc = p.recipe[g];
        if (rc.source < 0 || rc.source >= p.n_sources) return (int)cudaErrorInvalidValue;
        const AugmentSource& sc = p.source[rc.source];
        if (!sc.rows || !sc.offsets || sc.C < 4 || rc.channel < 0 || rc.channel >= sc.C) return (int)cudaErrorInvalidValue;
        any_nz |= rc.nonzero_only != 0;
    }
    size_t smem = any_nz ? (size_t)max_rows * sizeof(int) : 0;
    if (any_nz && max_rows <= 0) return (int)cudaErrorInvalidValue;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    ScopedTimer timer(TAG_AUGMENT, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(augment_kernel, dim3(p.B * p.G), dim3(THREADS), smem, st, p));
    return (int)cudaGetLastError();
}

}  // namespace facl
