// K1: farthest-point sampling, one thread block per cloud.
//
// Replaces NTU_RGBD_new.farthest_point_sampling_fast / fps_sample_data
// (reference training_code/cn3D_data_set.py:675-694 and :665-672), which run in numpy on the host.
//
// Each thread keeps its points AND their running min-distance in registers (the cloud is read from HBM exactly
// once: 12*N bytes); one pick costs a register pass + a redux.sync warp argmax + one __syncthreads.
// Tie rule = np.argmax: the lowest index among equal maxima.  Distances use the reference's association
// ((dx*dx + dy*dy) + dz*dz) without FMA contraction, so picks are bit-exact against the fp32 reference.
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

template <int T, int PPT>
__global__ void __launch_bounds__(T) fps_kernel(const float* __restrict__ pts, int N, int D, const int* __restrict__ start, int m,
                                                int* __restrict__ out) {
    pdl_prologue();
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    const float* base = pts + (long long)v * N * D;
    float x[PPT], y[PPT], z[PPT], md[PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        int i = s * T + tid;
        if (i < N) {
            x[s] = __ldg(base + (long long)i * D + 0);
            y[s] = __ldg(base + (long long)i * D + 1);
            z[s] = __ldg(base + (long long)i * D + 2);
        } else {
            x[s] = y[s] = z[s] = 0.f;
        }
        md[s] = 0.f;
    }
    __shared__ unsigned red_val[2][32];
    __shared__ unsigned red_idx[2][32];
    int cur = start[v];
    cur = cur < 0 ? 0 : (cur >= N ? N - 1 : cur);          // an out-of-range seed would index past the cloud: clamped (documented in the header)
    if (tid == 0) out[(long long)v * m] = cur;
    int par = 0;
    for (int it = 1; it < m; ++it) {
        const float cx = __ldg(base + (long long)cur * D + 0);
        const float cy = __ldg(base + (long long)cur * D + 1);
        const float cz = __ldg(base + (long long)cur * D + 2);
        unsigned bv = 0u, bi = 0xFFFFFFFFu;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            int i = s * T + tid;
            float d = sqdist_ref(x[s], y[s], z[s], cx, cy, cz);
            md[s] = (it == 1) ? d : fminf(md[s], d);
            unsigned b = __float_as_uint(md[s]);          // md >= 0: the bit pattern orders like the value
            if (i < N && (bi == 0xFFFFFFFFu || b > bv)) {  // strict '>' while walking ascending indices = first maximum
                bv = b;
                bi = (unsigned)i;
            }
        }
        unsigned wv = __reduce_max_sync(0xFFFFFFFFu, bv);
        unsigned wi = __reduce_min_sync(0xFFFFFFFFu, (bv == wv) ? bi : 0xFFFFFFFFu);
        if (lane == 0) {
            red_val[par][warp] = wv;
            red_idx[par][warp] = wi;
        }
        __syncthreads();
        unsigned rv = (lane < NW) ? red_val[par][lane] : 0u;
        unsigned ri = (lane < NW) ? red_idx[par][lane] : 0xFFFFFFFFu;
        unsigned gv = __reduce_max_sync(0xFFFFFFFFu, rv);
        unsigned gi = __reduce_min_sync(0xFFFFFFFFu, (rv == gv) ? ri : 0xFFFFFFFFu);
        cur = (int)gi;
        if (tid == 0) out[(long long)v * m + it] = cur;
        par ^= 1;
    }
}

// rows [0,m) <- the picks in pick order, rows [m,N) <- the remaining points in ascending index order
// (cn3D_data_set.py:669-671: concatenate(picks, setdiff1d(arange(N), picks))[:N]).  FPS repeats an index once a cloud has fewer
// than m distinct points (the loader resamples with replacement); the unpicked rows then number N - uniq > N - m and the
// reference TRUNCATES the concatenation to N rows (`new_idx[:NUM_POINT]`, :671) -- so does the tail copy here.  Picks outside
// [0, N) cannot be reported from the device without a synchronisation: they are clamped into the cloud.
__global__ void __launch_bounds__(256) fps_reorder_kernel(const float* __restrict__ pts, int N, int D, const int* __restrict__ picks,
                                                          int m, float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ unsigned char flag[];   // N bytes, then 256 ints
    int* cnt = reinterpret_cast<int*>(flag + ((N + 15) & ~15));
    const int v = blockIdx.x, tid = threadIdx.x;
    const float* src = pts + (long long)v * N * D;
    float* dst = out + (long long)v * N * D;
    for (int i = tid; i < N; i += 256) flag[i] = 0;
    __syncthreads();
    for (int j = tid; j < m; j += 256) {
        int i = picks[(long long)v * m + j];
        i = i < 0 ? 0 : (i >= N ? N - 1 : i);
        flag[i] = 1;
    }
    __syncthreads();
    const int chunk = (N + 255) / 256;
    const int lo = tid * chunk, hi = min(N, lo + chunk);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += flag[i] ? 0 : 1;
    cnt[tid] = c;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < 256; ++t) {
            int q = cnt[t];
            cnt[t] = run;
            run += q;
        }
    }
    __syncthreads();
    int pos = m + cnt[tid];
    for (int i = lo; i < hi && pos < N; ++i) {           // pos < N: the reference's truncation when picks repeat
        if (!flag[i]) {
            for (int d = 0; d < D; ++d) dst[(long long)pos * D + d] = src[(long long)i * D + d];
            ++pos;
        }
    }
    for (int j = tid; j < m; j += 256) {
        int i = picks[(long long)v * m + j];
        i = i < 0 ? 0 : (i >= N ? N - 1 : i);
        for (int d = 0; d < D; ++d) dst[(long long)j * D + d] = src[(long long)i * D + d];
    }
}

template <int T, int PPT>
static int launch_fps(const float* pts, int V, int N, int D, const int* start, int m, int* out, cudaStream_t st) {
    ScopedTimer timer(TAG_FPS, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(fps_kernel<T, PPT>, dim3(V), dim3(T), 0, st, pts, N, D, start, m, out));
    return (int)cudaGetLastError();
}

int fps_launch(const float* pts, int V, int N, int D, const int* start, int m, int* out, cudaStream_t st) {
    if (V <= 0 || N <= 0 || m <= 0 || D < 3) return (int)cudaErrorInvalidValue;
    if (N <= 256) return launch_fps<128, 2>(pts, V, N, D, start, m, out, st);
    if (N <= 512) return launch_fps<128, 4>(pts, V, N, D, start, m, out, st);
    if (N <= 1024) return launch_fps<128, 8>(pts, V, N, D, start, m, out, st);
    if (N <= 2048) return launch_fps<256, 8>(pts, V, N, D, start, m, out, st);
    if (N <= 4096) return launch_fps<512, 8>(pts, V, N, D, start, m, out, st);
    if (N <= 8192) return launch_fps<1024, 8>(pts, V, N, D, start, m, out, st);
    if (N <= 16384) return launch_fps<1024, 16>(pts, V, N, D, start, m, out, st);
    if (N <= 32768) return launch_fps<1024, 32>(pts, V, N, D, start, m, out, st);
    return (int)cudaErrorInvalidValue;
}

int fps_reorder_launch(const float* pts, int V, int N, int D, const int* picks, int m, float* out, cudaStream_t st) {
    if (V <= 0 || N <= 0 || m <= 0 || m > N) return (int)cudaErrorInvalidValue;
    size_t smem = ((N + 15) & ~15) + 256 * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    ScopedTimer timer(TAG_FPS, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(fps_reorder_kernel, dim3(V), dim3(256), smem, st, pts, N, D, picks, m, out));
    return (int)cudaGetLastError();
}

}  // namespace facl
