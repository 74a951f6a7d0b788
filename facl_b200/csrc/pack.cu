// Weight packing: fp32 matrix -> bf16 (hi, lo) operand-tile images in exactly the shared-memory layout the
// tcgen05 GEMM consumes (128-row x 64-column K-major tiles, 128B swizzle), so the GEMM can stage a weight tile
// with one 1-D bulk TMA copy per half.  Rows / columns beyond (Md, Kd) are zero.
//   image[m_tile][k_block][half][128 x 128 B]      half 0 = bf16(w), half 1 = bf16(w - hi)
#include "common.cuh"
#include "facl_internal.h"
#include "umma.cuh"

namespace facl {

namespace {
__global__ void pack_weight_kernel(const float* __restrict__ src, long long sm, long long sk, int Md, int Kd, int KBp, long long tasks,
                                   uint8_t* __restrict__ img) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tasks) return;
    int j = (int)(t & 7);
    int r = (int)((t >> 3) & 127);
    long long tile = t >> 10;           // mt * KBp + kb
    int kb = (int)(tile % KBp);
    int mt = (int)(tile / KBp);
    int m = mt * 128 + r;
    int k0 = kb * 64 + j * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (m < Md && k0 + e < Kd) ? __ldg(src + (long long)m * sm + (long long)(k0 + e) * sk) : 0.f;
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    uint8_t* base = img + tile * (2ll * 16384);
    uint32_t off = sw128_offset((uint32_t)r, (uint32_t)j);
    *reinterpret_cast<uint4*>(base + off) = hi;
    *reinterpret_cast<uint4*>(base + 16384 + off) = lo;
}

// several matrices in one launch (the per-step re-packing of all encoder weights): job table passed by value
__global__ void pack_weights_batched_kernel(const PackTable tbl) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tbl.total) return;
    int ji = 0;
#pragma unroll 1
    while (ji + 1 < tbl.n && t >= tbl.job[ji + 1].task0) ++ji;
    const PackJob& jb = tbl.job[ji];
    t -= jb.task0;
    int j = (int)(t & 7);
    int r = (int)((t >> 3) & 127);
    long long tile = t >> 10;
    int kb = (int)(tile % jb.KBp);
    int mt = (int)(tile / jb.KBp);
    int m = mt * 128 + r;
    int k0 = kb * 64 + j * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
        v[e] = (m < jb.Md && k0 + e < jb.Kd) ? __ldg(jb.src + (long long)m * jb.sm + (long long)(k0 + e) * jb.sk) : 0.f;
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    uint8_t* base = jb.img + tile * (2ll * 16384);
    uint32_t off = sw128_offset((uint32_t)r, (uint32_t)j);
    *reinterpret_cast<uint4*>(base + off) = hi;
    *reinterpret_cast<uint4*>(base + 16384 + off) = lo;
}
}  // namespace

void pack_table_add(PackTable& tbl, const float* src, long long sm, long long sk, int Md, int Kd, void* image) {
    PackJob& jb = tbl.job[tbl.n++];
    jb.src = src; jb.sm = sm; jb.sk = sk; jb.Md = Md; jb.Kd = Kd; jb.KBp = (Kd + 63) / 64;
    jb.task0 = tbl.total;
    jb.img = reinterpret_cast<uint8_t*>(image);
    tbl.total += (long long)((Md + 127) / 128) * jb.KBp * 1024;
}

int pack_table_launch(const PackTable& tbl, cudaStream_t st) {
    if (tbl.n <= 0) return 0;
    ScopedTimer timer(TAG_PACK, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(pack_weights_batched_kernel, dim3(div_up(tbl.total, 256)), dim3(256), 0, st, tbl));
    return (int)cudaGetLastError();
}

size_t packed_weight_bytes(int Md, int Kd) {
    size_t numMT = (Md + 127) / 128, KBp = (Kd + 63) / 64;
    return numMT * KBp * 2 * 16384;
}

int pack_weight_launch(const float* src, long long sm, long long sk, int Md, int Kd, void* image, cudaStream_t st) {
    if (Md <= 0 || Kd <= 0) return (int)cudaErrorInvalidValue;
    int numMT = (Md + 127) / 128, KBp = (Kd + 63) / 64;
    long long tasks = (long long)numMT * KBp * 1024;
    ScopedTimer timer(TAG_PACK, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(pack_weight_kernel, dim3(div_up(tasks, 256)), dim3(256), 0, st, src, sm, sk, Md, Kd, KBp, tasks, reinterpret_cast<uint8_t*>(image)));
    return (int)cudaGetLastError();
}

}  // namespace facl
