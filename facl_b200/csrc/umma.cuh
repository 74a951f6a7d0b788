// sm_100a building blocks: mbarrier, 1-D bulk TMA (cp.async.bulk), tcgen05 MMA / TMEM, UMMA descriptors.
// Hand-written inline PTX; bit layouts follow the PTX ISA "tcgen05" matrix / instruction descriptors.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace facl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
#ifdef FACL_WAIT_SLEEP_NS
        __nanosleep(FACL_WAIT_SLEEP_NS);     // polling costs issue slots and shared-memory port time that the other roles need
#endif
    }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies read smem through it)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// 1-D bulk TMA: global -> shared, completion on an mbarrier (SASS: UBLKCP)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ask L2 to fetch a range that a later bulk copy will read (no shared memory, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMEM allocation (whole warp executes these)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand stored as rows of 128 bytes (64 bf16) with the
// 128-byte swizzle: 16-byte chunk j of row r lives at chunk (j ^ (r & 7)); 8-row groups are 1024 B apart.
//   [0,14) start>>4   [16,30) LBO>>4 (unused for swizzled K-major, 1)   [32,46) SBO>>4 = 64
//   [46,48) version = 1 (Blackwell)   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Shared-memory descriptor for an MN-major operand with the 128-byte swizzle.  Canonical layout (16-byte units):
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- a 128-byte row holds 64 consecutive MN elements of ONE k; 8 consecutive
// k form a 1024-byte swizzle atom (chunk index XOR (k & 7)); the next 8 k are SBO bytes further, the next 64 MN
// elements LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of the 16-byte chunk holding MN elements [8*chunk, 8*chunk+8) of row k in such a tile
__device__ __forceinline__ uint32_t mn_sw128_offset(uint32_t k, uint32_t chunk, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (chunk >> 3) * lbo_bytes + (k >> 3) * sbo_bytes + (k & 7u) * 128u + (((chunk & 7u) ^ (k & 7u)) << 4);
}
constexpr uint32_t UMMA_B_MN_MAJOR = 1u << 16;   // instruction-descriptor bit: B operand is MN-major
constexpr uint32_t UMMA_A_MN_MAJOR = 1u << 15;

// Instruction descriptor, kind::f16 with BF16 A/B (both K-major), FP32 accumulate.
//   [4,6) D fmt = 1 (F32)   [7,10) A fmt = 1 (BF16)   [10,13) B fmt = 1 (BF16)   [15] A major = 0   [16] B major = 0
//   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A tile (128 lanes x K, two bf16 per 32-bit column, K-major) is read from tensor memory,
// so the instruction pulls only the B tile from shared memory.  One k-step (K = 16) spans 8 columns.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ----------------------------------------------------------------------------------------------
// Lean issue path.  Measured with clock64 role accounting (profiles/r2_role_profile.md): the ONE thread that issues the
// tcgen05.mma stream of a fused kernel was bound by its own scalar instructions -- ~17 SASS instructions per MMA at ~6 cycles
// each: (i) every 64-bit descriptor was rebuilt (shift, mask, or) from a byte address, (ii) under `if (lane == 0)` ptxas
// cannot prove that a single thread is active and wraps every warp-level instruction (UTCHMMA, UTCBAR, UBLKCP) in an
// ELECT / PLOP3 / BRA.U.ANY loop over the active threads.  Here a descriptor is a (lo, hi) pair of 32-bit words built once per
// operand; a k-step or half-image offset is an immediate added to `lo` (the 14-bit address field counts 16-byte units and the
// shared window is < 256 KB, so the sum never carries out of the field), and the issuing branch is taken through elect.sync,
// after which ptxas emits the bare instruction: 1-3 SASS instructions per MMA.
// ----------------------------------------------------------------------------------------------
// true in exactly one lane of a fully converged warp (all 32 lanes must execute it)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
struct UDesc {
    uint32_t lo, hi;
};
// K-major, 128-byte rows, 128B swizzle (same bits as umma_desc_sw128)
__device__ __forceinline__ UDesc udesc_k(uint32_t smem_addr) {
    return UDesc{((smem_addr >> 4) & 0x3FFFu) | (1u << 16), 64u | (1u << 14) | (2u << 29)};
}
// MN-major, 128B swizzle (same bits as umma_desc_mn_sw128)
__device__ __forceinline__ UDesc udesc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return UDesc{((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16), ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)};
}
// `off16`: byte offset / 16 added to the operand's start address (k-step, hi/lo half, stage)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, UDesc a, uint32_t a_off16, UDesc b, uint32_t b_off16, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a.lo + a_off16), "r"(a.hi), "r"(b.lo + b_off16), "r"(b.hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, UDesc b, uint32_t b_off16, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b.lo + b_off16), "r"(b.hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Make the mbarrier track completion of all tcgen05.mma issued so far by this thread (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane = TMEM lane, reg i = column i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variant (lane = TMEM lane, reg i = column i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&u)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]),
        "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]),
        "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]),
        "r"(u[31])
        : "memory");
}
// 8-column variant (a 128 x 16 bf16 A tile: two bf16 per column)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&u)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]),
                 "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// operand tile helpers (K-major, 128-byte rows, 128B swizzle)
// ----------------------------------------------------------------------------------------------
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
    return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// split 8 fp32 values into bf16 "hi" (round-to-nearest) and bf16 "lo" (rounded residual): v ~= hi + lo, |err| <~ 2^-17 |v|
__device__ __forceinline__ void split_bf16x8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);                       // one packed convert for the pair
        const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);          // bf16 -> fp32 is a 16-bit shift
        const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
        l[i] = pack_bf16x2(r0, r1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

}  // namespace facl
