// Shared host/device helpers for the facl_b200 CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FACL_CHECK(expr)                                   \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)

#define FACL_CHECK_LAUNCH()                                \
    do {                                                   \
        cudaError_t _e = cudaGetLastError();               \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)

namespace facl {

constexpr int kNumSMs = 148;

// squared distance exactly as the reference evaluates it: ((dx*dx + dy*dy) + dz*dz), no FMA contraction
// (utils_my.py:265-268 / cn3D_data_set.py:682-683)
__device__ __forceinline__ float sqdist_ref(float px, float py, float pz, float cx, float cy, float cz) {
    float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int div_up_dev(int a, int b) { return (a + b - 1) / b; }

}  // namespace facl

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A training step is ~90 kernels on one stream, a third of them a few microseconds long; with
// plain stream ordering every boundary costs the drain of the previous grid plus the launch of the next.  Every kernel of this
// library starts with pdl_prologue(): `griddepcontrol.wait` returns only when the preceding grid has COMPLETED and its memory
// is visible -- so no kernel body ever runs early -- and `griddepcontrol.launch_dependents` lets the next grid be scheduled
// (its CTAs become resident and park in their own wait) while this one is still running.  Launches go through launch_pdl(),
// which sets cudaLaunchAttributeProgrammaticStreamSerialization; FACL_PDL=0 in the environment disables it (A/B, debugging).
// Without the attribute both instructions are no-ops.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

#include <cstdlib>
#include <utility>
namespace facl {
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("FACL_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace facl
// a failed launch returns its error code from the enclosing launcher (all of them return the cudaError_t as int)
#define FACL_LAUNCH_OK(expr)                                   \
    do {                                                       \
        cudaError_t _le = (expr);                              \
        if (_le != cudaSuccess) return (int)_le;               \
    } while (0)

// One-time PER-DEVICE kernel configuration: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a property of the function on the
// current device, so a process that touches a second GPU (a model moved to cuda:1, nn.DataParallel replicas) must configure
// again there.  `need()` is true until `done()` has been called on the current device; a race between two host threads only
// repeats an idempotent attribute call.
#include <atomic>
namespace facl {
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0ull};
    int dev = 0;
    bool need() {
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) { dev = -1; return true; }   // unknown device: configure every time
        return ((mask.load(std::memory_order_acquire) >> dev) & 1ull) == 0ull;
    }
    void done() {
        if (dev >= 0) mask.fetch_or(1ull << dev, std::memory_order_release);
    }
};
}  // namespace facl
