// Optional in-library kernel timing: when enabled, tagged launch sites are bracketed by CUDA events on the launch
// stream; facl_timing_collect() synchronises and returns total ms and launch count per tag.  Also counts every
// kernel launch the library makes (facl_launch_count), enabled or not.
#include <vector>

#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {
struct Range {
    int tag;
    cudaEvent_t a, b;
};
bool g_enabled = false;
std::vector<Range> g_ranges;
std::vector<cudaEvent_t> g_pool;
long long g_launches = 0;

cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void count_launch(int n) { g_launches += n; }

ScopedTimer::ScopedTimer(int tag, cudaStream_t st) : tag_(tag), st_(st), active_(g_enabled && tag >= 0) {
    if (active_) {
        a_ = get_event();
        b_ = get_event();
        cudaEventRecord(a_, st_);
    }
}
ScopedTimer::~ScopedTimer() {
    if (active_) {
        cudaEventRecord(b_, st_);
        g_ranges.push_back(Range{tag_, a_, b_});
    }
}

}  // namespace facl

extern "C" {

void facl_timing_enable(int on) { facl::g_enabled = on != 0; }

int facl_timing_collect(float* ms_per_tag, int* count_per_tag, int ntags) {
    for (int i = 0; i < ntags; ++i) {
        ms_per_tag[i] = 0.f;
        count_per_tag[i] = 0;
    }
    for (auto& r : facl::g_ranges) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return (int)e;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (r.tag < ntags) {
            ms_per_tag[r.tag] += ms;
            count_per_tag[r.tag] += 1;
        }
        facl::g_pool.push_back(r.a);
        facl::g_pool.push_back(r.b);
    }
    facl::g_ranges.clear();
    return 0;
}

long long facl_launch_count(void) { return facl::g_launches; }

void facl_debug_l1_dump(unsigned char* mask1, unsigned char* mask2) {
    facl::l1_set_debug_dump(mask1, mask2);
}

}  // extern "C"
