// Library-wide state and the entry points that are not kernels.
#include "common.cuh"

#include <mutex>
#include <vector>

namespace gc {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof[kProfSlots];
static thread_local cudaEvent_t g_prof_open[kProfSlots];

void prof_begin(int slot, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t e0 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess) return;
    cudaEventRecord(e0, st);
    g_prof_open[slot] = e0;
}
void prof_end(int slot, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed) || g_prof_open[slot] == nullptr) return;
    cudaEvent_t e1 = nullptr;
    if (cudaEventCreate(&e1) != cudaSuccess) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof[slot].emplace_back(g_prof_open[slot], e1);
    g_prof_open[slot] = nullptr;
}
}  // namespace gc

extern "C" int gc_profile_enable(int on) {
    gc::g_prof_on.store(on ? 1 : 0);
    return 0;
}

extern "C" int gc_profile_collect(int slot, double *total_ms, long long *launches) {
    GC_REQUIRE(slot >= 0 && slot < gc::kProfSlots && total_ms && launches, "bad arguments");
    std::lock_guard<std::mutex> lk(gc::g_prof_mu);
    double tot = 0.0;
    long long n = 0;
    for (auto &pr : gc::g_prof[slot]) {
        GC_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        GC_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        ++n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    gc::g_prof[slot].clear();
    *total_ms = tot;
    *launches = n;
    return 0;
}

extern "C" int gc_abi_version(void) { return GC_ABI_VERSION; }
extern "C" const char *gc_last_error(void) { return gc::g_err; }
extern "C" long long gc_launch_count(void) { return gc::g_launches.load(); }
