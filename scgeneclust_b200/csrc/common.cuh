// Shared helpers for libgeneclust_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "geneclust_b200.h"

namespace gc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(const char *what, const char *detail) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail);
    return 1;
}

inline int check_launch(const char *kernel) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(kernel, cudaGetErrorString(e));
    return 0;
}

#define GC_REQUIRE(cond, msg)                        \
    do {                                             \
        if (!(cond)) return ::gc::fail(__func__, msg); \
    } while (0)

#define GC_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::gc::fail(#expr, cudaGetErrorString(_e)); \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute is per DEVICE: caches of "already configured" are indexed by the current device, so one process
// (or thread) that drives several GPUs configures the kernel on each of them.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}

// CUDA-event bracket around a dominant kernel (gc_profile_enable / gc_profile_collect); no-ops when disabled.
enum ProfSlot { kProfRsvdPass = 0, kProfMiPairs = 1, kProfSlots = 2 };
void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, cudaStream_t st);

// scratch need of the SIMT cross-check pass (embed_kernels.cu); gc_rsvd_scratch_bytes covers both kernels
int64_t simt_scratch_bytes(int64_t C, int64_t G);

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Pearson residual of one count, float32 arithmetic in the order of the oracle (oracle/stages.py
// pearson_residuals): mu = r*s/T ; (x - mu) / sqrt(mu + mu*mu/theta) ; clip.
__device__ __forceinline__ float pearson_z(float x, float r, float s_over_T, float inv_theta, float clip) {
    float mu = r * s_over_T;
    float var = fmaf(mu * mu, inv_theta, mu);
    float z = (x - mu) * rsqrtf(var);
    return fminf(fmaxf(z, -clip), clip);
}

}  // namespace gc
