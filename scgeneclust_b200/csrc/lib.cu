// Library-wide state and the entry points that are not kernels.
#include "common.cuh"

namespace gc {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace gc

extern "C" int gc_rsvd_apply_simt(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q,
                                  void *partials, void *stream);

extern "C" int gc_abi_version(void) { return GC_ABI_VERSION; }
extern "C" const char *gc_last_error(void) { return gc::g_err; }
extern "C" long long gc_launch_count(void) { return gc::g_launches.load(); }

extern "C" int gc_rsvd_apply(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *partials,
                             void *stream) {
    return gc_rsvd_apply_simt(M, trans, Qin, Y, q, partials, stream);
}
