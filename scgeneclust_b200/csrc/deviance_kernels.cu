// Binomial deviance of the singleton-cluster genes (GeneClust-fast post-hoc filtering), on the device:
//   compute_deviance(X) of scGeneClust/tl/selection.py:82-89 applied to the float32 residual columns X (C x m):
//     pi = X.sum(0) / X.sum() ; n = X.sum(1)[:, None]
//     left  = nansum(X * log(X / (n pi^T)), axis=0) ; right = nansum((n - X) * log((n - X) / (n (1 - pi)^T)), axis=0)
//     deviance = 2 * (left + right)
// The host version costs ~0.1 us per matrix entry (17 ms at 200 000 cells) and needs the columns on the host.
// Every float32 operation and every summation ORDER is numpy's (checked against numpy in tests/test_gpu_e2e.py):
//   * X.sum() / a single column's sum : pairwise summation over the contiguous data (loops_utils.h.src)
//   * X.sum(1)                         : pairwise summation per row
//   * sums over axis 0 with m > 1      : row after row, out[j] += a[i, j] (sequential float32 adds)
// so the result is bit-identical up to the last-bit behaviour of logf (numpy's SIMD logf vs CUDA's, both < 1 ulp).
// Compiled with -fmad=false.
#include "common.cuh"
#include "np_sum.cuh"

namespace gc {

constexpr int kDevDepth = 10;             // 1024 subtrees of the pairwise recursion are summed in parallel

// Subtree k (bits of k, most significant first = left/right turns) of numpy's pairwise recursion over n elements
__device__ __forceinline__ void pairwise_subtree(int64_t n, int depth, int k, int64_t &lo, int64_t &len) {
    lo = 0; len = n;
    for (int level = depth - 1; level >= 0; --level) {
        int64_t n2 = len / 2;
        n2 -= n2 % 8;
        if ((k >> level) & 1) { lo += n2; len -= n2; } else { len = n2; }
    }
}

// depth such that every node above the leaves of the parallel split still has more than 128 elements
__host__ __device__ inline int pairwise_depth(int64_t n) {
    int d = 0;
    while (d < kDevDepth && (n >> (d + 1)) > 160) ++d;
    return d;
}

// partial[which][k] = pairwise sum of subtree k of array `which` (stride-1 float data of length n)
__global__ void pairwise_partial_kernel(const float *__restrict__ a0, const float *__restrict__ a1, const float *__restrict__ a2,
                                        int64_t n, int depth, float *__restrict__ partial) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const float *a = blockIdx.y == 0 ? a0 : (blockIdx.y == 1 ? a1 : a2);
    if (k >= (1 << depth) || a == nullptr) return;
    int64_t lo, len;
    pairwise_subtree(n, depth, k, lo, len);
    auto get = [&](int64_t i) -> float { return a[i]; };
    partial[(int64_t)blockIdx.y * (1 << kDevDepth) + k] = np_pairwise_sum<float>(get, lo, len);
}
// combine the subtree sums in the order of the recursion: result = pairwise(left) + pairwise(right)
__global__ void pairwise_combine_kernel(float *__restrict__ partial, int depth, float *__restrict__ out) {
    float *p = partial + (int64_t)blockIdx.x * (1 << kDevDepth);
    for (int level = depth; level > 0; --level) {
        const int half = 1 << (level - 1);
        float v = 0.f;
        if ((int)threadIdx.x < half) v = p[2 * threadIdx.x] + p[2 * threadIdx.x + 1];
        __syncthreads();
        if ((int)threadIdx.x < half) p[threadIdx.x] = v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = p[0];
}

// out[j] = a[0, j] + a[1, j] + ... sequentially (numpy add.reduce over the outer axis of a C-ordered (C, m) array, m > 1).
// The C-long dependent float32 add chain per column is the cost (4 cycles per row is the floor); everything else is
// kept off it: one CTA per block of columns (all of them when m <= 256, else 64: longer stages, more CTAs), whole row
// blocks stream into a 3-stage shared-memory ring with
// cp.async two stages ahead (16-byte copies when the block is the full row width, i.e. m <= 256, and the array is
// 16-byte aligned), thread j adds column j of the stage row after row.
constexpr int kColStages = 3, kColStageFloats = 4096, kColThreads = 256;

__device__ __forceinline__ void cp_async4(float *dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kColThreads)
colsum_sequential_kernel(const float *__restrict__ a0, const float *__restrict__ a1, int64_t C, int m, int cb,
                         float *__restrict__ out0, float *__restrict__ out1) {
    // blockIdx.y selects the array: the left and right halves of the deviance are summed by concurrent CTAs
    const float *__restrict__ a = blockIdx.y == 0 ? a0 : a1;
    float *__restrict__ out = blockIdx.y == 0 ? out0 : out1;
    __shared__ __align__(16) float ring[kColStages][kColStageFloats];
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * cb, mc = min(cb, m - j0);            // cb <= kColThreads columns per CTA
    const bool flat = mc == m;                                       // the block is whole rows: one contiguous range per stage
    const bool vec = flat && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    int R = max(1, kColStageFloats / mc);                            // rows per stage
    if (R >= 8) R -= R % 8;                                          // stage starts stay 16-byte aligned (R m floats)
    else if (R >= 4) R = 4;
    const int64_t n_stage = ceil_div<int64_t>(C, R);

    auto issue = [&](int64_t s) {
        if (s < n_stage) {
            const int64_t i0 = s * R;
            const int n = (int)min((int64_t)R, C - i0) * mc;
            float *dst = ring[s % kColStages];
            if (flat) {
                const float *src = a + i0 * m;
                int done = 0;
                if (vec) {
                    done = n & ~3;
                    for (int e = 4 * tid; e < done; e += 4 * kColThreads) cp_async16(dst + e, src + e);
                }
                for (int e = done + tid; e < n; e += kColThreads) cp_async4(dst + e, src + e);
            } else {
                for (int e = tid; e < n; e += kColThreads) {
                    const int r = e / mc, c = e - r * mc;
                    cp_async4(dst + e, a + (i0 + r) * m + j0 + c);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");          // one group per stage, empty past the end
    };

    for (int s = 0; s < kColStages - 1; ++s) issue(s);
    float acc = 0.f;
    for (int64_t s = 0; s < n_stage; ++s) {
        issue(s + kColStages - 1);                                    // into the slot the previous iteration consumed
        asm volatile("cp.async.wait_group %0;" ::"n"(kColStages - 1) : "memory");
        __syncthreads();                                              // stage s of every thread has landed
        if (tid < mc) {
            const float *p = ring[s % kColStages] + tid;
            const int rows = (int)min((int64_t)R, C - s * R);
            int r = 0;
            if (s == 0) { acc = p[0]; r = 1; }                        // the reduction starts from the first row
            // blocks of 8 rows in two register sets: the shared-memory loads of the next block are issued (unconditionally,
            // row index clamped) before the adds of the current one, so only the adds are on the dependent chain
            const int nb = (rows - r) >> 3, last = rows - 1;
            float c[8], nx[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = p[min(r + k, last) * mc];
            int b = 0;
#pragma unroll 1
            for (; b + 1 < nb; b += 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) nx[k] = p[min(r + 8 + k, last) * mc];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = acc + c[k];
#pragma unroll
                for (int k = 0; k < 8; ++k) c[k] = p[min(r + 16 + k, last) * mc];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = acc + nx[k];
                r += 16;
            }
            if (b < nb) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = acc + c[k];
                r += 8;
            }
            for (; r < rows; ++r) acc = acc + p[r * mc];
        }
        __syncthreads();                                              // slot free for the stage issued next iteration
    }
    if (tid < mc) out[j0 + tid] = acc;
}

// per row: n_i = pairwise sum of the row; L[i,j], R[i,j] with NaN -> 0 (np.nansum's _replace_nan)
__global__ void deviance_terms_kernel(const float *__restrict__ X, int64_t C, int m, const float *__restrict__ colsum,
                                      const float *__restrict__ total, float *__restrict__ L, float *__restrict__ R) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C) return;
    const float *x = X + i * m;
    auto get = [&](int64_t j) -> float { return x[j]; };
    const float n = np_pairwise_sum<float>(get, 0, m);
    const float tot = total[0];
    for (int j = 0; j < m; ++j) {
        const float pi = colsum[j] / tot;
        const float xv = x[j];
        float l = xv * logf(xv / (n * pi));
        const float nx = n - xv;
        float r = nx * logf(nx / (n * (1.0f - pi)));
        L[i * m + j] = isnan(l) ? 0.f : l;
        R[i * m + j] = isnan(r) ? 0.f : r;
    }
}

__global__ void deviance_finish_kernel(const float *__restrict__ left, const float *__restrict__ right, int m, float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = 2.0f * (left[j] + right[j]);
}

}  // namespace gc

using namespace gc;

static int64_t dev_align(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t gc_binomial_deviance_scratch_bytes(int64_t C, int m) {
    return 2 * dev_align(C * m * 4) + dev_align(3 * (1 << kDevDepth) * 4) + 4 * dev_align((int64_t)m * 4 + 16) + 512;
}

extern "C" int gc_binomial_deviance(const float *X, int64_t C, int m, float *out, void *scratch, void *stream) {
    GC_REQUIRE(X && out && scratch, "null pointer");
    GC_REQUIRE(C >= 1 && m >= 1, "bad shape");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char *p = (char *)scratch;
    float *L = (float *)p;         p += dev_align(C * m * 4);
    float *R = (float *)p;         p += dev_align(C * m * 4);
    float *partial = (float *)p;   p += dev_align(3 * (1 << kDevDepth) * 4);
    float *colsum = (float *)p;    p += dev_align((int64_t)m * 4 + 16);
    float *total = (float *)p;     p += dev_align((int64_t)m * 4 + 16);
    float *left = (float *)p;      p += dev_align((int64_t)m * 4 + 16);
    float *right = (float *)p;
    const int64_t n_all = C * (int64_t)m;
    const int cb = m <= kColThreads ? kColThreads : 64;      // columns per CTA of the sequential column sums
    const int depth = pairwise_depth(n_all);
    const int n_sub = 1 << depth;
    // X.sum() (and, for m == 1, the column sum is the same number)
    pairwise_partial_kernel<<<dim3((unsigned)ceil_div(n_sub, 128), 1), 128, 0, st>>>(X, nullptr, nullptr, n_all, depth, partial);
    if (int rc = check_launch("pairwise_partial_kernel")) return rc;
    pairwise_combine_kernel<<<1, 512, 0, st>>>(partial, depth, total);
    if (int rc = check_launch("pairwise_combine_kernel")) return rc;
    if (m == 1) {
        GC_CUDA(cudaMemcpyAsync(colsum, total, sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        colsum_sequential_kernel<<<dim3((unsigned)ceil_div(m, cb), 1), kColThreads, 0, st>>>(X, nullptr, C, m, cb, colsum, nullptr);
        if (int rc = check_launch("colsum_sequential_kernel")) return rc;
    }
    deviance_terms_kernel<<<(unsigned)ceil_div<int64_t>(C, 128), 128, 0, st>>>(X, C, m, colsum, total, L, R);
    if (int rc = check_launch("deviance_terms_kernel")) return rc;
    if (m == 1) {
        pairwise_partial_kernel<<<dim3((unsigned)ceil_div(n_sub, 128), 2), 128, 0, st>>>(L, R, nullptr, n_all, depth, partial);
        if (int rc = check_launch("pairwise_partial_kernel")) return rc;
        pairwise_combine_kernel<<<2, 512, 0, st>>>(partial, depth, left);     // left[0], then right via the second block
        if (int rc = check_launch("pairwise_combine_kernel")) return rc;
        // the second block wrote left[1]; move it to right[0]
        GC_CUDA(cudaMemcpyAsync(right, left + 1, sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        colsum_sequential_kernel<<<dim3((unsigned)ceil_div(m, cb), 2), kColThreads, 0, st>>>(L, R, C, m, cb, left, right);
        if (int rc = check_launch("colsum_sequential_kernel")) return rc;
    }
    deviance_finish_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(left, right, m, out);
    return check_launch("deviance_finish_kernel");
}
