// Microbenchmark (developer tool): issue/execution rate of tcgen05.mma kind::f16 (bf16, M = 128, K = 16) as a function
// of N and of whether consecutive MMAs share an accumulator.  One CTA per SM, one issuing thread, operands = zeros in
// shared memory (K-major A, MN-major B, SWIZZLE_128B).  Prints cycles per MMA.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// executed by the WHOLE warp (uniform control flow); one elected lane issues
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, 1, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(32, 1) rate(int n, int rotate, int reps, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 32) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    {
        const uint64_t da = desc(base, 16, 1024), db = desc(base + 32768, 8192, 1024);
        const uint32_t step = rotate ? n : 0;     // N <= 128: 4 accumulators fit in 512 columns
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
            mma(tmem, da, db, idesc);
            mma(tmem + step, da + 2, db + 128, idesc);
            mma(tmem + 2 * step, da + 4, db + 256, idesc);
            mma(tmem + 3 * step, da + 6, db + 384, idesc);
        }
        long long t1 = clock64();
        if (threadIdx.x == 0)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        while (!try_wait(smem_u32(&bar), 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {32, 64, 128})
        for (int rot : {0, 1})
            for (int grid : {1, 148}) {
                const int reps = 4096;
                rate<<<grid, 32, 50 * 1024>>>(n, rot, reps, d);
                cudaDeviceSynchronize();
                rate<<<grid, 32, 50 * 1024>>>(n, rot, reps, d);
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N=%3d rotate=%d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)  %s\n", n, rot, grid, (double)h[0] / reps,
                       (double)h[1] / reps, 128.0 * n * 16 / 4096.0, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
