// Randomized-SVD pass on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   Y[n_out x 64] = Mop[n_out x n_k] @ Q[n_k x 64],   Mop = M (trans 0, rows = cells) or M^T (trans 1, rows = genes)
//   M[c,g] = clip((x - mu)/sqrt(mu + mu^2/theta), +-clip) - row_shift[c] - col_shift[g],  mu = r_c * s_g / T
//
// replaces the BLAS passes of sklearn randomized SVD (utils/extmath.py:377-383, :606) reached from sc.pp.pca at
// scGeneClust/pp/_preprocessing.py:52,54.  The residual matrix never exists in memory: producer warps read the
// uint16 counts from HBM (the only large stream, 2 B per matrix entry per pass), form the residual in registers,
// split it into bf16 hi + bf16 lo (16 mantissa bits together) and store both into shared memory in the UMMA
// canonical SWIZZLE_128B layout; one elected thread issues three tcgen05.mma per 16-deep k slice
//   D += A_hi*Q_hi + A_hi*Q_lo + A_lo*Q_hi        (kind::f16, bf16 inputs, fp32 accumulation in TMEM)
// which carries ~2^-16 relative operand error - single-pass bf16 or tf32 cannot hold the 1e-3 PCA-score tolerance
// (measured: tf32 2e-3, bf16 O(1); hi+lo 8e-5).  The Q panel is pre-split once per pass (tiny, L2 resident).
//
// CTA = one 128-row output tile x one k-slab (split-K; slabs are reduced in a fixed order by a second kernel, so
// results are deterministic), one CTA per SM, 576 threads:
//   warps 0-15  producers (then epilogue), organised as FOUR GROUPS of four warps: group g owns the stages it = g, g+4,
//               ... completely (128 rows x 64 k = 8 chunks of 8 entries per thread).  A group pays the per-stage fixed
//               costs (proxy fence, barrier round trips) once per 64 entries per thread, and the groups drift apart in
//               time, so one group's fence / wait bubbles are filled by the other groups' math (with all sixteen warps
//               on the same stage - the previous layout - every warp hit the same bubble at the same time).
//               Counts come straight from HBM into registers, prefetched one whole group-stage (= four stages of the
//               CTA) ahead: 64 KB of reads in flight per SM.
//   warp 16     MMA issuer + TMEM owner
//   warp 17     bulk-copy issuer for the panel tiles (Qh/Ql, stored PRE-SWIZZLED per 64-row tile by the prep kernel, so
//               one 8 KB cp.async.bulk lands a tile in the UMMA SWIZZLE_128B layout)
// Four UMMA stages of 48 KB (A hi, A lo, B hi, B lo): the slot a group needs was last used by its own previous stage.
// Bound: HBM (algorithmic 2 B/entry); the producer ALU work (12 lane-instructions per entry) is the practical limiter.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace gc {

constexpr int TC_TM = 128;           // output rows per CTA (UMMA M)
constexpr int TC_TN = 64;            // panel width (UMMA N)
constexpr int TC_TK = 64;            // k per stage (one 128-byte swizzle row of bf16)
constexpr int TC_STAGES = 4;         // UMMA operand stages = producer groups
constexpr int TC_GROUPS = 4;
constexpr int TC_PRODUCER_WARPS = 16;
constexpr int TC_GROUP_WARPS = TC_PRODUCER_WARPS / TC_GROUPS;
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_THREADS = TC_PRODUCERS + 64;
constexpr int TC_CHUNKS = 1024 / (TC_GROUP_WARPS * 32);   // 16-byte count chunks per producer thread and stage
constexpr int TC_EPI_WARPS = 16;     // producer warps that also run the epilogue (4 TMEM lane groups x 4 column quarters)
constexpr int TC_A_BYTES = TC_TM * TC_TK * 2;      // 16 KB per half (hi or lo)
constexpr int TC_B_BYTES = TC_TN * TC_TK * 2;      // 8 KB per half
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB
constexpr int TC_BAR_BYTES = 256;
// TMEM: two accumulator sets (k-slice parity) x {D1: 128 columns = [A_hi Q_hi | A_hi Q_lo], D2: 64 columns = A_lo Q_hi}:
// the hi.hi and hi.lo products are one N = 128 MMA (64 cycles, full rate; N = 64 runs at 48 cycles, 67 %).
constexpr int TC_TMEM_SET = 192, TC_TMEM_COLS = 512;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + TC_BAR_BYTES;

struct TcArgs {
    const uint16_t *counts; int64_t C, G, ld;
    const float *rpad;                   // [ceil128(C)]  r_c  (1.0 beyond C)
    const float *tpad;                   // [ceil128(ld)] s_g / T (1.0 beyond G)
    float inv_theta, clip;
    uint32_t magic;                      // 0x4B000000 (2^23) as a run-time value: keeps it in a register so that PRMT
                                         // takes its selector as the immediate (ptxas otherwise re-materialises it)
    const __nv_bfloat16 *Qh, *Ql;        // pre-split panel, [ceil64(n_k)][64] each, zero rows beyond n_k
    float *partials;                     // [n_split][n_out][64]
    int64_t n_out, n_k; int n_split; int64_t k_per_split;   // multiple of TC_TK
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU - after ~seconds of spinning the kernel traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// Same, for waits that are expected to last: back off between polls so the spinning warp does not take issue slots
// from the warps doing the residual math on the same scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        __nanosleep(64);
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Same, with the 64-bit shared-memory descriptors given as (lo, hi) words - the issuer keeps the constant hi words and
// only adds byte offsets (>> 4) to the lo words - and the accumulate flag as a compile-time constant.
// MUST be executed by a whole converged warp with warp-uniform operands: one elected lane issues.  (Issued from a
// divergent `if (lane == 0)` region instead, ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop and feeds its uniform
// registers through R2UR chains - measured 150 cycles per MMA on the issuing thread against 48 (N = 64) / 64 (N = 128)
// for the tensor core itself, scripts/microbench/umma_rate.cu.)
template <bool ACCUM>
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACCUM ? 1 : 0) : "memory");
}
// whole-warp variant of umma_commit: one elected lane commits
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b_format BF16 (1)
// [7,10)/[10,13), a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int a_mn_major, int b_mn_major, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

// 16-byte chunk `chunk` (0..7) of 128-byte row `row` inside a 1024-byte-aligned SWIZZLE_128B region
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t chunk) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));     // one MUFU.RSQ
    return y;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Eight residuals of one 16-byte chunk of counts (x_e = uint16 e of `raw`), row constant r, column constants t[e]:
//   mu = r t ; z = min((x - mu) * rsqrt(mu (1 + mu/theta)), clip)            (12 lane-instructions per entry)
// The lower clip is dead code when clip >= sqrt(theta): z >= -sqrt(theta mu / (theta + mu)) > -sqrt(theta).
// Output: packed bf16 hi (truncation) and bf16 lo (rounded remainder) chunks.
template <bool LO_CLIP, bool ROW_R>
__device__ __forceinline__ void resid8(const uint4 &raw, const float (&cr)[8], float fixed, float inv_theta, float clip,
                                       uint32_t magic, uint4 &hi, uint4 &lo) {
    // ROW_R: `fixed` is the row constant r and cr[] the per-column t;  else `fixed`... (same product either way)
#ifdef TC_DBG_NOMATH      // ablation build (scripts/build_variant.py): keep the stores / MMAs, drop the residual math
    hi = raw; lo = raw; return;
#endif
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float z[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // exact uint16 -> float: splice the 16 bits under the 2^23 exponent, subtract 2^23
            const float xf = __uint_as_float(__byte_perm(w[p], magic, e == 0 ? 0x7410 : 0x7432)) - 8388608.0f;
            const float mu = fixed * cr[2 * p + e];
            const float var = mu * fmaf(mu, inv_theta, 1.0f);
            float v = (xf - mu) * rsqrt_approx(var);
            v = fminf(v, clip);
            if (LO_CLIP) v = fmaxf(v, -clip);
            z[e] = v;
        }
        const uint32_t b0 = __float_as_uint(z[0]), b1 = __float_as_uint(z[1]);
        h[p] = __byte_perm(b0, b1, 0x7632);                                  // {hi16(z0), hi16(z1)}
        const float l0 = z[0] - __uint_as_float(b0 & 0xffff0000u);
        const float l1 = z[1] - __uint_as_float(b1 & 0xffff0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(l0, l1);             // .x = l0 (low half)
        l[p] = *reinterpret_cast<const uint32_t *>(&lp);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------------------------
// per-pass preparation (one launch): the input panel Q fp32 [n_k][64] -> Qh, Ql bf16 [ceil64(n_k)][64], stored as
// pre-swizzled 64-row tiles (zero rows
// beyond n_k, so out-of-range k never contributes), block-partial column sums of Q and of k_shift[k] * Q[k,:]
// (for the rank-1 centring correction applied by the reduce kernel), and the padded constant vectors rpad / tpad.
// Block b owns PREP_ROWS consecutive rows; thread = (row lane 0..15, float4 column group 0..15).
// ---------------------------------------------------------------------------------------------------------
constexpr int PREP_ROWS = 128;
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float *__restrict__ Q, int64_t n_k, int64_t n_k_pad, const float *__restrict__ k_shift,
               __nv_bfloat16 *__restrict__ Qh, __nv_bfloat16 *__restrict__ Ql, double *__restrict__ colsum_part,
               const float *__restrict__ row_sum_f, int64_t C, int64_t C_pad, const float *__restrict__ col_sum_f,
               int64_t G, int64_t G_pad, float inv_total, float *__restrict__ rpad, float *__restrict__ tpad) {
    __shared__ double red[16][128];
    const int rl = threadIdx.x >> 4, cg = threadIdx.x & 15;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const int64_t r_base = (int64_t)blockIdx.x * PREP_ROWS;
#pragma unroll
    for (int i = 0; i < PREP_ROWS / 16; ++i) {
        const int64_t r = r_base + rl + 16 * i;
        if (r >= n_k_pad) break;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_k) q = reinterpret_cast<const float4 *>(Q)[r * 16 + cg];
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(q.x, q.y), h1 = __floats2bfloat162_rn(q.z, q.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(q.x - f0.x, q.y - f0.y);
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(q.z - f1.x, q.w - f1.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t *>(&h0); hv.y = *reinterpret_cast<const uint32_t *>(&h1);
        lv.x = *reinterpret_cast<const uint32_t *>(&l0); lv.y = *reinterpret_cast<const uint32_t *>(&l1);
        // pre-swizzled 64-row tiles: 16-byte chunk c of row rr sits where SWIZZLE_128B wants it in shared memory
        const int64_t off = (r >> 6) * 8192 + swz128((uint32_t)(r & 63), (uint32_t)(cg >> 1)) + (cg & 1) * 8;
        *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(Qh) + off) = hv;
        *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(Ql) + off) = lv;
        if (r < n_k) {
            const double sh = k_shift ? (double)k_shift[r] : 0.0;
            s1[0] += q.x; s1[1] += q.y; s1[2] += q.z; s1[3] += q.w;
            s2[0] += sh * q.x; s2[1] += sh * q.y; s2[2] += sh * q.z; s2[3] += sh * q.w;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[rl][cg * 4 + j] = s1[j]; red[rl][64 + cg * 4 + j] = s2[j]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += red[k][threadIdx.x];
        colsum_part[(int64_t)blockIdx.x * 128 + threadIdx.x] = s;
    }
    // padded constants (grid-stride)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < C_pad; i += stride)
        rpad[i] = i < C ? row_sum_f[i] : 1.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < G_pad; i += stride)
        tpad[i] = i < G ? col_sum_f[i] * inv_total : 1.0f;
}

// colsum[0..63] = sum_k Q[k,:], colsum[64..127] = sum_k k_shift[k] Q[k,:]   (fixed order => deterministic):
// one CTA per output value e; thread t adds blocks t, t + 256, ... in order, then a fixed shared-memory tree.  (A single
// CTA for all 128 values took 0.1 ms per pass at 400 000 rows.)
__global__ void __launch_bounds__(256) tc_colsum_kernel(const double *__restrict__ part, int n_blocks, float *__restrict__ colsum) {
    __shared__ double red[256];
    const int e = blockIdx.x, t = threadIdx.x;
    double s = 0.0;
    for (int b = t; b < n_blocks; b += 256) s += part[(int64_t)b * 128 + e];
    red[t] = s;
    __syncthreads();
#pragma unroll
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
    }
    if (t == 0) colsum[e] = (float)red[0];
}

// ---------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------
template <int TRANS, bool LO_CLIP>
__global__ void __launch_bounds__(TC_THREADS, 1)
rsvd_apply_tc_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
    // barriers (8 B each): full[S] | empty[S] | accum ; then the TMEM address slot
    const uint32_t bar_full = bar_base, bar_empty = bar_full + 8 * TC_STAGES, bar_accum = bar_empty + 8 * TC_STAGES;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + (bar_accum + 8 - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_TM;
    const int split = blockIdx.y;
    const int64_t k_lo = split * a.k_per_split, k_hi = min(a.n_k, k_lo + a.k_per_split);
    const int n_it = (int)ceil_div<int64_t>(max((int64_t)0, k_hi - k_lo), TC_TK);

    if (warp == TC_PRODUCER_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) {
                mbar_init(bar_full + 8 * s, TC_GROUP_WARPS + 1);      // one arrive per warp of the owning group + the panel copier
                mbar_init(bar_empty + 8 * s, 1);                      // tcgen05.commit
            }
            mbar_init(bar_accum, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32((const void *)tmem_slot), TC_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        // =============================== producers ===============================
        // group = warp / 4 owns stages it = group, group + 4, ...; tg = thread index inside the group (0..127)
        // TRANS 0: A tile = 128 cells x 64 genes, K-major.   chunk col = tg%8  (8 genes), rows  tg/8  + 16 j, j < 8
        // TRANS 1: A tile = 128 genes x 64 cells, MN-major.  chunk col = tg%16 (8 genes), cells tg/16 +  8 j, j < 8
        // No bounds predicates: row / cell indices are clamped to the matrix (their products meet zero panel rows or are
        // never stored), columns are clamped to ld - 8, rpad / tpad are 1 beyond the matrix so every residual is finite.
        const int group = warp / TC_GROUP_WARPS, tg = tid & (TC_GROUP_WARPS * 32 - 1);
        const int a_col = TRANS == 0 ? (tg & 7) : (tg & 15);
        const int a_row0 = TRANS == 0 ? (tg >> 3) : (tg >> 4);
        constexpr int A_ROW_STEP = TRANS == 0 ? TC_GROUP_WARPS * 32 / 8 : TC_GROUP_WARPS * 32 / 16;
        // shared-memory chunk offsets are affine in j (row + 16 j / cell + 8 j keeps row & 7): base + immediate
        constexpr uint32_t A_OFF_STEP = (uint32_t)(A_ROW_STEP / 8) * 1024u;
        const uint32_t a_off0 = TRANS == 0 ? swz128(a_row0, a_col)
                                           : (uint32_t)(a_col >> 3) * 8192u + swz128(a_row0, a_col & 7);
        const int32_t c_last = (int32_t)(a.C - 1);
        const int64_t ld = a.ld;

        float fix[8];                  // TRANS 0: r of the 8 cells; TRANS 1: t of the 8 genes
        const uint16_t *col_base;      // TRANS 1: counts + first gene of this thread's chunk
        if (TRANS == 0) {
#pragma unroll
            for (int j = 0; j < TC_CHUNKS; ++j) fix[j] = __ldg(a.rpad + m0 + a_row0 + A_ROW_STEP * j);   // rpad covers ceil128(C)
            col_base = a.counts;
        } else {
            const int64_t g_fixed = min(m0 + a_col * 8, ld - 8);
            const float4 t0 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g_fixed));
            const float4 t1 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g_fixed) + 1);
            fix[0] = t0.x; fix[1] = t0.y; fix[2] = t0.z; fix[3] = t0.w;
            fix[4] = t1.x; fix[5] = t1.y; fix[6] = t1.z; fix[7] = t1.w;
            col_base = a.counts + g_fixed;
        }

        uint4 raw[TC_CHUNKS];
        float cst[8];                  // TRANS 0: t of the stage's 8 genes of this thread; TRANS 1: r of its 8 cells
        auto load_chunk = [&](int it, int j) {
#ifdef TC_DBG_NOLOAD      // ablation build: only the first group-stage is read from HBM
            if (it >= TC_GROUPS) return;
#endif
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
            if constexpr (TRANS == 0) {
                const int32_t row = min((int32_t)(m0 + a_row0 + A_ROW_STEP * j), c_last);
                const int64_t g0 = min(kk + a_col * 8, ld - 8);
                raw[j] = __ldg(reinterpret_cast<const uint4 *>(col_base + (int64_t)row * ld + g0));
            } else {
                const int32_t row = min((int32_t)(kk + a_row0 + A_ROW_STEP * j), c_last);
                raw[j] = __ldg(reinterpret_cast<const uint4 *>(col_base + (int64_t)row * ld));
            }
        };
        auto load_cst = [&](int it) {
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
            if constexpr (TRANS == 0) {
                const int64_t g0 = min(kk + a_col * 8, ld - 8);
                const float4 t0 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g0));
                const float4 t1 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g0) + 1);
                cst[0] = t0.x; cst[1] = t0.y; cst[2] = t0.z; cst[3] = t0.w;
                cst[4] = t1.x; cst[5] = t1.y; cst[6] = t1.z; cst[7] = t1.w;
            } else {
#pragma unroll
                for (int j = 0; j < TC_CHUNKS; ++j) cst[j] = __ldg(a.rpad + kk + a_row0 + A_ROW_STEP * j);   // rpad covers ceil64(C)
            }
        };

        if (group < n_it) {
#pragma unroll
            for (int j = 0; j < TC_CHUNKS; ++j) load_chunk(group, j);
            load_cst(group);
        }
        for (int it = group; it < n_it; it += TC_GROUPS) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            const uint32_t st = smem_base + s * TC_STAGE_BYTES + a_off0;
            const bool more = it + TC_GROUPS < n_it;
            uint4 hi, lo;
            // chunk 0 is computed before the wait for the UMMA slot: the slot was last used by this group's own previous
            // stage, whose MMAs were issued a few hundred cycles ago
            if constexpr (TRANS == 0) resid8<LO_CLIP, true>(raw[0], cst, fix[0], a.inv_theta, a.clip, a.magic, hi, lo);
            else resid8<LO_CLIP, false>(raw[0], fix, cst[0], a.inv_theta, a.clip, a.magic, hi, lo);
            if (more) load_chunk(it + TC_GROUPS, 0);
            mbar_wait_backoff(bar_empty + 8 * s, ph ^ 1);     // passes at once on the first lap
            sts128(st, hi);
            sts128(st + TC_A_BYTES, lo);
#pragma unroll
            for (int j = 1; j < TC_CHUNKS; ++j) {
                if constexpr (TRANS == 0) resid8<LO_CLIP, true>(raw[j], cst, fix[j], a.inv_theta, a.clip, a.magic, hi, lo);
                else resid8<LO_CLIP, false>(raw[j], fix, cst[j], a.inv_theta, a.clip, a.magic, hi, lo);
                if (more) load_chunk(it + TC_GROUPS, j);      // the register of the chunk just consumed takes the next one
                sts128(st + A_OFF_STEP * j, hi);
                sts128(st + TC_A_BYTES + A_OFF_STEP * j, lo);
            }
            if (more) load_cst(it + TC_GROUPS);
            fence_proxy_async_smem();                         // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }

        // =============================== epilogue ===============================
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (= output rows), column quarter w / 4
        if (n_it > 0) {
            mbar_wait(bar_accum, 0);
            tc_fence_after();
        }
        if (warp >= TC_EPI_WARPS) goto epilogue_done;
        {
        const int lane_grp = warp & 3, col_q = warp >> 2;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        if (n_it > 0) {
            // fixed summation order: set 0 {hi.hi, hi.lo, lo.hi}, then set 1
            const uint32_t t0 = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(col_q * 16);
#pragma unroll
            for (int part = 0; part < 6; ++part) {
                uint32_t v[16];
                tmem_ld16(t0 + (part / 3) * TC_TMEM_SET + (part % 3) * 64, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
            }
        }
        const int64_t row = m0 + lane_grp * 32 + lane;
        if (row < a.n_out) {
            float4 *dst = reinterpret_cast<float4 *>(a.partials + ((int64_t)split * a.n_out + row) * TC_TN + col_q * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
        }
    epilogue_done:;
    } else if (warp == TC_PRODUCER_WARPS) {
        // =============================== MMA issuer (whole warp, one elected lane issues) =============
        {
            constexpr uint32_t idesc128 = umma_idesc(TRANS == 0 ? 0 : 1, 1, 128), idesc64 = umma_idesc(TRANS == 0 ? 0 : 1, 1, 64);
            // A: K-major (TRANS 0): SBO = 1024 (next 8 rows), K step = 32 B.
            //    MN-major (TRANS 1): LBO = 8192 (next 64 genes), SBO = 1024 (next 8 cells), K step = 2048 B.
            // B: MN-major, LBO = 8192 (next 64 panel columns: B_lo sits right behind B_hi, so ONE N = 128 descriptor covers
            //    [Q_hi | Q_lo]), SBO = 1024 (next 8 k), K step = 2048 B.
            // The issuing thread is on the critical path of every stage, so the descriptors are built once: per MMA only
            // the start-address field of the lo word changes (byte offset >> 4, no carry out of its 14 bits because
            // shared memory is < 256 KB).
            constexpr uint32_t a_lbo = TRANS == 0 ? 16u : 8192u, a_sbo = 1024u, a_kstep = TRANS == 0 ? 32u : 2048u;
            constexpr uint32_t b_lbo = 8192u, b_sbo = 1024u, b_kstep = 2048u;
            const uint64_t da0 = umma_desc(smem_base, a_lbo, a_sbo), db0 = umma_desc(smem_base + 2 * TC_A_BYTES, b_lbo, b_sbo);
            const uint32_t a_hiw = (uint32_t)(da0 >> 32), b_hiw = (uint32_t)(db0 >> 32);
            const uint32_t a_low0 = (uint32_t)da0, b_low0 = (uint32_t)db0;
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_st = a_low0 + (uint32_t)s * (TC_STAGE_BYTES >> 4), b_st = b_low0 + (uint32_t)s * (TC_STAGE_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < TC_TK / 16; ++k) {
                    const uint32_t ah = a_st + k * (a_kstep >> 4), al = ah + (TC_A_BYTES >> 4);
                    const uint32_t bh = b_st + k * (b_kstep >> 4);
                    const uint32_t d1 = tmem_d + (k & 1) * TC_TMEM_SET, d2 = d1 + 128;
                    if (k < 2) {       // first use of this accumulator set in the stage: overwrite on the first stage
                        if (it == 0) {
                            umma_bf16_w<false>(d1, ah, a_hiw, bh, b_hiw, idesc128);
                            umma_bf16_w<false>(d2, al, a_hiw, bh, b_hiw, idesc64);
                        } else {
                            umma_bf16_w<true>(d1, ah, a_hiw, bh, b_hiw, idesc128);
                            umma_bf16_w<true>(d2, al, a_hiw, bh, b_hiw, idesc64);
                        }
                    } else {
                        umma_bf16_w<true>(d1, ah, a_hiw, bh, b_hiw, idesc128);
                        umma_bf16_w<true>(d2, al, a_hiw, bh, b_hiw, idesc64);
                    }
                }
                umma_commit_elect(bar_empty + 8 * s);         // frees the stage when these MMAs have read it
            }
            if (n_it > 0) umma_commit_elect(bar_accum);       // accumulator complete
        }
        __syncwarp();
    } else {
        // =============================== panel-tile copy issuer (one thread) ===============================
        if (lane == 0) {
            const char *qh = reinterpret_cast<const char *>(a.Qh) + (k_lo / TC_TK) * TC_B_BYTES;
            const char *ql = reinterpret_cast<const char *>(a.Ql) + (k_lo / TC_TK) * TC_B_BYTES;
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait_backoff(bar_empty + 8 * s, ph ^ 1);
                const uint32_t b_hi = smem_base + s * TC_STAGE_BYTES + 2 * TC_A_BYTES;
                mbar_arrive_expect_tx(bar_full + 8 * s, 2 * TC_B_BYTES);
                bulk_g2s(b_hi, qh + (int64_t)it * TC_B_BYTES, TC_B_BYTES, bar_full + 8 * s);
                bulk_g2s(b_hi + TC_B_BYTES, ql + (int64_t)it * TC_B_BYTES, TC_B_BYTES, bar_full + 8 * s);
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

// Y[row, :] = sum_p partials[p][row, :]  -  out_shift[row] * colsum[0..63]  -  colsum[64..127]
// (fixed summation order => deterministic; the last two terms are the implicit PCA centring, a rank-1 correction).
// n_ctas == 0: every slab wrote its own partial (slots 0 .. n_split-1).  n_ctas > 0 (the persistent TS kernel): a CTA
// accumulates the consecutive slabs of a tile in registers and writes ONE partial per (tile, CTA), into the slot of the
// first slab of its segment; the segments of tile t start at split 0 and wherever a CTA's item range
// [floor(b n_items / n_ctas), floor((b+1) n_items / n_ctas)) begins inside the tile's items [t n_split, (t+1) n_split).
__global__ void rsvd_tc_reduce_kernel(const float *__restrict__ partials, int64_t n_elem4, int n_split,
                                      const float *__restrict__ out_shift, const float *__restrict__ colsum,
                                      float *__restrict__ Y, int n_ctas, int64_t n_items) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_elem4) return;
    float4 s = reinterpret_cast<const float4 *>(partials)[i];
    if (n_ctas == 0) {
        for (int p = 1; p < n_split; ++p) {
            const float4 v = reinterpret_cast<const float4 *>(partials)[(int64_t)p * n_elem4 + i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    } else {
        const int64_t tile = (i >> 4) / TC_TM;                       // 16 float4 per output row, TC_TM rows per tile
        const int64_t lo = tile * n_split, hi = lo + n_split;        // the tile's items
        // first CTA whose range starts after `lo`:  floor(b n_items / n_ctas) > lo  <=>  b > ((lo + 1) n_ctas - 1) / n_items
        int64_t b = ((lo + 1) * n_ctas - 1) / n_items + 1;
        for (; b < n_ctas; ++b) {
            const int64_t start = b * n_items / n_ctas;
            if (start <= lo) continue;
            if (start >= hi) break;
            const float4 v = reinterpret_cast<const float4 *>(partials)[(start - lo) * n_elem4 + i];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    const int cg = (int)(i & 15);
    const float4 c1 = reinterpret_cast<const float4 *>(colsum)[cg];
    const float4 c2 = reinterpret_cast<const float4 *>(colsum)[16 + cg];
    const float sh = out_shift ? out_shift[i >> 4] : 0.f;
    s.x = s.x - sh * c1.x - c2.x; s.y = s.y - sh * c1.y - c2.y;
    s.z = s.z - sh * c1.z - c2.z; s.w = s.w - sh * c1.w - c2.w;
    reinterpret_cast<float4 *>(Y)[i] = s;
}

// split-K heuristic (one CTA per SM): minimise  waves x (stages per CTA + fixed cost of a CTA), the fixed cost (barrier /
// TMEM set-up, first loads, epilogue: ~3 us) counted as 4 stages; among choices within 1.5 % the fewest slabs win
// (less partial traffic).
// ACCURACY bound on the slab length: the tensor core's fp32 accumulate TRUNCATES, so an accumulator that takes n UMMAs
// ends up ~n x 1.4e-8 too small in magnitude (measured, scripts/pass_accuracy.py: 16 slabs -1.1e-6, 1 slab -2e-5 ... -5e-5
// at 50 000 x 20 000).  At config c5 the unbounded heuristic gave slabs of 700+ stages and the gene-sharded and
// single-GPU embeddings of the same matrix differed by 7e-2 in the trailing components.  Slabs are therefore capped at
// kMaxSlabStages stages (4 UMMAs per accumulator and stage); the slab partials are summed in round-to-nearest fp32 by
// the reduce kernel.
constexpr int kMaxSlabStages = 32;
__host__ int tc_choose_split(int64_t n_tiles, int64_t n_k) {
    static const int forced = []() { const char *e = getenv("GC_RSVD_SPLIT"); return e ? atoi(e) : 0; }();   // developer knobs
    static const int max_stages = []() { const char *e = getenv("GC_RSVD_MAX_STAGES"); return e ? atoi(e) : kMaxSlabStages; }();
    if (forced > 0) return forced;
    const int64_t slots = (int64_t)kNumSMs;
    const int64_t stages_total = ceil_div<int64_t>(n_k, TC_TK);
    const int s_min = (int)std::max<int64_t>(1, ceil_div<int64_t>(stages_total, max_stages));
    int best = s_min;
    int64_t best_cost = 0;
    for (int s = s_min; s <= s_min + 64; ++s) {
        const int64_t k_per = ceil_div<int64_t>(ceil_div<int64_t>(n_k, s), TC_TK) * TC_TK;
        if (s > s_min && k_per < 8 * TC_TK) break;
        const int64_t cost = ceil_div<int64_t>(n_tiles * s, slots) * (k_per / TC_TK + 4);
        if (s == s_min || cost * 1000 < best_cost * 985) { best_cost = cost; best = s; }
    }
    return best;
}

}  // namespace gc

#include "rsvd_ts.cuh"

using namespace gc;

static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }
static int64_t ceil64(int64_t x) { return (x + 63) / 64 * 64; }
static int64_t ceil128(int64_t x) { return (x + 127) / 128 * 128; }

struct TcScratch {
    float *partials; __nv_bfloat16 *Qh, *Ql; double *colsum_part; float *colsum, *rpad, *tpad;
    int n_split; int64_t k_per_split; int prep_blocks; int64_t bytes;
};

// carve the caller's scratch for one pass direction
static TcScratch tc_layout(char *base, int64_t C, int64_t ld, int64_t n_out, int64_t n_k) {
    TcScratch s;
    s.n_split = tc_choose_split(ceil_div<int64_t>(n_out, TC_TM), n_k);
    s.k_per_split = ceil_div<int64_t>(ceil_div<int64_t>(n_k, s.n_split), TC_TK) * TC_TK;
    const int64_t n_k_pad = ceil64(n_k);
    s.prep_blocks = (int)ceil_div<int64_t>(n_k_pad, PREP_ROWS);
    char *p = base;
    s.partials = (float *)p;             p += align256((int64_t)s.n_split * n_out * TC_TN * 4);
    s.Qh = (__nv_bfloat16 *)p;           p += align256(n_k_pad * TC_TN * 2);
    s.Ql = (__nv_bfloat16 *)p;           p += align256(n_k_pad * TC_TN * 2);
    s.colsum_part = (double *)p;         p += align256((int64_t)s.prep_blocks * 128 * 8);
    s.colsum = (float *)p;               p += 512;
    s.rpad = (float *)p;                 p += align256(ceil128(C) * 4);
    s.tpad = (float *)p;                 p += align256(ceil128(ld) * 4);
    s.bytes = p - base;
    return s;
}

extern "C" int64_t gc_rsvd_scratch_bytes(int64_t C, int64_t G) {
    // worst case over both pass directions (ld <= ceil64(G) + 64 covers any padding the caller uses)
    const int64_t ld = ceil64(G) + 64;
    const int64_t b0 = tc_layout(nullptr, C, ld, C, G).bytes, b1 = tc_layout(nullptr, C, ld, G, C).bytes;
    const int64_t simt = simt_scratch_bytes(C, G);   // the SIMT cross-check kernel shares the buffer
    return std::max(std::max(b0, b1), simt) + 1024;
}

template <int TRANS, bool LO_CLIP>
static int tc_launch(const TcArgs &a, dim3 grid, cudaStream_t st) {
    static thread_local bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev]) {
        GC_CUDA(cudaFuncSetAttribute(rsvd_apply_tc_kernel<TRANS, LO_CLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_BYTES));
        configured[dev] = true;
    }
    rsvd_apply_tc_kernel<TRANS, LO_CLIP><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a);
    return 0;
}

// mode 0: TS kernel (product path), 1: SS kernel
static int rsvd_apply_impl(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *scratch,
                           void *stream, int mode) {
    GC_REQUIRE(M && Qin && Y && scratch, "null pointer");
    GC_REQUIRE(M->counts != nullptr && M->C >= 1 && M->G >= 1, "empty matrix");
    GC_REQUIRE(M->ld >= M->G && M->ld % 8 == 0 && ((uintptr_t)M->counts & 15) == 0, "counts: ld % 8 == 0, 16-byte aligned");
    GC_REQUIRE(M->ld <= ceil64(M->G) + 64, "counts: leading dimension padded by more than 127 columns");
    GC_REQUIRE(q >= 1 && q <= TC_TN, "q must be in [1, 64]");
    GC_REQUIRE(trans == 0 || trans == 1, "trans must be 0 or 1");
    GC_REQUIRE(((uintptr_t)Qin & 15) == 0 && ((uintptr_t)Y & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "alignment");
    GC_REQUIRE(M->total > 0.f && M->theta > 0.f, "total and theta must be positive");
    cudaStream_t st = as_stream(stream);
    TcArgs a;
    a.counts = M->counts; a.C = M->C; a.G = M->G; a.ld = M->ld;
    a.inv_theta = 1.0f / M->theta; a.clip = M->clip; a.magic = 0x4B000000u;
    a.n_out = trans == 0 ? M->C : M->G;
    a.n_k = trans == 0 ? M->G : M->C;
    const TcScratch s = tc_layout((char *)scratch, M->C, M->ld, a.n_out, a.n_k);
    a.n_split = s.n_split; a.k_per_split = s.k_per_split;
    a.partials = s.partials; a.Qh = s.Qh; a.Ql = s.Ql; a.rpad = s.rpad; a.tpad = s.tpad;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_out, TC_TM);
    // centring: M = Z - row_shift[c] - col_shift[g]; the shift indexed by k goes into colsum[64..], the one indexed by
    // the output row multiplies colsum[0..63]
    const float *k_shift = trans == 0 ? M->col_shift : M->row_shift;
    const float *out_shift = trans == 0 ? M->row_shift : M->col_shift;

    tc_prep_kernel<<<s.prep_blocks, 256, 0, st>>>(Qin, a.n_k, ceil64(a.n_k), k_shift, s.Qh, s.Ql, s.colsum_part,
                                                  M->row_sum_f, M->C, ceil128(M->C), M->col_sum_f, M->G, ceil128(M->ld),
                                                  1.0f / M->total, s.rpad, s.tpad);
    if (int rc = check_launch("tc_prep_kernel")) return rc;
    tc_colsum_kernel<<<128, 256, 0, st>>>(s.colsum_part, s.prep_blocks, s.colsum);
    if (int rc = check_launch("tc_colsum_kernel")) return rc;

    GC_REQUIRE(a.n_split <= 65535, "too many k slabs");   // (1 M cells / (48 stages x 64) = 326 slabs at config c5)
    GC_REQUIRE(M->C < (1ll << 31), "more than 2^31 cells");
    dim3 grid((unsigned)n_tiles, (unsigned)a.n_split);
    const bool lo_clip = !(M->clip * M->clip >= M->theta);     // z > -sqrt(theta) always: lower clip is dead otherwise
    // kernel choice: the TS-mode kernel (A operand in tensor memory, TMA-fed count tiles) is the product path; the
    // SS-mode kernel stays reachable (gc_rsvd_apply_ss, or GC_RSVD_MODE=ss for a whole run) as the cross-check of the
    // tests and for A/B timing
    static const bool force_ss = []() { const char *e = getenv("GC_RSVD_MODE"); return e && strcmp(e, "ss") == 0; }();
    int rc;
    if (mode == 1 || force_ss) {
        prof_begin(kProfRsvdPass, st);
        if (trans == 0) rc = lo_clip ? tc_launch<0, true>(a, grid, st) : tc_launch<0, false>(a, grid, st);
        else rc = lo_clip ? tc_launch<1, true>(a, grid, st) : tc_launch<1, false>(a, grid, st);
        prof_end(kProfRsvdPass, st);
    } else {
        GC_REQUIRE(M->ld < (1ll << 31), "leading dimension too large for the tensor map coordinates");
        // persistent launch: one CTA per SM walks the (tile, slab) items
        grid = dim3((unsigned)std::min<int64_t>(n_tiles * a.n_split, (int64_t)kNumSMs), 1);
        CUtensorMap map;
        if (int rcm = make_counts_tensor_map(&map, M->counts, M->C, M->ld, trans)) return rcm;
        // clip mode: 2 = both bounds, 1 = upper only (lower unreachable when clip >= sqrt(theta)), 0 = none (upper proved
        // inactive by the caller from the per-gene count maxima)
        const int clip_mode = lo_clip ? 2 : (M->upper_clip_inactive ? 0 : 1);
        prof_begin(kProfRsvdPass, st);
        if (trans == 0) rc = clip_mode == 2 ? ts_launch<0, 2>(map, a, grid, st)
                           : clip_mode == 1 ? ts_launch<0, 1>(map, a, grid, st) : ts_launch<0, 0>(map, a, grid, st);
        else rc = clip_mode == 2 ? ts_launch<1, 2>(map, a, grid, st)
                : clip_mode == 1 ? ts_launch<1, 1>(map, a, grid, st) : ts_launch<1, 0>(map, a, grid, st);
        prof_end(kProfRsvdPass, st);
    }
    if (rc) return rc;
    if (int rc2 = check_launch("rsvd_apply_tc_kernel")) return rc2;
    const int64_t o4 = a.n_out * TC_TN / 4;
    const bool segmented = !(mode == 1 || force_ss);
    rsvd_tc_reduce_kernel<<<(unsigned)ceil_div<int64_t>(o4, 256), 256, 0, st>>>(
        a.partials, o4, a.n_split, out_shift, s.colsum, Y, segmented ? (int)grid.x : 0, n_tiles * a.n_split);
    return check_launch("rsvd_tc_reduce_kernel");
}

extern "C" int gc_rsvd_apply(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *scratch,
                             void *stream) {
    return rsvd_apply_impl(M, trans, Qin, Y, q, scratch, stream, 0);
}

extern "C" int gc_rsvd_apply_ss(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *scratch,
                                void *stream) {
    return rsvd_apply_impl(M, trans, Qin, Y, q, scratch, stream, 1);
}
