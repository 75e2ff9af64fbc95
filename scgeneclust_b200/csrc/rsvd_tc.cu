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
// results are deterministic).  288 threads: warps 0-7 producers (then epilogue), warp 8 MMA issuer + TMEM owner.
// 2 stages x 48 KB shared memory per CTA, two CTAs per SM so one CTA's prologue/epilogue hides under the other's
// main loop.  Bound: HBM (algorithmic 2 B/entry); the producer ALU work (~13 instr/entry) is the practical limiter.
#include <cuda_bf16.h>

#include "common.cuh"

namespace gc {

constexpr int TC_TM = 128;           // output rows per CTA (UMMA M)
constexpr int TC_TN = 64;            // panel width (UMMA N)
constexpr int TC_TK = 64;            // k per stage (one 128-byte swizzle row of bf16)
constexpr int TC_STAGES = 2;
constexpr int TC_PRODUCERS = 256;    // warps 0..7
constexpr int TC_THREADS = TC_PRODUCERS + 32;
constexpr int TC_A_BYTES = TC_TM * TC_TK * 2;      // 16 KB per half (hi or lo)
constexpr int TC_B_BYTES = TC_TN * TC_TK * 2;      // 8 KB per half
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;

struct TcArgs {
    const uint16_t *counts; int64_t C, G, ld;
    const float *row_sum_f, *col_sum_f; float inv_total, inv_theta, clip;
    const float *row_shift, *col_shift;
    const __nv_bfloat16 *Qh, *Ql;        // pre-split panel, [n_k][64] each
    float *partials;                     // [n_split][n_out][64]
    int64_t n_out, n_k; int n_split; int64_t k_per_split;   // multiple of TC_TK
    int *error_flag;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__host__ __device__ constexpr uint32_t umma_idesc(int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(TC_TN >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

// 16-byte chunk `chunk` (0..7) of 128-byte row `row` inside a 1024-byte-aligned SWIZZLE_128B region
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t chunk) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

// residual of one count (low 16 bits of x16), before the centring shifts
__device__ __forceinline__ float tc_resid(uint32_t x16, float r, float t, float inv_theta, float clip) {
    const float xf = __uint_as_float(0x4B000000u | x16) - 8388608.0f;   // exact uint16 -> float without I2F
    const float mu = r * t;
    const float var = mu * fmaf(mu, inv_theta, 1.0f);
    const float rs = rsqrtf(var);
    const float z = fmaf(xf, rs, -mu * rs);
    return fminf(fmaxf(z, -clip), clip);
}

// split 8 floats into packed bf16 hi (truncation) and bf16 lo (rounded remainder) chunks
__device__ __forceinline__ void split8(const float (&z)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t b0 = __float_as_uint(z[2 * p]), b1 = __float_as_uint(z[2 * p + 1]);
        h[p] = __byte_perm(b0, b1, 0x7632);                         // {hi16(b0), hi16(b1)}
        const float l0 = z[2 * p] - __uint_as_float(b0 & 0xffff0000u);
        const float l1 = z[2 * p + 1] - __uint_as_float(b1 & 0xffff0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(l0, l1);     // .x = l0 (low half)
        l[p] = *reinterpret_cast<const uint32_t *>(&lp);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------------------------
// panel pre-split: Q fp32 [n_k][64] -> Qh, Ql bf16 (round-to-nearest hi, rounded remainder lo)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_panel_kernel(const float *__restrict__ Q, int64_t n4, __nv_bfloat16 *__restrict__ Qh,
                   __nv_bfloat16 *__restrict__ Ql) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 q = reinterpret_cast<const float4 *>(Q)[i];
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(q.x, q.y), h1 = __floats2bfloat162_rn(q.z, q.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(q.x - f0.x, q.y - f0.y);
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(q.z - f1.x, q.w - f1.y);
    reinterpret_cast<__nv_bfloat162 *>(Qh)[2 * i] = h0;
    reinterpret_cast<__nv_bfloat162 *>(Qh)[2 * i + 1] = h1;
    reinterpret_cast<__nv_bfloat162 *>(Ql)[2 * i] = l0;
    reinterpret_cast<__nv_bfloat162 *>(Ql)[2 * i + 1] = l1;
}

// ---------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------
template <int TRANS>
__global__ void __launch_bounds__(TC_THREADS, 2)
rsvd_apply_tc_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
    // barriers: full[s] @ +8s, empty[s] @ +16+8s, accum @ +32 ; tmem address @ +40
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + TC_STAGES * TC_STAGE_BYTES + 40);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_TM;
    const int split = blockIdx.y;
    const int64_t k_lo = split * a.k_per_split, k_hi = min(a.n_k, k_lo + a.k_per_split);
    const int n_it = (int)ceil_div<int64_t>(max((int64_t)0, k_hi - k_lo), TC_TK);

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) {
                mbar_init(bar_base + 8 * s, TC_PRODUCERS / 32);   // one arrive per producer warp
                mbar_init(bar_base + 16 + 8 * s, 1);              // tcgen05.commit
            }
            mbar_init(bar_base + 32, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32((const void *)tmem_slot), TC_TN);    // 64 fp32 accumulator columns
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 8) {
        // =============================== producers ===============================
        // TRANS 0: A tile = 128 cells x 64 genes, K-major.   chunk col j = tid%8 (8 genes), rows tid/8 + 32 i
        // TRANS 1: A tile = 128 genes x 64 cells, MN-major.  chunk col jj = tid%16 (8 genes), cells tid/16 + 16 i
        // B tile  = 64 k x 64 n (hi, lo), MN-major: row = k, chunk = 8 panel columns; thread -> rows tid/8 + 32 i
        const int a_col = TRANS == 0 ? (tid & 7) : (tid & 15);
        const int a_row0 = TRANS == 0 ? (tid >> 3) : (tid >> 4);
        constexpr int A_ROW_STEP = TRANS == 0 ? 32 : 16;
        const int b_chunk = tid & 7, b_row0 = tid >> 3;

        // per-thread constants
        float r_c[4] = {0.f, 0.f, 0.f, 0.f}, sh_c[4] = {0.f, 0.f, 0.f, 0.f};   // TRANS 0: fixed cells
        float t_g[8], sh_g[8];                                                  // TRANS 1: fixed genes
        bool g_ok[8];
        if (TRANS == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t c = m0 + a_row0 + A_ROW_STEP * i;
                if (c < a.C) { r_c[i] = a.row_sum_f[c]; sh_c[i] = a.row_shift ? a.row_shift[c] : 0.f; }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int64_t g = m0 + a_col * 8 + e;
                g_ok[e] = g < a.G;
                t_g[e] = g_ok[e] ? a.col_sum_f[g] * a.inv_total : 1.f;
                sh_g[e] = (g_ok[e] && a.col_shift) ? a.col_shift[g] : 0.f;
            }
        }

        uint4 raw[4], qh[2], ql[2];
        float4 tgv[2], shv[2];     // TRANS 0: per-stage gene constants of this thread's 8 genes
        float rcs[4], shs[4];      // TRANS 1: per-stage cell constants of this thread's 4 cells
        auto load_stage = [&](int it) {
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
            if (TRANS == 0) {
                const int64_t g0 = kk + a_col * 8;
                const bool gin = g0 < a.ld;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t c = m0 + a_row0 + A_ROW_STEP * i;
                    raw[i] = (c < a.C && gin) ? __ldg(reinterpret_cast<const uint4 *>(a.counts + c * a.ld + g0))
                                              : make_uint4(0, 0, 0, 0);
                }
                float tt[8], ss[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int64_t g = g0 + e;
                    const bool ok = g < k_hi;          // k_hi <= G
                    tt[e] = ok ? __ldg(a.col_sum_f + g) * a.inv_total : 0.f;
                    ss[e] = (ok && a.col_shift) ? __ldg(a.col_shift + g) : 0.f;
                }
                tgv[0] = make_float4(tt[0], tt[1], tt[2], tt[3]); tgv[1] = make_float4(tt[4], tt[5], tt[6], tt[7]);
                shv[0] = make_float4(ss[0], ss[1], ss[2], ss[3]); shv[1] = make_float4(ss[4], ss[5], ss[6], ss[7]);
            } else {
                const int64_t g0 = m0 + a_col * 8;
                const bool gin = g0 < a.ld;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t c = kk + a_row0 + A_ROW_STEP * i;
                    const bool ok = c < k_hi;          // k_hi <= C
                    raw[i] = (ok && gin) ? __ldg(reinterpret_cast<const uint4 *>(a.counts + c * a.ld + g0))
                                         : make_uint4(0, 0, 0, 0);
                    rcs[i] = ok ? __ldg(a.row_sum_f + c) : 0.f;
                    shs[i] = (ok && a.row_shift) ? __ldg(a.row_shift + c) : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t kr = kk + b_row0 + 32 * i;
                if (kr < k_hi) {
                    qh[i] = __ldg(reinterpret_cast<const uint4 *>(a.Qh + kr * TC_TN + b_chunk * 8));
                    ql[i] = __ldg(reinterpret_cast<const uint4 *>(a.Ql + kr * TC_TN + b_chunk * 8));
                } else {
                    qh[i] = make_uint4(0, 0, 0, 0); ql[i] = make_uint4(0, 0, 0, 0);
                }
            }
        };

        auto store_stage = [&](int it, uint8_t *st) {
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
            uint8_t *a_hi = st, *a_lo = st + TC_A_BYTES, *b_hi = st + 2 * TC_A_BYTES, *b_lo = b_hi + TC_B_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
                float z[8];
                if (TRANS == 0) {
                    const float tt[8] = {tgv[0].x, tgv[0].y, tgv[0].z, tgv[0].w, tgv[1].x, tgv[1].y, tgv[1].z, tgv[1].w};
                    const float ss[8] = {shv[0].x, shv[0].y, shv[0].z, shv[0].w, shv[1].x, shv[1].y, shv[1].z, shv[1].w};
                    const bool row_ok = (m0 + a_row0 + A_ROW_STEP * i) < a.C;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t x16 = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                        const float v = tc_resid(x16, r_c[i], tt[e], a.inv_theta, a.clip) - sh_c[i] - ss[e];
                        z[e] = (row_ok && tt[e] > 0.f) ? v : 0.f;     // tt == 0 marks k >= k_hi
                    }
                } else {
                    const bool cell_ok = (kk + a_row0 + A_ROW_STEP * i) < k_hi;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t x16 = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                        const float v = tc_resid(x16, rcs[i], t_g[e], a.inv_theta, a.clip) - shs[i] - sh_g[e];
                        z[e] = (cell_ok && g_ok[e]) ? v : 0.f;
                    }
                }
                uint4 hi, lo;
                split8(z, hi, lo);
                uint32_t off;
                if (TRANS == 0) off = swz128(a_row0 + A_ROW_STEP * i, a_col);
                else off = (uint32_t)(a_col >> 3) * 8192u + swz128(a_row0 + A_ROW_STEP * i, a_col & 7);
                *reinterpret_cast<uint4 *>(a_hi + off) = hi;
                *reinterpret_cast<uint4 *>(a_lo + off) = lo;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t off = swz128(b_row0 + 32 * i, b_chunk);
                *reinterpret_cast<uint4 *>(b_hi + off) = qh[i];
                *reinterpret_cast<uint4 *>(b_lo + off) = ql[i];
            }
        };

        if (n_it > 0) load_stage(0);
        for (int it = 0; it < n_it; ++it) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            mbar_wait(bar_base + 16 + 8 * s, ph ^ 1);         // stage free (passes at once on the first lap)
            store_stage(it, smem + s * TC_STAGE_BYTES);
            if (it + 1 < n_it) load_stage(it + 1);            // next global loads fly while the MMA runs
            fence_proxy_async_smem();                         // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8 * s);
        }

        // =============================== epilogue ===============================
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (= output rows), column half w / 4
        if (n_it > 0) {
            mbar_wait(bar_base + 32, 0);
            tc_fence_after();
        }
        const int lane_grp = warp & 3, col_half = warp >> 2;
        uint32_t v[32];
        if (n_it > 0) {
            tmem_ld32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(col_half * 32), v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int64_t row = m0 + lane_grp * 32 + lane;
        if (row < a.n_out) {
            float4 *dst = reinterpret_cast<float4 *>(a.partials + ((int64_t)split * a.n_out + row) * TC_TN + col_half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
    } else {
        // =============================== MMA issuer (one thread) ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(TRANS == 0 ? 0 : 1, 1);
            // A: K-major (TRANS 0): SBO = 1024 (next 8 rows), K step = 32 B.
            //    MN-major (TRANS 1): LBO = 8192 (next 64 genes), SBO = 1024 (next 8 cells), K step = 2048 B.
            // B: MN-major: one 64-wide block, SBO = 1024 (next 8 k), K step = 2048 B.
            constexpr uint32_t a_lbo = TRANS == 0 ? 16u : 8192u, a_sbo = 1024u, a_kstep = TRANS == 0 ? 32u : 2048u;
            constexpr uint32_t b_lbo = 8192u, b_sbo = 1024u, b_kstep = 2048u;
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait(bar_base + 8 * s, ph);
                tc_fence_after();
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint32_t a_hi = st, a_lo = st + TC_A_BYTES, b_hi = st + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
                for (int k = 0; k < TC_TK / 16; ++k) {
                    const uint64_t dah = umma_desc(a_hi + k * a_kstep, a_lbo, a_sbo);
                    const uint64_t dal = umma_desc(a_lo + k * a_kstep, a_lbo, a_sbo);
                    const uint64_t dbh = umma_desc(b_hi + k * b_kstep, b_lbo, b_sbo);
                    const uint64_t dbl = umma_desc(b_lo + k * b_kstep, b_lbo, b_sbo);
                    umma_bf16(tmem_d, dah, dbh, idesc, (it | k) != 0 ? 1u : 0u);
                    umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                    umma_bf16(tmem_d, dal, dbh, idesc, 1u);
                }
                umma_commit(bar_base + 16 + 8 * s);           // frees the stage when these MMAs have read it
            }
            if (n_it > 0) umma_commit(bar_base + 32);         // accumulator complete
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_d, TC_TN);
}

__global__ void rsvd_tc_reduce_kernel(const float *__restrict__ partials, int64_t n_elem4, int n_split,
                                      float *__restrict__ Y) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_elem4) return;
    float4 s = reinterpret_cast<const float4 *>(partials)[i];
    for (int p = 1; p < n_split; ++p) {
        const float4 v = reinterpret_cast<const float4 *>(partials)[(int64_t)p * n_elem4 + i];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4 *>(Y)[i] = s;
}

// split-K heuristic: fill a whole number of waves of (148 SMs x 2 resident CTAs)
__host__ int tc_choose_split(int64_t n_tiles, int64_t n_k) {
    const int64_t slots = (int64_t)kNumSMs * 2;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64; ++s) {
        const int64_t k_per = ceil_div<int64_t>(ceil_div<int64_t>(n_k, s), TC_TK) * TC_TK;
        if (s > 1 && k_per < 16 * TC_TK) break;
        const int64_t ctas = n_tiles * s;
        const double eff = (double)ctas / (double)(ceil_div<int64_t>(ctas, slots) * slots);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

}  // namespace gc

using namespace gc;

static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t gc_rsvd_scratch_bytes(int64_t C, int64_t G) {
    // per direction: split-K partials + the pre-split bf16 panel (hi, lo); worst case over both directions
    auto need = [](int64_t n_out, int64_t n_k) {
        const int s = tc_choose_split(ceil_div<int64_t>(n_out, TC_TM), n_k);
        return align256((int64_t)s * n_out * TC_TN * 4) + 2 * align256(n_k * TC_TN * 2);
    };
    const int64_t b0 = need(C, G), b1 = need(G, C);
    const int64_t simt = simt_scratch_bytes(C, G);   // the SIMT cross-check kernel shares the buffer
    return std::max(std::max(b0, b1), simt) + 512;
}

extern "C" int gc_rsvd_apply(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *scratch,
                             void *stream) {
    GC_REQUIRE(M && Qin && Y && scratch, "null pointer");
    GC_REQUIRE(M->counts != nullptr && M->C >= 1 && M->G >= 1, "empty matrix");
    GC_REQUIRE(M->ld >= M->G && M->ld % 8 == 0 && ((uintptr_t)M->counts & 15) == 0, "counts: ld % 8 == 0, 16-byte aligned");
    GC_REQUIRE(q >= 1 && q <= TC_TN, "q must be in [1, 64]");
    GC_REQUIRE(trans == 0 || trans == 1, "trans must be 0 or 1");
    GC_REQUIRE(((uintptr_t)Qin & 15) == 0 && ((uintptr_t)Y & 15) == 0 && ((uintptr_t)scratch & 255) == 0, "alignment");
    cudaStream_t st = as_stream(stream);
    TcArgs a;
    a.counts = M->counts; a.C = M->C; a.G = M->G; a.ld = M->ld;
    a.row_sum_f = M->row_sum_f; a.col_sum_f = M->col_sum_f;
    a.inv_total = 1.0f / M->total; a.inv_theta = 1.0f / M->theta; a.clip = M->clip;
    a.row_shift = M->row_shift; a.col_shift = M->col_shift;
    a.n_out = trans == 0 ? M->C : M->G;
    a.n_k = trans == 0 ? M->G : M->C;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_out, TC_TM);
    a.n_split = tc_choose_split(n_tiles, a.n_k);
    a.k_per_split = ceil_div<int64_t>(ceil_div<int64_t>(a.n_k, a.n_split), TC_TK) * TC_TK;
    char *p = (char *)scratch;
    a.partials = (float *)p;                         p += align256((int64_t)a.n_split * a.n_out * TC_TN * 4);
    __nv_bfloat16 *Qh = (__nv_bfloat16 *)p;          p += align256(a.n_k * TC_TN * 2);
    __nv_bfloat16 *Ql = (__nv_bfloat16 *)p;
    a.Qh = Qh; a.Ql = Ql;
    a.error_flag = nullptr;

    const int64_t n4 = a.n_k * TC_TN / 4;
    split_panel_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, st>>>(Qin, n4, Qh, Ql);
    if (int rc = check_launch("split_panel_kernel")) return rc;

    static thread_local bool configured = false;
    if (!configured) {
        GC_CUDA(cudaFuncSetAttribute(rsvd_apply_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        GC_CUDA(cudaFuncSetAttribute(rsvd_apply_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        configured = true;
    }
    GC_REQUIRE(a.n_split <= 65535, "too many k slabs");
    dim3 grid((unsigned)n_tiles, (unsigned)a.n_split);
    prof_begin(kProfRsvdPass, st);
    if (trans == 0) rsvd_apply_tc_kernel<0><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a);
    else rsvd_apply_tc_kernel<1><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a);
    prof_end(kProfRsvdPass, st);
    if (int rc = check_launch("rsvd_apply_tc_kernel")) return rc;
    const int64_t o4 = a.n_out * TC_TN / 4;
    rsvd_tc_reduce_kernel<<<(unsigned)ceil_div<int64_t>(o4, 256), 256, 0, st>>>(a.partials, o4, a.n_split, Y);
    return check_launch("rsvd_tc_reduce_kernel");
}
