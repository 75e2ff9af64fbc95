// EXPERIMENTAL DRAFT - NOT PART OF THE PRODUCT BUILD, NEVER RUN ON A GPU YET (written when the round's GPU budget was
// spent; it compiles for sm_100a, so ptxas accepts every tcgen05 form used below - nothing more is claimed).
//
// `M Q` direction of the randomized-SVD pass with the A operand (the Pearson-residual tile) in TENSOR MEMORY instead of
// shared memory: DESIGN.md section 8, "Design note for the next pass-kernel version".  Why: the shipped SS-mode kernel
// moves 88 KB per 128 x 64 stage through shared memory (32 KB written by the producers, 56 KB fetched by the UMMAs) and
// that alone costs 0.32 ms per pass; with A in TMEM only the 24 KB of panel-tile fetches remain.
//
// How to try it (next round):
//   python scripts/build_variant.py ts --replace scripts/experimental/rsvd_tc_ts.cu
//   GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/libts.so python scripts/pass_times.py        # timing + checksum
//   GENECLUST_B200_LIB=... python -m pytest tests/test_gpu_embed.py -m gpu                           # parity
// This file includes the product translation unit with its entry point renamed and re-exports `gc_rsvd_apply`:
// trans == 0 goes to the TS kernel below, trans == 1 to the shipped SS kernel.
//
// Layout decisions to verify first (in this order):
//   1. TMEM A operand of kind::f16, M = 128: row m of the tile = TMEM lane m; one 32-bit column holds two consecutive
//      K elements, lower k in the low half; a K = 16 slice = 8 columns.  (If results are permuted, try the other order.)
//   2. tcgen05.st.32x32b.x4 by warp w writes lanes 32 (w % 4) .. +31, columns [col, col + 4): lane i <- thread i.
//   3. tcgen05.wait::st + tcgen05.fence::before_thread_sync before the mbarrier arrive is enough for the MMA (issued by
//      another warp after the barrier wait + fence::after_thread_sync) to see the stores.
#define gc_rsvd_apply gc_rsvd_apply_ss
#include "../../scgeneclust_b200/csrc/rsvd_tc.cu"
#undef gc_rsvd_apply

namespace gc {

constexpr int TS_STAGES = 5;                     // A stages in TMEM (64 columns each) = panel stages in shared memory
constexpr int TS_ACC_COLS = 192;                 // D1 = [0, 128): A_hi [Q_hi | Q_lo];  D2 = [128, 192): A_lo Q_hi
constexpr int TS_A_COLS = 64;                    // per stage: hi = 32 columns (64 bf16), lo = 32 columns
constexpr int TS_STAGE_BYTES = 2 * TC_B_BYTES;   // shared memory per stage: the panel tile [Q_hi | Q_lo], 16 KB
constexpr int TS_SMEM_BYTES = TS_STAGES * TS_STAGE_BYTES + 1024 + TC_BAR_BYTES;
static_assert(TS_ACC_COLS + TS_STAGES * TS_A_COLS == TC_TMEM_COLS, "TMEM budget");

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4 &v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]; whole converged warp, one elected lane issues (see umma_bf16_w)
template <bool ACCUM>
__device__ __forceinline__ void umma_bf16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACCUM ? 1 : 0) : "memory");
}

// CTA = 128 cells x one gene slab, 576 threads as in the SS kernel: warps 0-15 producers (4 groups x 4 warps, group g owns
// the stages it = g, g + 4, ...), warp 16 MMA issuer, warp 17 panel copier.  Inside a group, warp w owns TMEM lanes
// 32 (w % 4) .. +31: a thread owns ONE cell row of the tile and the 64 genes of the stage (128 contiguous bytes of counts).
template <bool LO_CLIP>
__global__ void __launch_bounds__(TC_THREADS, 1)
rsvd_apply_ts_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + TS_STAGES * TS_STAGE_BYTES;
    const uint32_t bar_full = bar_base, bar_empty = bar_full + 8 * TS_STAGES, bar_accum = bar_empty + 8 * TS_STAGES;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + (bar_accum + 8 - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_TM;
    const int split = blockIdx.y;
    const int64_t k_lo = split * a.k_per_split, k_hi = min(a.n_k, k_lo + a.k_per_split);
    const int n_it = (int)ceil_div<int64_t>(max((int64_t)0, k_hi - k_lo), TC_TK);

    if (warp == TC_PRODUCER_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < TS_STAGES; ++s) {
                mbar_init(bar_full + 8 * s, TC_GROUP_WARPS + 1);      // the four warps of the owning group + the panel copier
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
        const int group = warp / TC_GROUP_WARPS, quarter = warp % TC_GROUP_WARPS;
        const int64_t ld = a.ld;
        const int32_t row = min((int32_t)(m0 + quarter * 32 + lane), (int32_t)(a.C - 1));
        const float r = __ldg(a.rpad + m0 + quarter * 32 + lane);               // rpad covers ceil128(C), 1.0 beyond C
        const uint16_t *row_base = a.counts + (int64_t)row * ld;
        const uint32_t lane_base = tmem_d + ((uint32_t)(quarter * 32) << 16);    // this warp's TMEM lane quarter

        uint4 raw[TC_CHUNKS];
        auto load_stage = [&](int it) {
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
#pragma unroll
            for (int c = 0; c < TC_CHUNKS; ++c)
                raw[c] = __ldg(reinterpret_cast<const uint4 *>(row_base + min(kk + c * 8, ld - 8)));
        };
        if (group < n_it) load_stage(group);
        for (int it = group; it < n_it; it += TC_GROUPS) {
            const int s = it % TS_STAGES;
            const uint32_t ph = (it / TS_STAGES) & 1;
            const int64_t kk = k_lo + (int64_t)it * TC_TK;
            const uint32_t a_col = lane_base + TS_ACC_COLS + (uint32_t)s * TS_A_COLS;
            const bool more = it + TC_GROUPS < n_it;
            mbar_wait_backoff(bar_empty + 8 * s, ph ^ 1);                // slot last read by the MMAs of stage it - 5
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < TC_CHUNKS; ++c) {
                // the 8 column constants of this chunk: the same address for the whole warp (broadcast, L1 resident)
                const int64_t g0 = min(kk + c * 8, ld - 8);
                const float4 t0 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g0));
                const float4 t1 = __ldg(reinterpret_cast<const float4 *>(a.tpad + g0) + 1);
                const float cst[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                uint4 hi, lo;
                resid8<LO_CLIP, true>(raw[c], cst, r, a.inv_theta, a.clip, a.magic, hi, lo);
                if (more) {
                    const int64_t kn = kk + (int64_t)TC_GROUPS * TC_TK;
                    raw[c] = __ldg(reinterpret_cast<const uint4 *>(row_base + min(kn + c * 8, ld - 8)));
                }
                tmem_st4(a_col + 4 * c, hi);                             // 8 bf16 = 4 columns per chunk
                tmem_st4(a_col + 32 + 4 * c, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }

        // =============================== epilogue (as in the SS kernel, three partial accumulators) ===============
        if (n_it > 0) {
            mbar_wait(bar_accum, 0);
            tc_fence_after();
        }
        {
            const int lane_grp = warp & 3, col_q = warp >> 2;
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            if (n_it > 0) {
                const uint32_t t0 = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(col_q * 16);
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    uint32_t v[16];
                    tmem_ld16(t0 + part * 64, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
                }
            }
            const int64_t orow = m0 + lane_grp * 32 + lane;
            if (orow < a.n_out) {
                float4 *dst = reinterpret_cast<float4 *>(a.partials + ((int64_t)split * a.n_out + orow) * TC_TN + col_q * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
        }
    } else if (warp == TC_PRODUCER_WARPS) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc128 = umma_idesc(0, 1, 128), idesc64 = umma_idesc(0, 1, 64);   // A K-major (TMEM), B MN-major
        constexpr uint32_t b_lbo = 8192u, b_sbo = 1024u, b_kstep = 2048u;
        const uint64_t db0 = umma_desc(smem_base, b_lbo, b_sbo);
        const uint32_t b_hiw = (uint32_t)(db0 >> 32), b_low0 = (uint32_t)db0;
        for (int it = 0; it < n_it; ++it) {
            const int s = it % TS_STAGES;
            const uint32_t ph = (it / TS_STAGES) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t b_st = b_low0 + (uint32_t)s * (TS_STAGE_BYTES >> 4);
            const uint32_t a_st = tmem_d + TS_ACC_COLS + (uint32_t)s * TS_A_COLS;
#pragma unroll
            for (int k = 0; k < TC_TK / 16; ++k) {
                const uint32_t bh = b_st + k * (b_kstep >> 4);
                const uint32_t ah = a_st + 8 * k, al = a_st + 32 + 8 * k;        // K = 16 bf16 = 8 columns
                if (it == 0 && k == 0) {
                    umma_bf16_ts_w<false>(tmem_d, ah, bh, b_hiw, idesc128);
                    umma_bf16_ts_w<false>(tmem_d + 128, al, bh, b_hiw, idesc64);
                } else {
                    umma_bf16_ts_w<true>(tmem_d, ah, bh, b_hiw, idesc128);
                    umma_bf16_ts_w<true>(tmem_d + 128, al, bh, b_hiw, idesc64);
                }
            }
            umma_commit_elect(bar_empty + 8 * s);
        }
        if (n_it > 0) umma_commit_elect(bar_accum);
        __syncwarp();
    } else {
        // =============================== panel-tile copy issuer ===============================
        if (lane == 0) {
            const char *qh = reinterpret_cast<const char *>(a.Qh) + (k_lo / TC_TK) * TC_B_BYTES;
            const char *ql = reinterpret_cast<const char *>(a.Ql) + (k_lo / TC_TK) * TC_B_BYTES;
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TS_STAGES;
                const uint32_t ph = (it / TS_STAGES) & 1;
                mbar_wait_backoff(bar_empty + 8 * s, ph ^ 1);
                const uint32_t b_hi = smem_base + s * TS_STAGE_BYTES;
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

}  // namespace gc

using namespace gc;

template <bool LO_CLIP>
static int ts_launch(const TcArgs &a, dim3 grid, cudaStream_t st) {
    static thread_local bool configured = false;
    if (!configured) {
        GC_CUDA(cudaFuncSetAttribute(rsvd_apply_ts_kernel<LO_CLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        configured = true;
    }
    rsvd_apply_ts_kernel<LO_CLIP><<<grid, TC_THREADS, TS_SMEM_BYTES, st>>>(a);
    return 0;
}

extern "C" int gc_rsvd_apply(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *scratch,
                             void *stream) {
    if (trans != 0) return gc_rsvd_apply_ss(M, trans, Qin, Y, q, scratch, stream);
    GC_REQUIRE(M && Qin && Y && scratch, "null pointer");
    GC_REQUIRE(M->counts != nullptr && M->C >= 1 && M->G >= 1, "empty matrix");
    GC_REQUIRE(M->ld >= M->G && M->ld % 8 == 0 && ((uintptr_t)M->counts & 15) == 0, "counts: ld % 8 == 0, 16-byte aligned");
    GC_REQUIRE(M->ld <= ceil64(M->G) + 64, "counts: leading dimension padded by more than 127 columns");
    GC_REQUIRE(q >= 1 && q <= TC_TN, "q must be in [1, 64]");
    cudaStream_t st = as_stream(stream);
    TcArgs a;
    a.counts = M->counts; a.C = M->C; a.G = M->G; a.ld = M->ld;
    a.inv_theta = 1.0f / M->theta; a.clip = M->clip; a.magic = 0x4B000000u;
    a.n_out = M->C; a.n_k = M->G;
    const TcScratch s = tc_layout((char *)scratch, M->C, M->ld, a.n_out, a.n_k);
    a.n_split = s.n_split; a.k_per_split = s.k_per_split;
    a.partials = s.partials; a.Qh = s.Qh; a.Ql = s.Ql; a.rpad = s.rpad; a.tpad = s.tpad;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_out, TC_TM);
    tc_prep_kernel<<<s.prep_blocks, 256, 0, st>>>(Qin, a.n_k, ceil64(a.n_k), M->col_shift, s.Qh, s.Ql, s.colsum_part,
                                                  M->row_sum_f, M->C, ceil128(M->C), M->col_sum_f, M->G, ceil128(M->ld),
                                                  1.0f / M->total, s.rpad, s.tpad);
    if (int rc = check_launch("tc_prep_kernel")) return rc;
    tc_colsum_kernel<<<1, 1024, 0, st>>>(s.colsum_part, s.prep_blocks, s.colsum);
    if (int rc = check_launch("tc_colsum_kernel")) return rc;
    dim3 grid((unsigned)n_tiles, (unsigned)a.n_split);
    const bool lo_clip = !(M->clip * M->clip >= M->theta);
    prof_begin(kProfRsvdPass, st);
    const int rc = lo_clip ? ts_launch<true>(a, grid, st) : ts_launch<false>(a, grid, st);
    prof_end(kProfRsvdPass, st);
    if (rc) return rc;
    if (int rc2 = check_launch("rsvd_apply_ts_kernel")) return rc2;
    const int64_t o4 = a.n_out * TC_TN / 4;
    rsvd_tc_reduce_kernel<<<(unsigned)ceil_div<int64_t>(o4, 256), 256, 0, st>>>(a.partials, o4, a.n_split, M->row_shift,
                                                                               s.colsum, Y);
    return check_launch("rsvd_tc_reduce_kernel");
}
