// Randomized-SVD pass, version 4: the A operand (the Pearson-residual tile) lives in TENSOR MEMORY (tcgen05.mma with
// A from TMEM, "TS mode"), the raw uint16 count tiles arrive through a TMA-fed shared-memory ring.
// Included by rsvd_tc.cu (shares its PTX wrappers, the prep / column-sum / reduce kernels and the scratch layout).
//
// Why: the SS-mode kernel above moves 88 KB through shared memory per 128 x 64 stage (32 KB of
// A hi/lo written by the producers, 56 KB fetched by the UMMAs) and its producers spend ~1/3 of their instructions on
// global-load address arithmetic, bounds clamps and shared-memory stores.  Here
//   * one elected thread issues ONE tensor-map TMA per stage for the 16 KB raw count tile (zero fill outside the matrix,
//     so there is no bounds code at all) + bulk copies of the stage's 64 per-k constants and of the panel tile;
//   * a producer thread owns one output row of the tile (= its TMEM lane): it reads the row's counts from shared memory
//     (M Q: eight swizzled LDS.128; M^T Q: 64 conflict-free LDS.U16 down a column of the tile), forms the residuals in
//     registers, splits them into bf16 hi + lo and writes both straight into TMEM with tcgen05.st (256 B/clk);
//   * the tensor core reads A from TMEM and only the panel tile [Q_hi | Q_lo] from shared memory (24 KB per stage), and an
//     A-in-TMEM M128 x N64 UMMA runs at its 32-cycle floor (48 in SS mode);
//   * BOTH pass directions use the same kernel - the direction only changes how the raw tile is read.
// Accumulators: D1 (128 columns) += A_hi [Q_hi | Q_lo] (one N = 128 UMMA), D2 (64 columns) += A_lo Q_hi; the epilogue adds
// the three partial results.  The small A_lo Q_hi products get their OWN accumulator: the tensor core's fp32 accumulate
// truncates, so adding them into the large hi.hi sums (tried: 128 accumulator columns, 6 stages) loses them - the PCA
// scores moved from 8e-5 to 8e-4 of the oracle's.  TMEM: 192 accumulator columns + 5 A stages x 64 columns (hi: 32
// columns = 64 bf16, lo: 32) = 512.
//
// CTA = one 128-row output tile x one k-slab, 22 warps:
//   warps 0-19  producers, five groups of four warps (warp w of a group owns TMEM lanes 32 (w % 4) .. +31); group g owns the
//               stages it = g, g + 5, ... and TMEM A slot g;  the first 16 also run the epilogue
//   warp 20     MMA issuer + TMEM owner
//   warp 21     TMA issuer of the raw count tiles + per-k constants (nine tiles in flight: the ring is deeper than the
//               number of groups, so a group's next tile is already in shared memory when it finishes the current one)
//   warp 22     bulk-copy issuer of the panel tiles
#pragma once
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)

namespace gc {

#ifndef TS_GROUPS_N          // developer knobs (scripts/build_variant.py -DTS_GROUPS_N=.. -DTS_A_STAGES_N=..)
#define TS_GROUPS_N 4
#endif
#ifndef TS_A_STAGES_N
#define TS_A_STAGES_N 5
#endif
#ifndef TS_WAIT_SLEEP_NS_N
#define TS_WAIT_SLEEP_NS_N 0
#endif
#ifndef TS_I2F_WORDS_N      // of the 4 count words (8 entries) of a chunk, how many are converted by I2F.U16 (XU pipe)
#define TS_I2F_WORDS_N 1
#endif
constexpr int TS_I2F_WORDS = TS_I2F_WORDS_N;
constexpr int TS_EPI_WARPS_N = 4;
constexpr int TS_GROUPS = TS_GROUPS_N;                           // producer groups of four warps
constexpr int TS_A_STAGES = TS_A_STAGES_N;                       // A-operand slots in tensor memory (64 columns each)
constexpr int TS_PRODUCER_WARPS = 4 * TS_GROUPS;
constexpr int TS_EPI_WARPS = TS_EPI_WARPS_N;                     // one per TMEM lane quarter
constexpr int TS_MMA_WARP = TS_PRODUCER_WARPS + TS_EPI_WARPS;
constexpr int TS_THREADS = (TS_PRODUCER_WARPS + TS_EPI_WARPS + 3) * 32;
// Raw count tiles in flight (the TMA runs this far ahead).  MUST be a multiple of the number of producer groups: stage gs
// goes to group gs % TS_GROUPS and to slot gs % TS_RAW_STAGES, and a parity wait on the slot's full barrier is only safe
// for a waiter that has itself waited for the slot's PREVIOUS load.  With 7 (or 9) slots and 4 groups the previous use
// of a slot belonged to another group: when that earlier load was still in flight - TMA loads complete out of order, and
// under memory-system contention (NVLink traffic of a collective next door) one was seen to land later than the load
// issued three stages after it - the barrier was one phase behind, the parity test passed at once and the group consumed
// a stale tile: a 1e-5 wobble of one pass, run-to-run differences of the gene-sharded PCA and, at 8 GPUs with a million
// cells, a rank-deficient panel.  (Single-GPU runs never showed it: 41 of 41 repetitions bit-identical.)
constexpr int TS_RAW_STAGES = 8;
constexpr int TS_B_STAGES = 4;
constexpr int TS_RAW_BYTES = TC_TM * TC_TK * 2;                  // 16 KB raw uint16 tile
constexpr int TS_PANEL_BYTES = 2 * TC_B_BYTES;                   // 16 KB [Q_hi | Q_lo]
constexpr int TS_KC_BYTES = TC_TK * 4;                           // 256 B of per-k constants
constexpr int TS_OFF_PANEL = TS_RAW_STAGES * TS_RAW_BYTES;
constexpr int TS_OFF_KC = TS_OFF_PANEL + TS_B_STAGES * TS_PANEL_BYTES;
constexpr int TS_OFF_ACC = TS_OFF_KC + TS_RAW_STAGES * TS_KC_BYTES;         // epilogue running sums: 4 warps x 64 x 32 fp32
constexpr int TS_ACC_BYTES = TS_EPI_WARPS_N * TC_TN * 32 * 4;
constexpr int TS_OFF_BAR = TS_OFF_ACC + TS_ACC_BYTES;
constexpr int TS_N_BARS = 2 * TS_RAW_STAGES + 2 * TS_A_STAGES + 2 * TS_B_STAGES + 2;
// the tiles want a 1024-byte aligned base (SWIZZLE_128B).  The dynamic shared-memory window of a kernel without static
// shared memory starts 1024-byte aligned; the slack covers an offset of up to 512 bytes and the kernel traps beyond it
constexpr int TS_SMEM_USED = TS_OFF_BAR + 8 * TS_N_BARS + 16;
constexpr int TS_SMEM_BYTES = TS_SMEM_USED + 512;
static_assert(TS_RAW_STAGES % TS_GROUPS == 0, "every raw slot must belong to one producer group (see TS_RAW_STAGES)");
constexpr int TS_ACC_COLS = 192, TS_A_COLS = 64;   // D1 = [0, 128): A_hi [Q_hi | Q_lo];  D2 = [128, 192): A_lo Q_hi
static_assert(TS_ACC_COLS + TS_A_STAGES * TS_A_COLS <= TC_TMEM_COLS, "TMEM budget");
static_assert(TS_THREADS <= 1024, "too many warps");
static_assert(TS_SMEM_BYTES <= 227 * 1024, "shared-memory budget");

// mbarrier wait that lets the HARDWARE park the warp: try_wait with a suspend-time hint returns only when the phase
// completed or the hint ran out, so a waiting warp issues a handful of instructions per microsecond instead of a
// try_wait / nanosleep / branch loop (the kernel is instruction-issue bound: polling loops of the waiting warps were 15 %
// of all issued instructions).  Bounded: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity) {
#if TS_WAIT_SLEEP_NS_N > 0      // developer knob: plain try_wait + fixed nanosleep between polls
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        __nanosleep(TS_WAIT_SLEEP_NS_N);
        if (spin > (1u << 22)) __trap();
    }
    return;
#endif
    uint32_t ok;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if (spin > (1u << 20)) __trap();
    }
}

// Alternative wait for the warps that wait LONG (epilogue warps: a whole slab; copy issuers and MMA issuer: most of every
// stage): a non-blocking test and a plain nanosleep in between, instead of the parked try_wait, which wakes on every
// barrier event of the CTA and re-polls (ncu: ~6 of the 22 lane-instructions per matrix entry are these polls).
// Measured (50 000 x 20 000, both directions): parked 0.586 / 0.611 ms, epilogue 400 ns 0.590 / 0.617, + issuers 100 ns
// 0.590 / 0.617, + MMA 30 ns 0.590 / 0.615, 1000 / 150 / 200 ns 0.591 / 0.616 - no difference: the polls use issue slots
// nobody else wants.  Default 0 = parked; the knobs stay for the next kernel that might care.
#ifndef TS_EPI_SLEEP_NS_N
#define TS_EPI_SLEEP_NS_N 0
#endif
#ifndef TS_MMA_SLEEP_NS_N
#define TS_MMA_SLEEP_NS_N 0
#endif
#ifndef TS_ISSUER_SLEEP_NS_N
#define TS_ISSUER_SLEEP_NS_N 0
#endif
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    if constexpr (SLEEP_NS <= 0) {
        mbar_wait_park(bar, parity);
    } else {
        uint32_t ok;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (ok) return;
            __nanosleep(SLEEP_NS);
            if (spin > (1u << 24)) __trap();
        }
    }
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
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

// All UMMAs of one 128 x 64 stage (four k slices x {A_hi [Q_hi | Q_lo] -> D1, A_lo Q_hi -> D2}) and the two commits that
// release the stage's TMEM slot and panel tile, issued by ONE elected lane from ONE asm block.  The MMA-issuing warp is a
// serial resource: it shares its scheduler with four producer warps that are always ready to issue, so every instruction
// in its loop costs ~5 issue cycles.  Issuing each UMMA from its own wrapper (elect + predicate vote + four vector ->
// uniform register moves each) made that loop ~200 instructions per stage, i.e. as long as the stage itself, and the
// producers waited for TMEM slots (25-32 % of all stall samples); here it is one elect and straight-line UMMAs.
//   d: accumulator base (D1 = d, D2 = d + 128); a: the stage's A slot (hi at +0, lo at +32 columns; K = 16 is 8 columns);
//   b_lo / b_hi: shared-memory descriptor of the panel tile (K step = 2048 B >> 4 = 128); first != 0 accumulates slice 0.
__device__ __forceinline__ void umma_stage_ts(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc128,
                                              uint32_t idesc64, uint32_t accumulate_first, uint32_t bar_a, uint32_t bar_b) {
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, q;\n\t"
        ".reg .b32 d2, a1, a2, a3, l0, l1, l2, l3, b1, b2, b3;\n\t"
        ".reg .b64 db0, db1, db2, db3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p0, %6, 0;\n\t"
        "setp.eq.b32 p1, 0, 0;\n\t"
        "add.u32 d2, %0, 128;\n\t"
        "add.u32 a1, %1, 8;\n\t"
        "add.u32 a2, %1, 16;\n\t"
        "add.u32 a3, %1, 24;\n\t"
        "add.u32 l0, %1, 32;\n\t"
        "add.u32 l1, %1, 40;\n\t"
        "add.u32 l2, %1, 48;\n\t"
        "add.u32 l3, %1, 56;\n\t"
        "add.u32 b1, %2, 128;\n\t"
        "add.u32 b2, %2, 256;\n\t"
        "add.u32 b3, %2, 384;\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db0, %4, p0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d2], [l0], db0, %5, p0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], db1, %4, p1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d2], [l1], db1, %5, p1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], db2, %4, p1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d2], [l2], db2, %5, p1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], db3, %4, p1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [d2], [l3], db3, %5, p1;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "}"
        ::"r"(d), "r"(a), "r"(b_lo), "r"(b_hi), "r"(idesc128), "r"(idesc64), "r"(accumulate_first), "r"(bar_a), "r"(bar_b)
        : "memory");
}

// 2-D tiled TMA load global -> shared, completion counted in bytes on an mbarrier; streaming data: evict-first in L2 so
// that the panel tiles and constants (re-read by every CTA) stay resident
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int32_t x, int32_t y, uint32_t bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// residual of one entry: xf = the count as float, mu = r t
// CLIP: 0 = no test at all (the caller proved that no residual reaches +clip, gc_residual_matrix::upper_clip_inactive, and
// clip >= sqrt(theta) makes the lower bound unreachable), 1 = upper test only, 2 = both
template <int CLIP>
__device__ __forceinline__ float resid1(float xf, float mu, float inv_theta, float clip) {
    const float var = mu * fmaf(mu, inv_theta, 1.0f);
    float v = (xf - mu) * rsqrt_approx(var);
    if (CLIP >= 1) v = fminf(v, clip);
    if (CLIP >= 2) v = fmaxf(v, -clip);
    return v;
}
// two residuals -> one packed bf16 hi word (truncation) and one packed bf16 lo word (rounded remainder); the entry with
// the lower k sits in the low half, which is the order the tensor core expects for a K-major 16-bit operand
__device__ __forceinline__ void split_pair(float z0, float z1, uint32_t &h, uint32_t &l) {
    const uint32_t b0 = __float_as_uint(z0), b1 = __float_as_uint(z1);
    h = __byte_perm(b0, b1, 0x7632);
    const float l0 = z0 - __uint_as_float(b0 & 0xffff0000u);
    const float l1 = z1 - __uint_as_float(b1 & 0xffff0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(l0, l1);
    l = *reinterpret_cast<const uint32_t *>(&lp);
}

// One work item = one 128-row output tile x one k-slab.  The kernel is PERSISTENT: the grid is one CTA per SM and a CTA
// walks the items  blockIdx.x, blockIdx.x + gridDim.x, ...  with its pipeline (rings, barrier phases, TMEM) running
// straight through the item boundaries: the producers start the next item's stages while the UMMAs of the last stages
// and the epilogue of the previous item are still in flight.  Short slabs are what the accumulation accuracy wants
// (kMaxSlabStages); with one CTA launch per slab their fixed costs (launch, TMEM allocation, pipeline fill and drain,
// ~4 us) were 7 % of the pass at 48 stages per slab.
// Item order: item = tile * n_split + split, and CTA b takes the CONTIGUOUS range [floor(b n_items / gridDim.x),
// floor((b + 1) n_items / gridDim.x)): the slabs of a tile are consecutive items of (mostly) one CTA, whose epilogue warps
// add the slab results in registers (round-to-nearest fp32) and write one partial per (tile, CTA) instead of one per slab
// (at 32 stages per slab the per-slab partials were 130 MB per pass at config c3 and 470 MB per pass and GPU at config c5).
template <int TRANS, int CLIP>
__global__ void __launch_bounds__(TS_THREADS, 1)
rsvd_apply_ts_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (smem_base - smem_u32(smem_raw) > (uint32_t)(TS_SMEM_BYTES - TS_SMEM_USED)) __trap();     // see TS_SMEM_BYTES
    const uint32_t bar_base = smem_base + TS_OFF_BAR;
    const uint32_t bar_raw_full = bar_base, bar_raw_empty = bar_raw_full + 8 * TS_RAW_STAGES;
    const uint32_t bar_a_full = bar_raw_empty + 8 * TS_RAW_STAGES, bar_a_empty = bar_a_full + 8 * TS_A_STAGES;
    const uint32_t bar_b_full = bar_a_empty + 8 * TS_A_STAGES, bar_b_empty = bar_b_full + 8 * TS_B_STAGES;
    const uint32_t bar_acc_full = bar_b_empty + 8 * TS_B_STAGES, bar_acc_free = bar_acc_full + 8;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + (bar_acc_free + 8 - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_out, TC_TM);
    const int64_t n_items = n_tiles * a.n_split;
    // stages of an item: every slab has k_per_split / 64 stages except the last one of the k range
    auto item_stages = [&](int64_t item) -> int {
        const int64_t k_lo = (item % a.n_split) * a.k_per_split, k_hi = min(a.n_k, k_lo + a.k_per_split);
        return (int)ceil_div<int64_t>(max((int64_t)0, k_hi - k_lo), TC_TK);
    };
    const int64_t item_begin = (int64_t)blockIdx.x * n_items / gridDim.x;
    const int64_t item_end = ((int64_t)blockIdx.x + 1) * n_items / gridDim.x;

    if (warp == TS_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < TS_RAW_STAGES; ++s) {
                mbar_init(bar_raw_full + 8 * s, 1);            // the copy issuer's arrive.expect_tx
                mbar_init(bar_raw_empty + 8 * s, 4);           // one arrive per warp of the group that read the slot
            }
            for (int s = 0; s < TS_A_STAGES; ++s) {
                mbar_init(bar_a_full + 8 * s, 4);
                mbar_init(bar_a_empty + 8 * s, 1);             // tcgen05.commit
            }
            for (int s = 0; s < TS_B_STAGES; ++s) {
                mbar_init(bar_b_full + 8 * s, 1);
                mbar_init(bar_b_empty + 8 * s, 1);
            }
            mbar_init(bar_acc_full, 1);                        // tcgen05.commit after the last stage of an item
            mbar_init(bar_acc_free, TS_EPI_WARPS);             // the epilogue warps have read the accumulators
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32((const void *)tmem_slot), TC_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp < TS_PRODUCER_WARPS) {
        // =============================== producers ===============================
        const int group = warp >> 2, quarter = warp & 3;
        const int m = quarter * 32 + lane;                                        // row of the tile = TMEM lane
        const uint32_t a_lane = tmem_d + ((uint32_t)(quarter * 32) << 16) + TS_ACC_COLS;
        // TRANS 0: raw tile = 128 cells x 64 genes, 128-byte rows, TMA SWIZZLE_128B: 16-byte chunk c of row m sits at
        //          chunk c ^ (m & 7) - eight consecutive lanes read eight different bank groups.
        // TRANS 1: raw tile = 64 cells x 128 genes, 256-byte rows, no swizzle: a warp reads 32 consecutive uint16.
        const uint32_t row_off = TRANS == 0 ? (uint32_t)m * 128u : (uint32_t)m * 2u;
        // TRANS 0: byte offsets of this row's eight swizzled 16-byte chunks inside a raw tile, computed once
        uint32_t chunk_off[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) chunk_off[c] = row_off + ((((uint32_t)c) ^ (uint32_t)(m & 7)) << 4);
        int64_t g0 = 0;                                            // stages of this CTA's earlier items
        for (int64_t item = item_begin; item < item_end; ++item) {
            const int n_it = item_stages(item);
            const int64_t m0 = (item / a.n_split) * TC_TM;
            // the constant of this thread's output row (TRANS 0: r of its cell, TRANS 1: s/T of its gene); the padded
            // vectors hold 1.0 beyond the matrix, where TMA delivered zero counts: every residual stays finite
            const float fixed = __ldg((TRANS == 0 ? a.rpad : a.tpad) + m0 + m);
            // group g takes the stages whose running number gs = g0 + it is congruent to g: the rotation continues
            // across items, so the four groups stay evenly loaded whatever the slab length
            for (int it = (int)((group - g0 % TS_GROUPS + TS_GROUPS) % TS_GROUPS); it < n_it; it += TS_GROUPS) {
                const int64_t gs = g0 + it;
                const int rs = (int)(gs % TS_RAW_STAGES);         // raw slot of this stage (the TMA runs ahead of the groups)
                // TMEM A slot of this stage.  There are more slots than groups: the slot was last used by stage
                // gs - TS_A_STAGES, an EARLIER stage than this group's previous one, so its UMMAs have normally retired
                const int as = (int)(gs % TS_A_STAGES);
                const uint32_t a_col = a_lane + (uint32_t)as * TS_A_COLS;
                const uint32_t raw_tile = smem_base + rs * TS_RAW_BYTES;
                const uint32_t raw_row = raw_tile + row_off;
                const uint32_t kc_s = smem_base + TS_OFF_KC + rs * TS_KC_BYTES;
                // (the slot belongs to this group alone - see TS_RAW_STAGES - so the group has itself waited for the slot's
                // previous load and the parity cannot be read one lap off)
                mbar_wait_park(bar_raw_full + 8 * rs, (uint32_t)(gs / TS_RAW_STAGES) & 1);
#pragma unroll
                for (int c2 = 0; c2 < TC_TK / 16; ++c2) {             // 16 entries (k = 16 c2 .. 16 c2 + 15) per round
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c = 2 * c2 + half;                  // 16-byte chunk of 8 entries
                        const uint4 k0 = lds128(kc_s + c * 32), k1 = lds128(kc_s + c * 32 + 16);   // warp-wide broadcast
                        const float kc[8] = {__uint_as_float(k0.x), __uint_as_float(k0.y), __uint_as_float(k0.z),
                                             __uint_as_float(k0.w), __uint_as_float(k1.x), __uint_as_float(k1.y),
                                             __uint_as_float(k1.z), __uint_as_float(k1.w)};
                        float z[8];
                        if constexpr (TRANS == 0) {
                            const uint4 raw = lds128(raw_tile + chunk_off[c]);
                            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                // exact uint16 -> float.  Two ways, mixed to balance the pipes: splice the 16 bits under
                                // the 2^23 exponent and subtract 2^23 (PRMT on the ALU pipe + FADD: two issue slots), or one
                                // I2F.U16 with a half-word selector (one issue slot, but on the quarter-rate XU pipe that
                                // also runs the MUFU.RSQ of every entry)
                                float x0, x1;
                                if (p < TS_I2F_WORDS) {
                                    x0 = (float)(unsigned short)(w[p] & 0xffffu);
                                    x1 = (float)(unsigned short)(w[p] >> 16);
                                } else {
                                    x0 = __uint_as_float(__byte_perm(w[p], a.magic, 0x7410)) - 8388608.0f;
                                    x1 = __uint_as_float(__byte_perm(w[p], a.magic, 0x7432)) - 8388608.0f;
                                }
                                z[2 * p] = resid1<CLIP>(x0, fixed * kc[2 * p], a.inv_theta, a.clip);
                                z[2 * p + 1] = resid1<CLIP>(x1, fixed * kc[2 * p + 1], a.inv_theta, a.clip);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const uint32_t v = lds_u16(raw_row + (uint32_t)(c * 8 + e) * 256u);
                                const float xf = e < 2 * TS_I2F_WORDS ? (float)(unsigned short)v
                                                                      : __uint_as_float(v | a.magic) - 8388608.0f;
                                z[e] = resid1<CLIP>(xf, fixed * kc[e], a.inv_theta, a.clip);
                            }
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p) split_pair(z[2 * p], z[2 * p + 1], hi[4 * half + p], lo[4 * half + p]);
                    }
                    if (c2 == 0) {
                        mbar_wait_park(bar_a_empty + 8 * as, ((uint32_t)(gs / TS_A_STAGES) & 1) ^ 1);   // passes at once on the first lap
                        tc_fence_after();
                    }
                    tmem_st8(a_col + 8 * c2, hi);                     // 16 bf16 = 8 columns
                    tmem_st8(a_col + 32 + 8 * c2, lo);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_raw_empty + 8 * rs);              // the raw slot may be refilled
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_full + 8 * as);
            }
            g0 += n_it;
        }
    } else if (warp < TS_PRODUCER_WARPS + TS_EPI_WARPS) {
        // =============================== epilogue warps ===============================
        // warp q reads TMEM lanes 32 q .. +31 (= output rows) of the three partial accumulators (hi.hi, hi.lo, lo.hi, summed
        // in that fixed order) and writes the slab's partial result
        const int quarter = warp - TS_PRODUCER_WARPS;
        const uint32_t t_lane = tmem_d + ((uint32_t)(quarter * 32) << 16);
        uint32_t ph = 0;
        // this lane's output row, summed over the slabs of the segment, lives in shared memory (64 running sums per lane
        // in registers pushed the whole kernel over its register budget): element j of lane l at [j][l], conflict free
        const uint32_t acc_s = smem_base + TS_OFF_ACC + (uint32_t)quarter * (TC_TN * 32 * 4) + (uint32_t)lane * 4u;
        for (int j = 0; j < TC_TN; ++j) asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_s + j * 128), "f"(0.f) : "memory");
        int64_t seg_split = item_begin % a.n_split;                 // slot = split index of the segment's first slab
        for (int64_t item = item_begin; item < item_end; ++item) {
            const int n_it = item_stages(item);
            if (n_it > 0) {
                mbar_wait_sleep<TS_EPI_SLEEP_NS_N>(bar_acc_full, ph);
                tc_fence_after();
#pragma unroll 1
                for (int cq = 0; cq < 4; ++cq) {
                    // fixed order inside a slab: hi.hi + hi.lo + lo.hi, then onto the running sum (round-to-nearest fp32)
                    float t[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t[j] = 0.f;
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        uint32_t v[16];
                        tmem_ld16(t_lane + (uint32_t)(part * 64 + cq * 16), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) t[j] += __uint_as_float(v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float r;
                        const uint32_t ad = acc_s + (uint32_t)(cq * 16 + j) * 128u;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(ad) : "memory");
                        r += t[j];
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ad), "f"(r) : "memory");
                    }
                }
                tc_fence_before();                                  // the TMEM reads before the next item's overwrite
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_free);
                ph ^= 1;
            }
            const int64_t tile = item / a.n_split;
            if (item + 1 == item_end || (item + 1) / a.n_split != tile) {      // last slab of this CTA's segment of the tile
                const int64_t row = tile * TC_TM + quarter * 32 + lane;
                float4 *dst = reinterpret_cast<float4 *>(a.partials + (seg_split * a.n_out + row) * TC_TN);
#pragma unroll 4
                for (int j = 0; j < TC_TN / 4; ++j) {
                    float4 o;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o.x) : "r"(acc_s + (4 * j) * 128) : "memory");
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o.y) : "r"(acc_s + (4 * j + 1) * 128) : "memory");
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o.z) : "r"(acc_s + (4 * j + 2) * 128) : "memory");
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o.w) : "r"(acc_s + (4 * j + 3) * 128) : "memory");
                    if (row < a.n_out) dst[j] = o;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_s + (4 * j) * 128), "f"(0.f) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_s + (4 * j + 1) * 128), "f"(0.f) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_s + (4 * j + 2) * 128), "f"(0.f) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc_s + (4 * j + 3) * 128), "f"(0.f) : "memory");
                }
                seg_split = 0;                                      // the next segment starts a new tile
            }
        }
    } else if (warp == TS_MMA_WARP) {
        // =============================== MMA issuer (whole warp, one elected lane issues) =============
        constexpr uint32_t idesc128 = umma_idesc(0, 1, 128), idesc64 = umma_idesc(0, 1, 64);   // A K-major (TMEM), B MN-major
        constexpr uint32_t b_lbo = 8192u, b_sbo = 1024u;          // K step 2048 B: see umma_stage_ts
        const uint64_t db0 = umma_desc(smem_base + TS_OFF_PANEL, b_lbo, b_sbo);
        const uint32_t b_hiw = (uint32_t)(db0 >> 32), b_low0 = (uint32_t)db0;
        // warp-uniform copy of the TMEM base (it came out of shared memory through a vector load): everything derived from
        // it can then live in uniform registers
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
        int64_t g0 = 0;
        uint32_t ph = 0;
        for (int64_t item = item_begin; item < item_end; ++item) {
            const int n_it = item_stages(item);
            if (n_it == 0) continue;
            mbar_wait_sleep<TS_MMA_SLEEP_NS_N>(bar_acc_free, ph ^ 1);   // the previous item's accumulators have been read
            tc_fence_after();
            for (int it = 0; it < n_it; ++it) {
                const int64_t gs = g0 + it;
                const int s = (int)(gs % TS_A_STAGES), sb = (int)(gs % TS_B_STAGES);
                mbar_wait_sleep<TS_MMA_SLEEP_NS_N>(bar_b_full + 8 * sb, (uint32_t)(gs / TS_B_STAGES) & 1);
                mbar_wait_sleep<TS_MMA_SLEEP_NS_N>(bar_a_full + 8 * s, (uint32_t)(gs / TS_A_STAGES) & 1);
                tc_fence_after();
                umma_stage_ts(tmem_u, tmem_u + TS_ACC_COLS + (uint32_t)s * TS_A_COLS,
                              b_low0 + (uint32_t)sb * (TS_PANEL_BYTES >> 4), b_hiw, idesc128, idesc64, it > 0 ? 1u : 0u,
                              bar_a_empty + 8 * s, bar_b_empty + 8 * sb);
            }
            umma_commit_elect(bar_acc_full);                        // accumulators of this item complete
            g0 += n_it;
            ph ^= 1;
        }
        __syncwarp();
    } else if (warp == TS_MMA_WARP + 1) {
        // =============================== raw-tile TMA issuer (one thread) ===============================
        if (lane == 0) {
            const uint64_t policy = l2_evict_first_policy();
            int64_t g0 = 0;
            for (int64_t item = item_begin; item < item_end; ++item) {
                const int n_it = item_stages(item);
                const int64_t m0 = (item / a.n_split) * TC_TM, k_lo = (item % a.n_split) * a.k_per_split;
                const float *kconst = (TRANS == 0 ? a.tpad : a.rpad) + k_lo;
                for (int it = 0; it < n_it; ++it) {
                    const int64_t gs = g0 + it;
                    const int s = (int)(gs % TS_RAW_STAGES);
                    const int32_t kk = (int32_t)(k_lo + (int64_t)it * TC_TK);
                    mbar_wait_sleep<TS_ISSUER_SLEEP_NS_N>(bar_raw_empty + 8 * s, ((uint32_t)(gs / TS_RAW_STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(bar_raw_full + 8 * s, TS_RAW_BYTES + TS_KC_BYTES);
                    if (TRANS == 0) tma_load_2d(smem_base + s * TS_RAW_BYTES, &tmap, kk, (int32_t)m0, bar_raw_full + 8 * s, policy);
                    else tma_load_2d(smem_base + s * TS_RAW_BYTES, &tmap, (int32_t)m0, kk, bar_raw_full + 8 * s, policy);
                    bulk_g2s(smem_base + TS_OFF_KC + s * TS_KC_BYTES, kconst + (int64_t)it * TC_TK, TS_KC_BYTES, bar_raw_full + 8 * s);
                }
                g0 += n_it;
            }
        }
        __syncwarp();
    } else {
        // =============================== panel-tile copy issuer (one thread) ===============================
        if (lane == 0) {
            int64_t g0 = 0;
            for (int64_t item = item_begin; item < item_end; ++item) {
                const int n_it = item_stages(item);
                const int64_t k_lo = (item % a.n_split) * a.k_per_split;
                const char *qh = reinterpret_cast<const char *>(a.Qh) + (k_lo / TC_TK) * TC_B_BYTES;
                const char *ql = reinterpret_cast<const char *>(a.Ql) + (k_lo / TC_TK) * TC_B_BYTES;
                for (int it = 0; it < n_it; ++it) {
                    const int64_t gs = g0 + it;
                    const int sb = (int)(gs % TS_B_STAGES);
                    mbar_wait_sleep<TS_ISSUER_SLEEP_NS_N>(bar_b_empty + 8 * sb, ((uint32_t)(gs / TS_B_STAGES) & 1) ^ 1);
                    const uint32_t b_hi = smem_base + TS_OFF_PANEL + sb * TS_PANEL_BYTES;
                    mbar_arrive_expect_tx(bar_b_full + 8 * sb, TS_PANEL_BYTES);
                    bulk_g2s(b_hi, qh + (int64_t)it * TC_B_BYTES, TC_B_BYTES, bar_b_full + 8 * sb);
                    bulk_g2s(b_hi + TC_B_BYTES, ql + (int64_t)it * TC_B_BYTES, TC_B_BYTES, bar_b_full + 8 * sb);
                }
                g0 += n_it;
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TS_MMA_WARP) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

// ---- host side: the tensor map of the count matrix -------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// counts: uint16 [C][ld]; box = (64 genes x 128 cells, SWIZZLE_128B) for M Q, (128 genes x 64 cells, no swizzle) for M^T Q
static int make_counts_tensor_map(CUtensorMap *map, const uint16_t *counts, int64_t C, int64_t ld, int trans) {
    EncodeTiledFn enc = tensor_map_encoder();
    GC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)C};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {trans == 0 ? (cuuint32_t)TC_TK : (cuuint32_t)TC_TM, trans == 0 ? (cuuint32_t)TC_TM : (cuuint32_t)TC_TK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t *>(counts), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, trans == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        char msg[96];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
        return fail("make_counts_tensor_map", msg);
    }
    return 0;
}

template <int TRANS, int CLIP>
static int ts_launch(const CUtensorMap &map, const TcArgs &a, dim3 grid, cudaStream_t st) {
    static thread_local bool configured[kMaxDevices] = {};
    const int dev = current_device();
    if (!configured[dev]) {
        GC_CUDA(cudaFuncSetAttribute(rsvd_apply_ts_kernel<TRANS, CLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TS_SMEM_BYTES));
        configured[dev] = true;
    }
    rsvd_apply_ts_kernel<TRANS, CLIP><<<grid, TS_THREADS, TS_SMEM_BYTES, st>>>(map, a);
    return 0;
}

}  // namespace gc
