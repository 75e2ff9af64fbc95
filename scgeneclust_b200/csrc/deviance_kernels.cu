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

// ---------------------------------------------------------------------------------------------------------
// The same sequential float32 column sums, piece-parallel and still bit-exact (m <= kSpecMaxCols).
//
// While the accumulator stays inside one binade [2^e, 2^(e+1)) (ulp u = 2^(e-23)) and keeps its sign,
// fl(acc + x) = acc + sgn * rn_u(sgn * x) in integer arithmetic in units of u; only a tie (|x| / u has fraction exactly
// 1/2) needs the parity of acc.  So a piece of kSpecPiece rows is summarised independently of what came before it as
// two integer deltas (acc even / odd on entry) plus the extremes of the running offset, given a GUESS of (sign, e) -
// taken from float64 prefix sums of the pieces.  One warp per column then walks the pieces: it checks the guess
// against the true accumulator (same sign and exponent, the running offset never leaves the binade's interior) and
// applies the delta, or replays the piece add by add when the check fails.  Exact by construction - the guess only
// decides the speed: the deviance terms grow steadily (a dozen replays in 1500 pieces), a column sum that wanders
// around zero replays most pieces and costs what the plain chain costs.  Prototype and adversarial test (ties, sign
// changes, binade crossings): scripts/proto/spec_colsum.py.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSpecPiece = 256, kSpecMaxCols = 64;

struct SpecSummary {
    int32_t ok, sgn, e, pad;
    int32_t d[2], mn[2], mx[2];
};

__device__ __forceinline__ const float *spec_array(const float *a0, const float *a1) { return blockIdx.y == 0 ? a0 : a1; }

// sums[arr][p * m + j] = float64 sum of column j over piece p
__global__ void __launch_bounds__(kSpecPiece)
colsum_spec_sums_kernel(const float *__restrict__ a0, const float *__restrict__ a1, int64_t C, int m, int64_t n_pieces,
                        double *__restrict__ sums) {
    const float *__restrict__ a = spec_array(a0, a1);
    __shared__ double warp_part[kSpecPiece / 32][kSpecMaxCols];
    const int64_t row = (int64_t)blockIdx.x * kSpecPiece + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < m; ++j) {
        double v = row < C ? (double)a[row * m + j] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) warp_part[warp][j] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < m) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kSpecPiece / 32; ++w) t += warp_part[w][threadIdx.x];
        sums[((int64_t)blockIdx.y * n_pieces + blockIdx.x) * m + threadIdx.x] = t;
    }
}

// in place: sums -> exclusive prefix over the pieces (the float64 estimate of the accumulator on entry of a piece).
// One warp per column: every lane scans a contiguous range of pieces, the range totals are scanned by shuffles.
__global__ void __launch_bounds__(32)
colsum_spec_prefix_kernel(double *__restrict__ sums, int m, int64_t n_pieces) {
    const int j = blockIdx.x, lane = threadIdx.x;
    double *s = sums + (int64_t)blockIdx.y * n_pieces * m + j;
    const int64_t per = (n_pieces + 31) / 32, p_lo = min(n_pieces, lane * per), p_hi = min(n_pieces, p_lo + per);
    double tot = 0.0;
#pragma unroll 4
    for (int64_t p = p_lo; p < p_hi; ++p) tot += s[p * m];
    double incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    double run = incl - tot;
    for (int64_t p = p_lo; p < p_hi; ++p) {
        const double t = s[p * m];
        s[p * m] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(128)
colsum_spec_summary_kernel(const float *__restrict__ a0, const float *__restrict__ a1, int64_t C, int m, int64_t n_pieces,
                           const double *__restrict__ est, SpecSummary *__restrict__ summ) {
    const float *__restrict__ a = spec_array(a0, a1);
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= n_pieces * m) return;
    const int64_t p = idx / m;
    const int j = (int)(idx - p * m);
    const double g = est[(int64_t)blockIdx.y * n_pieces * m + idx];
    SpecSummary out;
    out.ok = 0; out.sgn = 0; out.e = 0; out.pad = 0;
    out.d[0] = out.d[1] = out.mn[0] = out.mn[1] = out.mx[0] = out.mx[1] = 0;
    bool ok = p > 0 && isfinite(g) && g != 0.0;
    const int sgn = g > 0.0 ? 1 : -1;
    const int e = ok ? ilogb(fabs(g)) : 0;
    if (e < -100 || e > 100) ok = false;
    long long d0 = 0, d1 = 0, mn0 = 0, mn1 = 0, mx0 = 0, mx1 = 0;
    if (ok) {
        const int64_t r_lo = p * kSpecPiece, r_hi = min(C, r_lo + kSpecPiece);
        for (int64_t r = r_lo; r < r_hi; ++r) {
            const uint32_t b = __float_as_uint(a[r * m + j]);
            const int ex_raw = (b >> 23) & 0xff;
            uint32_t M = b & 0x7fffffu;
            if (ex_raw == 255) { ok = false; break; }                       // inf / nan: replay
            int ex;
            if (ex_raw == 0) ex = -149; else { M |= 0x800000u; ex = ex_raw - 150; }
            if (M == 0) continue;
            const int s = (b >> 31) ? -sgn : sgn;
            const int sh = (e - 23) - ex;
            long long inc0, inc1;
            if (sh <= 0) {
                if (sh < -2) { ok = false; break; }                          // |x| >= 4 |acc|: leaves the binade anyway
                inc0 = inc1 = (long long)M << (-sh);
            } else if (sh >= 25) {
                continue;                                                    // |x| < u / 2
            } else {
                const uint32_t q = M >> sh, rem = M & ((1u << sh) - 1u), half = 1u << (sh - 1);
                if (rem > half) {
                    inc0 = inc1 = (long long)q + 1;
                } else if (rem == half) {                                    // tie: to the even result
                    inc0 = (long long)q + (((0 + d0 + s * (long long)q) & 1) ? 1 : 0);
                    inc1 = (long long)q + (((1 + d1 + s * (long long)q) & 1) ? 1 : 0);
                } else {
                    inc0 = inc1 = q;
                }
            }
            d0 += s * inc0; d1 += s * inc1;
            mn0 = min(mn0, d0); mx0 = max(mx0, d0);
            mn1 = min(mn1, d1); mx1 = max(mx1, d1);
        }
        const long long lim = 1LL << 28;
        if (mn0 < -lim || mn1 < -lim || mx0 > lim || mx1 > lim) ok = false;
    }
    if (ok) {
        out.ok = 1; out.sgn = sgn; out.e = e;
        out.d[0] = (int32_t)d0; out.d[1] = (int32_t)d1;
        out.mn[0] = (int32_t)mn0; out.mn[1] = (int32_t)mn1;
        out.mx[0] = (int32_t)mx0; out.mx[1] = (int32_t)mx1;
    }
    summ[(int64_t)blockIdx.y * n_pieces * m + idx] = out;
}

// One warp per column; every lane carries the same accumulator.  The summaries are fetched 32 pieces at a time (one
// per lane) and handed round by shuffles.  A replayed piece is eight blocks of 32 rows, one row per lane and block,
// all loaded before the first add (and already during the previous piece when the summary says the piece cannot be
// taken in one step), handed to the chain by shuffles: four cycles per row, the FADD latency.
__global__ void __launch_bounds__(32)
colsum_spec_chain_kernel(const float *__restrict__ a0, const float *__restrict__ a1, int64_t C, int m, int64_t n_pieces,
                         const SpecSummary *__restrict__ summ, float *__restrict__ out0, float *__restrict__ out1,
                         int32_t *__restrict__ replays) {
    const float *__restrict__ a = spec_array(a0, a1);
    float *__restrict__ out = blockIdx.y == 0 ? out0 : out1;
    const int j = blockIdx.x, lane = threadIdx.x;
    const SpecSummary *S = summ + (int64_t)blockIdx.y * n_pieces * m + j;
    constexpr int kBlocks = kSpecPiece / 32;
    float acc = a[j];                                                        // the reduction starts from the first row
    int n_replay = 0;
    float nxt[kBlocks];
    bool have_nxt = false;
    auto load_piece = [&](int64_t p, float (&v)[kBlocks]) {
        const int64_t r_lo = p * kSpecPiece;
#pragma unroll
        for (int b = 0; b < kBlocks; ++b) {
            const int64_t r = r_lo + b * 32 + lane;
            v[b] = r < C ? a[r * m + j] : 0.f;
        }
    };
    for (int64_t p0 = 0; p0 < n_pieces; p0 += 32) {
        SpecSummary mine;
        mine.ok = 0;
        if (p0 + lane < n_pieces) mine = S[(p0 + lane) * m];
        const int n_here = (int)min((int64_t)32, n_pieces - p0);
        for (int k = 0; k < n_here; ++k) {
            const int64_t p = p0 + k;
            const int ok = __shfl_sync(0xffffffffu, mine.ok, k);
            // rows of the NEXT piece, when it is known to need a replay (same batch only: its summary is at hand)
            float cur[kBlocks];
            const bool have_cur = have_nxt;
            if (have_cur) {
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) cur[b] = nxt[b];
            }
            have_nxt = false;
            if (k + 1 < n_here && __shfl_sync(0xffffffffu, mine.ok, k + 1) == 0) {
                load_piece(p + 1, nxt);
                have_nxt = true;
            }
            bool fast = false;
            if (ok) {
                const int sgn = __shfl_sync(0xffffffffu, mine.sgn, k), e = __shfl_sync(0xffffffffu, mine.e, k);
                const uint32_t bits = __float_as_uint(acc);
                const int ex = (bits >> 23) & 0xff;
                if (ex != 0 && ex != 255 && ((bits >> 31) ? -1 : 1) == sgn && ex - 127 == e) {
                    const int32_t ai = (int32_t)((bits & 0x7fffffu) | 0x800000u);
                    const int par = ai & 1;
                    const int32_t d = __shfl_sync(0xffffffffu, par ? mine.d[1] : mine.d[0], k);
                    const int32_t mn = __shfl_sync(0xffffffffu, par ? mine.mn[1] : mine.mn[0], k);
                    const int32_t mx = __shfl_sync(0xffffffffu, par ? mine.mx[1] : mine.mx[0], k);
                    if (ai + mn >= (1 << 23) + 1 && ai + mx <= (1 << 24) - 2) {
                        acc = __uint_as_float((bits & 0xff800000u) | ((uint32_t)(ai + d) & 0x7fffffu));
                        fast = true;
                    }
                }
            }
            if (!fast) {
                ++n_replay;
                if (!have_cur) load_piece(p, cur);
                const int64_t r_lo = p * kSpecPiece, r_hi = min(C, r_lo + kSpecPiece);
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) {
                    const int64_t r0 = r_lo + b * 32;
                    const int first = (p == 0 && b == 0) ? 1 : 0;              // row 0 started the accumulator
                    const int cnt = (int)max((int64_t)0, min((int64_t)32, r_hi - r0));
                    if (cnt == 32 && first == 0) {
                        // all 32 broadcasts first (independent), then the dependent adds: the shuffles of the next block
                        // overlap the add chain of this one
                        float v[32];
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = __shfl_sync(0xffffffffu, cur[b], q);
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc = __fadd_rn(acc, v[q]);
                    } else {
                        for (int q = first; q < cnt; ++q) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, cur[b], q));
                    }
                }
            }
        }
    }
    if (lane == 0) {
        out[j] = acc;
        if (replays) replays[blockIdx.y * m + j] = n_replay;
    }
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

static int64_t spec_pieces(int64_t C) { return (C + kSpecPiece - 1) / kSpecPiece; }
static int64_t spec_bytes(int64_t C, int m) {
    if (m > kSpecMaxCols) return 0;
    return 2 * (dev_align(spec_pieces(C) * m * 8) + dev_align(spec_pieces(C) * m * (int64_t)sizeof(SpecSummary))) +
           dev_align(2 * kSpecMaxCols * 4);
}

extern "C" int64_t gc_binomial_deviance_scratch_bytes(int64_t C, int m) {
    return 2 * dev_align(C * m * 4) + dev_align(3 * (1 << kDevDepth) * 4) + 4 * dev_align((int64_t)m * 4 + 16) + 512 +
           spec_bytes(C, m);
}

// out0[j] (and out1[j]) = sequential float32 sum of column j of a0 (a1): piece-parallel when m <= kSpecMaxCols
static int colsum_axis0(const float *a0, const float *a1, int64_t C, int m, float *out0, float *out1, char *spec,
                        cudaStream_t st) {
    const unsigned n_arr = a1 ? 2 : 1;
    if (m > kSpecMaxCols) {
        const int cb = m <= kColThreads ? kColThreads : 64;      // columns per CTA of the sequential column sums
        colsum_sequential_kernel<<<dim3((unsigned)ceil_div(m, cb), n_arr), kColThreads, 0, st>>>(a0, a1, C, m, cb, out0, out1);
        return check_launch("colsum_sequential_kernel");
    }
    const int64_t P = spec_pieces(C);
    double *sums = (double *)spec;
    SpecSummary *summ = (SpecSummary *)(spec + 2 * dev_align(P * m * 8));
    int32_t *replays = (int32_t *)(spec + 2 * dev_align(P * m * 8) + 2 * dev_align(P * m * (int64_t)sizeof(SpecSummary)));
    colsum_spec_sums_kernel<<<dim3((unsigned)P, n_arr), kSpecPiece, 0, st>>>(a0, a1, C, m, P, sums);
    if (int rc = check_launch("colsum_spec_sums_kernel")) return rc;
    colsum_spec_prefix_kernel<<<dim3((unsigned)m, n_arr), 32, 0, st>>>(sums, m, P);
    if (int rc = check_launch("colsum_spec_prefix_kernel")) return rc;
    colsum_spec_summary_kernel<<<dim3((unsigned)ceil_div<int64_t>(P * m, 128), n_arr), 128, 0, st>>>(a0, a1, C, m, P, sums, summ);
    if (int rc = check_launch("colsum_spec_summary_kernel")) return rc;
    colsum_spec_chain_kernel<<<dim3((unsigned)m, n_arr), 32, 0, st>>>(a0, a1, C, m, P, summ, out0, out1, replays);
    return check_launch("colsum_spec_chain_kernel");
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
    float *right = (float *)p;     p += dev_align((int64_t)m * 4 + 16) + 256;
    char *spec = p;
    const int64_t n_all = C * (int64_t)m;
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
        if (int rc = colsum_axis0(X, nullptr, C, m, colsum, nullptr, spec, st)) return rc;
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
        if (int rc = colsum_axis0(L, R, C, m, left, right, spec, st)) return rc;
    }
    deviance_finish_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(left, right, m, out);
    return check_launch("deviance_finish_kernel");
}

extern "C" int gc_sum_axis0_f32(const float *a, int64_t C, int m, float *out, int32_t *replays, void *scratch, void *stream) {
    GC_REQUIRE(a && out && scratch, "null pointer");
    GC_REQUIRE(C >= 1 && m >= 2, "bad shape (m == 1 is numpy's pairwise sum: gc_binomial_deviance handles it)");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char *spec = (char *)scratch;
    if (int rc = colsum_axis0(a, nullptr, C, m, out, nullptr, spec, st)) return rc;
    if (replays && m <= kSpecMaxCols) {
        const int64_t P = spec_pieces(C);
        const int32_t *src = (const int32_t *)(spec + 2 * dev_align(P * m * 8) + 2 * dev_align(P * m * (int64_t)sizeof(SpecSummary)));
        GC_CUDA(cudaMemcpyAsync(replays, src, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}
