// GeneClust-fast gene k-means kernels: replace scikit-learn's Cython/OpenMP/BLAS kernels reached from
// MiniBatchKMeans.fit_predict (scGeneClust/tl/cluster.py:66-70) and paired_distances (:101).
// Compiled with -fmad=false; where sklearn's order of float32 operations is known it is reproduced.
#include "common.cuh"

#include <cstdlib>

namespace gc {

constexpr int kDMax = 64;        // feature dimension bound (GeneClust uses 50 PCs)

// ---------------------------------------------------------------------------------------------------------
// assignment: one WARP per sample; lane l scores centres l, l+32, ... (centres staged in shared memory with a padded
// row stride, so the 32 lanes read 32 different banks), then a warp arg-min with first-index ties.
// labels = argmin_j (|c_j|^2 - 2 x.c_j), strict '<' => first index on ties (_k_means_lloyd.pyx:196-211)
// sqdist = sum (x-c)^2 in the 4-way unrolled order of _euclidean_dense_dense (_k_means_common.pyx:16-43)
// The float32 operation order of every (sample, centre) score is the one of the original one-thread-per-sample
// kernel (two interleaved fma chains, one fma with the norm): results are bit-identical to it.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAssignWarps = 8;
__global__ void __launch_bounds__(kAssignWarps * 32)
kmeans_assign_kernel(const float *__restrict__ X, int64_t ldx, const int32_t *__restrict__ gather, int64_t n,
                     const float *__restrict__ centers, int K, int d, int ds /* padded row stride, odd */,
                     int32_t *__restrict__ labels, float *__restrict__ sqdist) {
    extern __shared__ float smem[];
    float *sc = smem;             // K * ds
    float *sn = smem + K * ds;    // K
    for (int i = threadIdx.x; i < K * d; i += blockDim.x) sc[(i / d) * ds + (i % d)] = centers[i];
    __syncthreads();
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        float s = 0.f;
        for (int f = 0; f < d; ++f) s = fmaf(sc[j * ds + f], sc[j * ds + f], s);
        sn[j] = s;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = (int64_t)blockIdx.x * kAssignWarps + warp; s < n; s += (int64_t)gridDim.x * kAssignWarps) {
        const int64_t row = gather ? gather[s] : s;
        const float *xp = X + row * ldx;
        float x[kDMax];
#pragma unroll
        for (int f = 0; f < kDMax; ++f) x[f] = f < d ? __ldg(xp + f) : 0.f;   // same address in all lanes: broadcast

        float best = 0.f;
        int lab = -1;
        for (int j = lane; j < K; j += 32) {
            const float *c = sc + j * ds;
            float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
            for (int f = 0; f < kDMax; f += 2) {
                if (f < d) dot0 = fmaf(x[f], c[f], dot0);
                if (f + 1 < d) dot1 = fmaf(x[f + 1], c[f + 1], dot1);
            }
            const float pd = fmaf(-2.0f, dot0 + dot1, sn[j]);
            if (lab < 0 || pd < best) { best = pd; lab = j; }     // ascending j within the lane: first index on ties
        }
        // warp arg-min, ties -> smaller index (lanes without a centre carry lab = -1)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ol = __shfl_xor_sync(0xffffffffu, lab, o);
            if (ol >= 0 && (lab < 0 || ob < best || (ob == best && ol < lab))) { best = ob; lab = ol; }
        }
        if (lane == 0) {
            labels[s] = lab;
            if (sqdist != nullptr) {
                const float *c = sc + lab * ds;
                float result = 0.f;
                const int n4 = d / 4;
#pragma unroll
                for (int i = 0; i < kDMax / 4; ++i) {
                    if (i < n4) {
                        const float t0 = x[4 * i] - c[4 * i], t1 = x[4 * i + 1] - c[4 * i + 1];
                        const float t2 = x[4 * i + 2] - c[4 * i + 2], t3 = x[4 * i + 3] - c[4 * i + 3];
                        result = result + (((t0 * t0 + t1 * t1) + t2 * t2) + t3 * t3);
                    }
                }
#pragma unroll
                for (int f = 0; f < kDMax; ++f) {
                    if (f >= 4 * n4 && f < d) { const float t = x[f] - c[f]; result = result + t * t; }
                }
                sqdist[s] = result;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// minibatch centre update: one warp per cluster, lane l owns features l and l + 32; the batch labels are scanned 32
// at a time (coalesced load + ballot) and the matching samples are accumulated in batch order
// (_k_means_minibatch.pyx:60-108 with unit sample weights: sequential float32 accumulation in sample order)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
kmeans_minibatch_update_kernel(const float *__restrict__ X, int64_t ldx, const int32_t *__restrict__ gather,
                               int64_t n, const int32_t *__restrict__ labels,
                               const float *__restrict__ centers_old, float *__restrict__ centers_new,
                               float *__restrict__ weight_sums, int K, int d) {
    extern __shared__ int32_t hits[];          // rows of this cluster's samples, in batch order (<= n entries)
    const int k = blockIdx.x, lane = threadIdx.x;
    const int f0 = lane, f1 = lane + 32;
    const float w_old = weight_sums[k];
    float acc0 = f0 < d ? centers_old[k * d + f0] * w_old : 0.f;
    float acc1 = f1 < d ? centers_old[k * d + f1] * w_old : 0.f;
    // pass 1: compact the matching batch positions (independent coalesced loads, unrolled so their latencies overlap)
    int cnt = 0;
#pragma unroll 4
    for (int64_t s0 = 0; s0 < n; s0 += 32) {
        const int64_t s = s0 + lane;
        const bool hit = s < n && __ldg(labels + s) == k;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) hits[cnt + __popc(m & ((1u << lane) - 1u))] = gather ? __ldg(gather + s) : (int32_t)s;
        cnt += __popc(m);
    }
    __syncwarp();
    // pass 2: sequential float32 accumulation in batch order; the row loads of four samples are issued together
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
        const int64_t r0 = hits[j], r1 = hits[j + 1], r2 = hits[j + 2], r3 = hits[j + 3];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        if (f0 < d) { a0 = X[r0 * ldx + f0]; a1 = X[r1 * ldx + f0]; a2 = X[r2 * ldx + f0]; a3 = X[r3 * ldx + f0]; }
        if (f1 < d) { b0 = X[r0 * ldx + f1]; b1 = X[r1 * ldx + f1]; b2 = X[r2 * ldx + f1]; b3 = X[r3 * ldx + f1]; }
        acc0 = acc0 + a0; acc0 = acc0 + a1; acc0 = acc0 + a2; acc0 = acc0 + a3;
        acc1 = acc1 + b0; acc1 = acc1 + b1; acc1 = acc1 + b2; acc1 = acc1 + b3;
    }
    for (; j < cnt; ++j) {
        const int64_t r = hits[j];
        if (f0 < d) acc0 = acc0 + X[r * ldx + f0];
        if (f1 < d) acc1 = acc1 + X[r * ldx + f1];
    }
    if (cnt > 0) {
        const float w_new = w_old + (float)cnt;      // = w_old + 1 + 1 + ... exactly: counts stay far below 2^24
        const float alpha = 1.0f / w_new;
        if (f0 < d) centers_new[k * d + f0] = acc0 * alpha;
        if (f1 < d) centers_new[k * d + f1] = acc1 * alpha;
        if (lane == 0) weight_sums[k] = w_new;
    } else {
        if (f0 < d) centers_new[k * d + f0] = centers_old[k * d + f0];
        if (f1 < d) centers_new[k * d + f1] = centers_old[k * d + f1];
    }
}

// ---------------------------------------------------------------------------------------------------------
// squared distances to candidate rows in float64, rounded to float32, clamped at 0
// (metrics/pairwise.py::_euclidean_distances_upcast: d = -2*x.y ; d += |x|^2 ; d += |y|^2 ; astype(float32))
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqdist_upcast(const float *__restrict__ xc, double xx, const float *__restrict__ y,
                                               int d) {
    double dot = 0.0, yy = 0.0;
    for (int f = 0; f < d; ++f) {
        const double yv = (double)y[f];
        dot = fma((double)xc[f], yv, dot);
        yy = fma(yv, yv, yy);
    }
    double v = -2.0 * dot;
    v = v + xx;
    v = v + yy;
    const float r = (float)v;
    return r > 0.f ? r : 0.f;
}

__global__ void sqdist_to_rows_kernel(const float *__restrict__ X, int64_t ldx, int64_t n, int d,
                                      const int32_t *__restrict__ cand, int ncand, float *__restrict__ out) {
    __shared__ float sc[kDMax];
    __shared__ double sxx;
    const int c = blockIdx.y;
    const float *xc = X + (int64_t)cand[c] * ldx;
    if (threadIdx.x < d) sc[threadIdx.x] = xc[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        double xx = 0.0;
        for (int f = 0; f < d; ++f) xx = fma((double)sc[f], (double)sc[f], xx);
        sxx = xx;
    }
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[(int64_t)c * n + i] = sqdist_upcast(sc, sxx, X + i * ldx, d);
}

// ---------------------------------------------------------------------------------------------------------
// k-means++ (cluster/_kmeans.py:180-275), device resident: all K-1 greedy rounds are enqueued back to back,
// the host only supplies the first centre id and the uniforms it drew from RandomState (their number does
// not depend on the data).  State lives in `ws` (see gc_kmeanspp_scratch_bytes).
// ---------------------------------------------------------------------------------------------------------
struct KppState {
    float current_pot;
    int32_t best;      // index into cand[] of the winning trial of the previous round
    int32_t pad[2];
};

// round step A (1 CTA): adopt the previous winner, sequential float32 cumsum of closest_dist_sq (np.cumsum),
// searchsorted(rand_vals) -> candidate ids.  The cumsum is the order-exact sequential one: chunks are staged
// through shared memory (coalesced loads/stores by all threads), one thread runs the dependent add chain.
constexpr int kPickChunk = 8192;
__global__ void __launch_bounds__(1024)
kpp_pick_kernel(KppState *st, float *__restrict__ closest, const float *__restrict__ trial_dist, int64_t n,
                const double *__restrict__ uniforms, int n_trials, int32_t *__restrict__ cand,
                float *__restrict__ cumsum, int32_t *__restrict__ indices_out, int round) {
    __shared__ float buf[kPickChunk];
    __shared__ float carry;
    // closest <- trial_dist[best] (np.minimum result of the winning candidate), round >= 2 only
    const bool adopt = round >= 2;
    const float *src = adopt ? trial_dist + (int64_t)st->best * n : closest;
    if (adopt && threadIdx.x == 0) indices_out[round - 1] = cand[st->best];
    if (threadIdx.x == 0) carry = 0.f;
    for (int64_t c0 = 0; c0 < n; c0 += kPickChunk) {
        const int cn = (int)min((int64_t)kPickChunk, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += blockDim.x) {
            const float v = src[c0 + i];
            buf[i] = v;
            if (adopt) closest[c0 + i] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float acc = carry;
            for (int i = 0; i < cn; ++i) { acc = acc + buf[i]; buf[i] = acc; }
            carry = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += blockDim.x) cumsum[c0 + i] = buf[i];
    }
    __syncthreads();
    if (threadIdx.x < n_trials) {
        const double rv = uniforms[(int64_t)(round - 1) * n_trials + threadIdx.x] * (double)st->current_pot;
        // np.searchsorted(cumsum, rv, side='left'): first index with cumsum[idx] >= rv
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((double)cumsum[mid] < rv) lo = mid + 1; else hi = mid;
        }
        if (lo > n - 1) lo = n - 1;  // np.clip(candidate_ids, None, n - 1)
        cand[threadIdx.x] = (int32_t)lo;
    }
}

// round step B: trial_dist[c][i] = min(closest[i], dist(X[cand[c]], X[i])); per-CTA partial potentials (fp64)
constexpr int kKppThreads = 256;
__global__ void __launch_bounds__(kKppThreads)
kpp_trial_kernel(const float *__restrict__ X, int64_t ldx, int64_t n, int d, const int32_t *__restrict__ cand,
                 const float *__restrict__ closest, float *__restrict__ trial_dist, double *__restrict__ partial) {
    __shared__ float sc[kDMax];
    __shared__ double sxx;
    __shared__ double red[kKppThreads];
    const int c = blockIdx.y;
    const float *xc = X + (int64_t)cand[c] * ldx;
    if (threadIdx.x < d) sc[threadIdx.x] = xc[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        double xx = 0.0;
        for (int f = 0; f < d; ++f) xx = fma((double)sc[f], (double)sc[f], xx);
        sxx = xx;
    }
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < n) {
        const float dd = fminf(closest[i], sqdist_upcast(sc, sxx, X + i * ldx, d));
        trial_dist[(int64_t)c * n + i] = dd;
        v = (double)dd;
    }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = kKppThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(int64_t)c * gridDim.x + blockIdx.x] = red[0];
}

// Once per k-means++ call: Xt[f][i] = (double)X[i][f] (feature-major, so a warp reads consecutive samples) and
// yy[i] = sum_f fma(y_f, y_f, .) in the order of sqdist_upcast.  The round kernel then spends 50 DFMAs per
// (candidate, sample) instead of 100 DFMAs + 100 float->double conversions; the distances are bit-identical.
__global__ void kpp_prep_kernel(const float *__restrict__ X, int64_t ldx, int64_t n, int d, double *__restrict__ Xt,
                                double *__restrict__ yy) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *y = X + i * ldx;
    double s = 0.0;
    for (int f = 0; f < d; ++f) {
        const double yv = (double)y[f];
        Xt[(int64_t)f * n + i] = yv;
        s = fma(yv, yv, s);
    }
    yy[i] = s;
}
// same value as sqdist_upcast(xc, xx, X + i*ldx, d), from the prepared arrays
__device__ __forceinline__ float sqdist_prepared(const float *__restrict__ xc, double xx, const double *__restrict__ Xt,
                                                 const double *__restrict__ yy, int64_t n, int64_t i, int d) {
    double dot = 0.0;
    for (int f = 0; f < d; ++f) dot = fma((double)xc[f], Xt[(int64_t)f * n + i], dot);
    double v = -2.0 * dot;
    v = v + xx;
    v = v + yy[i];
    const float r = (float)v;
    return r > 0.f ? r : 0.f;
}

// One greedy round as ONE kernel: every CTA computes its trial-distance tile (as kpp_trial_kernel); the LAST CTA to
// finish (atomic ticket) then runs the tail of the round: candidates_pot = fixed-order sum of the partials, first
// argmin, new current_pot and - unless this was the last round - the head of the next one (what kpp_pick_kernel does
// for round 1): adopt the winner's distances, sequential float32 cumsum staged through shared memory, searchsorted of
// the next uniforms inside shared memory.
constexpr int kRoundChunk = 6144;     // floats of cumsum staged per pass of the tail (24 KB)
constexpr int kSpecP = 64;            // pieces of the speculative running sum (see the tail of kpp_round_kernel)

// exact sequential float32 sum of q[0..len) started from acc; loads batched ahead of the dependent adds
__device__ __forceinline__ float chain_sum(const float *q, int len, float acc) {
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = q[i + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = acc + v[k];
    }
    for (; i < len; ++i) acc = acc + q[i];
    return acc;
}
// Returns (to every thread of the CTA) whether this CTA ran the tail of the round.
__device__ __forceinline__ bool
kpp_round_step(KppState *st, const float *__restrict__ X, int64_t ldx, const double *__restrict__ Xt,
               const double *__restrict__ yy, int64_t n, int d, int32_t *cand,
               float *closest, float *trial_dist, double *partial, const double *__restrict__ uniforms, int n_trials,
               int32_t *__restrict__ indices_out, int round, int last_round, unsigned int *ticket) {
    __shared__ float sc[kDMax];
    __shared__ double sxx;
    __shared__ double red[kKppThreads];
    __shared__ __align__(16) float buf[kRoundChunk + 2 * kSpecP + 32];    // + zero padding up to kSpecP pieces of odd length
    __shared__ float sp_sum[kSpecP], sp_d0[kSpecP], sp_d1[kSpecP], sp_start[kSpecP], sp_end[kSpecP];
    __shared__ int sp_e[kSpecP];
    __shared__ float carry;
    __shared__ int s_best, s_is_last;
    __shared__ float s_pot;
    __shared__ double s_rv[32];
    __shared__ int64_t s_found[32];
    const int c = blockIdx.y;
    {
        // cand[] and closest[] were written by the tail CTA of the previous round - in the persistent launch that is
        // another CTA of the SAME kernel, so they are read through L2 (ld.cg), never from a stale L1 line
        const float *xc = X + (int64_t)__ldcg(cand + c) * ldx;
        if (threadIdx.x < d) sc[threadIdx.x] = xc[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            double xx = 0.0;
            for (int f = 0; f < d; ++f) xx = fma((double)sc[f], (double)sc[f], xx);
            sxx = xx;
        }
        __syncthreads();
        const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        double v = 0.0;
        if (i < n) {
            const float dd = fminf(__ldcg(closest + i), sqdist_prepared(sc, sxx, Xt, yy, n, i, d));
            trial_dist[(int64_t)c * n + i] = dd;
            v = (double)dd;
        }
        red[threadIdx.x] = v;
        __syncthreads();
        for (int s = kKppThreads / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(int64_t)c * gridDim.x + blockIdx.x] = red[0];
    }
    // ---- last CTA to finish runs the sequential tail
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int t = atomicAdd(ticket, 1u);
        s_is_last = (t == total - 1);
        if (s_is_last) *ticket = 0;              // re-arm for the next launch (stream ordered)
    }
    __syncthreads();
    if (!s_is_last) return false;
    __threadfence();
    const int n_blocks = gridDim.x;
    // choose: potentials in a fixed order, first argmin
    for (int t = threadIdx.x >> 5; t < n_trials; t += kKppThreads / 32) {   // one warp per trial, fixed reduction tree
        double s = 0.0;
        for (int b = threadIdx.x & 31; b < n_blocks; b += 32) s += __ldcg(partial + (int64_t)t * n_blocks + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[t] = (double)(float)s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best_pot = 0.f;
        int best = 0;
        for (int t = 0; t < n_trials; ++t) {
            const float pot = (float)red[t];
            if (t == 0 || pot < best_pot) { best_pot = pot; best = t; }
        }
        st->current_pot = best_pot;
        st->best = best;
        s_best = best;
        s_pot = best_pot;
        indices_out[round] = __ldcg(cand + best);     // centre chosen in this round
        carry = 0.f;
    }
    __syncthreads();
    if (last_round) return true;
    // pick for round + 1: closest <- trial_dist[best]; cumsum; searchsorted
    const float *src = trial_dist + (int64_t)s_best * n;
    if (threadIdx.x < n_trials) {
        s_rv[threadIdx.x] = uniforms[(int64_t)round * n_trials + threadIdx.x] * (double)s_pot;
        s_found[threadIdx.x] = -1;
    }
    for (int64_t c0 = 0; c0 < n; c0 += kRoundChunk) {
        const int cn = (int)min((int64_t)kRoundChunk, n - c0);
        __syncthreads();
        // adopt: closest <- trial_dist[best] for this chunk, staged in shared memory (loads issued eight at a time)
        {
            const int cn4 = cn >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src + c0);      // n_trials * n floats, 256-B aligned base;
            float4 *dst4 = reinterpret_cast<float4 *>(closest + c0);             // c0 is a multiple of kRoundChunk
            float4 *b4w = reinterpret_cast<float4 *>(buf);
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(src4) | reinterpret_cast<uintptr_t>(dst4)) & 15) == 0;
            if (vec_ok) {
                for (int i0 = threadIdx.x; i0 < cn4; i0 += 4 * blockDim.x) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i0 + u * blockDim.x < cn4) v[u] = __ldcg(src4 + i0 + u * blockDim.x);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * blockDim.x < cn4) { b4w[i0 + u * blockDim.x] = v[u]; dst4[i0 + u * blockDim.x] = v[u]; }
                }
                for (int i = 4 * cn4 + threadIdx.x; i < cn; i += blockDim.x) {
                    const float v = __ldcg(src + c0 + i);
                    buf[i] = v;
                    closest[c0 + i] = v;
                }
            } else {
                for (int i = threadIdx.x; i < cn; i += blockDim.x) {
                    const float v = __ldcg(src + c0 + i);
                    buf[i] = v;
                    closest[c0 + i] = v;
                }
            }
            for (int i = cn + threadIdx.x; i < cn + 2 * kSpecP + 32; i += blockDim.x) buf[i] = 0.f;   // x + 0 = x
        }
        __syncthreads();
        // np.cumsum(closest) is a dependent float32 add chain (4 cycles per element, 6 000 elements) and only its
        // crossings of the n_trials targets are needed.  Speculative exact evaluation (scripts/proto/spec_cumsum.py is
        // the CPU prototype, checked bit for bit against np.cumsum): the chunk is cut into kSpecP pieces of odd length
        // (bank-conflict free).  While the running sum stays inside one binade [2^e, 2^(e+1)), every add rounds to a
        // multiple of u = ulp and only round-half-to-even looks at the sum itself - at its last mantissa bit.  So the
        // amount a piece adds depends on the start value only through (e, parity): each thread runs its piece from the
        // two SURROGATE starts 2^e and 2^e + u (e guessed from approximate prefix sums) and records the increments
        // D[parity]; one thread then walks the pieces, end = start + D[parity(start)], accepting a piece iff the
        // start really has exponent e and the end still has it (monotone sums: no add left the binade; the
        // surrogates are <= the true start, so they did not leave it either), and summing the piece for real
        // otherwise (~6 of 64 pieces: the ones in which the sum crosses a power of two).  Finally each target is
        // located inside its piece by a real chain from the exact piece start.
        const int len = ((cn + kSpecP - 1) / kSpecP) | 1;
        if (threadIdx.x < kSpecP) {
            const float *q = buf + threadIdx.x * len;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int i = 0;
            for (; i + 4 <= len; i += 4) { s0 += q[i]; s1 += q[i + 1]; s2 += q[i + 2]; s3 += q[i + 3]; }
            for (; i < len; ++i) s0 += q[i];
            sp_sum[threadIdx.x] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        if (threadIdx.x < kSpecP) {
            const int p = threadIdx.x;
            const float *q = buf + p * len;
            if (p == 0) {
                sp_start[0] = carry;
                sp_end[0] = chain_sum(q, len, carry);
            } else {
                float a0 = carry, a1 = 0.f, a2 = 0.f, a3 = 0.f;         // approximate start (any order)
                int j = 0;
                for (; j + 4 <= p; j += 4) { a0 += sp_sum[j]; a1 += sp_sum[j + 1]; a2 += sp_sum[j + 2]; a3 += sp_sum[j + 3]; }
                for (; j < p; ++j) a0 += sp_sum[j];
                const int e = (int)((__float_as_uint((a0 + a1) + (a2 + a3)) >> 23) & 0xffu);
                const float lo0 = __uint_as_float((uint32_t)e << 23), lo1 = __uint_as_float(((uint32_t)e << 23) + 1u);
                float x0 = lo0, x1 = lo1;
                int i = 0;
                for (; i + 4 <= len; i += 4) {
                    const float v0 = q[i], v1 = q[i + 1], v2 = q[i + 2], v3 = q[i + 3];
                    x0 = x0 + v0; x1 = x1 + v0; x0 = x0 + v1; x1 = x1 + v1;
                    x0 = x0 + v2; x1 = x1 + v2; x0 = x0 + v3; x1 = x1 + v3;
                }
                for (; i < len; ++i) { const float v = q[i]; x0 = x0 + v; x1 = x1 + v; }
                sp_d0[p] = x0 - lo0;
                sp_d1[p] = x1 - lo1;
                sp_e[p] = (e >= 1 && e <= 253) ? e : -1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float start = sp_end[0];
#pragma unroll 4
            for (int p = 1; p < kSpecP; ++p) {
                sp_start[p] = start;
                const uint32_t sb = __float_as_uint(start);
                const int es = (int)((sb >> 23) & 0xffu);
                float end = start + ((sb & 1u) ? sp_d1[p] : sp_d0[p]);
                const bool ok = es == sp_e[p] && (int)((__float_as_uint(end) >> 23) & 0xffu) == es;
                if (!ok) end = chain_sum(buf + p * len, len, start);
                sp_end[p] = end;
                start = end;
            }
            carry = start;
        }
        __syncthreads();
        if (threadIdx.x < n_trials && s_found[threadIdx.x] < 0) {
            // np.searchsorted(cumsum, rv, side='left') restricted to this chunk (earlier chunks are all < rv).  For a
            // float32 c:  (double)c >= rv  <=>  c >= rv rounded UP to float32
            const float rvf = __double2float_ru(s_rv[threadIdx.x]);
            if (sp_end[kSpecP - 1] >= rvf) {
                int lo = 0, hi = kSpecP - 1;                           // first piece whose end reaches the target
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (sp_end[mid] < rvf) lo = mid + 1; else hi = mid;
                }
                const float *q = buf + lo * len;
                float acc = sp_start[lo];
                int i = 0;
                for (; i + 8 <= len; i += 8) {                         // blocks of 8: one test per block on the chain
                    const float before = acc;
                    const float nxt = chain_sum(q + i, 8, acc);
                    if (nxt >= rvf) { acc = before; break; }
                    acc = nxt;
                }
                for (; i < len; ++i) { acc = acc + q[i]; if (acc >= rvf) break; }
                s_found[threadIdx.x] = c0 + lo * len + i;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < n_trials) {
        int64_t f = s_found[threadIdx.x];
        if (f < 0 || f > n - 1) f = n - 1;           // np.clip(candidate_ids, None, n - 1)
        cand[threadIdx.x] = (int32_t)f;
    }
    return true;
}

// one launch per round (fallback when the grid cannot be co-resident)
__global__ void __launch_bounds__(kKppThreads)
kpp_round_kernel(KppState *st, const float *__restrict__ X, int64_t ldx, const double *__restrict__ Xt,
                 const double *__restrict__ yy, int64_t n, int d, int32_t *cand,
                 float *closest, float *trial_dist, double *partial, const double *__restrict__ uniforms, int n_trials,
                 int32_t *__restrict__ indices_out, int round, int last_round, unsigned int *ticket) {
    kpp_round_step(st, X, ldx, Xt, yy, n, d, cand, closest, trial_dist, partial, uniforms, n_trials, indices_out, round,
                   last_round, ticket);
}

// ALL greedy rounds in one cooperative launch (every CTA is resident, so CTAs may wait for one another): per round the
// CTA that finishes its distance tile last runs the sequential tail and then publishes the round number; the other CTAs
// wait for it.  Saves the 198 launches and their inter-kernel gaps (~4 us each of the 23 us a round took).
__global__ void __launch_bounds__(kKppThreads)
kpp_persistent_kernel(KppState *st, const float *__restrict__ X, int64_t ldx, const double *__restrict__ Xt,
                      const double *__restrict__ yy, int64_t n, int d, int32_t *cand, float *closest, float *trial_dist,
                      double *partial, const double *__restrict__ uniforms, int n_trials,
                      int32_t *__restrict__ indices_out, int K, unsigned int *ticket, int *round_flag) {
    for (int round = 1; round < K; ++round) {
        const bool ran_tail = kpp_round_step(st, X, ldx, Xt, yy, n, d, cand, closest, trial_dist, partial, uniforms, n_trials,
                                             indices_out, round, round == K - 1 ? 1 : 0, ticket);
        if (round == K - 1) break;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (ran_tail) {
                __threadfence();                                  // the tail's global writes before the flag
                atomicExch(round_flag, round);
            } else {
                unsigned int spins = 0;
                while (atomicAdd(round_flag, 0) < round) {        // L2 read; the tail takes a few microseconds
                    __nanosleep(200);
                    if (++spins > (1u << 24)) __trap();
                }
                __threadfence();
            }
        }
        __syncthreads();
    }
}

// first centre: closest = dist(X[first]), current_pot = sum
__global__ void kpp_first_kernel(KppState *st, const double *__restrict__ partial, int n_blocks,
                                 int32_t *__restrict__ cand, int32_t first, int32_t *__restrict__ indices_out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; ++b) s += partial[b];
        st->current_pot = (float)s;
        st->best = 0;
        cand[0] = first;
        indices_out[0] = first;
    }
}

__global__ void kpp_set_cand_kernel(int32_t *cand, int32_t first) { cand[0] = first; }

__global__ void gather_rows_kernel(const float *__restrict__ X, int64_t ldx, const int32_t *__restrict__ idx,
                                   int n_rows, int d, float *__restrict__ out) {
    const int r = blockIdx.x;
    if (r < n_rows && threadIdx.x < d) out[r * d + threadIdx.x] = X[(int64_t)idx[r] * ldx + threadIdx.x];
}

// paired euclidean distance to the assigned centre (row_norms(X - Y), float32)
__global__ void paired_dist_kernel(const float *__restrict__ X, int64_t ldx, int64_t n, int d,
                                   const float *__restrict__ centers, const int32_t *__restrict__ labels,
                                   float *__restrict__ dist) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float *x = X + s * ldx, *c = centers + (int64_t)labels[s] * d;
    float acc = 0.f;
    for (int f = 0; f < d; ++f) { const float t = x[f] - c[f]; acc = acc + t * t; }
    dist[s] = sqrtf(acc);
}

// per-cluster min / max of non-negative float32 distances via integer atomics on the bit patterns (order
// preserving for values >= 0), then closeness = 1 - (d * scale + min_) with sklearn MinMaxScaler's float32 steps
__global__ void closeness_init_kernel(int32_t *scratch, int K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < K) { scratch[j] = 0x7f800000; scratch[K + j] = 0; }
}
__global__ void closeness_minmax_kernel(const float *__restrict__ dist, const int32_t *__restrict__ labels, int64_t n,
                                        int K, int32_t *scratch) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int b = __float_as_int(dist[s]), l = labels[s];
    atomicMin(&scratch[l], b);
    atomicMax(&scratch[K + l], b);
}
__global__ void closeness_apply_kernel(const float *__restrict__ dist, const int32_t *__restrict__ labels, int64_t n,
                                       int K, const int32_t *__restrict__ scratch, float *__restrict__ out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int l = labels[s];
    const float dmin = __int_as_float(scratch[l]), dmax = __int_as_float(scratch[K + l]);
    float range = dmax - dmin;
    if (range < 10.0f * 1.1920928955078125e-07f) range = 1.0f;   // _handle_zeros_in_scale
    const float scale = 1.0f / range;
    const float min_ = 0.0f - dmin * scale;
    float v = dist[s] * scale;
    v = v + min_;
    out[s] = 1.0f - v;
}

// Per-cluster size, first arg-max and first arg-min of a non-negative float32 value (the closeness in [0, 1]): what
// `groupby('cluster')[col].nlargest(1) / .nsmallest(1)` (keep='first') select in scGeneClust/tl/selection.py:57-59.
// Non-negative floats order like their bit patterns, so the extremes are integer atomics; the first index attaining an
// extreme is an atomicMin over the indices that hold it.  out: counts[K] | first_max[K] | first_min[K] (int32).
__global__ void cluster_extremes_init_kernel(int32_t *out, int32_t *ext, int K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < K) { out[j] = 0; out[K + j] = 0x7fffffff; out[2 * K + j] = 0x7fffffff; ext[j] = -1; ext[K + j] = 0x7fffffff; }
}
__global__ void cluster_extremes_value_kernel(const float *__restrict__ v, const int32_t *__restrict__ labels, int64_t n,
                                              int K, int32_t *out, int32_t *ext) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int l = labels[s], b = __float_as_int(v[s]);
    atomicAdd(&out[l], 1);
    atomicMax(&ext[l], b);
    atomicMin(&ext[K + l], b);
}
__global__ void cluster_extremes_index_kernel(const float *__restrict__ v, const int32_t *__restrict__ labels, int64_t n,
                                              int K, int32_t *out, const int32_t *__restrict__ ext) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int l = labels[s], b = __float_as_int(v[s]);
    if (b == ext[l]) atomicMin(&out[K + l], (int32_t)s);
    if (b == ext[K + l]) atomicMin(&out[2 * K + l], (int32_t)s);
}

// idx[i] = np.searchsorted(cdf, u[i], side='right') = first index with cdf[index] > u[i] (n if none): the mini-batch
// sampling of RandomState.choice(n, size, p=p) given its uniform draws (numpy mtrand.pyx)
__global__ void searchsorted_right_kernel(const double *__restrict__ cdf, int64_t n, const double *__restrict__ u, int m,
                                          int32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double x = u[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= x) lo = mid + 1; else hi = mid;
    }
    idx[i] = (int32_t)lo;
}

// random reassignment of starved centres (_mini_batch_step, cluster/_kmeans.py:1655-1682): centre clusters[r] <- the batch
// sample picks[r] (a position in the mini-batch, i.e. row idx[picks[r]] of X), its count <- min_count
__global__ void kmeans_reassign_kernel(const float *__restrict__ X, int64_t ldx, int d, const int32_t *__restrict__ idx,
                                       const int32_t *__restrict__ clusters, const int32_t *__restrict__ picks,
                                       float min_count, float *__restrict__ centers_new, float *__restrict__ counts) {
    const int r = blockIdx.x, c = clusters[r];
    const float *src = X + (int64_t)idx[picks[r]] * ldx;
    for (int f = threadIdx.x; f < d; f += blockDim.x) centers_new[(int64_t)c * d + f] = src[f];
    if (threadIdx.x == 0) counts[c] = min_count;
}

// deterministic float64 sum of a float32 vector (single CTA, fixed tree), result as float64 and float32
__global__ void __launch_bounds__(1024) sum_f32_kernel(const float *__restrict__ v, int64_t n, double *out) {
    __shared__ double red[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

}  // namespace gc

using namespace gc;

extern "C" int gc_kmeans_assign(const float *X, int64_t ldx, const int32_t *gather, int64_t n, const float *centers,
                                int K, int d, int32_t *labels, float *sqdist, void *stream) {
    GC_REQUIRE(X && centers && labels, "null pointer");
    GC_REQUIRE(d >= 1 && d <= kDMax, "d must be in [1, 64]");
    GC_REQUIRE(K >= 1 && n >= 1 && ldx >= d, "bad shape");
    const int ds = d | 1;                                            // odd row stride: conflict-free lane-per-centre reads
    const size_t smem = (size_t)(K * ds + K) * sizeof(float);
    GC_REQUIRE(smem <= 200 * 1024, "K*d too large for the shared-memory centre cache");
    static thread_local size_t configured[kMaxDevices] = {};
    const int dev_id = current_device();
    if (smem > 48 * 1024 && smem > configured[dev_id]) {
        GC_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_id] = smem;
    }
    // every CTA stages the centres once: a few samples per warp amortise that without starving the SMs
    const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(n, kAssignWarps), (int64_t)kNumSMs * 4));
    kmeans_assign_kernel<<<(unsigned)ctas, kAssignWarps * 32, smem, as_stream(stream)>>>(X, ldx, gather, n, centers, K, d, ds,
                                                                                     labels, sqdist);
    return check_launch("kmeans_assign_kernel");
}

extern "C" int gc_kmeans_minibatch_update(const float *X, int64_t ldx, const int32_t *gather, int64_t n,
                                          const int32_t *labels, const float *centers_old, float *centers_new,
                                          float *weight_sums, int K, int d, void *stream) {
    GC_REQUIRE(X && labels && centers_old && centers_new && weight_sums, "null pointer");
    GC_REQUIRE(d >= 1 && d <= kDMax && K >= 1 && n >= 1, "bad shape");
    const size_t hit_bytes = (size_t)n * 4;
    GC_REQUIRE(hit_bytes <= 200 * 1024, "mini-batch larger than 51200 samples");
    static thread_local size_t upd_configured[kMaxDevices] = {};
    const int upd_dev = current_device();
    if (hit_bytes > 48 * 1024 && hit_bytes > upd_configured[upd_dev]) {
        GC_CUDA(cudaFuncSetAttribute(kmeans_minibatch_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hit_bytes));
        upd_configured[upd_dev] = hit_bytes;
    }
    kmeans_minibatch_update_kernel<<<K, 32, hit_bytes, as_stream(stream)>>>(X, ldx, gather, n, labels, centers_old,
                                                                      centers_new, weight_sums, K, d);
    return check_launch("kmeans_minibatch_update_kernel");
}

extern "C" int gc_sqdist_to_rows(const float *X, int64_t ldx, int64_t n, int d, const int32_t *cand, int ncand,
                                 float *out, void *stream) {
    GC_REQUIRE(X && cand && out, "null pointer");
    GC_REQUIRE(d >= 1 && d <= kDMax && n >= 1 && ncand >= 1, "bad shape");
    dim3 grid((unsigned)ceil_div<int64_t>(n, 256), (unsigned)ncand);
    sqdist_to_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, ldx, n, d, cand, ncand, out);
    return check_launch("sqdist_to_rows_kernel");
}

extern "C" int gc_paired_dist(const float *X, int64_t ldx, int64_t n, int d, const float *centers,
                              const int32_t *labels, float *dist, void *stream) {
    GC_REQUIRE(X && centers && labels && dist, "null pointer");
    paired_dist_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(X, ldx, n, d, centers,
                                                                                          labels, dist);
    return check_launch("paired_dist_kernel");
}

extern "C" int gc_closeness(const float *dist, const int32_t *labels, int64_t n, int K, int32_t *scratch,
                            float *closeness, void *stream) {
    GC_REQUIRE(dist && labels && scratch && closeness && n >= 1 && K >= 1, "bad arguments");
    cudaStream_t st = as_stream(stream);
    closeness_init_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(scratch, K);
    if (int rc = check_launch("closeness_init_kernel")) return rc;
    closeness_minmax_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(dist, labels, n, K, scratch);
    if (int rc = check_launch("closeness_minmax_kernel")) return rc;
    closeness_apply_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(dist, labels, n, K, scratch, closeness);
    return check_launch("closeness_apply_kernel");
}

extern "C" int gc_searchsorted_right(const double *cdf, int64_t n, const double *u, int m, int32_t *idx, void *stream) {
    GC_REQUIRE(cdf && u && idx && n >= 1 && m >= 1, "bad arguments");
    searchsorted_right_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, as_stream(stream)>>>(cdf, n, u, m, idx);
    return check_launch("searchsorted_right_kernel");
}

extern "C" int gc_cluster_extremes(const float *value, const int32_t *labels, int64_t n, int K, int32_t *out,
                                   int32_t *scratch, void *stream) {
    GC_REQUIRE(value && labels && out && scratch && n >= 1 && K >= 1 && n < (1ll << 31), "bad arguments");
    cudaStream_t st = as_stream(stream);
    const unsigned nb = (unsigned)ceil_div<int64_t>(n, 256);
    cluster_extremes_init_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(out, scratch, K);
    if (int rc = check_launch("cluster_extremes_init_kernel")) return rc;
    cluster_extremes_value_kernel<<<nb, 256, 0, st>>>(value, labels, n, K, out, scratch);
    if (int rc = check_launch("cluster_extremes_value_kernel")) return rc;
    cluster_extremes_index_kernel<<<nb, 256, 0, st>>>(value, labels, n, K, out, scratch);
    return check_launch("cluster_extremes_index_kernel");
}

extern "C" int gc_kmeans_minibatch_step(const float *X, int64_t ldx, int64_t n, int d, int K, const double *cdf,
                                        const double *u, int batch, int32_t *idx, int32_t *labels_b, float *sq_b,
                                        const float *centers, float *centers_new, float *counts, double *inertia,
                                        void *stream) {
    // one `_mini_batch_step` of MiniBatchKMeans.fit (cluster/_kmeans.py:2169-2196) as ONE call: the batch indices from the
    // host's uniform draws, assignment to the current centres, batch inertia, running-mean update of the centres
    if (int rc = gc_searchsorted_right(cdf, n, u, batch, idx, stream)) return rc;
    if (int rc = gc_kmeans_assign(X, ldx, idx, batch, centers, K, d, labels_b, sq_b, stream)) return rc;
    if (int rc = gc_sum_f32(sq_b, batch, inertia, stream)) return rc;
    return gc_kmeans_minibatch_update(X, ldx, idx, batch, labels_b, centers, centers_new, counts, K, d, stream);
}

extern "C" int gc_kmeans_reassign(const float *X, int64_t ldx, int d, const int32_t *idx, const int32_t *clusters,
                                  const int32_t *picks, int n_reassign, float min_count, float *centers_new, float *counts,
                                  void *stream) {
    GC_REQUIRE(X && idx && clusters && picks && centers_new && counts && d >= 1, "bad arguments");
    if (n_reassign <= 0) return 0;
    kmeans_reassign_kernel<<<n_reassign, 64, 0, as_stream(stream)>>>(X, ldx, d, idx, clusters, picks, min_count, centers_new, counts);
    return check_launch("kmeans_reassign_kernel");
}

extern "C" int gc_sum_f32(const float *v, int64_t n, double *out, void *stream) {
    GC_REQUIRE(v && out && n >= 1, "bad arguments");
    sum_f32_kernel<<<1, 1024, 0, as_stream(stream)>>>(v, n, out);
    return check_launch("sum_f32_kernel");
}

extern "C" int64_t gc_kmeanspp_scratch_bytes(int64_t n, int n_trials) {
    const int64_t n_blocks = ceil_div<int64_t>(n, kKppThreads);
    // state | cand[n_trials] | closest[n] | cumsum[n] | trial_dist[n_trials*n] | partial[n_trials*n_blocks]
    return 256 + 256 + (int64_t)sizeof(float) * (2 * n + (int64_t)n_trials * n) + 256 * 3 +
           (int64_t)sizeof(double) * n_trials * n_blocks + 256 + 256 /* ticket */ +
           (int64_t)sizeof(double) * (n * (kDMax + 1)) + 512 /* feature-major float64 copy + squared norms */;
}

extern "C" int gc_kmeanspp_init(const float *X, int64_t ldx, int64_t n, int d, int K, int32_t first_center,
                                const double *uniforms, int n_trials, void *ws, float *centers_out,
                                int32_t *indices_out, void *stream) {
    GC_REQUIRE(X && uniforms && ws && centers_out && indices_out, "null pointer");
    GC_REQUIRE(d >= 1 && d <= kDMax && K >= 1 && n >= K, "bad shape");
    GC_REQUIRE(n_trials >= 1 && n_trials <= 32, "n_local_trials must be in [1, 32]");
    GC_REQUIRE(first_center >= 0 && first_center < n, "first centre out of range");
    cudaStream_t st = as_stream(stream);
    const int n_blocks = (int)ceil_div<int64_t>(n, kKppThreads);
    auto align = [](char *p) { return (char *)(((uintptr_t)p + 255) & ~(uintptr_t)255); };
    char *p = align((char *)ws);
    KppState *state = (KppState *)p;            p = align(p + sizeof(KppState));
    int32_t *cand = (int32_t *)p;               p = align(p + sizeof(int32_t) * 32);
    float *closest = (float *)p;                p = align(p + sizeof(float) * n);
    float *cumsum = (float *)p;                 p = align(p + sizeof(float) * n);
    float *trial = (float *)p;                  p = align(p + sizeof(float) * n * n_trials);
    double *partial = (double *)p;

    // first centre: closest_dist_sq = dist(centre0, X) (no min), current_pot = its sum
    kpp_set_cand_kernel<<<1, 1, 0, st>>>(cand, first_center);
    if (int rc = check_launch("kpp_set_cand_kernel")) return rc;
    {
        dim3 grid((unsigned)n_blocks, 1);
        // closest starts at +inf so that min(closest, d) = d
        GC_CUDA(cudaMemsetAsync(closest, 0x7f, sizeof(float) * n, st));  // 0x7f7f7f7f = 3.39e38 > any distance
        kpp_trial_kernel<<<grid, kKppThreads, 0, st>>>(X, ldx, n, d, cand, closest, trial, partial);
        if (int rc = check_launch("kpp_trial_kernel")) return rc;
        kpp_first_kernel<<<1, 32, 0, st>>>(state, partial, n_blocks, cand, first_center, indices_out);
        if (int rc = check_launch("kpp_first_kernel")) return rc;
        GC_CUDA(cudaMemcpyAsync(closest, trial, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    }
    unsigned int *ticket = (unsigned int *)align((char *)(partial + (int64_t)n_trials * n_blocks));
    GC_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
    double *Xt = (double *)align((char *)(ticket + 64));
    double *yy = (double *)align((char *)(Xt + (int64_t)d * n));
    kpp_prep_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(X, ldx, n, d, Xt, yy);
    if (int rc = check_launch("kpp_prep_kernel")) return rc;
    if (K > 1) {
        // head of round 1 (no winner to adopt yet), then one launch per round
        kpp_pick_kernel<<<1, 1024, 0, st>>>(state, closest, trial, n, uniforms, n_trials, cand, cumsum, indices_out, 1);
        if (int rc = check_launch("kpp_pick_kernel")) return rc;
        dim3 grid((unsigned)n_blocks, (unsigned)n_trials);
        int per_sm = 0, coop = 0, dev = 0;
        GC_CUDA(cudaGetDevice(&dev));
        GC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        GC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kpp_persistent_kernel, kKppThreads, 0));
        static const bool no_persistent = getenv("GC_KPP_PER_ROUND") != nullptr;       // A/B switch for the tests
        if (coop && !no_persistent && (int64_t)per_sm * kNumSMs >= (int64_t)n_blocks * n_trials && K > 2) {
            int *round_flag = (int *)(ticket + 16);
            GC_CUDA(cudaMemsetAsync(round_flag, 0, sizeof(int), st));
            int last_dummy = 0;
            (void)last_dummy;
            void *args[] = {(void *)&state, (void *)&X, (void *)&ldx, (void *)&Xt, (void *)&yy, (void *)&n, (void *)&d,
                            (void *)&cand, (void *)&closest, (void *)&trial, (void *)&partial, (void *)&uniforms,
                            (void *)&n_trials, (void *)&indices_out, (void *)&K, (void *)&ticket, (void *)&round_flag};
            GC_CUDA(cudaLaunchCooperativeKernel((const void *)kpp_persistent_kernel, grid, dim3(kKppThreads), args, 0, st));
            if (int rc = check_launch("kpp_persistent_kernel")) return rc;
        } else {
            for (int round = 1; round < K; ++round) {
                kpp_round_kernel<<<grid, kKppThreads, 0, st>>>(state, X, ldx, Xt, yy, n, d, cand, closest, trial, partial,
                                                               uniforms, n_trials, indices_out, round, round == K - 1 ? 1 : 0,
                                                               ticket);
                if (int rc = check_launch("kpp_round_kernel")) return rc;
            }
        }
    }
    gather_rows_kernel<<<K, kDMax, 0, st>>>(X, ldx, indices_out, K, d, centers_out);
    return check_launch("gather_rows_kernel");
}
