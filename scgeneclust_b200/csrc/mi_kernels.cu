// kNN mutual-information kernels (GeneClust-ps): input preparation, KSG (continuous-continuous) and Ross
// (continuous-discrete) estimators, exact fp64, integer neighbour counts bit-identical to scikit-learn 1.9.0
// (feature_selection/_mutual_info.py:20-153, 294-315) as called from scGeneClust/tl/information.py:24,63,96.
//
// This translation unit is compiled with -fmad=false: every floating-point operation below must round exactly
// like the numpy/sklearn expression it restates (no FMA contraction).
#include <cstdlib>

#include "common.cuh"
#include "np_sum.cuh"

namespace gc {

// nextafter(eps, 0) for eps >= 0 (np.nextafter(radius, 0), _mutual_info.py:61,131)
__device__ __forceinline__ double next_toward_zero(double eps) {
    long long b = __double_as_longlong(eps);
    return b > 0 ? __longlong_as_double(b - 1) : 0.0;
}

// keep the three smallest values seen so far, m0 <= m1 <= m2
__device__ __forceinline__ void insert3(double d, double &m0, double &m1, double &m2) {
    const bool c2 = d < m2, c1 = d < m1, c0 = d < m0;
    m2 = c1 ? m1 : (c2 ? d : m2);
    m1 = c0 ? m0 : (c1 ? d : m1);
    m0 = c0 ? d : m0;
}

// ---------------------------------------------------------------------------------------------------------
// gc_mi_prepare: one thread per (gene, segment).  Reads Z[order[r], g] (coalesced across the genes of a warp).
// ---------------------------------------------------------------------------------------------------------
__global__ void mi_prepare_kernel(const float *__restrict__ Z, int64_t ldz, const int32_t *__restrict__ order,
                                  const int32_t *__restrict__ seg_off, int n_seg, int64_t G, int64_t n_tot,
                                  const double *__restrict__ xi1, const double *__restrict__ xi2,
                                  double *__restrict__ XR, double *__restrict__ YR) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (g >= G || c >= n_seg) return;
    const int lo = seg_off[c], n = seg_off[c + 1] - lo;
    if (n <= 0) return;
    const int32_t *ord = order + lo;
    auto zf = [&](int i) -> float { return Z[(int64_t)ord[i] * ldz + g]; };

    // ---- feature role: float64 (X.astype(float64); scale(with_mean=False); += 1e-10*max(1,mean|x|)*xi1) ----
    {
        auto x64 = [&](int i) -> double { return (double)zf(i); };
        const double mean = np_pairwise_sum<double>(x64, 0, n) / (double)n;
        auto dev2 = [&](int i) -> double { double d = x64(i) - mean; return d * d; };
        const double var = np_pairwise_sum<double>(dev2, 0, n) / (double)n;
        double sd = sqrt(var);
        if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;  // _handle_zeros_in_scale (array branch)
        auto xabs = [&](int i) -> double { return fabs(x64(i) / sd); };
        const double mabs = np_pairwise_sum<double>(xabs, 0, n) / (double)n;
        const double coef = 1e-10 * fmax(1.0, mabs);
        double *out = XR + g * n_tot + lo;
        const double *xi = xi1 + lo;
        for (int i = 0; i < n; ++i) out[i] = x64(i) / sd + coef * xi[i];
    }
    // ---- target role: float32 scale, float32 noise coefficient, float64 add rounded back to float32 ----------
    if (YR != nullptr) {
        auto y32 = [&](int i) -> float { return zf(i); };
        const float mean = np_pairwise_sum<float>(y32, 0, n) / (float)n;
        auto dev2 = [&](int i) -> float { float d = y32(i) - mean; return d * d; };
        const float var = np_pairwise_sum<float>(dev2, 0, n) / (float)n;
        float sd = sqrtf(var);
        if (sd == 0.0f) sd = 1.0f;  // _handle_zeros_in_scale (scalar branch: 1-D target)
        auto yabs = [&](int i) -> float { return fabsf(y32(i) / sd); };
        const float mabs = np_pairwise_sum<float>(yabs, 0, n) / (float)n;
        const float coef32 = 1e-10f * fmaxf(1.0f, mabs);
        const double coef = (double)coef32;
        double *out = YR + g * n_tot + lo;
        const double *xi = xi2 + lo;
        for (int i = 0; i < n; ++i) {
            const float ys = y32(i) / sd;
            out[i] = (double)(float)((double)ys + coef * xi[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gc_mi_cc_pairs: one warp per gene pair, n <= 64 samples.  Lane L owns samples L and L+32.
// Pass 1: k-th (k=3) smallest Chebyshev distance to the other points.  Pass 2: marginal counts.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPairWarps = 8;

__device__ __forceinline__ void pair_from_index(long long p, long long N, long long &i, long long &j) {
    // condensed index of itertools.combinations(range(N), 2): offset(i) = i*N - i*(i+1)/2
    const double b = 2.0 * (double)N - 1.0;
    long long ii = (long long)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > N - 2) ii = N - 2;
    while (ii > 0 && ii * N - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * N - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = ii;
    j = p - (ii * N - ii * (ii + 1) / 2) + ii + 1;
}

__global__ void __launch_bounds__(kPairWarps * 32)
mi_cc_pairs_kernel(const double *__restrict__ XR, const double *__restrict__ YR, long long N, int n, int64_t ldr,
                   const double *__restrict__ psi, long long p_begin, long long p_end,
                   double *__restrict__ mi_condensed, double *__restrict__ dense, uint8_t *__restrict__ nx_out,
                   uint8_t *__restrict__ ny_out) {
    __shared__ double sx[kPairWarps][64];
    __shared__ double sy[kPairWarps][64];
    __shared__ double spsi[65];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int m = threadIdx.x; m <= n; m += blockDim.x) spsi[m] = psi[m];
    __syncthreads();
    const double psi_n = spsi[n], psi_k = spsi[3];
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const bool v0 = lane < n, v1 = lane + 32 < n;

    const long long n_pairs = p_end - p_begin;
    const long long warp_global = (long long)blockIdx.x * kPairWarps + w;
    const long long n_warps = (long long)gridDim.x * kPairWarps;
    for (long long q = warp_global; q < n_pairs; q += n_warps) {
        long long i, j;
        pair_from_index(p_begin + q, N, i, j);
        const double *xg = XR + i * ldr, *yg = YR + j * ldr;
        const double xs0 = v0 ? xg[lane] : 0.0, ys0 = v0 ? yg[lane] : 0.0;
        const double xs1 = v1 ? xg[lane + 32] : 0.0, ys1 = v1 ? yg[lane + 32] : 0.0;
        __syncwarp();
        sx[w][lane] = xs0; sy[w][lane] = ys0;
        sx[w][lane + 32] = xs1; sy[w][lane + 32] = ys1;
        __syncwarp();

        double a0 = INF, a1 = INF, a2 = INF, b0 = INF, b1 = INF, b2 = INF;
#pragma unroll 5
        for (int t = 0; t < n; ++t) {
            const double xt = sx[w][t], yt = sy[w][t];
            double da = fmax(fabs(xs0 - xt), fabs(ys0 - yt));
            double db = fmax(fabs(xs1 - xt), fabs(ys1 - yt));
            if (t == lane) da = INF;        // exclude the point itself
            if (t == lane + 32) db = INF;
            insert3(da, a0, a1, a2);
            insert3(db, b0, b1, b2);
        }
        const double ra = next_toward_zero(a2), rb = next_toward_zero(b2);
        int nxa = 0, nya = 0, nxb = 0, nyb = 0;
#pragma unroll 5
        for (int t = 0; t < n; ++t) {
            const double xt = sx[w][t], yt = sy[w][t];
            nxa += fabs(xs0 - xt) <= ra; nya += fabs(ys0 - yt) <= ra;
            nxb += fabs(xs1 - xt) <= rb; nyb += fabs(ys1 - yt) <= rb;
        }
        // counts include the point itself (distance 0 <= r): nx+1 == the count, so psi(nx+1) = psi[count]
        if (nx_out != nullptr) {
            if (v0) { nx_out[q * n + lane] = (uint8_t)(nxa - 1); ny_out[q * n + lane] = (uint8_t)(nya - 1); }
            if (v1) { nx_out[q * n + lane + 32] = (uint8_t)(nxb - 1); ny_out[q * n + lane + 32] = (uint8_t)(nyb - 1); }
        }
        __syncwarp();
        // reuse the smem rows for psi(nx+1), psi(ny+1); then numpy-order means on lane 0
        sx[w][lane] = v0 ? spsi[nxa] : 0.0; sy[w][lane] = v0 ? spsi[nya] : 0.0;
        sx[w][lane + 32] = v1 ? spsi[nxb] : 0.0; sy[w][lane + 32] = v1 ? spsi[nyb] : 0.0;
        __syncwarp();
        if (lane < 2) {
            const double *src = lane == 0 ? sx[w] : sy[w];
            auto get = [&](int64_t s) -> double { return src[s]; };
            const double mean = np_pairwise_sum<double>(get, 0, n) / (double)n;
            const double other = __shfl_sync(0x3u, mean, 1);
            if (lane == 0) {
                double mi = ((psi_n + psi_k) - mean) - other;
                mi = mi > 0.0 ? mi : 0.0;
                mi_condensed[q] = mi;
                if (dense != nullptr) { dense[i * N + j] = mi; dense[j * N + i] = mi; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gc_mi_cc_pairs_sorted: the same estimator with 1-D sorted orders prepared once per gene.  The k-th neighbour
// search of sample s walks outward from s in x-sorted order and stops as soon as the next |dx| cannot beat the
// current third-smallest Chebyshev distance (every unvisited point has d >= |dx|), ~13 candidates instead of 49 for
// n = 50; the marginal counts are run lengths in the x- resp. y-sorted arrays.  Distances, radii and counts are the
// brute-force kernel's bit for bit (same subtractions, same order statistic); tests compare the two kernels.
// Measured SLOWER than brute force on GeneClust's 50-d PCA embeddings (50.9 vs 37.7 ms per 7 998 000 pairs: isolated
// large-coordinate samples walk the whole array and every warp waits for its slowest lane), so it serves as the
// independent cross-check, not as the production path.
// ---------------------------------------------------------------------------------------------------------
// one thread per (row, role): stable insertion sort of the row; sorted values, order (sample id at position) and
// rank (position of sample) as uint8, rows padded to 64
__global__ void mi_sort_rows_kernel(const double *__restrict__ XR, const double *__restrict__ YR, int64_t N, int n,
                                    int64_t ldr, double *__restrict__ xs, uint8_t *__restrict__ xo,
                                    double *__restrict__ ys, uint8_t *__restrict__ yrank) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= N) return;
    const bool is_y = blockIdx.y == 1;
    const double *src = (is_y ? YR : XR) + g * ldr;
    double v[64];
    uint8_t id[64];
    for (int s = 0; s < n; ++s) {
        const double x = src[s];
        int p = s;
        while (p > 0 && v[p - 1] > x) { v[p] = v[p - 1]; id[p] = id[p - 1]; --p; }    // strict >: stable
        v[p] = x; id[p] = (uint8_t)s;
    }
    double *dst = (is_y ? ys : xs) + g * 64;
    for (int p = 0; p < n; ++p) dst[p] = v[p];
    if (is_y) { for (int p = 0; p < n; ++p) yrank[g * 64 + id[p]] = (uint8_t)p; }
    else { for (int p = 0; p < n; ++p) xo[g * 64 + p] = id[p]; }
}

// k = 3 neighbour radius and marginal counts of the sample at x-sorted position p (y-sorted position yr)
__device__ __forceinline__ void ksg_sample_sorted(const double *__restrict__ Xs, const double *__restrict__ Yx,
                                                  const double *__restrict__ Ys, int n, int p, int yr, int &nx, int &ny) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const double xp = Xs[p], yp = Yx[p];
    int l = p - 1, r = p + 1;
    double m0 = INF, m1 = INF, m2 = INF;
    double dl = l >= 0 ? xp - Xs[l] : INF, dr = r < n ? Xs[r] - xp : INF;
    while (true) {
        const bool take_l = dl <= dr;
        const double dx = take_l ? dl : dr;
        if (!(dx < m2)) break;                     // every unvisited point has d >= |dx| >= the current third smallest
        const int q = take_l ? l : r;
        const double d = fmax(dx, fabs(yp - Yx[q]));
        insert3(d, m0, m1, m2);
        if (take_l) { --l; dl = l >= 0 ? xp - Xs[l] : INF; }
        else { ++r; dr = r < n ? Xs[r] - xp : INF; }
    }
    const double rad = next_toward_zero(m2);
    int c = 1;                                     // counts include the point itself, as in the brute-force kernel
    for (int q = p - 1; q >= 0 && xp - Xs[q] <= rad; --q) ++c;
    for (int q = p + 1; q < n && Xs[q] - xp <= rad; ++q) ++c;
    nx = c;
    c = 1;
    for (int q = yr - 1; q >= 0 && yp - Ys[q] <= rad; --q) ++c;
    for (int q = yr + 1; q < n && Ys[q] - yp <= rad; ++q) ++c;
    ny = c;
}

__global__ void __launch_bounds__(kPairWarps * 32)
mi_cc_pairs_sorted_kernel(const double *__restrict__ YR, long long N, int n, int64_t ldr,
                          const double *__restrict__ XS, const uint8_t *__restrict__ XO,
                          const double *__restrict__ YS, const uint8_t *__restrict__ YRK,
                          const double *__restrict__ psi, long long p_begin, long long p_end,
                          double *__restrict__ mi_condensed, double *__restrict__ dense, uint8_t *__restrict__ nx_out,
                          uint8_t *__restrict__ ny_out) {
    __shared__ double sXs[kPairWarps][64];     // x, x-sorted
    __shared__ double sYx[kPairWarps][64];     // y of the sample at each x-sorted position
    __shared__ double sYs[kPairWarps][64];     // y, y-sorted
    __shared__ double sP[kPairWarps][64];      // psi(nx + 1) by sample id
    __shared__ double sQ[kPairWarps][64];      // psi(ny + 1) by sample id
    __shared__ double spsi[65];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int m = threadIdx.x; m <= n; m += blockDim.x) spsi[m] = psi[m];
    __syncthreads();
    const double psi_n = spsi[n], psi_k = spsi[3];
    const bool v0 = lane < n, v1 = lane + 32 < n;

    const long long n_pairs = p_end - p_begin;
    const long long warp_global = (long long)blockIdx.x * kPairWarps + w;
    const long long n_warps = (long long)gridDim.x * kPairWarps;
    for (long long q = warp_global; q < n_pairs; q += n_warps) {
        long long i, j;
        pair_from_index(p_begin + q, N, i, j);
        const double *yg = YR + j * ldr;
        int sid0 = 0, sid1 = 0, yr0 = 0, yr1 = 0;
        __syncwarp();
        if (v0) {
            sid0 = XO[i * 64 + lane];
            sXs[w][lane] = XS[i * 64 + lane]; sYx[w][lane] = yg[sid0]; sYs[w][lane] = YS[j * 64 + lane];
            yr0 = YRK[j * 64 + sid0];
        }
        if (v1) {
            sid1 = XO[i * 64 + lane + 32];
            sXs[w][lane + 32] = XS[i * 64 + lane + 32]; sYx[w][lane + 32] = yg[sid1]; sYs[w][lane + 32] = YS[j * 64 + lane + 32];
            yr1 = YRK[j * 64 + sid1];
        }
        __syncwarp();
        int nxa = 1, nya = 1, nxb = 1, nyb = 1;
        if (v0) ksg_sample_sorted(sXs[w], sYx[w], sYs[w], n, lane, yr0, nxa, nya);
        if (v1) ksg_sample_sorted(sXs[w], sYx[w], sYs[w], n, lane + 32, yr1, nxb, nyb);
        if (nx_out != nullptr) {
            if (v0) { nx_out[q * n + sid0] = (uint8_t)(nxa - 1); ny_out[q * n + sid0] = (uint8_t)(nya - 1); }
            if (v1) { nx_out[q * n + sid1] = (uint8_t)(nxb - 1); ny_out[q * n + sid1] = (uint8_t)(nyb - 1); }
        }
        if (v0) { sP[w][sid0] = spsi[nxa]; sQ[w][sid0] = spsi[nya]; }
        if (v1) { sP[w][sid1] = spsi[nxb]; sQ[w][sid1] = spsi[nyb]; }
        __syncwarp();
        if (lane < 2) {
            const double *src = lane == 0 ? sP[w] : sQ[w];
            auto get = [&](int64_t s) -> double { return src[s]; };
            const double mean = np_pairwise_sum<double>(get, 0, n) / (double)n;
            const double other = __shfl_sync(0x3u, mean, 1);
            if (lane == 0) {
                double mi = ((psi_n + psi_k) - mean) - other;
                mi = mi > 0.0 ? mi : 0.0;
                mi_condensed[q] = mi;
                if (dense != nullptr) { dense[i * N + j] = mi; dense[j * N + i] = mi; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gc_mi_cc_problems: one CTA per variable-length KSG problem (complementarity).  Threads own samples s,
// the partner points stream through shared memory in tiles.  nx/ny are written to global memory; the
// numpy-order digamma means are taken by a second kernel (one thread per problem).
// ---------------------------------------------------------------------------------------------------------
constexpr int kProbThreads = 256;
constexpr int kProbTile = 1024;

__global__ void __launch_bounds__(kProbThreads)
mi_cc_problems_kernel(const double *__restrict__ XR, const double *__restrict__ YR,
                      const int64_t *__restrict__ x_off, const int64_t *__restrict__ y_off,
                      const int32_t *__restrict__ len, const int64_t *__restrict__ out_off,
                      int32_t *__restrict__ nx, int32_t *__restrict__ ny) {
    __shared__ double tx[kProbTile];
    __shared__ double ty[kProbTile];
    const int q = blockIdx.x;
    const int n = len[q];
    const double *x = XR + x_off[q], *y = YR + y_off[q];
    int32_t *onx = nx + out_off[q], *ony = ny + out_off[q];
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int s_base = 0; s_base < n; s_base += kProbThreads) {
        const int s = s_base + threadIdx.x;
        const bool valid = s < n;
        const double xs = valid ? x[s] : 0.0, ys = valid ? y[s] : 0.0;
        double m0 = INF, m1 = INF, m2 = INF;
        for (int t0 = 0; t0 < n; t0 += kProbTile) {
            const int tn = min(kProbTile, n - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < tn; t += kProbThreads) { tx[t] = x[t0 + t]; ty[t] = y[t0 + t]; }
            __syncthreads();
            for (int t = 0; t < tn; ++t) {
                double d = fmax(fabs(xs - tx[t]), fabs(ys - ty[t]));
                if (t0 + t == s) d = INF;
                insert3(d, m0, m1, m2);
            }
        }
        const double r = next_toward_zero(m2);
        int cx = 0, cy = 0;
        for (int t0 = 0; t0 < n; t0 += kProbTile) {
            const int tn = min(kProbTile, n - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < tn; t += kProbThreads) { tx[t] = x[t0 + t]; ty[t] = y[t0 + t]; }
            __syncthreads();
            for (int t = 0; t < tn; ++t) {
                cx += fabs(xs - tx[t]) <= r;
                cy += fabs(ys - ty[t]) <= r;
            }
        }
        if (valid) { onx[s] = cx - 1; ony[s] = cy - 1; }
    }
}

__global__ void mi_cc_finalize_kernel(const int32_t *__restrict__ nx, const int32_t *__restrict__ ny,
                                      const int32_t *__restrict__ len, const int64_t *__restrict__ out_off,
                                      int64_t n_prob, const double *__restrict__ psi, double *__restrict__ mi) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_prob) return;
    const int n = len[q];
    const int32_t *a = nx + out_off[q], *b = ny + out_off[q];
    auto ga = [&](int s) -> double { return psi[a[s] + 1]; };
    auto gb = [&](int s) -> double { return psi[b[s] + 1]; };
    const double ma = np_pairwise_sum<double>(ga, 0, n) / (double)n;
    const double mb = np_pairwise_sum<double>(gb, 0, n) / (double)n;
    const double v = ((psi[n] + psi[3]) - ma) - mb;
    mi[q] = v > 0.0 ? v : 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// gc_mi_cd: Ross estimator, one CTA per gene (brute force over the n samples, tiled through shared memory).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kProbThreads)
mi_cd_counts_kernel(const double *__restrict__ XR, int64_t ldr, int64_t G, int n,
                    const int32_t *__restrict__ label, const int32_t *__restrict__ label_k,
                    int32_t *__restrict__ m_all) {
    __shared__ double tc[kProbTile];
    __shared__ int32_t tl[kProbTile];
    const int64_t g = blockIdx.x;
    const double *c = XR + g * ldr;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int s_base = 0; s_base < n; s_base += kProbThreads) {
        const int s = s_base + threadIdx.x;
        const bool valid = s < n;
        const double cs = valid ? c[s] : 0.0;
        const int ls = valid ? label[s] : -1;
        const int ks = valid ? label_k[ls] : 0;   // 0 => unique label, sample ignored
        double m0 = INF, m1 = INF, m2 = INF;
        for (int t0 = 0; t0 < n; t0 += kProbTile) {
            const int tn = min(kProbTile, n - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < tn; t += kProbThreads) { tc[t] = c[t0 + t]; tl[t] = label[t0 + t]; }
            __syncthreads();
            for (int t = 0; t < tn; ++t) {
                double d = fabs(cs - tc[t]);
                if (tl[t] != ls || t0 + t == s) d = INF;
                insert3(d, m0, m1, m2);
            }
        }
        const double eps = ks >= 3 ? m2 : (ks == 2 ? m1 : m0);
        const double r = next_toward_zero(eps);
        int cnt = 0;
        for (int t0 = 0; t0 < n; t0 += kProbTile) {
            const int tn = min(kProbTile, n - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < tn; t += kProbThreads) {
                tc[t] = c[t0 + t];
                tl[t] = label_k[label[t0 + t]];  // > 0: point takes part in the global count
            }
            __syncthreads();
            for (int t = 0; t < tn; ++t) cnt += (tl[t] > 0) & (fabs(cs - tc[t]) <= r);
        }
        if (valid) m_all[(int64_t)s * G + g] = ks > 0 ? cnt : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same Ross counts in O(n log n) per gene: the 1-D estimator only needs, per sample, the k-th nearest value of
// its own class and the number of values of any class within that radius.  One CTA per gene: bitonic sort of the
// kept samples by value in shared memory; a stable partition by label gives every class its sorted sub-sequence, so
// the k-th same-class neighbour is among the k predecessors / successors there; the global count is two binary
// searches in the fully sorted array with the exact float64 predicate  v_s - v_t <= r  (monotone in v_t).
// Distances, radii (nextafter) and counts are those of the brute-force kernel bit for bit; tests compare the two.
// n <= 16384 (shared memory: 14 bytes per padded sample).
// ---------------------------------------------------------------------------------------------------------
constexpr int kCdThreads = 1024;
__global__ void __launch_bounds__(kCdThreads)
mi_cd_counts_sorted_kernel(const double *__restrict__ XR, int64_t ldr, int64_t G, int n, int np2,
                           const int32_t *__restrict__ label, const int32_t *__restrict__ label_k, int n_label,
                           int32_t *__restrict__ m_all) {
    extern __shared__ __align__(16) unsigned char cd_smem[];
    double *V = reinterpret_cast<double *>(cd_smem);                          // [np2] values, sorted ascending
    uint16_t *ID = reinterpret_cast<uint16_t *>(V + np2);                     // [np2] sample id at sorted position
    uint16_t *CQ = ID + np2;                                                  // [np2] class-order position of sorted position
    uint16_t *CPOS = CQ + np2;                                                // [np2] sorted position of class-order position
    __shared__ int s_warp[32];
    __shared__ int s_total, s_kept;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t g = blockIdx.x;
    const double *c = XR + g * ldr;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (tid == 0) s_kept = 0;
    __syncthreads();
    int kept_local = 0;
    for (int s = tid; s < np2; s += kCdThreads) {
        bool kept = false;
        if (s < n) {
            kept = label_k[label[s]] > 0;
            if (!kept) m_all[(int64_t)s * G + g] = 0;
        }
        V[s] = kept ? c[s] : INF;
        ID[s] = (uint16_t)(s < n ? s : 0xFFFF);
        kept_local += kept;
    }
    for (int o = 16; o > 0; o >>= 1) kept_local += __shfl_xor_sync(0xffffffffu, kept_local, o);
    if (lane == 0 && kept_local) atomicAdd(&s_kept, kept_local);
    __syncthreads();
    const int nk = s_kept;                                                    // kept samples sort to [0, nk)
    // ---- bitonic sort by value (ties in any order: equal values give equal distances and counts)
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += kCdThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = V[i], b = V[ixj];
                    const bool up = (i & k) == 0;
                    if (up ? (a > b) : (a < b)) {
                        V[i] = b; V[ixj] = a;
                        const uint16_t t = ID[i]; ID[i] = ID[ixj]; ID[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- stable partition by label: class l occupies class-order positions [off_l, off_l + cnt_l)
    const int per = (nk + kCdThreads - 1) / kCdThreads;                       // contiguous slice per thread
    const int p_lo = min(nk, tid * per), p_hi = min(nk, p_lo + per);
    int off = 0;
    for (int l = 0; l < n_label; ++l) {
        if (label_k[l] <= 0) continue;
        int cnt = 0;
        for (int p = p_lo; p < p_hi; ++p) cnt += label[ID[p]] == l;
        int incl = cnt;                                                       // warp inclusive scan
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
            int wi = w;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            s_warp[lane] = wi - w;                                            // exclusive warp offsets
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        int pos = off + s_warp[wid] + incl - cnt;
        for (int p = p_lo; p < p_hi; ++p) {
            if (label[ID[p]] == l) { CQ[p] = (uint16_t)pos; CPOS[pos] = (uint16_t)p; ++pos; }
        }
        off += s_total;
        __syncthreads();
    }
    // ---- per sample: k-th same-class neighbour distance, then the global count within nextafter(radius)
    for (int p = tid; p < nk; p += kCdThreads) {
        const int sid = ID[p];
        const int l = label[sid];
        const int ks = label_k[l];
        const double v = V[p];
        const int q = CQ[p];
        // class segment bounds: walk is bounded by k <= 3 steps, membership checked through the label of the neighbour
        double dl[3] = {INF, INF, INF}, dr[3] = {INF, INF, INF};
#pragma unroll
        for (int j = 1; j <= 3; ++j) {
            if (q - j >= 0) { const int pp = CPOS[q - j]; if (label[ID[pp]] == l) dl[j - 1] = v - V[pp]; }
            if (q + j < nk) { const int pp = CPOS[q + j]; if (label[ID[pp]] == l) dr[j - 1] = V[pp] - v; }
        }
        // k-th smallest of the two ascending lists
        int il = 0, ir = 0;
        double eps = INF;
        for (int t = 0; t < ks; ++t) {
            const double a = il < 3 ? dl[il] : INF, b = ir < 3 ? dr[ir] : INF;
            if (a <= b) { eps = a; ++il; } else { eps = b; ++ir; }
        }
        const double r = next_toward_zero(eps);
        // smallest i in [0, p] with v - V[i] <= r ; largest j in [p, nk) with V[j] - v <= r
        int lo = 0, hi = p;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (v - V[mid] <= r) hi = mid; else lo = mid + 1; }
        const int first = lo;
        lo = p; hi = nk - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (V[mid] - v <= r) lo = mid; else hi = mid - 1; }
        m_all[(int64_t)sid * G + g] = lo - first + 1;
    }
}

__global__ void mi_cd_finalize_kernel(const int32_t *__restrict__ m_all, int64_t G, int n,
                                      const int32_t *__restrict__ label, const int32_t *__restrict__ label_k,
                                      double const_term, const double *__restrict__ psi,
                                      double *__restrict__ mi) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    // mean over the kept samples in their original order; kept samples are compacted implicitly by walking
    // a running index, so the pairwise tree is built over the compacted sequence like numpy does.
    // The compaction map is identical for all genes, so it is recomputed from the labels.
    int n_kept = 0;
    for (int s = 0; s < n; ++s) n_kept += label_k[label[s]] > 0;
    // kept index -> sample index is monotone; emulate with a forward-only cursor (accesses are monotone
    // within every leaf and across the recursion order of np_pairwise_sum).
    int cur_kept = -1, cur_s = -1;
    auto get = [&](int kidx) -> double {
        if (kidx < cur_kept) { cur_kept = -1; cur_s = -1; }
        while (cur_kept < kidx) { ++cur_s; cur_kept += label_k[label[cur_s]] > 0; }
        return psi[m_all[(int64_t)cur_s * G + g]];
    };
    const double mm = np_pairwise_sum<double>(get, 0, n_kept) / (double)n_kept;
    const double v = const_term - mm;
    mi[g] = v > 0.0 ? v : 0.0;
}

}  // namespace gc

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
using namespace gc;

extern "C" int gc_mi_prepare(const float *Z, int64_t ldz, const int32_t *order, const int32_t *seg_off, int n_seg,
                             int64_t G, int64_t n_tot, const double *xi1, const double *xi2, double *XR, double *YR,
                             void *stream) {
    GC_REQUIRE(Z && order && seg_off && xi1 && XR, "null pointer");
    GC_REQUIRE((YR == nullptr) == (xi2 == nullptr), "YR and xi2 must be given together");
    GC_REQUIRE(n_seg >= 1 && G >= 1 && n_tot >= 1, "empty problem");
    dim3 grid((unsigned)ceil_div<int64_t>(G, 128), (unsigned)n_seg);
    mi_prepare_kernel<<<grid, 128, 0, as_stream(stream)>>>(Z, ldz, order, seg_off, n_seg, G, n_tot, xi1, xi2, XR, YR);
    return check_launch("mi_prepare_kernel");
}

extern "C" int gc_mi_cc_pairs(const double *XR, const double *YR, int64_t N, int n, int64_t ldr, const double *psi,
                              int64_t p_begin, int64_t p_end, double *mi_condensed, double *dense,
                              uint8_t *nx_out, uint8_t *ny_out, void *stream) {
    GC_REQUIRE(XR && YR && psi && mi_condensed, "null pointer");
    GC_REQUIRE(n >= 4 && n <= 64, "n must be in [4, 64] (k = 3 neighbours)");
    GC_REQUIRE(N >= 2 && ldr >= n, "bad shape");
    GC_REQUIRE(p_begin >= 0 && p_end >= p_begin && p_end <= N * (N - 1) / 2, "pair range out of bounds");
    GC_REQUIRE((nx_out == nullptr) == (ny_out == nullptr), "nx_out and ny_out must be given together");
    if (p_end == p_begin) return 0;
    const long long n_pairs = p_end - p_begin;
    long long blocks = ceil_div<long long>(n_pairs, kPairWarps);
    const long long cap = (long long)kNumSMs * 8 * 4;  // 8 CTAs of 8 warps per SM, 4 waves; grid-stride beyond
    if (blocks > cap) blocks = cap;
    prof_begin(kProfMiPairs, as_stream(stream));
    mi_cc_pairs_kernel<<<(unsigned)blocks, kPairWarps * 32, 0, as_stream(stream)>>>(
        XR, YR, N, n, ldr, psi, p_begin, p_end, mi_condensed, dense, nx_out, ny_out);
    prof_end(kProfMiPairs, as_stream(stream));
    return check_launch("mi_cc_pairs_kernel");
}

extern "C" int64_t gc_mi_cc_pairs_scratch_bytes(int64_t N, int n) {
    (void)n;
    return 2 * N * 64 * 8 + 2 * N * 64 + 512;
}

extern "C" int gc_mi_cc_pairs_sorted(const double *XR, const double *YR, int64_t N, int n, int64_t ldr, const double *psi,
                                     int64_t p_begin, int64_t p_end, double *mi_condensed, double *dense,
                                     uint8_t *nx_out, uint8_t *ny_out, void *scratch, void *stream) {
    GC_REQUIRE(XR && YR && psi && mi_condensed && scratch, "null pointer");
    GC_REQUIRE(n >= 4 && n <= 64, "n must be in [4, 64] (k = 3 neighbours)");
    GC_REQUIRE(N >= 2 && ldr >= n, "bad shape");
    GC_REQUIRE(p_begin >= 0 && p_end >= p_begin && p_end <= N * (N - 1) / 2, "pair range out of bounds");
    GC_REQUIRE((nx_out == nullptr) == (ny_out == nullptr), "nx_out and ny_out must be given together");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    if (p_end == p_begin) return 0;
    cudaStream_t st = as_stream(stream);
    char *p = (char *)scratch;
    double *xs = (double *)p;     p += N * 64 * 8;
    double *ys = (double *)p;     p += N * 64 * 8;
    uint8_t *xo = (uint8_t *)p;   p += N * 64;
    uint8_t *yrank = (uint8_t *)p;
    mi_sort_rows_kernel<<<dim3((unsigned)ceil_div<int64_t>(N, 64), 2), 64, 0, st>>>(XR, YR, N, n, ldr, xs, xo, ys, yrank);
    if (int rc = check_launch("mi_sort_rows_kernel")) return rc;
    const long long n_pairs = p_end - p_begin;
    long long blocks = ceil_div<long long>(n_pairs, kPairWarps);
    const long long cap = (long long)kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    prof_begin(kProfMiPairs, st);
    mi_cc_pairs_sorted_kernel<<<(unsigned)blocks, kPairWarps * 32, 0, st>>>(YR, N, n, ldr, xs, xo, ys, yrank, psi, p_begin, p_end,
                                                                           mi_condensed, dense, nx_out, ny_out);
    prof_end(kProfMiPairs, st);
    return check_launch("mi_cc_pairs_sorted_kernel");
}

extern "C" int gc_mi_cc_problems(const double *XR, const double *YR, const int64_t *x_off, const int64_t *y_off,
                                 const int32_t *len, const int64_t *out_off, int64_t n_prob, const double *psi,
                                 int32_t *nx, int32_t *ny, double *mi, void *stream) {
    GC_REQUIRE(XR && YR && x_off && y_off && len && out_off && psi && nx && ny && mi, "null pointer");
    if (n_prob == 0) return 0;
    GC_REQUIRE(n_prob > 0 && n_prob < (1LL << 31), "bad problem count");
    mi_cc_problems_kernel<<<(unsigned)n_prob, kProbThreads, 0, as_stream(stream)>>>(XR, YR, x_off, y_off, len,
                                                                                   out_off, nx, ny);
    if (int rc = check_launch("mi_cc_problems_kernel")) return rc;
    mi_cc_finalize_kernel<<<(unsigned)ceil_div<int64_t>(n_prob, 64), 64, 0, as_stream(stream)>>>(
        nx, ny, len, out_off, n_prob, psi, mi);
    return check_launch("mi_cc_finalize_kernel");
}

extern "C" int gc_mi_cd(const double *XR, int64_t ldr, int64_t G, int n, const int32_t *label,
                        const int32_t *label_k, int n_label, double const_term, const double *psi, int32_t *m_all,
                        double *mi, void *stream) {
    GC_REQUIRE(XR && label && label_k && psi && m_all && mi, "null pointer");
    GC_REQUIRE(G >= 1 && n >= 2 && ldr >= n && n_label >= 1, "bad shape");
    int np2 = 64;
    while (np2 < n) np2 <<= 1;
    static const bool force_brute = getenv("GC_MI_CD_BRUTE_FORCE") != nullptr;      // cross-check switch for the tests
    if (n <= 16384 && !force_brute) {
        const size_t smem = (size_t)np2 * 14;
        static thread_local size_t configured[kMaxDevices] = {};
        const int dev_id = current_device();
        if (smem > 48 * 1024 && smem > configured[dev_id]) {
            GC_CUDA(cudaFuncSetAttribute(mi_cd_counts_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev_id] = smem;
        }
        mi_cd_counts_sorted_kernel<<<(unsigned)G, kCdThreads, smem, as_stream(stream)>>>(XR, ldr, G, n, np2, label, label_k,
                                                                                      n_label, m_all);
        if (int rc = check_launch("mi_cd_counts_sorted_kernel")) return rc;
    } else {
        mi_cd_counts_kernel<<<(unsigned)G, kProbThreads, 0, as_stream(stream)>>>(XR, ldr, G, n, label, label_k, m_all);
        if (int rc = check_launch("mi_cd_counts_kernel")) return rc;
    }
    mi_cd_finalize_kernel<<<(unsigned)ceil_div<int64_t>(G, 64), 64, 0, as_stream(stream)>>>(
        m_all, G, n, label, label_k, const_term, psi, mi);
    return check_launch("mi_cd_finalize_kernel");
}
