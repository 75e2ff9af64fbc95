// q x q (q <= 64) float64 dense algebra on one CTA: the panel normaliser and the thin SVD of randomized SVD
// (replaces LAPACK getrf/geqrf/gesdd reached from sklearn utils/extmath.py:377-386, :606-619 via
// scGeneClust/pp/_preprocessing.py:52).  Kept on the device so the 16 passes run without host round trips.
#include "common.cuh"

#include <cstdlib>

namespace gc {

constexpr int kN = 64;

// Cholesky Gm = L L^T of the leading q x q block, then Rinv = (L^-1)^T as float32 (64 x 64, zero padded), so that
// P @ Rinv has orthonormal columns when Gm = P^T P.  status[0] |= 1 on a non-positive pivot.
// 1024 threads = 64 rows x 16 lanes: the length-j dot product of every row is split over 16 lanes (k = lane, lane+16,
// ...) and folded with shuffles in a fixed order, so a column costs ~4 dependent FMAs instead of up to 60.  Same for
// the forward substitutions of the inverse (one column of L^-1 per 16-lane group).
// LANES lanes per row (16: 1024 threads, 4: 256 threads - fewer warps make the three barriers per column cheaper).
template <int LANES>
__global__ void __launch_bounds__(kN * LANES)
chol_inv_kernel(const double *__restrict__ Gm, int q, float *__restrict__ Rinv, int32_t *status) {
    extern __shared__ double dyn_smem[];
    double (*L)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem);
    double (*Li)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem + kN * (kN + 1));
    __shared__ double rdiag[kN];                                   // 1 / L[j][j]
    const int i = threadIdx.x / LANES, t = threadIdx.x % LANES;    // row / column owner, lane inside its group
    for (int j = t; j < kN; j += LANES) { L[i][j] = (i < q && j < q) ? Gm[i * kN + j] : 0.0; Li[i][j] = 0.0; }
    __syncthreads();
    for (int j = 0; j < q; ++j) {
        double s = 0.0;
        if (i >= j && i < q) {
            for (int k = t; k < j; k += LANES) s = fma(L[i][k], L[j][k], s);
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LANES);
        __syncthreads();                                            // all reads of column j's inputs are done
        if (t == 0 && i >= j && i < q) {
            double v = L[i][j] - s;
            if (i == j) {
                if (!(v > 0.0)) { atomicOr(status, 1); v = 1.0; }
                const double rs = rsqrt(v);                         // one reciprocal square root per column: the rows
                rdiag[j] = rs;                                      // multiply by it instead of dividing by the pivot
                L[j][j] = v * rs;
            } else {
                L[i][j] = v;
            }
        }
        __syncthreads();
        if (t == 0 && i > j && i < q) L[i][j] = L[i][j] * rdiag[j];
        __syncthreads();
    }
    // inverse of the lower-triangular L: the 16-lane group `i` solves column i of L X = I by forward substitution
    // (two groups share a warp and run different trip counts: group-wide masks)
    const unsigned gmask = (LANES == 32 ? 0xFFFFFFFFu : ((1u << LANES) - 1u)) << ((threadIdx.x & 31) / LANES * LANES);
    if (i < q) {
        for (int r = i; r < q; ++r) {
            double s = 0.0;
            for (int k = i + t; k < r; k += LANES) s = fma(L[r][k], Li[k][i], s);
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o, LANES);
            if (t == 0) Li[r][i] = ((r == i ? 1.0 : 0.0) - s) * rdiag[r];
            __syncwarp(gmask);
        }
    }
    __syncthreads();
    // Rinv[a][b] = (L^-1)^T[a][b] = Li[b][a]
    for (int j = t; j < kN; j += LANES) Rinv[i * kN + j] = (i < q && j < q) ? (float)Li[j][i] : 0.f;
}

// Symmetric eigendecomposition of the leading q x q block of Gm by cyclic two-sided Jacobi with the round-robin
// parallel ordering.  Outputs eigenvalues in descending order (evals[64], zero padded) and the matching
// eigenvectors as columns of U (float32 and float64, 64 x 64 row-major, zero padded).
constexpr int kEigThreads = 1024;   // ~1800 row/column updates per Jacobi round: two per thread
__global__ void __launch_bounds__(kEigThreads)
sym_eig_kernel(const double *__restrict__ Gm, int q, double *__restrict__ evals, float *__restrict__ U32,
               double *__restrict__ U64) {
    extern __shared__ double dyn_smem[];
    double (*A)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem);
    double (*V)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem + kN * (kN + 1));
    __shared__ double cs[kN / 2][2];
    __shared__ int pp[kN / 2], qq[kN / 2];
    __shared__ double off2, diag2;
    __shared__ double rowsum[kN];
    __shared__ int order[kN];
    const int tid = threadIdx.x;
    const int m = (q + 1) & ~1;          // even number of players (one dummy when q is odd)
    const int n_pairs = m / 2;
    for (int idx = tid; idx < kN * kN; idx += kEigThreads) {
        const int r = idx / kN, c = idx % kN;
        A[r][c] = (r < q && c < q) ? Gm[r * kN + c] : 0.0;
        V[r][c] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        // off-diagonal / diagonal squared norms: thread r < q sums its row, thread 0 adds the q row sums in order
        if (tid < q) {
            double o = 0.0;
            for (int c = tid + 1; c < q; ++c) o += A[tid][c] * A[tid][c];
            rowsum[tid] = o;
        }
        __syncthreads();
        if (tid == 0) {
            double o = 0.0, d = 0.0;
            for (int r = 0; r < q; ++r) { o += rowsum[r]; d += A[r][r] * A[r][r]; }
            off2 = o; diag2 = d;
        }
        __syncthreads();
        // 1e-27: the off-diagonal mass cannot fall below ~q^2 eps^2 = 2e-29 of the diagonal mass in float64, so a
        // stricter test would only burn the remaining sweeps; at 1e-27 the eigenvectors are converged to ~3e-14
        if (off2 <= 1e-27 * diag2) break;
        for (int round = 0; round < m - 1; ++round) {
            if (tid < n_pairs) {
                // round-robin tournament: player m-1 fixed, the others rotate
                int a = tid == 0 ? m - 1 : (round + tid) % (m - 1);
                int b = (round + m - 1 - tid) % (m - 1);
                if (a > b) { const int t = a; a = b; b = t; }
                double c = 1.0, s = 0.0;
                if (b < q && A[a][b] != 0.0) {
                    const double apq = A[a][b], tau = (A[b][b] - A[a][a]) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                pp[tid] = a; qq[tid] = b; cs[tid][0] = c; cs[tid][1] = s;
            }
            __syncthreads();
            // A <- A J  (columns), V <- V J
            for (int task = tid; task < n_pairs * q; task += kEigThreads) {
                const int pr = task / q, r = task % q;
                const int a = pp[pr], b = qq[pr];
                if (b >= q) continue;
                const double c = cs[pr][0], s = cs[pr][1];
                const double x = A[r][a], y = A[r][b];
                A[r][a] = c * x - s * y; A[r][b] = s * x + c * y;
                const double vx = V[r][a], vy = V[r][b];
                V[r][a] = c * vx - s * vy; V[r][b] = s * vx + c * vy;
            }
            __syncthreads();
            // A <- J^T A  (rows)
            for (int task = tid; task < n_pairs * q; task += kEigThreads) {
                const int pr = task / q, col = task % q;
                const int a = pp[pr], b = qq[pr];
                if (b >= q) continue;
                const double c = cs[pr][0], s = cs[pr][1];
                const double x = A[a][col], y = A[b][col];
                A[a][col] = c * x - s * y; A[b][col] = s * x + c * y;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // descending order by selection (rank counting), ties by index
    if (tid < kN) {
        if (tid < q) {
            int rank = 0;
            const double v = A[tid][tid];
            for (int j = 0; j < q; ++j) {
                const double w = A[j][j];
                rank += (w > v) || (w == v && j < tid);
            }
            order[rank] = tid;
        }
    }
    __syncthreads();
    if (tid < kN) evals[tid] = tid < q ? A[order[tid]][order[tid]] : 0.0;
    for (int idx = tid; idx < kN * kN; idx += kEigThreads) {
        const int r = idx / kN, c = idx % kN;
        const double v = (r < q && c < q) ? V[r][order[c]] : 0.0;
        U32[idx] = (float)v;
        if (U64) U64[idx] = v;
    }
}

// R[:, j] *= sign(val[j]) * scale[j] for the leading q columns (svd_flip + optional column scaling)
__global__ void scale_cols_kernel(float *__restrict__ R, const float *__restrict__ sign_src,
                                  const double *__restrict__ scale, int q) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kN * kN) return;
    const int c = idx % kN;
    if (c >= q) { R[idx] = 0.f; return; }
    double f = 1.0;
    if (sign_src) f = sign_src[c] < 0.f ? -1.0 : (sign_src[c] > 0.f ? 1.0 : 0.0);   // np.sign
    if (scale) f *= scale[c];
    R[idx] = (float)((double)R[idx] * f);
}

}  // namespace gc

using namespace gc;

extern "C" int gc_chol_inv(const double *Gm, int q, float *Rinv, int32_t *status, void *stream) {
    GC_REQUIRE(Gm && Rinv && status, "null pointer");
    GC_REQUIRE(q >= 1 && q <= kN, "q must be in [1, 64]");
    const int smem = 2 * kN * (kN + 1) * (int)sizeof(double);
    static const int lanes = []() { const char *e = getenv("GC_CHOL_LANES"); return e ? atoi(e) : 8; }();   // developer knob (4 / 8 / 16 measured alike)
    static thread_local bool configured[kMaxDevices] = {};
    const int dev_id = current_device();
    if (!configured[dev_id]) {
        GC_CUDA(cudaFuncSetAttribute(chol_inv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GC_CUDA(cudaFuncSetAttribute(chol_inv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GC_CUDA(cudaFuncSetAttribute(chol_inv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev_id] = true;
    }
    if (lanes == 16) chol_inv_kernel<16><<<1, kN * 16, smem, as_stream(stream)>>>(Gm, q, Rinv, status);
    else if (lanes == 8) chol_inv_kernel<8><<<1, kN * 8, smem, as_stream(stream)>>>(Gm, q, Rinv, status);
    else chol_inv_kernel<4><<<1, kN * 4, smem, as_stream(stream)>>>(Gm, q, Rinv, status);
    return check_launch("chol_inv_kernel");
}

extern "C" int gc_sym_eig(const double *Gm, int q, double *evals, float *U32, double *U64, void *stream) {
    GC_REQUIRE(Gm && evals && U32, "null pointer");
    GC_REQUIRE(q >= 1 && q <= kN, "q must be in [1, 64]");
    const int smem = 2 * kN * (kN + 1) * (int)sizeof(double);
    static thread_local bool configured[kMaxDevices] = {};
    const int dev_id = current_device();
    if (!configured[dev_id]) {
        GC_CUDA(cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev_id] = true;
    }
    sym_eig_kernel<<<1, kEigThreads, smem, as_stream(stream)>>>(Gm, q, evals, U32, U64);
    return check_launch("sym_eig_kernel");
}

extern "C" int gc_scale_cols(float *R, const float *sign_src, const double *scale, int q, void *stream) {
    GC_REQUIRE(R != nullptr, "null pointer");
    GC_REQUIRE(q >= 1 && q <= kN, "q must be in [1, 64]");
    scale_cols_kernel<<<kN * kN / 256, 256, 0, as_stream(stream)>>>(R, sign_src, scale, q);
    return check_launch("scale_cols_kernel");
}
