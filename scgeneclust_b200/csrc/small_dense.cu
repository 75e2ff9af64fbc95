// q x q (q <= 64) float64 dense algebra on one CTA: the panel normaliser and the thin SVD of randomized SVD
// (replaces LAPACK getrf/geqrf/gesdd reached from sklearn utils/extmath.py:377-386, :606-619 via
// scGeneClust/pp/_preprocessing.py:52).  Kept on the device so the 16 passes run without host round trips.
#include "common.cuh"

#include <cstdlib>

namespace gc {

constexpr int kN = 64;

// 1 / sqrt(x) to ~2e-16 relative for x > 0: float32 seed (2^-22) and two Newton steps in float64 - a third of the
// latency of rsqrt(double), which sits on the serial path of every Cholesky column and every Jacobi round
__device__ __forceinline__ double fast_rsqrt(double x) {
    const float xf = (float)x;
    if (!(xf > 1e-30f && xf < 1e30f)) return rsqrt(x);
    double y = (double)rsqrtf(xf);
    const double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

// Cholesky Gm = L L^T of the leading q x q block, then Rinv = (L^-1)^T as float32 (64 x 64, zero padded), so that
// P @ Rinv has orthonormal columns when Gm = P^T P.  status[0] |= 1 on a non-positive pivot.
//
// Right-looking (outer-product) Cholesky fused with the column-oriented forward substitution L X = I, ONE barrier per
// column.  At step j the Schur complement A holds the pivot column j; with rs = 1/sqrt(A[j][j]) the column of L is
// A[:, j] rs and row j of L^-1 is Xraw[j][:] rs.  Both updates of step j,
//     A[i][k] -= (A[i][j] rs) (A[k][j] rs)        j < k <= i
//     X[i][c] -= (A[i][j] rs) (Xraw[j][c] rs)     c <= j < i
// read only column j of A and row j of X, which step j does not write (they keep their unscaled values for good: the
// scaling by rs is applied once, when Rinv is written), and write only entries no other thread of the step reads.
// 1024 threads = 64 rows x 16 lanes, four columns each.  (Round 1's left-looking version needed three barriers per
// column plus a sequential substitution: 62 us per call, 16 calls per selection.)
__global__ void __launch_bounds__(1024)
chol_inv_kernel(const double *__restrict__ Gm, int q, float *__restrict__ Rinv, int32_t *status) {
    extern __shared__ double dyn_smem[];
    double (*A)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem);
    double (*X)[kN + 1] = reinterpret_cast<double (*)[kN + 1]>(dyn_smem + kN * (kN + 1));
    __shared__ double rdiag[kN];                                   // 1 / L[j][j]
    const int i = threadIdx.x >> 4, l = threadIdx.x & 15;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int c = l + 16 * m;
        A[i][c] = (i < q && c < q) ? Gm[i * kN + c] : 0.0;
        X[i][c] = i == c ? 1.0 : 0.0;
    }
    __syncthreads();
    // 1 / sqrt(pivot) is formed by ONE thread - the owner of the next pivot, right after it has updated it - while the
    // others finish their updates of the step (a float64 rsqrt in all 1024 threads made the FP64 pipe the bottleneck)
    auto pivot_rsqrt = [&](double piv, int j) {
        if (!(piv > 0.0)) { atomicOr(status, 1); piv = 1.0; }
        rdiag[j] = fast_rsqrt(piv);
    };
    if (threadIdx.x == 0) pivot_rsqrt(A[0][0], 0);
    __syncthreads();
    for (int j = 0; j < q; ++j) {
        const double rs = rdiag[j];
        if (i > j && i < q) {
            const double w = -(A[i][j] * (rs * rs));
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int c = l + 16 * m;
                if (c <= j) {
                    X[i][c] = fma(w, X[j][c], X[i][c]);
                } else if (c <= i) {
                    const double v = fma(w, A[c][j], A[i][c]);
                    A[i][c] = v;
                    if (i == j + 1 && c == i) pivot_rsqrt(v, i);     // the next pivot: its owner forms 1 / sqrt right away
                }
            }
        }
        __syncthreads();
    }
    // Rinv[a][b] = (L^-1)[b][a] = Xraw[b][a] / L[b][b]   (a <= b)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int c = l + 16 * m;
        Rinv[i * kN + c] = (i < q && c < q && i <= c) ? (float)(X[c][i] * rdiag[c]) : 0.f;
    }
}

// Symmetric eigendecomposition of the leading q x q block of Gm by cyclic two-sided Jacobi with the round-robin
// parallel ordering.  Outputs eigenvalues in descending order (evals[64], zero padded) and the matching
// eigenvectors as columns of U (float32 and float64, 64 x 64 row-major, zero padded).
//
// A round has two barriers.  (1) One thread per pair (a, b) forms its rotation.  The angle is formed in float32
// (t = sgn(tau) / (|tau| + sqrt(1 + tau^2)) from the float32 casts of the three entries) and then c = rsqrt(1 + t^2),
// s = t c in float64: the rotation is orthogonal to float64 accuracy, the annihilated entry is left at ~1e-7 of its
// size instead of 0, which the next sweep removes - the sweep count is the same, and the float64 divide / sqrt chains
// (~5 x 150 cycles, the serial part of every round) are gone.  (2) A <- J^T A J is done 2 x 2 block-wise: the pairs
// partition the indices, so block (pair i, pair j) is read and written by one thread only, with both rotations
// applied in registers; V <- V J runs in the same phase.
constexpr int kEigThreads = 1024;
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
    // fixed work assignment: one 2 x 2 block of A per thread, row vr of V for pairs vp0 and vp0 + 16
    const int bi = tid / n_pairs, bj = tid % n_pairs;
    const bool has_block = bi < n_pairs;
    const int vr = tid & (kN - 1), vp0 = tid >> 6;
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
                int a = round + tid, b = round + m - 1 - tid;           // both < 2 (m - 1): one conditional subtraction
                if (a >= m - 1) a -= m - 1;
                if (b >= m - 1) b -= m - 1;
                if (tid == 0) a = m - 1;
                if (a > b) { const int t = a; a = b; b = t; }
                double c = 1.0, s = 0.0;
                if (b < q) {
                    const float apq = (float)A[a][b], diff = (float)(A[b][b] - A[a][a]);
                    if (apq != 0.f) {
                        const float tau = __fdividef(diff, 2.f * apq);
                        const float tf = __fdividef(copysignf(1.f, tau), fabsf(tau) + __fsqrt_rn(fmaf(tau, tau, 1.f)));
                        if (tf == tf) {                                  // (0 / 0 cannot occur; guard anyway)
                            const double t = (double)tf;
                            c = fast_rsqrt(fma(t, t, 1.0));
                            s = t * c;
                        }
                    }
                }
                pp[tid] = a; qq[tid] = b; cs[tid][0] = c; cs[tid][1] = s;
            }
            __syncthreads();
            if (has_block) {
                const int ai = pp[bi], bi_ = qq[bi], aj = pp[bj], bj_ = qq[bj];
                const double ci = cs[bi][0], si = cs[bi][1], cj = cs[bj][0], sj = cs[bj][1];
                const double x00 = A[ai][aj], x01 = A[ai][bj_], x10 = A[bi_][aj], x11 = A[bi_][bj_];
                // columns (pair j), then rows (pair i)
                const double y00 = cj * x00 - sj * x01, y01 = sj * x00 + cj * x01;
                const double y10 = cj * x10 - sj * x11, y11 = sj * x10 + cj * x11;
                A[ai][aj] = ci * y00 - si * y10;  A[ai][bj_] = ci * y01 - si * y11;
                A[bi_][aj] = si * y00 + ci * y10; A[bi_][bj_] = si * y01 + ci * y11;
            }
            if (vr < q) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pr = vp0 + 16 * h;
                    if (pr < n_pairs) {
                        const int a = pp[pr], b = qq[pr];
                        const double c = cs[pr][0], s = cs[pr][1];
                        const double vx = V[vr][a], vy = V[vr][b];
                        V[vr][a] = c * vx - s * vy; V[vr][b] = s * vx + c * vy;
                    }
                }
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
    static thread_local bool configured[kMaxDevices] = {};
    const int dev_id = current_device();
    if (!configured[dev_id]) {
        GC_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev_id] = true;
    }
    chol_inv_kernel<<<1, 1024, smem, as_stream(stream)>>>(Gm, q, Rinv, status);
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
